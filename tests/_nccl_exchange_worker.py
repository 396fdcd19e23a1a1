"""Worker of tests/test_gpu_exchange.py: one process per GPU, NCCL.  Each rank sweeps its own
domain on its GPU, the library exchanges boundary fluxes with ncclSend/ncclRecv (moc_exchange),
and every rank checks its slab and leakage against the all-ranks CPU model of comms.c."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import simplemoc_b200 as m  # noqa: E402
from simplemoc_b200 import api  # noqa: E402
from oracle_lib import CASES, CommGrid, OracleCase, make_grid as oracle_grid, rel_l2  # noqa: E402


def main():
    cx, cy, cz = (int(v) for v in sys.argv[1].split(","))
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    vals = CASES["exch"]
    host = m.HostProblem(m.derive(m.input_from_values(vals)), seed=21 + rank)
    dev = m.DeviceProblem(host, device=local)
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(api.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    dev.comm_init(world, rank, bytes(buf.cpu().numpy().tobytes()))

    # the model: all domains on the CPU
    cases = [OracleCase(vals, seed=21 + r) for r in range(world)]
    for c in cases:
        c.sweep()
    assert dev.sweep() == cases[rank].I.segments_processed
    # start the exchange from identical slabs so that the comparison below is bit-exact
    dev.set(api.ARR_PSI, cases[rank].psi)
    grids = (CommGrid * world)(*[oracle_grid(cx, cy, cz, r) for r in range(world)])
    hs = (C.c_void_p * world)(*[c.h for c in cases])
    assert OracleCase.lib().oracle_exchange(hs, grids, world) == 0
    dev.exchange(m.make_grid(cx, cy, cz, rank))
    got = dev.get(api.ARR_PSI)
    assert np.array_equal(got, cases[rank].psi), f"rank {rank}: slab differs after the exchange"
    assert dev.leakage == cases[rank].leakage[0], (dev.leakage, cases[rank].leakage[0])
    # the scalar reductions across ranks (solver.c:1190-1195, 1394-1418): one domain-sum each
    dev.set(api.ARR_FINE_FLUX, cases[rank].fine_flux)
    dev.renormalize()
    dev.update_sources(1.0)
    k = dev.compute_keff()
    # model: sum the per-rank partial sums in double (the reference's MPI reduction order is
    # implementation-defined), compare within FP32 tolerance
    fis = absr = leak = 0.0
    for c in cases:
        G, F = c.I.n_egroups, c.I.fai
        x = c.xs[c.xs_index]                                  # [N][G][3]
        flux = c.fine_flux.astype(np.float64)                 # [N][F][G]
        fis += float((flux * c.vol[:, None, None] * x[:, None, :, 0]).sum())
    norm = 1.0 / fis
    for c in cases:
        x = c.xs[c.xs_index]
        flux = c.fine_flux.astype(np.float64) * (norm * 4 * np.pi * c.I.fai / c.vol[:, None, None].astype(np.float64))
        absr += float((flux * x[:, None, :, 1]).sum())
        leak += float(c.leakage[0])
    fis2 = sum(float((c.fine_flux.astype(np.float64) * (norm * 4 * np.pi * c.I.fai / c.vol[:, None, None]) *
                      c.xs[c.xs_index][:, None, :, 0]).sum()) for c in cases)
    k_model = fis2 / (absr + leak)
    assert abs(k - k_model) <= 1e-4 * abs(k_model), (k, k_model)
    print(f"rank {rank}/{world}: exchange bit-exact, leakage {dev.leakage:.6g}, keff {k:.6f} (model {k_model:.6f}), ok",
          flush=True)
    # the overlapped form (moc_sweep_exchange) against the two separate calls, from the same host data
    def fresh():
        d = m.DeviceProblem(host, device=local)
        b2 = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            b2.copy_(torch.frombuffer(bytearray(api.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(b2, 0)
        d.comm_init(world, rank, bytes(b2.cpu().numpy().tobytes()))
        return d
    one, two = fresh(), fresh()
    grid = m.make_grid(cx, cy, cz, rank)
    n1 = one.sweep()
    one.exchange(grid)
    n2 = two.sweep_exchange(grid)
    assert n1 == n2
    assert np.array_equal(one.get(api.ARR_PSI), two.get(api.ARR_PSI)), f"rank {rank}: overlapped exchange differs"
    assert one.leakage == two.leakage
    one.close(); two.close()
    dist.barrier()
    dev.close(); host.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
