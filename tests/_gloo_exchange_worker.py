"""Worker of tests/test_exchange_gloo.py: one process per spatial domain, gloo backend (CPU).

Executes the product's boundary-exchange schedule (moc_exchange_plan + moc_make_grid, the host
logic moc_exchange drives NCCL with) over torch.distributed point-to-point operations, in the same
order and with the same staging as simplemoc_b200/csrc/moc_device.cu::exchange_on_stream, and
checks the result on every rank against the all-ranks CPU model of comms.c (oracle_exchange)."""
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
from oracle_lib import CASES, CommGrid, OracleCase, make_grid as oracle_grid  # noqa: E402


def main():
    cx, cy, cz = (int(v) for v in sys.argv[1].split(","))
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert world == cx * cy * cz
    inp = m.derive(m.input_from_values(CASES["exch"]))
    T3, G = inp.ntracks, inp.n_egroups
    rng = np.random.default_rng(1000 + rank)
    slab = rng.standard_normal(2 * T3 * G).astype(np.float32)       # this rank's [T3][2][G] flux slab
    mine = torch.from_numpy(slab.copy())

    # ---- the product's schedule, executed with gloo the way moc_exchange executes it with NCCL
    grid = m.make_grid(cx, cy, cz, rank)
    L = api.lib()
    n = L.moc_exchange_plan(C.byref(inp), C.byref(grid), None, 0)
    ops = (api.ExchangeOp * max(n, 1))()
    assert L.moc_exchange_plan(C.byref(inp), C.byref(grid), ops, n) == n and n > 0
    leakage = np.float32(0)
    reqs, stage = [], {}
    for k in range(n):
        op = ops[k]
        chunk = mine[op.offset: op.offset + op.count]
        if op.send_to >= 0:
            reqs.append(dist.isend(chunk.clone(), op.send_to))      # per-peer FIFO order pairs them up
        if op.recv_from >= 0:
            stage[k] = torch.empty(op.count, dtype=torch.float32)
            reqs.append(dist.irecv(stage[k], op.recv_from))
    for r in reqs:
        r.wait()
    for k in range(n):
        op = ops[k]
        if op.recv_from >= 0:
            mine[op.offset: op.offset + op.count] = stage[k]
        else:
            mine[op.offset: op.offset + op.count] = 0

    # ---- the model: every rank's original slab gathered, comms.c restated over all of them
    everyone = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(everyone, torch.from_numpy(slab))
    cases = [OracleCase(CASES["exch"], seed=1) for _ in range(world)]
    for c, s in zip(cases, everyone):
        c.psi.ravel()[:] = s.numpy()
    grids = (CommGrid * world)(*[oracle_grid(cx, cy, cz, r) for r in range(world)])
    hs = (C.c_void_p * world)(*[c.h for c in cases])
    assert OracleCase.lib().oracle_exchange(hs, grids, world) == 0
    want = cases[rank].psi.ravel()
    assert np.array_equal(mine.numpy(), want), f"rank {rank}: exchanged slab differs from the comms.c model"
    moved = int((want != slab).sum())
    assert moved > 0
    print(f"rank {rank}/{world}: {n} chunks, {moved} floats replaced, ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
