"""Boundary exchange on the GPU (moc_exchange, reference src/comms.c:5-196)."""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import simplemoc_b200 as m
from simplemoc_b200 import api
from oracle_lib import CASES, CommGrid, OracleCase, make_grid as oracle_grid

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_single_domain_every_face_leaks(built):
    """1x1x1: no neighbours.  Every chunk is pairwise-summed into the leakage in (round,
    direction) order and replaced by zeros; the rest of the slab is untouched.  Bit-exact."""
    vals = CASES["exch"]
    host = m.HostProblem(m.derive(m.input_from_values(vals)), seed=6)
    dev = m.DeviceProblem(host, device=0)
    oracle = OracleCase(vals, seed=6)
    assert dev.sweep() == oracle.sweep()
    dev.set(api.ARR_PSI, oracle.psi)
    grids = (CommGrid * 1)(oracle_grid(1, 1, 1, 0))
    hs = (C.c_void_p * 1)(oracle.h)
    assert OracleCase.lib().oracle_exchange(hs, grids, 1) == 0
    before = dev.launch_count
    dev.exchange(m.make_grid(1, 1, 1, 0))
    assert dev.launch_count - before == 3            # chunk sums, ordered accumulation, scatter
    assert np.array_equal(dev.get(api.ARR_PSI), oracle.psi)
    assert dev.leakage == oracle.leakage[0] and dev.leakage != 0
    # k-eff now sees the leakage (solver.c:1425)
    dev.set(api.ARR_FINE_FLUX, oracle.fine_flux)
    assert dev.compute_keff() == oracle.compute_keff()
    dev.close(); host.close(); oracle.close()


def test_overlapped_sweep_exchange_equals_two_calls(built):
    """moc_sweep_exchange starts the exchange once the boundary stacks are swept and runs it under
    the interior stacks on a second stream: same slab, same leakage, same ray state as
    moc_sweep followed by moc_exchange -- over two iterations."""
    vals = CASES["exch"]
    host = m.HostProblem(m.derive(m.input_from_values(vals)), seed=8)
    a, b = m.DeviceProblem(host, device=0), m.DeviceProblem(host, device=0)
    grid = m.make_grid(1, 1, 1, 0)
    for it in range(2):
        na = a.sweep()
        a.exchange(grid)
        nb = b.sweep_exchange(grid)
        assert na == nb
        assert b.timing().n_batches >= 2              # boundary stacks, then interior stacks
        assert np.array_equal(a.get(api.ARR_PSI), b.get(api.ARR_PSI)), it
        assert np.array_equal(a.get(api.ARR_Z_HEIGHT), b.get(api.ARR_Z_HEIGHT))
        assert a.leakage == b.leakage and a.leakage != 0
        # the scalar flux is accumulated with floating-point atomics (order varies run to run):
        # continue both from the same tallies so the next iteration is comparable bit for bit
        b.set(api.ARR_FINE_FLUX, a.get(api.ARR_FINE_FLUX))
        a.renormalize(); b.renormalize()
    a.close(); b.close(); host.close()


def test_dropin_loop_overlaps_the_exchange_when_told_the_grid(built):
    """The reference's loop (main.c:57-92) under the drop-in names, resident: with
    moc_dropin_set_grid the exchange runs inside transport_sweep and the following
    fast_transfer_boundary_fluxes only collects it -- same slab and leakage as the handle API."""
    L = api.lib()
    vals = CASES["exch"]
    host = m.HostProblem(m.derive(m.input_from_values(vals)), seed=8)
    ref = m.DeviceProblem(host, device=0)
    grid = m.make_grid(1, 1, 1, 0)
    n_ref = ref.sweep()
    ref.exchange(grid)
    L.moc_dropin_configure(host.seed, host.rand_calls, 0, 48)
    L.moc_set_resident(1)
    L.moc_dropin_set_grid(C.byref(grid))
    inp = type(host.I).from_buffer_copy(host.I)
    L.transport_sweep(C.byref(host.P), C.byref(inp))
    mirror = L.moc_handle_of(C.byref(host.P))
    before = L.moc_get_launch_count(mirror)
    L.fast_transfer_boundary_fluxes(host.P, inp, grid)              # nothing left to do
    assert L.moc_get_launch_count(mirror) == before
    assert inp.segments_processed == n_ref
    assert L.moc_sync_to_host(C.byref(host.P)) == 0
    assert np.array_equal(host.get(api.ARR_PSI).ravel(), ref.get(api.ARR_PSI).ravel())
    assert host.P.leakage[0] == ref.leakage and ref.leakage != 0
    L.moc_dropin_set_grid(None)
    L.moc_set_resident(0)
    assert L.moc_release(C.byref(host.P)) == 0
    ref.close(); host.close()


def test_exchange_needs_a_communicator(built):
    host = m.HostProblem(m.derive(m.input_from_values(CASES["exch"])), seed=6)
    dev = m.DeviceProblem(host, device=0)
    with pytest.raises(m.MocError, match="moc_comm_init"):
        dev.exchange(m.make_grid(2, 1, 1, 0))
    dev.close(); host.close()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("dims", ["2,1,1", "2,2,1", "2,2,2"])
def test_nccl_exchange_between_domains(built, dims):
    world = eval(dims.replace(",", "*"))
    if api.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "_nccl_exchange_worker.py"), dims]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count(", ok") == world
