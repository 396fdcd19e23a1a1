"""Pins the boundary-exchange oracle (oracle_exchange, oracle_make_grid, the all-rank reductions of
oracle/moc_oracle.c) against the reference's OWN communication code.

The image has no MPI, so oracle/Makefile compiles the reference's sources UNMODIFIED with -DMPI
against oracle/mpi_stub/mpi.h -- an in-process MPI: one pthread per rank, eager mailbox, MPICH's
MPI_PROC_NULL = -1 -- into oracle/_ref/libsimplemoc_ref_mpi.so.  What executes here is therefore the reference's
comms.c:5-196 (fast_transfer_boundary_fluxes), init.c:162-225 (init_mpi_grid, its hard-coded 2x2x1 grid) and the
MPI branches of solver.c:1186-1196 / 1391-1425, on four (or 1, 2, 8) ranks with different domains.  The oracle's
restatement must reproduce every rank's flux slab and leakage BIT FOR BIT; the chain then continues
oracle_exchange == the product's schedule over gloo (tests/test_exchange_gloo.py, CPU) == moc_exchange over NCCL
(tests/test_gpu_exchange.py, bench.py's exchange_parity).
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_lib import (CASES, CommGrid, OracleCase, RefCase, REF_DIR, ensure_ref_built, make_grid as oracle_grid,
                        write_input_file)

GRID_FIELDS = [n for n, _ in CommGrid._fields_]
HAVE = os.path.exists(os.path.join(REF_DIR, "libsimplemoc_ref_mpi.so")) or os.path.isdir("/root/reference/src")
pytestmark = pytest.mark.skipif(not HAVE, reason="oracle/_ref/libsimplemoc_ref_mpi.so not built and /root/reference absent")


def mpi_lib():
    assert ensure_ref_built()
    L = RefCase.lib("_mpi")
    L.ref_mpi_case_create.restype = C.c_void_p
    L.ref_mpi_case_create.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int]
    L.ref_mpi_run.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_float)]
    L.ref_mpi_get_grid.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.ref_mpi_set_grid.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.ref_mpi_select_stream.argtypes = [C.c_uint64, C.c_uint64]
    return L


class MpiWorld:
    """nranks domains of the -DMPI reference build, and the same domains in the oracle"""

    def __init__(self, values, nranks, seed0, tmp_path):
        self.L = mpi_lib()
        self.n = nranks
        path = write_input_file(str(tmp_path / "case.in"), values).encode()
        self.ref, self.ora = [], []
        for r in range(nranks):
            h = self.L.ref_mpi_case_create(path, seed0 + r, r, nranks)
            c = RefCase.__new__(RefCase)
            c.variant, c.h = "_mpi", h
            assert c.I.mype == r                      # init.c:7-11: the rank comes from MPI_Comm_rank
            self.ref.append(c)
            self.ora.append(OracleCase(values, seed=seed0 + r))
        self.seed0 = seed0

    def sweep_all(self):
        for r, (c, o) in enumerate(zip(self.ref, self.ora)):
            self.L.ref_mpi_select_stream(self.seed0 + r, c.init_rand_calls)   # rand() is process-global in the reference
            assert c.sweep() == o.sweep()
            assert np.array_equal(c.psi, o.psi)

    def run(self, what):
        hs = (C.c_void_p * self.n)(*[c.h for c in self.ref])
        keff = (C.c_float * self.n)()
        assert self.L.ref_mpi_run(self.n, hs, what, keff) == 0
        return list(keff)

    def ref_grid(self, r):
        g = (C.c_int * 12)()
        self.L.ref_mpi_get_grid(self.ref[r].h, g)
        return list(g)

    def set_grids(self, grids):
        for c, g in zip(self.ref, grids):
            self.L.ref_mpi_set_grid(c.h, (C.c_int * 12)(*[getattr(g, f) for f in GRID_FIELDS]))

    def close(self):
        for c in self.ref:
            c.close()
        for o in self.ora:
            o.close()


def check_exchange(world, grids):
    """the reference's fast_transfer_boundary_fluxes on every rank against oracle_exchange"""
    before = [o.psi.copy() for o in world.ora]
    world.run(2)
    arr = (CommGrid * world.n)(*grids)
    hs = (C.c_void_p * world.n)(*[o.h for o in world.ora])
    assert OracleCase.lib().oracle_exchange(hs, arr, world.n) == 0
    moved = 0
    for r in range(world.n):
        assert np.array_equal(world.ref[r].psi, world.ora[r].psi), f"rank {r}: slab differs from comms.c"
        assert world.ref[r].leakage[0] == world.ora[r].leakage[0], f"rank {r}: leakage differs from comms.c"
        assert world.ora[r].leakage[0] != 0        # every rank of these grids has at least one border face
        moved += int((world.ora[r].psi != before[r]).sum())
    assert moved > 0
    return moved


def test_reference_grid_and_exchange_on_its_own_2x2x1_grid(tmp_path):
    """init_mpi_grid (init.c:162-225: MPI_Cart_create {2,2,1} + six MPI_Cart_shift) and the exchange on it"""
    w = MpiWorld(CASES["exch"], 4, 21, tmp_path)
    w.sweep_all()
    w.run(1)                                              # the reference builds its own neighbour tables
    grids = [oracle_grid(2, 2, 1, r) for r in range(4)]
    for r in range(4):
        assert w.ref_grid(r) == [getattr(grids[r], f) for f in GRID_FIELDS], r
    # rank 0 of a 2x2x1 grid: neighbours 2 (x+) and 1 (y+), borders elsewhere
    assert (grids[0].x_pos_dest, grids[0].y_pos_dest, grids[0].x_neg_dest, grids[0].z_pos_dest) == (2, 1, -1, -1)
    check_exchange(w, grids)
    # a second exchange on the exchanged slabs: the leakage keeps accumulating (comms.c:120, never reset)
    check_exchange(w, grids)
    # the MPI reductions: renormalize_flux (MPI_Allreduce, solver.c:1190-1195), compute_keff (3 x MPI_Reduce)
    w.run(4)
    hs = (C.c_void_p * 4)(*[o.h for o in w.ora])
    OracleCase.lib().oracle_renormalize_all(hs, 4)
    for r in range(4):
        assert np.array_equal(w.ref[r].fine_flux, w.ora[r].fine_flux, equal_nan=True)
        assert np.array_equal(w.ref[r].psi, w.ora[r].psi, equal_nan=True)
    k_ref = w.run(8)[0]                                   # rank 0 holds the result (solver.c:1394-1425)
    OracleCase.lib().oracle_compute_keff_all.restype = C.c_float
    k_ora = OracleCase.lib().oracle_compute_keff_all(hs, 4)
    assert k_ref == k_ora and np.isfinite(k_ref)
    w.close()


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (1, 1, 2), (2, 2, 2)])
def test_exchange_on_other_grids(tmp_path, dims):
    """comms.c itself does not care where its twelve neighbour ranks come from: 1x1x1 (every face leaks), and the
    grids the product adds (moc_make_grid generalises init.c's {2,2,1}), tables written into CommGrid by hand"""
    n = dims[0] * dims[1] * dims[2]
    w = MpiWorld(CASES["exch"], n, 40, tmp_path)
    w.sweep_all()
    grids = [oracle_grid(*dims, r) for r in range(n)]
    w.set_grids(grids)
    check_exchange(w, grids)
    w.close()


def test_product_grid_equals_reference_grid(built):
    """moc_make_grid(2, 2, 1, r) -- the product's init_mpi_grid -- against the tables the reference built above"""
    import simplemoc_b200 as m
    for dims in ((2, 2, 1), (2, 2, 2), (1, 1, 1), (3, 2, 2)):
        for r in range(dims[0] * dims[1] * dims[2]):
            a, b = m.make_grid(*dims, r), oracle_grid(*dims, r)
            assert [getattr(a, f) for f in GRID_FIELDS] == [getattr(b, f) for f in GRID_FIELDS]
