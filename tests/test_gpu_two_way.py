"""GPU parity of moc_two_way_sweep against the oracle's restatement of two_way_transport_sweep (solver.c:556-891,
SURVEY 8f row f3).  The oracle itself is pinned to the reference's own function in tests/test_oracle_vs_ref.py.

Integers are bit-exact on every case: the count the function returns (both passes), per-track segment counts, the
digest of (serial segment index, tally row) pairs of the forward pass, the digest of (track, segment, tally row) of
the backward pass, the random-stream position (seen through the NEXT sweep's source regions) and the ray heights.
Floating point: within 1e-4 where the reference's function is defined (no step of negative length: "polar1",
"zone").  Elsewhere the reference reads its heap; the oracle and the GPU both answer those lookups from the first table
cell, and the comparison is norm-wise only (negative optical lengths make exp(-tau) grow along the ray).
"""
import numpy as np
import pytest

from simplemoc_b200 import api
from oracle_lib import rel_l2
from test_gpu_parity import NOISE_CAP, TOL, check_state, make_pair

pytestmark = pytest.mark.gpu


def integers_match(dev, oracle, what):
    assert np.array_equal(dev.get(api.ARR_SEG_COUNT), oracle.seg_count), what
    assert np.array_equal(dev.get(api.ARR_QSR_DIGEST), oracle.digest), what
    assert np.array_equal(dev.get(api.ARR_QSR_DIGEST_BACK), oracle.digest_back), what
    assert np.array_equal(dev.get(api.ARR_Z_HEIGHT), oracle.z_height), what       # every ray back at its start height


@pytest.mark.parametrize("case,seed,defined", [("polar1", 2, True), ("zone", 2, True), ("tiny", 3, False),
                                               ("mini104", 5, False), ("tiny_flat", 2, False), ("odd", 4, False),
                                               ("g130", 1, False)])
def test_two_way_sweep(built, case, seed, defined):
    host, dev, oracle = make_pair(case, seed=seed)
    for sweep in range(2):
        n_gpu, n_cpu = dev.two_way_sweep(), oracle.two_way_sweep()
        assert n_gpu == n_cpu
        assert (oracle.table_oob == 0) == defined
        integers_match(dev, oracle, f"{case} two-way sweep {sweep}")
        psi = dev.get(api.ARR_PSI)
        assert psi[:, 1].any()                                                    # the backward rows are in use
        if defined and sweep == 0:
            check_state(dev, oracle, f"{case} two-way sweep", noise_cap=NOISE_CAP)
        elif defined:
            assert rel_l2(dev.get(api.ARR_FINE_FLUX), oracle.fine_flux) <= TOL
            assert rel_l2(psi, oracle.psi) <= TOL
        else:
            ok = np.isfinite(oracle.fine_flux).all() and np.isfinite(oracle.psi).all()
            if ok:
                assert rel_l2(dev.get(api.ARR_FINE_FLUX), oracle.fine_flux) <= 1e-3
                assert rel_l2(psi, oracle.psi) <= 1e-3
        if defined:
            dev.renormalize(); oracle.renormalize()
            dev.update_sources(0.95); oracle.update_sources(0.95)
    # the one-way sweep continues from the state the two-way sweep left: same random stream, same heights
    assert dev.sweep() == oracle.sweep()
    assert np.array_equal(dev.get(api.ARR_SEG_COUNT), oracle.seg_count)
    assert np.array_equal(dev.get(api.ARR_QSR_DIGEST), oracle.digest)
    dev.close(); host.close(); oracle.close()


def test_two_way_sweep_in_batches(built):
    """a record buffer that holds a few z-stacks at a time: same results"""
    host, dev, oracle = make_pair("polar1", seed=2, batch=300)
    assert dev.two_way_sweep() == oracle.two_way_sweep()
    assert dev.timing().n_batches > 1
    integers_match(dev, oracle, "polar1 batches")
    check_state(dev, oracle, "polar1 two-way sweep in batches", noise_cap=NOISE_CAP)
    dev.close(); host.close(); oracle.close()


def test_two_way_sweep_sfu_mode(built):
    """The SFU exponential shares the integers.  Floating point as in test_sfu_mode_against_exact_exp_oracle: against
    the reference's formula with libm's expf the scalar flux is dominated by a handful of cancelling elements, so the
    bar is the median element and the (non-cancelling) angular flux."""
    host, dev, oracle = make_pair("polar1", seed=6, exp_mode=1)
    assert dev.two_way_sweep() == oracle.two_way_sweep()
    integers_match(dev, oracle, "polar1 sfu")
    flux, ref = dev.get(api.ARR_FINE_FLUX).astype(np.float64), np.asarray(oracle.fine_flux, np.float64)
    rel = np.abs(flux - ref).ravel() / np.maximum(np.abs(ref).ravel(), 1e-300)
    p = rel_l2(dev.get(api.ARR_PSI), oracle.psi)
    print(f"two-way, SFU: scalar flux median rel. difference {np.median(rel):.2e}, 90 % {np.quantile(rel, 0.9):.2e}, "
          f"rel-L2 {rel_l2(flux, ref):.2e}; angular flux rel-L2 {p:.2e}")
    # measured on B200: median 1.4e-7, 90 % 8.6e-7, rel-L2 7.4e-3 (cancelling elements), angular flux 4.9e-6
    assert np.quantile(rel, 0.9) <= TOL
    assert p <= TOL
    dev.close(); host.close(); oracle.close()


def test_two_way_dropin_name_against_the_reference_own_function(built):
    """two_way_transport_sweep under the reference's name, on the Params/Input the UNMODIFIED reference's build_tracks()
    allocated, against the UNMODIFIED reference's own two_way_transport_sweep on a twin problem (oracle/_ref) -- on a
    case where that function is defined (no table lookup in front of the table)."""
    import ctypes as C
    from oracle_lib import CASES, RefCase, frac_within, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref did not travel")
    L = api.lib()
    vals, seed = CASES["polar1"], 2
    mine, theirs = RefCase(vals, seed=seed), RefCase(vals, seed=seed)
    L.moc_dropin_configure(seed, mine.init_rand_calls, 0, 48)
    L.moc_set_resident(0)
    params = api.Params.from_address(mine._ptr("params"))
    inp = api.Input.from_address(mine._ptr("input_mut"))
    for sweep in range(2):
        n_cpu = theirs.two_way_sweep()                                        # the reference, on the CPU
        L.two_way_transport_sweep(C.byref(params), C.byref(inp))              # the library, on the GPU
        assert inp.segments_processed == n_cpu
        assert np.array_equal(mine.z_height, theirs.z_height)
        for name in ("fine_flux", "psi"):
            a, b = getattr(mine, name), getattr(theirs, name)
            assert rel_l2(a, b) <= TOL, (name, sweep)
            assert sweep or frac_within(a, b, TOL) >= 0.999, name
    # the one-way sweep after it: the random stream moved by one draw per forward segment on both sides
    L.transport_sweep(C.byref(params), C.byref(inp))
    assert inp.segments_processed == theirs.sweep()
    assert np.array_equal(mine.z_height, theirs.z_height)
    assert L.moc_release(C.byref(params)) == 0
