"""GPU parity on the SHAPES of the BASELINE configurations (BASELINE.json `configs`).

The other parity files use toy problems; the kernels the bench runs are chosen by the real sizes: the z-stack
height picks the ray-trace kernel (Z = 80: one warp per stack, three rays per lane; Z = 2000 / 800: 16 / 7 warps per
stack), the group count picks the attenuation instantiation (G = 104 / 100 / 128 / 64 / 32 are compiled with G as a
constant), and a source slab larger than the L2 (config 5) switches the attenuation to the per-segment fit.  Here
every configuration keeps ALL of its parameters -- source regions, axial intervals, stack height, groups, polar
angles, segments per track -- and only the number of 2D tracks is cut (`limit_tracks_2D`, the same cut on both
sides) so that the serial oracle finishes in seconds.  Checked: segment total, per-track counts, the digest of
(serial segment index, tally row) pairs and the ray heights bit-exact over two sweeps; flux, angular flux and
sources within 1e-4 (rel-L2 and >= 99.9 % of elements); k-eff within 1e-4.

Reference: init.c:33-103 (the two built-in input sets), default.in, solver.c:283-552.
"""
import numpy as np
import pytest

import simplemoc_b200 as m
from simplemoc_b200 import api
from oracle_lib import INPUT_FILE_FIELDS, OracleCase
from test_gpu_parity import NOISE_CAP, TOL, check_state

pytestmark = pytest.mark.gpu

# the 18 input-file values (io.c:210-267 order) of each configuration
DEFAULT = [17, 17, 27, 5, 2, 0.05, 0.25, 64, 10, 104, 1, 20, 120, 21.42, 400.0, 0.01, 5000, 0]   # init.c:33-74
SMALL = [15, 15, 5, 3, 2, 0.5, 0.2, 5, 5, 104, 0, 1, 120, 1.26 * 17, 400.0, 0.01, 3000, 0]         # init.c:77-103
DEFAULT_IN = [17, 17, 9, 5, 2, 0.05, 0.25, 64, 10, 100, 1, 20, 20, 21.42, 400.0, 0.01, 5000, 0]    # default.in:1-18


def with_(values, **kw):
    v = list(values)
    for k, x in kw.items():
        v[INPUT_FILE_FIELDS.index(k)] = x
    return v


# name -> (values, 2D tracks kept, expected z_stacked, expected source regions, per-segment fit expected)
SHAPES = {
    "default": (DEFAULT, 4, 80, 6750, False),
    "small": (SMALL, 2, 2000, 15000, False),
    "default_in": (DEFAULT_IN, 8, 80, 2250, False),
    "default_g32": (with_(DEFAULT, n_egroups=32), 4, 80, 6750, False),
    "default_g64": (with_(DEFAULT, n_egroups=64), 4, 80, 6750, False),
    "default_g128": (with_(DEFAULT, n_egroups=128), 4, 80, 6750, False),
    # config 5: decomp_assemblies_ax = 2 -> Z = 800, N = 67 500, source slab 380 MB > L2
    "config5": (with_(DEFAULT, decomp_assemblies_ax=2), 2, 800, 67500, True),
}


def test_built_in_input_sets_are_the_reference_ones(built):
    """the value lists above are what moc_set_default_input / moc_set_small_input return (init.c:33-103)"""
    d = m.default_input()
    assert [getattr(d, f) for f in INPUT_FILE_FIELDS][:15] == pytest.approx(DEFAULT[:15])
    assert d.n_2D_source_regions_per_assembly == DEFAULT[16] and d.precision == pytest.approx(DEFAULT[15])
    s = m.small_input()
    assert [getattr(s, f) for f in INPUT_FILE_FIELDS][:15] == pytest.approx(SMALL[:15])
    assert s.n_2D_source_regions_per_assembly == SMALL[16]


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_baseline_shape_slice(built, name):
    values, keep, Z, N, per_segment_fit = SHAPES[name]
    seed = 1
    inp = m.derive(m.input_from_values(values), limit_tracks_2D=keep)
    assert (inp.z_stacked, inp.n_source_regions_per_node, inp.ntracks_2D) == (Z, N, keep)
    host = m.HostProblem(inp, seed=seed)
    dev = m.DeviceProblem(host, device=0)
    dev.set_option(api.OPT_DIGEST, 1)
    oracle = OracleCase(values, seed=seed, limit_tracks_2D=keep)
    assert oracle.I.ntracks == inp.ntracks and oracle.init_rand_calls == host.rand_calls
    # the paths the full-size run takes
    assert dev.get_option(api.OPT_FIT_PER_SEGMENT) == (1 if per_segment_fit else 0)
    assert dev.get_option(api.OPT_WALK_KERNEL) == 0 and dev.get_option(api.OPT_LANES) == 0

    n_gpu, n_cpu = dev.sweep(), oracle.sweep()
    assert n_gpu == n_cpu
    assert np.array_equal(dev.get(api.ARR_SEG_COUNT), oracle.seg_count)
    assert np.array_equal(dev.get(api.ARR_QSR_DIGEST), oracle.digest)
    assert np.array_equal(dev.get(api.ARR_Z_HEIGHT), oracle.z_height)
    check_state(dev, oracle, f"{name} sweep", noise_cap=NOISE_CAP)
    dev.renormalize(); oracle.renormalize()
    check_state(dev, oracle, f"{name} renormalize")
    r_gpu, r_cpu = dev.update_sources(1.0), oracle.update_sources(1.0)
    assert abs(r_gpu - r_cpu) <= 1e-3 * abs(r_cpu)
    k_gpu, k_cpu = dev.compute_keff(), oracle.compute_keff()
    assert abs(k_gpu - k_cpu) <= TOL * abs(k_cpu), (k_gpu, k_cpu)
    # second sweep: stale ray heights, moved random stream, iterated sources -- the integers stay exact
    assert dev.sweep() == oracle.sweep()
    assert np.array_equal(dev.get(api.ARR_SEG_COUNT), oracle.seg_count)
    assert np.array_equal(dev.get(api.ARR_QSR_DIGEST), oracle.digest)
    assert np.array_equal(dev.get(api.ARR_Z_HEIGHT), oracle.z_height)
    print(f"{name}: {n_gpu} segments x {inp.n_egroups} groups, k-eff {k_gpu:.7f} (oracle {k_cpu:.7f})")
    dev.close(); host.close(); oracle.close()


@pytest.mark.parametrize("name", ["default", "config5"])
def test_baseline_shape_slice_sfu_mode_integers(built, name):
    """the SFU-exponential mode (the north star's performance mode) shares the ray trace: same integers"""
    values, keep, Z, N, _ = SHAPES[name]
    inp = m.derive(m.input_from_values(values), limit_tracks_2D=keep)
    host = m.HostProblem(inp, seed=2)
    dev = m.DeviceProblem(host, device=0, exp_mode=api.EXP_SFU)
    dev.set_option(api.OPT_DIGEST, 1)
    oracle = OracleCase(values, seed=2, exp_mode=1, limit_tracks_2D=keep)
    assert dev.sweep() == oracle.sweep()
    assert np.array_equal(dev.get(api.ARR_SEG_COUNT), oracle.seg_count)
    assert np.array_equal(dev.get(api.ARR_QSR_DIGEST), oracle.digest)
    dev.close(); host.close(); oracle.close()

