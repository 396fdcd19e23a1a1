"""Host side of the library (moc_host.c): configuration, CLI, derived sizes, synthetic problem
construction in the reference's draw order, the boundary-exchange schedule.  CPU only."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import simplemoc_b200 as m
from simplemoc_b200 import api
from oracle_lib import CASES, OracleCase, write_input_file

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_digests.json")) as f:
    _G = json.load(f)
GOLDEN, GOLDEN_TRACKS = _G["table"], _G["tracks"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_default_and_small_inputs_derive_the_surveyed_sizes(built):
    d = m.derive(m.default_input())                      # SURVEY 8 "D"
    assert (d.ntracks_2D, d.z_stacked, d.ntracks, d.n_source_regions_per_node) == (19386, 80, 15508800, 6750)
    assert (d.n_egroups, d.n_polar_angles, d.fai, d.cai) == (104, 10, 5, 27)
    s = m.derive(m.small_input())                        # SURVEY 8 "S"
    assert (s.ntracks_2D, s.z_stacked, s.ntracks, s.n_source_regions_per_node) == (120, 2000, 1200000, 15000)


def check_build(g, track_file=None):
    host = m.HostProblem(m.derive(m.input_from_values(g["values"], track_file)), seed=g["seed"])
    assert host.rand_calls == g["init_rand_calls"]
    I = host.I
    derived = {"ntracks_2D": I.ntracks_2D, "z_stacked": I.z_stacked, "ntracks": I.ntracks,
               "n_source_regions_per_node": I.n_source_regions_per_node}
    if track_file:
        derived.update({"n_azimuthal": I.n_azimuthal, "radial_ray_sep": float(I.radial_ray_sep),
                        "segments_per_track": I.segments_per_track})
    assert derived == g["derived"]
    got = {"az_weight": sha(host.get(api.HOST_AZ_WEIGHT)), "n_segments": sha(host.get(api.HOST_N_SEGMENTS)),
           "seg_lengths": sha(host.get(api.HOST_SEG_LENGTHS)), "p_weight": sha(host.get(api.ARR_P_WEIGHT)),
           "z_height": sha(host.get(api.ARR_Z_HEIGHT)), "xs": sha(host.get(api.HOST_XS)),
           "scatter": sha(host.get(api.HOST_SCATTER)), "fine_source": sha(host.get(api.ARR_FINE_SOURCE)),
           "sigT": sha(host.get(api.ARR_SIGT)), "xs_index": sha(host.get(api.HOST_XS_INDEX)),
           "vol": sha(host.get(api.HOST_VOL)), "table": sha(host.get(api.HOST_TABLE))}
    assert got == g["init"]
    assert not host.get(api.ARR_PSI).any() and not host.get(api.ARR_FINE_FLUX).any()
    host.close()


@pytest.mark.parametrize("case", sorted(GOLDEN))
def test_build_tracks_reproduces_the_reference_draw_for_draw(built, case):
    """moc_build_tracks == build_tracks (init.c:106-159) of the unmodified reference under the
    same counter random stream: every array, and the number of draws consumed."""
    check_build(GOLDEN[case])


@pytest.mark.parametrize("case", sorted(GOLDEN_TRACKS))
def test_build_tracks_from_a_track_file_matches_the_reference(built, case):
    """`-d <file>`: moc_build_tracks with I->load_tracks == the reference's build_tracks with its own
    load_OpenMOC_tracks (tracks.c:170-323) on tests/golden/tracks_44.bin: updated Input, every array,
    the number of draws."""
    g = GOLDEN_TRACKS[case]
    check_build(g, os.path.join(HERE, "golden", g["track_file"]))


def test_track_file_errors_are_reported_not_undefined(built, tmp_path):
    """the reference reads whatever fread leaves behind; the library returns MOC_EIO"""
    from openmoc_tracks import synthetic_tracks, write_track_file
    vals = CASES["tiny"]
    with pytest.raises(m.MocError, match="cannot open track file"):
        m.HostProblem(m.derive(m.input_from_values(vals, str(tmp_path / "missing.bin"))), seed=1)
    good = write_track_file(str(tmp_path / "good.bin"), [3, 2], synthetic_tracks(1, [3, 2], 6))
    blob = open(good, "rb").read()
    for cut in (2, 10, 40, len(blob) // 2, len(blob) - 3):
        bad = tmp_path / f"cut{cut}.bin"
        bad.write_bytes(blob[:cut])
        with pytest.raises(m.MocError, match="track file"):
            m.HostProblem(m.derive(m.input_from_values(vals, str(bad))), seed=1)
    # a negative segment count / an absurd azimuthal count
    import struct
    hdr = len(struct.pack("=i", 0)) + len(b"test geometry")
    evil = bytearray(blob)
    evil[hdr:hdr + 4] = struct.pack("=i", 2 ** 30)
    (tmp_path / "evil.bin").write_bytes(bytes(evil))
    with pytest.raises(m.MocError, match="azimuthal"):
        m.HostProblem(m.derive(m.input_from_values(vals, str(tmp_path / "evil.bin"))), seed=1)
    # the intact file loads, CMFD flavour included (two more ints per segment, tracks.c:300-304)
    host = m.HostProblem(m.derive(m.input_from_values(vals, good)), seed=1)
    assert host.I.ntracks_2D == 5 and host.I.n_azimuthal == 2
    n_seg = host.get(api.HOST_N_SEGMENTS)
    lengths = host.get(api.HOST_SEG_LENGTHS)
    host.close()
    cm = write_track_file(str(tmp_path / "cmfd.bin"), [3, 2], synthetic_tracks(1, [3, 2], 6), cmfd=True)
    inp = m.derive(m.input_from_values(vals))
    t2 = C.c_void_p()
    total = C.c_long()
    L = api.lib()
    L.moc_load_openmoc_tracks.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64,
                                          C.POINTER(C.c_void_p), C.POINTER(C.c_long)]
    assert L.moc_load_openmoc_tracks(cm.encode(), 1, C.byref(inp), 1, 0, C.byref(t2), C.byref(total)) == 0
    assert total.value == int(n_seg.sum()) == lengths.size and inp.ntracks_2D == 5


def test_cli_dash_d_selects_the_track_file(built):
    L = api.lib()
    argv = (C.c_char_p * 4)(b"SimpleMOC", b"-d", b"some/tracks.bin", b"-s")
    inp = m.default_input()
    assert L.moc_read_CLI(4, argv, C.byref(inp)) == 0
    assert inp.load_tracks and inp.track_file == b"some/tracks.bin"       # src/io.c:160-168


def test_input_file_and_cli_order(built, tmp_path):
    L = api.lib()
    f = write_input_file(str(tmp_path / "case.in"), CASES["odd"])
    inp = m.read_input_file(f)
    assert (inp.cai, inp.fai, inp.n_egroups, inp.segments_per_track) == (3, 4, 10, 8)
    assert abs(inp.radial_ray_sep - 2.5) < 1e-7 and inp.decompose is False
    # options apply in command-line order (io.c:115-181): "-i f -s" ends small, "-s -i f" ends with f
    def cli(*args):
        argv = (C.c_char_p * (len(args) + 1))(b"SimpleMOC", *[a.encode() for a in args])
        inp = m.default_input()
        assert L.moc_read_CLI(len(args) + 1, argv, C.byref(inp)) == 0, L.moc_last_error()
        return inp
    a = cli("-i", f, "-s")
    assert (a.n_egroups, a.cai) == (104, 5)
    b = cli("-s", "-i", f, "-t", "7")
    assert (b.n_egroups, b.cai, b.nthreads) == (10, 3, 7)
    # the shipped default.in carries 20 lines; only the first 18 are read
    with open(f, "a") as fh:
        fh.write("1 - extra\n2 - extra\n")
    assert m.read_input_file(f).papi_event_set == 0
    # short file
    short = tmp_path / "short.in"
    short.write_text("1\n2\n3\n")
    with pytest.raises(m.MocError, match="expected 18 values"):
        m.read_input_file(str(short))


def test_time_per_intersection_is_the_reference_formula(built):
    inp = m.derive(m.default_input())
    inp.segments_processed = 1926588096
    ns = api.lib().moc_time_per_intersection(C.byref(inp), 360.0)      # 1.797 ns
    assert abs(ns - 360.0 / 1926588096 * 1e9 / 104) < 1e-12
    assert abs(ns - 1.797) < 1e-3


def test_est_mem_usage_of_the_default_problem_is_about_13_gb(built):
    inp = m.derive(m.default_input())
    gb = api.lib().moc_est_mem_usage(C.byref(inp)) / 1e9
    assert 12.5 < gb < 14.5        # README.txt:146-147 "around 13 GB"


@pytest.mark.parametrize("dims", [(2, 2, 1), (2, 2, 2), (1, 1, 1), (3, 1, 2)])
def test_make_grid_matches_the_cart_shift_model(built, dims):
    from oracle_lib import make_grid as oracle_grid
    cx, cy, cz = dims
    fields = [n for n, _ in api.CommGrid._fields_]
    for rank in range(cx * cy * cz):
        a, b = m.make_grid(cx, cy, cz, rank), oracle_grid(cx, cy, cz, rank)
        assert [getattr(a, f) for f in fields] == [getattr(b, f) for f in fields]
    # the reference's own grid: dims {2,2,1} (init.c:169); rank 0 sits at (0,0,0)
    g = m.make_grid(2, 2, 1, 0)
    assert (g.x_pos_dest, g.x_pos_src, g.y_pos_dest, g.z_pos_dest, g.z_neg_dest) == (2, -1, 1, -1, -1)


def exchange_plan(inp, grid):
    L = api.lib()
    n = L.moc_exchange_plan(C.byref(inp), C.byref(grid), None, 0)
    ops = (api.ExchangeOp * max(n, 1))()
    assert L.moc_exchange_plan(C.byref(inp), C.byref(grid), ops, n) == n
    return list(ops)[:n]


def test_exchange_plan_follows_comms_c(built):
    inp = m.derive(m.default_input())
    grid = m.make_grid(2, 2, 2, 5)
    ops = exchange_plan(inp, grid)
    nmsg = (C.c_long * 6)()
    from oracle_lib import Input as OracleInput
    OracleCase.lib().oracle_exchange_plan(C.byref(OracleInput.from_buffer_copy(inp)), nmsg)   # comms.c:12-28 restated
    assert len(ops) == sum(nmsg)
    # ~270 messages per axial face, ~252 per radial face on the built-in default
    assert [nmsg[d] for d in range(6)] == [252, 252, 252, 252, 270, 270]
    at = 0
    k = 0
    for i in range(max(nmsg)):
        for d in range(6):
            if i >= nmsg[d]:
                continue
            op = ops[k]
            assert (op.round, op.direction, op.offset, op.count) == (i, d, at, 10000 * 104)
            at += op.count
            k += 1
    dest = [grid.x_pos_dest, grid.x_neg_dest, grid.y_pos_dest, grid.y_neg_dest, grid.z_pos_dest, grid.z_neg_dest]
    src = [grid.x_pos_src, grid.x_neg_src, grid.y_pos_src, grid.y_neg_src, grid.z_pos_src, grid.z_neg_src]
    assert all(op.send_to == dest[op.direction] and op.recv_from == src[op.direction] for op in ops)
    assert at * 4 <= 2 * inp.ntracks * 104 * 4          # inside the flux slab
    # a problem too small for a single message exchanges nothing
    tiny = m.derive(m.input_from_values(CASES["tiny"]))
    assert exchange_plan(tiny, grid) == []


def test_build_tracks_matches_the_oracle_on_random_configurations(built):
    """The digests above pin moc_build_tracks on the named cases; here the same arrays are compared with the
    oracle's construction (itself bit-identical to the reference's on these very configurations:
    tests/test_oracle_vs_ref.py) for 80 random inputs -- 1-6 coarse / fine axial intervals, flat and quadratic
    source, odd group counts, decomposed nodes."""
    from oracle_lib import OracleCase
    from test_oracle_vs_ref import _random_configurations
    checked = 0
    for vals, seed in _random_configurations(80, 20261017):
        host = m.HostProblem(m.derive(m.input_from_values(vals)), seed=seed)
        o = OracleCase(vals, seed=seed)
        if o.n_segments.min() >= 0:     # negative draws: the library clamps what the reference leaves undefined
            what = f"vals={vals} seed={seed}"
            assert host.rand_calls == o.init_rand_calls, what
            for name, mine, theirs in (("az_weight", api.HOST_AZ_WEIGHT, o.az_weight), ("n_segments", api.HOST_N_SEGMENTS, o.n_segments),
                                       ("seg_lengths", api.HOST_SEG_LENGTHS, o.seg_lengths), ("p_weight", api.ARR_P_WEIGHT, o.p_weight),
                                       ("z_height", api.ARR_Z_HEIGHT, o.z_height), ("xs", api.HOST_XS, o.xs),
                                       ("scatter", api.HOST_SCATTER, o.scatter), ("fine_source", api.ARR_FINE_SOURCE, o.fine_source),
                                       ("sigT", api.ARR_SIGT, o.sigT), ("xs_index", api.HOST_XS_INDEX, o.xs_index),
                                       ("vol", api.HOST_VOL, o.vol)):
                assert np.array_equal(np.asarray(host.get(mine)).ravel(), np.asarray(theirs).ravel()), f"{name}: {what}"
            checked += 1
        host.close(); o.close()
    assert checked >= 60


@pytest.mark.parametrize("case", ["tiny", "mini104", "tiny_flat", "tall", "ragged", "polar1"])
def test_oracle_geometry_only_and_reversed_order_keep_every_integer(case):
    """Two variants of the oracle's sweep used as yardsticks (never as the parity oracle itself):
    geometry-only (oracle_create_geometry: what pins the FULL-SIZE problems' integers,
    tests/golden/make_full_size_counts.py) and 2D tracks in reverse order (tools/tolerance_anchor.py: the same
    tallies added in another order).  Both must reproduce the full oracle's segment counts, draws, digest and ray
    heights bit for bit over two sweeps; the reversed order also its angular flux (only the scalar flux is order-
    dependent)."""
    from oracle_lib import CASES, OracleCase
    seed = 2 if case in ("ragged", "polar1") else 3
    full, geo, rev = (OracleCase(CASES[case], seed=seed), OracleCase(CASES[case], seed=seed, geometry_only=True),
                      OracleCase(CASES[case], seed=seed))
    assert full.init_rand_calls == geo.init_rand_calls
    for sweep in range(2):
        n = full.sweep()
        assert geo.sweep() == n and rev.sweep_reversed() == n
        for other in (geo, rev):
            assert other.rand_calls == full.rand_calls
            assert np.array_equal(other.seg_count, full.seg_count)
            assert np.array_equal(other.digest, full.digest)
            assert np.array_equal(other.z_height, full.z_height)
        if sweep == 0:
            assert np.array_equal(rev.psi, full.psi)
            assert rel_l2_(rev.fine_flux, full.fine_flux) < 1e-5
            rev.fine_flux[...] = full.fine_flux          # continue from identical tallies
    full.close(); geo.close(); rev.close()


def rel_l2_(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))
