"""The C driver SimpleMOC-b200 (simplemoc_b200/csrc/driver_main.c): the reference's main() (src/main.c:3-147)
written against include/moc_b200.h.  Same command line (-i, -s, -d, -t), same loop, same report."""
import os
import re
import subprocess

import pytest

from oracle_lib import CASES, OracleCase, write_input_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "simplemoc_b200", "SimpleMOC-b200")
TRACKS = os.path.join(ROOT, "tests", "golden", "tracks_44.bin")


def run_driver(*args, expect_ok=True):
    p = subprocess.run([DRIVER, *args], capture_output=True, text=True, timeout=600)
    if expect_ok:
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p


def test_driver_without_a_gpu_stops_with_an_error(built, tmp_path):
    """no CPU fallback: without a CUDA device the program says so and exits 1 (it never computes)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    f = write_input_file(str(tmp_path / "tiny.in"), CASES["tiny"])
    p = run_driver("-i", f, expect_ok=False)
    assert p.returncode == 1
    assert "no CUDA device" in p.stderr and "keff" not in p.stdout


def test_driver_rejects_a_bad_command_line(built):
    p = run_driver("-i", expect_ok=False)          # src/io.c:184-195: usage + exit
    assert p.returncode != 0 and "usage" in (p.stdout + p.stderr).lower()


@pytest.mark.gpu
@pytest.mark.parametrize("case,extra,track_file", [("tiny", [], None), ("mini104", ["--host-buffers"], None),
                                                   ("tiny", [], TRACKS), ("tiny_flat", ["--iters", "2"], None)])
def test_driver_reports_what_the_oracle_computes(built, tmp_path, case, extra, track_file):
    """segments processed (bit-exact) and k-eff (1e-4) of every iteration, printed by the driver, against the
    oracle run through the same loop (main.c:57-92: sweep, renormalise, update with the previous k, k-eff)"""
    f = write_input_file(str(tmp_path / f"{case}.in"), CASES[case])
    args = ["-i", f, "--seed", "11", *extra] + (["-d", track_file] if track_file else [])
    out = run_driver(*args).stdout
    iters = int(extra[extra.index("--iters") + 1]) if "--iters" in extra else 1
    o = OracleCase(CASES[case], seed=11, track_file=track_file)
    k, segs, ks = 1.0, 0, []
    for _ in range(iters):
        segs += o.sweep()
        o.renormalize()
        o.update_sources(k)
        k = o.compute_keff()
        ks.append(k)
    assert int(re.search(r"Segments processed:\s+(\d+)", out).group(1)) == segs
    printed = [float(x) for x in re.findall(r"^keff = (\S+)", out, flags=re.M)]
    assert len(printed) == iters
    assert len(re.findall(r"source residual = ", out)) == (iters if iters > 1 else 0)
    for a, b in zip(printed, ks):
        assert abs(a - b) <= 1e-4 * abs(b) + 5e-7, (printed, ks)      # "%f" prints six decimals
    assert f"3D tracks:" in out and str(o.I.ntracks) in out
    if track_file:
        assert "Reading track data from" in out and "2D tracks:" in out
    o.close()


REF_MAIN = os.path.join(ROOT, "oracle", "_ref", "SimpleMOC-dropin")


@pytest.mark.gpu
@pytest.mark.parametrize("case,seed", [("tiny", 11), ("mini104", 4), ("tiny_flat", 7)])
def test_the_reference_own_main_linked_against_the_library(built, tmp_path, case, seed):
    """oracle/_ref/SimpleMOC-dropin = the reference's main.c, init.c, io.c, tracks.c, source.c, utils.c --
    unmodified -- with solver.c and comms.c replaced by libmoc_b200.so (oracle/Makefile `dropin`).  It prints
    the k-eff the all-CPU reference computes from the same pinned random stream."""
    from oracle_lib import RefCase, ref_available
    if not (os.path.exists(REF_MAIN) and ref_available()):
        pytest.skip("oracle/_ref/SimpleMOC-dropin did not travel")
    f = write_input_file(str(tmp_path / f"{case}.in"), CASES[case])
    p = subprocess.run([REF_MAIN, "-i", f], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, SMOC_SEED=str(seed)))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    k_gpu = float(re.search(r"^keff = (\S+)", p.stdout, flags=re.M).group(1))
    r = RefCase(CASES[case], seed=seed)
    r.sweep(); r.renormalize(); r.update_sources(1.0)
    k_cpu = r.compute_keff()
    assert abs(k_gpu - k_cpu) <= 1e-4 * abs(k_cpu) + 5e-7, (k_gpu, k_cpu)
    assert "Time per Intersection" in p.stdout          # the reference's own report (main.c:130-134)
    r.close()
