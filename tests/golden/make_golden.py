"""Generates tests/golden/reference_digests.json by running the UNMODIFIED reference
(oracle/_ref/libsimplemoc_ref.so = /root/reference/src compiled by oracle/Makefile with
rand()/time() pinned by oracle/ref_shim.c) on the named small cases.

The reference ships no golden vectors; these digests are the pin that travels
to machines where /root/reference does not exist (the GPU box).  Run in the build container:

    python tests/golden/make_golden.py

A digest is a SHA-256 of the raw bytes of each output array, plus a few scalars kept in
clear so that a mismatch can be localised.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import CASES, RefCase, ensure_ref_built  # noqa: E402

SEEDS = {name: 11 for name in CASES}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


TRACK_FILE = os.path.join(HERE, "tracks_44.bin")   # tests/golden/make_track_file.py


def run_case(name, variant="", track_file=None):
    r = RefCase(CASES[name], seed=SEEDS[name], variant=variant, track_file=track_file)
    out = {"seed": SEEDS[name], "values": CASES[name], "init_rand_calls": int(r.init_rand_calls)}
    I = r.I
    out["derived"] = {"ntracks_2D": I.ntracks_2D, "z_stacked": I.z_stacked, "ntracks": I.ntracks,
                      "n_source_regions_per_node": I.n_source_regions_per_node}
    if track_file:
        # what load_OpenMOC_tracks (tracks.c:170-323) changes in the Input
        out["track_file"] = os.path.basename(track_file)
        out["derived"].update({"n_azimuthal": I.n_azimuthal, "radial_ray_sep": float(I.radial_ray_sep),
                               "segments_per_track": I.segments_per_track})
    az, ns, ln = r.tracks_2D()
    pw, zh = r.tracks()
    idx, vol = r.source_meta()
    out["init"] = {"az_weight": sha(az), "n_segments": sha(ns), "seg_lengths": sha(ln),
                   "p_weight": sha(pw), "z_height": sha(zh), "xs": sha(r.xs), "scatter": sha(r.scatter),
                   "fine_source": sha(r.fine_source), "sigT": sha(r.sigT), "xs_index": sha(idx),
                   "vol": sha(vol), "table": sha(r.table[0])}
    out["segments_processed"] = int(r.sweep())
    out["sweep_rand_calls"] = int(r.rand_calls - r.init_rand_calls)
    out["after_sweep"] = {"fine_flux": sha(r.fine_flux), "psi": sha(r.psi), "z_height": sha(r.z_height),
                          "fine_flux_sum": float(np.sum(r.fine_flux, dtype=np.float64)),
                          "fine_flux_l2": float(np.linalg.norm(r.fine_flux.astype(np.float64)))}
    r.renormalize()
    out["after_renormalize"] = {"fine_flux": sha(r.fine_flux), "psi": sha(r.psi)}
    res = r.update_sources(1.0)
    out["after_update_sources"] = {"residual": float(res), "fine_source": sha(r.fine_source)}
    out["keff"] = float(r.compute_keff())
    r.close()
    return out


def main():
    assert ensure_ref_built(), "oracle/_ref is not built and /root/reference is absent"
    golden = {"generator": "tests/golden/make_golden.py",
              "reference": "ANL-CESAR/SimpleMOC v4 sources, unmodified, -O2 -ffp-contract=off, serial, "
                           "rand()=moc_rand31(seed, call index) (oracle/ref_shim.c)",
              "table": {n: run_case(n) for n in CASES},
              "expf": {n: run_case(n, "_expf") for n in ("tiny", "mini104")},
              # 2D tracks read by the reference's own load_OpenMOC_tracks from tests/golden/tracks_44.bin
              "tracks": {n: run_case(n, track_file=TRACK_FILE) for n in ("tiny", "mini104", "tiny_flat")}}
    with open(os.path.join(HERE, "reference_digests.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    print("wrote reference_digests.json:", {k: v["segments_processed"] for k, v in golden["table"].items()})


if __name__ == "__main__":
    main()
