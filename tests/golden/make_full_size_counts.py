"""Generates tests/golden/full_size_counts.json: the INTEGER results of one transport sweep of the full-size
BASELINE problems (no slice), computed by the CPU oracle in geometry-only mode (oracle/moc_oracle.c
oracle_create_geometry: the ray trace, the source-region draws and the digest of solver.c:347-529 without the
attenuation arithmetic or the 13-129 GB of flux arrays).  The oracle's ray trace is bit-identical to the
reference's on every case of tests/test_oracle_vs_ref.py; the geometry-only mode is checked against the full
oracle in tests/test_host.py.  The GPU test tests/test_gpu_full_size.py compares the CUDA path against this file.

python tests/golden/make_full_size_counts.py [name ...]      (minutes of CPU per problem; config5 ~ 10x default)
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
from oracle_lib import OracleCase

PROBLEMS = {
    # the 18 input-file values (io.c:210-267 order)
    "default": [17, 17, 27, 5, 2, 0.05, 0.25, 64, 10, 104, 1, 20, 120, 21.42, 400.0, 0.01, 5000, 0],      # init.c:33-74
    "small": [15, 15, 5, 3, 2, 0.5, 0.2, 5, 5, 104, 0, 1, 120, 1.26 * 17, 400.0, 0.01, 3000, 0],            # init.c:77-103
    "default_in": [17, 17, 9, 5, 2, 0.05, 0.25, 64, 10, 100, 1, 20, 20, 21.42, 400.0, 0.01, 5000, 0],       # default.in
    "config5": [17, 17, 27, 5, 2, 0.05, 0.25, 64, 10, 104, 1, 2, 120, 21.42, 400.0, 0.01, 5000, 0],         # decomp_assemblies_ax = 2
}
SEED = 1
OUT = os.path.join(HERE, "full_size_counts.json")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    names = sys.argv[1:] or ["small", "default_in", "default"]
    table = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        t0 = time.time()
        o = OracleCase(PROBLEMS[name], seed=SEED, geometry_only=True)
        I = o.I
        row = {"values": PROBLEMS[name], "seed": SEED, "ntracks_2D": I.ntracks_2D, "z_stacked": I.z_stacked, "ntracks": I.ntracks,
               "n_source_regions_per_node": I.n_source_regions_per_node, "init_rand_calls": int(o.init_rand_calls), "sweeps": []}
        for sweep in range(2):
            n = o.sweep()
            row["sweeps"].append({"segments_processed": int(n), "rand_calls": int(o.rand_calls),
                                  "digest": [int(v) for v in o.digest], "seg_count_sha256": sha(o.seg_count),
                                  "z_height_sha256": sha(o.z_height), "longest_track": int(o.seg_count.max())})
            print(f"{name} sweep {sweep + 1}: {n} segments, {time.time() - t0:.0f} s", flush=True)
        o.close()
        table[name] = row
        with open(OUT, "w") as f:
            json.dump(table, f, indent=1)


if __name__ == "__main__":
    main()
