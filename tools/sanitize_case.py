"""One small problem through every kernel of the library (both ray-trace kernels, both exponential modes,
flat and quadratic source, reductions, overlapped emit), for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import simplemoc_b200 as m
from simplemoc_b200 import api
from oracle_lib import CASES

for case, opts in (("tiny", {}), ("tiny", {api.OPT_WALK_KERNEL: 1}), ("tiny_flat", {}), ("odd", {api.OPT_FIT_PER_SEGMENT: 1}),
                   ("ragged", {api.OPT_FILL_OVERLAP: 1, api.OPT_FILL_BATCHES: 4}), ("mini104", {api.OPT_EXP_MODE: 1})):
    host = m.HostProblem(m.derive(m.input_from_values(CASES[case])), seed=3)
    dev = m.DeviceProblem(host, device=0)
    for k, v in opts.items():
        dev.set_option(k, v)
    n = dev.sweep(); dev.renormalize(); r = dev.update_sources(1.0); k = dev.compute_keff()
    n2 = dev.sweep()
    print(case, opts, n, n2, k, flush=True)
    dev.close(); host.close()
syn = m.DeviceProblem.synthetic(m.derive(m.input_from_values(CASES["tiny"])), seed=3, device=0)
print("synthetic", syn.sweep(), flush=True)
syn.close()
