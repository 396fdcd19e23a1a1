"""Timing of the resident sweep with the ray trace's emitting pass under the attenuation (diagnostic):
python tools/probe_overlap.py [exp_mode]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplemoc_b200 as m
from simplemoc_b200 import api

exp_mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
inp = m.default_input()
if len(sys.argv) > 2:
    inp.n_egroups = int(sys.argv[2])          # python tools/probe_overlap.py <exp_mode> <G>
inp = m.derive(inp)
dev = m.DeviceProblem.synthetic(inp, seed=1, device=0, exp_mode=exp_mode)
for ctas, batches in ((0, 8), (1, 8), (2, 8), (3, 8), (4, 8), (2, 16), (0, 8)):
    dev.set_option(api.OPT_FILL_OVERLAP, ctas)
    dev.set_option(api.OPT_FILL_BATCHES, batches)
    res = []
    for rep in range(3):
        n = dev.sweep()
        t = dev.timing()
        res.append(t.total_ms)
    print(f"exp_mode={exp_mode} overlap_ctas={ctas} batches={batches}: total {min(res[1:]):.1f} ms (runs {', '.join('%.1f' % r for r in res)}); last: count {t.count_ms:.1f} "
          f"fill {t.fill_ms:.1f} attenuate {t.attenuate_ms:.1f} n_batches {t.n_batches} launches {t.launches}", flush=True)
