"""Quick timing probe of the resident sweep (not the bench): python tools/probe.py [small|default|<case>] [limit_tracks_2D]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import simplemoc_b200 as m
from simplemoc_b200 import api

which = sys.argv[1] if len(sys.argv) > 1 else "default"
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if which == "default":
    inp = m.default_input()
elif which == "small":
    inp = m.small_input()
else:
    from oracle_lib import CASES
    inp = m.input_from_values(CASES[which])
inp = m.derive(inp, limit)
t0 = time.time()
host = m.HostProblem(inp, seed=1)
t1 = time.time()
dev = m.DeviceProblem(host)
t2 = time.time()
print(f"{which}: T2={inp.ntracks_2D} Z={inp.z_stacked} T3={inp.ntracks} G={inp.n_egroups} N={inp.n_source_regions_per_node} "
      f"host build {t1-t0:.1f}s upload {t2-t1:.1f}s", flush=True)
for variant in (1,):
    for mode in (0, 1, 0, 1):
        dev.set_option(api.OPT_EXP_MODE, mode)
        n = dev.sweep()
        t = dev.timing()
        integ = n * inp.n_egroups
        print(f"  kernel={variant} exp_mode={mode} segments={n} total {t.total_ms:.1f} ms (count {t.count_ms:.1f} scan {t.scan_ms:.2f} fill {t.fill_ms:.1f} "
              f"attenuate {t.attenuate_ms:.1f}; batches {t.n_batches}) -> {integ/t.total_ms/1e6:.1f} G integ/s, "
              f"{t.total_ms*1e6/integ:.5f} ns/integ", flush=True)
t3 = time.time(); dev.renormalize(); t4 = time.time(); r = dev.update_sources(1.0); t5 = time.time(); k = dev.compute_keff(); t6 = time.time()
print(f"  renormalize {1e3*(t4-t3):.2f} ms, update_sources {1e3*(t5-t4):.2f} ms, keff {1e3*(t6-t5):.2f} ms  (res {r:.4g}, keff {k:.6f})")
