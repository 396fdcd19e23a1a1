"""Diagnostic: the scalar-flux elements of a BASELINE-shape slice that differ most from the oracle, with the two
accumulation scales of each (sum of |tally|; sum of the magnitudes of the terms inside every tally) and the segments
that tallied into them.  python tools/diag_worst.py <shape> [seed]   (shapes: tests/test_gpu_baseline_shapes.py)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import simplemoc_b200 as m
from simplemoc_b200 import api
from oracle_lib import OracleCase, frac_within, noise_units, rel_l2
from test_gpu_baseline_shapes import SHAPES

name = sys.argv[1]
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
values, keep, Z, N, _ = SHAPES[name]
inp = m.derive(m.input_from_values(values), limit_tracks_2D=keep)
host = m.HostProblem(inp, seed=seed)
dev = m.DeviceProblem(host, device=0)
if len(sys.argv) > 3:
    dev.set_option(api.OPT_FIT_PER_SEGMENT, int(sys.argv[3]))
ora = OracleCase(values, seed=seed, limit_tracks_2D=keep)
ora.enable_trace(1 << 23)
assert dev.sweep() == ora.sweep()
f, o = dev.get(api.ARR_FINE_FLUX), ora.fine_flux
G, F = inp.n_egroups, inp.fai
diff = np.abs(f.astype(np.float64) - o)
rel_ok = diff <= 1e-4 * np.abs(o)
u_tally = noise_units(f, o, ora.abs_flux).reshape(f.shape)
u_terms = noise_units(f, o, ora.abs_terms).reshape(f.shape)
print(f"{name} seed {seed}: fit per segment {dev.get_option(api.OPT_FIT_PER_SEGMENT)}; relL2 {rel_l2(f, o):.2e}, "
      f"within 1e-4: {frac_within(f, o, 1e-4):.5f}; elements outside 1e-4: {int((~rel_ok).sum())} of {rel_ok.size}")
for k in (4, 16, 64, 256):
    print(f"  outside 1e-4 AND more than {k:3d} eps of sum|tally|: {int((~rel_ok & (u_tally > k)).sum()):6d}   "
          f"of the sum of term magnitudes: {int((~rel_ok & (u_terms > k)).sum()):6d}")
tt, tr, tds, tz = ora.trace()
order = np.argsort(np.where(rel_ok, 0.0, u_tally).ravel())[::-1][:6]
sig = ora.sigT
for e in order:
    r, g = divmod(int(e), G)
    reg, fine = divmod(r, F)
    segs = np.nonzero(tr == r)[0]
    print(f"element region {reg} fine {fine} group {g}: gpu {f.ravel()[e]:.7e} oracle {o.ravel()[e]:.7e} diff {diff.ravel()[e]:.2e} "
          f"sum|tally| {ora.abs_flux.ravel()[e]:.3e} ({u_tally.ravel()[e]:.0f} eps) sum|terms| {ora.abs_terms.ravel()[e]:.3e} "
          f"({u_terms.ravel()[e]:.1f} eps); sigT {sig[reg, g]:.4e}; {len(segs)} segments, ds " +
          " ".join(f"{tds[s]:.3e}" for s in segs[:8]))
