"""Run-to-run spread of the parity statistics (atomic tally order is not deterministic): diagnostic."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import simplemoc_b200 as m
from simplemoc_b200 import api
from oracle_lib import CASES, OracleCase, frac_within, rel_l2
case, seed, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
vals = CASES[case]
for sync in (0, 1):
    fr = []
    for r in range(reps):
        host = m.HostProblem(m.derive(m.input_from_values(vals)), seed=seed)
        dev = m.DeviceProblem(host, device=0)
        ora = OracleCase(vals, seed=seed)
        dev.sweep(); ora.sweep()
        dev.renormalize(); ora.renormalize(); dev.update_sources(1.0); ora.update_sources(1.0)
        dev.compute_keff(); ora.compute_keff()
        if sync:   # second sweep from the oracle's state: no amplified first-sweep rounding
            dev.set(api.ARR_FINE_SOURCE, ora.fine_source); dev.set(api.ARR_PSI, ora.psi); dev.set(api.ARR_FINE_FLUX, ora.fine_flux)
        dev.sweep(); ora.sweep()
        f, o = dev.get(api.ARR_FINE_FLUX), ora.fine_flux
        fr.append((frac_within(f, o, 1e-4), rel_l2(f, o)))
        dev.close(); host.close(); ora.close()
    print(case, seed, "synced" if sync else "free", "frac min/median/max %.5f %.5f %.5f" % (min(x[0] for x in fr), sorted(x[0] for x in fr)[len(fr)//2], max(x[0] for x in fr)),
          "relL2 max %.2e" % max(x[1] for x in fr), flush=True)
