"""Parity statistics of the CUDA path against the CPU oracle, phase by phase (diagnostic, not a test):
python tools/parity_stats.py [case ...]   (library: MOC_B200_LIB or the in-tree one)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import simplemoc_b200 as m
from simplemoc_b200 import api
from oracle_lib import CASES, OracleCase, frac_within, rel_l2

cases = sys.argv[1:] or ["tiny", "mini104", "odd", "mini_default_in"]
for case in cases:
    for seed in (11, 3):
        vals = CASES[case]
        host = m.HostProblem(m.derive(m.input_from_values(vals)), seed=seed)
        dev = m.DeviceProblem(host, device=0)
        ora = OracleCase(vals, seed=seed)
        line = []
        for sw in range(3):
            assert dev.sweep() == ora.sweep()
            f, o = dev.get(api.ARR_FINE_FLUX), ora.fine_flux
            line.append(f"sweep{sw+1}: relL2 {rel_l2(f, o):.2e} frac {frac_within(f, o, 1e-4):.5f}")
            dev.renormalize(); ora.renormalize()
            dev.update_sources(1.0); ora.update_sources(1.0)
            s, so = dev.get(api.ARR_FINE_SOURCE), ora.fine_source
            line.append(f"src frac {frac_within(s, so, 1e-4):.5f}")
            dev.compute_keff(); ora.compute_keff()
        print(f"{case:16s} seed {seed:2d}  " + "  ".join(line), flush=True)
        dev.close(); host.close(); ora.close()
