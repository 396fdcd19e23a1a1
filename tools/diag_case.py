"""Diagnostic: parity statistics and bit checksums of one case, repeated (python tools/diag_case.py case seed reps)."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import simplemoc_b200 as m
from simplemoc_b200 import api
from oracle_lib import CASES, OracleCase, frac_within, noise_units, rel_l2
case, seed, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
vals = CASES[case]
ora = OracleCase(vals, seed=seed)
ora.sweep()
for r in range(reps):
    host = m.HostProblem(m.derive(m.input_from_values(vals)), seed=seed)
    dev = m.DeviceProblem(host, device=0)
    dev.sweep()
    out = []
    for name, arr, ref in (("flux", api.ARR_FINE_FLUX, ora.fine_flux), ("psi", api.ARR_PSI, ora.psi)):
        a = dev.get(arr)
        out.append(f"{name} frac {frac_within(a, ref, 1e-4):.5f} relL2 {rel_l2(a, ref):.2e} crc {zlib.crc32(np.ascontiguousarray(a).tobytes()):08x}")
    f = dev.get(api.ARR_FINE_FLUX)
    units = noise_units(f, ora.fine_flux, ora.abs_flux)
    rel_ok = np.abs(f.astype(np.float64) - ora.fine_flux).ravel() <= 1e-4 * np.abs(np.asarray(ora.fine_flux, np.float64)).ravel()
    out.append("flux within 1e-4 or k eps of own accumulation: " + " ".join(f"k={k}: {float((rel_ok | (units <= k)).mean()):.5f}" for k in (2, 4, 8, 16))
               + f" worst {units[~rel_ok].max() if (~rel_ok).any() else 0:.0f}")
    z = dev.get(api.ARR_Z_HEIGHT)
    out.append(f"z crc {zlib.crc32(np.ascontiguousarray(z).tobytes()):08x} digest_ok {np.array_equal(dev.get(api.ARR_QSR_DIGEST), ora.digest)}")
    print(case, seed, os.environ.get("MOC_B200_LIB", "default"), " | ".join(out), flush=True)
    dev.close(); host.close()
