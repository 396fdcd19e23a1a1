"""One FULL-SIZE problem (default: the reference's `-s` problem, 1.2 M tracks, 1.41e8 segments x 104 groups) on the GPU
against the serial CPU oracle -- every number, not a slice: segment total, per-track counts, region digest and ray heights
bit-exact; scalar flux, angular flux and sources by rel-L2 and fraction within 1e-4; k-eff.  ~3 minutes of one host
core for the oracle.   python tools/full_size_parity.py [small|default_in] [> profiles/rNN_full_size_parity_small.log]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import simplemoc_b200 as m
from simplemoc_b200 import api
from oracle_lib import OracleCase, frac_within, noise_units, rel_l2

PROBLEMS = {
    "small": [15, 15, 5, 3, 2, 0.5, 0.2, 5, 5, 104, 0, 1, 120, 1.26 * 17, 400.0, 0.01, 3000, 0],            # init.c:77-103
    "default_in": [17, 17, 9, 5, 2, 0.05, 0.25, 64, 10, 100, 1, 20, 20, 21.42, 400.0, 0.01, 5000, 0],       # default.in
}
name = sys.argv[1] if len(sys.argv) > 1 else "small"
values, seed = PROBLEMS[name], 1
inp = m.derive(m.input_from_values(values))
t0 = time.time()
host = m.HostProblem(inp, seed=seed)
dev = m.DeviceProblem(host, device=0)
dev.set_option(api.OPT_DIGEST, 1)
ora = OracleCase(values, seed=seed)
print(f"{name}: {inp.ntracks} tracks, {inp.n_source_regions_per_node} source regions, G = {inp.n_egroups}; built in {time.time() - t0:.0f} s", flush=True)
t0 = time.time(); n_gpu = dev.sweep(); t_gpu = time.time() - t0
t0 = time.time(); n_cpu = ora.sweep(); t_cpu = time.time() - t0
print(f"segments: gpu {n_gpu} oracle {n_cpu} ({'bit-exact' if n_gpu == n_cpu else 'DIFFERENT'}); sweep {dev.timing().total_ms:.1f} ms on the GPU, "
      f"{t_cpu:.0f} s on one host core")
print("per-track segment counts:", "bit-exact" if np.array_equal(dev.get(api.ARR_SEG_COUNT), ora.seg_count) else "DIFFERENT")
print("(serial index, tally row) digest:", "bit-exact" if np.array_equal(dev.get(api.ARR_QSR_DIGEST), ora.digest) else "DIFFERENT")
print("ray heights after the sweep:", "bit-exact" if np.array_equal(dev.get(api.ARR_Z_HEIGHT), ora.z_height) else "DIFFERENT")
flux = dev.get(api.ARR_FINE_FLUX)
rel_ok = np.abs(flux.astype(np.float64) - ora.fine_flux) <= 1e-4 * np.abs(ora.fine_flux)
units = noise_units(flux, ora.fine_flux, ora.abs_terms).reshape(flux.shape)
print(f"scalar flux: rel-L2 {rel_l2(flux, ora.fine_flux):.2e}, within 1e-4: {frac_within(flux, ora.fine_flux, 1e-4):.5f}, worst element outside 1e-4: "
      f"{(units[~rel_ok].max() if (~rel_ok).any() else 0):.1f} eps x its running error scale")
psi = dev.get(api.ARR_PSI)
print(f"angular flux: rel-L2 {rel_l2(psi, ora.psi):.2e}, within 1e-4: {frac_within(psi, ora.psi, 1e-4):.5f}")
dev.renormalize(); ora.renormalize()
r_gpu, r_cpu = dev.update_sources(1.0), ora.update_sources(1.0)
src = dev.get(api.ARR_FINE_SOURCE)
print(f"sources after renormalise + update_sources: rel-L2 {rel_l2(src, ora.fine_source):.2e}, within 1e-4: {frac_within(src, ora.fine_source, 1e-4):.5f}; "
      f"residual gpu {r_gpu:.6e} oracle {r_cpu:.6e}")
k_gpu, k_cpu = dev.compute_keff(), ora.compute_keff()
print(f"k-eff: gpu {k_gpu:.7f} oracle {k_cpu:.7f} relative difference {abs(k_gpu - k_cpu) / abs(k_cpu):.2e}")
n2g, n2c = dev.sweep(), ora.sweep()
print(f"second sweep: segments gpu {n2g} oracle {n2c}; counts {'bit-exact' if np.array_equal(dev.get(api.ARR_SEG_COUNT), ora.seg_count) else 'DIFFERENT'}; "
      f"digest {'bit-exact' if np.array_equal(dev.get(api.ARR_QSR_DIGEST), ora.digest) else 'DIFFERENT'}")
