#!/usr/bin/env python
"""bench.py -- segment-group integrations/sec of the 3D MOC transport sweep on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl moc|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" is one iteration of the reference's loop (reference src/main.c:57-92) over one
spatial domain per GPU: transport_sweep -> [fast_transfer_boundary_fluxes when N > 1] ->
renormalize_flux -> update_sources -> compute_keff.  The metric is the reference's own
(src/utils.c:147-155): integrations = segments_processed * n_egroups; `value` is the
whole-job integrations per second of step time (all N domains; weak scaling: every rank owns
a full-size domain, exactly as every MPI rank of the reference does).

Workload at N = 1: the default strawman problem (BASELINE.json configs[1]): the built-in
defaults of set_default_input (src/init.c:33-74; G=104, 120 2D segments per track,
15.5 M 3D tracks, 12.9 GB of angular flux, ~1.93e9 3D segments = 2.0e11 integrations per
sweep).  `--workload default_in` runs the shipped default.in values instead (G=100, 20
segments per track); `--workload small` the -s problem.

The JSON line also carries
  e2e          the same metric through the reference's own entry point
               transport_sweep(Params*, Input*) on HOST structures: every step uploads the
               mutable state from pinned host memory and downloads what the reference
               function mutates (inside the timed region)
  roofline     the attenuation kernel (the dominant launch) against measured HBM bandwidth,
               plus its FP32 and L2-level rates (what actually bounds it)
  cpu_baseline the UNMODIFIED reference (oracle/_ref, OpenMP, all host cores) timed here on
               a bounded sample of the same workload (N = 1 only)
  clocks       SM clocks / throttle reasons sampled with nvidia-smi during the timed region

`--impl reference` times the reference's own CPU implementation (oracle/_ref OpenMP build;
falls back to the serial C restatement oracle/_build when the reference library did not
travel) -- the only place this file touches oracle/.
"""
import argparse
import ctypes as C
import glob
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "segment-group integrations/sec"
UNIT = "integrations/s"
FLOP_PER_INTEGRATION = 62      # attenuate_fluxes as written, axial_exp = 2
L2_BYTES_PER_INTEGRATION = 24  # 3 source rows + sigT (16 B gathered) + 8 B tally read-modify-write
FP32_PEAK_TFLOPS_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12   # 74.4: 148 SMs x 128 FMA lanes x 1965 MHz

# values of the shipped default.in (reference src/default.in:1-18), in input-file order
DEFAULT_IN = [17, 17, 9, 5, 2, 0.05, 0.25, 64, 10, 100, 1, 20, 20, 21.42, 400.0, 0.01, 5000, 0]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="moc", choices=["moc", "reference"])
    ap.add_argument("--workload", default="default", choices=["default", "default_in", "small"])
    ap.add_argument("--exp", default="table", choices=["table", "sfu"],
                    help="table = the reference's exponential table (parity mode, default); "
                         "sfu = MUFU.EX2 (__expf)")
    ap.add_argument("--limit-tracks-2d", type=int, default=0,
                    help="shrink the number of 2D tracks (profiling under ncu only; not a bench value)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0,
                    help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--egroups", type=int, default=0,
                    help="override n_egroups (BASELINE config 4: 32 / 64 / 128 on the default geometry)")
    ap.add_argument("--decomp-ax", type=int, default=0,
                    help="override decomp_assemblies_ax (BASELINE config 5: 2 = ~131 GB per GPU)")
    ap.add_argument("--device-build", action="store_true",
                    help="generate the problem on the device (moc_create_synthetic); implies --no-e2e: there "
                         "are no host structures to run the drop-in call on")
    ap.add_argument("--grid", default="", help="cx,cy,cz (default: 1x1x1, 2x1x1, 2x2x1, 2x2x2)")
    ap.add_argument("--full-size", action="store_true",
                    help="--impl reference: sweep the WHOLE workload instead of a bounded sample (minutes per step; "
                         "the once-per-round check of the sampled figure, profiles/)")
    ap.add_argument("--no-full-loop", action="store_true", help="skip the e2e_full_loop leg")
    ap.add_argument("--noclamp", type=int, default=1, choices=[0, 1],
                    help="0: keep the table's x > maxVal test in every attenuation launch (A/B timing)")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="N > 1: do not bind each rank to the NUMA node of its GPU")
    ap.add_argument("--nccl-max-ctas", type=int, default=8,
                    help="N > 1: NCCL_MAX_CTAS for the boundary exchange (unless the environment already sets it)")
    ap.add_argument("--staged", type=int, default=1, choices=[0, 1],
                    help="1 (default): TMA-staged attenuation kernel; 0: the direct-gather kernel (A/B timing)")
    ap.add_argument("--fill-overlap", type=int, default=0,
                    help="ray-trace CTAs per SM emitting batch b+1's records under the attenuation of batch b (A/B timing; "
                         "0: the ray trace runs in front of the attenuation)")
    ap.add_argument("--fill-batches", type=int, default=0, help="with --fill-overlap: batches per sweep")
    return ap.parse_args()


# ------------------------------------------------------------------ helpers

def workload_input(m, name, egroups=0, decomp_ax=0):
    inp, label = _workload_input(m, name)
    if egroups:
        inp.n_egroups = egroups
        label += f"; n_egroups overridden to {egroups}"
    if decomp_ax:
        inp.decomp_assemblies_ax = decomp_ax
        label += f"; decomp_assemblies_ax overridden to {decomp_ax}"
    return inp, label


def _workload_input(m, name):
    if name == "default":
        inp = m.default_input()
    elif name == "default_in":
        inp = m.input_from_values(DEFAULT_IN)
    else:
        inp = m.small_input()
    return inp, WORKLOAD_LABELS[name]


def bind_to_gpu_numa_node(torch, local):
    """One process per GPU on a two-socket host: run this rank's host threads -- and, through first touch, place its
    pinned host buffers (7 GB up + 7 GB down per step on the default problem) -- on the NUMA node the GPU hangs off,
    instead of wherever the launcher left the process.  Returns a short description for the JSON line."""
    try:
        props = torch.cuda.get_device_properties(local)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return f"gpu {bus}: no NUMA node reported"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return f"gpu {bus}: node {node} has no CPU this process may use"
        os.sched_setaffinity(0, allowed)
        return f"gpu {bus}: bound to NUMA node {node} ({len(allowed)} CPUs)"
    except (OSError, ValueError, AttributeError) as e:
        return f"not bound ({type(e).__name__}: {e})"


def grid_for(n, spec):
    if spec:
        cx, cy, cz = (int(v) for v in spec.split(","))
    else:
        cx, cy, cz = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(n, (n, 1, 1))
    assert cx * cy * cz == n, f"grid {cx}x{cy}x{cz} does not have {n} domains"
    return cx, cy, cz


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="moc_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def finite(obj):
    """strict JSON has no NaN/Infinity: non-finite floats become null"""
    if isinstance(obj, float):
        return obj if obj == obj and abs(obj) != float("inf") else None
    if isinstance(obj, dict):
        return {k: finite(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [finite(v) for v in obj]
    return obj


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ the reference on the host cores

def reference_input_struct():
    """ctypes image of the reference's Input (src/SimpleMOC_header.h:28-76), declared here so that the reference
    arm does not import the product package (its process then maps oracle/_ref only)"""
    class Input(C.Structure):
        _fields_ = [
            ("x_assemblies", C.c_int), ("y_assemblies", C.c_int), ("cai", C.c_int), ("fai", C.c_int),
            ("axial_exp", C.c_int), ("radial_ray_sep", C.c_float), ("axial_z_sep", C.c_float),
            ("n_azimuthal", C.c_int), ("n_polar_angles", C.c_int), ("n_egroups", C.c_int), ("decompose", C.c_bool),
            ("decomp_assemblies_ax", C.c_int), ("segments_per_track", C.c_long), ("assembly_width", C.c_float),
            ("height", C.c_float), ("domain_height", C.c_float), ("precision", C.c_float), ("mype", C.c_long),
            ("ntracks_2D", C.c_long), ("z_stacked", C.c_int), ("ntracks", C.c_long), ("nthreads", C.c_int),
            ("papi_event_set", C.c_int), ("n_2D_source_regions_per_assembly", C.c_long),
            ("n_source_regions_per_node", C.c_long), ("load_tracks", C.c_bool), ("track_file", C.c_char_p),
            ("segments_processed", C.c_long)]
    return Input


WORKLOAD_LABELS = {
    "default": "default strawman problem, built-in set_default_input (src/init.c:33-74): G=104, 120 2D segments/track",
    "default_in": "default.in as shipped (src/default.in): G=100, cai=9, 20 2D segments/track",
    "small": "small problem (-s, src/init.c:77-103)",
}


class ReferenceCPU:
    """oracle/_ref/libsimplemoc_ref_omp.so: the unmodified reference, stock OpenMP flags."""

    def __init__(self):
        path = os.path.join(ROOT, "oracle", "_ref", "libsimplemoc_ref_omp.so")
        self.kind = "reference"
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = C.CDLL(path, mode=C.RTLD_LOCAL)
        L.ref_case_create.restype = C.c_void_p
        L.ref_case_create.argtypes = [C.c_char_p, C.c_int, C.c_uint64, C.c_int, C.c_long]
        L.ref_case_destroy.argtypes = [C.c_void_p]
        L.ref_time_transport_sweep.restype = C.c_double
        L.ref_time_transport_sweep.argtypes = [C.c_void_p]
        L.ref_input.restype = C.c_void_p
        L.ref_input.argtypes = [C.c_void_p]
        L.ref_psi.restype = C.c_void_p
        L.ref_psi.argtypes = [C.c_void_p]
        L.ref_source_data.restype = C.c_void_p
        L.ref_source_data.argtypes = [C.c_void_p]
        self.L = L

    def make(self, workload, nthreads, limit_tracks_2d):
        Input = reference_input_struct()     # the reference arm maps oracle/_ref only, never libmoc_b200.so
        path = b""
        tmp = None
        if workload == "default_in":
            fd, tmp = tempfile.mkstemp(prefix="moc_ref_", suffix=".in")
            with os.fdopen(fd, "w") as f:
                for v in DEFAULT_IN:
                    f.write(f"{v}\n")
            path = tmp.encode()
        h = self.L.ref_case_create(path, int(workload == "small"), 1, nthreads, limit_tracks_2d)
        if tmp:
            os.unlink(tmp)
        inp = C.cast(self.L.ref_input(h), C.POINTER(Input)).contents
        return h, inp

    def reset(self, h, inp):
        """zero the angular and scalar flux (what a fresh run starts from), outside any timed region"""
        T3, G, F, N = inp.ntracks, inp.n_egroups, inp.fai, inp.n_source_regions_per_node
        C.memset(self.L.ref_psi(h), 0, 4 * 2 * T3 * G)
        C.memset(self.L.ref_source_data(h) + 4 * N * F * G, 0, 4 * N * F * G)

    def sweep_seconds(self, h):
        return self.L.ref_time_transport_sweep(h)

    def destroy(self, h):
        self.L.ref_case_destroy(h)


def time_reference(args, steps, warmup, seconds_per_step, full=False):
    """Times transport_sweep of the reference on a sample of `workload` sized for
    ~seconds_per_step (full: on the whole workload); returns (integrations/s, s/step, description dict)."""
    ref = ReferenceCPU()
    cores = os.cpu_count() or 1
    # calibrate on a sliver
    probe_tracks = 16
    h, inp = ref.make(args.workload, cores, probe_tracks)
    t = ref.sweep_seconds(h)
    integ = inp.segments_processed * inp.n_egroups
    ref.destroy(h)
    rate = integ / max(t, 1e-6)
    per_2d_track = integ / probe_tracks
    want = int(max(16, seconds_per_step * rate / per_2d_track)) // 2 * 2
    if full:
        want = 0          # ref_case_create: 0 = every 2D track of the workload
    h, inp = ref.make(args.workload, cores, want)
    want = inp.ntracks_2D   # clipped to the full problem if the sample would exceed it
    times, integs = [], []
    for s in range(warmup + steps):
        ref.reset(h, inp)
        t = ref.sweep_seconds(h)
        if s >= warmup:
            times.append(t)
            integs.append(inp.segments_processed * inp.n_egroups)
    ref.destroy(h)
    total_t, total_i = sum(times), sum(integs)
    desc = {"kind": ref.kind, "cores": cores,
            "sample": f"transport_sweep of the unmodified reference (OpenMP, -Ofast) over "
                      f"{want} of the workload's 2D tracks ({integs[0]:.3e} integrations/step, "
                      f"{total_t / len(times):.2f} s/step, {len(times)} steps after {warmup} warm-up)",
            "value": total_i / total_t, "unit": UNIT,
            "ns_per_integration": 1e9 * total_t / total_i}
    return total_i / total_t, total_t / len(times), desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the whole run should end within a few minutes
    per_step = max(2.0, min(20.0, 150.0 / (args.steps + args.warmup)))
    value, sec_per_step, desc = time_reference(args, args.steps, args.warmup, per_step, full=args.full_size)
    label = WORKLOAD_LABELS[args.workload]          # the reference arm runs the named workload as shipped
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": label, "sampled": not args.full_size}, "impl": "reference",
            "ns_per_integration": 1e9 / value, "cpu_baseline": desc,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_line(json.dumps(finite(line), allow_nan=False))



# ------------------------------------------------------------------ roofline of K1

# FMA-pipe cycles one segment x one lane costs in the G = 104 instantiation of attenuate_kernel (13 groups per lane),
# counted in the SASS of the built library (tools/sass_histogram: packed FFMA2/FMUL2/FADD2 hold the pipe for two
# cycles per warp instruction, scalar FFMA/FMUL/FADD/IMAD for one); ncu's sm__pipe_fma_cycles_active of the same
# launch is the cross-check (profiles/).  Keep in step with moc_attenuate.cuh.
K1_PIPE_MIX = {"table": (210, 60), "sfu": (192, 45)}     # (packed, scalar) per segment and lane
# ... and the instructions it issues per segment and lane (loop body of the staged kernel, table mode 437 with the
# per-segment skip of the x > maxVal test, SFU mode 385); a packed FP32x2 instruction holds the issue port of its
# sub-partition for a second cycle (tools/ubench/issue_slots.cu: 2.04 cycles): issue cycles = instructions + packed.
# The arithmetic alone, operands already in shared memory, runs AT this limit (profiles/r02_K1_findings.md).
K1_ISSUED = {"table": 437, "sfu": 385}


def k1_roofline(args, api, dev_opts, inp, state, step_ms_total, clocks, l2_probe, torch, local, world):
    """attenuate_kernel against what bounds it.  L2-resident source slab (configs 1-4): the FP32 (FMA) pipe --
    `frac` is the fraction of FMA-pipe cycles the launch keeps busy; HBM sits beside it (`hbm`), with the DRAM
    traffic ncu measured against SURVEY 8(d)'s algorithmic bytes (angular flux in + out).  Source slab larger
    than the L2 (config 5): random 128-byte DRAM gathers -- `bound` is "hbm" and `frac` the HBM fraction."""
    G, T3 = inp.n_egroups, inp.ntracks
    n_launch = max(args.steps, 1)
    att_s = state["att_ms"] * 1e-3
    att_per_launch = att_s / n_launch
    my_integ = state["segments"] * G
    integ_per_launch = my_integ / n_launch
    hbm_peak, peak_src = measured_peaks()
    # algorithmic HBM bytes = the angular flux of every track read and written once
    alg_bytes = 8.0 * T3 * G
    record_bytes = 12.0 * state["segments"] / n_launch       # K0 -> K1 segment records (this design's own stream)
    # DRAM bytes of one launch from the committed ncu capture of this workload (profiles/), if there is one
    traffic, traffic_src = None, None
    if not (args.egroups or args.limit_tracks_2d) and world == 1:
        want = "config5" if args.decomp_ax == 2 else args.workload
        for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_K1_dram_traffic*.json")), reverse=True):
            try:
                with open(f) as fh:
                    t = json.load(fh)
                if t.get("workload") == want and t.get("exp") == args.exp:
                    traffic, traffic_src = t["traffic"], os.path.relpath(f, ROOT)
                    if t.get("scale_to_full"):      # captured on a slice of 2D tracks (ncu replays ~40x): per-track figure
                        traffic *= T3 / float(t["ntracks"])
                    break
            except (OSError, ValueError, KeyError):
                pass
    Gp = (G + 31) // 32 * 32
    gather_set = 4.0 * (2 * inp.fai + 1) * inp.n_source_regions_per_node * Gp
    l2_size = float(torch.cuda.get_device_properties(local).L2_cache_size)
    slab_in_dram = inp.axial_exp == 2 and dev_opts["fit_per_segment"] == 1 and gather_set > l2_size
    hbm = {"algorithmic_bytes": alg_bytes, "segment_record_bytes": record_bytes,
           "achieved_gbs": alg_bytes / att_per_launch / 1e9 if att_per_launch > 0 else None,
           "peak_gbs": hbm_peak, "peak_source": peak_src}
    hbm["frac"] = hbm["achieved_gbs"] / hbm_peak if hbm["achieved_gbs"] else None
    if traffic:
        hbm["traffic_over_algorithmic"] = traffic / alg_bytes
        hbm["traffic_gbs"] = traffic / att_per_launch / 1e9 if att_per_launch > 0 else None
        hbm["traffic_frac_of_peak"] = hbm["traffic_gbs"] / hbm_peak if hbm["traffic_gbs"] else None
    fp32 = {"flop_per_integration": FLOP_PER_INTEGRATION,
            "algorithmic_tflops": my_integ * FLOP_PER_INTEGRATION / att_s / 1e12 if att_s else None,
            "peak_tflops_nominal": FP32_PEAK_TFLOPS_NOMINAL}
    fp32["algorithmic_frac_of_nominal"] = fp32["algorithmic_tflops"] / FP32_PEAK_TFLOPS_NOMINAL if att_s else None
    roof = {"kernel": "attenuate_kernel", "ms_per_launch": 1e3 * att_per_launch,
            "share_of_step": state["att_ms"] / step_ms_total if step_ms_total else None,
            "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": alg_bytes, "hbm": hbm, "fp32": fp32,
            "l2": {"achieved_gbs": my_integ * L2_BYTES_PER_INTEGRATION / att_s / 1e9 if att_s else None,
                   "bytes_per_integration": L2_BYTES_PER_INTEGRATION}}
    # the measured ceiling of the kernel's memory side: the same gathers + vector reductions on the
    # same (L2-resident) slab without the arithmetic, timed live (moc_probe_l2_gather)
    if isinstance(l2_probe, tuple):
        probe_rd, probe_mix = l2_probe
        iface = my_integ * 20.0 / att_s / 1e9 if att_s else None   # 16 B gathered + 4 B reduced per integration
        roof["l2"].update({"sm_l2_interface_bytes_per_integration": 20, "achieved_interface_gbs": iface,
                           "probe_gather_gbs": probe_rd / 1e9, "probe_gather_plus_red_gbs": probe_mix / 1e9,
                           "frac_of_probe": iface / (probe_mix / 1e9) if iface else None,
                           "probe": "moc_probe_l2_gather: K1's access pattern on the same slab, no arithmetic"})
    elif l2_probe is not None:
        roof["l2"]["probe_error"] = l2_probe
    if slab_in_dram:
        # every per-integration gather (16 B read + 8 B read-modify-write) is a DRAM access, except for
        # the share of the working set the L2 holds.  With a committed ncu capture `achieved` is measured DRAM
        # traffic / launch time; without one it is the modelled figure, and says so.
        miss = 1.0 - l2_size / gather_set
        model = alg_bytes + record_bytes + L2_BYTES_PER_INTEGRATION * miss * integ_per_launch
        used = traffic if traffic else model
        roof.update({"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_source": peak_src,
                     "achieved": used / att_per_launch / 1e9 if att_per_launch > 0 else None,
                     "achieved_from": ("ncu dram__bytes of the same launch (" + traffic_src + ")") if traffic else
                                      f"modelled: flux + records + {L2_BYTES_PER_INTEGRATION} B/integration x L2 miss share {miss:.2f}",
                     "modelled_bytes": model,
                     "note": f"gather working set {gather_set / 1e6:.0f} MB > L2 {l2_size / 1e6:.0f} MB: random 128-byte DRAM gathers"})
    else:
        # FMA-pipe occupancy: pipe cycles the launch needs (static instruction mix x integrations) against
        # 128 FP32 lanes per SM at the SM clock sampled during the timed region
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        packed, scalar = K1_PIPE_MIX[args.exp]
        cycles = (2 * packed + scalar) / 13.0
        exact = G == 104 and inp.axial_exp == 2
        peak = n_sm * 128 * sm_mhz * 1e6 * 2 / 1e12          # TFLOP/s the FMA pipes can issue at this clock
        achieved = my_integ * cycles * 2 / att_s / 1e12 if att_s else None
        issue_cycles = (K1_ISSUED[args.exp] + packed) / 13.0
        roof["issue"] = {"issue_cycles_per_integration": issue_cycles,
                         "frac_of_issue_slots": my_integ * issue_cycles / att_s / (n_sm * 4 * 32 * sm_mhz * 1e6) if att_s else None,
                         "note": "instructions issued + one extra cycle per packed FP32x2 instruction, against 4 sub-partitions x "
                                 "1 warp-instruction (32 lanes) per cycle: what the loop is actually bound by"}
        roof.update({"bound": "fp32", "unit": "TFLOP/s", "achieved": achieved, "peak": peak,
                     "peak_source": f"{n_sm} SMs x 128 FP32 lanes x 2 FLOP x {sm_mhz:.0f} MHz (SM clock sampled during the run)",
                     "fma_pipe_lane_cycles_per_integration": cycles,
                     "achieved_is": "FMA-pipe lane-cycles the kernel occupies x 2 (packed FP32x2 instructions hold the pipe two "
                                    "cycles): frac = sm__pipe_fma_cycles_active" + ("" if exact else
                                    " (instruction mix of the G = 104 instantiation applied to this group count: approximate)"),
                     "note": "not HBM-bound: the gathered working set is L2-resident (hbm.frac is tiny by design); the binding "
                             "pipe is FP32/FMA, then L1TEX/L2 gathers (l2.frac_of_probe)"})
    roof["frac"] = roof["achieved"] / roof["peak"] if roof.get("achieved") else None
    return roof


# ------------------------------------------------------------------ the CUDA path

def run_moc(args):
    import torch
    import simplemoc_b200 as m
    from simplemoc_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs `python -m torch.distributed.run "
                             f"--nproc-per-node {args.gpus} bench.py ...` (one process per GPU)")
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {args.gpus}")
    if not torch.cuda.is_available() or api.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the MOC path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local) if (world > 1 and not args.no_numa_bind) else "not requested"
    dist = None
    if world > 1:
        # the exchange runs UNDER the interior sweep: NCCL's copy kernels need a few CTAs, not half the machine
        # (3.2 GB per GPU at 2x2x2 against a 390 ms sweep); the user's own settings win
        os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_max_ctas))
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()

    inp, label = workload_input(m, args.workload, args.egroups, args.decomp_ax)
    inp.mype = rank
    inp = m.derive(inp, args.limit_tracks_2d)
    t0 = time.time()
    exp_mode = api.EXP_SFU if args.exp == "sfu" else api.EXP_TABLE_REF
    if args.device_build:
        args.no_e2e = True
        host = None
        dev = m.DeviceProblem.synthetic(inp, seed=1 + rank, device=local, exp_mode=exp_mode)
    else:
        host = m.HostProblem(inp, seed=1 + rank)          # every rank: a full-size domain of its own
        dev = m.DeviceProblem(host, device=local, exp_mode=exp_mode)
    build_s = time.time() - t0
    if not args.staged:
        dev.set_option(api.OPT_STAGED, 0)
    if not args.noclamp:
        dev.set_option(api.OPT_NOCLAMP, 0)
    if args.fill_overlap:
        dev.set_option(api.OPT_FILL_OVERLAP, args.fill_overlap)
    if args.fill_batches:
        dev.set_option(api.OPT_FILL_BATCHES, args.fill_batches)
    cx, cy, cz = grid_for(world, args.grid)
    grid = m.make_grid(cx, cy, cz, rank)
    if world > 1:
        # the 128-byte NCCL id travels over torch.distributed; the exchange itself is the library's
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(api.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        dev.comm_init(world, rank, bytes(buf.cpu().numpy().tobytes()))

    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local))
    G = inp.n_egroups
    state = {"keff": 1.0, "segments": 0, "att_ms": 0.0, "fill_ms": 0.0, "count_ms": 0.0, "scan_ms": 0.0, "sweep_ms": 0.0}

    def step(accumulate):
        # N > 1: the exchange runs under the sweep of the interior z-stacks (moc_sweep_exchange)
        n = dev.sweep_exchange(grid) if world > 1 else dev.sweep()
        dev.renormalize()
        # main.c:81,89 feeds each iteration's k-eff to the next update_sources.  With neighbours the
        # reference adds UN-normalised boundary flux sums to the leakage (comms.c:120) and never
        # resets it, so k = fission/(absorption+leakage) collapses after one iteration (it runs
        # exactly one, main.c:41); keep the source scale finite so later steps time real numbers
        k_in = state["keff"]
        dev.update_sources(k_in if (k_in == k_in and 1e-2 < abs(k_in) < 1e2) else 1.0)
        state["keff"] = dev.compute_keff()
        if accumulate:
            t = dev.timing()
            state["segments"] += n
            state["att_ms"] += t.attenuate_ms
            state["fill_ms"] += t.fill_ms
            state["count_ms"] += t.count_ms
            state["scan_ms"] += t.scan_ms
            state["sweep_ms"] += t.total_ms
        return n

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local)
    barrier(); torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    launches0 = dev.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step(True)
    e1.record(stream)
    torch.cuda.synchronize(); barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = dev.launch_count - launches0
    segs = state["segments"]
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        s = torch.tensor([segs, launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        segs, launches = int(s[0].item()), int(s[1].item())
    integrations = segs * G
    value = integrations / (ms * 1e-3)
    # N > 1: every rank's own sweep and attenuation time.  The step synchronises the ranks (exchange, two all-reduces),
    # so the step time is the SLOWEST GPU's sweep plus the reductions: the spread says how much of the scaling loss is
    # board-to-board variation under the power cap rather than communication.
    per_rank = None
    if dist is not None:
        n_l = max(args.steps, 1)
        mine = torch.tensor([state["sweep_ms"] / n_l, state["att_ms"] / n_l], dtype=torch.float64, device="cuda")
        everyone = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(everyone, mine)
        per_rank = {"sweep_ms": [round(float(x[0].item()), 1) for x in everyone],
                    "attenuate_ms": [round(float(x[1].item()), 1) for x in everyone]}

    leakage = dev.leakage
    # ---- the same step with the other exponential (north_star: MUFU.EX2 instead of the reference's table),
    # two timed steps; informational: the headline `value` is the mode --exp names (default: parity mode)
    other = None
    if world == 1 and not args.limit_tracks_2d:
        try:
            dev.set_option(api.OPT_EXP_MODE, api.EXP_TABLE_REF if args.exp == "sfu" else api.EXP_SFU)
            saved = dict(state)
            step(False)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            f0.record(stream)
            n_other = step(False) + step(False)
            f1.record(stream)
            torch.cuda.synchronize()
            o_ms = f0.elapsed_time(f1)
            other = {"exp": "table" if args.exp == "sfu" else "sfu", "steps": 2, "ms_per_step": o_ms / 2,
                     "value": n_other * G / (o_ms * 1e-3), "unit": UNIT,
                     "attenuate_ms": dev.timing().attenuate_ms}
            state.update(saved)
        finally:
            dev.set_option(api.OPT_EXP_MODE, exp_mode)
    # ---- the measured ceiling of K1's memory side (needs the resident handle: before the e2e leg)
    l2_probe = None
    try:
        if inp.fai >= 3 and inp.axial_exp == 2:
            l2_probe = (dev.probe_l2_gather(0), dev.probe_l2_gather(1))
    except Exception as e:   # diagnostics only
        l2_probe = str(e)

    dev_opts = {"fit_per_segment": dev.get_option(api.OPT_FIT_PER_SEGMENT)}
    T3 = inp.ntracks
    n_launch = max(args.steps, 1)

    # ---- N > 1: is what NCCL moved what comms.c moves?  The boundary-exchange case through moc_exchange on this
    # communicator's ranks against the all-ranks model of comms.c (pinned on the reference's own comms.c,
    # tests/test_exchange_pin.py), before anything else is torn down
    exchange_parity = None
    if world > 1:
        exchange_parity = check_exchange_parity(m, api, torch, dist, world, rank, local, (cx, cy, cz))

    # ---- end to end through the reference's own entry point on host structures (rank-local)
    e2e = None
    full_loop = None
    if not args.no_e2e:
        e2e, full_loop = measure_e2e(args, m, api, host, dev, torch, dist, world, rank, local, grid)

    # ---- roofline of the dominant kernel (attenuate_kernel), per launch
    roof = k1_roofline(args, api, dev_opts, inp, state, ms, clocks, l2_probe, torch, local, world)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            _, _, cpu = time_reference(args, 1, 0, args.cpu_seconds)
        except FileNotFoundError as e:
            cpu = {"unavailable": f"{e} (oracle/_ref did not travel)"}

    if dist is not None:
        dist.barrier()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": label, "domains": f"{cx}x{cy}x{cz}", "exp": args.exp,
                           "step": "transport_sweep" + (" + boundary exchange (overlapped)" if world > 1 else "") +
                                   " + renormalize_flux + update_sources + compute_keff",
                           "ntracks_per_gpu": T3, "n_egroups": G,
                           "segments_per_sweep_per_gpu": state["segments"] // n_launch,
                           "l2": f"inputs larger than L2 ({8e-9 * T3 * G:.1f} GB angular flux + "
                                 f"{12e-9 * state['segments'] / n_launch:.1f} GB segment records streamed per "
                                 "step, 126 MB L2); no flush",
                           "build_s": round(build_s, 1), "built_on": "device" if args.device_build else "host",
                           "numa": numa, "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS") if world > 1 else None},
                "ns_per_integration": 1e9 / value, "keff": state["keff"], "leakage": leakage,
                "sweep_ms": state["sweep_ms"] / n_launch,
                "phases_ms": {"count": state["count_ms"] / n_launch, "scan": state["scan_ms"] / n_launch,
                              "fill": state["fill_ms"] / n_launch, "attenuate": state["att_ms"] / n_launch},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "e2e": e2e,
                "e2e_full_loop": full_loop, "other_exp_mode": other,
                "cpu_baseline": cpu}
        if exchange_parity is not None:
            line["exchange_parity"] = exchange_parity
        if per_rank is not None:
            line["per_rank_ms"] = per_rank
        line["config"]["keff_feedback"] = ("k-eff of step n feeds update_sources of step n+1 (main.c:81,89); replaced by 1.0 when it "
                                          "leaves [1e-2, 1e2] (with neighbours the reference adds un-normalised flux sums to the "
                                          "leakage and never resets it, comms.c:120: k collapses after its single iteration)")
        if args.limit_tracks_2d:
            line["config"]["limit_tracks_2d"] = args.limit_tracks_2d
        emit_line(json.dumps(finite(line), allow_nan=False))
    dev.close()
    if host is not None:
        host.close()
    if dist is not None:
        dist.destroy_process_group()


def measure_e2e(args, m, api, host, dev, torch, dist, world, rank, local, grid):
    """transport_sweep(Params*, Input*) -- the reference's own prototype (src/solver.c:283) --
    on the host structures, non-resident: upload of the step's inputs from pinned host memory,
    the sweep, download of what transport_sweep mutates; all inside the timed region.
    Returns (e2e, e2e_full_loop): the second times the reference's whole iteration (main.c:57-92:
    transport_sweep, [fast_transfer_boundary_fluxes,] renormalize_flux, update_sources, compute_keff) through
    the five drop-in names on the same host structures, every call moving what it reads and writes."""
    L = api.lib()
    inp = host.I
    T3, G, F, N = inp.ntracks, inp.n_egroups, inp.fai, inp.n_source_regions_per_node
    # release the resident copy first: the drop-in path builds its own mirror of this Params
    dev.close()
    L.moc_set_resident(0)
    L.moc_dropin_configure(host.seed, host.rand_calls, 1 if args.exp == "sfu" else 0, 48)
    steps = max(1, args.e2e_steps)
    I2 = type(inp).from_buffer_copy(inp)
    L.transport_sweep(C.byref(host.P), C.byref(I2))           # warm-up: builds the mirror
    torch.cuda.synchronize()
    mirror = L.moc_handle_of(C.byref(host.P))
    if world > 1:
        # the drop-in mirror is a handle of its own: it needs its own communicator for the exchange
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(api.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        assert L.moc_comm_init(mirror, world, rank, bytes(buf.cpu().numpy().tobytes())) == 0
    stream = torch.cuda.ExternalStream(L.moc_get_stream(mirror), device=torch.device("cuda", local))

    def timed(body):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        segs = 0
        for _ in range(steps):
            segs += body()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dt = e0.elapsed_time(e1) * 1e-3
        if dist is not None:
            t = torch.tensor([dt, wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, wall = float(t[0].item()), float(t[1].item())
            s = torch.tensor([segs], dtype=torch.int64, device="cuda")
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
            segs = int(s.item())
        return segs, dt, wall

    def sweep_only():
        L.transport_sweep(C.byref(host.P), C.byref(I2))       # returns after the download completed
        return I2.segments_processed

    segs, dt, wall = timed(sweep_only)
    h2d = 40 * T3 + 4 * T3 * G + 4 * (2 * F + 1) * N * G     # Track image, forward flux rows, source slab
    d2h = 40 * T3 + 4 * T3 * G + 4 * F * N * G               # Track image, forward flux rows, scalar flux
    e2e = {"value": segs * G / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": steps, "ms_per_step": 1e3 * dt / steps, "host_wall_ms_per_step": 1e3 * wall / steps,
           "copy_gbs_per_rank_each_way": h2d * steps / dt / 1e9,
           "call": "transport_sweep(Params*, Input*) on host structures (drop-in C-ABI), CUDA events on "
                   "the library's stream around upload + sweep + download; times the sweep only, like the "
                   "metric (utils.c:147-155) and the reference arm -- the whole iteration is e2e_full_loop"}
    # the same call when the caller promises not to modify its structures between calls (moc_dropin_trust_device):
    # uploads are skipped, every result is still written back.  NOT the headline e2e (its inputs do not travel).
    L.moc_dropin_trust_device(1)
    segs, dt, wall = timed(sweep_only)
    L.moc_dropin_trust_device(0)
    e2e["trust_device"] = {"value": segs * G / dt, "ms_per_step": 1e3 * dt / steps, "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": d2h,
                           "what": "moc_dropin_trust_device(1): the caller only reads its structures between calls; uploads "
                                   "skipped, downloads kept (informational: the headline e2e is the strict mode above)"}
    full = None
    if not args.no_full_loop:
        keff = [1.0]

        def iteration():
            L.transport_sweep(C.byref(host.P), C.byref(I2))
            if world > 1:
                L.fast_transfer_boundary_fluxes(host.P, I2, grid)
            L.renormalize_flux(host.P, I2, grid)
            k = keff[0]
            L.update_sources(host.P, I2, k if (k == k and 1e-2 < abs(k) < 1e2) else 1.0)
            keff[0] = L.compute_keff(host.P, I2, grid)
            return I2.segments_processed

        iteration()                                            # warm-up of the four other entry points
        segs, dt, wall = timed(iteration)
        full = {"value": segs * G / dt, "unit": UNIT, "steps": steps, "ms_per_step": 1e3 * dt / steps,
                "host_wall_ms_per_step": 1e3 * wall / steps,
                "call": "the reference's iteration (main.c:57-92) through the five drop-in names on host structures, "
                        "host authoritative between the calls (moc_set_resident(0), the default)"}
    L.moc_release(C.byref(host.P))
    return e2e, full


def check_exchange_parity(m, api, torch, dist, world, rank, local, dims):
    """The boundary exchange of this run's ranks against the CPU model of comms.c (oracle_exchange, pinned bit for
    bit on the reference's own comms.c by tests/test_exchange_pin.py): a small problem with one 10 000-track
    message per face, every rank's own domain swept on its GPU, moc_exchange over NCCL, slab and leakage compared
    on every rank.  The oracle is the checker here, outside every timed region."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from oracle_lib import CASES, CommGrid, OracleCase, make_grid as oracle_grid
    vals = CASES["exch"]
    host = m.HostProblem(m.derive(m.input_from_values(vals)), seed=21 + rank)
    dev = m.DeviceProblem(host, device=local)
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(api.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    dev.comm_init(world, rank, bytes(buf.cpu().numpy().tobytes()))
    cases = [OracleCase(vals, seed=21 + r) for r in range(world)]
    for c in cases:
        c.sweep()
    ok = dev.sweep() == cases[rank].I.segments_processed
    dev.set(api.ARR_PSI, cases[rank].psi)               # identical slabs in, so the comparison is bit for bit
    grids = (CommGrid * world)(*[oracle_grid(*dims, r) for r in range(world)])
    hs = (C.c_void_p * world)(*[c.h for c in cases])
    ok = ok and OracleCase.lib().oracle_exchange(hs, grids, world) == 0
    dev.exchange(m.make_grid(*dims, rank))
    ok = ok and bool(np.array_equal(dev.get(api.ARR_PSI), cases[rank].psi))
    ok = ok and dev.leakage == float(cases[rank].leakage[0])
    moved = int((cases[rank].psi != 0).sum())
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dev.close(); host.close()
    for c in cases:
        c.close()
    verdict = "bit-exact" if int(flag.item()) == 1 else "MISMATCH"
    return {"result": verdict, "ranks": world, "grid": "x".join(str(d) for d in dims),
            "checked": "flux slab after moc_exchange and leakage on every rank == oracle_exchange (comms.c:5-196 model, "
                       "pinned on the reference's own comms.c under an in-process MPI)", "case": "exch (96 000 tracks, G = 8)"}


_RESULT_FD = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result.  Native libraries write there too (NCCL prints
    "NCCL version ..." to stdout when NCCL_DEBUG is set in the environment): keep a private duplicate of the
    real stdout for the result and point fd 1 at stderr for everybody else."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit_line(text):
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (text + "\n").encode())


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_moc(args)


if __name__ == "__main__":
    main()
