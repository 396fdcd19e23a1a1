"""ctypes binding of libmoc_b200.so (include/moc_b200.h).

This is the host-side mirror of the reference's C interface for Python callers
(tests, bench.py): same structures (reference src/SimpleMOC_header.h:28-158), same
function names and argument meaning.  It holds no arithmetic: every compute call goes
to the CUDA library and raises MocError if the library or a GPU is missing -- there is
no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# MOC_B200_LIB: an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("MOC_B200_LIB") or os.path.join(HERE, "libmoc_b200.so")


class MocError(RuntimeError):
    pass


# ---------------------------------------------------------------- structures

class Input(C.Structure):
    """src/SimpleMOC_header.h:28-76"""
    _fields_ = [
        ("x_assemblies", C.c_int), ("y_assemblies", C.c_int), ("cai", C.c_int),
        ("fai", C.c_int), ("axial_exp", C.c_int), ("radial_ray_sep", C.c_float),
        ("axial_z_sep", C.c_float), ("n_azimuthal", C.c_int),
        ("n_polar_angles", C.c_int), ("n_egroups", C.c_int), ("decompose", C.c_bool),
        ("decomp_assemblies_ax", C.c_int), ("segments_per_track", C.c_long),
        ("assembly_width", C.c_float), ("height", C.c_float),
        ("domain_height", C.c_float), ("precision", C.c_float), ("mype", C.c_long),
        ("ntracks_2D", C.c_long), ("z_stacked", C.c_int), ("ntracks", C.c_long),
        ("nthreads", C.c_int), ("papi_event_set", C.c_int),
        ("n_2D_source_regions_per_assembly", C.c_long),
        ("n_source_regions_per_node", C.c_long), ("load_tracks", C.c_bool),
        ("track_file", C.c_char_p), ("segments_processed", C.c_long),
    ]


class Table(C.Structure):
    """src/SimpleMOC_header.h:123-128"""
    _fields_ = [("values", C.POINTER(C.c_float)), ("dx", C.c_float),
                ("maxVal", C.c_float), ("N", C.c_int)]


class Params(C.Structure):
    """src/SimpleMOC_header.h:131-138 (pointer members kept opaque)"""
    _fields_ = [("tracks_2D", C.c_void_p), ("tracks", C.c_void_p), ("sources", C.c_void_p),
                ("polar_angles", C.POINTER(C.c_float)), ("leakage", C.POINTER(C.c_float)),
                ("expTable", Table)]


class CommGrid(C.Structure):
    """src/SimpleMOC_header.h:141-158 without the MPI members"""
    _fields_ = [(n, C.c_int) for n in (
        "x_pos_src", "x_pos_dest", "x_neg_src", "x_neg_dest",
        "y_pos_src", "y_pos_dest", "y_neg_src", "y_neg_dest",
        "z_pos_src", "z_pos_dest", "z_neg_src", "z_neg_dest")]


class ExchangeOp(C.Structure):
    """moc_exchange_op: one chunk of the boundary exchange schedule (src/comms.c:100-183)"""
    _fields_ = [("offset", C.c_longlong), ("count", C.c_longlong), ("round", C.c_int),
                ("direction", C.c_int), ("send_to", C.c_int), ("recv_from", C.c_int)]


class SweepTiming(C.Structure):
    _fields_ = [("count_ms", C.c_float), ("scan_ms", C.c_float), ("fill_ms", C.c_float),
                ("attenuate_ms", C.c_float), ("total_ms", C.c_float),
                ("n_batches", C.c_long), ("launches", C.c_long)]


# option / array ids (include/moc_b200.h)
OPT_EXP_MODE, OPT_SEED, OPT_RAND_BASE, OPT_BATCH_SEGMENTS, OPT_SOURCE_STRIDE, OPT_LANES = 1, 2, 3, 4, 5, 6
OPT_STREAM_CHUNKS = 7
OPT_WALK_KERNEL = 8
OPT_FILL_OVERLAP, OPT_FILL_BATCHES = 9, 10
OPT_DIGEST = 100
OPT_EXACT_RAY_TRACE = 102
OPT_FIT_PER_SEGMENT = 103   # diagnostic: quadratic fit per segment (the path of source slabs larger than the L2)
OPT_NOCLAMP = 105           # set 0: keep the x > maxVal test of the table in every launch; get: 1 if the last sweep dropped it
OPT_STAGED = 104            # 1 (default): TMA-staged attenuation kernel where it applies; 0: direct gathers (A/B timing)
EXP_TABLE_REF, EXP_SFU = 0, 1
ARR_FINE_SOURCE, ARR_FINE_FLUX, ARR_SIGT, ARR_PSI, ARR_Z_HEIGHT, ARR_P_WEIGHT, ARR_SEG_COUNT, ARR_QSR_DIGEST = \
    1, 2, 3, 4, 5, 6, 7, 8
ARR_QSR_DIGEST_BACK = 9     # the backward pass of two_way_sweep
(HOST_AZ_WEIGHT, HOST_N_SEGMENTS, HOST_SEG_LENGTHS, HOST_XS, HOST_SCATTER, HOST_XS_INDEX, HOST_VOL,
 HOST_POLAR, HOST_TABLE) = range(20, 29)
_HOST_DTYPE = {HOST_N_SEGMENTS: np.int64, HOST_XS_INDEX: np.int32, ARR_SEG_COUNT: np.uint32}

# every symbol include/moc_b200.h declares (checked by tests/test_abi.py)
EXPORTED = [
    "transport_sweep", "two_way_transport_sweep", "renormalize_flux", "update_sources", "compute_keff",
    "fast_transfer_boundary_fluxes", "moc_dropin_configure", "moc_set_device", "moc_handle_of",
    "moc_set_resident", "moc_dropin_trust_device", "moc_dropin_set_grid", "moc_sync_to_host", "moc_release",
    "moc_create", "moc_create_synthetic", "moc_destroy", "moc_set_option", "moc_get_option", "moc_sweep", "moc_two_way_sweep",
    "moc_renormalize", "moc_update_sources", "moc_compute_keff", "moc_exchange", "moc_sweep_exchange",
    "moc_get_sweep_timing", "moc_get_array", "moc_set_array", "moc_download", "moc_upload",
    "moc_get_leakage", "moc_synchronize", "moc_get_stream", "moc_get_launch_count", "moc_probe_l2_gather", "moc_comm_get_unique_id", "moc_comm_init",
    "moc_make_grid", "moc_exchange_plan", "moc_last_error", "moc_device_count", "moc_set_default_input",
    "moc_set_small_input", "moc_read_input_file", "moc_read_CLI",
    "moc_calculate_derived_inputs", "moc_est_mem_usage", "moc_build_tracks", "moc_load_openmoc_tracks",
    "moc_free_tracks", "moc_time_per_intersection", "moc_params_get", "moc_params_set",
]

_lib = None


def lib():
    """Load libmoc_b200.so; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MocError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    vp, ip, lp = C.c_void_p, C.POINTER(Input), C.POINTER(C.c_long)
    L.moc_last_error.restype = C.c_char_p
    L.moc_device_count.restype = C.c_int
    L.moc_set_default_input.restype = Input
    L.moc_set_small_input.argtypes = [ip]
    L.moc_read_input_file.argtypes = [ip, C.c_char_p]
    L.moc_read_CLI.argtypes = [C.c_int, C.POINTER(C.c_char_p), ip]
    L.moc_calculate_derived_inputs.argtypes = [ip]
    L.moc_est_mem_usage.restype = C.c_size_t
    L.moc_est_mem_usage.argtypes = [ip]
    L.moc_build_tracks.argtypes = [ip, C.c_uint64, C.POINTER(Params), C.POINTER(C.c_uint64)]
    L.moc_free_tracks.argtypes = [ip, C.POINTER(Params)]
    L.moc_time_per_intersection.restype = C.c_double
    L.moc_time_per_intersection.argtypes = [ip, C.c_double]
    L.moc_params_get.restype = C.c_long
    L.moc_params_get.argtypes = [ip, C.POINTER(Params), C.c_int, vp, C.c_size_t]
    L.moc_params_set.restype = C.c_long
    L.moc_params_set.argtypes = [ip, C.POINTER(Params), C.c_int, vp, C.c_size_t]
    L.moc_create.argtypes = [ip, C.POINTER(Params), C.c_int, C.POINTER(vp)]
    L.moc_create_synthetic.argtypes = [ip, C.c_uint64, C.c_int, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.moc_destroy.argtypes = [vp]
    L.moc_set_option.argtypes = [vp, C.c_int, C.c_long]
    L.moc_get_option.restype = C.c_long
    L.moc_get_option.argtypes = [vp, C.c_int]
    L.moc_sweep.argtypes = [vp, lp]
    L.moc_two_way_sweep.argtypes = [vp, lp]
    L.moc_renormalize.argtypes = [vp]
    L.moc_update_sources.argtypes = [vp, C.c_float, C.POINTER(C.c_float)]
    L.moc_compute_keff.argtypes = [vp, C.POINTER(C.c_float)]
    L.moc_exchange.argtypes = [vp, C.POINTER(CommGrid)]
    L.moc_sweep_exchange.argtypes = [vp, C.POINTER(CommGrid), lp]
    L.moc_get_sweep_timing.argtypes = [vp, C.POINTER(SweepTiming)]
    L.moc_get_array.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.moc_set_array.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.moc_download.argtypes = [vp, C.POINTER(Params)]
    L.moc_upload.argtypes = [vp, C.POINTER(Params)]
    L.moc_get_leakage.restype = C.c_float
    L.moc_get_leakage.argtypes = [vp]
    L.moc_synchronize.argtypes = [vp]
    L.moc_get_stream.restype = vp
    L.moc_get_stream.argtypes = [vp]
    L.moc_get_launch_count.restype = C.c_long
    L.moc_get_launch_count.argtypes = [vp]
    L.moc_probe_l2_gather.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.moc_comm_get_unique_id.argtypes = [C.c_char_p]
    L.moc_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.moc_make_grid.argtypes = [C.c_int] * 4 + [C.POINTER(CommGrid)]
    L.moc_exchange_plan.restype = C.c_long
    L.moc_exchange_plan.argtypes = [ip, C.POINTER(CommGrid), C.POINTER(ExchangeOp), C.c_long]
    # drop-in names (structures by value where the reference passes them by value)
    L.transport_sweep.restype = None
    L.transport_sweep.argtypes = [C.POINTER(Params), ip]
    L.two_way_transport_sweep.restype = None
    L.two_way_transport_sweep.argtypes = [C.POINTER(Params), ip]
    L.renormalize_flux.restype = None
    L.renormalize_flux.argtypes = [Params, Input, CommGrid]
    L.update_sources.restype = C.c_float
    L.update_sources.argtypes = [Params, Input, C.c_float]
    L.compute_keff.restype = C.c_float
    L.compute_keff.argtypes = [Params, Input, CommGrid]
    L.fast_transfer_boundary_fluxes.restype = None
    L.fast_transfer_boundary_fluxes.argtypes = [Params, Input, CommGrid]
    L.moc_set_resident.argtypes = [C.c_int]
    L.moc_dropin_trust_device.argtypes = [C.c_int]
    L.moc_set_device.argtypes = [C.c_int]
    L.moc_sync_to_host.argtypes = [C.POINTER(Params)]
    L.moc_dropin_set_grid.argtypes = [C.POINTER(CommGrid)]
    L.moc_release.argtypes = [C.POINTER(Params)]
    L.moc_dropin_configure.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int]
    L.moc_handle_of.restype = vp
    L.moc_handle_of.argtypes = [C.POINTER(Params)]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise MocError(f"{what} failed ({rc}): {lib().moc_last_error().decode()}")


def device_count():
    return lib().moc_device_count()


# ---------------------------------------------------------------- configuration

# positional values of an input file (src/io.c:210-267)
INPUT_FILE_FIELDS = [
    "x_assemblies", "y_assemblies", "cai", "fai", "axial_exp", "radial_ray_sep",
    "axial_z_sep", "n_azimuthal", "n_polar_angles", "n_egroups", "decompose",
    "decomp_assemblies_ax", "segments_per_track", "assembly_width", "height",
    "precision", "n_2D_source_regions_per_assembly", "papi_event_set"]


def default_input():
    return lib().moc_set_default_input()


def small_input():
    inp = default_input()
    lib().moc_set_small_input(C.byref(inp))
    return inp


def input_from_values(values, track_file=None):
    """track_file: an OpenMOC track file, what `-d <file>` sets (src/io.c:160-168)"""
    inp = default_input()
    for name, v in zip(INPUT_FILE_FIELDS, values):
        setattr(inp, name, bool(v) if name == "decompose" else v)
    if track_file:
        inp.load_tracks = True
        inp.track_file = os.fsencode(track_file)
    return inp


def read_input_file(path, base=None):
    inp = base if base is not None else default_input()
    _check(lib().moc_read_input_file(C.byref(inp), path.encode()), "moc_read_input_file")
    return inp


def derive(inp, limit_tracks_2D=0):
    lib().moc_calculate_derived_inputs(C.byref(inp))
    if limit_tracks_2D and limit_tracks_2D < inp.ntracks_2D:
        inp.ntracks_2D = 2 * (limit_tracks_2D // 2)
        inp.ntracks = inp.ntracks_2D * inp.n_polar_angles * inp.z_stacked
    return inp


def make_grid(cx, cy, cz, rank):
    g = CommGrid()
    _check(lib().moc_make_grid(cx, cy, cz, rank, C.byref(g)), "moc_make_grid")
    return g


# ---------------------------------------------------------------- problems

class HostProblem:
    """The host-side Params of one domain, built like build_tracks() (src/init.c:106-159)."""

    def __init__(self, inp, seed=1):
        self.I = inp
        self.P = Params()
        calls = C.c_uint64(0)
        _check(lib().moc_build_tracks(C.byref(self.I), seed, C.byref(self.P), C.byref(calls)),
               "moc_build_tracks")
        self.seed = seed
        self.rand_calls = calls.value
        self._alive = True

    def get(self, which):
        """flat copy of one array of the host structures (moc_params_get)"""
        n = lib().moc_params_get(C.byref(self.I), C.byref(self.P), which, None, 0)
        if n < 0:
            _check(int(n), "moc_params_get")
        dt = _HOST_DTYPE.get(which, np.float32)
        out = np.empty(n // np.dtype(dt).itemsize, dtype=dt)
        lib().moc_params_get(C.byref(self.I), C.byref(self.P), which, out.ctypes.data, out.nbytes)
        return out

    def set(self, which, arr):
        a = np.ascontiguousarray(arr, dtype=_HOST_DTYPE.get(which, np.float32))
        n = lib().moc_params_set(C.byref(self.I), C.byref(self.P), which, a.ctypes.data, a.nbytes)
        if n != a.nbytes:
            raise MocError(f"moc_params_set({which}): expected {n} bytes, got {a.nbytes}")

    def close(self):
        if self._alive:
            lib().moc_release(C.byref(self.P))
            lib().moc_free_tracks(C.byref(self.I), C.byref(self.P))
            self._alive = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceProblem:
    """A problem resident in HBM (handle API, include/moc_b200.h PART B2)."""

    def __init__(self, host, device=0, exp_mode=EXP_TABLE_REF):
        self.host = host
        self.I = host.I
        self.h = C.c_void_p()
        _check(lib().moc_create(C.byref(host.I), C.byref(host.P), device, C.byref(self.h)), "moc_create")
        self.set_option(OPT_SEED, host.seed)
        self.set_option(OPT_RAND_BASE, host.rand_calls)
        self.set_option(OPT_EXP_MODE, exp_mode)

    @classmethod
    def synthetic(cls, inp, seed=1, device=0, exp_mode=EXP_TABLE_REF):
        """the problem of HostProblem(inp, seed) generated on the device (moc_create_synthetic)"""
        self = cls.__new__(cls)
        self.host = None
        self.I = inp
        self.h = C.c_void_p()
        calls = C.c_uint64(0)
        _check(lib().moc_create_synthetic(C.byref(inp), seed, device, C.byref(self.h), C.byref(calls)),
               "moc_create_synthetic")
        self.seed, self.rand_calls = seed, calls.value
        self.set_option(OPT_EXP_MODE, exp_mode)
        return self

    def close(self):
        if self.h:
            lib().moc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, opt, value):
        _check(lib().moc_set_option(self.h, opt, int(value)), "moc_set_option")

    def get_option(self, opt):
        return lib().moc_get_option(self.h, opt)

    def sweep(self):
        n = C.c_long(0)
        _check(lib().moc_sweep(self.h, C.byref(n)), "moc_sweep")
        return n.value

    def two_way_sweep(self):
        """two_way_transport_sweep (solver.c:556-891): forward walk + backward retrace of every ray"""
        n = C.c_long(0)
        _check(lib().moc_two_way_sweep(self.h, C.byref(n)), "moc_two_way_sweep")
        return n.value

    def renormalize(self):
        _check(lib().moc_renormalize(self.h), "moc_renormalize")

    def update_sources(self, keff):
        r = C.c_float(0)
        _check(lib().moc_update_sources(self.h, keff, C.byref(r)), "moc_update_sources")
        return r.value

    def compute_keff(self):
        k = C.c_float(0)
        _check(lib().moc_compute_keff(self.h, C.byref(k)), "moc_compute_keff")
        return k.value

    def exchange(self, grid):
        _check(lib().moc_exchange(self.h, C.byref(grid)), "moc_exchange")

    def probe_l2_gather(self, mode):
        """bytes/s of the attenuation kernel's gather (+ reduction) pattern alone (diagnostic)"""
        v = C.c_double(0)
        _check(lib().moc_probe_l2_gather(self.h, mode, C.byref(v)), "moc_probe_l2_gather")
        return v.value

    def sweep_exchange(self, grid):
        """sweep + boundary exchange, the exchange overlapped with the interior stacks"""
        n = C.c_long(0)
        _check(lib().moc_sweep_exchange(self.h, C.byref(grid), C.byref(n)), "moc_sweep_exchange")
        return n.value

    def comm_init(self, nranks, rank, unique_id):
        _check(lib().moc_comm_init(self.h, nranks, rank, unique_id), "moc_comm_init")

    def timing(self):
        t = SweepTiming()
        _check(lib().moc_get_sweep_timing(self.h, C.byref(t)), "moc_get_sweep_timing")
        return t

    @property
    def leakage(self):
        return lib().moc_get_leakage(self.h)

    @property
    def stream(self):
        """cudaStream_t (as int) the kernels of this problem are launched on"""
        return lib().moc_get_stream(self.h)

    @property
    def launch_count(self):
        return lib().moc_get_launch_count(self.h)

    def synchronize(self):
        _check(lib().moc_synchronize(self.h), "moc_synchronize")

    def _shape(self, which):
        I = self.I
        T3, G, F, N = I.ntracks, I.n_egroups, I.fai, I.n_source_regions_per_node
        return {
            ARR_FINE_SOURCE: ((N, F, G), np.float32), ARR_FINE_FLUX: ((N, F, G), np.float32),
            ARR_SIGT: ((N, G), np.float32), ARR_PSI: ((T3, 2, G), np.float32),
            ARR_Z_HEIGHT: ((T3,), np.float32), ARR_P_WEIGHT: ((T3,), np.float32),
            ARR_SEG_COUNT: ((T3,), np.uint32), ARR_QSR_DIGEST: ((4,), np.uint64),
            ARR_QSR_DIGEST_BACK: ((4,), np.uint64)}[which]

    def get(self, which):
        shape, dt = self._shape(which)
        out = np.empty(shape, dtype=dt)
        _check(lib().moc_get_array(self.h, which, out.ctypes.data, out.nbytes), "moc_get_array")
        return out

    def set(self, which, arr):
        shape, dt = self._shape(which)
        a = np.ascontiguousarray(arr, dtype=dt).reshape(shape)
        _check(lib().moc_set_array(self.h, which, a.ctypes.data, a.nbytes), "moc_set_array")


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().moc_comm_get_unique_id(buf), "moc_comm_get_unique_id")
    return buf.raw
