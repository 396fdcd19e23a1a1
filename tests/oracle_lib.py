"""ctypes access to the two CPU checkers (TEST INFRASTRUCTURE, never the product):

  OracleCase  -> oracle/_build/libmoc_oracle.so   this repo's C restatement
  RefCase     -> oracle/_ref/libsimplemoc_ref*.so the unmodified reference + rand shim

Both expose the same attribute names (numpy views on the library's own buffers)
so tests can compare them with each other and with the CUDA path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libmoc_oracle.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REFERENCE_SRC = "/root/reference/src"


class Input(C.Structure):
    """reference src/SimpleMOC_header.h:28-76 (== include/moc_b200.h Input)"""
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


class CommGrid(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "x_pos_src", "x_pos_dest", "x_neg_src", "x_neg_dest",
        "y_pos_src", "y_pos_dest", "y_neg_src", "y_neg_dest",
        "z_pos_src", "z_pos_dest", "z_neg_src", "z_neg_dest")]


# The 18 positional values of a SimpleMOC input file (reference src/io.c:210-267),
# in file order.
INPUT_FILE_FIELDS = [
    "x_assemblies", "y_assemblies", "cai", "fai", "axial_exp", "radial_ray_sep",
    "axial_z_sep", "n_azimuthal", "n_polar_angles", "n_egroups", "decompose",
    "decomp_assemblies_ax", "segments_per_track", "assembly_width", "height",
    "precision", "n_2D_source_regions_per_assembly", "papi_event_set"]

# named problem sizes used across the tests (values in INPUT_FILE_FIELDS order)
CASES = {
    # SURVEY Appendix B "tiny.in": 60 2D tracks, 24 000 3D tracks, G=16
    "tiny": [3, 3, 4, 3, 2, 2.0, 4.0, 8, 4, 16, 0, 1, 10, 21.42, 400.0, 0.01, 200, 0],
    # G=104 like the real configs, decomposed axially so rays cross many fine intervals
    "mini104": [3, 3, 6, 5, 2, 3.0, 1.0, 6, 6, 104, 1, 8, 12, 21.42, 400.0, 0.01, 120, 0],
    # flat-source variant (axial_exp == 0)
    "tiny_flat": [3, 3, 4, 3, 0, 2.0, 4.0, 8, 4, 16, 0, 1, 10, 21.42, 400.0, 0.01, 200, 0],
    # G not a multiple of 4 (generic lane mapping), odd polar count
    "odd": [3, 3, 3, 4, 2, 2.5, 5.0, 6, 3, 10, 0, 1, 8, 21.42, 400.0, 0.01, 96, 0],
    # G=100 / 20 segments per track as in the shipped default.in, scaled down
    "mini_default_in": [17, 17, 9, 5, 2, 2.0, 0.25, 8, 10, 100, 1, 20, 20, 21.42, 400.0, 0.01, 80, 0],
    # z-stacks of 25 rays (one ray per lane of the warp-per-stack ray trace) and of 200 rays
    # (more than a warp can own: the CTA-per-stack ray trace)
    "short": [3, 3, 4, 3, 2, 2.0, 16.0, 8, 4, 16, 0, 1, 10, 21.42, 400.0, 0.01, 200, 0],
    "tall": [3, 3, 4, 3, 2, 4.0, 2.0, 8, 4, 16, 0, 1, 10, 21.42, 400.0, 0.01, 200, 0],
    # edge cases.  ragged: ~5 segments per 2D track, some tracks get NONE (seeds 2 and 4; a negative
    # draw -- seeds 5, 6 -- is undefined behaviour in the reference itself, tracks.c:29-47, and is not a
    # case).  polar1: a single polar angle, pi/2: every ray travels downward with cos ~ -4e-8.
    # polar2: one upward and one downward angle.  onecell: one coarse axial interval (fai = 3 fine
    # ones: every segment uses an edge stencil or the only interior row).  zone: one ray per stack.
    "ragged": [3, 3, 4, 3, 2, 2.0, 4.0, 8, 4, 16, 0, 1, 5, 21.42, 400.0, 0.01, 200, 0],
    "polar1": [3, 3, 4, 3, 2, 2.0, 4.0, 8, 1, 16, 0, 1, 10, 21.42, 400.0, 0.01, 200, 0],
    "polar2": [3, 3, 2, 3, 2, 2.0, 8.0, 8, 2, 12, 0, 1, 10, 21.42, 400.0, 0.01, 200, 0],
    "onecell": [3, 3, 1, 3, 2, 2.0, 4.0, 8, 4, 16, 0, 1, 10, 21.42, 400.0, 0.01, 200, 0],
    "zone": [3, 3, 4, 3, 2, 2.0, 400.0, 8, 4, 16, 0, 1, 10, 21.42, 400.0, 0.01, 200, 0],
    # flat source over one and over two fine intervals per coarse one: legal in the reference, fai < 3
    "flat_f1": [3, 3, 4, 1, 0, 2.0, 4.0, 8, 4, 16, 0, 1, 10, 21.42, 400.0, 0.01, 200, 0],
    "flat_f2": [3, 3, 4, 2, 0, 2.0, 4.0, 8, 4, 16, 0, 1, 10, 21.42, 400.0, 0.01, 200, 0],
    # group counts whose pairwise_sum trees are uneven (utils.c:29-45: 33 = 16 + 17 -> 17 = 8 + 9;
    # 130 = 65 + 65 -> 32 + 33 -> ...; 200): the reductions spread those trees over lanes
    "g33": [3, 3, 3, 4, 2, 2.5, 8.0, 6, 2, 33, 0, 1, 6, 21.42, 400.0, 0.01, 96, 0],
    "g130": [3, 3, 3, 3, 2, 2.5, 10.0, 6, 2, 130, 0, 1, 5, 21.42, 400.0, 0.01, 96, 0],
    "g200": [3, 3, 2, 3, 2, 3.0, 16.0, 6, 2, 200, 0, 1, 4, 21.42, 400.0, 0.01, 96, 0],
    # 96 000 short 3D tracks, G=8: large enough for one 10 000-track message per face
    # (comms.c:12-28), cheap to sweep -- the boundary-exchange case
    "exch": [3, 3, 4, 3, 2, 2.0, 0.05, 8, 4, 8, 1, 20, 10, 21.42, 400.0, 0.01, 200, 0],
}


def write_input_file(path, values):
    with open(path, "w") as f:
        for name, v in zip(INPUT_FILE_FIELDS, values):
            f.write(f"{v} - {name}\n")
    return path


def ensure_oracle_built():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


def ref_available(variant=""):
    return os.path.exists(os.path.join(REF_DIR, f"libsimplemoc_ref{variant}.so"))


def ensure_ref_built():
    """Build oracle/_ref from /root/reference if it is there; otherwise use what travelled."""
    if os.path.isdir(REFERENCE_SRC) and not ref_available():
        subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)
    return ref_available()


def _view(ptr, shape, dtype=np.float32):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    ctype = {np.float32: C.c_float, np.int32: C.c_int, np.int64: C.c_long,
             np.uint32: C.c_uint32, np.uint64: C.c_uint64, np.float64: C.c_double}[dtype]
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,))
    return arr.reshape(shape)


class _Base:
    """Common numpy views; subclasses set self._p(name) -> pointer getters."""

    def _sizes(self):
        I = self.I
        return (I.ntracks_2D, I.n_polar_angles, I.z_stacked, I.ntracks, I.n_egroups,
                I.fai, I.n_source_regions_per_node)

    @property
    def psi(self):
        T2, P, Z, T3, G, F, N = self._sizes()
        return _view(self._ptr("psi"), (T3, 2, G))

    @property
    def source_data(self):
        T2, P, Z, T3, G, F, N = self._sizes()
        return _view(self._ptr("source_data"), ((2 * F + 1) * N * G,))

    @property
    def fine_source(self):
        T2, P, Z, T3, G, F, N = self._sizes()
        return self.source_data[: N * F * G].reshape(N, F, G)

    @property
    def fine_flux(self):
        T2, P, Z, T3, G, F, N = self._sizes()
        return self.source_data[N * F * G: 2 * N * F * G].reshape(N, F, G)

    @property
    def sigT(self):
        T2, P, Z, T3, G, F, N = self._sizes()
        return self.source_data[2 * N * F * G:].reshape(N, G)

    @property
    def xs(self):
        G = self.I.n_egroups
        return _view(self._ptr("xs_data"), (self.n_xs, G, 3))

    @property
    def scatter(self):
        G = self.I.n_egroups
        return _view(self._ptr("scatter_data"), (self.n_xs, G, G))

    @property
    def polar_angles(self):
        return _view(self._ptr("polar_angles"), (self.I.n_polar_angles,))

    @property
    def table(self):
        dx, mx, n = C.c_float(), C.c_float(), C.c_int()
        self._table_info(C.byref(dx), C.byref(mx), C.byref(n))
        return _view(self._ptr("table_values"), (2 * n.value,)), dx.value, mx.value, n.value


class OracleCase(_Base):
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(ensure_oracle_built(), mode=C.RTLD_LOCAL)
            L.oracle_create.restype = C.c_void_p
            L.oracle_create.argtypes = [C.POINTER(Input), C.c_uint64]
            L.oracle_create_geometry.restype = C.c_void_p
            L.oracle_create_geometry.argtypes = [C.POINTER(Input), C.c_uint64]
            L.oracle_sweep_reversed.restype = C.c_long
            L.oracle_sweep_reversed.argtypes = [C.c_void_p]
            for n in ("psi64", "flux64"):
                getattr(L, "oracle_" + n).restype = C.c_void_p
                getattr(L, "oracle_" + n).argtypes = [C.c_void_p]
            L.oracle_destroy.argtypes = [C.c_void_p]
            L.oracle_derive.argtypes = [C.POINTER(Input)]
            L.oracle_sweep.restype = C.c_long
            L.oracle_sweep.argtypes = [C.c_void_p]
            L.oracle_two_way_sweep.restype = C.c_long
            L.oracle_two_way_sweep.argtypes = [C.c_void_p]
            L.oracle_renormalize.argtypes = [C.c_void_p]
            L.oracle_update_sources.restype = C.c_float
            L.oracle_update_sources.argtypes = [C.c_void_p, C.c_float]
            L.oracle_compute_keff.restype = C.c_float
            L.oracle_compute_keff.argtypes = [C.c_void_p]
            L.oracle_set_exp_mode.argtypes = [C.c_void_p, C.c_int]
            L.oracle_input.restype = C.POINTER(Input)
            L.oracle_input.argtypes = [C.c_void_p]
            for n in ("rand_calls", "init_rand_calls"):
                getattr(L, "oracle_" + n).restype = C.c_uint64
                getattr(L, "oracle_" + n).argtypes = [C.c_void_p]
            for n in ("n_xs_regions", "total_2d_segments", "trace_len", "table_oob"):
                getattr(L, "oracle_" + n).restype = C.c_long
                getattr(L, "oracle_" + n).argtypes = [C.c_void_p]
            for n in ("psi", "source_data", "xs_data", "scatter_data", "polar_angles",
                      "p_weight", "z_height", "az_weight", "n_segments", "seg_lengths",
                      "xs_index", "vol", "table_values", "leakage", "seg_count", "digest", "digest_back", "abs_flux", "abs_terms", "abs_psi",
                      "trace_track", "trace_row", "trace_ds", "trace_zstart"):
                getattr(L, "oracle_" + n).restype = C.c_void_p
                getattr(L, "oracle_" + n).argtypes = [C.c_void_p]
            L.oracle_table_info.argtypes = [C.c_void_p] + [C.c_void_p] * 3
            L.oracle_enable_trace.argtypes = [C.c_void_p, C.c_long]
            L.oracle_exchange.restype = C.c_int
            L.oracle_exchange.argtypes = [C.POINTER(C.c_void_p), C.POINTER(CommGrid), C.c_int]
            L.oracle_exchange_plan.argtypes = [C.POINTER(Input), C.POINTER(C.c_long)]
            L.oracle_make_grid.argtypes = [C.c_int] * 4 + [C.POINTER(CommGrid)]
            cls._lib = L
        return cls._lib

    def __init__(self, values, seed=1, exp_mode=0, limit_tracks_2D=0, track_file=None, geometry_only=False):
        """exp_mode 0: the reference's table, 1: 1 - expf(-x), 2: the attenuation in double with exp() (psi64 / flux64).
        geometry_only: ray trace, source-region draws and digest only -- no flux or material arrays (full-size pins)"""
        L = self.lib()
        inp = input_from_values(values, track_file)
        L.oracle_derive(C.byref(inp))
        if limit_tracks_2D and limit_tracks_2D < inp.ntracks_2D:
            inp.ntracks_2D = 2 * (limit_tracks_2D // 2)
            inp.ntracks = inp.ntracks_2D * inp.n_polar_angles * inp.z_stacked
        self.h = (L.oracle_create_geometry if geometry_only else L.oracle_create)(C.byref(inp), seed)
        if not self.h:
            raise RuntimeError(f"oracle_create failed (track file {track_file!r})")
        L.oracle_set_exp_mode(self.h, exp_mode)
        self.seed = seed

    def close(self):
        if self.h:
            self.lib().oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def I(self):
        return self.lib().oracle_input(self.h).contents

    def _ptr(self, name):
        return getattr(self.lib(), "oracle_" + name)(self.h)

    def _table_info(self, *a):
        self.lib().oracle_table_info(self.h, *a)

    @property
    def n_xs(self):
        return self.lib().oracle_n_xs_regions(self.h)

    @property
    def init_rand_calls(self):
        return self.lib().oracle_init_rand_calls(self.h)

    @property
    def rand_calls(self):
        return self.lib().oracle_rand_calls(self.h)

    @property
    def p_weight(self):
        return _view(self._ptr("p_weight"), (self.I.ntracks,))

    @property
    def z_height(self):
        return _view(self._ptr("z_height"), (self.I.ntracks,))

    @property
    def az_weight(self):
        return _view(self._ptr("az_weight"), (self.I.ntracks_2D,))

    @property
    def n_segments(self):
        return _view(self._ptr("n_segments"), (self.I.ntracks_2D,), np.int64)

    @property
    def seg_lengths(self):
        return _view(self._ptr("seg_lengths"), (self.lib().oracle_total_2d_segments(self.h),))

    @property
    def xs_index(self):
        return _view(self._ptr("xs_index"), (self.I.n_source_regions_per_node,), np.int32)

    @property
    def vol(self):
        return _view(self._ptr("vol"), (self.I.n_source_regions_per_node,))

    @property
    def leakage(self):
        return _view(self._ptr("leakage"), (1,))

    @property
    def seg_count(self):
        return _view(self._ptr("seg_count"), (self.I.ntracks,), np.uint32)

    @property
    def abs_flux(self):
        I = self.I
        return _view(self._ptr("abs_flux"), (I.n_source_regions_per_node, I.fai, I.n_egroups))

    @property
    def abs_terms(self):
        """per element of fine_flux: sum over its tallies of the magnitudes of the terms the reference's formula
        adds to form each tally -- the scale of the reference's own rounding error (cancellation inside a tally)"""
        I = self.I
        return _view(self._ptr("abs_terms"), (I.n_source_regions_per_node, I.fai, I.n_egroups))

    @property
    def abs_psi(self):
        """the same scale for the angular flux, carried along every track (a running error bound)"""
        I = self.I
        return _view(self._ptr("abs_psi"), (I.ntracks, 2, I.n_egroups))

    @property
    def digest(self):
        return _view(self._ptr("digest"), (4,), np.uint64)

    def sweep(self):
        return self.lib().oracle_sweep(self.h)

    def sweep_reversed(self):
        """the same sweep, 2D tracks in reverse order: same draws and ray states, tallies added in another order"""
        return self.lib().oracle_sweep_reversed(self.h)

    @property
    def psi64(self):
        T2, P, Z, T3, G, F, N = self._sizes()
        return _view(self._ptr("psi64"), (T3, 2, G), np.float64)

    @property
    def flux64(self):
        T2, P, Z, T3, G, F, N = self._sizes()
        return _view(self._ptr("flux64"), (N, F, G), np.float64)

    def two_way_sweep(self):
        """solver.c:556-891 (the v1 sweep, forward + backward flux)"""
        return self.lib().oracle_two_way_sweep(self.h)

    @property
    def table_oob(self):
        """two-way sweep: table lookups below cell 0 (out-of-bounds reads in the reference)"""
        return self.lib().oracle_table_oob(self.h)

    @property
    def digest_back(self):
        return _view(self._ptr("digest_back"), (4,), np.uint64)

    def renormalize(self):
        self.lib().oracle_renormalize(self.h)

    def update_sources(self, keff):
        return self.lib().oracle_update_sources(self.h, keff)

    def compute_keff(self):
        return self.lib().oracle_compute_keff(self.h)

    def enable_trace(self, cap):
        self.lib().oracle_enable_trace(self.h, cap)

    def trace(self):
        n = self.lib().oracle_trace_len(self.h)
        return (_view(self._ptr("trace_track"), (n,), np.uint32),
                _view(self._ptr("trace_row"), (n,), np.uint32),
                _view(self._ptr("trace_ds"), (n,)),
                _view(self._ptr("trace_zstart"), (n,)))


def input_from_values(values, track_file=None):
    """An Input as main() would have it after reading an input file (before derive);
    track_file: what `-d <file>` adds (src/io.c:160-168)."""
    inp = Input()
    # set_default_input (reference src/init.c:33-74) for what the file does not carry
    inp.mype = 0
    inp.nthreads = 1
    inp.load_tracks = False
    for name, v in zip(INPUT_FILE_FIELDS, values):
        if name == "decompose":
            inp.decompose = bool(v)
        else:
            setattr(inp, name, v)
    if track_file:
        inp.load_tracks = True
        inp.track_file = os.fsencode(track_file)
    return inp


class RefCase(_Base):
    """The unmodified reference (oracle/_ref).  variant: "", "_expf" or "_omp"."""
    _libs = {}

    @classmethod
    def lib(cls, variant=""):
        if variant not in cls._libs:
            path = os.path.join(REF_DIR, f"libsimplemoc_ref{variant}.so")
            L = C.CDLL(path, mode=C.RTLD_LOCAL)
            L.ref_case_create.restype = C.c_void_p
            L.ref_case_create.argtypes = [C.c_char_p, C.c_int, C.c_uint64, C.c_int, C.c_long]
            L.ref_case_create_tracks.restype = C.c_void_p
            L.ref_case_create_tracks.argtypes = [C.c_char_p, C.c_int, C.c_uint64, C.c_int, C.c_long, C.c_char_p]
            L.ref_case_destroy.argtypes = [C.c_void_p]
            L.ref_transport_sweep.restype = C.c_long
            L.ref_transport_sweep.argtypes = [C.c_void_p]
            L.ref_two_way_transport_sweep.restype = C.c_long
            L.ref_two_way_transport_sweep.argtypes = [C.c_void_p]
            L.ref_time_transport_sweep.restype = C.c_double
            L.ref_time_transport_sweep.argtypes = [C.c_void_p]
            L.ref_renormalize_flux.argtypes = [C.c_void_p]
            L.ref_update_sources.restype = C.c_float
            L.ref_update_sources.argtypes = [C.c_void_p, C.c_float]
            L.ref_compute_keff.restype = C.c_float
            L.ref_compute_keff.argtypes = [C.c_void_p]
            L.ref_input.restype = C.POINTER(Input)
            L.ref_input.argtypes = [C.c_void_p]
            L.ref_sizeof.restype = C.c_long
            L.ref_init_rand_calls.restype = C.c_uint64
            L.ref_init_rand_calls.argtypes = [C.c_void_p]
            L.ref_rand_calls.restype = C.c_uint64
            for n in ("n_xs_regions", "total_2d_segments"):
                getattr(L, "ref_" + n).restype = C.c_long
                getattr(L, "ref_" + n).argtypes = [C.c_void_p]
            for n in ("psi", "source_data", "xs_data", "scatter_data", "polar_angles",
                      "leakage", "table_values", "params", "input_mut"):
                getattr(L, "ref_" + n).restype = C.c_void_p
                getattr(L, "ref_" + n).argtypes = [C.c_void_p]
            L.ref_table_info.argtypes = [C.c_void_p] * 4
            L.ref_copy_tracks.argtypes = [C.c_void_p] * 3
            L.ref_copy_tracks_2D.argtypes = [C.c_void_p] * 4
            L.ref_copy_source_meta.argtypes = [C.c_void_p] * 3
            cls._libs[variant] = L
        return cls._libs[variant]

    def __init__(self, values=None, seed=1, variant="", small=False, nthreads=1,
                 limit_tracks_2D=0, tmpdir="/tmp", track_file=None):
        self.variant = variant
        L = self.lib(variant)
        path = b""
        if values is not None:
            fn = os.path.join(tmpdir, f"moc_case_{os.getpid()}_{id(self)}.in")
            write_input_file(fn, values)
            path = fn.encode()
        self._track_file = os.fsencode(track_file) if track_file else None   # the reference keeps the pointer
        self.h = L.ref_case_create_tracks(path, int(small), seed, nthreads, limit_tracks_2D, self._track_file)
        if path:
            os.unlink(path.decode())

    def close(self):
        if self.h:
            self.lib(self.variant).ref_case_destroy(self.h)
            self.h = None

    @property
    def I(self):
        return self.lib(self.variant).ref_input(self.h).contents

    def _ptr(self, name):
        return getattr(self.lib(self.variant), "ref_" + name)(self.h)

    def _table_info(self, *a):
        self.lib(self.variant).ref_table_info(self.h, *a)

    @property
    def n_xs(self):
        return self.lib(self.variant).ref_n_xs_regions(self.h)

    @property
    def init_rand_calls(self):
        return self.lib(self.variant).ref_init_rand_calls(self.h)

    @property
    def rand_calls(self):
        return self.lib(self.variant).ref_rand_calls()

    def tracks(self):
        n = self.I.ntracks
        pw = np.empty(n, np.float32)
        zh = np.empty(n, np.float32)
        self.lib(self.variant).ref_copy_tracks(self.h, pw.ctypes.data, zh.ctypes.data)
        return pw, zh

    @property
    def p_weight(self):
        return self.tracks()[0]

    @property
    def z_height(self):
        return self.tracks()[1]

    def tracks_2D(self):
        T2 = self.I.ntracks_2D
        az = np.empty(T2, np.float32)
        ns = np.empty(T2, np.int64)
        ln = np.empty(self.lib(self.variant).ref_total_2d_segments(self.h), np.float32)
        self.lib(self.variant).ref_copy_tracks_2D(self.h, az.ctypes.data, ns.ctypes.data,
                                                  ln.ctypes.data)
        return az, ns, ln

    def source_meta(self):
        N = self.I.n_source_regions_per_node
        idx = np.empty(N, np.int32)
        vol = np.empty(N, np.float32)
        self.lib(self.variant).ref_copy_source_meta(self.h, idx.ctypes.data, vol.ctypes.data)
        return idx, vol

    @property
    def leakage(self):
        return _view(self._ptr("leakage"), (1,))

    def sweep(self):
        return self.lib(self.variant).ref_transport_sweep(self.h)

    def two_way_sweep(self):
        """the reference's own two_way_transport_sweep (solver.c:556-891)"""
        return self.lib(self.variant).ref_two_way_transport_sweep(self.h)

    def time_sweep(self):
        return self.lib(self.variant).ref_time_transport_sweep(self.h)

    def renormalize(self):
        self.lib(self.variant).ref_renormalize_flux(self.h)

    def update_sources(self, keff):
        return self.lib(self.variant).ref_update_sources(self.h, keff)

    def compute_keff(self):
        return self.lib(self.variant).ref_compute_keff(self.h)


def make_grid(cx, cy, cz, rank):
    g = CommGrid()
    OracleCase.lib().oracle_make_grid(cx, cy, cz, rank, C.byref(g))
    return g


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a - b))


def noise_units(a, b, magnitude):
    """|a-b| in units of FP32 epsilon times the accumulation magnitude of each element
    (backward-error view of a sum whose terms cancel)."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    mag = np.maximum(np.asarray(magnitude, np.float64).ravel(), 1e-300)
    return np.abs(a - b) / (np.finfo(np.float32).eps * mag)


def frac_within(a, b, tol):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    ok = np.abs(a - b) <= tol * np.abs(b)
    return float(ok.mean()) if ok.size else 1.0
