/* moc_b200.h -- C-ABI of libmoc_b200.so: a B200 (sm_100a) implementation of the
 * SimpleMOC 3D MOC transport sweep behind the reference's own C interface.
 *
 * There is no plugin/FFI layer in the reference: the boundary of the hot path is
 * the set of C prototypes in reference src/SimpleMOC_header.h:224-237,250 called
 * from the iteration loop in src/main.c:57-92.  This header therefore has two
 * parts:
 *
 *   PART A  the data model of the reference (Input, Track, Source, Params, ...),
 *           layout-compatible field by field with src/SimpleMOC_header.h:28-158
 *           (serial build: no MPI, no OPENMP, no PAPI members) so that a driver
 *           written against the reference compiles against this header and the
 *           reference's own objects can hand their structures to this library.
 *
 *   PART B  the entry points.
 *           B1: the five hot functions under the reference's names and
 *               signatures (host structures in, host structures out).
 *           B2: a handle API that keeps the problem resident in HBM between
 *               phases (what the driver uses for speed).
 *           B3: host-side helpers mirroring init.c / io.c / tracks.c / source.c
 *               (configuration, CLI, synthetic problem construction).
 *
 * Every function that can fail returns 0 on success, a negative MOC_E* code on
 * failure; moc_last_error() gives the message.  The reference itself has no
 * error returns (it printf()s and exit(1)s: src/solver.c:506-511,
 * src/io.c:184-195); the drop-in names B1 keep that behaviour.
 *
 * No CPU fallback exists: every compute entry point fails with MOC_ENODEVICE if
 * no CUDA device is usable.
 */
#ifndef MOC_B200_H
#define MOC_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* =========================================================================
 * PART A -- data model (reference src/SimpleMOC_header.h)
 * ========================================================================= */

/* User inputs + derived sizes.  src/SimpleMOC_header.h:28-76.  sizeof == 152. */
typedef struct {
    int x_assemblies;
    int y_assemblies;
    int cai;                    /* coarse axial intervals                        */
    int fai;                    /* fine axial intervals per coarse interval      */
    int axial_exp;              /* axial source expansion order: 0 or 2          */
    float radial_ray_sep;
    float axial_z_sep;
    int n_azimuthal;
    int n_polar_angles;
    int n_egroups;
    bool decompose;
    int decomp_assemblies_ax;
    long segments_per_track;
    float assembly_width;
    float height;
    float domain_height;
    float precision;
    long mype;
    long ntracks_2D;            /* derived */
    int z_stacked;              /* derived */
    long ntracks;               /* derived */
    int nthreads;
    int papi_event_set;
    long n_2D_source_regions_per_assembly;
    long n_source_regions_per_node;   /* derived */
    bool load_tracks;
    char *track_file;
    long segments_processed;    /* OUT of transport_sweep (src/solver.c:549)     */
} Input;

/* src/SimpleMOC_header.h:86-89.  sizeof == 16. */
typedef struct {
    float length;
    long source_id;
} Segment;

/* src/SimpleMOC_header.h:92-97.  sizeof == 32. */
typedef struct {
    float az_weight;
    long n_segments;
    Segment *segments;
    int n_3D_segments;
} Track2D;

/* src/SimpleMOC_header.h:100-107.  sizeof == 40. */
typedef struct {
    float p_weight;
    float z_height;
    long rank_in;
    long rank_out;
    float *f_psi;
    float *b_psi;
} Track;

/* src/SimpleMOC_header.h:110-120 (serial build, no omp locks).  sizeof == 48. */
typedef struct {
    float **fine_flux;          /* [fai] -> [G]                                  */
    float **fine_source;        /* [fai] -> [G]                                  */
    float vol;
    float *sigT;                /* [G]                                           */
    float **XS;                 /* [G] -> {nu*SigmaF, SigmaA, Chi}               */
    float **scattering_matrix;  /* [G] -> [G]                                    */
} Source;

/* src/SimpleMOC_header.h:123-128. */
typedef struct {
    float *values;
    float dx;
    float maxVal;
    int N;
} Table;

/* src/SimpleMOC_header.h:131-138. */
typedef struct {
    Track2D *tracks_2D;
    Track ***tracks;
    Source *sources;
    float *polar_angles;
    float *leakage;
    Table expTable;
} Params;

/* src/SimpleMOC_header.h:141-158 without the MPI members.  A neighbour of -1 is
 * a domain border (MPICH's MPI_PROC_NULL, src/comms.c:118,146). */
typedef struct {
    int x_pos_src, x_pos_dest;
    int x_neg_src, x_neg_dest;
    int y_pos_src, y_pos_dest;
    int y_neg_src, y_neg_dest;
    int z_pos_src, z_pos_dest;
    int z_neg_src, z_neg_dest;
} CommGrid;

/* =========================================================================
 * PART B1 -- the hot path under the reference's names (drop-in)
 *
 * These take the reference's host structures.  The first call on a given
 * Params uploads the problem to the current CUDA device; every call runs on
 * the GPU and writes back exactly what the reference function mutates on the
 * host, unless residency is enabled (moc_set_resident), in which case the
 * write-back is deferred to moc_sync_to_host().
 * ========================================================================= */

/* src/solver.c:283-552.  Mutates tracks[][][].z_height, f_psi, sources[].fine_flux;
 * writes I->segments_processed. */
void transport_sweep(Params *params, Input *I);
/* src/solver.c:556-891 (defined in the reference, declared in none of its headers, called from nowhere).
 * Mutates f_psi AND b_psi, sources[].fine_flux; leaves every z_height at its start value; writes
 * I->segments_processed (both passes; 0 for the flat source). */
void two_way_transport_sweep(Params *params, Input *I);
/* src/solver.c:1143-1230.  Scales fine_flux and every f_psi/b_psi. */
void renormalize_flux(Params params, Input I, CommGrid grid);
/* src/solver.c:1235-1320.  Rewrites fine_source; returns the residual. */
float update_sources(Params params, Input I, float keff);
/* src/solver.c:1324-1437. */
float compute_keff(Params params, Input I, CommGrid grid);
/* src/comms.c:5-196.  Exchanges/zeroes the leading chunks of the psi slab and
 * accumulates *params.leakage on border faces.  Needs moc_comm_init() first
 * when any neighbour is not -1. */
void fast_transfer_boundary_fluxes(Params params, Input I, CommGrid grid);

/* What the reference takes from process-global state, the drop-in names take from here:
 * the seed and position of the random stream (the reference: srand(time(NULL)) + the rand()
 * calls made so far), the exponential mode (MOC_OPT_EXP_MODE) and sizeof(Source) of the caller
 * (48, or 56 when it was compiled with -DOPENMP).  Applies to mirrors created afterwards. */
void moc_dropin_configure(unsigned long long seed, unsigned long long rand_base, int exp_mode,
                          int source_stride);
/* A host program that is NOT edited (the reference's main.c linked as it is) never calls the function
 * above.  If it -- or anything linked into it -- exports
 *     void moc_host_rand_state(unsigned long long *seed, unsigned long long *calls);
 * the library calls it (dlsym) when the first drop-in call builds the mirror, and takes the seed and
 * position of the host's rand() stream from there.  oracle/dropin_glue.c is such an export. */
/* CUDA device the drop-in names create their mirror on (cudaSetDevice) */
int moc_set_device(int device);
/* 1: keep results on the device between the calls above; 0 (default): write back */
void moc_set_resident(int on);
/* Half way between the two: 1 = the caller promises not to MODIFY the host structures between drop-in calls
 * except through these calls (reading them is fine: the reference's main.c only prints).  Uploads are then
 * skipped -- the device copy is what the library last wrote back -- and every call still writes its results to the
 * host structures: half the traffic of the default.  0 (default): the host is authoritative, every call uploads
 * what it reads. */
void moc_dropin_trust_device(int on);
/* Resident mode only: tell the library the neighbour table before the loop.  transport_sweep then
 * starts the boundary exchange (comms.c:5-196) as soon as the z-stacks it moves are swept and runs it
 * under the sweep of the interior stacks; the fast_transfer_boundary_fluxes call that follows with the
 * same grid only collects it.  NULL switches the overlap off. */
void moc_dropin_set_grid(const CommGrid *grid);
/* download everything the phases run so far have mutated into the host Params */
int moc_sync_to_host(Params *params);
/* forget (and free) the device mirror that belongs to this Params */
int moc_release(Params *params);

/* =========================================================================
 * PART B2 -- handle API (problem resident in HBM)
 * ========================================================================= */

typedef struct moc_handle moc_handle;

enum {
    MOC_OK = 0,
    MOC_EINVAL = -1,     /* bad argument / unsupported configuration            */
    MOC_ENODEVICE = -2,  /* no usable CUDA device (there is no CPU fallback)    */
    MOC_ECUDA = -3,      /* CUDA runtime error (see moc_last_error)             */
    MOC_ENOMEM = -4,
    MOC_ELAYOUT = -5,    /* host Params is not laid out as the reference's slabs */
    MOC_ECOMM = -6,      /* NCCL error or communicator missing                  */
    MOC_EIO = -7         /* track file missing, truncated or corrupt            */
};

/* options for moc_set_option */
enum {
    MOC_OPT_EXP_MODE = 1,       /* 0 = TABLE_REF (bit-faithful table, default), 1 = SFU (__expf) */
    MOC_OPT_SEED = 2,           /* counter-RNG seed for source-region ids (default 1)            */
    MOC_OPT_RAND_BASE = 3,      /* rand() calls made before the sweep (serial-stream position)   */
    MOC_OPT_BATCH_SEGMENTS = 4, /* max 3D segments staged per batch (scratch size), default 2^28 */
    MOC_OPT_SOURCE_STRIDE = 5,  /* sizeof(Source) of the caller: 48 (default) or 56 (OPENMP)     */
    MOC_OPT_LANES_PER_TRACK = 6,/* override the lane mapping of the attenuation kernel (0=auto)  */
    MOC_OPT_STREAM_CHUNKS = 7,  /* z-stack chunks the host-side transport_sweep moves the angular
                                   flux in (copies overlap kernels), default 16                  */
    MOC_OPT_WALK_KERNEL = 8,    /* axial ray trace: 0 = auto (warps per z-stack: one for z_stacked <= 128,
                                   ceil(z_stacked / 128) up to 2048; one thread per ray above),
                                   1 = always one thread per ray, 2 = same as 0                  */
    MOC_OPT_FILL_OVERLAP = 9,   /* ray-trace CTAs per SM that emit the segment records of batch b+1
                                   under the attenuation of batch b (0 = emit in front of it);
                                   warp-per-stack ray trace only.  Default 0: same sweep time on
                                   a power-capped B200, but 2/batches of the record memory       */
    MOC_OPT_FILL_BATCHES = 10   /* batches per chunk of z-stacks when the two overlap, default 8 */
};

/* arrays for moc_get_array / moc_set_array (flat, the reference's slab order) */
enum {
    MOC_ARR_FINE_SOURCE = 1, /* float [N][fai][G]                                */
    MOC_ARR_FINE_FLUX = 2,   /* float [N][fai][G]                                */
    MOC_ARR_SIGT = 3,        /* float [N][G]                                     */
    MOC_ARR_PSI = 4,         /* float [ntracks][2][G]  (f row, b row)            */
    MOC_ARR_Z_HEIGHT = 5,    /* float [ntracks]                                  */
    MOC_ARR_P_WEIGHT = 6,    /* float [ntracks]                                  */
    MOC_ARR_SEG_COUNT = 7,   /* uint32 [ntracks]: 3D segments of each track in the last sweep */
    MOC_ARR_QSR_DIGEST = 8,  /* uint64 [4]: n, sum(row), sum(row*idx mix), xor-hash: see DESIGN.md */
    MOC_ARR_QSR_DIGEST_BACK = 9 /* uint64 [4]: the same over the backward pass of moc_two_way_sweep, keyed track*4096+segment */
};

/* timing of the last moc_sweep, CUDA events, milliseconds */
typedef struct {
    float count_ms;      /* geometry pass 1 (segment counts)                     */
    float scan_ms;       /* prefix sums                                          */
    float fill_ms;       /* geometry pass 2 (segment records), summed over batches */
    float attenuate_ms;  /* attenuation kernel, summed over batches              */
    float total_ms;      /* first launch to last completion                      */
    long n_batches;
    long launches;       /* kernels launched by the sweep                        */
} moc_sweep_timing;

/* Upload the problem described by the reference-layout host structures. */
int moc_create(const Input *I, const Params *P, int device, moc_handle **out);
/* The same synthetic problem as moc_build_tracks + moc_create, generated on the device (the
 * reference's build_tracks(), init.c:106-159, with every rand() draw taken from its position in the
 * counter stream): no host copy of the 3D-track and source arrays ever exists.  Only the handle API
 * works on such a handle (there are no host structures to write back to).  *rand_calls receives the
 * stream position after construction (the handle already uses it).  With I->load_tracks the 2D tracks
 * come from I->track_file and *I is updated like build_tracks(Input *) updates it (init.c:119-124). */
int moc_create_synthetic(Input *I, unsigned long long seed, int device, moc_handle **out,
                         unsigned long long *rand_calls);
int moc_destroy(moc_handle *h);
int moc_set_option(moc_handle *h, int option, long value);
long moc_get_option(moc_handle *h, int option);

int moc_sweep(moc_handle *h, long *segments_processed);        /* transport_sweep   */
/* two_way_transport_sweep (src/solver.c:556-891; compiled into the reference, called from nowhere): every ray is
 * walked forward into the forward angular flux, then retraced last segment first into the backward one.
 * *segments_processed counts both passes, and only for axial_exp == 2 (solver.c:752, 826).  Where a step length
 * comes out negative (solver.c:686) the reference's exponential table is read in front of its first cell; this
 * library answers from cell 0, like oracle/moc_oracle.c. */
int moc_two_way_sweep(moc_handle *h, long *segments_processed);
int moc_renormalize(moc_handle *h);                            /* renormalize_flux  */
int moc_update_sources(moc_handle *h, float keff, float *res); /* update_sources    */
int moc_compute_keff(moc_handle *h, float *keff);              /* compute_keff      */
int moc_exchange(moc_handle *h, const CommGrid *grid);         /* fast_transfer_... */
/* transport_sweep + fast_transfer_boundary_fluxes (main.c:60-70) as one call: the exchange starts
 * as soon as the z-stacks whose flux it moves (the first tracks of the slab, comms.c:100-183) are
 * swept and runs on a second stream under the sweep of the remaining stacks.  Same results as
 * moc_sweep followed by moc_exchange. */
int moc_sweep_exchange(moc_handle *h, const CommGrid *grid, long *segments_processed);

int moc_get_sweep_timing(moc_handle *h, moc_sweep_timing *t);
int moc_get_array(moc_handle *h, int which, void *dst, size_t bytes);
int moc_set_array(moc_handle *h, int which, const void *src, size_t bytes);
int moc_download(moc_handle *h, Params *P);   /* device -> host structures         */
int moc_upload(moc_handle *h, const Params *P); /* host structures -> device (mutable state) */
float moc_get_leakage(moc_handle *h);
int moc_synchronize(moc_handle *h);
/* the cudaStream_t every kernel of this handle is launched on (for CUDA-event timing by the caller) */
void *moc_get_stream(moc_handle *h);
/* diagnostics: bytes/s the L2 delivers for the attenuation kernel's access pattern on this
 * handle's source slab without the arithmetic (mode 0: row gathers, 1: gathers + vector
 * reductions) -- the measured ceiling bench.py reports the kernel's L2-level rate against */
int moc_probe_l2_gather(moc_handle *h, int mode, double *bytes_per_second);
/* kernels launched through this handle since moc_create (the library counts its own launches) */
long moc_get_launch_count(moc_handle *h);

/* multi-GPU: one process per GPU, one spatial domain per process.  The caller
 * distributes the 128-byte NCCL unique id (e.g. with torch.distributed). */
int moc_comm_get_unique_id(char id_out[128]);
int moc_comm_init(moc_handle *h, int nranks, int rank, const char id[128]);
/* neighbour table of a cx*cy*cz non-periodic Cartesian grid with MPI_Cart_shift
 * semantics (src/init.c:162-225; the reference hard-codes 2x2x1) */
int moc_make_grid(int cx, int cy, int cz, int rank, CommGrid *out);
/* The schedule of fast_transfer_boundary_fluxes (src/comms.c:12-28,75-183) for one rank, in the
 * reference's (round, direction) order: operation k covers `count` floats of the flux slab at
 * `offset`; they are sent to send_to (or, at a border, -1: pairwise-summed into the leakage,
 * src/comms.c:118-121) and replaced by what recv_from sent from the same offset of its own slab
 * (or, at a border, -1: by zeros, src/comms.c:146-149,179-181).  Pure host arithmetic: usable
 * without a GPU.  Returns the number of operations (ops may be NULL to query it), or MOC_EINVAL
 * if the plan does not fit the slab. */
typedef struct {
    long long offset, count;
    int round, direction;      /* direction: 0 x+, 1 x-, 2 y+, 3 y-, 4 z+, 5 z-  (src/comms.c:53-71) */
    int send_to, recv_from;
} moc_exchange_op;
long moc_exchange_plan(const Input *I, const CommGrid *grid, moc_exchange_op *ops, long max_ops);

/* the handle behind a Params used through the drop-in names (timing queries, moc_comm_init) */
moc_handle *moc_handle_of(Params *params);

const char *moc_last_error(void);
int moc_device_count(void);

/* =========================================================================
 * PART B3 -- host helpers (configuration and synthetic problem construction)
 * ========================================================================= */

Input moc_set_default_input(void);                 /* src/init.c:33-74            */
void moc_set_small_input(Input *I);                /* src/init.c:77-103           */
int moc_read_input_file(Input *I, const char *fname);   /* src/io.c:198-270       */
/* src/io.c:115-181.  Accepts -t -i -s -d (and the new long options of the driver);
 * returns MOC_EINVAL instead of exit(1) on a bad command line. */
int moc_read_CLI(int argc, char *argv[], Input *I);
void moc_calculate_derived_inputs(Input *I);       /* src/init.c:4-30             */
size_t moc_est_mem_usage(const Input *I);          /* src/utils.c:97-143          */
/* src/init.c:106-159 (+ tracks.c, source.c, utils.c:48-78): same slabs, same
 * draw order, draws taken from moc_rand31(seed, counter).  *rand_calls receives
 * the number of draws consumed (the serial-stream position at sweep start).
 * Takes Input * like build_tracks(Input *): a track file (I->load_tracks) changes *I. */
int moc_build_tracks(Input *I, uint64_t seed, Params *out, uint64_t *rand_calls);
/* src/tracks.c:170-323 (the -d option): the 2D tracks of an OpenMOC track file.  Updates
 * n_azimuthal, radial_ray_sep, ntracks_2D, segments_per_track and ntracks of *I as the reference
 * does; az_weight of track u is draw az_weight_at + u of the counter stream (tracks.c:283).
 * moc_build_tracks calls it when I->load_tracks is set (init.c:119-124).  A missing, truncated or
 * implausible file returns MOC_EIO (the reference does not check).  Free with moc_free_tracks or
 * free(tracks[0].segments); free(tracks). */
int moc_load_openmoc_tracks(const char *fname, int cmfd, Input *I, uint64_t seed, uint64_t az_weight_at,
                            Track2D **tracks, long *total_segments);
void moc_free_tracks(const Input *I, Params *P);
double moc_time_per_intersection(const Input *I, double seconds);   /* src/utils.c:147-155 */

/* Flat copies out of / into the pointer-rich host structures (tools and tests).
 * `which` is a MOC_ARR_* id (same shapes as moc_get_array) or one of the host-only ids
 * below.  Returns the number of bytes of the array; copies only if dst != NULL and
 * bytes matches. */
enum {
    MOC_HOST_AZ_WEIGHT = 20,   /* float [T2]                    */
    MOC_HOST_N_SEGMENTS = 21,  /* long  [T2]                    */
    MOC_HOST_SEG_LENGTHS = 22, /* float [sum n_segments]        */
    MOC_HOST_XS = 23,          /* float [N/8][G][3]             */
    MOC_HOST_SCATTER = 24,     /* float [N/8][G][G]             */
    MOC_HOST_XS_INDEX = 25,    /* int   [N]                     */
    MOC_HOST_VOL = 26,         /* float [N]                     */
    MOC_HOST_POLAR = 27,       /* float [P]                     */
    MOC_HOST_TABLE = 28        /* float [2*expTable.N]          */
};
long moc_params_get(const Input *I, const Params *P, int which, void *dst, size_t bytes);
long moc_params_set(const Input *I, Params *P, int which, const void *src, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* MOC_B200_H */
