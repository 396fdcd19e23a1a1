/* ref_harness.c -- TEST INFRASTRUCTURE (oracle).  Not part of the product.
 *
 * A thin in-process entry layer over the UNMODIFIED reference objects
 * (compiled from /root/reference/src/{solver,source,tracks,init,io,utils,
 * comms,test,papi}.c by oracle/Makefile; main.c is left out).  It replaces
 * only the reference's main() (src/main.c:3-147): it builds a problem the way
 * main() does, runs the individual phases on request and hands the reference's
 * own buffers to the caller so tests can compare them with the CUDA path.
 *
 * Built three ways (see oracle/Makefile):
 *   libsimplemoc_ref.so       serial, -O2 -ffp-contract=off, rand()/time()
 *                             pinned by ref_shim.c               (parity oracle)
 *   libsimplemoc_ref_expf.so  same, interpolateTable overridden by
 *                             ref_expf_override.c                (exact-exp oracle)
 *   libsimplemoc_ref_omp.so   stock Makefile flags, OpenMP, libc rand_r
 *                             (timing baseline only; its keff is NaN)
 *   libsimplemoc_ref_mpi.so   as the first, every source compiled with -DMPI against the
 *                             in-process MPI of oracle/mpi_stub (ranks = pthreads): the pin of
 *                             comms.c:5-196, init.c:162-225 and the MPI reductions of solver.c
 *   libsimplemoc_ref_ofast.so as the first with -Ofast -ffast-math -mfma: reference-vs-reference
 *                             noise floor (tools/tolerance_anchor.py)
 */
#include "SimpleMOC_header.h"
#include <fcntl.h>
#include <stdint.h>

#ifdef REF_WITH_SHIM
void ref_shim_reset(uint64_t seed);
uint64_t ref_shim_calls(void);
#endif

typedef struct {
    Input I;
    Params P;
    CommGrid grid;
    long n_xs_regions;
    long total_2d_segments;
    uint64_t init_rand_calls;
} RefCase;

/* ---- stdout silencing: the reference prints progress from inside the hot loop */
static int g_verbose = 0;
void ref_set_verbose(int v) { g_verbose = v; }

static int quiet_begin(void)
{
    if (g_verbose) return -1;
    fflush(stdout);
    int saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    if (nul >= 0) { dup2(nul, 1); close(nul); }
    return saved;
}
static void quiet_end(int saved)
{
    if (saved < 0) return;
    fflush(stdout);
    dup2(saved, 1);
    close(saved);
}

/* Mirrors main(): set_default_input -> (-s) -> (-i file) -> derived -> build_tracks.
 * limit_tracks_2D > 0 shrinks the number of 2D tracks AFTER the derived inputs are
 * computed (bounded-sample timing; everything per-track is unchanged). */
RefCase *ref_case_create_tracks(const char *input_file, int small, uint64_t seed,
                                int nthreads, long limit_tracks_2D, const char *track_file);

RefCase *ref_case_create(const char *input_file, int small, uint64_t seed,
                         int nthreads, long limit_tracks_2D)
{
    return ref_case_create_tracks(input_file, small, seed, nthreads, limit_tracks_2D, NULL);
}

/* track_file != NULL: what "-d <file>" does (src/io.c:160-168): the reference's own
 * load_OpenMOC_tracks (src/tracks.c:170-323) supplies the 2D tracks. */
RefCase *ref_case_create_tracks(const char *input_file, int small, uint64_t seed,
                                int nthreads, long limit_tracks_2D, const char *track_file)
{
    RefCase *c = (RefCase *)calloc(1, sizeof(RefCase));
    int saved = quiet_begin();
#ifdef REF_WITH_SHIM
    ref_shim_reset(seed);
#else
    srand((unsigned)seed);
#endif
    c->I = set_default_input();
    c->I.nthreads = nthreads > 0 ? nthreads : 1;
    c->I.segments_processed = 0;
    c->I.track_file = NULL;
    c->I.papi_event_set = 0;
    if (small) set_small_input(&c->I);
    if (input_file && input_file[0]) read_input_file(&c->I, (char *)input_file);
    c->I.nthreads = nthreads > 0 ? nthreads : 1;
    if (track_file && track_file[0]) {
        c->I.track_file = (char *)track_file;
        c->I.load_tracks = true;
    }
    calculate_derived_inputs(&c->I);
    if (limit_tracks_2D > 0 && limit_tracks_2D < c->I.ntracks_2D) {
        c->I.ntracks_2D = 2 * (limit_tracks_2D / 2);
        c->I.ntracks = c->I.ntracks_2D * c->I.n_polar_angles * c->I.z_stacked;
    }
#ifdef OPENMP
    omp_set_num_threads(c->I.nthreads);
#endif
    c->P = build_tracks(&c->I);
#ifndef MPI
    c->grid = init_mpi_grid(c->I);   /* without MPI: returns an unset grid (init.c:166) */
#else
    memset(&c->grid, 0, sizeof c->grid);   /* collective under MPI: ref_mpi_run(..., what & 1) or ref_mpi_set_grid */
#endif
    c->n_xs_regions = c->I.n_source_regions_per_node / 8;
    long tot = 0;
    for (long i = 0; i < c->I.ntracks_2D; i++) tot += c->P.tracks_2D[i].n_segments;
    c->total_2d_segments = tot;
#ifdef REF_WITH_SHIM
    c->init_rand_calls = ref_shim_calls();
#endif
    quiet_end(saved);
    return c;
}

void ref_case_destroy(RefCase *c)
{
    if (!c) return;
    free(c->P.tracks[0][0][0].f_psi);
    free(&c->P.tracks[0][0][0]);
    free(c->P.tracks[0]);
    free_tracks(c->P.tracks);
    free_2D_tracks(c->P.tracks_2D);
    /* not free_sources(): src/source.c:216-230 frees fine_flux[0], which points into
     * the middle of the shared source slab (src/source.c:141-144) -- it is never
     * called by the reference's main() either.  Free the slabs by their bases. */
    {
        Source *s0 = &c->P.sources[0];
        free(s0->XS[0]);
        free(s0->XS);
        free(s0->scattering_matrix[0]);
        free(s0->scattering_matrix);
        free(s0->fine_source[0]);
        free(s0->fine_source);
        free(s0->fine_flux);
    }
    free(c->P.sources);
    free(c->P.polar_angles);
    free(c->P.leakage);
    free(c->P.expTable.values);
    free(c);
}

/* ---- phases (reference src/main.c:57-92) ---- */
long ref_transport_sweep(RefCase *c)
{
    int saved = quiet_begin();
    transport_sweep(&c->P, &c->I);
    quiet_end(saved);
    return c->I.segments_processed;
}

/* solver.c:556-891: compiled into the reference but not declared in its header (never called by main.c) */
void two_way_transport_sweep(Params *params, Input *I);
long ref_two_way_transport_sweep(RefCase *c)
{
    int saved = quiet_begin();
    two_way_transport_sweep(&c->P, &c->I);
    quiet_end(saved);
    return c->I.segments_processed;
}

double ref_time_transport_sweep(RefCase *c)
{
    struct timespec t0, t1;
    int saved = quiet_begin();
    clock_gettime(CLOCK_MONOTONIC, &t0);
    transport_sweep(&c->P, &c->I);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    quiet_end(saved);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

void ref_renormalize_flux(RefCase *c)
{
    int saved = quiet_begin();
    renormalize_flux(c->P, c->I, c->grid);
    quiet_end(saved);
}

float ref_update_sources(RefCase *c, float keff) { return update_sources(c->P, c->I, keff); }
float ref_compute_keff(RefCase *c) { return compute_keff(c->P, c->I, c->grid); }

/* ---- accessors: the reference's own buffers ---- */
const Input *ref_input(RefCase *c) { return &c->I; }
long ref_sizeof(int what)
{
    switch (what) {
    case 0: return (long)sizeof(Input);
    case 1: return (long)sizeof(Track);
    case 2: return (long)sizeof(Track2D);
    case 3: return (long)sizeof(Segment);
    case 4: return (long)sizeof(Source);
    case 5: return (long)sizeof(Params);
    case 6: return (long)sizeof(CommGrid);
    case 7: return (long)sizeof(Table);
    }
    return -1;
}
uint64_t ref_init_rand_calls(RefCase *c) { return c->init_rand_calls; }
uint64_t ref_rand_calls(void)
{
#ifdef REF_WITH_SHIM
    return ref_shim_calls();
#else
    return 0;
#endif
}
long ref_n_xs_regions(RefCase *c) { return c->n_xs_regions; }
long ref_total_2d_segments(RefCase *c) { return c->total_2d_segments; }

/* psi slab: [ntracks][2][G], forward rows first in each pair (src/tracks.c:106-138) */
float *ref_psi(RefCase *c) { return c->P.tracks[0][0][0].f_psi; }
/* source slab: fine_source[N][fai][G] | fine_flux[N][fai][G] | sigT[N][G]
 * (src/source.c:121-152) */
float *ref_source_data(RefCase *c) { return c->P.sources[0].fine_source[0]; }
float *ref_xs_data(RefCase *c) { return c->P.sources[0].XS[0]; }
float *ref_scatter_data(RefCase *c) { return c->P.sources[0].scattering_matrix[0]; }
float *ref_polar_angles(RefCase *c) { return c->P.polar_angles; }
float *ref_leakage(RefCase *c) { return c->P.leakage; }
float *ref_table_values(RefCase *c) { return c->P.expTable.values; }
void ref_table_info(RefCase *c, float *dx, float *maxVal, int *N)
{
    *dx = c->P.expTable.dx;
    *maxVal = c->P.expTable.maxVal;
    *N = c->P.expTable.N;
}

void ref_copy_tracks(RefCase *c, float *p_weight, float *z_height)
{
    Track *t = &c->P.tracks[0][0][0];
    for (long i = 0; i < c->I.ntracks; i++) {
        p_weight[i] = t[i].p_weight;
        z_height[i] = t[i].z_height;
    }
}

void ref_copy_tracks_2D(RefCase *c, float *az_weight, long *n_segments, float *lengths)
{
    long idx = 0;
    for (long i = 0; i < c->I.ntracks_2D; i++) {
        az_weight[i] = c->P.tracks_2D[i].az_weight;
        n_segments[i] = c->P.tracks_2D[i].n_segments;
        for (long n = 0; n < c->P.tracks_2D[i].n_segments; n++)
            lengths[idx++] = c->P.tracks_2D[i].segments[n].length;
    }
}

void ref_copy_source_meta(RefCase *c, int *xs_index, float *vol)
{
    float *xs0 = c->P.sources[0].XS[0];
    long G = c->I.n_egroups;
    for (long i = 0; i < c->I.n_source_regions_per_node; i++) {
        xs_index[i] = (int)((c->P.sources[i].XS[0] - xs0) / (3 * G));
        vol[i] = c->P.sources[i].vol;
    }
}

#ifdef MPI
/* ---- the communication path (src/main.c:66-71, 73-91 under -DMPI), one pthread per rank ----
 * cases[r] is rank r's domain, built on the calling thread with mpi_stub_set_rank(r) (see
 * ref_mpi_case_create).  what: 1 = init_mpi_grid (init.c:162-225; dims {2,2,1}: four ranks),
 * 2 = fast_transfer_boundary_fluxes (comms.c:5-196), 4 = renormalize_flux, 8 = compute_keff
 * (keff_out[r]; only rank 0's is defined, solver.c:1394-1425).  Without bit 1 the grids the cases
 * already hold are used (ref_mpi_set_grid: a 1x1x1 or any other neighbour table). */
typedef struct { RefCase **cases; int what; float *keff_out; } MpiJob;

static void mpi_rank_main(int rank, void *arg)
{
    MpiJob *job = (MpiJob *)arg;
    RefCase *c = job->cases[rank];
    if (job->what & 1) c->grid = init_mpi_grid(c->I);
    if (job->what & 2) fast_transfer_boundary_fluxes(c->P, c->I, c->grid);
    if (job->what & 4) renormalize_flux(c->P, c->I, c->grid);
    if (job->what & 8) job->keff_out[rank] = compute_keff(c->P, c->I, c->grid);
}

int ref_mpi_run(int nranks, RefCase **cases, int what, float *keff_out)
{
    MpiJob job = { cases, what, keff_out };
    int saved = quiet_begin();
    int rc = mpi_stub_run(nranks, mpi_rank_main, &job);
    quiet_end(saved);
    return rc;
}

/* rank `rank` of `nranks`: calculate_derived_inputs asks MPI_Comm_rank (init.c:7-11) */
RefCase *ref_mpi_case_create(const char *input_file, uint64_t seed, int rank, int nranks)
{
    mpi_stub_world(nranks);
    mpi_stub_set_rank(rank);
    RefCase *c = ref_case_create_tracks(input_file, 0, seed, 1, 0, NULL);
    mpi_stub_set_rank(0);
    return c;
}

/* the twelve neighbour ranks in CommGrid order (x_pos_src .. z_neg_dest) */
void ref_mpi_get_grid(RefCase *c, int out[12])
{
    const int v[12] = { c->grid.x_pos_src, c->grid.x_pos_dest, c->grid.x_neg_src, c->grid.x_neg_dest,
                        c->grid.y_pos_src, c->grid.y_pos_dest, c->grid.y_neg_src, c->grid.y_neg_dest,
                        c->grid.z_pos_src, c->grid.z_pos_dest, c->grid.z_neg_src, c->grid.z_neg_dest };
    memcpy(out, v, sizeof v);
}
void ref_mpi_set_grid(RefCase *c, const int in[12])
{
    c->grid.cart_comm_3d = 0;
    MPI_Type_contiguous(c->I.n_egroups, MPI_FLOAT, &c->grid.Flux_Array);   /* init.c:215-219 */
    c->grid.x_pos_src = in[0]; c->grid.x_pos_dest = in[1]; c->grid.x_neg_src = in[2]; c->grid.x_neg_dest = in[3];
    c->grid.y_pos_src = in[4]; c->grid.y_pos_dest = in[5]; c->grid.y_neg_src = in[6]; c->grid.y_neg_dest = in[7];
    c->grid.z_pos_src = in[8]; c->grid.z_pos_dest = in[9]; c->grid.z_neg_src = in[10]; c->grid.z_neg_dest = in[11];
}
/* the random stream is process-global in the reference: continue case c's stream where it stood */
void ref_mpi_select_stream(uint64_t seed, uint64_t calls)
{
    void ref_shim_set_calls(uint64_t c);
    ref_shim_reset(seed);
    ref_shim_set_calls(calls);
}
#endif

/* the host-side Params/Input themselves, for driving the drop-in C-ABI of the
 * product with the reference's own (pointer-rich) structures */
Params *ref_params(RefCase *c) { return &c->P; }
Input *ref_input_mut(RefCase *c) { return &c->I; }
