/* moc_host.c -- host side of libmoc_b200.so: configuration, command line and
 * synthetic problem construction (PART B3 of include/moc_b200.h).
 *
 * Mirrors the behaviour of the reference's init.c / io.c / tracks.c / source.c
 * (cited per function, paths relative to /root/reference/src) but is organised
 * around the counter-based random stream of include/moc_rng.h: every array is
 * filled from a closed-form stream position (moc_draw_layout), so the
 * construction can be evaluated in any order while consuming draws in exactly
 * the order the reference's serial code does.
 *
 * The big slabs come from moc_host_alloc() (pinned when a CUDA device is
 * present, so the drop-in path can DMA straight out of them).
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "moc_b200.h"
#include "moc_internal.h"
#include "moc_rng.h"

/* ------------------------------------------------------------------ errors */
static __thread char g_error[512];

void moc_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

const char *moc_last_error(void) { return g_error; }

/* ------------------------------------------------------------------ defaults */

/* init.c:33-74 -- the "default strawman" problem (~13 GB) */
Input moc_set_default_input(void)
{
    Input in;
    memset(&in, 0, sizeof in);
    in.x_assemblies = 17;
    in.y_assemblies = 17;
    in.cai = 27;
    in.fai = 5;
    in.axial_exp = 2;
    in.radial_ray_sep = 0.05;
    in.axial_z_sep = 0.25;
    in.n_azimuthal = 64;
    in.n_polar_angles = 10;
    in.n_egroups = 104;
    in.decompose = true;
    in.decomp_assemblies_ax = 20;
    in.segments_per_track = 120;
    in.assembly_width = 21.42;
    in.height = 400.0;
    in.precision = 0.01;
    in.mype = 0;
    in.n_2D_source_regions_per_assembly = 5000;
    in.nthreads = 1;
    in.load_tracks = false;
    in.track_file = NULL;
    return in;
}

/* init.c:77-103 -- the "-s" problem (~1 GB) */
void moc_set_small_input(Input *in)
{
    in->x_assemblies = 15;
    in->y_assemblies = 15;
    in->cai = 5;
    in->fai = 3;
    in->axial_exp = 2;
    in->radial_ray_sep = 0.5;
    in->axial_z_sep = 0.2;
    in->n_azimuthal = 5;
    in->n_polar_angles = 5;
    in->n_egroups = 104;
    in->decompose = false;
    in->decomp_assemblies_ax = 1;
    in->segments_per_track = 120;
    in->assembly_width = 1.26 * 17;
    in->height = 400.0;
    in->precision = 0.01;
    in->n_2D_source_regions_per_assembly = 3000;
}

/* init.c:4-30 */
void moc_calculate_derived_inputs(Input *in)
{
    in->n_azimuthal /= 2;   /* forward/backward tracking shares a 2D track */
    in->ntracks_2D = in->n_azimuthal * (in->assembly_width * sqrt(2) / in->radial_ray_sep);
    in->ntracks_2D = 2 * (in->ntracks_2D / 2);
    in->z_stacked = (int)(in->height / (in->axial_z_sep * in->decomp_assemblies_ax));
    in->ntracks = in->ntracks_2D * in->n_polar_angles * in->z_stacked;
    in->domain_height = in->height / in->decomp_assemblies_ax;
    in->n_source_regions_per_node =
        in->n_2D_source_regions_per_assembly * in->cai / in->decomp_assemblies_ax;
}

/* io.c:198-270 -- 18 values, each the leading token of a line; the rest of the line
 * is commentary.  Integers are read like "%d"/"%ld", reals like "%f". */
int moc_read_input_file(Input *in, const char *fname)
{
    enum { KI, KL, KF, KB };
    struct { int kind; void *dst; } slot[18] = {
        { KI, &in->x_assemblies }, { KI, &in->y_assemblies }, { KI, &in->cai },
        { KI, &in->fai }, { KI, &in->axial_exp }, { KF, &in->radial_ray_sep },
        { KF, &in->axial_z_sep }, { KI, &in->n_azimuthal }, { KI, &in->n_polar_angles },
        { KI, &in->n_egroups }, { KB, &in->decompose }, { KI, &in->decomp_assemblies_ax },
        { KL, &in->segments_per_track }, { KF, &in->assembly_width }, { KF, &in->height },
        { KF, &in->precision }, { KL, &in->n_2D_source_regions_per_assembly },
        { KI, &in->papi_event_set } };

    FILE *fp = fopen(fname, "r");
    if (!fp) {
        moc_set_error("cannot open input file '%s'", fname);
        return MOC_EINVAL;
    }
    char line[512];
    int got = 0;
    while (got < 18 && fgets(line, sizeof line, fp)) {
        char *p = line, *end = NULL;
        while (*p == ' ' || *p == '\t') p++;
        if (*p == '\n' || *p == '\r' || *p == '\0') continue;
        switch (slot[got].kind) {
        case KI: *(int *)slot[got].dst = (int)strtol(p, &end, 10); break;
        case KL: *(long *)slot[got].dst = strtol(p, &end, 10); break;
        case KF: *(float *)slot[got].dst = strtof(p, &end); break;
        case KB: *(bool *)slot[got].dst = strtol(p, &end, 10) != 0; break;
        }
        if (end == p) {
            fclose(fp);
            moc_set_error("input file '%s': line for value %d does not start with a number",
                          fname, got + 1);
            return MOC_EINVAL;
        }
        got++;
    }
    fclose(fp);
    if (got < 18) {
        moc_set_error("input file '%s': expected 18 values, found %d", fname, got);
        return MOC_EINVAL;
    }
    return MOC_OK;
}

/* io.c:115-181.  Options are applied in command-line order, as in the reference
 * (so "-i f -s" ends with the small set, "-s -i f" with the file's). */
int moc_read_CLI(int argc, char *argv[], Input *in)
{
    in->nthreads = 1;
    for (int a = 1; a < argc; a++) {
        const char *opt = argv[a];
        int has_value = (a + 1 < argc);
        if (strcmp(opt, "-t") == 0) {
            if (!has_value) goto usage;
            in->nthreads = atoi(argv[++a]);   /* accepted for compatibility; the GPU path ignores it */
        } else if (strcmp(opt, "-i") == 0) {
            if (!has_value) goto usage;
            int rc = moc_read_input_file(in, argv[++a]);
            if (rc) return rc;
        } else if (strcmp(opt, "-s") == 0) {
            moc_set_small_input(in);
        } else if (strcmp(opt, "-d") == 0) {
            if (!has_value) goto usage;
            in->track_file = argv[++a];
            in->load_tracks = true;
        } else if (strcmp(opt, "-p") == 0) {
            if (!has_value) goto usage;
            ++a;                              /* PAPI event name: CPU counters do not apply */
        } else {
            goto usage;
        }
    }
    if (in->nthreads < 1) goto usage;
    return MOC_OK;
usage:
    moc_set_error("usage: SimpleMOC-b200 [-t <threads>] [-i <input file>] [-s] [-d <OpenMOC track file>]");
    return MOC_EINVAL;
}

/* utils.c:97-143 -- the reference's own footprint estimate, same terms */
size_t moc_est_mem_usage(const Input *in)
{
    const size_t T2 = (size_t)in->ntracks_2D, T3 = (size_t)in->ntracks;
    const size_t P = (size_t)in->n_polar_angles, G = (size_t)in->n_egroups;
    const size_t N = (size_t)in->n_source_regions_per_node, F = (size_t)in->fai;
    const size_t X = N / 8;
    const size_t Z = (size_t)(int)(in->height / (in->axial_z_sep * in->decomp_assemblies_ax));
    const size_t nthr = (size_t)in->nthreads, Zs = (size_t)in->z_stacked;
    const size_t spt = (size_t)in->segments_per_track;
    size_t b = 0;
    b += T2 * sizeof(Track2D) + spt * T2 * sizeof(Segment);
    b += T2 * sizeof(Track **) + T2 * P * sizeof(Track *) + T3 * sizeof(Track);
    b += T2 * P * Z * G * sizeof(float) * 2;                       /* angular flux */
    b += N * sizeof(Source);
    b += 3 * X * sizeof(float **) + X * G * G * sizeof(float);
    b += X * G * sizeof(float *) + X * G * 3 * sizeof(float);
    b += 2 * (N * sizeof(float **) + N * F * sizeof(float *));
    b += N * F * G * sizeof(float);
    /* per-thread two-way tracking scratch the reference still counts */
    b += nthr * Zs * (sizeof(double *) + sizeof(Source **) + 2 * sizeof(int));
    b += nthr * Zs * 2 * spt * (sizeof(double) + sizeof(Source *));
    return b;
}

/* utils.c:147-155 */
double moc_time_per_intersection(const Input *in, double seconds)
{
    return seconds / (double)in->segments_processed * 1.0e9 / (double)in->n_egroups;
}

/* ------------------------------------------------------------------ boundary exchange schedule */

/* comms.c:12-28,75-83: tracks per face from the surface-area ratio, in whole messages of
 * 10000 tracks; then the (round, direction) order of comms.c:100-183. */
long moc_exchange_plan(const Input *in, const CommGrid *grid, moc_exchange_op *ops, long max_ops)
{
    const int tracks_per_msg = 10000;
    const float hgt = in->domain_height;
    const float x = in->assembly_width;
    long per_axial = in->ntracks * x / (2 * x + 4 * hgt);
    long per_radial = in->ntracks * hgt / (2 * x + 4 * hgt);
    const long remaining = in->ntracks - 2 * per_axial - 4 * per_radial;
    long add_radial = remaining * (4 * hgt / (2 * x + 4 * hgt));
    add_radial = 4 * (add_radial / 4);
    per_radial += add_radial / 4;
    const long add_axial = remaining - add_radial;
    per_axial += add_axial / 2;
    long nmsg[6], rounds = 0, total = 0;
    for (int d = 0; d < 4; d++) nmsg[d] = per_radial / tracks_per_msg;
    for (int d = 4; d < 6; d++) nmsg[d] = per_axial / tracks_per_msg;
    for (int d = 0; d < 6; d++) {
        if (nmsg[d] > rounds) rounds = nmsg[d];
        total += nmsg[d];
    }
    const long long chunk = (long long)in->n_egroups * tracks_per_msg;
    if (total * chunk > 2ll * in->ntracks * in->n_egroups) {
        moc_set_error("exchange plan (%ld messages of %lld floats) exceeds the flux slab", total, chunk);
        return MOC_EINVAL;
    }
    if (!ops) return total;
    const int dest[6] = { grid->x_pos_dest, grid->x_neg_dest, grid->y_pos_dest,
                          grid->y_neg_dest, grid->z_pos_dest, grid->z_neg_dest };
    const int from[6] = { grid->x_pos_src, grid->x_neg_src, grid->y_pos_src,
                          grid->y_neg_src, grid->z_pos_src, grid->z_neg_src };
    long k = 0;
    long long at = 0;
    for (long i = 0; i < rounds; i++)
        for (int d = 0; d < 6; d++) {
            if (i >= nmsg[d]) continue;
            if (k < max_ops) {
                ops[k].offset = at;
                ops[k].count = chunk;
                ops[k].round = (int)i;
                ops[k].direction = d;
                ops[k].send_to = dest[d];
                ops[k].recv_from = from[d];
            }
            k++;
            at += chunk;
        }
    return k;
}

/* ------------------------------------------------------------------ construction */

/* utils.c:11-26 with explicit stream positions */
static float normal_draw(uint64_t seed, uint64_t at, float mean, float sigma)
{
    float u_radius = moc_urand(seed, at);
    float u_angle = moc_urand(seed, at + 1);
    float x = sqrt(-2 * log(u_radius)) * cos(2 * M_PI * u_angle);
    return x * sigma + mean;
}

/* utils.c:48-78 */
static int make_exp_table(Table *tab, float precision, float maxVal)
{
    int cells = (int)(maxVal * sqrt(1.0 / (8.0 * precision * 0.01)));
    float dx = maxVal / (float)cells;
    float *v = (float *)malloc(sizeof(float) * 2 * (size_t)cells);
    if (!v) return MOC_ENOMEM;
    for (int n = 0; n < cells; n++) {
        float e = exp(-n * dx);
        v[2 * n] = -e;                       /* slope, as the reference stores it */
        v[2 * n + 1] = 1 + (n * dx - 1) * e; /* intercept */
    }
    tab->values = v;
    tab->dx = dx;
    tab->maxVal = maxVal - dx;
    tab->N = cells;
    return MOC_OK;
}

/* ---- OpenMOC track files (tracks.c:170-323, the -d option) ----
 * On-disk layout, as that reader consumes it (native endianness, no padding):
 *   int    string_length;  char geometry[string_length];
 *   int    n_azimuthal;    double spacing;
 *   int    num_tracks[n_azimuthal], num_x[n_azimuthal], num_y[n_azimuthal];
 *   double azim_weights[n_azimuthal];
 *   per track (azimuthal angle major):  double x0, y0, x1, y1, phi;  int azim_angle_index;
 *                                       int num_segments;
 *     per segment:  double length;  int material_id;  int region_id;  (+ 2 ints if cmfd)
 * What the reference keeps: n_azimuthal, (float)spacing -> radial_ray_sep, per track the segment
 * count, per segment (float)length and region_id -> source_id; it draws az_weight = urand() per
 * track (tracks.c:283), recomputes segments_per_track = total / ntracks_2D (integer division,
 * tracks.c:311) and ntracks.  Everything else in the file is skipped.  Unlike the reference, a
 * missing, truncated or implausible file is an error (MOC_EIO), not undefined behaviour. */
typedef struct {
    FILE *f;
    const char *name;
    int failed;
} track_reader;

static void rd(track_reader *r, void *dst, size_t size, size_t count)
{
    if (r->failed || count == 0) return;
    if (fread(dst, size, count, r->f) != count) r->failed = 1;
}

static void rd_skip(track_reader *r, long bytes)
{
    if (r->failed) return;
    if (fseek(r->f, bytes, SEEK_CUR) != 0) r->failed = 1;
}

int moc_load_openmoc_tracks(const char *fname, int cmfd, Input *in, uint64_t seed, uint64_t az_weight_at,
                            Track2D **tracks_out, long *total_segments_out)
{
    if (!fname || !in || !tracks_out) {
        moc_set_error("moc_load_openmoc_tracks: null argument");
        return MOC_EINVAL;
    }
    track_reader r = { fopen(fname, "rb"), fname, 0 };
    if (!r.f) {
        moc_set_error("cannot open track file '%s'", fname);
        return MOC_EIO;
    }
    fseek(r.f, 0, SEEK_END);
    const long file_bytes = ftell(r.f);
    fseek(r.f, 0, SEEK_SET);
    int rc = MOC_EIO;
    Track2D *t2 = NULL;
    Segment *segs = NULL;
    int *per_angle = NULL;

    int string_length = 0, n_azim = 0;
    double spacing = 0;
    rd(&r, &string_length, sizeof(int), 1);
    if (r.failed || string_length < 0 || string_length > file_bytes) {
        moc_set_error("track file '%s': bad geometry-string length %d", fname, string_length);
        goto done;
    }
    rd_skip(&r, string_length);
    rd(&r, &n_azim, sizeof(int), 1);
    rd(&r, &spacing, sizeof(double), 1);
    if (r.failed || n_azim <= 0 || (long)n_azim * 20 > file_bytes) {
        moc_set_error("track file '%s': bad number of azimuthal angles %d", fname, n_azim);
        goto done;
    }
    per_angle = (int *)malloc(sizeof(int) * (size_t)n_azim);
    if (!per_angle) { rc = MOC_ENOMEM; goto done; }
    rd(&r, per_angle, sizeof(int), (size_t)n_azim);
    rd_skip(&r, (long)n_azim * (2 * (long)sizeof(int) + (long)sizeof(double)));   /* num_x, num_y, azim_weights */
    long T2 = 0;
    for (int a = 0; a < n_azim && !r.failed; a++) {
        if (per_angle[a] < 0) r.failed = 1;
        T2 += per_angle[a];
    }
    /* a track takes at least 5 doubles + 2 ints on disk */
    if (r.failed || T2 <= 0 || T2 * 48 > file_bytes) {
        moc_set_error("track file '%s': bad track counts (%ld tracks in a %ld-byte file)", fname, T2, file_bytes);
        goto done;
    }

    /* pass 1: segment total (tracks.c:225-250) */
    const long body = ftell(r.f);
    const long seg_bytes = (long)sizeof(double) + 2 * (long)sizeof(int) + (cmfd ? 2 * (long)sizeof(int) : 0);
    long total = 0;
    for (long u = 0; u < T2 && !r.failed; u++) {
        int n = 0;
        rd_skip(&r, 5 * (long)sizeof(double) + (long)sizeof(int));
        rd(&r, &n, sizeof(int), 1);
        if (n < 0 || (long)n * seg_bytes > file_bytes) r.failed = 1;
        total += n;
        rd_skip(&r, (long)n * seg_bytes);
    }
    if (r.failed || ftell(r.f) > file_bytes) {
        moc_set_error("track file '%s' is truncated or corrupt (segment counts run past its %ld bytes)", fname, file_bytes);
        goto done;
    }

    /* pass 2: the data (tracks.c:259-308) */
    t2 = (Track2D *)calloc((size_t)T2, sizeof(Track2D));
    segs = (Segment *)calloc((size_t)(total > 0 ? total : 1), sizeof(Segment));
    if (!t2 || !segs) { rc = MOC_ENOMEM; goto done; }
    fseek(r.f, body, SEEK_SET);
    long first = 0;
    for (long u = 0; u < T2 && !r.failed; u++) {
        int n = 0;
        rd_skip(&r, 5 * (long)sizeof(double) + (long)sizeof(int));
        rd(&r, &n, sizeof(int), 1);
        t2[u].n_segments = n;
        t2[u].segments = segs + first;
        t2[u].az_weight = moc_urand(seed, az_weight_at + (uint64_t)u);     /* tracks.c:283 */
        for (int sgm = 0; sgm < n && !r.failed; sgm++) {
            double length = 0;
            int ids[2] = { 0, 0 };                                       /* material_id, region_id */
            rd(&r, &length, sizeof(double), 1);
            rd(&r, ids, sizeof(int), 2);
            if (cmfd) rd_skip(&r, 2 * (long)sizeof(int));
            segs[first + sgm].length = (float)length;
            segs[first + sgm].source_id = (long)ids[1];
        }
        first += n;
    }
    if (r.failed) {
        moc_set_error("track file '%s': read error in the track data", fname);
        goto done;
    }
    in->n_azimuthal = n_azim;
    in->radial_ray_sep = (float)spacing;
    in->ntracks_2D = T2;
    in->segments_per_track = total / T2;
    in->ntracks = T2 * in->n_polar_angles * in->z_stacked;
    *tracks_out = t2;
    if (total_segments_out) *total_segments_out = total;
    t2 = NULL;
    segs = NULL;
    rc = MOC_OK;
done:
    free(per_angle);
    free(t2);
    free(segs);
    fclose(r.f);
    return rc;
}

/* 2D tracks of build_tracks() (init.c:119-125): from the track file when -d was given, else the
 * synthetic ones of tracks.c:4-58.  Fills the first three stream positions of *at and at->p_weight. */
static int make_tracks_2d(Input *in, uint64_t seed, moc_draw_layout *at, Track2D **out)
{
    at->az_weight = 0;
    if (in->load_tracks) {
        long total = 0;
        int rc = moc_load_openmoc_tracks(in->track_file, 0, in, seed, at->az_weight, out, &total);
        if (rc) return rc;
        /* one draw per track and nothing else (tracks.c:283) */
        at->n_segments = at->seg_length = at->p_weight = at->az_weight + (uint64_t)in->ntracks_2D;
        return MOC_OK;
    }
    const long T2 = in->ntracks_2D;
    at->n_segments = at->az_weight + (uint64_t)T2;
    at->seg_length = at->n_segments + 2 * (uint64_t)T2;
    Track2D *t2 = (Track2D *)calloc((size_t)T2, sizeof(Track2D));
    if (!t2) return MOC_ENOMEM;
    long total_segments = 0;
    for (long i = 0; i < T2; i++) {
        t2[i].az_weight = moc_urand(seed, at->az_weight + (uint64_t)i);
        t2[i].n_segments = normal_draw(seed, at->n_segments + 2 * (uint64_t)i,
                                       in->segments_per_track, sqrt(in->segments_per_track));
        if (t2[i].n_segments < 0) t2[i].n_segments = 0;   /* cannot index a negative count */
        total_segments += t2[i].n_segments;
    }
    Segment *segs = (Segment *)calloc((size_t)(total_segments > 0 ? total_segments : 1), sizeof(Segment));
    if (!segs) {
        free(t2);
        return MOC_ENOMEM;
    }
    long first = 0;
    for (long i = 0; i < T2; i++) {
        t2[i].segments = segs + first;
        for (long n = 0; n < t2[i].n_segments; n++)
            segs[first + n].length = moc_urand(seed, at->seg_length + (uint64_t)(first + n))
                                     * in->assembly_width / t2[i].n_segments;
        first += t2[i].n_segments;
    }
    at->p_weight = at->seg_length + (uint64_t)total_segments;
    *out = t2;
    return MOC_OK;
}

static int check_sizes(const Input *in)
{
    if ((!in->load_tracks && in->ntracks_2D <= 0) || in->z_stacked <= 0 || in->n_polar_angles <= 0 ||
        in->n_egroups <= 0 || in->fai <= 0 || in->n_source_regions_per_node < 8) {
        moc_set_error("degenerate problem: T2=%ld Z=%d P=%d G=%d fai=%d N=%ld (need N >= 8)",
                      in->ntracks_2D, in->z_stacked, in->n_polar_angles, in->n_egroups,
                      in->fai, in->n_source_regions_per_node);
        return MOC_EINVAL;
    }
    return MOC_OK;
}

/* The cheap, libm-dependent part of build_tracks(): 2D tracks (tracks.c:4-58), polar angles
 * (tracks.c:159-168), the exponential table (utils.c:48-78), the leakage cell -- plus the stream
 * positions of every later block of draws.  Params.tracks and Params.sources stay NULL: the
 * device fills those arrays itself (moc_create_synthetic). */
int moc_build_tracks_2d(Input *in, uint64_t seed, Params *out, moc_draw_layout *layout)
{
    int rc0 = check_sizes(in);
    if (rc0) return rc0;
    memset(out, 0, sizeof *out);
    moc_draw_layout at;
    Track2D *t2 = NULL;
    if ((rc0 = make_tracks_2d(in, seed, &at, &t2))) return rc0;
    const long T3 = in->ntracks, N = in->n_source_regions_per_node;
    const long X = N / 8;
    const int P = in->n_polar_angles, G = in->n_egroups, F = in->fai;
    out->tracks_2D = t2;
    float *polar = (float *)malloc(sizeof(float) * (size_t)P);
    if (!polar) return MOC_ENOMEM;
    for (int j = 0; j < P; j++) polar[j] = M_PI * (j + 0.5) / P;
    out->polar_angles = polar;
    at.scatter = at.p_weight + (uint64_t)T3;
    at.xs = at.scatter + (uint64_t)X * G * G;
    at.fine_source = at.xs + (uint64_t)X * G * 3;
    at.sigT = at.fine_source + (uint64_t)N * F * G;
    at.regions = at.sigT + (uint64_t)N * G;
    at.end = at.regions + 2 * (uint64_t)N - 1;
    out->leakage = (float *)calloc(1, sizeof(float));
    int rc = make_exp_table(&out->expTable, in->precision, 10.0);
    if (rc) return rc;
    if (layout) *layout = at;
    return MOC_OK;
}

/* init.c:106-159 = tracks.c:4-58 + tracks.c:75-168 + source.c:4-214 + utils.c:48-78 */
int moc_build_tracks(Input *in, uint64_t seed, Params *out, uint64_t *rand_calls)
{
    int rc0 = check_sizes(in);
    if (rc0) return rc0;
    memset(out, 0, sizeof *out);

    /* ---- 2D tracks (the track file may change ntracks_2D, ntracks, n_azimuthal, ...) ---- */
    moc_draw_layout at;
    Track2D *t2 = NULL;
    if ((rc0 = make_tracks_2d(in, seed, &at, &t2))) return rc0;
    const long T2 = in->ntracks_2D, T3 = in->ntracks, N = in->n_source_regions_per_node;
    const long X = N / 8;                                   /* source.c:12 */
    const int P = in->n_polar_angles, Z = in->z_stacked, G = in->n_egroups, F = in->fai;
    out->tracks_2D = t2;

    /* ---- 3D tracks: [i][j][k] views over one Track array and one flux slab ---- */
    Track ***by_i = (Track ***)malloc(sizeof(Track **) * (size_t)T2);
    Track **by_ij = (Track **)malloc(sizeof(Track *) * (size_t)T2 * P);
    Track *trk = (Track *)moc_host_alloc(sizeof(Track) * (size_t)T3);
    float *flux = (float *)moc_host_alloc(sizeof(float) * 2 * (size_t)T3 * G);   /* zero-filled */
    if (!by_i || !by_ij || !trk || !flux) return MOC_ENOMEM;
    for (long i = 0; i < T2; i++) {
        by_i[i] = by_ij + i * P;
        for (int j = 0; j < P; j++) by_i[i][j] = trk + (i * P + j) * Z;
    }
    for (long t = 0; t < T3; t++) {
        const int k = (int)(t % Z), j = (int)((t / Z) % P);
        /* upward rays start at the bottom of their slot, downward ones at the top */
        trk[t].z_height = (j < P / 2) ? in->axial_z_sep * k : in->axial_z_sep * (k + 1);
        trk[t].p_weight = moc_urand(seed, at.p_weight + (uint64_t)t);
        trk[t].f_psi = flux + 2 * t * G;
        trk[t].b_psi = flux + (2 * t + 1) * G;
    }
    out->tracks = by_i;

    float *polar = (float *)malloc(sizeof(float) * (size_t)P);
    for (int j = 0; j < P; j++) polar[j] = M_PI * (j + 0.5) / P;       /* tracks.c:159-168 */
    out->polar_angles = polar;

    /* ---- sources ---- */
    at.scatter = at.p_weight + (uint64_t)T3;
    at.xs = at.scatter + (uint64_t)X * G * G;
    at.fine_source = at.xs + (uint64_t)X * G * 3;
    at.sigT = at.fine_source + (uint64_t)N * F * G;
    at.regions = at.sigT + (uint64_t)N * G;
    at.end = at.regions + 2 * (uint64_t)N - 1;

    float *scat = (float *)moc_host_alloc(sizeof(float) * (size_t)X * G * G);
    float *xs = (float *)moc_host_alloc(sizeof(float) * (size_t)X * G * 3);
    /* one slab: fine_source[N][fai][G] | fine_flux[N][fai][G] | sigT[N][G]  (source.c:121-152) */
    float *slab = (float *)moc_host_alloc(sizeof(float) * (size_t)(2 * F + 1) * N * G);
    float **scat_rows = (float **)malloc(sizeof(float *) * (size_t)X * G);
    float **xs_rows = (float **)malloc(sizeof(float *) * (size_t)X * G);
    float **src_rows = (float **)malloc(sizeof(float *) * (size_t)N * F);
    float **flux_rows = (float **)malloc(sizeof(float *) * (size_t)N * F);
    Source *regions = (Source *)calloc((size_t)N, sizeof(Source));
    if (!scat || !xs || !slab || !scat_rows || !xs_rows || !src_rows || !flux_rows || !regions)
        return MOC_ENOMEM;

    for (long e = 0; e < X * G * G; e++) scat[e] = moc_urand(seed, at.scatter + (uint64_t)e);
    for (long e = 0; e < X * G * 3; e++) xs[e] = moc_urand(seed, at.xs + (uint64_t)e);
    for (long e = 0; e < N * F * G; e++) slab[e] = moc_urand(seed, at.fine_source + (uint64_t)e);
    float *sigT = slab + 2 * N * F * G;
    for (long e = 0; e < N * G; e++) sigT[e] = moc_urand(seed, at.sigT + (uint64_t)e);
    for (long r = 0; r < X * G; r++) {
        scat_rows[r] = scat + r * G;
        xs_rows[r] = xs + r * 3;
    }
    for (long r = 0; r < N * F; r++) {
        src_rows[r] = slab + r * G;
        flux_rows[r] = slab + (N * F + r) * G;
    }
    for (long i = 0; i < N; i++) {
        /* region 0 takes material 0 without a draw; region i>0 draws (index, volume) */
        long material = 0;
        uint64_t vol_at = at.regions;
        if (i > 0) {
            material = (long)moc_rand31(seed, at.regions + 2 * (uint64_t)i - 1) % X;
            vol_at = at.regions + 2 * (uint64_t)i;
        }
        regions[i].scattering_matrix = scat_rows + material * G;
        regions[i].XS = xs_rows + material * G;
        regions[i].fine_flux = flux_rows + i * F;
        regions[i].fine_source = src_rows + i * F;
        regions[i].sigT = sigT + i * G;
        regions[i].vol = moc_urand(seed, vol_at);
    }
    out->sources = regions;

    out->leakage = (float *)calloc(1, sizeof(float));
    int rc = make_exp_table(&out->expTable, in->precision, 10.0);
    if (rc) return rc;
    if (rand_calls) *rand_calls = at.end;
    return MOC_OK;
}

void moc_free_tracks(const Input *in, Params *p)
{
    (void)in;
    if (!p) return;
    if (p->tracks_2D) {
        free(p->tracks_2D[0].segments);
        free(p->tracks_2D);
    }
    if (p->tracks) {
        Track *trk = p->tracks[0][0];
        moc_host_free(trk[0].f_psi);
        moc_host_free(trk);
        free(p->tracks[0]);
        free(p->tracks);
    }
    if (p->sources) {
        Source *s0 = &p->sources[0];
        moc_host_free(s0->scattering_matrix[0]);
        free(s0->scattering_matrix);
        moc_host_free(s0->XS[0]);
        free(s0->XS);
        moc_host_free(s0->fine_source[0]);
        free(s0->fine_source);
        free(s0->fine_flux);
        free(p->sources);
    }
    free(p->polar_angles);
    free(p->leakage);
    free(p->expTable.values);
    memset(p, 0, sizeof *p);
}

/* ------------------------------------------------------------------ flat views */

static long params_copy(const Input *in, const Params *p, int which, void *dst, const void *src,
                        size_t bytes)
{
    const long T2 = in->ntracks_2D, T3 = in->ntracks, N = in->n_source_regions_per_node, X = N / 8;
    const int P = in->n_polar_angles, G = in->n_egroups, F = in->fai;
    const Source *s0 = &p->sources[0];
    Track *trk = p->tracks[0][0];
    void *flat = NULL;      /* contiguous arrays: plain memcpy */
    size_t n = 0;
    switch (which) {
    case MOC_ARR_FINE_SOURCE: flat = s0->fine_source[0]; n = sizeof(float) * N * F * G; break;
    case MOC_ARR_FINE_FLUX: flat = s0->fine_flux[0]; n = sizeof(float) * N * F * G; break;
    case MOC_ARR_SIGT: flat = s0->sigT; n = sizeof(float) * N * G; break;
    case MOC_ARR_PSI: flat = trk[0].f_psi; n = sizeof(float) * 2 * T3 * G; break;
    case MOC_HOST_XS: flat = s0->XS[0]; n = sizeof(float) * X * G * 3; break;
    case MOC_HOST_SCATTER: flat = s0->scattering_matrix[0]; n = sizeof(float) * X * G * G; break;
    case MOC_HOST_POLAR: flat = p->polar_angles; n = sizeof(float) * P; break;
    case MOC_HOST_TABLE: flat = p->expTable.values; n = sizeof(float) * 2 * p->expTable.N; break;
    case MOC_ARR_Z_HEIGHT:
    case MOC_ARR_P_WEIGHT: n = sizeof(float) * T3; break;
    case MOC_HOST_AZ_WEIGHT: n = sizeof(float) * T2; break;
    case MOC_HOST_N_SEGMENTS: n = sizeof(long) * T2; break;
    case MOC_HOST_SEG_LENGTHS: {
        long tot = 0;
        for (long i = 0; i < T2; i++) tot += p->tracks_2D[i].n_segments;
        n = sizeof(float) * tot;
        break;
    }
    case MOC_HOST_XS_INDEX: n = sizeof(int) * N; break;
    case MOC_HOST_VOL: n = sizeof(float) * N; break;
    default:
        moc_set_error("moc_params_get/set: unknown array id %d", which);
        return MOC_EINVAL;
    }
    if ((!dst && !src) || bytes != n) return (long)n;
    if (flat) {
        if (dst) memcpy(dst, flat, n);
        else memcpy(flat, src, n);
        return (long)n;
    }
    /* strided members of the AoS structures */
    switch (which) {
    case MOC_ARR_Z_HEIGHT:
        for (long t = 0; t < T3; t++) {
            if (dst) ((float *)dst)[t] = trk[t].z_height;
            else trk[t].z_height = ((const float *)src)[t];
        }
        break;
    case MOC_ARR_P_WEIGHT:
        for (long t = 0; t < T3; t++) {
            if (dst) ((float *)dst)[t] = trk[t].p_weight;
            else trk[t].p_weight = ((const float *)src)[t];
        }
        break;
    case MOC_HOST_AZ_WEIGHT:
        for (long i = 0; i < T2; i++) {
            if (dst) ((float *)dst)[i] = p->tracks_2D[i].az_weight;
            else p->tracks_2D[i].az_weight = ((const float *)src)[i];
        }
        break;
    case MOC_HOST_N_SEGMENTS:
        if (!dst) { moc_set_error("n_segments is read-only"); return MOC_EINVAL; }
        for (long i = 0; i < T2; i++) ((long *)dst)[i] = p->tracks_2D[i].n_segments;
        break;
    case MOC_HOST_SEG_LENGTHS: {
        long e = 0;
        for (long i = 0; i < T2; i++)
            for (long k = 0; k < p->tracks_2D[i].n_segments; k++, e++) {
                if (dst) ((float *)dst)[e] = p->tracks_2D[i].segments[k].length;
                else p->tracks_2D[i].segments[k].length = ((const float *)src)[e];
            }
        break;
    }
    case MOC_HOST_XS_INDEX:
        if (!dst) { moc_set_error("xs_index is read-only"); return MOC_EINVAL; }
        for (long i = 0; i < N; i++)
            ((int *)dst)[i] = (int)((p->sources[i].XS[0] - s0->XS[0]) / (3 * G));
        break;
    case MOC_HOST_VOL:
        for (long i = 0; i < N; i++) {
            if (dst) ((float *)dst)[i] = p->sources[i].vol;
            else p->sources[i].vol = ((const float *)src)[i];
        }
        break;
    }
    return (long)n;
}

long moc_params_get(const Input *in, const Params *p, int which, void *dst, size_t bytes)
{
    return params_copy(in, p, which, dst, NULL, bytes);
}

long moc_params_set(const Input *in, Params *p, int which, const void *src, size_t bytes)
{
    return params_copy(in, p, which, NULL, src, bytes);
}
