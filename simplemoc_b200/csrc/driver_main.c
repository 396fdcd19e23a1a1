/* driver_main.c -- SimpleMOC-b200: the C driver of the B200 transport sweep.
 *
 * Keeps the interface of the reference's main() (reference src/main.c:3-147): the same command
 * line (-t, -i <file>, -s, -p, -d; src/io.c:115-181), the same Input/Params structures, the same
 * iteration loop
 *      transport_sweep -> fast_transfer_boundary_fluxes -> renormalize_flux
 *                      -> update_sources -> compute_keff
 * called under the reference's own function names, and the same report (phase timers and "Time
 * per Intersection", src/utils.c:147-155).  The five functions are the drop-in entry points of
 * libmoc_b200.so (include/moc_b200.h PART B1): every one of them runs on the GPU; without a CUDA
 * device the program stops with an error (there is no CPU path).
 *
 * New long options (none of them changes what the old ones mean):
 *   --seed N          counter-RNG seed (the reference seeds from time(NULL), src/main.c:20)
 *   --exp table|sfu   exponential: the reference's table (default) or MUFU.EX2
 *   --iters N         iterations of the loop (the reference hard-codes 1, src/main.c:41); N > 1 also
 *                     prints the source residual update_sources returns (dropped by main.c:81)
 *   --host-buffers    strict drop-in mode: every phase uploads from / downloads to the host
 *                     structures; default keeps the problem resident in HBM between phases and
 *                     synchronises the host structures once at the end
 *   --grid cx,cy,cz --rank R --id-file PATH
 *                     one process per GPU / spatial domain (the reference hard-codes a 2x2x1 MPI
 *                     grid, src/init.c:169); rank 0 writes the NCCL id to PATH, the others read it
 *   --device D        CUDA device (default: rank % device count)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "moc_b200.h"

static double now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static void rule(void) { puts("==============================================================================="); }
static void title(const char *s)
{
    int pad = (79 - (int)strlen(s)) / 2;
    printf("%*s%s\n", pad > 0 ? pad : 0, "", s);
}

static void die(const char *what)
{
    fprintf(stderr, "SimpleMOC-b200: %s: %s\n", what, moc_last_error());
    exit(1);
}

int main(int argc, char *argv[])
{
    unsigned long long seed = 1;
    int exp_mode = 0, iters = 1, host_buffers = 0, device = -1;
    int cx = 1, cy = 1, cz = 1, rank = 0;
    const char *id_file = NULL;

    /* split our long options from the reference's short ones */
    char **ref_argv = (char **)calloc((size_t)argc + 1, sizeof(char *));
    int ref_argc = 0;
    ref_argv[ref_argc++] = argv[0];
    for (int a = 1; a < argc; a++) {
        const char *o = argv[a];
        const char *v = (a + 1 < argc) ? argv[a + 1] : NULL;
        if (!strcmp(o, "--seed") && v) { seed = strtoull(v, NULL, 10); a++; }
        else if (!strcmp(o, "--exp") && v) { exp_mode = !strcmp(v, "sfu"); a++; }
        else if (!strcmp(o, "--iters") && v) { iters = atoi(v); a++; }
        else if (!strcmp(o, "--host-buffers")) host_buffers = 1;
        else if (!strcmp(o, "--grid") && v) { if (sscanf(v, "%d,%d,%d", &cx, &cy, &cz) != 3) { fprintf(stderr, "bad --grid\n"); return 1; } a++; }
        else if (!strcmp(o, "--rank") && v) { rank = atoi(v); a++; }
        else if (!strcmp(o, "--id-file") && v) { id_file = v; a++; }
        else if (!strcmp(o, "--device") && v) { device = atoi(v); a++; }
        else ref_argv[ref_argc++] = argv[a];
    }
    const int nranks = cx * cy * cz;

    Input input = moc_set_default_input();
    if (moc_read_CLI(ref_argc, ref_argv, &input)) die("command line");
    input.mype = rank;
    moc_calculate_derived_inputs(&input);

    if (moc_device_count() <= 0) {
        fprintf(stderr, "SimpleMOC-b200: no CUDA device: this program has no CPU path\n");
        return 1;
    }
    if (device < 0) device = rank % moc_device_count();
    if (moc_set_device(device)) die("device");

    if (rank == 0) {
        rule();
        title("SimpleMOC-b200 : 3D MOC transport sweep on NVIDIA B200 (sm_100a)");
        rule();
        if (input.load_tracks) printf("Reading track data from:\n%s\n", input.track_file);
    }

    Params params;
    unsigned long long draws = 0;
    double t0 = now();
    if (moc_build_tracks(&input, seed + (unsigned long long)rank, &params, (uint64_t *)&draws)) die("build_tracks");
    if (rank == 0) printf("Problem construction (host):          %6.2f sec\n", now() - t0);

    if (rank == 0) {
        /* after build_tracks, as main.c:32-36 does: a track file (-d) changes the track counts */
        title("INPUT SUMMARY");
        rule();
        printf("%-38s%d x %d x %d (this is rank %d)\n", "Spatial domains (GPUs):", cx, cy, cz, rank);
        printf("%-38s%d / %d\n", "Coarse / fine axial intervals:", input.cai, input.fai);
        printf("%-38s%d\n", "Axial source expansion order:", input.axial_exp);
        printf("%-38s%d (half-space) x %d\n", "Azimuthal x polar angles:", input.n_azimuthal, input.n_polar_angles);
        printf("%-38s%d\n", "Energy groups:", input.n_egroups);
        printf("%-38s%ld\n", "2D tracks:", input.ntracks_2D);
        printf("%-38s%d\n", "z-stacked rays per 2D track:", input.z_stacked);
        printf("%-38s%ld\n", "3D tracks:", input.ntracks);
        printf("%-38s%ld\n", "Source regions per domain:", input.n_source_regions_per_node);
        printf("%-38s%.2f MB\n", "Estimated memory (reference formula):", (double)moc_est_mem_usage(&input) / 1024.0 / 1024.0);
        printf("%-38s%s\n", "Exponential:", exp_mode ? "MUFU.EX2 (SFU)" : "reference table");
        printf("%-38s%llu\n", "Random stream seed:", seed);
        rule();
    }

    CommGrid grid;
    if (moc_make_grid(cx, cy, cz, rank, &grid)) die("grid");
    moc_dropin_configure(seed + (unsigned long long)rank, draws, exp_mode, 48);
    moc_set_resident(!host_buffers);

    float res = 0.f, keff = 1.0f;
    double t_sweep = 0, t_exch = 0, t_renorm = 0, t_update = 0, t_keff = 0, a, b;
    long segments = 0;

    if (rank == 0) { title("SIMULATION"); rule(); }
    for (int it = 0; it < iters; it++) {
        a = now();
        transport_sweep(&params, &input);
        b = now();
        t_sweep += b - a;
        segments += input.segments_processed;
        if (it == 0 && nranks > 1) {
            /* the communicator needs the device mirror, which the first call above created */
            char id[128];
            if (!id_file) { fprintf(stderr, "--grid with more than one domain needs --id-file\n"); return 1; }
            if (rank == 0) {
                if (moc_comm_get_unique_id(id)) die("nccl id");
                char tmp[4096];
                snprintf(tmp, sizeof tmp, "%s.tmp", id_file);
                FILE *f = fopen(tmp, "wb");
                if (!f || fwrite(id, 1, 128, f) != 128) { perror("id file"); return 1; }
                fclose(f);
                rename(tmp, id_file);
            } else {
                FILE *f = NULL;
                for (int tries = 0; tries < 6000 && !(f = fopen(id_file, "rb")); tries++) usleep(10000);
                if (!f || fread(id, 1, 128, f) != 128) { fprintf(stderr, "cannot read %s\n", id_file); return 1; }
                fclose(f);
            }
            if (moc_comm_init(moc_handle_of(&params), nranks, rank, id)) die("nccl init");
            /* from the next iteration on the exchange runs under the sweep of the interior stacks */
            if (!host_buffers) moc_dropin_set_grid(&grid);
        }
        if (nranks > 1) {
            a = now();
            fast_transfer_boundary_fluxes(params, input, grid);
            b = now();
            t_exch += b - a;
        }
        a = now();
        renormalize_flux(params, input, grid);
        b = now();
        t_renorm += b - a;
        a = now();
        res = update_sources(params, input, keff);
        b = now();
        t_update += b - a;
        a = now();
        keff = compute_keff(params, input, grid);
        b = now();
        t_keff += b - a;
        if (rank == 0) printf("keff = %f\n", keff);
        /* the reference computes the source residual and drops it (main.c:81); with more than the
         * reference's single iteration it is what tells whether the source iteration converges */
        if (rank == 0 && iters > 1) printf("iteration %d: source residual = %.6e\n", it + 1, res);
    }
    if (!host_buffers && moc_sync_to_host(&params)) die("sync to host");

    const double total = t_sweep + t_exch + t_renorm + t_update + t_keff;
    if (rank == 0) {
        rule();
        title("RESULTS SUMMARY");
        rule();
        printf("Transport Sweep Time:         %9.4lf sec   (%4.1lf%%)\n", t_sweep, 100 * t_sweep / total);
        printf("Domain Flux Exchange Time:    %9.4lf sec   (%4.1lf%%)\n", t_exch, 100 * t_exch / total);
        printf("Flux Renormalization Time:    %9.4lf sec   (%4.1lf%%)\n", t_renorm, 100 * t_renorm / total);
        printf("Update Source Time:           %9.4lf sec   (%4.1lf%%)\n", t_update, 100 * t_update / total);
        printf("K-Effective Calc Time:        %9.4lf sec   (%4.1lf%%)\n", t_keff, 100 * t_keff / total);
        printf("Total Time:                   %9.4lf sec\n", total);
        input.segments_processed = segments;
        printf("Segments processed:           %ld\n", segments);
        printf("Time per Intersection:          %.5lf ns\n", moc_time_per_intersection(&input, t_sweep));
        printf("Integrations per second:        %.4g\n", (double)segments * input.n_egroups / t_sweep);
        moc_sweep_timing tm;
        if (!moc_get_sweep_timing(moc_handle_of(&params), &tm))
            printf("Last sweep on the device:       %.2f ms (count %.2f, scan %.2f, records %.2f, attenuate %.2f; %ld kernels)\n",
                   tm.total_ms, tm.count_ms, tm.scan_ms, tm.fill_ms, tm.attenuate_ms, tm.launches);
        rule();
    }
    moc_release(&params);
    moc_free_tracks(&input, &params);
    free(ref_argv);
    return 0;
}
