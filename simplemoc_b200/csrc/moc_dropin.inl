/* moc_dropin.inl -- part of moc_device.cu (one translation unit; included there, in this order):
 * the drop-in entry points under the reference's names (include/moc_b200.h PART B1) and their device mirrors. */
// ------------------------------------------------------------------ drop-in entry points (PART B1)

struct Mirror {
    moc_handle *h = nullptr;
    bool dirty_sweep = false;   // device holds newer psi/z/flux than the host
    bool dirty_all = false;     // device holds newer everything
    std::vector<void *> registered;   // host slabs this library page-locked (cudaHostRegister)
    const void *host_tracks = nullptr, *host_psi = nullptr, *host_src = nullptr;   // the slabs the mirror was built from
    bool exchanged = false;           // the last transport_sweep already ran the boundary exchange
};
static std::mutex g_mirror_mutex;
static std::unordered_map<const void *, Mirror> g_mirrors;   // keyed by Params.tracks
static int g_resident = 0;
static int g_trust_device = 0;
static unsigned long long g_dropin_seed = 1, g_dropin_rand_base = 0;
static bool g_dropin_configured = false;
static int g_dropin_exp_mode = 0, g_dropin_source_stride = 48;
static CommGrid g_dropin_grid;
static bool g_dropin_grid_set = false;

// With the grid known in advance (resident mode), transport_sweep starts the boundary exchange under
// the sweep of the interior stacks and the following fast_transfer_boundary_fluxes only collects it.
extern "C" void moc_dropin_set_grid(const CommGrid *grid)
{
    g_dropin_grid_set = grid != nullptr;
    if (grid) g_dropin_grid = *grid;
}

extern "C" void moc_set_resident(int on) { g_resident = on ? 1 : 0; }

// Between resident (nothing moves) and the default (every call uploads what it reads and downloads what it writes):
// the caller promises not to modify the host structures between drop-in calls except through these calls.  The device
// copy then always equals the host copy, uploads are skipped (a mirror still takes everything when it is built), and
// every call still writes its results back -- a caller that only READS its structures between calls (prints, dumps,
// convergence checks: the reference's own main.c) sees exactly what the default gives it, for half the traffic.
extern "C" void moc_dropin_trust_device(int on) { g_trust_device = on ? 1 : 0; }

extern "C" int moc_set_device(int device)
{
    int rc = require_device(device);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    return MOC_OK;
}

// options applied to mirrors created by the drop-in entry points
extern "C" void moc_dropin_configure(unsigned long long seed, unsigned long long rand_base, int exp_mode,
                                     int source_stride)
{
    g_dropin_configured = true;
    g_dropin_seed = seed;
    g_dropin_rand_base = rand_base;
    g_dropin_exp_mode = exp_mode;
    g_dropin_source_stride = source_stride;
}

// The reference allocates its slabs with malloc (tracks.c:87-115, source.c:121).  Asynchronous
// copies that overlap kernels need page-locked memory, so slabs that are not already pinned
// (moc_host_alloc pins) are registered in place, once per mirror; failure is not an error -- the
// copies then simply run synchronously.  MOC_B200_NO_PIN=1 disables it.
static void pin_range(Mirror &m, const void *p, size_t bytes)
{
    if (!p || !bytes) return;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    if (at.type != cudaMemoryTypeUnregistered) return;
    if (cudaHostRegister((void *)p, bytes, cudaHostRegisterPortable) == cudaSuccess) m.registered.push_back((void *)p);
    else cudaGetLastError();
}

static void pin_host_slabs(Mirror &m, const HostLayout &L)
{
    const char *off = getenv("MOC_B200_NO_PIN");
    if (off && off[0] == '1') return;
    const moc_handle *h = m.h;
    pin_range(m, L.tracks, sizeof(TrackImage) * (size_t)h->T3);
    pin_range(m, L.psi, sizeof(float) * 2 * (size_t)h->T3 * (size_t)h->G);
    pin_range(m, L.src, sizeof(float) * (size_t)(2 * h->F + 1) * (size_t)h->N * (size_t)h->G);
}

[[noreturn]] static void die(const char *where)
{
    // the reference has no error returns on this path: it prints and exits (solver.c:506-511)
    fprintf(stderr, "libmoc_b200: %s: %s\n", where, moc_last_error());
    exit(1);
}

// What a drop-in call reads from (uploads) or writes to (downloads) the host structures.  The host is
// authoritative between calls (moc_set_resident(0), the default): every call moves exactly the mutable state the
// reference function of the same name reads and writes -- not the whole problem.  Static data (2D tracks, polar
// angles, weights, materials, volumes, the table) travels once, when the mirror is built.
enum {
    PART_TRACKS = 1,    // ray heights (inside the 40-byte Track image)
    PART_PSI_F = 2,     // forward angular flux rows
    PART_PSI_B = 4,     // backward angular flux rows
    PART_SOURCE = 8,    // fine_source
    PART_FLUX = 16,     // fine_flux
    PART_SIGT = 32,     // sigT
    PART_LEAKAGE = 64,
    PART_PSI_HEAD = 128 // the leading floats of the flux slab the boundary exchange moves (comms.c:100-183)
};

static long long exchange_head_floats(const moc_handle *h, const CommGrid *grid)
{
    const long n_ops = moc_exchange_plan(&h->I, grid, nullptr, 0);
    if (n_ops <= 0) return 0;
    return std::min<long long>((long long)n_ops * 10000ll * h->G, 2ll * h->T3 * h->G);
}

static int move_parts(moc_handle *h, const HostLayout &L, const Params *P, int parts, bool up, long long head_floats = 0)
{
    const size_t T3 = (size_t)h->T3, G = (size_t)h->G, N = (size_t)h->N, F = (size_t)h->F;
    const cudaMemcpyKind kind = up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    const int threads = 256;
    if (parts & PART_TRACKS) {
        int rc = need_track_image(h);
        if (rc) return rc;
        if (up) {
            CUDA_TRY(cudaMemcpyAsync(h->d.track_image, L.tracks, sizeof(TrackImage) * T3, kind, h->stream));
            unpack_tracks_kernel<<<(unsigned)((T3 + threads - 1) / threads), threads, 0, h->stream>>>(
                h->d.track_image, (long long)T3, h->d.p_weight, h->d.z_height);
        } else {
            patch_tracks_kernel<<<(unsigned)((T3 + threads - 1) / threads), threads, 0, h->stream>>>(
                h->d.track_image, (long long)T3, h->d.z_height);
            CUDA_TRY(cudaMemcpyAsync((void *)L.tracks, h->d.track_image, sizeof(TrackImage) * T3, kind, h->stream));
        }
        h->launch_count++;
    }
    if ((parts & PART_PSI_F) && (parts & PART_PSI_B)) {
        if (up) CUDA_TRY(cudaMemcpyAsync(h->d.psi, L.psi, sizeof(float) * 2 * T3 * G, kind, h->stream));
        else CUDA_TRY(cudaMemcpyAsync(L.psi, h->d.psi, sizeof(float) * 2 * T3 * G, kind, h->stream));
    } else if (parts & (PART_PSI_F | PART_PSI_B)) {
        const size_t off = (parts & PART_PSI_B) ? G : 0;   // one row of every [t][2][G] pair: pitch 2 G floats
        if (up) CUDA_TRY(cudaMemcpy2DAsync(h->d.psi + off, sizeof(float) * 2 * G, L.psi + off, sizeof(float) * 2 * G,
                                           sizeof(float) * G, T3, kind, h->stream));
        else CUDA_TRY(cudaMemcpy2DAsync(L.psi + off, sizeof(float) * 2 * G, h->d.psi + off, sizeof(float) * 2 * G,
                                        sizeof(float) * G, T3, kind, h->stream));
    } else if ((parts & PART_PSI_HEAD) && head_floats > 0) {
        if (up) CUDA_TRY(cudaMemcpyAsync(h->d.psi, L.psi, sizeof(float) * (size_t)head_floats, kind, h->stream));
        else CUDA_TRY(cudaMemcpyAsync(L.psi, h->d.psi, sizeof(float) * (size_t)head_floats, kind, h->stream));
    }
    // the source slab: fine_source rows [0, NF), fine_flux rows [NF, 2NF), sigT rows [2NF, 2NF + N)
    struct { int part; size_t row0, rows; } slab[3] = {{PART_SOURCE, 0, N * F}, {PART_FLUX, N * F, N * F}, {PART_SIGT, 2 * N * F, N}};
    for (const auto &q : slab)
        if (parts & q.part) {
            if (up) CUDA_TRY(slab_to_device(h, q.row0, q.rows, L.src + q.row0 * G));
            else CUDA_TRY(slab_to_host(h, q.row0, q.rows, L.src + q.row0 * G));
        }
    if ((parts & PART_LEAKAGE) && P->leakage) {
        if (up) CUDA_TRY(cudaMemcpyAsync(h->d.leakage, P->leakage, sizeof(float), kind, h->stream));
        else CUDA_TRY(cudaMemcpyAsync(P->leakage, h->d.leakage, sizeof(float), kind, h->stream));
    }
    CUDA_TRY(cudaGetLastError());
    return MOC_OK;
}

// Find (or build) the device mirror of a host Params.  Non-resident mode uploads `reads` (PART_* mask) on every
// call (host is authoritative); resident mode uploads everything once, when the mirror is built.
static Mirror &mirror_for(const Params *P, const Input *I, HostLayout &L, int reads, const char *where,
                          const CommGrid *grid = nullptr)
{
    std::lock_guard<std::mutex> lock(g_mirror_mutex);
    Mirror &m = g_mirrors[(const void *)P->tracks];
    bool created = false;
    if (m.h) {
        // Mirrors are keyed by the address of Params.tracks: a caller that frees its problem and builds another one
        // may get the same address back from malloc.  A mirror whose sizes or slabs are not the incoming ones is stale.
        if (inspect_layout(I, P, m.h->source_stride, L)) die(where);
        const moc_handle *h = m.h;
        if (h->T2 != I->ntracks_2D || h->T3 != I->ntracks || h->N != I->n_source_regions_per_node || h->P != I->n_polar_angles ||
            h->Z != I->z_stacked || h->G != I->n_egroups || h->F != I->fai || h->I.axial_exp != I->axial_exp ||
            m.host_tracks != (const void *)L.tracks || m.host_psi != (const void *)L.psi || m.host_src != (const void *)L.src) {
            moc_destroy(m.h);
            for (void *p : m.registered) cudaHostUnregister(p);
            m = Mirror();
        }
    }
    if (!m.h) {
        int device = 0;
        cudaGetDevice(&device);
        if (create_common(I, P, device, g_dropin_source_stride, &m.h, L)) die(where);
        if (const char *c = getenv("MOC_B200_STREAM_CHUNKS"))
            if (atoi(c) >= 1 && atoi(c) <= 4096) m.h->stream_chunks = atoi(c);
        m.h->seed = g_dropin_seed;
        m.h->rand_base = g_dropin_rand_base;
        m.h->exp_mode = g_dropin_exp_mode;
        if (!g_dropin_configured) {
            // A host program that cannot be edited to call moc_dropin_configure (the reference's own main.c,
            // linked unmodified) may export  void moc_host_rand_state(unsigned long long *seed,
            // unsigned long long *calls)  instead: where ITS rand() stream stands at the first sweep.
            typedef void (*rand_state_fn)(unsigned long long *, unsigned long long *);
            if (rand_state_fn f = (rand_state_fn)dlsym(RTLD_DEFAULT, "moc_host_rand_state")) {
                unsigned long long seed = g_dropin_seed, calls = g_dropin_rand_base;
                f(&seed, &calls);
                m.h->seed = seed;
                m.h->rand_base = calls;
            }
        }
        created = true;
        m.host_tracks = L.tracks;
        m.host_psi = L.psi;
        m.host_src = L.src;
    }
    if (created) pin_host_slabs(m, L);
    // a new mirror takes the whole mutable state once (whatever this call reads): later calls, and a later switch to
    // resident mode, find every array defined
    int parts = reads;
    if (created) parts = PART_TRACKS | PART_PSI_F | PART_PSI_B | PART_SOURCE | PART_FLUX | PART_SIGT | PART_LEAKAGE;
    else if (g_resident || g_trust_device) parts = 0;
    if (parts && move_parts(m.h, L, P, parts, true, (parts & PART_PSI_HEAD) && grid ? exchange_head_floats(m.h, grid) : 0))
        die(where);
    return m;
}

// Wait for the stream (downloads included); the reference's functions return when their results are in place.
static void finish(moc_handle *h, const char *where)
{
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) {
        moc_set_error("%s: %s", where, cudaGetErrorString(cudaGetLastError()));
        die(where);
    }
}

// renormalize_flux on host structures (solver.c:1143-1230), host authoritative: the scalar flux (14 MB on the default
// problem) goes up, is reduced and scaled and comes back; the angular flux -- every forward and backward row, 12.9 GB
// -- travels in chunks: upload of chunk c+1, scaling of chunk c and download of chunk c-1 overlap on three streams.
static int renormalize_streamed(Mirror &m, const HostLayout &L, const Params *P)
{
    moc_handle *h = m.h;
    int rc;
    if ((rc = renormalize_scalar_flux(h))) return rc;                 // the scalar flux was uploaded by mirror_for
    if ((rc = move_parts(h, L, P, PART_FLUX, false))) return rc;
    if (!h->up_stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->up_stream, cudaStreamNonBlocking));
    if (!h->down_stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->down_stream, cudaStreamNonBlocking));
    const long long n = 2 * h->T3 * h->G;
    const long long chunks = std::max<long long>(1, std::min<long long>(h->stream_chunks, n / 4096 + 1));
    const long long per = ((n + chunks - 1) / chunks + 3) / 4 * 4;
    size_t ev_next = 0;
    cudaEvent_t e_begin, e_home;
    if ((rc = event_at(h, ev_next++, &e_begin))) return rc;
    CUDA_TRY(cudaEventRecord(e_begin, h->stream));
    CUDA_TRY(cudaStreamWaitEvent(h->up_stream, e_begin, 0));
    for (long long first = 0; first < n; first += per) {
        const long long count = std::min(per, n - first);
        cudaEvent_t e_up, e_done;
        if ((rc = event_at(h, ev_next++, &e_up))) return rc;
        if ((rc = event_at(h, ev_next++, &e_done))) return rc;
        if (!g_trust_device)
            CUDA_TRY(cudaMemcpyAsync(h->d.psi + first, L.psi + first, sizeof(float) * (size_t)count, cudaMemcpyHostToDevice, h->up_stream));
        CUDA_TRY(cudaEventRecord(e_up, h->up_stream));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, e_up, 0));
        scale_psi_range(h, first, count, h->stream);
        CUDA_TRY(cudaEventRecord(e_done, h->stream));
        CUDA_TRY(cudaStreamWaitEvent(h->down_stream, e_done, 0));
        CUDA_TRY(cudaMemcpyAsync(L.psi + first, h->d.psi + first, sizeof(float) * (size_t)count, cudaMemcpyDeviceToHost, h->down_stream));
    }
    if ((rc = event_at(h, ev_next++, &e_home))) return rc;
    CUDA_TRY(cudaEventRecord(e_home, h->down_stream));
    CUDA_TRY(cudaStreamWaitEvent(h->stream, e_home, 0));
    CUDA_TRY(cudaGetLastError());
    return MOC_OK;
}

extern "C" void transport_sweep(Params *params, Input *I)
{
    HostLayout L;
    // non-resident: the sweep moves its own inputs and outputs, chunk by chunk, overlapped with the kernels (sweep_core)
    Mirror &m = mirror_for(params, I, L, 0, "transport_sweep");
    long segs = 0;
    const CommGrid *ahead = nullptr;
    if (g_resident && g_dropin_grid_set) {
        const int *nb = &g_dropin_grid.x_pos_src;
        bool peers = false;
        for (int q = 0; q < 12; q++) peers = peers || nb[q] >= 0;
        if (!peers || m.h->nccl_comm) ahead = &g_dropin_grid;   // neighbours need moc_comm_init first
    }
    if (sweep_core(m.h, &segs, g_resident ? nullptr : &L, ahead, !g_trust_device)) die("transport_sweep");
    m.exchanged = ahead != nullptr;
    I->segments_processed = segs;
    if (g_resident) m.dirty_sweep = true;
}

// solver.c:556-891: the reference compiles this sweep and calls it from nowhere (it is not in SimpleMOC_header.h either);
// a host program that does call it finds the name here.  Plain copies around moc_two_way_sweep: everything it reads
// (ray heights and weights, both angular-flux rows, sources, scalar flux, sigT) up, everything it writes (both
// angular-flux rows, scalar flux, the reset ray heights) down.
extern "C" void two_way_transport_sweep(Params *params, Input *I)
{
    HostLayout L;
    Mirror &m = mirror_for(params, I, L, PART_TRACKS | PART_PSI_F | PART_PSI_B | PART_SOURCE | PART_FLUX | PART_SIGT,
                           "two_way_transport_sweep");
    long segs = 0;
    if (moc_two_way_sweep(m.h, &segs)) die("two_way_transport_sweep");
    I->segments_processed = segs;
    if (g_resident) {
        m.dirty_all = true;
        return;
    }
    if (move_parts(m.h, L, params, PART_TRACKS | PART_PSI_F | PART_PSI_B | PART_FLUX, false)) die("two_way_transport_sweep");
    finish(m.h, "two_way_transport_sweep");
}

extern "C" void renormalize_flux(Params params, Input I, CommGrid grid)
{
    (void)grid;
    HostLayout L;
    // reads the scalar flux (here) and every angular flux (streamed below)
    Mirror &m = mirror_for(&params, &I, L, PART_FLUX, "renormalize_flux");
    if (g_resident) {
        if (moc_renormalize(m.h)) die("renormalize_flux");
        m.dirty_all = true;
        return;
    }
    if (renormalize_streamed(m, L, &params)) die("renormalize_flux");
    finish(m.h, "renormalize_flux");
}

extern "C" float update_sources(Params params, Input I, float keff)
{
    HostLayout L;
    // reads the scalar flux and the old sources (residual, solver.c:1290-1297), writes the sources
    Mirror &m = mirror_for(&params, &I, L, PART_FLUX | PART_SOURCE, "update_sources");
    float res = 0.f;
    if (moc_update_sources(m.h, keff, &res)) die("update_sources");
    if (g_resident) m.dirty_all = true;
    else {
        if (move_parts(m.h, L, &params, PART_SOURCE, false)) die("update_sources");
        finish(m.h, "update_sources");
    }
    return res;
}

extern "C" float compute_keff(Params params, Input I, CommGrid grid)
{
    (void)grid;
    HostLayout L;
    Mirror &m = mirror_for(&params, &I, L, PART_FLUX | PART_LEAKAGE, "compute_keff");
    float k = 0.f;
    if (moc_compute_keff(m.h, &k)) die("compute_keff");
    return k;
}

extern "C" void fast_transfer_boundary_fluxes(Params params, Input I, CommGrid grid)
{
    HostLayout L;
    // reads and writes the leading chunks of the flux slab (comms.c:100-183) and the leakage
    Mirror &m = mirror_for(&params, &I, L, PART_PSI_HEAD | PART_LEAKAGE, "fast_transfer_boundary_fluxes", &grid);
    if (g_resident && m.exchanged && g_dropin_grid_set && memcmp(&grid, &g_dropin_grid, sizeof(CommGrid)) == 0) {
        m.exchanged = false;   // done under the sweep (moc_dropin_set_grid)
        m.dirty_all = true;
        return;
    }
    if (moc_exchange(m.h, &grid)) die("fast_transfer_boundary_fluxes");
    if (g_resident) m.dirty_all = true;
    else {
        if (move_parts(m.h, L, &params, PART_PSI_HEAD | PART_LEAKAGE, false, exchange_head_floats(m.h, &grid)))
            die("fast_transfer_boundary_fluxes");
        finish(m.h, "fast_transfer_boundary_fluxes");
    }
}

extern "C" int moc_sync_to_host(Params *params)
{
    if (!params) return MOC_EINVAL;
    std::lock_guard<std::mutex> lock(g_mirror_mutex);
    auto it = g_mirrors.find((const void *)params->tracks);
    if (it == g_mirrors.end() || !it->second.h) {
        moc_set_error("moc_sync_to_host: no device mirror for this Params");
        return MOC_EINVAL;
    }
    Mirror &m = it->second;
    HostLayout L;
    int rc = inspect_layout(&m.h->I, params, m.h->source_stride, L);
    if (rc) return rc;
    rc = download_into(m.h, L, params, 2);
    if (!rc) m.dirty_sweep = m.dirty_all = false;
    return rc;
}

extern "C" int moc_release(Params *params)
{
    if (!params) return MOC_EINVAL;
    std::lock_guard<std::mutex> lock(g_mirror_mutex);
    auto it = g_mirrors.find((const void *)params->tracks);
    if (it == g_mirrors.end()) return MOC_OK;
    moc_destroy(it->second.h);
    for (void *p : it->second.registered) cudaHostUnregister(p);
    g_mirrors.erase(it);
    return MOC_OK;
}

// the handle behind a Params used through the drop-in names (for timing queries)
extern "C" moc_handle *moc_handle_of(Params *params)
{
    std::lock_guard<std::mutex> lock(g_mirror_mutex);
    auto it = g_mirrors.find((const void *)params->tracks);
    return it == g_mirrors.end() ? nullptr : it->second.h;
}
