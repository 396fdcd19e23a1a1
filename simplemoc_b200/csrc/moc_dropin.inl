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

// Find (or build) the device mirror of a host Params.  Non-resident mode re-uploads the
// mutable state on every call (host is authoritative); resident mode uploads once.
static Mirror &mirror_for(const Params *P, const Input *I, HostLayout &L, bool need_backward_psi,
                          const char *where, bool caller_streams = false)
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
    if (created || !g_resident) {
        // caller_streams: the non-resident transport_sweep moves the mutable state itself, chunk by
        // chunk, overlapped with the kernels (sweep_core); nothing to upload here
        if (!(caller_streams && !g_resident)) {
            if (upload_mutable(m.h, L, created || need_backward_psi)) die(where);
        } else if (created) {
            // the streamed sweep moves forward rows only: a new mirror still needs the backward rows once (they are
            // scaled by renormalize_flux and moved by the exchange if the caller later switches to resident mode)
            const size_t T3 = (size_t)m.h->T3, G = (size_t)m.h->G;
            if (cudaMemcpy2DAsync(m.h->d.psi + G, sizeof(float) * 2 * G, L.psi + G, sizeof(float) * 2 * G, sizeof(float) * G,
                                  T3, cudaMemcpyHostToDevice, m.h->stream) != cudaSuccess) {
                moc_set_error("upload of the backward angular flux failed");
                die(where);
            }
        }
        if (P->leakage)
            cudaMemcpyAsync(m.h->d.leakage, P->leakage, sizeof(float), cudaMemcpyHostToDevice, m.h->stream);
    }
    return m;
}

extern "C" void transport_sweep(Params *params, Input *I)
{
    HostLayout L;
    Mirror &m = mirror_for(params, I, L, false, "transport_sweep", true);
    long segs = 0;
    // non-resident: uploads, kernels and downloads are pipelined inside the sweep; the call
    // returns after the last byte is back in the host structures
    const CommGrid *ahead = nullptr;
    if (g_resident && g_dropin_grid_set) {
        const int *nb = &g_dropin_grid.x_pos_src;
        bool peers = false;
        for (int q = 0; q < 12; q++) peers = peers || nb[q] >= 0;
        if (!peers || m.h->nccl_comm) ahead = &g_dropin_grid;   // neighbours need moc_comm_init first
    }
    if (sweep_core(m.h, &segs, g_resident ? nullptr : &L, ahead)) die("transport_sweep");
    m.exchanged = ahead != nullptr;
    I->segments_processed = segs;
    if (g_resident) m.dirty_sweep = true;
}

extern "C" void renormalize_flux(Params params, Input I, CommGrid grid)
{
    (void)grid;
    HostLayout L;
    Mirror &m = mirror_for(&params, &I, L, true, "renormalize_flux");
    if (moc_renormalize(m.h)) die("renormalize_flux");
    if (g_resident) m.dirty_all = true;
    else if (download_into(m.h, L, &params, 2)) die("renormalize_flux");
}

extern "C" float update_sources(Params params, Input I, float keff)
{
    HostLayout L;
    Mirror &m = mirror_for(&params, &I, L, false, "update_sources");
    float res = 0.f;
    if (moc_update_sources(m.h, keff, &res)) die("update_sources");
    if (g_resident) m.dirty_all = true;
    else {
        // only fine_source changes
        if (slab_to_host(m.h, 0, (size_t)m.h->N * m.h->F, L.src) != cudaSuccess ||
            cudaStreamSynchronize(m.h->stream) != cudaSuccess) {
            moc_set_error("download of fine_source failed");
            die("update_sources");
        }
    }
    return res;
}

extern "C" float compute_keff(Params params, Input I, CommGrid grid)
{
    (void)grid;
    HostLayout L;
    Mirror &m = mirror_for(&params, &I, L, false, "compute_keff");
    float k = 0.f;
    if (moc_compute_keff(m.h, &k)) die("compute_keff");
    return k;
}

extern "C" void fast_transfer_boundary_fluxes(Params params, Input I, CommGrid grid)
{
    HostLayout L;
    Mirror &m = mirror_for(&params, &I, L, true, "fast_transfer_boundary_fluxes");
    if (g_resident && m.exchanged && g_dropin_grid_set && memcmp(&grid, &g_dropin_grid, sizeof(CommGrid)) == 0) {
        m.exchanged = false;   // done under the sweep (moc_dropin_set_grid)
        m.dirty_all = true;
        return;
    }
    if (moc_exchange(m.h, &grid)) die("fast_transfer_boundary_fluxes");
    if (g_resident) m.dirty_all = true;
    else if (download_into(m.h, L, &params, 2)) die("fast_transfer_boundary_fluxes");
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
