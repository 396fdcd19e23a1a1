/* moc_sweep.inl -- part of moc_device.cu (one translation unit; included there, in this order):
 * the transport sweep: ray-trace launches, batching of z-stacks, the attenuation launch,
 * host<->device streaming of the drop-in call, the exchange under the interior sweep, the L2 probe. */
// ------------------------------------------------------------------ the sweep

static WalkParams walk_params(const moc_handle *h)
{
    WalkParams w;
    memset(&w, 0, sizeof w);
    const Input &I = h->I;
    w.seg_len = h->d.seg_len;
    w.seg_start = h->d.seg_start;
    w.n_seg = h->d.n_seg;
    w.cos_p = h->d.cos_p;
    w.sin_p = h->d.sin_p;
    w.z_height = h->d.z_height;
    w.seg_count = h->d.seg_count;
    w.pair_count = h->d.pair_count;
    w.pair_base = h->d.pair_base;
    w.rec_base = h->d.rec_base;
    w.pair_max = h->d.pair_max;
    w.Zs = (h->Z + 7) / 8 * 8;
    w.rec_ds = h->d.rec_ds;
    w.rec_zin = h->d.rec_zin;
    w.rec_code = h->d.rec_code;
    w.digest = h->want_digest ? h->d.digest : nullptr;
    w.P = h->P;
    w.Z = h->Z;
    w.fai = h->F;
    w.axial_exp = I.axial_exp;
    w.n_regions = (unsigned int)h->N;
    w.mod_magic = h->mod_magic;
    w.mod_shift = h->mod_shift;
    w.mod_fast = h->mod_fast;
    w.fai_magic = (unsigned int)((1ull << 32) / (unsigned long long)std::max(h->F, 1)) + 1u;
    w.z_sep = I.axial_z_sep;
    // solver.c:288-289: float / int, widened; then double / int
    const double node_dz = (double)(float)(I.height / I.decomp_assemblies_ax);
    const double fine_dz = node_dz / (I.cai * I.fai);
    w.node_dz = node_dz;
    w.fine_dz = fine_dz;
    w.dz_interval = (float)fine_dz;
    // solver.c:38: float / int
    w.dz_fine = I.height / (I.fai * I.decomp_assemblies_ax * I.cai);
    w.node_dz_f = (float)node_dz;
    w.flags = reinterpret_cast<unsigned int *>(h->d.digest + 4);
    w.iv_fast = h->iv_fast;
    w.fine_fast = h->fine_fast;
    w.iv_lo = h->iv_lo;
    w.iv_hi = h->iv_hi;
    w.iv_rdz = 1.0f / w.dz_interval;
    w.fine_rdz = 1.0f / w.dz_fine;
    w.seed = h->seed;
    w.rand_base = h->rand_base;
    return w;
}

template <bool FILL>
static void launch_walk(const moc_handle *h, const WalkParams &w, long long n_pairs, cudaStream_t st = nullptr,
                        unsigned max_ctas = 0)
{
    if (n_pairs <= 0) return;
    if (!st) st = h->stream;
    const int Z = h->Z;
    if (h->walk_kernel != 1 && Z <= 128) {
        // short stacks: one warp per stack, 4 stacks per CTA, one launch per ray direction
        const long long P = h->P, H = P / 2, p0 = w.first_pair, p1 = w.first_pair + n_pairs;
        auto ups_before = [&](long long p) { return (p / P) * H + std::min<long long>(p % P, H); };
        const long long up0 = ups_before(p0), n_up = ups_before(p1) - up0;
        const long long down0 = p0 - up0, n_down = n_pairs - n_up;
        const int kpt = (Z + 31) / 32;
        const bool fast = h->iv_fast && h->fine_fast;
#define MOC_WALK(K, UP, before, n)                                                                                \
    if (kpt == K && (n) > 0) {                                                                                    \
        unsigned grid = (unsigned)(((n) + 3) / 4);                                                                \
        if (max_ctas && grid > max_ctas) grid = max_ctas; /* resident grid: warps stride over the stacks */       \
        if (fast) stack_walk_warp_kernel<K, FILL, UP, true><<<grid, 128, 0, st>>>(w, before, n);                  \
        else stack_walk_warp_kernel<K, FILL, UP, false><<<grid, 128, 0, st>>>(w, before, n);                      \
        h->launch_count++;                                                                                        \
    }
        MOC_WALK(1, true, up0, n_up) MOC_WALK(2, true, up0, n_up) MOC_WALK(3, true, up0, n_up) MOC_WALK(4, true, up0, n_up)
        MOC_WALK(1, false, down0, n_down) MOC_WALK(2, false, down0, n_down) MOC_WALK(3, false, down0, n_down)
        MOC_WALK(4, false, down0, n_down)
#undef MOC_WALK
        return;
    }
    if (h->walk_kernel != 1 && Z <= 2048) {
        // taller stacks: the same walk with ceil(Z / 128) warps per stack (one CTA per stack and direction)
        const long long P = h->P, H = P / 2, p0 = w.first_pair, p1 = w.first_pair + n_pairs;
        auto ups_before = [&](long long p) { return (p / P) * H + std::min<long long>(p % P, H); };
        const long long up0 = ups_before(p0), n_up = ups_before(p1) - up0;
        const long long down0 = p0 - up0, n_down = n_pairs - n_up;
        const unsigned threads = 32u * (unsigned)((Z + 127) / 128);
        const bool fast = h->iv_fast && h->fine_fast;
#define MOC_WALK_BLOCK(UP, before, n)                                                                          \
    if ((n) > 0) {                                                                                              \
        unsigned grid = (unsigned)std::min<long long>((n), 0x7fffffffll);                                       \
        if (max_ctas && grid > max_ctas) grid = max_ctas;                                                       \
        if (fast) stack_walk_block_kernel<FILL, UP, true><<<grid, threads, 0, st>>>(w, before, n);              \
        else stack_walk_block_kernel<FILL, UP, false><<<grid, threads, 0, st>>>(w, before, n);                  \
        h->launch_count++;                                                                                      \
    }
        MOC_WALK_BLOCK(true, up0, n_up)
        MOC_WALK_BLOCK(false, down0, n_down)
#undef MOC_WALK_BLOCK
        return;
    }
    int kpt = 1;
    while (kpt < 16 && (Z + kpt - 1) / kpt > 256) kpt *= 2;
    int threads = ((Z + kpt - 1) / kpt + 31) / 32 * 32;
    if (threads > 1024) threads = 1024;
    const unsigned grid = (unsigned)n_pairs;
    h->launch_count++;
    switch (kpt) {
    case 1: stack_walk_kernel<1, FILL><<<grid, threads, 0, st>>>(w); break;
    case 2: stack_walk_kernel<2, FILL><<<grid, threads, 0, st>>>(w); break;
    case 4: stack_walk_kernel<4, FILL><<<grid, threads, 0, st>>>(w); break;
    case 8: stack_walk_kernel<8, FILL><<<grid, threads, 0, st>>>(w); break;
    default: stack_walk_kernel<16, FILL><<<grid, threads, 0, st>>>(w); break;
    }
}

// lane mapping of the attenuation kernel for G groups (LaneMap: moc_device.cu)
static LaneMap choose_lanes(int G, int lanes_override)
{
    if (lanes_override == 0) {
        if (G % 4 == 0) {
            if (G == 104 || G == 100) return {8, 3, 1};
            if (G == 128) return {8, 4, 0};
            if (G == 96) return {8, 3, 0};
            if (G == 64) return {8, 2, 0};
            if (G == 32) return {8, 1, 0};
            if (G == 16) return {4, 1, 0};
        }
    } else if (lanes_override == 32 && G % 4 == 0 && G <= 128) {
        return {32, 1, 0};
    } else if (lanes_override == 16 && G % 4 == 0 && G <= 128) {
        return {16, 2, 0};
    }
    // generic: single groups only
    const int L = (lanes_override == 32 || G > 128) ? 32 : 8;
    int ns = (G + L - 1) / L;
    int r = 1;
    while (r < ns) r *= 2;
    return {L, 0, r};
}

template <int L, int NV4, int NS, int GC>
static int launch_attenuate_mode(const moc_handle *h, const AttenuateParams &a, unsigned grid, size_t smem)
{
    const bool flat = h->I.axial_exp == 0;
    // 0: table, IEEE division; 1: table, verified fast division; 2: SFU
    const int mode = h->exp_mode == 1 ? 2 : (h->fast_cell_ok ? 1 : 0);
    h->launch_count++;
#define MOC_LAUNCH(M, F, C)                                                                                     \
    do {                                                                                                       \
        /* SFU mode uses no shared memory: the whole 256 KB as L1 (more gather lines in flight; measured   */  \
        /* 313 -> 308 ms per launch).  The table modes keep the driver's default split.                     */  \
        static bool configured = false;                                                                        \
        if ((M) == 2 && !configured) {                                                                         \
            cudaFuncSetAttribute(attenuate_kernel<L, NV4, NS, M, F, GC, C>,                                    \
                                 cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);  \
            configured = true;                                                                                 \
        }                                                                                                      \
        /* tables finer than the default (Input.precision below ~2.7e-5: more than 48 KB) need the opt-in   */  \
        static size_t smem_allowed = 48 * 1024;                                                                \
        if ((M) != 2 && smem > smem_allowed) {                                                                 \
            if (cudaFuncSetAttribute(attenuate_kernel<L, NV4, NS, M, F, GC, C>,                                \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { \
                moc_set_error("exponential table of %d cells (%zu bytes) does not fit the shared memory of an SM", \
                              a.table_n, smem);                                                                \
                cudaGetLastError();                                                                            \
                return MOC_EINVAL;                                                                             \
            }                                                                                                  \
            smem_allowed = smem;                                                                               \
        }                                                                                                      \
        attenuate_kernel<L, NV4, NS, M, F, GC, C><<<grid, 128, (M) == 2 ? 0 : smem, h->stream>>>(a);           \
    } while (0)
    if (!flat && a.coef) {
        if (mode == 0) MOC_LAUNCH(0, false, true);
        else if (mode == 1) MOC_LAUNCH(1, false, true);
        else MOC_LAUNCH(2, false, true);
    } else if (!flat) {
        if (mode == 0) MOC_LAUNCH(0, false, false);
        else if (mode == 1) MOC_LAUNCH(1, false, false);
        else MOC_LAUNCH(2, false, false);
    } else {
        if (mode == 0) MOC_LAUNCH(0, true, false);
        else if (mode == 1) MOC_LAUNCH(1, true, false);
        else MOC_LAUNCH(2, true, false);
    }
#undef MOC_LAUNCH
    return MOC_OK;
}

// Does the TMA-staged attenuation kernel apply to this problem?  Quadratic source with the per-stencil coefficient
// slab (L2-resident working set), one of the group counts it is instantiated for, default lane mapping, and a
// shared-memory plan (table + two-stage ring per warp) that fits an SM.
static bool staged_applies(const moc_handle *h)
{
    if (!h->staged || h->I.axial_exp != 2 || !h->d.coef || h->fit_per_segment || h->lanes_override) return false;
    const int G = h->G;
    if (G != 104 && G != 100 && G != 128 && G != 64 && G != 32) return false;
    if ((double)h->N * (h->F - 2) * 4.0 * G >= 4294967296.0) return false;
    return staged_smem_bytes(G, h->table_n, 4, h->exp_mode != 1) <= 227 * 1024;
}

template <int L, int NV4, int NS, int GC>
static int launch_staged_mode(const moc_handle *h, const StagedParams &sp, unsigned grid)
{
    const int mode = h->exp_mode == 1 ? 2 : (h->fast_cell_ok ? 1 : 0);
    const int smem = staged_smem_bytes(GC, h->table_n, 4, mode != 2, L);
    h->launch_count++;
#define MOC_LAUNCH_STAGED(M)                                                                                        \
    do {                                                                                                           \
        static int smem_allowed = 0;                                                                               \
        if (smem > smem_allowed) {                                                                                 \
            if (cudaFuncSetAttribute(attenuate_staged_kernel<L, NV4, NS, M, GC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     smem) != cudaSuccess) {                                                       \
                moc_set_error("attenuate_staged_kernel: %d bytes of shared memory refused", smem);                 \
                cudaGetLastError();                                                                                \
                return MOC_ECUDA;                                                                                  \
            }                                                                                                      \
            smem_allowed = smem;                                                                                   \
        }                                                                                                          \
        attenuate_staged_kernel<L, NV4, NS, M, GC><<<grid, 128, smem, h->stream>>>(sp);                               \
    } while (0)
    if (mode == 0) MOC_LAUNCH_STAGED(0);
    else if (mode == 1) MOC_LAUNCH_STAGED(1);
    else MOC_LAUNCH_STAGED(2);
#undef MOC_LAUNCH_STAGED
    return MOC_OK;
}

static int launch_staged(const moc_handle *h, const AttenuateParams &a, long long n_tracks)
{
    StagedParams sp;
    sp.a = a;
    sp.coef4 = h->d.coef4;
    // (L = 4, eight tracks per warp, <4, 6, 2, 104>: 6.5 % fewer instructions per track-segment, but two CTAs of 255
    // registers per SM leave two warps per scheduler -- 471 instead of 343 ms, profiles/r02_K1_lanes_ab_L4.json.)
    const unsigned grid = (unsigned)((n_tracks + 15) / 16);
    switch (h->G) {
    case 104: return launch_staged_mode<8, 3, 1, 104>(h, sp, grid);
    case 100: return launch_staged_mode<8, 3, 1, 100>(h, sp, grid);
    case 128: return launch_staged_mode<8, 4, 0, 128>(h, sp, grid);
    case 64: return launch_staged_mode<8, 2, 0, 64>(h, sp, grid);
    case 32: return launch_staged_mode<8, 1, 0, 32>(h, sp, grid);
    }
    moc_set_error("no staged attenuation kernel for %d groups", h->G);
    return MOC_EINVAL;
}

static int launch_attenuate(const moc_handle *h, const AttenuateParams &a, long long n_tracks)
{
    if (n_tracks <= 0) return MOC_OK;
    if (h->staged_now) return launch_staged(h, a, n_tracks);
    const LaneMap m = choose_lanes(h->G, h->lanes_override);
    if (4 * m.L * m.NV4 + m.L * m.NS < h->G) {
        moc_set_error("no lane mapping for %d energy groups", h->G);
        return MOC_EINVAL;
    }
    const int tracks_per_block = 4 * (32 / m.L);
    const unsigned grid = (unsigned)((n_tracks + tracks_per_block - 1) / tracks_per_block);
    const size_t smem = sizeof(float) * 2 * ((size_t)h->table_n + 1);
    const int G = h->G;
    // the group counts of the named configurations get their own instantiation (row strides
    // become immediates); everything else takes G from the parameters (GC = 0)
#define MOC_CASE_G(l, v, s, gc) \
    if (m.L == l && m.NV4 == v && m.NS == s && G == gc) return launch_attenuate_mode<l, v, s, gc>(h, a, grid, smem);
#define MOC_CASE(l, v, s) \
    if (m.L == l && m.NV4 == v && m.NS == s) return launch_attenuate_mode<l, v, s, 0>(h, a, grid, smem);
    MOC_CASE_G(8, 3, 1, 104)
    MOC_CASE_G(8, 3, 1, 100)
    MOC_CASE_G(8, 4, 0, 128)
    MOC_CASE_G(8, 2, 0, 64)
    MOC_CASE_G(8, 1, 0, 32)
    MOC_CASE(8, 3, 1)
    MOC_CASE(8, 4, 0)
    MOC_CASE(8, 3, 0)
    MOC_CASE(8, 2, 0)
    MOC_CASE(8, 1, 0)
    MOC_CASE(4, 1, 0)
    MOC_CASE(32, 1, 0)
    MOC_CASE(16, 2, 0)
    MOC_CASE(8, 0, 1)
    MOC_CASE(8, 0, 2)
    MOC_CASE(8, 0, 4)
    MOC_CASE(8, 0, 8)
    MOC_CASE(8, 0, 16)
    MOC_CASE(32, 0, 1)
    MOC_CASE(32, 0, 2)
    MOC_CASE(32, 0, 4)
    MOC_CASE(32, 0, 8)
    MOC_CASE(32, 0, 16)
#undef MOC_CASE
#undef MOC_CASE_G
    moc_set_error("no attenuation kernel instantiated for lane map L=%d NV4=%d NS=%d", m.L, m.NV4, m.NS);
    return MOC_EINVAL;
}

static int ensure_record_capacity(moc_handle *h, long long records)
{
    if (records > h->rec_capacity) {
        if (h->d.rec_ds) cudaFree(h->d.rec_ds);
        if (h->d.rec_zin) cudaFree(h->d.rec_zin);
        if (h->d.rec_code) cudaFree(h->d.rec_code);
        h->d.rec_ds = h->d.rec_zin = nullptr;
        h->d.rec_code = nullptr;
        h->rec_capacity = 0;
        int rc;
        if ((rc = dev_alloc(&h->d.rec_ds, (size_t)records))) return rc;
        if ((rc = dev_alloc(&h->d.rec_zin, (size_t)records))) return rc;
        if ((rc = dev_alloc(&h->d.rec_code, (size_t)records))) return rc;
        h->rec_capacity = records;
    }
    return MOC_OK;
}

// everything of AttenuateParams that does not depend on the batch
static AttenuateParams attenuate_params(const moc_handle *h, const WalkParams &w)
{
    AttenuateParams a;
    memset(&a, 0, sizeof a);
    a.rec_ds = h->d.rec_ds;
    a.rec_zin = h->d.rec_zin;
    a.rec_code = h->d.rec_code;
    a.rec_base = h->d.rec_base;
    a.Zs = w.Zs;
    a.seg_count = h->d.seg_count;
    a.p_weight = h->d.p_weight;
    a.az_weight = h->d.az_weight;
    a.mu = h->d.mu;
    a.psi = h->d.psi;
    a.fine_source = h->d.src;
    a.coef = h->fit_per_segment ? nullptr : h->d.coef;
    a.coef_stencils = h->F - 2;
    a.inv_2dz = 1.0f / (2.f * w.dz_fine);
    a.inv_2dz2 = 1.0f / (2.f * w.dz_fine * w.dz_fine);
    a.fine_flux = h->d.src + (size_t)h->N * h->F * h->Gp;
    a.sigT = h->d.src + (size_t)2 * h->N * h->F * h->Gp;
    a.pitch = h->Gp;
    a.table = h->d.table;
    a.table_dx = h->table_dx;
    a.table_rdx = 1.0f / h->table_dx;
    a.table_max = h->table_max;
    a.table_half_dx = 0.5f * h->table_dx;
    a.table_n = h->table_n;
    a.P = h->P;
    a.Z = h->Z;
    a.G = h->G;
    a.fai = h->F;
    return a;
}

static int exchange_on_stream(moc_handle *h, const CommGrid *grid, cudaStream_t st);   // comms section
static int ensure_exchange_stage(moc_handle *h, long n_recv, long long chunk);
static long exchange_receives(moc_handle *h, const CommGrid *grid, long long *chunk);

// events of the per-batch pipeline, created on demand and kept for the next sweep
static int event_at(moc_handle *h, size_t idx, cudaEvent_t *out)
{
    while (h->ev_pool.size() <= idx) {
        cudaEvent_t e = nullptr;
        CUDA_TRY(cudaEventCreate(&e));
        h->ev_pool.push_back(e);
    }
    *out = h->ev_pool[idx];
    return MOC_OK;
}

// One transport sweep.  io == nullptr: the problem is resident in HBM (moc_sweep).
// io != nullptr: the host structures are authoritative (the drop-in transport_sweep):
// the Track image and the source slab are uploaded first, the forward angular flux
// travels in `stream_chunks` chunks of whole z-stacks on a copy stream while earlier
// chunks are swept, and every finished chunk (flux rows, ray heights) goes back on a
// third stream -- host<->device copies overlap the kernels in both directions.
//
// overlap_grid != nullptr (resident problem only): the boundary exchange of comms.c is started on a
// second stream as soon as the z-stacks whose angular flux it moves -- the first tracks of the
// slab, comms.c:100-183 -- have been swept, and runs under the sweep of the interior stacks.
// io_upload = false (moc_dropin_trust_device): the device copy is what the library last wrote to / read from the host
// structures and the caller promises it has not changed them since: only the downloads happen.
static int sweep_core(moc_handle *h, long *segments_processed, const HostLayout *io,
                      const CommGrid *overlap_grid = nullptr, bool io_upload = true)
{
    CUDA_TRY(cudaSetDevice(h->device));
    const long long pairs = h->T2 * h->P;
    const size_t G = (size_t)h->G;
    cudaEvent_t e_start = h->ev[0], e_count = h->ev[1], e_scan = h->ev[2], e_end = h->ev[3];
    const long launches_before = h->launch_count;
    int rc;

    // ---- chunks of whole z-stacks (only the host-streamed sweep has more than one)
    std::vector<long long> chunk_first;   // first pair of every chunk, plus the end
    long long boundary_pairs = 0;         // z-stacks that hold the flux the boundary exchange moves
    if (overlap_grid && !io) {
        const long n_ops = moc_exchange_plan(&h->I, overlap_grid, nullptr, 0);
        if (n_ops < 0) return (int)n_ops;
        const long long floats = (long long)n_ops * 10000ll * h->G;              // comms.c:12-28: whole messages
        const long long tracks = (floats + 2ll * h->G - 1) / (2ll * h->G);       // [t][2][G] slab
        boundary_pairs = std::min<long long>((tracks + h->Z - 1) / h->Z, pairs);
    }
    if (overlap_grid && !io) {
        // the exchange's receive staging comes first: the record buffers below take what is left
        long long chunk = 0;
        const long n_recv = exchange_receives(h, overlap_grid, &chunk);
        if (n_recv < 0) return (int)n_recv;
        if (n_recv > 0 && (rc = ensure_exchange_stage(h, n_recv, chunk))) return rc;
    }
    if (boundary_pairs > 0 && boundary_pairs < pairs) {
        chunk_first = {0, boundary_pairs, pairs};
    } else {
        long long n = io ? std::min<long long>(std::max(h->stream_chunks, 1), std::max<long long>(pairs, 1)) : 1;
        const long long per = (pairs + n - 1) / std::max<long long>(n, 1);
        for (long long p = 0; p < pairs; p += std::max<long long>(per, 1)) chunk_first.push_back(p);
        chunk_first.push_back(pairs);
    }
    cudaEvent_t e_exchanged = nullptr;    // recorded on the communication stream after the exchange
    const size_t n_chunks = chunk_first.size() - 1;
    size_t ev_next = 0;
    std::vector<cudaEvent_t> ev_up(n_chunks);

    CUDA_TRY(cudaEventRecord(e_start, h->stream));
    if (io) {
        if (!h->up_stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->up_stream, cudaStreamNonBlocking));
        if (!h->down_stream) CUDA_TRY(cudaStreamCreateWithFlags(&h->down_stream, cudaStreamNonBlocking));
        // what the counting pass needs first: ray heights (inside the 40-byte Track image) and,
        // for the attenuation, the source slab
        cudaEvent_t e_img;
        if ((rc = event_at(h, ev_next++, &e_img))) return rc;
        CUDA_TRY(cudaStreamWaitEvent(h->up_stream, e_start, 0));
        if (io_upload) {
            CUDA_TRY(cudaMemcpyAsync(h->d.track_image, io->tracks, sizeof(TrackImage) * (size_t)h->T3,
                                     cudaMemcpyHostToDevice, h->up_stream));
            CUDA_TRY(cudaMemcpy2DAsync(h->d.src, sizeof(float) * h->Gp, io->src, sizeof(float) * G, sizeof(float) * G,
                                       (size_t)(2 * h->F + 1) * (size_t)h->N, cudaMemcpyHostToDevice, h->up_stream));
            h->sigT_known = false;
        }
        CUDA_TRY(cudaEventRecord(e_img, h->up_stream));
        for (size_t c = 0; c < n_chunks; c++) {
            const size_t t0 = (size_t)chunk_first[c] * h->Z, t1 = (size_t)chunk_first[c + 1] * h->Z;
            // forward rows only: row pitch 2*G floats on both sides
            if (io_upload)
                CUDA_TRY(cudaMemcpy2DAsync(h->d.psi + 2 * t0 * G, sizeof(float) * 2 * G, io->psi + 2 * t0 * G,
                                           sizeof(float) * 2 * G, sizeof(float) * G, t1 - t0, cudaMemcpyHostToDevice,
                                           h->up_stream));
            if ((rc = event_at(h, ev_next++, &ev_up[c]))) return rc;
            CUDA_TRY(cudaEventRecord(ev_up[c], h->up_stream));
        }
        // (Tried: downloads held back until the last upload of the call, because eight ranks of one box move 184 GB/s in
        // one direction against 2 x 64 GB/s in both at once -- profiles/r02_host_copy_probe_n8.json.  The downloads are the
        // slow direction, 83 GB/s for all eight ranks alone: 965 instead of 915 ms per call at 8 ranks.  Not kept.)
        CUDA_TRY(cudaStreamWaitEvent(h->stream, e_img, 0));
        if (io_upload) {
            const int threads = 256;
            unpack_tracks_kernel<<<(unsigned)((h->T3 + threads - 1) / threads), threads, 0, h->stream>>>(
                h->d.track_image, h->T3, h->d.p_weight, h->d.z_height);
            h->launch_count++;
        }
    }

    // the largest sigT of the slab, if the slab changed since it was last looked for (see ds_noclamp below)
    unsigned int *const sigt_host = reinterpret_cast<unsigned int *>(h->pair_base_pinned + 2 * (pairs + 1));
    const bool sigt_pending = h->allow_noclamp && !h->sigT_known;
    if (sigt_pending) {
        unsigned int *out = reinterpret_cast<unsigned int *>(h->d.digest + 5);
        sigt_host[0] = 0;
        sigt_host[1] = 1;
        CUDA_TRY(cudaMemsetAsync(out, 0, 2 * sizeof(unsigned int), h->stream));
        sigt_range_kernel<<<148 * 4, 256, 0, h->stream>>>(h->d.src + (size_t)2 * h->N * h->F * h->Gp, h->N, h->G, h->Gp, out);
        h->launch_count++;
        CUDA_TRY(cudaMemcpyAsync(sigt_host, out, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    }

    // ---- pass 1: segment counts per ray and per (2D track, polar angle) stack
    WalkParams w = walk_params(h);
    if (h->want_digest) cudaMemsetAsync(h->d.digest, 0, sizeof(unsigned long long) * 4, h->stream);
    CUDA_TRY(cudaMemsetAsync(h->d.pair_max, 0, sizeof(unsigned int) * (size_t)std::max<long long>(pairs, 1), h->stream));
    launch_walk<false>(h, w, pairs);
    if (h->iv_fast && h->fine_fast) {
        // a ray height outside the node (never produced by the sweep itself, but the host may hand us
        // anything) voids the range the fast interval arithmetic was verified on: count again exactly
        CUDA_TRY(cudaMemcpyAsync(&h->walk_flags_host, w.flags, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        if (h->walk_flags_host) {
            h->iv_fast = h->fine_fast = 0;
            cudaMemsetAsync(w.flags, 0, sizeof(unsigned int), h->stream);
            w = walk_params(h);
            launch_walk<false>(h, w, pairs);
        }
    }
    CUDA_TRY(cudaEventRecord(e_count, h->stream));
    pair_scan_kernel<<<1, 1024, 0, h->stream>>>(h->d.pair_count, h->d.pair_max, w.Zs, h->d.pair_base, h->d.rec_base, pairs);
    h->launch_count++;
    CUDA_TRY(cudaMemcpyAsync(h->pair_base_pinned, h->d.pair_base, sizeof(unsigned long long) * (size_t)(pairs + 1),
                             cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->pair_base_pinned + pairs + 1, h->d.rec_base, sizeof(unsigned long long) * (size_t)(pairs + 1),
                             cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaEventRecord(e_scan, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    // base[]: record slots (segment-major stacks: Zs * longest ray each) -- what the staging buffers
    // and the batches are sized by; the segment total is the last entry of the serial scan
    const unsigned long long *base = h->pair_base_pinned + pairs + 1;
    const unsigned long long total = h->pair_base_pinned[pairs];
    if (sigt_pending) {
        memcpy(&h->sigT_max, &sigt_host[0], sizeof(float));
        h->sigT_clean = sigt_host[1] == 0;
        h->sigT_known = true;
    }

    // ---- batches of whole stacks whose records fit the staging buffers (never across a chunk)
    unsigned long long largest_pair = 0, largest_chunk = 0;
    for (long long p = 0; p < pairs; p++) largest_pair = std::max(largest_pair, base[p + 1] - base[p]);
    for (size_t c = 0; c < n_chunks; c++)
        largest_chunk = std::max(largest_chunk, base[chunk_first[c + 1]] - base[chunk_first[c]]);
    if (largest_pair >= (1ull << 32)) {
        moc_set_error("a single z-stack needs %llu record slots (> 2^32)", largest_pair);
        return MOC_EINVAL;
    }
    // The emitting pass of the ray trace is issue-bound on the ALU/XU pipes, the attenuation on the FMA
    // pipe and the L2: with the warp-per-stack ray trace the records of batch b+1 are emitted by a few
    // resident CTAs per SM UNDER the attenuation of batch b (second stream, two record buffers) instead
    // of by a full grid in front of it.
    const bool overlap_fill = h->fill_overlap_ctas > 0 && h->walk_kernel != 1 && h->Z <= 128;
    const unsigned long long nbuf = overlap_fill ? 2 : 1;
    // 10 % headroom: the record rows a stack needs (its longest ray) drift from sweep to sweep (stale
    // ray heights, solver.c:514-523) and re-allocating multi-GB staging buffers costs ~0.2 s
    unsigned long long target = largest_chunk;
    if (overlap_fill) target = std::max(largest_pair, (largest_chunk + h->fill_batches - 1) / (unsigned long long)h->fill_batches);
    const unsigned long long want = target + target / 10 + 1024;
    long long cap = h->batch_segments;   // records per batch
    if (cap <= 0 && target * nbuf <= (unsigned long long)h->rec_capacity) {
        // the staging buffers of the previous sweep are large enough
        cap = (long long)std::min<unsigned long long>((unsigned long long)h->rec_capacity / nbuf, overlap_fill ? want : ~0ull);
    } else if (cap <= 0) {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        free_b += (size_t)h->rec_capacity * 12;   // what we already hold can be reused
        cap = (long long)((double)free_b * 0.7 / 12.0 / (double)nbuf);
    }
    if ((unsigned long long)cap < largest_pair) cap = (long long)largest_pair;
    if (cap >= (1ll << 32)) cap = (1ll << 32) - 1;
    const long long slot = (long long)std::min<unsigned long long>(std::max(want, largest_pair), (unsigned long long)cap);
    long long need = slot * (long long)nbuf;
    struct Batch {
        long long first, end;
        size_t chunk;
        bool last_of_chunk;
    };
    std::vector<Batch> batches;
    for (size_t c = 0; c < n_chunks; c++) {
        long long p = chunk_first[c];
        const long long pe = chunk_first[c + 1];
        while (p < pe) {
            const unsigned long long lim = base[p] + (unsigned long long)slot;
            // largest q with base[q] <= lim
            long long q = (long long)(std::upper_bound(base + p, base + pe + 1, lim) - base) - 1;
            if (q <= p) q = p + 1;
            batches.push_back({p, q, c, q == pe});
            p = q;
        }
    }
    if ((rc = ensure_record_capacity(h, std::max<long long>(need, 1)))) return rc;
    w = walk_params(h);   // record pointers may have changed

    AttenuateParams a = attenuate_params(h, w);

    // Which segments may skip the reference's x > maxVal test of the table (solver.c:1444-1445)?  Those whose optical
    // length cannot get there with ANY cross section of the slab: ds <= 0.99 maxVal / (largest sigT).  The largest sigT
    // was looked for when the slab last changed (sigt_range_kernel, queued in front of the ray trace and read with the
    // scan's results); a slab with a negative or non-finite value keeps the test everywhere.
    a.ds_noclamp = 0.f;
    if (h->allow_noclamp && h->sigT_known && h->sigT_clean && h->table_max > 0.f)
        a.ds_noclamp = h->sigT_max > 0.f ? (float)(0.99 * (double)h->table_max / (double)h->sigT_max) : 3.0e38f;
    h->noclamp_now = a.ds_noclamp > 0.f;
    h->staged_now = staged_applies(h);
    if (h->staged_now && !h->d.coef4 && (rc = dev_alloc(&h->d.coef4, (size_t)h->N * (h->F - 2) * 4 * (size_t)h->G))) return rc;
    if (h->staged_now) {
        // the source only changes between sweeps (update_sources, uploads): fit every stencil once, and pack
        // (c0, c1, c2, sigT) of a stencil into one contiguous block for the bulk copies of the staged kernel
        const long long cells = h->N * (h->F - 2) * (long long)h->G;
        fit_coefficients4_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, h->stream>>>(
            h->d.src, a.sigT, h->d.coef4, h->N, h->F, h->Gp, h->G, w.dz_fine);
        h->launch_count++;
    } else if (a.coef) {
        const long long cells = h->N * (h->F - 2) * (long long)h->Gp;
        fit_coefficients_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, h->stream>>>(
            h->d.src, h->d.coef, h->N, h->F, h->Gp, w.dz_fine);
        h->launch_count++;
    }

    // three events per batch: before the fill, after it (on the stream that ran it), after the attenuation
    std::vector<cudaEvent_t> ev_b(3 * batches.size());
    for (auto &e : ev_b)
        if ((rc = event_at(h, ev_next++, &e))) return rc;
    const bool two_streams = overlap_fill && batches.size() > 1;
    if (two_streams && !h->fill_stream) {
        // highest priority: the few ray-trace CTAs become resident as soon as attenuation CTAs retire
        int least = 0, greatest = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CUDA_TRY(cudaStreamCreateWithPriority(&h->fill_stream, cudaStreamNonBlocking, greatest));
    }
    if (two_streams && !h->n_sm) CUDA_TRY(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, h->device));
    // batch bi lives in record buffer bi & 1 (one buffer without the overlap)
    auto emit = [&](size_t bi, cudaStream_t st, unsigned max_ctas) {
        const Batch &b = batches[bi];
        const size_t off = two_streams ? (bi & 1) * (size_t)slot : 0;
        w.first_pair = b.first;
        w.batch_first_record = base[b.first];
        w.rec_ds = h->d.rec_ds + off;
        w.rec_zin = h->d.rec_zin + off;
        w.rec_code = h->d.rec_code + off;
        cudaEventRecord(ev_b[3 * bi], st);
        launch_walk<true>(h, w, b.end - b.first, st, max_ctas);
        cudaEventRecord(ev_b[3 * bi + 1], st);
    };
    size_t chunk_start_batch = 0;
    for (size_t bi = 0; bi < batches.size(); bi++) {
        const Batch &b = batches[bi];
        if (io && (bi == 0 || batches[bi - 1].chunk != b.chunk)) {
            CUDA_TRY(cudaStreamWaitEvent(h->stream, ev_up[b.chunk], 0));   // this chunk's flux has arrived
            chunk_start_batch = bi;
        }
        if (!two_streams || bi == 0) emit(bi, h->stream, 0);   // nothing to hide behind: full grid, in front
        if (two_streams) {
            if (bi + 1 < batches.size()) {
                // records of the next batch, under this batch's attenuation; its buffer was last read by
                // the attenuation of batch bi - 1
                if (bi >= 1) CUDA_TRY(cudaStreamWaitEvent(h->fill_stream, ev_b[3 * (bi - 1) + 2], 0));
                else CUDA_TRY(cudaStreamWaitEvent(h->fill_stream, e_scan, 0));
                emit(bi + 1, h->fill_stream, (unsigned)(h->n_sm * h->fill_overlap_ctas));
            }
            if (bi >= 1) CUDA_TRY(cudaStreamWaitEvent(h->stream, ev_b[3 * bi + 1], 0));
        }
        const size_t off = two_streams ? (bi & 1) * (size_t)slot : 0;
        a.rec_ds = h->d.rec_ds + off;
        a.rec_zin = h->d.rec_zin + off;
        a.rec_code = h->d.rec_code + off;
        a.batch_first_record = base[b.first];
        a.first_track = b.first * h->Z;
        a.end_track = b.end * h->Z;
        if ((rc = launch_attenuate(h, a, a.end_track - a.first_track))) return rc;
        CUDA_TRY(cudaEventRecord(ev_b[3 * bi + 2], h->stream));
        if (overlap_grid && !io && b.last_of_chunk && b.end == std::max<long long>(boundary_pairs, 1) &&
            (boundary_pairs < pairs || bi + 1 == batches.size()) && !e_exchanged) {
            // every track the exchange touches has its outgoing flux: exchange under the interior sweep
            if (!h->comm_stream) {
                // highest priority: the exchange's small kernels and NCCL's copy kernels take SM slots as
                // they free up instead of queueing behind the interior sweep's ~5e5 pending CTAs
                int least = 0, greatest = 0;
                CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
                CUDA_TRY(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, greatest));
            }
            CUDA_TRY(cudaStreamWaitEvent(h->comm_stream, ev_b[3 * bi + 2], 0));
            if ((rc = exchange_on_stream(h, overlap_grid, h->comm_stream))) return rc;
            if ((rc = event_at(h, ev_next++, &e_exchanged))) return rc;
            CUDA_TRY(cudaEventRecord(e_exchanged, h->comm_stream));
        }
        if (io && b.last_of_chunk) {
            // the finished chunk goes home while the next one is swept
            const size_t t0 = (size_t)batches[chunk_start_batch].first * h->Z, t1 = (size_t)b.end * h->Z;
            CUDA_TRY(cudaStreamWaitEvent(h->down_stream, ev_b[3 * bi + 2], 0));
            const int threads = 256;
            patch_tracks_kernel<<<(unsigned)((t1 - t0 + threads - 1) / threads), threads, 0, h->down_stream>>>(
                h->d.track_image + t0, (long long)(t1 - t0), h->d.z_height + t0);
            h->launch_count++;
                CUDA_TRY(cudaMemcpyAsync((void *)(io->tracks + t0), h->d.track_image + t0, sizeof(TrackImage) * (t1 - t0),
                                     cudaMemcpyDeviceToHost, h->down_stream));
            CUDA_TRY(cudaMemcpy2DAsync(io->psi + 2 * t0 * G, sizeof(float) * 2 * G, h->d.psi + 2 * t0 * G,
                                       sizeof(float) * 2 * G, sizeof(float) * G, t1 - t0, cudaMemcpyDeviceToHost,
                                       h->down_stream));
        }
    }
    if (io) {
        // scalar flux (the only part of the source slab the sweep writes), then join the streams
        const size_t NF = (size_t)h->N * h->F;
        if (batches.empty()) CUDA_TRY(cudaStreamWaitEvent(h->down_stream, e_scan, 0));
        CUDA_TRY(cudaMemcpy2DAsync(io->src + NF * G, sizeof(float) * G, h->d.src + NF * h->Gp, sizeof(float) * h->Gp,
                                   sizeof(float) * G, NF, cudaMemcpyDeviceToHost, h->down_stream));
        cudaEvent_t e_home;
        if ((rc = event_at(h, ev_next++, &e_home))) return rc;
        CUDA_TRY(cudaEventRecord(e_home, h->down_stream));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, e_home, 0));
    }
    if (overlap_grid && !io) {
        if (!e_exchanged) {
            // no interior to hide behind (the exchange covers every stack, or there are none)
            if ((rc = exchange_on_stream(h, overlap_grid, h->stream))) return rc;
        } else {
            CUDA_TRY(cudaStreamWaitEvent(h->stream, e_exchanged, 0));
        }
    }
    CUDA_TRY(cudaEventRecord(e_end, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    // with the overlap the two phases run concurrently: fill_ms is the time the emitting kernels were
    // resident, attenuate_ms the time from "records ready and previous batch done" to the batch's end
    float fill_ms = 0.f, att_ms = 0.f;
    for (size_t bi = 0; bi < batches.size(); bi++) {
        float f = 0, t = 0, t2 = 0;
        cudaEventElapsedTime(&f, ev_b[3 * bi], ev_b[3 * bi + 1]);
        cudaEventElapsedTime(&t, ev_b[3 * bi + 1], ev_b[3 * bi + 2]);
        if (two_streams && bi >= 1) {
            cudaEventElapsedTime(&t2, ev_b[3 * (bi - 1) + 2], ev_b[3 * bi + 2]);
            t = std::min(t, t2);
        }
        fill_ms += f;
        att_ms += t;
    }
    cudaEventElapsedTime(&h->timing.count_ms, e_start, e_count);
    cudaEventElapsedTime(&h->timing.scan_ms, e_count, e_scan);
    cudaEventElapsedTime(&h->timing.total_ms, e_start, e_end);
    h->timing.fill_ms = fill_ms;
    h->timing.attenuate_ms = att_ms;
    h->timing.n_batches = (long)batches.size();
    h->timing.launches = h->launch_count - launches_before;
    h->I.segments_processed = (long)total;
    h->rand_base += total;   // the serial rand() stream moves on by one draw per segment (solver.c:481)
    if (segments_processed) *segments_processed = (long)total;
    return MOC_OK;
}

extern "C" int moc_sweep(moc_handle *h, long *segments_processed)
{
    if (!h) {
        moc_set_error("moc_sweep: null handle");
        return MOC_EINVAL;
    }
    return sweep_core(h, segments_processed, nullptr);
}

extern "C" int moc_get_sweep_timing(moc_handle *h, moc_sweep_timing *t)
{
    if (!h || !t) return MOC_EINVAL;
    *t = h->timing;
    return MOC_OK;
}

extern "C" int moc_sweep_exchange(moc_handle *h, const CommGrid *grid, long *segments_processed)
{
    if (!h || !grid) {
        moc_set_error("moc_sweep_exchange: null argument");
        return MOC_EINVAL;
    }
    return sweep_core(h, segments_processed, nullptr, grid);
}

// Measured ceiling of the attenuation kernel's memory side: the same gathers (3 source rows + sigT,
// 128 bytes per 8 lanes) and vector reductions on the handle's own source slab, no arithmetic.
// mode 0: gathers only, 1: gathers + reductions.  The flux slab receives zeros only.
extern "C" int moc_probe_l2_gather(moc_handle *h, int mode, double *bytes_per_second)
{
    if (!h || !bytes_per_second || h->F < 3) {
        moc_set_error("moc_probe_l2_gather: needs a handle with fai >= 3");
        return MOC_EINVAL;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const int quads = h->G / 32 > 0 ? h->G / 32 : 1, pitch4 = h->Gp / 4, iters = 2000;
    int sm = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, h->device));
    const unsigned blocks = (unsigned)sm * 5 * 4;
    float4 *sink = nullptr, *zeros = nullptr;
    const size_t slab_rows = (size_t)h->N * h->F;
    CUDA_TRY(cudaMalloc((void **)&sink, 64));
    // reductions go to a scratch copy of the flux slab so the problem state is untouched
    CUDA_TRY(cudaMalloc((void **)&zeros, slab_rows * h->Gp * sizeof(float)));
    CUDA_TRY(cudaMemsetAsync(zeros, 0, slab_rows * h->Gp * sizeof(float), h->stream));
    const float4 *src = reinterpret_cast<const float4 *>(h->d.src);
    float ms = 0.f;
    for (int pass = 0; pass < 2; pass++) {   // first pass warms the L2
        CUDA_TRY(cudaEventRecord(h->ev[6], h->stream));
        if (mode == 0)
            l2_gather_probe_kernel<false><<<blocks, 128, 0, h->stream>>>(src, zeros, (uint32_t)h->N, (uint32_t)h->F, pitch4, quads, iters, sink);
        else
            l2_gather_probe_kernel<true><<<blocks, 128, 0, h->stream>>>(src, zeros, (uint32_t)h->N, (uint32_t)h->F, pitch4, quads, iters, sink);
        CUDA_TRY(cudaEventRecord(h->ev[7], h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        CUDA_TRY(cudaGetLastError());
        cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]);
    }
    cudaFree(sink);
    cudaFree(zeros);
    const double segs = (double)blocks * 16.0 * iters;
    const double bytes = segs * quads * 128.0 * (mode == 0 ? 4.0 : 5.0);
    *bytes_per_second = bytes / ((double)ms * 1e-3);
    return MOC_OK;
}
