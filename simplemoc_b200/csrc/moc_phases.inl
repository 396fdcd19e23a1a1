/* moc_phases.inl -- part of moc_device.cu (one translation unit; included there, in this order):
 * renormalize_flux / update_sources / compute_keff on the device, and array access (moc_get_array ...). */
// ------------------------------------------------------------------ reductions

static SourceParams source_params(const moc_handle *h)
{
    SourceParams p;
    p.fine_source = h->d.src;
    p.fine_flux = h->d.src + (size_t)h->N * h->F * h->Gp;
    p.pitch = h->Gp;
    p.xs = h->d.xs;
    p.scatter = h->d.scatter;
    p.xs_index = h->d.xs_index;
    p.vol = h->d.vol;
    p.N = h->N;
    p.G = h->G;
    p.fai = h->F;
    return p;
}

static int allreduce_scalars(moc_handle *h, float *dev, int count);   // comms section

// renormalize_flux, first half (solver.c:1143-1214): total fission rate (all ranks), scalar flux scaled
static int renormalize_scalar_flux(moc_handle *h)
{
    const SourceParams p = source_params(h);
    const unsigned rb = (unsigned)((h->N + 127) / 128);
    region_fission_rate_kernel<<<rb, 128, 0, h->stream>>>(p, h->d.per_region_a);
    pairwise_reduce_kernel<<<1, 256, 0, h->stream>>>(h->d.per_region_a, h->N, h->d.scalars, 0);
    if (h->nranks > 1) {
        int rc = allreduce_scalars(h, h->d.scalars, 1);   // solver.c:1190-1195
        if (rc) return rc;
    }
    const long long cells = h->N * h->F * h->Gp;
    scale_flux_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, h->stream>>>(p, h->d.scalars);
    h->launch_count += 3;
    CUDA_TRY(cudaGetLastError());
    return MOC_OK;
}

// renormalize_flux, second half (solver.c:1219-1226): floats [first, first + count) of the angular-flux slab
// *= 1 / (total fission rate); first must be a multiple of 4
static void scale_psi_range(moc_handle *h, long long first, long long count, cudaStream_t st)
{
    if (count <= 0) return;
    const long long n4 = count / 4;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(148 * 8, (n4 + 255) / 256));
    scale_psi_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<float4 *>(h->d.psi + first), n4, h->d.psi + first + 4 * n4,
                                           (int)(count - 4 * n4), h->d.scalars);
    h->launch_count++;
}

extern "C" int moc_renormalize(moc_handle *h)
{
    if (!h) return MOC_EINVAL;
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = renormalize_scalar_flux(h);
    if (rc) return rc;
    scale_psi_range(h, 0, 2 * h->T3 * h->G, h->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MOC_OK;
}

extern "C" int moc_update_sources(moc_handle *h, float keff, float *res)
{
    if (!h) return MOC_EINVAL;
    CUDA_TRY(cudaSetDevice(h->device));
    const SourceParams p = source_params(h);
    const float inverse_k = (float)(1.0 / (double)keff);   // solver.c:1241
    const long long rows = h->N * h->F;
    const int threads = std::min(128, (h->G + 31) / 32 * 32);
    const int depth = tree_depth(h->G);
    if (depth <= 5) {
        // the G-term sums of a row spread over 2^depth lanes each (same tree, same order of additions)
        update_sources_coop_kernel<<<(unsigned)rows, 256, sizeof(float) * (2 * (size_t)h->G + 1), h->stream>>>(
            p, inverse_k, h->d.per_fine, depth);
    } else {
        update_sources_kernel<<<(unsigned)rows, threads, sizeof(float) * 2 * (size_t)h->G, h->stream>>>(
            p, inverse_k, h->d.per_fine);
    }
    region_fold_kernel<<<(unsigned)((h->N + 127) / 128), 128, 0, h->stream>>>(h->d.per_fine, h->N, h->F,
                                                                             h->d.per_region_a);
    pairwise_reduce_kernel<<<1, 256, 0, h->stream>>>(h->d.per_region_a, h->N, h->d.scalars, 1);
    h->launch_count += 3;
    CUDA_TRY(cudaGetLastError());
    float r = 0.f;
    CUDA_TRY(cudaMemcpyAsync(&r, h->d.scalars + 1, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (res) *res = r;
    return MOC_OK;
}

extern "C" int moc_compute_keff(moc_handle *h, float *keff)
{
    if (!h) return MOC_EINVAL;
    CUDA_TRY(cudaSetDevice(h->device));
    const SourceParams p = source_params(h);
    region_reaction_rates_kernel<<<(unsigned)((h->N + 127) / 128), 128, 0, h->stream>>>(p, h->d.per_region_a,
                                                                                       h->d.per_region_b);
    pairwise_reduce_kernel<<<1, 256, 0, h->stream>>>(h->d.per_region_a, h->N, h->d.scalars, 2);   // absorption
    pairwise_reduce_kernel<<<1, 256, 0, h->stream>>>(h->d.per_region_b, h->N, h->d.scalars, 3);   // fission
    h->launch_count += 3;
    CUDA_TRY(cudaMemcpyAsync(h->d.scalars + 4, h->d.leakage, sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    if (h->nranks > 1) {
        int rc = allreduce_scalars(h, h->d.scalars + 2, 3);   // solver.c:1394-1418, one vector
        if (rc) return rc;
    }
    float v[3];
    CUDA_TRY(cudaMemcpyAsync(v, h->d.scalars + 2, sizeof(float) * 3, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    if (keff) *keff = v[1] / (v[0] + v[2]);   // solver.c:1423,1425
    return MOC_OK;
}

// ------------------------------------------------------------------ array access

// `rows` > 0: the array is `rows` rows of the padded source slab starting at slab row `row0`
static int array_span(moc_handle *h, int which, void **ptr, size_t *bytes, size_t *row0, size_t *rows)
{
    const size_t T3 = (size_t)h->T3, G = (size_t)h->G, N = (size_t)h->N, F = (size_t)h->F;
    *rows = 0;
    *row0 = 0;
    *ptr = nullptr;
    switch (which) {
    case MOC_ARR_FINE_SOURCE: *row0 = 0; *rows = N * F; *bytes = sizeof(float) * N * F * G; return MOC_OK;
    case MOC_ARR_FINE_FLUX: *row0 = N * F; *rows = N * F; *bytes = sizeof(float) * N * F * G; return MOC_OK;
    case MOC_ARR_SIGT: *row0 = 2 * N * F; *rows = N; *bytes = sizeof(float) * N * G; return MOC_OK;
    case MOC_ARR_PSI: *ptr = h->d.psi; *bytes = sizeof(float) * 2 * T3 * G; return MOC_OK;
    case MOC_ARR_Z_HEIGHT: *ptr = h->d.z_height; *bytes = sizeof(float) * T3; return MOC_OK;
    case MOC_ARR_P_WEIGHT: *ptr = h->d.p_weight; *bytes = sizeof(float) * T3; return MOC_OK;
    case MOC_ARR_SEG_COUNT: *ptr = h->d.seg_count; *bytes = sizeof(uint32_t) * T3; return MOC_OK;
    case MOC_ARR_QSR_DIGEST: *ptr = h->d.digest; *bytes = sizeof(unsigned long long) * 4; return MOC_OK;
    case MOC_ARR_QSR_DIGEST_BACK: *ptr = h->d.digest + 8; *bytes = sizeof(unsigned long long) * 4; return MOC_OK;
    }
    moc_set_error("unknown array id %d", which);
    return MOC_EINVAL;
}

extern "C" int moc_get_array(moc_handle *h, int which, void *dst, size_t bytes)
{
    if (!h || !dst) return MOC_EINVAL;
    CUDA_TRY(cudaSetDevice(h->device));
    void *p;
    size_t n, row0, rows;
    int rc = array_span(h, which, &p, &n, &row0, &rows);
    if (rc) return rc;
    if (bytes != n) {
        moc_set_error("moc_get_array(%d): buffer is %zu bytes, array is %zu", which, bytes, n);
        return MOC_EINVAL;
    }
    if (rows) CUDA_TRY(slab_to_host(h, row0, rows, (float *)dst));
    else CUDA_TRY(cudaMemcpyAsync(dst, p, n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MOC_OK;
}

extern "C" int moc_set_array(moc_handle *h, int which, const void *src, size_t bytes)
{
    if (!h || !src) return MOC_EINVAL;
    CUDA_TRY(cudaSetDevice(h->device));
    void *p;
    size_t n, row0, rows;
    int rc = array_span(h, which, &p, &n, &row0, &rows);
    if (rc) return rc;
    if (bytes != n || which == MOC_ARR_SEG_COUNT || which == MOC_ARR_QSR_DIGEST || which == MOC_ARR_QSR_DIGEST_BACK) {
        moc_set_error("moc_set_array(%d): read-only array or size mismatch (%zu vs %zu)", which, bytes, n);
        return MOC_EINVAL;
    }
    if (rows) CUDA_TRY(slab_to_device(h, row0, rows, (const float *)src));
    else CUDA_TRY(cudaMemcpyAsync(p, src, n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MOC_OK;
}

extern "C" float moc_get_leakage(moc_handle *h)
{
    if (!h) return 0.f;
    cudaSetDevice(h->device);
    float v = 0.f;
    cudaMemcpyAsync(&v, h->d.leakage, sizeof(float), cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    return v;
}

extern "C" void *moc_get_stream(moc_handle *h) { return h ? (void *)h->stream : nullptr; }
extern "C" long moc_get_launch_count(moc_handle *h) { return h ? h->launch_count : -1; }

extern "C" int moc_synchronize(moc_handle *h)
{
    if (!h) return MOC_EINVAL;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MOC_OK;
}

// what: 1 = forward psi + z_height + fine_flux (what transport_sweep mutates)
//       2 = everything mutable (psi both rows, z_height, whole source slab, leakage)
static int download_into(moc_handle *h, const HostLayout &L, Params *P, int what)
{
    const size_t T3 = (size_t)h->T3, G = (size_t)h->G, N = (size_t)h->N, F = (size_t)h->F;
    const int threads = 256;
    if (!h->d.track_image) {
        moc_set_error("this handle was generated on the device (moc_create_synthetic): there are no host Track structures to write back to");
        return MOC_EINVAL;
    }
    patch_tracks_kernel<<<(unsigned)((T3 + threads - 1) / threads), threads, 0, h->stream>>>(
        h->d.track_image, (long long)T3, h->d.z_height);
    CUDA_TRY(cudaMemcpyAsync((void *)L.tracks, h->d.track_image, sizeof(TrackImage) * T3, cudaMemcpyDeviceToHost, h->stream));
    if (what == 1) {
        CUDA_TRY(cudaMemcpy2DAsync(L.psi, sizeof(float) * 2 * G, h->d.psi, sizeof(float) * 2 * G, sizeof(float) * G,
                                   T3, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(slab_to_host(h, N * F, N * F, L.src + N * F * G));
    } else {
        CUDA_TRY(cudaMemcpyAsync(L.psi, h->d.psi, sizeof(float) * 2 * T3 * G, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(slab_to_host(h, 0, (2 * F + 1) * N, L.src));
        if (P->leakage)
            CUDA_TRY(cudaMemcpyAsync(P->leakage, h->d.leakage, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    return MOC_OK;
}

extern "C" int moc_download(moc_handle *h, Params *P)
{
    if (!h || !P) return MOC_EINVAL;
    CUDA_TRY(cudaSetDevice(h->device));
    HostLayout L;
    int rc = inspect_layout(&h->I, P, h->source_stride, L);
    if (rc) return rc;
    return download_into(h, L, P, 2);
}

extern "C" int moc_upload(moc_handle *h, const Params *P)
{
    if (!h || !P) return MOC_EINVAL;
    CUDA_TRY(cudaSetDevice(h->device));
    HostLayout L;
    int rc = inspect_layout(&h->I, P, h->source_stride, L);
    if (rc) return rc;
    if ((rc = upload_mutable(h, L, true))) return rc;
    if (P->leakage) CUDA_TRY(cudaMemcpyAsync(h->d.leakage, P->leakage, sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MOC_OK;
}
