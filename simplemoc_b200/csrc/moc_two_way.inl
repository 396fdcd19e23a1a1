/* moc_two_way.inl -- part of moc_device.cu: two_way_transport_sweep (solver.c:556-891) on the handle.
 * Same structure as sweep_core -- count pass, scan, batches of whole z-stacks (fill pass, attenuation) -- on the
 * plain kernels of moc_two_way.cuh, one stream, no host streaming and no exchange overlap: the reference never
 * calls this sweep (main.c:60), it is here so that every function on the path exists. */

template <bool FILL>
static void launch_two_way_walk(moc_handle *h, const WalkParams &w, long long n_pairs)
{
    if (n_pairs <= 0) return;
    const int Z = h->Z;
    int kpt = 1;
    while (kpt < 16 && (Z + kpt - 1) / kpt > 256) kpt *= 2;
    int threads = ((Z + kpt - 1) / kpt + 31) / 32 * 32;
    if (threads > 1024) threads = 1024;
    const unsigned grid = (unsigned)n_pairs;
    h->launch_count++;
    switch (kpt) {
    case 1: two_way_walk_kernel<1, FILL><<<grid, threads, 0, h->stream>>>(w); break;
    case 2: two_way_walk_kernel<2, FILL><<<grid, threads, 0, h->stream>>>(w); break;
    case 4: two_way_walk_kernel<4, FILL><<<grid, threads, 0, h->stream>>>(w); break;
    case 8: two_way_walk_kernel<8, FILL><<<grid, threads, 0, h->stream>>>(w); break;
    default: two_way_walk_kernel<16, FILL><<<grid, threads, 0, h->stream>>>(w); break;
    }
}

template <int MODE, bool FLAT>
static int launch_two_way_attenuate(moc_handle *h, const AttenuateParams &a, const TwoWayParams &tw, long long n_tracks)
{
    const size_t smem = MODE == 2 ? 0 : sizeof(float) * 2 * ((size_t)h->table_n + 1);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(attenuate_two_way_kernel<MODE, FLAT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess) {
        cudaGetLastError();
        moc_set_error("exponential table of %d cells (%zu bytes) does not fit the shared memory of an SM", h->table_n, smem);
        return MOC_EINVAL;
    }
    attenuate_two_way_kernel<MODE, FLAT><<<(unsigned)((n_tracks + 3) / 4), 128, smem, h->stream>>>(a, tw);
    h->launch_count++;
    return MOC_OK;
}

extern "C" int moc_two_way_sweep(moc_handle *h, long *segments_processed)
{
    if (!h) {
        moc_set_error("moc_two_way_sweep: null handle");
        return MOC_EINVAL;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const long long pairs = h->T2 * h->P;
    const int axial_exp = h->I.axial_exp;
    if (axial_exp != 2 && axial_exp != 0) {
        moc_set_error("moc_two_way_sweep: axial_exp %d (the reference has the quadratic and the flat source, solver.c:748-760)",
                      axial_exp);
        return MOC_EINVAL;
    }
    if (h->G > 512) {
        moc_set_error("moc_two_way_sweep: %d energy groups (at most 512)", h->G);
        return MOC_EINVAL;
    }
    int rc;
    const long launches_before = h->launch_count;
    unsigned long long *const digest_back = h->d.digest + 8;

    // ---- pass 1: segment counts
    WalkParams w = walk_params(h);
    CUDA_TRY(cudaMemsetAsync(h->d.digest, 0, sizeof(unsigned long long) * 5, h->stream));   // digest + flags
    CUDA_TRY(cudaMemsetAsync(digest_back, 0, sizeof(unsigned long long) * 4, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->d.pair_max, 0, sizeof(unsigned int) * (size_t)std::max<long long>(pairs, 1), h->stream));
    launch_two_way_walk<false>(h, w, pairs);
    pair_scan_kernel<<<1, 1024, 0, h->stream>>>(h->d.pair_count, h->d.pair_max, w.Zs, h->d.pair_base, h->d.rec_base, pairs);
    h->launch_count++;
    CUDA_TRY(cudaMemcpyAsync(h->pair_base_pinned, h->d.pair_base, sizeof(unsigned long long) * (size_t)(pairs + 1),
                             cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->pair_base_pinned + pairs + 1, h->d.rec_base, sizeof(unsigned long long) * (size_t)(pairs + 1),
                             cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(&h->walk_flags_host, w.flags, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    if (h->walk_flags_host & 2u) {
        // before any record is sized by these counts
        h->walk_flags_host = 0;
        CUDA_TRY(cudaMemsetAsync(w.flags, 0, sizeof(unsigned int), h->stream));
        moc_set_error("moc_two_way_sweep: a ray made more than %d steps inside one 2D segment (the reference would not "
                      "return from this input)", TWO_WAY_GUARD);
        return MOC_EINVAL;
    }
    const unsigned long long *base = h->pair_base_pinned + pairs + 1;
    const unsigned long long total = h->pair_base_pinned[pairs];

    // ---- batches of whole stacks
    unsigned long long largest_pair = 0;
    for (long long p = 0; p < pairs; p++) largest_pair = std::max(largest_pair, base[p + 1] - base[p]);
    if (largest_pair >= (1ull << 32)) {
        moc_set_error("a single z-stack needs %llu record slots (> 2^32)", largest_pair);
        return MOC_EINVAL;
    }
    long long cap = h->batch_segments;
    if (cap <= 0) {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        free_b += (size_t)h->rec_capacity * 12;
        cap = (long long)((double)free_b * 0.7 / 12.0);
    }
    cap = std::max<long long>(cap, (long long)largest_pair);
    cap = std::min<long long>(cap, (1ll << 32) - 1);
    const long long slot = (long long)std::min<unsigned long long>(base[pairs], (unsigned long long)cap);
    if ((rc = ensure_record_capacity(h, std::max<long long>(slot, 1)))) return rc;
    w = walk_params(h);

    AttenuateParams a = attenuate_params(h, w);
    a.coef = nullptr;   // the fit is done per segment
    TwoWayParams tw;
    tw.z_height = h->d.z_height;
    tw.z_sep = h->I.axial_z_sep;
    tw.dz_fine = w.dz_fine;
    tw.axial_exp = axial_exp;
    tw.digest_back = h->want_digest ? digest_back : nullptr;
    tw.flags = w.flags;

    long n_batches = 0;
    for (long long p = 0; p < pairs;) {
        const unsigned long long lim = base[p] + (unsigned long long)slot;
        long long q = (long long)(std::upper_bound(base + p, base + pairs + 1, lim) - base) - 1;
        if (q <= p) q = p + 1;
        w.first_pair = p;
        w.batch_first_record = base[p];
        launch_two_way_walk<true>(h, w, q - p);
        a.batch_first_record = base[p];
        a.first_track = p * h->Z;
        a.end_track = q * h->Z;
        const bool flat = axial_exp == 0, sfu = h->exp_mode == 1;
        if (!flat && !sfu) rc = launch_two_way_attenuate<6, false>(h, a, tw, a.end_track - a.first_track);
        else if (!flat) rc = launch_two_way_attenuate<2, false>(h, a, tw, a.end_track - a.first_track);
        else if (!sfu) rc = launch_two_way_attenuate<6, true>(h, a, tw, a.end_track - a.first_track);
        else rc = launch_two_way_attenuate<2, true>(h, a, tw, a.end_track - a.first_track);
        if (rc) return rc;
        n_batches++;
        p = q;
    }
    h->walk_flags_host = 0;
    CUDA_TRY(cudaMemsetAsync(w.flags, 0, sizeof(unsigned int), h->stream));   // bit 2 (a ray below the node) is informational
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    memset(&h->timing, 0, sizeof h->timing);   // no phase timers on this path
    h->timing.n_batches = n_batches;
    h->timing.launches = h->launch_count - launches_before;
    // solver.c:752, 826: only the quadratic branch counts, once per pass
    const long counted = axial_exp == 2 ? (long)(2 * total) : 0;
    h->I.segments_processed = counted;
    h->rand_base += total;   // one rand() per forward segment (solver.c:745)
    if (segments_processed) *segments_processed = counted;
    return MOC_OK;
}
