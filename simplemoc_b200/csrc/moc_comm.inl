/* moc_comm.inl -- part of moc_device.cu (one translation unit; included there, in this order):
 * boundary exchange (comms.c:5-196): NCCL loaded lazily, the exchange itself, the scalar all-reduces. */
// ------------------------------------------------------------------ communication (NCCL, loaded lazily)

// Minimal NCCL surface, resolved with dlopen so that single-GPU use has no NCCL dependency
// and so that the library shares whichever libnccl the host process already loaded.
typedef struct { char internal[128]; } nccl_unique_id;
typedef int (*nccl_get_unique_id_t)(nccl_unique_id *);
typedef int (*nccl_comm_init_rank_t)(void **, int, nccl_unique_id, int);
typedef int (*nccl_comm_destroy_t)(void *);
typedef int (*nccl_send_t)(const void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_recv_t)(void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_all_reduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_group_t)(void);
typedef const char *(*nccl_error_string_t)(int);

static struct {
    void *lib = nullptr;
    nccl_get_unique_id_t get_unique_id = nullptr;
    nccl_comm_init_rank_t comm_init_rank = nullptr;
    nccl_comm_destroy_t comm_destroy = nullptr;
    nccl_send_t send = nullptr;
    nccl_recv_t recv = nullptr;
    nccl_all_reduce_t all_reduce = nullptr;
    nccl_group_t group_start = nullptr, group_end = nullptr;
    nccl_error_string_t error_string = nullptr;
} g_nccl;

static int load_nccl()
{
    if (g_nccl.lib) return MOC_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) {
        moc_set_error("cannot dlopen libnccl.so.2: %s", dlerror());
        return MOC_ECOMM;
    }
#define MOC_SYM(field, name)                                                  \
    g_nccl.field = (decltype(g_nccl.field))dlsym(g_nccl.lib, name);           \
    if (!g_nccl.field) {                                                      \
        moc_set_error("libnccl lacks %s", name);                              \
        return MOC_ECOMM;                                                     \
    }
    MOC_SYM(get_unique_id, "ncclGetUniqueId")
    MOC_SYM(comm_init_rank, "ncclCommInitRank")
    MOC_SYM(comm_destroy, "ncclCommDestroy")
    MOC_SYM(send, "ncclSend")
    MOC_SYM(recv, "ncclRecv")
    MOC_SYM(all_reduce, "ncclAllReduce")
    MOC_SYM(group_start, "ncclGroupStart")
    MOC_SYM(group_end, "ncclGroupEnd")
    MOC_SYM(error_string, "ncclGetErrorString")
#undef MOC_SYM
    return MOC_OK;
}

#define NCCL_TRY(expr)                                                                         \
    do {                                                                                       \
        int res__ = (expr);                                                                    \
        if (res__ != 0) {                                                                      \
            moc_set_error("%s failed: %s", #expr, g_nccl.error_string ? g_nccl.error_string(res__) : "?"); \
            return MOC_ECOMM;                                                                  \
        }                                                                                      \
    } while (0)

extern "C" int moc_comm_get_unique_id(char id_out[128])
{
    int rc = load_nccl();
    if (rc) return rc;
    nccl_unique_id id;
    NCCL_TRY(g_nccl.get_unique_id(&id));
    memcpy(id_out, id.internal, 128);
    return MOC_OK;
}

extern "C" int moc_comm_init(moc_handle *h, int nranks, int rank, const char id_in[128])
{
    if (!h || nranks < 1 || rank < 0 || rank >= nranks) {
        moc_set_error("moc_comm_init: bad arguments");
        return MOC_EINVAL;
    }
    int rc = load_nccl();
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(h->device));
    nccl_unique_id id;
    memcpy(id.internal, id_in, 128);
    comm_release(h);   // a second moc_comm_init replaces the communicator
    NCCL_TRY(g_nccl.comm_init_rank(&h->nccl_comm, nranks, id, rank));
    h->nranks = nranks;
    h->rank = rank;
    return MOC_OK;
}

static void comm_release(moc_handle *h)
{
    if (h->nccl_comm && g_nccl.comm_destroy) g_nccl.comm_destroy(h->nccl_comm);
    h->nccl_comm = nullptr;
}

static int allreduce_scalars(moc_handle *h, float *dev, int count)
{
    if (!h->nccl_comm) {
        moc_set_error("multi-rank reduction without moc_comm_init");
        return MOC_ECOMM;
    }
    NCCL_TRY(g_nccl.all_reduce(dev, dev, (size_t)count, /*ncclFloat*/ 7, /*ncclSum*/ 0, h->nccl_comm, h->stream));
    return MOC_OK;
}

// Receive staging of the boundary exchange: n_recv chunks.  On problems that fill the HBM (SURVEY config 5:
// 129 GB of flux, 32 GB of staging at 2x2x2) the segment-record buffers of the last sweep may be in the
// way: they are scratch, so they are given back and the allocation is tried again.
static int ensure_exchange_stage(moc_handle *h, long n_recv, long long chunk)
{
    if (n_recv <= h->stage_chunks) return MOC_OK;
    if (h->recv_stage) cudaFree(h->recv_stage);
    h->recv_stage = nullptr;
    h->stage_chunks = 0;
    const size_t bytes = sizeof(float) * (size_t)n_recv * (size_t)chunk;
    if (cudaMalloc((void **)&h->recv_stage, bytes) != cudaSuccess) {
        cudaGetLastError();
        h->recv_stage = nullptr;
        if (h->d.rec_ds) cudaFree(h->d.rec_ds);
        if (h->d.rec_zin) cudaFree(h->d.rec_zin);
        if (h->d.rec_code) cudaFree(h->d.rec_code);
        h->d.rec_ds = h->d.rec_zin = nullptr;
        h->d.rec_code = nullptr;
        h->rec_capacity = 0;
        CUDA_TRY(cudaMalloc((void **)&h->recv_stage, bytes));
    }
    h->stage_chunks = n_recv;
    return MOC_OK;
}

// receives of one exchange under `grid` (chunks that arrive from a neighbour) and the chunk size in floats
static long exchange_receives(moc_handle *h, const CommGrid *grid, long long *chunk)
{
    const long n_ops = moc_exchange_plan(&h->I, grid, nullptr, 0);
    if (n_ops <= 0) return n_ops;
    std::vector<moc_exchange_op> ops((size_t)n_ops);
    moc_exchange_plan(&h->I, grid, ops.data(), n_ops);
    long n_recv = 0;
    for (const moc_exchange_op &op : ops) n_recv += op.recv_from >= 0;
    *chunk = ops[0].count;
    return n_recv;
}

// fast_transfer_boundary_fluxes (comms.c:5-196) on the device, driven by the host schedule
// moc_exchange_plan() (moc_host.c).  Chunks sit at the head of the flux slab in (round,
// direction) order.  Border faces: the chunk's pairwise sum goes to the leakage, zeros come
// back.  Interior faces: ncclSend of the chunk to send_to, ncclRecv from recv_from into a
// staging buffer (a chunk is sent and overwritten at the same offset, so it cannot be
// received in place), scattered back over the same offsets after the group.
// Three kernels + one NCCL group per call, whatever the number of chunks.
static int exchange_on_stream(moc_handle *h, const CommGrid *grid, cudaStream_t st)
{
    const long n_ops = moc_exchange_plan(&h->I, grid, nullptr, 0);
    if (n_ops < 0) return (int)n_ops;
    if (n_ops == 0) return MOC_OK;
    std::vector<moc_exchange_op> ops((size_t)n_ops);
    moc_exchange_plan(&h->I, grid, ops.data(), n_ops);
    const long long chunk = ops[0].count;
    if (chunk % 4 != 0) {
        moc_set_error("exchange chunk of %lld floats is not a multiple of 4", chunk);
        return MOC_EINVAL;
    }
    // device-side tables: [0,n) destination offsets (float4 units), [n,2n) staging offsets or -1,
    // [2n, 2n+n_border) offsets (floats) of the chunks that leak
    std::vector<long long> tab((size_t)3 * n_ops);
    long n_border = 0, n_recv = 0;
    bool any_peer = false;
    for (long k = 0; k < n_ops; k++) {
        tab[(size_t)k] = ops[(size_t)k].offset / 4;
        if (ops[(size_t)k].recv_from >= 0) tab[(size_t)(n_ops + k)] = (n_recv++) * (chunk / 4);
        else tab[(size_t)(n_ops + k)] = -1;
        if (ops[(size_t)k].send_to < 0) tab[(size_t)(2 * n_ops + n_border++)] = ops[(size_t)k].offset;
        any_peer = any_peer || ops[(size_t)k].send_to >= 0 || ops[(size_t)k].recv_from >= 0;
    }
    if (any_peer && !h->nccl_comm) {
        moc_set_error("moc_exchange: neighbours present but moc_comm_init was not called");
        return MOC_ECOMM;
    }
    if (h->exch_capacity < n_ops) {
        if (h->exch_table) cudaFree(h->exch_table);
        if (h->exch_sums) cudaFree(h->exch_sums);
        h->exch_table = nullptr;
        h->exch_sums = nullptr;
        CUDA_TRY(cudaMalloc((void **)&h->exch_table, sizeof(long long) * 3 * (size_t)n_ops));
        h->exch_table_ready = false;
        CUDA_TRY(cudaMalloc((void **)&h->exch_sums, sizeof(float) * (size_t)n_ops));
        h->exch_capacity = n_ops;
    }
    int rc_stage = ensure_exchange_stage(h, n_recv, chunk);
    if (rc_stage) return rc_stage;
    // The tables depend on the grid only: upload once.  (A pageable cudaMemcpyAsync synchronises the
    // host with the stream first -- the overlapped form must not wait for the boundary sweep here.)
    if (!h->exch_table_ready || memcmp(&h->exch_grid, grid, sizeof(CommGrid)) != 0 || h->exch_table_ops != n_ops) {
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemcpy(h->exch_table, tab.data(), sizeof(long long) * 3 * (size_t)n_ops, cudaMemcpyHostToDevice));
        h->exch_grid = *grid;
        h->exch_table_ops = n_ops;
        h->exch_table_ready = true;
    }
    // 1) leakage of the border faces, in the reference's accumulation order
    if (n_border > 0) {
        border_chunk_sums_kernel<<<(unsigned)n_border, 256, 0, st>>>(h->d.psi, h->exch_table + 2 * n_ops, chunk,
                                                                    h->exch_sums);
        leakage_accumulate_kernel<<<1, 1, 0, st>>>(h->exch_sums, (int)n_border, h->d.leakage);
        h->launch_count += 2;
    }
    // 2) every send and receive of every round in one NCCL group.  The reference tags messages
    //    with the direction; here the per-peer FIFO order -- (round, direction) on both sides --
    //    pairs them up.
    if (any_peer) {
        NCCL_TRY(g_nccl.group_start());
        long r = 0;
        for (long k = 0; k < n_ops; k++) {
            const moc_exchange_op &op = ops[(size_t)k];
            if (op.send_to >= 0)
                NCCL_TRY(g_nccl.send(h->d.psi + op.offset, (size_t)chunk, /*ncclFloat*/ 7, op.send_to, h->nccl_comm, st));
            if (op.recv_from >= 0)
                NCCL_TRY(g_nccl.recv(h->recv_stage + (size_t)(r++) * (size_t)chunk, (size_t)chunk, 7, op.recv_from,
                                     h->nccl_comm, st));
        }
        NCCL_TRY(g_nccl.group_end());
    }
    // 3) received chunks (or zeros) replace the sent ones
    {
        const dim3 grid3(16, (unsigned)n_ops);
        exchange_scatter_kernel<<<grid3, 256, 0, st>>>(reinterpret_cast<float4 *>(h->d.psi),
                                                      reinterpret_cast<const float4 *>(h->recv_stage), h->exch_table,
                                                      h->exch_table + n_ops, chunk / 4);
        h->launch_count++;
    }
    CUDA_TRY(cudaGetLastError());
    return MOC_OK;
}

extern "C" int moc_exchange(moc_handle *h, const CommGrid *grid)
{
    if (!h || !grid) return MOC_EINVAL;
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = exchange_on_stream(h, grid, h->stream);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    return MOC_OK;
}

// init.c:162-225 generalised from the hard-coded {2,2,1} to cx*cy*cz (MPI_Cart_create row-major
// ranks, MPI_Cart_shift neighbours, -1 at the non-periodic border)
extern "C" int moc_make_grid(int cx, int cy, int cz, int rank, CommGrid *g)
{
    if (!g || cx < 1 || cy < 1 || cz < 1 || rank < 0 || rank >= cx * cy * cz) {
        moc_set_error("moc_make_grid: bad grid %dx%dx%d rank %d", cx, cy, cz, rank);
        return MOC_EINVAL;
    }
    const int dims[3] = {cx, cy, cz};
    const int at[3] = {rank / (cy * cz), (rank / cz) % cy, rank % cz};
    auto rank_of = [&](int a, int delta) {
        int c[3] = {at[0], at[1], at[2]};
        c[a] += delta;
        if (c[a] < 0 || c[a] >= dims[a]) return -1;
        return (c[0] * cy + c[1]) * cz + c[2];
    };
    int *pos_src[3] = {&g->x_pos_src, &g->y_pos_src, &g->z_pos_src};
    int *pos_dest[3] = {&g->x_pos_dest, &g->y_pos_dest, &g->z_pos_dest};
    int *neg_src[3] = {&g->x_neg_src, &g->y_neg_src, &g->z_neg_src};
    int *neg_dest[3] = {&g->x_neg_dest, &g->y_neg_dest, &g->z_neg_dest};
    for (int a = 0; a < 3; a++) {
        *pos_src[a] = rank_of(a, -1);   // MPI_Cart_shift(+1): receive from below, send up
        *pos_dest[a] = rank_of(a, +1);
        *neg_src[a] = rank_of(a, +1);   // MPI_Cart_shift(-1): receive from above, send down
        *neg_dest[a] = rank_of(a, -1);
    }
    return MOC_OK;
}
