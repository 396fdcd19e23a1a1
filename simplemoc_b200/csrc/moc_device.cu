/* moc_device.cu -- the C-ABI of libmoc_b200.so (include/moc_b200.h PART B1/B2):
 * device mirrors of the reference's slabs, kernel launches, transfers.
 *
 * Data layout in HBM (one problem = one spatial domain = one GPU):
 *   psi        float [T3][2][G]   the reference's flux slab verbatim (forward row, backward
 *                                 row per 3D track, tracks.c:106-138) so the boundary exchange
 *                                 can treat it as the flat array comms.c does
 *   src slab   float fine_source[N][fai][G] | fine_flux[N][fai][G] | sigT[N][G]  (source.c:121-152)
 *   xs         float [X][G][3],  scatter float [X][G][G],  xs_index int [N],  vol float [N]
 *   tracks     SoA: p_weight[T3], z_height[T3] (unpacked on the device from the 40-byte AoS)
 *   2D tracks  SoA: az_weight[T2], n_seg[T2], seg_start[T2+1], seg_len[S2]
 *   sweep scratch: seg_count u32[T3], pair_count/pair_base/rec_base u64[T2*P(+1)], pair_max u32[T2*P],
 *                  and per batch the segment records rec_ds f32[], rec_zin f32[], rec_code u32[]
 *                  (segment-major inside a z-stack: record j of ray k at rec_base[stack] + j*Zs + k)
 *
 * There is no CPU implementation behind any entry point: without a usable CUDA device
 * every compute call fails with MOC_ENODEVICE.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "moc_b200.h"
#include "moc_internal.h"
#include "moc_kernels.cuh"

using namespace moc;

// ------------------------------------------------------------------ small utilities

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t err__ = (expr);                                                           \
        if (err__ != cudaSuccess) {                                                           \
            moc_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, \
                          __LINE__);                                                          \
            return err__ == cudaErrorMemoryAllocation ? MOC_ENOMEM : MOC_ECUDA;               \
        }                                                                                     \
    } while (0)

static int usable_devices()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int moc_device_count(void) { return usable_devices(); }

static std::mutex g_pin_mutex;
static std::unordered_set<void *> g_pinned;

extern "C" void *moc_host_alloc(size_t bytes)
{
    if (bytes == 0) bytes = 1;
    if (usable_devices() > 0) {
        void *p = nullptr;
        if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) {
            memset(p, 0, bytes);
            std::lock_guard<std::mutex> lock(g_pin_mutex);
            g_pinned.insert(p);
            return p;
        }
        cudaGetLastError();
    }
    return calloc(1, bytes);
}

extern "C" void moc_host_free(void *p)
{
    if (!p) return;
    bool pinned = false;
    {
        std::lock_guard<std::mutex> lock(g_pin_mutex);
        pinned = g_pinned.erase(p) > 0;
    }
    if (pinned) cudaFreeHost(p);
    else free(p);
}

// ------------------------------------------------------------------ the handle

struct DeviceBuffers {
    // 2D tracks
    float *az_weight = nullptr;
    int *n_seg = nullptr;
    long long *seg_start = nullptr;
    float *seg_len = nullptr;
    // polar
    double *cos_p = nullptr, *sin_p = nullptr;
    float *mu = nullptr;
    // 3D tracks
    float *p_weight = nullptr, *z_height = nullptr, *psi = nullptr;
    TrackImage *track_image = nullptr;   // only for the drop-in path
    // sources
    float *src = nullptr, *xs = nullptr, *scatter = nullptr, *vol = nullptr, *table = nullptr;
    float *coef = nullptr;               // quadratic fit coefficients of every stencil, rebuilt before each sweep
    int *xs_index = nullptr;
    // sweep scratch
    uint32_t *seg_count = nullptr, *pair_max = nullptr, *rec_code = nullptr;
    unsigned long long *pair_count = nullptr, *pair_base = nullptr, *rec_base = nullptr, *digest = nullptr;
    float *rec_ds = nullptr, *rec_zin = nullptr;
    // reductions
    float *per_region_a = nullptr, *per_region_b = nullptr, *per_fine = nullptr, *scalars = nullptr;
    float *leakage = nullptr;
};

struct moc_handle {
    int device = 0;
    Input I;
    long long T2 = 0, T3 = 0, N = 0, X = 0, S2 = 0;
    int P = 0, Z = 0, G = 0, F = 0;
    int Gp = 0;                    // device row pitch of the source slab: G rounded up to 32 floats (128 B)
    Table table_host;              // values pointer owned by the caller's Params; copied
    float table_dx = 0, table_max = 0;
    int table_n = 0;
    DeviceBuffers d;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    // host-streamed sweep (drop-in transport_sweep on host structures): copies overlap the kernels
    cudaStream_t up_stream = nullptr, down_stream = nullptr;
    std::vector<cudaEvent_t> ev_pool;   // per-batch events, grown on demand
    int stream_chunks = 16;             // z-stack chunks the flux slab travels in
    // options
    int exp_mode = 0;
    unsigned long long seed = 1, rand_base = 0;
    long long batch_segments = 0;  // 0 = choose from free memory
    int source_stride = 48;
    int lanes_override = 0;
    int fast_cell_ok = 0;          // table_cell_check_kernel found no mismatch (see moc_kernels.cuh)
    unsigned int mod_magic = 0, mod_shift = 0;
    int mod_fast = 0;              // fastmod_check_kernel found no mismatch (moc_walk_warp.cuh)
    float max_seg_len = 0.f;
    unsigned int walk_flags_host = 0;
    int iv_fast = 0, fine_fast = 0;   // interval_check_kernel found no mismatch (moc_walk_warp.cuh)
    float iv_lo = 0.f, iv_hi = 0.f;
    int walk_kernel = 0;           // 0 = auto, 1 = one CTA per z-stack, 2 = one warp per z-stack (Z <= 128)
    // ray-trace CTAs per SM resident under the attenuation of the previous batch.  0 = off, the default:
    // measured on the default problem the overlapped sweep takes the same time (400 vs 398-408 ms,
    // profiles/r01_fill_overlap.log) -- the chip runs at its power cap, so hiding one kernel under the
    // other only slows the other down -- but it needs 2/8 of the record memory (6 GB instead of 25 GB)
    int fill_overlap_ctas = 0;
    int fill_batches = 8;          // batches per chunk of z-stacks when the two overlap
    int fit_per_segment = 0;       // diagnostic: 1 = never use the coefficient slab
    cudaStream_t fill_stream = nullptr;
    int n_sm = 0;
    int want_digest = 0;
    // scratch capacity
    long long rec_capacity = 0;
    std::vector<unsigned long long> pair_base_host;
    unsigned long long *pair_base_pinned = nullptr;
    moc_sweep_timing timing;
    mutable long launch_count = 0;   // kernels launched through this handle
    float leakage_host = 0.f;
    // comms
    void *nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    float *recv_stage = nullptr;
    long stage_chunks = 0;
    long long *exch_table = nullptr;
    float *exch_sums = nullptr;
    long exch_capacity = 0;
    CommGrid exch_grid;
    long exch_table_ops = 0;
    bool exch_table_ready = false;
    cudaStream_t comm_stream = nullptr;
};

static int require_device(int device)
{
    const int n = usable_devices();
    if (n <= 0) {
        moc_set_error("no usable CUDA device: libmoc_b200 has no CPU fallback");
        return MOC_ENODEVICE;
    }
    if (device < 0 || device >= n) {
        moc_set_error("device %d out of range (%d visible)", device, n);
        return MOC_EINVAL;
    }
    return MOC_OK;
}

template <class T>
static int dev_alloc(T **p, size_t count)
{
    CUDA_TRY(cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T)));
    return MOC_OK;
}

static void free_buffers(DeviceBuffers &d)
{
    void *all[] = {d.az_weight, d.n_seg, d.seg_start, d.seg_len, d.cos_p, d.sin_p, d.mu, d.p_weight,
                   d.z_height, d.psi, d.track_image, d.src, d.coef, d.xs, d.scatter, d.vol, d.table,
                   d.xs_index, d.seg_count, d.pair_max, d.rec_code, d.pair_count, d.pair_base, d.rec_base,
                   d.digest, d.rec_ds, d.rec_zin, d.per_region_a, d.per_region_b, d.per_fine,
                   d.scalars, d.leakage};
    for (void *p : all)
        if (p) cudaFree(p);
    d = DeviceBuffers();
}


// The source slab on the device keeps the reference's order (fine_source | fine_flux | sigT,
// source.c:121-152) but pads every row of G floats to Gp so that rows start on 128-byte lines.
// rows [row0, row0 + rows) of the slab <-> a dense host array of G-float rows.
static cudaError_t slab_to_device(moc_handle *h, size_t row0, size_t rows, const float *host)
{
    return cudaMemcpy2DAsync(h->d.src + row0 * h->Gp, sizeof(float) * h->Gp, host, sizeof(float) * h->G,
                             sizeof(float) * h->G, rows, cudaMemcpyHostToDevice, h->stream);
}
static cudaError_t slab_to_host(moc_handle *h, size_t row0, size_t rows, float *host)
{
    return cudaMemcpy2DAsync(host, sizeof(float) * h->G, h->d.src + row0 * h->Gp, sizeof(float) * h->Gp,
                             sizeof(float) * h->G, rows, cudaMemcpyDeviceToHost, h->stream);
}

// ------------------------------------------------------------------ host layout checks

// The reference stores everything in a handful of contiguous slabs and hands out
// pointer-rich views.  The device mirror is built from the slabs;
// these checks make sure the host Params really has that shape.
struct HostLayout {
    const Track *tracks;      // [T3]
    float *psi;               // [T3][2][G]
    float *src;               // source slab
    float *xs, *scatter;      // material slabs
    std::vector<int> xs_index;
    std::vector<float> vol;
};

static inline const Source *source_at(const Params *P, long long i, int stride)
{
    return reinterpret_cast<const Source *>(reinterpret_cast<const char *>(P->sources) + (size_t)i * stride);
}

static int inspect_layout(const Input *I, const Params *P, int source_stride, HostLayout &L)
{
    const long long T2 = I->ntracks_2D, T3 = I->ntracks, N = I->n_source_regions_per_node;
    const int Pn = I->n_polar_angles, Z = I->z_stacked, G = I->n_egroups, F = I->fai;
    if (!P->tracks || !P->tracks_2D || !P->sources || !P->polar_angles || !P->expTable.values) {
        moc_set_error("Params has null members");
        return MOC_EINVAL;
    }
    L.tracks = P->tracks[0][0];
    for (long long i = 0; i < T2; i++)
        for (int j = 0; j < Pn; j++)
            if (P->tracks[i][j] != L.tracks + (i * Pn + j) * Z) {
                moc_set_error("tracks[%lld][%d] is not inside one contiguous [T2][P][Z] Track array "
                              "(reference tracks.c:87-104)", i, j);
                return MOC_ELAYOUT;
            }
    L.psi = L.tracks[0].f_psi;
    const long long probe[3] = {0, T3 / 2, T3 - 1};
    for (long long t : probe)
        if (L.tracks[t].f_psi != L.psi + 2 * t * G || L.tracks[t].b_psi != L.psi + (2 * t + 1) * G) {
            moc_set_error("angular flux of track %lld is not at [t][2][G] in one slab "
                          "(reference tracks.c:106-138)", t);
            return MOC_ELAYOUT;
        }
    const Source *s0 = source_at(P, 0, source_stride);
    L.src = s0->fine_source[0];
    L.xs = s0->XS[0];
    L.scatter = s0->scattering_matrix[0];
    L.xs_index.resize((size_t)N);
    L.vol.resize((size_t)N);
    const long long X = N / 8;
    for (long long i = 0; i < N; i++) {
        const Source *s = source_at(P, i, source_stride);
        if (s->fine_source[0] != L.src + i * F * G || s->fine_flux[0] != L.src + (N + i) * F * G ||
            s->sigT != L.src + 2 * N * F * G + i * G ||
            (F > 1 && s->fine_source[1] != s->fine_source[0] + G)) {
            moc_set_error("source region %lld does not view the source|flux|sigT slab "
                          "(reference source.c:121-152); OPENMP builds need MOC_OPT_SOURCE_STRIDE=56", i);
            return MOC_ELAYOUT;
        }
        const long long m = (s->XS[0] - L.xs) / (3 * G);
        const long long m2 = (s->scattering_matrix[0] - L.scatter) / ((long long)G * G);
        if (m < 0 || m >= X || m != m2 || s->XS[0] != L.xs + m * 3 * G) {
            moc_set_error("source region %lld: material pointers are not rows of the XS/scattering "
                          "slabs (reference source.c:31-86,183-193)", i);
            return MOC_ELAYOUT;
        }
        L.xs_index[(size_t)i] = (int)m;
        L.vol[(size_t)i] = s->vol;
    }
    return MOC_OK;
}

// ------------------------------------------------------------------ create / destroy

static int upload_static(moc_handle *h, const Params *P, const HostLayout &L, bool synthetic = false)
{
    const long long T2 = h->T2;
    const int Pn = h->P, G = h->G;
    // 2D tracks -> SoA
    std::vector<float> az((size_t)T2);
    std::vector<int> ns((size_t)T2);
    std::vector<long long> start((size_t)T2 + 1);
    long long total = 0;
    for (long long i = 0; i < T2; i++) {
        az[(size_t)i] = P->tracks_2D[i].az_weight;
        const long n = P->tracks_2D[i].n_segments;
        ns[(size_t)i] = (int)(n > 0 ? n : 0);
        start[(size_t)i] = total;
        total += ns[(size_t)i];
    }
    start[(size_t)T2] = total;
    h->S2 = total;
    h->max_seg_len = 0.f;
    for (long long i = 0; i < T2; i++)
        for (int n = 0; n < ns[(size_t)i]; n++)
            h->max_seg_len = std::max(h->max_seg_len, P->tracks_2D[i].segments[n].length);
    std::vector<float> len((size_t)std::max<long long>(total, 1));
    for (long long i = 0; i < T2; i++)
        for (int n = 0; n < ns[(size_t)i]; n++)
            len[(size_t)(start[(size_t)i] + n)] = P->tracks_2D[i].segments[n].length;
    int rc;
    if ((rc = dev_alloc(&h->d.az_weight, (size_t)T2))) return rc;
    if ((rc = dev_alloc(&h->d.n_seg, (size_t)T2))) return rc;
    if ((rc = dev_alloc(&h->d.seg_start, (size_t)T2 + 1))) return rc;
    if ((rc = dev_alloc(&h->d.seg_len, (size_t)total))) return rc;
    CUDA_TRY(cudaMemcpy(h->d.az_weight, az.data(), sizeof(float) * (size_t)T2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d.n_seg, ns.data(), sizeof(int) * (size_t)T2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d.seg_start, start.data(), sizeof(long long) * ((size_t)T2 + 1), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d.seg_len, len.data(), sizeof(float) * (size_t)total, cudaMemcpyHostToDevice));

    // polar angles: the reference evaluates cos()/sin() of the float angle in double
    // (solver.c:373,383,417,449); do it here with the same libm and ship the doubles.
    std::vector<double> c((size_t)Pn), s((size_t)Pn);
    std::vector<float> mu((size_t)Pn);
    for (int j = 0; j < Pn; j++) {
        const float ang = P->polar_angles[j];
        // explicit widening: in C++ cos(float) would pick the float overload, the reference is C
        c[(size_t)j] = cos((double)ang);
        s[(size_t)j] = sin((double)ang);
        mu[(size_t)j] = (float)cos((double)ang);
    }
    if ((rc = dev_alloc(&h->d.cos_p, (size_t)Pn))) return rc;
    if ((rc = dev_alloc(&h->d.sin_p, (size_t)Pn))) return rc;
    if ((rc = dev_alloc(&h->d.mu, (size_t)Pn))) return rc;
    CUDA_TRY(cudaMemcpy(h->d.cos_p, c.data(), sizeof(double) * (size_t)Pn, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d.sin_p, s.data(), sizeof(double) * (size_t)Pn, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d.mu, mu.data(), sizeof(float) * (size_t)Pn, cudaMemcpyHostToDevice));

    // materials
    if ((rc = dev_alloc(&h->d.xs, (size_t)h->X * G * 3))) return rc;
    if ((rc = dev_alloc(&h->d.scatter, (size_t)h->X * G * G))) return rc;
    if ((rc = dev_alloc(&h->d.xs_index, (size_t)h->N))) return rc;
    if ((rc = dev_alloc(&h->d.vol, (size_t)h->N))) return rc;
    if (!synthetic) {
        CUDA_TRY(cudaMemcpy(h->d.xs, L.xs, sizeof(float) * (size_t)h->X * G * 3, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(h->d.scatter, L.scatter, sizeof(float) * (size_t)h->X * G * G, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(h->d.xs_index, L.xs_index.data(), sizeof(int) * (size_t)h->N, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(h->d.vol, L.vol.data(), sizeof(float) * (size_t)h->N, cudaMemcpyHostToDevice));
    }

    // exponential table (utils.c:48-78), as built by the caller
    h->table_dx = P->expTable.dx;
    h->table_max = P->expTable.maxVal;
    h->table_n = P->expTable.N;
    if ((rc = dev_alloc(&h->d.table, (size_t)2 * h->table_n))) return rc;
    CUDA_TRY(cudaMemcpy(h->d.table, P->expTable.values, sizeof(float) * 2 * (size_t)h->table_n, cudaMemcpyHostToDevice));

    // May the attenuation kernel pick table cells with the 3-instruction division?  Only if it
    // agrees with the IEEE division for every float the table can be asked about.
    {
        unsigned long long *bad = nullptr, bad_host = 1;
        CUDA_TRY(cudaMalloc((void **)&bad, sizeof(unsigned long long)));
        CUDA_TRY(cudaMemset(bad, 0, sizeof(unsigned long long)));
        unsigned int bits_max;
        const float x_max = h->table_max;
        memcpy(&bits_max, &x_max, sizeof bits_max);
        table_cell_check_kernel<<<148 * 16, 256>>>(bits_max, h->table_dx, 1.0f / h->table_dx, 0.5f * h->table_dx, bad);
        CUDA_TRY(cudaMemcpy(&bad_host, bad, sizeof bad_host, cudaMemcpyDeviceToHost));
        cudaFree(bad);
        h->fast_cell_ok = (bad_host == 0 && x_max > 0.f);
    }
    // May the ray-trace kernel reduce rand() draws modulo n_regions with a multiplication?  Only if
    // it agrees with the hardware remainder for every possible draw (all 2^31 of them).
    h->mod_fast = 0;
    if (h->N > 1 && h->N < (1ll << 24)) {
        const uint32_t n = (uint32_t)h->N;
        uint32_t L = 0;
        while ((1ull << L) < n) L++;
        const unsigned long long magic = ((1ull << (31 + L)) / n) + 1ull;
        if (magic <= 0xffffffffull && L >= 1) {
            unsigned long long *bad = nullptr, bad_host = 1;
            CUDA_TRY(cudaMalloc((void **)&bad, sizeof(unsigned long long)));
            CUDA_TRY(cudaMemset(bad, 0, sizeof(unsigned long long)));
            fastmod_check_kernel<<<148 * 16, 256>>>(n, (uint32_t)magic, L - 1, bad);
            CUDA_TRY(cudaMemcpy(&bad_host, bad, sizeof bad_host, cudaMemcpyDeviceToHost));
            cudaFree(bad);
            if (bad_host == 0) {
                h->mod_magic = (uint32_t)magic;
                h->mod_shift = L - 1;
                h->mod_fast = 1;
            }
        }
    }
    return MOC_OK;
}

static WalkParams walk_params(const moc_handle *h);

// May the ray trace compute fine intervals with the FMA quotient (moc_walk_warp.cuh)?  Only if it
// returns the integer of the IEEE division for every float a ray height can take: checked
// exhaustively over [-node_dz, 2 node_dz] (~3e9 floats, a few ms), once per handle.
static int verify_fast_intervals(moc_handle *h)
{
    const WalkParams w = walk_params(h);
    h->iv_fast = h->fine_fast = 0;
    const float hi = 2.0f * (float)w.node_dz, lo = -(float)w.node_dz;
    if (!(hi > 0.f) || !(w.dz_interval > 0.f) || !(w.dz_fine > 0.f)) return MOC_OK;
    if ((double)hi / (double)w.dz_interval >= 4194304.0 || (double)hi / (double)w.dz_fine >= 4194304.0) return MOC_OK;
    unsigned int bits_hi, bits_lo;
    memcpy(&bits_hi, &hi, 4);
    memcpy(&bits_lo, &lo, 4);
    unsigned long long *bad = nullptr, bad_host[3] = {1, 1, 1};
    CUDA_TRY(cudaMalloc((void **)&bad, 3 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(bad, 0, 3 * sizeof(unsigned long long)));
    const int grid = 148 * 16;
    interval_check_kernel<<<grid, 256>>>(0u, bits_hi, w.dz_interval, 1.0f / w.dz_interval, 0, bad);
    interval_check_kernel<<<grid, 256>>>(0u, bits_hi, w.dz_interval, 1.0f / w.dz_interval, 1, bad + 1);
    interval_check_kernel<<<grid, 256>>>(0x80000000u, bits_lo, w.dz_interval, 1.0f / w.dz_interval, 1, bad + 1);
    interval_check_kernel<<<grid, 256>>>(0u, bits_hi, w.dz_fine, 1.0f / w.dz_fine, 0, bad + 2);
    CUDA_TRY(cudaMemcpy(bad_host, bad, sizeof bad_host, cudaMemcpyDeviceToHost));
    cudaFree(bad);
    // no trial height z + s cos(polar) may leave the verified range: s <= max length / min |sin|
    double min_sin = 1.0;
    {
        std::vector<double> sp((size_t)h->P);
        cudaMemcpy(sp.data(), h->d.sin_p, sizeof(double) * (size_t)h->P, cudaMemcpyDeviceToHost);
        for (double v : sp) min_sin = std::min(min_sin, fabs(v));
    }
    const bool advance_ok = min_sin > 0.0 && (double)h->max_seg_len / min_sin <= w.node_dz;
    h->iv_fast = (bad_host[0] == 0 && bad_host[1] == 0 && advance_ok);
    h->fine_fast = (bad_host[2] == 0);
    h->iv_lo = lo;
    h->iv_hi = hi;
    return MOC_OK;
}

static int need_track_image(moc_handle *h)
{
    if (h->d.track_image) return MOC_OK;
    return dev_alloc(&h->d.track_image, (size_t)h->T3);
}

static int upload_mutable(moc_handle *h, const HostLayout &L, bool with_backward_psi)
{
    const size_t T3 = (size_t)h->T3, G = (size_t)h->G;
    int rc_img = need_track_image(h);
    if (rc_img) return rc_img;
    // Track AoS image -> SoA on the device
    CUDA_TRY(cudaMemcpyAsync(h->d.track_image, L.tracks, sizeof(TrackImage) * T3, cudaMemcpyHostToDevice, h->stream));
    const int threads = 256;
    unpack_tracks_kernel<<<(unsigned)((T3 + threads - 1) / threads), threads, 0, h->stream>>>(
        h->d.track_image, (long long)T3, h->d.p_weight, h->d.z_height);
    if (with_backward_psi) {
        CUDA_TRY(cudaMemcpyAsync(h->d.psi, L.psi, sizeof(float) * 2 * T3 * G, cudaMemcpyHostToDevice, h->stream));
    } else {
        // forward rows only: row pitch 2*G floats
        CUDA_TRY(cudaMemcpy2DAsync(h->d.psi, sizeof(float) * 2 * G, L.psi, sizeof(float) * 2 * G,
                                   sizeof(float) * G, T3, cudaMemcpyHostToDevice, h->stream));
    }
    CUDA_TRY(slab_to_device(h, 0, (size_t)(2 * h->F + 1) * (size_t)h->N, L.src));
    CUDA_TRY(cudaGetLastError());
    return MOC_OK;
}

extern "C" int moc_destroy(moc_handle *h)
{
    if (!h) return MOC_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_buffers(h->d);
    for (auto &e : h->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : h->ev_pool)
        if (e) cudaEventDestroy(e);
    if (h->up_stream) cudaStreamDestroy(h->up_stream);
    if (h->down_stream) cudaStreamDestroy(h->down_stream);
    if (h->pair_base_pinned) cudaFreeHost(h->pair_base_pinned);
    if (h->recv_stage) cudaFree(h->recv_stage);
    if (h->exch_table) cudaFree(h->exch_table);
    if (h->exch_sums) cudaFree(h->exch_sums);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    if (h->fill_stream) cudaStreamDestroy(h->fill_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MOC_OK;
}

// synthetic: Params holds only the 2D tracks, the polar angles and the table (moc_build_tracks_2d);
// the 3D-track and source arrays are generated on the device by the caller (moc_create_synthetic)
static int create_common(const Input *I, const Params *P, int device, int source_stride, moc_handle **out,
                         HostLayout &L, bool synthetic = false)
{
    int rc = require_device(device);
    if (rc) return rc;
    if (I->axial_exp != 0 && I->axial_exp != 2) {
        // the reference prints this and exit(1)s from inside the sweep (solver.c:506-511)
        moc_set_error("Error: invalid axial expansion order %d. Please input 0 or 2", I->axial_exp);
        return MOC_EINVAL;
    }
    if (I->axial_exp == 2 && I->fai < 3) {
        moc_set_error("axial_exp=2 needs fai >= 3 (the edge stencil of solver.c:55-112 reads three rows)");
        return MOC_EINVAL;
    }
    if (I->fai > 63 || I->n_source_regions_per_node >= (1 << 24) || I->z_stacked > 16384) {
        moc_set_error("unsupported size: fai=%d (max 63), N=%ld (max 2^24-1), z_stacked=%d (max 16384)",
                      I->fai, I->n_source_regions_per_node, I->z_stacked);
        return MOC_EINVAL;
    }
    if ((double)(2 * I->fai + 1) * (double)I->n_source_regions_per_node * (double)((I->n_egroups + 31) / 32 * 32) >= 4294967296.0) {
        moc_set_error("source slab has more than 2^32 elements (the attenuation kernel indexes it with 32 bits)");
        return MOC_EINVAL;
    }
    if (I->axial_exp == 2 && 3.0 * (double)(I->fai - 2) * (double)I->n_source_regions_per_node *
                                 (double)((I->n_egroups + 31) / 32 * 32) >= 4294967296.0) {
        moc_set_error("fit-coefficient slab has more than 2^32 elements (the attenuation kernel indexes it with 32 bits)");
        return MOC_EINVAL;
    }
    if (!synthetic && (rc = inspect_layout(I, P, source_stride, L))) return rc;
    CUDA_TRY(cudaSetDevice(device));
    moc_handle *h = new moc_handle();
    h->device = device;
    h->I = *I;
    h->T2 = I->ntracks_2D;
    h->T3 = I->ntracks;
    h->N = I->n_source_regions_per_node;
    h->X = h->N / 8;
    h->P = I->n_polar_angles;
    h->Z = I->z_stacked;
    h->G = I->n_egroups;
    h->Gp = (I->n_egroups + 31) / 32 * 32;
    h->F = I->fai;
    h->source_stride = source_stride;
    memset(&h->timing, 0, sizeof h->timing);
    auto fail = [&](int code) {
        moc_destroy(h);
        return code;
    };
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        moc_set_error("cudaStreamCreate failed");
        return fail(MOC_ECUDA);
    }
    for (auto &e : h->ev)
        if (cudaEventCreate(&e) != cudaSuccess) {
            moc_set_error("cudaEventCreate failed");
            return fail(MOC_ECUDA);
        }
    if ((rc = upload_static(h, P, L, synthetic))) return fail(rc);
    if ((rc = verify_fast_intervals(h))) return fail(rc);
    const size_t T3 = (size_t)h->T3, G = (size_t)h->G, N = (size_t)h->N, F = (size_t)h->F;
    const size_t pairs = (size_t)h->T2 * h->P;
    // the 40-byte Track image only exists for problems that live in host structures
    if (!synthetic && (rc = dev_alloc(&h->d.track_image, T3))) return fail(rc);
    if ((rc = dev_alloc(&h->d.p_weight, T3))) return fail(rc);
    if ((rc = dev_alloc(&h->d.z_height, T3))) return fail(rc);
    if ((rc = dev_alloc(&h->d.psi, 2 * T3 * G))) return fail(rc);
    if ((rc = dev_alloc(&h->d.src, (2 * F + 1) * N * (size_t)h->Gp))) return fail(rc);
    cudaMemsetAsync(h->d.src, 0, sizeof(float) * (2 * F + 1) * N * (size_t)h->Gp, h->stream);   // padding columns stay 0
    {
        // K1 gathers per-stencil fit coefficients if what it then gathers from -- coefficients, sigT, scalar
        // flux -- still fits the L2; on larger problems (SURVEY config 5: 518 MB against 380 MB) every gather
        // goes to DRAM anyway and three more rows per region only add traffic, so the fit stays per segment
        int l2_bytes = 0;
        cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, device);
        const double coef_set = (3.0 * (F - 2) + F + 1) * (double)N * h->Gp * sizeof(float);
        if (I->axial_exp == 2 && coef_set <= (double)l2_bytes &&
            (rc = dev_alloc(&h->d.coef, 3 * (size_t)(F - 2) * N * (size_t)h->Gp)))
            return fail(rc);
    }
    if ((rc = dev_alloc(&h->d.seg_count, T3))) return fail(rc);
    if ((rc = dev_alloc(&h->d.pair_count, pairs))) return fail(rc);
    if ((rc = dev_alloc(&h->d.pair_base, pairs + 1))) return fail(rc);
    if ((rc = dev_alloc(&h->d.rec_base, pairs + 1))) return fail(rc);
    if ((rc = dev_alloc(&h->d.pair_max, pairs))) return fail(rc);
    if ((rc = dev_alloc(&h->d.digest, 5))) return fail(rc);   // [4]: ray-trace flags
    if ((rc = dev_alloc(&h->d.per_region_a, N))) return fail(rc);
    if ((rc = dev_alloc(&h->d.per_region_b, N))) return fail(rc);
    if ((rc = dev_alloc(&h->d.per_fine, N * F))) return fail(rc);
    if ((rc = dev_alloc(&h->d.scalars, 8))) return fail(rc);
    if ((rc = dev_alloc(&h->d.leakage, 1))) return fail(rc);
    if (cudaHostAlloc((void **)&h->pair_base_pinned, sizeof(unsigned long long) * 2 * (pairs + 1), cudaHostAllocDefault) != cudaSuccess) {
        moc_set_error("cudaHostAlloc(pair_base) failed");
        return fail(MOC_ENOMEM);
    }
    cudaMemsetAsync(h->d.seg_count, 0, sizeof(uint32_t) * T3, h->stream);
    cudaMemsetAsync(h->d.digest, 0, sizeof(unsigned long long) * 5, h->stream);
    cudaMemsetAsync(h->d.scalars, 0, sizeof(float) * 8, h->stream);
    h->leakage_host = P->leakage ? *P->leakage : 0.f;
    cudaMemcpyAsync(h->d.leakage, &h->leakage_host, sizeof(float), cudaMemcpyHostToDevice, h->stream);
    *out = h;
    return MOC_OK;
}

extern "C" int moc_create(const Input *I, const Params *P, int device, moc_handle **out)
{
    if (!I || !P || !out) {
        moc_set_error("moc_create: null argument");
        return MOC_EINVAL;
    }
    HostLayout L;
    int rc = create_common(I, P, device, 48, out, L);
    if (rc) return rc;
    moc_handle *h = *out;
    if ((rc = upload_mutable(h, L, true))) {
        moc_destroy(h);
        *out = nullptr;
        return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MOC_OK;
}

// the synthetic problem of build_tracks() (init.c:106-159) generated where it is
// used.  The reference fills 13-129 GB of arrays with rand() on one host thread; with the counter
// RNG (include/moc_rng.h) every element knows its own draw number, so the
// device fills them in parallel -- bit-identical to moc_build_tracks + moc_create
// (tests/test_gpu_parity.py::test_device_construction_is_bit_identical).  The 2D tracks, the polar
// angles and the exponential table stay on the host: they are small and depend on the host libm.
extern "C" int moc_create_synthetic(Input *I, unsigned long long seed, int device, moc_handle **out,
                                    unsigned long long *rand_calls)
{
    if (!I || !out) {
        moc_set_error("moc_create_synthetic: null argument");
        return MOC_EINVAL;
    }
    Params lite;
    moc_draw_layout at;
    int rc = moc_build_tracks_2d(I, seed, &lite, &at);
    if (rc) return rc;
    HostLayout L;
    rc = create_common(I, &lite, device, 48, out, L, true);
    Params tmp = lite;
    moc_free_tracks(I, &tmp);   // the handle has copied what it needs
    if (rc) return rc;
    moc_handle *h = *out;
    const long long T3 = h->T3, N = h->N, X = h->X;
    const int G = h->G, F = h->F, threads = 256;
    auto blocks = [&](long long n) { return (unsigned)std::min<long long>((n + threads - 1) / threads, 148ll * 64); };
    synth_tracks_kernel<<<blocks(T3), threads, 0, h->stream>>>(h->d.p_weight, h->d.z_height, T3, h->Z, h->P,
                                                              h->I.axial_z_sep, seed, at.p_weight);
    CUDA_TRY(cudaMemsetAsync(h->d.psi, 0, sizeof(float) * 2 * (size_t)T3 * G, h->stream));   // tracks.c:106-115
    synth_rows_kernel<<<blocks(N * F * G), threads, 0, h->stream>>>(h->d.src, N * F, G, h->Gp, seed, at.fine_source);
    synth_rows_kernel<<<blocks(N * G), threads, 0, h->stream>>>(h->d.src + (size_t)2 * N * F * h->Gp, N, G, h->Gp, seed, at.sigT);
    synth_rows_kernel<<<blocks(X * G * G), threads, 0, h->stream>>>(h->d.scatter, X * G * G, 1, 1, seed, at.scatter);
    synth_rows_kernel<<<blocks(X * G * 3), threads, 0, h->stream>>>(h->d.xs, X * G * 3, 1, 1, seed, at.xs);
    synth_regions_kernel<<<blocks(N), threads, 0, h->stream>>>(h->d.xs_index, h->d.vol, N, X, seed, at.regions);
    h->launch_count += 7;
    h->seed = seed;
    h->rand_base = at.end;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaGetLastError());
    if (rand_calls) *rand_calls = at.end;
    return MOC_OK;
}

extern "C" int moc_set_option(moc_handle *h, int option, long value)
{
    if (!h) return MOC_EINVAL;
    switch (option) {
    case MOC_OPT_EXP_MODE:
        if (value != 0 && value != 1) break;
        h->exp_mode = (int)value;
        return MOC_OK;
    case MOC_OPT_SEED: h->seed = (unsigned long long)value; return MOC_OK;
    case MOC_OPT_RAND_BASE: h->rand_base = (unsigned long long)value; return MOC_OK;
    case MOC_OPT_BATCH_SEGMENTS:
        if (value < 0) break;
        h->batch_segments = value;
        return MOC_OK;
    case MOC_OPT_SOURCE_STRIDE:
        if (value != 48 && value != 56) break;
        h->source_stride = (int)value;
        return MOC_OK;
    case MOC_OPT_LANES_PER_TRACK:
        if (value != 0 && value != 4 && value != 8 && value != 16 && value != 32) break;
        h->lanes_override = (int)value;
        return MOC_OK;
    case MOC_OPT_WALK_KERNEL:
        if (value < 0 || value > 2) break;
        h->walk_kernel = (int)value;
        return MOC_OK;
    case MOC_OPT_STREAM_CHUNKS:
        if (value < 1 || value > 4096) break;
        h->stream_chunks = (int)value;
        return MOC_OK;
    case MOC_OPT_FILL_OVERLAP:
        if (value < 0 || value > 8) break;
        h->fill_overlap_ctas = (int)value;
        return MOC_OK;
    case MOC_OPT_FILL_BATCHES:
        if (value < 2 || value > 1024) break;
        h->fill_batches = (int)value;
        return MOC_OK;
    case 103: h->fit_per_segment = value != 0; return MOC_OK;   // diagnostic: quadratic fit per segment (large-slab path)
    case 100: h->want_digest = value != 0; return MOC_OK;   // MOC_OPT_DIGEST (diagnostic)
    case 102:                                                // diagnostic: 1 = ray trace with IEEE divisions / hardware remainders only
        if (value) h->iv_fast = h->fine_fast = h->mod_fast = 0;
        return MOC_OK;
    case 101:                                                // MOC_OPT_EXACT_DIV (diagnostic): 1 = never use the fast cell selection
        if (value) h->fast_cell_ok = 0;
        return MOC_OK;
    }
    moc_set_error("moc_set_option: bad option %d / value %ld", option, value);
    return MOC_EINVAL;
}

extern "C" long moc_get_option(moc_handle *h, int option)
{
    if (!h) return -1;
    switch (option) {
    case MOC_OPT_EXP_MODE: return h->exp_mode;
    case MOC_OPT_SEED: return (long)h->seed;
    case MOC_OPT_RAND_BASE: return (long)h->rand_base;
    case MOC_OPT_BATCH_SEGMENTS: return (long)h->batch_segments;
    case MOC_OPT_SOURCE_STRIDE: return h->source_stride;
    case MOC_OPT_LANES_PER_TRACK: return h->lanes_override;
    case MOC_OPT_STREAM_CHUNKS: return h->stream_chunks;
    case MOC_OPT_WALK_KERNEL: return h->walk_kernel;
    case MOC_OPT_FILL_OVERLAP: return h->fill_overlap_ctas;
    case MOC_OPT_FILL_BATCHES: return h->fill_batches;
    case 103: return h->fit_per_segment || !h->d.coef;
    case 100: return h->want_digest;
    case 101: return !h->fast_cell_ok;
    case 102: return !(h->iv_fast && h->fine_fast && h->mod_fast);
    }
    return -1;
}

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

// lane mapping of the attenuation kernel for G groups
struct LaneMap {
    int L, NV4, NS;
};
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

static int launch_attenuate(const moc_handle *h, const AttenuateParams &a, long long n_tracks)
{
    if (n_tracks <= 0) return MOC_OK;
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
static int sweep_core(moc_handle *h, long *segments_processed, const HostLayout *io,
                      const CommGrid *overlap_grid = nullptr)
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
        CUDA_TRY(cudaMemcpyAsync(h->d.track_image, io->tracks, sizeof(TrackImage) * (size_t)h->T3,
                                 cudaMemcpyHostToDevice, h->up_stream));
        CUDA_TRY(cudaMemcpy2DAsync(h->d.src, sizeof(float) * h->Gp, io->src, sizeof(float) * G, sizeof(float) * G,
                                   (size_t)(2 * h->F + 1) * (size_t)h->N, cudaMemcpyHostToDevice, h->up_stream));
        CUDA_TRY(cudaEventRecord(e_img, h->up_stream));
        for (size_t c = 0; c < n_chunks; c++) {
            const size_t t0 = (size_t)chunk_first[c] * h->Z, t1 = (size_t)chunk_first[c + 1] * h->Z;
            // forward rows only: row pitch 2*G floats on both sides
            CUDA_TRY(cudaMemcpy2DAsync(h->d.psi + 2 * t0 * G, sizeof(float) * 2 * G, io->psi + 2 * t0 * G,
                                       sizeof(float) * 2 * G, sizeof(float) * G, t1 - t0, cudaMemcpyHostToDevice,
                                       h->up_stream));
            if ((rc = event_at(h, ev_next++, &ev_up[c]))) return rc;
            CUDA_TRY(cudaEventRecord(ev_up[c], h->up_stream));
        }
        CUDA_TRY(cudaStreamWaitEvent(h->stream, e_img, 0));
        const int threads = 256;
        unpack_tracks_kernel<<<(unsigned)((h->T3 + threads - 1) / threads), threads, 0, h->stream>>>(
            h->d.track_image, h->T3, h->d.p_weight, h->d.z_height);
        h->launch_count++;
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

    if (a.coef) {
        // the source only changes between sweeps (update_sources, uploads): fit every stencil once
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

extern "C" int moc_renormalize(moc_handle *h)
{
    if (!h) return MOC_EINVAL;
    CUDA_TRY(cudaSetDevice(h->device));
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
    const long long n = 2 * h->T3 * h->G;
    const long long n4 = n / 4;
    scale_psi_kernel<<<148 * 8, 256, 0, h->stream>>>(reinterpret_cast<float4 *>(h->d.psi), n4, h->d.psi + 4 * n4,
                                                    (int)(n - 4 * n4), h->d.scalars);
    h->launch_count += 4;
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
    update_sources_kernel<<<(unsigned)rows, threads, sizeof(float) * 2 * (size_t)h->G, h->stream>>>(
        p, inverse_k, h->d.per_fine);
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
    if (bytes != n || which == MOC_ARR_SEG_COUNT || which == MOC_ARR_QSR_DIGEST) {
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
    NCCL_TRY(g_nccl.comm_init_rank(&h->nccl_comm, nranks, id, rank));
    h->nranks = nranks;
    h->rank = rank;
    return MOC_OK;
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

// ------------------------------------------------------------------ drop-in entry points (PART B1)

struct Mirror {
    moc_handle *h = nullptr;
    bool dirty_sweep = false;   // device holds newer psi/z/flux than the host
    bool dirty_all = false;     // device holds newer everything
    std::vector<void *> registered;   // host slabs this library page-locked (cudaHostRegister)
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
    } else if (inspect_layout(I, P, m.h->source_stride, L)) {
        die(where);
    }
    if (created) pin_host_slabs(m, L);
    if (created || !g_resident) {
        // caller_streams: the non-resident transport_sweep moves the mutable state itself, chunk by
        // chunk, overlapped with the kernels (sweep_core); nothing to upload here
        if (!(caller_streams && !g_resident) && upload_mutable(m.h, L, created || need_backward_psi)) die(where);
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
