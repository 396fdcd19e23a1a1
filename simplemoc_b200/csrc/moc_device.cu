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
    float *coef4 = nullptr;              // the same + sigT packed per (region, stencil) for the TMA-staged attenuation
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
    int walk_kernel = 0;           // 0 / 2 = warp(s) per z-stack (Z <= 2048), 1 = always one thread per ray
    // ray-trace CTAs per SM resident under the attenuation of the previous batch.  0 = off, the default:
    // measured on the default problem the overlapped sweep takes the same time (400 vs 398-408 ms,
    // profiles/r01_fill_overlap.log) -- the chip runs at its power cap, so hiding one kernel under the
    // other only slows the other down -- but it needs 2/8 of the record memory (6 GB instead of 25 GB)
    int fill_overlap_ctas = 0;
    int fill_batches = 8;          // batches per chunk of z-stacks when the two overlap
    int fit_per_segment = 0;       // diagnostic: 1 = never use the coefficient slab
    int staged = 1;                // 1 (default): TMA-staged attenuation where it applies, 0: direct gathers
    int allow_noclamp = 1;         // 0: always keep the x > maxVal test of the table (A/B timing, option 105)
    bool sigT_known = false;       // sigT_max / sigT_clean describe the slab on the device
    float sigT_max = 0.f;
    bool sigT_clean = false;
    double min_abs_sin = 0.0, min_abs_cos = 0.0;   // over the polar angles
    mutable bool noclamp_now = false;   // this sweep's attenuation launches skip the x > maxVal test
    mutable bool staged_now = false;   // what the current sweep uses
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
                   d.z_height, d.psi, d.track_image, d.src, d.coef, d.coef4, d.xs, d.scatter, d.vol, d.table,
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
    if (row0 + rows > (size_t)2 * h->N * h->F) h->sigT_known = false;   // the copy reaches the sigT rows
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
    h->min_abs_sin = h->min_abs_cos = Pn > 0 ? 1.0 : 0.0;
    for (int j = 0; j < Pn; j++) {
        h->min_abs_sin = std::min(h->min_abs_sin, fabs(s[(size_t)j]));
        h->min_abs_cos = std::min(h->min_abs_cos, fabs(c[(size_t)j]));
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
// lane mapping of the attenuation kernel for G groups (moc_sweep.inl)
struct LaneMap {
    int L, NV4, NS;
};
static LaneMap choose_lanes(int G, int lanes_override);
static void comm_release(moc_handle *h);   // moc_comm.inl: ncclCommDestroy

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
    comm_release(h);
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
    if (I->fai < 1) {
        moc_set_error("unsupported size: fai=%d (min 1)", I->fai);
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
    {
        // what every sweep would otherwise only find out after both ray-trace passes: a lane mapping for G ...
        const LaneMap lm = choose_lanes(I->n_egroups, 0);
        if (I->n_egroups < 1 || 4 * lm.L * lm.NV4 + lm.L * lm.NS < I->n_egroups || lm.NS > 16) {
            moc_set_error("unsupported size: n_egroups=%d (1 .. 512: the attenuation kernel keeps a track's angular flux in "
                          "the registers of at most 32 lanes x 16 groups)", I->n_egroups);
            return MOC_EINVAL;
        }
        // ... and an exponential table that fits the shared memory of an SM (utils.c:55: N = 10 sqrt(1 / (0.08 precision));
        // the default precision 0.01 gives 353 cells = 2.8 KB; 227 KB hold 29 055 cells, i.e. precision >= 1.5e-6)
        if (P->expTable.N < 1 || sizeof(float) * 2 * ((size_t)P->expTable.N + 1) > 227 * 1024) {
            moc_set_error("unsupported size: exponential table of %d cells (Input.precision %g): at most %d cells fit the "
                          "227 KB of shared memory the attenuation kernel can keep it in", P->expTable.N, (double)I->precision,
                          (int)(227 * 1024 / 8 - 1));
            return MOC_EINVAL;
        }
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
    if ((rc = dev_alloc(&h->d.digest, 12))) return fail(rc);   // [4]: ray-trace flags, [5]: sigT range (sweep_core),
                                                                // [8..11]: backward digest of the two-way sweep
    if ((rc = dev_alloc(&h->d.per_region_a, N))) return fail(rc);
    if ((rc = dev_alloc(&h->d.per_region_b, N))) return fail(rc);
    if ((rc = dev_alloc(&h->d.per_fine, N * F))) return fail(rc);
    if ((rc = dev_alloc(&h->d.scalars, 8))) return fail(rc);
    if ((rc = dev_alloc(&h->d.leakage, 1))) return fail(rc);
    if (cudaHostAlloc((void **)&h->pair_base_pinned, sizeof(unsigned long long) * (2 * (pairs + 1) + 1), cudaHostAllocDefault) != cudaSuccess) {
        moc_set_error("cudaHostAlloc(pair_base) failed");
        return fail(MOC_ENOMEM);
    }
    cudaMemsetAsync(h->d.seg_count, 0, sizeof(uint32_t) * T3, h->stream);
    cudaMemsetAsync(h->d.digest, 0, sizeof(unsigned long long) * 12, h->stream);
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
    case 104: h->staged = value != 0; return MOC_OK;            // 0 = attenuation with direct gathers (A/B timing)
    case 105: h->allow_noclamp = value != 0; return MOC_OK;     // 0 = keep the x > maxVal test in every launch (A/B timing)
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
    case 104: return h->staged;
    case 105: return h->noclamp_now;
    case 100: return h->want_digest;
    case 101: return !h->fast_cell_ok;
    case 102: return !(h->iv_fast && h->fine_fast && h->mod_fast);
    }
    return -1;
}

// The rest of this translation unit, in the order it is compiled:
#include "moc_sweep.inl"    // the transport sweep
#include "moc_two_way.inl"  // two_way_transport_sweep
#include "moc_phases.inl"   // renormalise, update_sources, k-eff, array access
#include "moc_comm.inl"     // boundary exchange over NCCL
#include "moc_dropin.inl"   // the reference's names on host structures
