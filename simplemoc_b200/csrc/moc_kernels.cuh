/* moc_kernels.cuh -- hand-written sm_100a kernels of the SimpleMOC hot path.
 *
 *   K0 stack_walk_kernel<KPT, FILL>   axial ray trace of one (2D track, polar angle)
 *                                     z-stack per CTA: segment counts (pass 1) and
 *                                     segment records (pass 2).    solver.c:347-529
 *   K1 attenuate_kernel<...>          multigroup attenuation + scalar-flux tally of
 *                                     the staged segments; lanes = energy groups.
 *                                     solver.c:14-280, 1040-1138, 1441-1464
 *   K2 renormalisation                solver.c:1143-1230
 *   K3 source update                  solver.c:1235-1320
 *   K4 k-effective                    solver.c:1324-1437
 *
 * (reference paths relative to /root/reference/src)
 *
 * Arithmetic contract
 *   - Everything that decides an INTEGER (fine axial interval, segment count,
 *     z-stack window, source-region id) is evaluated with the same IEEE
 *     operations, in the same precision and order, as the serial reference
 *     compiled without FMA contraction: explicit __f*_rn / __d*_rn intrinsics.
 *   - The per-group flux arithmetic is FP32 with FMA contraction and reciprocal
 *     reuse (tolerance contract).
 *   - Reductions reproduce the reference's pairwise_sum tree exactly.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "moc_rng.h"

// minimum resident CTAs per SM the attenuation kernel is compiled for (register budget)
#ifndef MOC_ATT_MIN_BLOCKS
#define MOC_ATT_MIN_BLOCKS 5
#endif

namespace moc {

// ------------------------------------------------------------------ parameters

struct WalkParams {
    // 2D tracks
    const float *seg_len;             // [S2]
    const long long *seg_start;       // [T2+1]
    const int *n_seg;                 // [T2]
    // per polar angle, computed on the host with libm exactly as the reference does
    const double *cos_p;              // cos((double)polar[j])
    const double *sin_p;              // sin((double)polar[j])
    // 3D track state
    float *z_height;                  // [T3]  (written by the FILL pass only)
    uint32_t *seg_count;              // [T3]  written by pass 1, read by pass 2
    unsigned long long *pair_count;   // [T2*P] pass 1 output: 3D segments of the stack
    unsigned int *pair_max;           // [T2*P] pass 1 output: most 3D segments on one ray of the stack
    const unsigned long long *pair_base;  // [T2*P+1] exclusive scan of pair_count (serial segment index)
    const unsigned long long *rec_base;   // [T2*P+1] exclusive scan of Zs * pair_max (record slots)
    // records (pass 2)
    float *rec_ds;
    float *rec_zin;
    uint32_t *rec_code;
    unsigned long long *digest;       // [4] optional (nullptr = off)
    unsigned long long batch_first_record;   // rec_base[first pair of the batch]
    int Zs;                           // record row pitch of a stack: Z rounded up to 8 (32-byte sectors)
    long long first_pair;             // first (i*P+j) of this launch
    int P, Z, fai, axial_exp;
    unsigned int n_regions;
    // x % n_regions and x % fai by multiplication (moc_walk_warp.cuh); mod_fast = 0: hardware remainder
    unsigned int mod_magic, mod_shift, fai_magic;
    int mod_fast;
    // fine intervals without the division sequence (moc_walk_warp.cuh): valid for z in [iv_lo, iv_hi]
    int iv_fast, fine_fast;           // verified for dz_interval (both roundings) / for dz_fine (truncation)
    float iv_rdz, fine_rdz, iv_lo, iv_hi;
    float node_dz_f;                  // node_dz is a float value widened (solver.c:288)
    unsigned int *flags;              // bit 0: a ray height outside [0, node_dz] met the fast intervals
    float z_sep;                      // axial_z_sep
    float dz_interval;                // (float)fine_delta_z : what get_*_interval receive
    double fine_dz, node_dz;          // solver.c:288-289
    float dz_fine;                    // attenuate_fluxes' own float dz (solver.c:38)
    unsigned long long seed, rand_base;
};

struct AttenuateParams {
    const float *rec_ds;
    const float *rec_zin;
    const uint32_t *rec_code;
    const unsigned long long *rec_base;   // [T2*P+1] first record slot of every stack
    unsigned long long batch_first_record;
    int Zs;                           // records of a stack: slot(ray k, segment j) = base + j * Zs + k
    const uint32_t *seg_count;        // [T3]
    const float *p_weight;            // [T3]
    const float *az_weight;           // [T2]
    const float *mu;                  // [P] (float)cos(polar[j])
    float *psi;                       // [T3][2][G]
    const float *fine_source;         // [N][fai][pitch]   rows padded to 128 bytes (pitch = G rounded up to 32)
    const float *coef;                // [N][coef_stencils][3][pitch]  (c0, c1, c2) of solver.c:74-76 per stencil
    int coef_stencils;                // fai - 2: stencil r0 covers source rows r0 .. r0+2
    float *fine_flux;                 // [N][fai][pitch]
    const float *sigT;                // [N][pitch]
    int pitch;
    const float *table;               // [2*table_n]
    float table_dx, table_rdx, table_max, table_half_dx;
    int table_n;
    long long first_track, end_track; // tracks of this batch
    int P, Z, G, fai;
    float inv_2dz, inv_2dz2;          // 1/(2 dz), 1/(2 dz dz) of solver.c:75-76 (only when the fit is done per segment)
    float ds_noclamp;                 // segments no longer than this cannot reach the end of the exponential table with
                                      // any sigT of the slab (0: unknown, every segment keeps the x > maxVal test)
};

// record code: | which:2 | r0:6 | qsr:24 |   (stencil rows r0..r0+2, tally row r0+which)
__host__ __device__ __forceinline__ uint32_t pack_code(uint32_t qsr, uint32_t r0, uint32_t which)
{
    return qsr | (r0 << 24) | (which << 30);
}

// ------------------------------------------------------------------ block scan

// exclusive prefix sum of one 64-bit value per thread across the CTA.
// scratch: 34 x u64 of shared memory.  Returns the prefix; total -> all threads.
__device__ __forceinline__ unsigned long long
block_exclusive_scan(unsigned long long v, unsigned long long *scratch, unsigned long long &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_warps = (blockDim.x + 31) >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < n_warps ? scratch[lane] : 0ull;
        unsigned long long wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long up = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += up;
        }
        scratch[lane] = wi - w;             // exclusive prefix of each warp
        if (lane == 31) scratch[32] = wi;   // grand total
    }
    __syncthreads();
    const unsigned long long prefix = incl - v + scratch[warp];
    total = scratch[32];
    __syncthreads();   // scratch is reused by the caller's next scan
    return prefix;
}

// ------------------------------------------------------------------ K0: axial ray trace

// solver.c:895-907.  Both helpers receive dz narrowed to float.
template <bool UP>
__device__ __forceinline__ int axial_interval(float z, float dz)
{
    const float q = __fdiv_rn(z, dz);
    return UP ? (int)q : (int)ceilf(q);
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// One 2D segment of one z-ray: the while(!seg_complete) loop of solver.c:409-525.
// EMIT=false only counts.  zh is the ray's z_height (updated in place), `home` its
// reset height.  Returns the number of 3D segments; left_domain reports the exit.
template <bool UP, bool EMIT>
__device__ __forceinline__ int walk_segment(const WalkParams &w, float &zh, float home, float s_full,
                                            double cos_p, bool last_2d_segment, bool &left_domain,
                                            unsigned long long serial, unsigned long long slot,
                                            unsigned long long *dg)
{
    float s = s_full;
    int cell = axial_interval<UP>(zh, w.dz_interval);
    int made = 0;
    bool finished = false;
    left_domain = false;
    while (!finished) {
        bool out = false;
        // float z = z_height + s * cos(p_angle)   -- double product, double sum, narrowed
        float z = (float)__dadd_rn((double)zh, __dmul_rn((double)s, cos_p));
        const int cell_to = axial_interval<UP>(z, w.dz_interval);
        float ds;
        if (cell_to == cell) {
            finished = true;
            ds = s;
        } else {
            cell += UP ? 1 : -1;
            z = (float)__dmul_rn(w.fine_dz, (double)(float)cell);
            ds = (float)__ddiv_rn((double)__fsub_rn(z, zh), cos_p);
            s = __fsub_rn(s, ds);
            if (s <= 0.0f) finished = true;
            if (z <= 0.0f || (double)z >= w.node_dz) {
                finished = true;
                out = true;
                left_domain = true;
            }
        }
        if (EMIT) {
            // attenuate_fluxes' view of the start point: solver.c:38-45, 55-58, 84-87
            const float q = __fdiv_rn(zh, w.dz_fine);
            const int iq = (int)q;
            float zin = __fsub_rn(zh, __fmul_rn(w.dz_fine, __fadd_rn((float)iq, 0.5f)));
            const int fine = iq % w.fai;
            const unsigned int qsr =
                moc_rand31(w.seed, w.rand_base + serial + (unsigned)made) % w.n_regions;
            int r0 = fine, which = 0;
            if (w.axial_exp == 2) {
                if (fine == 0) { r0 = 0; zin = __fsub_rn(zin, w.dz_fine); }
                else if (fine == w.fai - 1) { r0 = w.fai - 3; zin = __fadd_rn(zin, w.dz_fine); }
                else r0 = fine - 1;
                which = fine - r0;
            }
            const unsigned long long at = slot + (unsigned long long)made * w.Zs;   // segment-major inside the stack
            w.rec_ds[at] = ds;
            w.rec_zin[at] = zin;
            w.rec_code[at] = pack_code(qsr, (uint32_t)r0, (uint32_t)which);
            if (dg) {
                const unsigned long long m = serial + (unsigned)made;
                const unsigned long long row = (unsigned long long)qsr * w.fai + fine;
                dg[0] += 1ull;
                dg[1] += row;
                dg[2] += (row + 1ull) * (2ull * m + 1ull);
                dg[3] ^= mix64(m * 0x100000001B3ULL + row);
            }
        }
        made++;
        zh = (last_2d_segment || out) ? home : z;
    }
    return made;
}

// CTA = one (2D track i, polar angle j) z-stack; thread owns KPT consecutive z-rays.
// Reproduces the moving [begin_stacked, end_stacked) window of solver.c:376-377,389,466-469
// with prefix sums over the stack.
template <int KPT, bool FILL>
__global__ void stack_walk_kernel(const WalkParams w)
{
    __shared__ unsigned long long scratch[34];
    __shared__ unsigned long long step_total;

    const long long pair = w.first_pair + blockIdx.x;
    const long long i = pair / w.P;
    const int j = (int)(pair % w.P);
    const bool up = j < w.P / 2;
    const int n_seg = w.n_seg[i];
    const float *len = w.seg_len + w.seg_start[i];
    const double cos_p = w.cos_p[j], sin_p = w.sin_p[j];
    const long long t0 = pair * w.Z;
    const int k0 = threadIdx.x * KPT;

    float zh[KPT], home[KPT];
    uint32_t made_total[KPT];
    unsigned long long cursor[KPT];
    unsigned long long dg[4] = {0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < KPT; r++) {
        const int k = k0 + r;
        made_total[r] = 0;
        cursor[r] = 0;
        if (k < w.Z) {
            zh[r] = w.z_height[t0 + k];
            home[r] = up ? __fmul_rn(w.z_sep, (float)k) : __fmul_rn(w.z_sep, (float)(k + 1));
        } else {
            zh[r] = 0.f;
            home[r] = 0.f;
        }
    }

    unsigned long long serial_at = 0;   // serial index of the first segment of this step
    if (FILL) {
        // records of a stack are segment-major: slot(ray k, its j-th segment) = base + j * Zs + k, so
        // the rays of a stack write neighbouring words at every step (full 32-byte sectors)
        serial_at = w.pair_base[pair];
        const unsigned long long base = w.rec_base[pair] - w.batch_first_record;
#pragma unroll
        for (int r = 0; r < KPT; r++) cursor[r] = base + k0 + r;
    }

    int lo = 0, hi = w.Z;
    for (int n = 0; n < n_seg; n++) {
        // s_full = length / sin(p_angle): float / double -> double -> float  (solver.c:382-383)
        const float s_full = (float)__ddiv_rn((double)len[n], sin_p);
        const bool last = (n == n_seg - 1);

        // ---- phase 1: every ray still in the window walks tentatively (no side effects)
        uint32_t cnt[KPT];
        uint32_t exits[KPT];
        float z_after[KPT];
        unsigned long long mine = 0;
#pragma unroll
        for (int r = 0; r < KPT; r++) {
            const int k = k0 + r;
            cnt[r] = 0;
            exits[r] = 0;
            z_after[r] = zh[r];
            if (k >= lo && k < hi) {
                bool left;
                float z = zh[r];
                cnt[r] = up ? walk_segment<true, false>(w, z, home[r], s_full, cos_p, last, left, 0, 0, nullptr)
                            : walk_segment<false, false>(w, z, home[r], s_full, cos_p, last, left, 0, 0, nullptr);
                exits[r] = left ? 1u : 0u;
                z_after[r] = z;
                mine += ((unsigned long long)exits[r] << 32) | cnt[r];
            }
        }
        if (threadIdx.x == 0) step_total = 0ull;
        unsigned long long everything;
        unsigned long long before = block_exclusive_scan(mine, scratch, everything);

        // ---- who is really processed: upward rays see end_stacked shrink as lower rays exit
        bool take[KPT];
        {
            unsigned long long run = before;
#pragma unroll
            for (int r = 0; r < KPT; r++) {
                const int k = k0 + r;
                const int exits_below = (int)(run >> 32);
                const bool in_window = (k >= lo && k < hi);
                take[r] = in_window && (!up || (k + exits_below < hi));
                if (up && take[r]) {
                    // the last processed ray publishes the totals of the processed prefix
                    const unsigned long long incl = run + (((unsigned long long)exits[r] << 32) | cnt[r]);
                    const int k1 = k + 1;
                    const bool next_taken = (k1 < hi) && (k1 + (int)(incl >> 32) < hi);
                    if (!next_taken) step_total = incl;
                }
                run += ((unsigned long long)exits[r] << 32) | cnt[r];
            }
        }
        __syncthreads();
        const unsigned long long done = up ? step_total : everything;

        // ---- phase 2: commit
        {
            unsigned long long run = before;
#pragma unroll
            for (int r = 0; r < KPT; r++) {
                if (take[r]) {
                    if (FILL) {
                        bool left;
                        float z = zh[r];
                        const unsigned long long serial = serial_at + (run & 0xffffffffull);
                        if (up) walk_segment<true, true>(w, z, home[r], s_full, cos_p, last, left, serial,
                                                         cursor[r], w.digest ? dg : nullptr);
                        else walk_segment<false, true>(w, z, home[r], s_full, cos_p, last, left, serial,
                                                       cursor[r], w.digest ? dg : nullptr);
                        cursor[r] += (unsigned long long)cnt[r] * w.Zs;
                    }
                    zh[r] = z_after[r];
                    made_total[r] += cnt[r];
                }
                run += ((unsigned long long)exits[r] << 32) | cnt[r];
            }
        }
        if (up) hi -= (int)(done >> 32);
        else lo += (int)(done >> 32);
        serial_at += (done & 0xffffffffull);
        __syncthreads();   // step_total is rewritten next step
    }

    if (FILL) {
#pragma unroll
        for (int r = 0; r < KPT; r++)
            if (k0 + r < w.Z) w.z_height[t0 + k0 + r] = zh[r];
        if (w.digest) {
            // order-independent digest: three sums and one xor, reduced per warp first
#pragma unroll
            for (int q = 0; q < 4; q++) {
                unsigned long long v = dg[q];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d);
                    v = (q == 3) ? (v ^ o) : (v + o);
                }
                if ((threadIdx.x & 31) == 0) {
                    if (q == 3) atomicXor(w.digest + q, v);
                    else atomicAdd(w.digest + q, v);
                }
            }
        }
    } else {
        unsigned long long mine = 0;
#pragma unroll
        for (int r = 0; r < KPT; r++) {
            if (k0 + r < w.Z) w.seg_count[t0 + k0 + r] = made_total[r];
            mine += made_total[r];
        }
        unsigned long long tot;
        block_exclusive_scan(mine, scratch, tot);
        if (threadIdx.x == 0) w.pair_count[pair] = tot;
        // the longest ray decides how many record rows the stack needs
        uint32_t longest = 0;
#pragma unroll
        for (int r = 0; r < KPT; r++) longest = max(longest, made_total[r]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) longest = max(longest, __shfl_xor_sync(0xffffffffu, longest, d));
        if ((threadIdx.x & 31) == 0) atomicMax(w.pair_max + pair, longest);
    }
}

#include "moc_walk_warp.cuh"

// exclusive scans over the stacks, single CTA (n ~ 2e5): pair_count -> pair_base (serial segment
// index of a stack's first segment), Zs * pair_max -> rec_base (its first record slot)
__global__ void pair_scan_kernel(const unsigned long long *count, const unsigned int *longest, int Zs,
                                 unsigned long long *serial_base, unsigned long long *rec_base, long long n)
{
    __shared__ unsigned long long scratch[34];
    const long long per = (n + blockDim.x - 1) / blockDim.x;
    const long long a = (long long)threadIdx.x * per;
    const long long b = (a + per < n) ? a + per : n;
    for (int which = 0; which < 2; which++) {
        unsigned long long mine = 0;
        for (long long e = a; e < b; e++) mine += which ? (unsigned long long)longest[e] * Zs : count[e];
        unsigned long long tot;
        unsigned long long run = block_exclusive_scan(mine, scratch, tot);
        unsigned long long *out = which ? rec_base : serial_base;
        for (long long e = a; e < b; e++) {
            out[e] = run;
            run += which ? (unsigned long long)longest[e] * Zs : count[e];
        }
        if (threadIdx.x == 0) out[n] = tot;
    }
}

// ------------------------------------------------------------------ K1: attenuation
#include "moc_attenuate.cuh"
#include "moc_two_way.cuh"   // two_way_transport_sweep (coverage path)

// ------------------------------------------------------------------ exact pairwise sums

// utils.c:29-45 over a virtual vector f(0..n-1): same tree, same order of additions.
template <class F>
__device__ float pairwise_sum(const F &f, long long lo, long long n)
{
    if (n <= 16) {
        float s = 0.f;
        for (int e = 0; e < (int)n; e++) s = __fadd_rn(s, f(lo + e));
        return s;
    }
    const long long half = n / 2;
    const float left = pairwise_sum(f, lo, half);
    const float right = pairwise_sum(f, lo + half, n - half);
    return __fadd_rn(left, right);
}

struct ArrayReader {
    const float *v;
    __device__ float operator()(long long e) const { return v[e]; }
};

// pairwise_sum of a device array by ONE CTA of 256 threads, bit-identical to the serial
// recursion: the top 8 levels of the tree are spread over the threads (subtrees are
// summed serially, then combined level by level in the recursion's own order).
// blockDim.x must be 256.
__device__ float pairwise_sum_cta(const float *v, long long n, float *slots /*[256] shared*/)
{
    constexpr int D = 8;
    const int t = threadIdx.x;
    long long lo = 0, sz = n;
    int depth = 0;
    for (; depth < D; depth++) {
        if (sz <= 16) break;
        const long long half = sz / 2;
        if ((t >> (D - 1 - depth)) & 1) { lo += half; sz -= half; }
        else sz = half;
    }
    if ((t & ((1 << (D - depth)) - 1)) == 0) {
        ArrayReader rd{v};
        slots[t] = pairwise_sum(rd, lo, sz);
    }
    __syncthreads();
    for (int level = D - 1; level >= 0; level--) {
        const int span = 1 << (D - level);
        if ((t & (span - 1)) == 0) {
            // size of node (level, t / span): descend `level` steps from the root
            long long s2 = n;
            bool split_all_the_way = true;
            for (int d = 0; d < level; d++) {
                if (s2 <= 16) { split_all_the_way = false; break; }
                const long long half = s2 / 2;
                s2 = ((t >> (D - 1 - d)) & 1) ? s2 - half : half;
            }
            if (split_all_the_way && s2 > 16) slots[t] = __fadd_rn(slots[t], slots[t + span / 2]);
        }
        __syncthreads();
    }
    return slots[0];
}

// ------------------------------------------------------------------ K2..K4

struct SourceParams {
    float *fine_source;     // [N][fai][pitch]
    float *fine_flux;       // [N][fai][pitch]
    int pitch;              // row pitch in floats (G rounded up to 32: 128-byte aligned rows)
    const float *xs;        // [X][G][3]
    const float *scatter;   // [X][G][G]
    const int *xs_index;    // [N]
    const float *vol;       // [N]
    long long N;
    int G, fai;
};

// per-region partial of the fission rate (solver.c:1161-1173) -- one thread per region
__global__ void region_fission_rate_kernel(const SourceParams p, float *per_region)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    const float *flux = p.fine_flux + (size_t)i * p.fai * p.pitch;
    const float *x = p.xs + (size_t)p.xs_index[i] * p.G * 3;
    const float vol = p.vol[i];
    const int G = p.G, W = p.pitch;
    auto per_fine = [&](long long jf) {
        auto per_group = [&](long long g) {
            return __fmul_rn(__fmul_rn(flux[jf * W + g], vol), x[3 * g]);
        };
        return pairwise_sum(per_group, 0, G);
    };
    per_region[i] = pairwise_sum(per_fine, 0, p.fai);
}

// per-region absorption (XS[g][1]) and fission (XS[g][0]) rates (solver.c:1335-1380)
__global__ void region_reaction_rates_kernel(const SourceParams p, float *absorption, float *fission)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    const float *flux = p.fine_flux + (size_t)i * p.fai * p.pitch;
    const float *x = p.xs + (size_t)p.xs_index[i] * p.G * 3;
    const int G = p.G, W = p.pitch;
    for (int col = 0; col < 2; col++) {
        auto per_fine = [&](long long jf) {
            auto per_group = [&](long long g) { return __fmul_rn(x[3 * g + col], flux[jf * W + g]); };
            return pairwise_sum(per_group, 0, G);
        };
        const float r = pairwise_sum(per_fine, 0, p.fai);
        if (col == 0) fission[i] = r;
        else absorption[i] = r;
    }
}

// out[slot] = pairwise_sum(v[0..n))   -- <<<1, 256>>>
__global__ void pairwise_reduce_kernel(const float *v, long long n, float *out, int slot)
{
    __shared__ float slots[256];
    const float s = pairwise_sum_cta(v, n, slots);
    if (threadIdx.x == 0) out[slot] = s;
}

// fine_flux *= norm * 4 pi fai / vol   (solver.c:1207-1214); scal[0] = total fission rate
__global__ void scale_flux_kernel(const SourceParams p, const float *scal)
{
    const long long per = (long long)p.fai * p.pitch;   // padding columns hold zeros and stay zero
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.N * per) return;
    const float norm = (float)(1.0 / (double)scal[0]);                     // solver.c:1203
    const float vol = p.vol[e / per];
    const float n4 = __fmul_rn(norm, 4.0f);
    const float adjust = (float)(((double)n4 * 3.14159265358979323846 * (double)p.fai) / (double)vol);
    p.fine_flux[e] = __fmul_rn(p.fine_flux[e], adjust);
}

// every angular flux (forward and backward rows) *= norm   (solver.c:1219-1226)
__global__ void scale_psi_kernel(float4 *psi4, long long n4, float *tail, int n_tail, const float *scal)
{
    const float norm = (float)(1.0 / (double)scal[0]);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += stride) {
        float4 v = psi4[e];
        v.x *= norm; v.y *= norm; v.z *= norm; v.w *= norm;
        psi4[e] = v;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) tail[threadIdx.x] *= norm;
}

// new source of one (region, fine interval) per CTA (solver.c:1259-1302).
// dynamic shared memory: 2*G floats (flux row, residual terms)
__global__ void update_sources_kernel(const SourceParams p, float inverse_k, float *fine_residual)
{
    extern __shared__ float sh[];
    float *phi = sh;
    float *res_g = sh + p.G;
    const long long row = blockIdx.x;          // i * fai + j
    const long long i = row / p.fai;
    const int G = p.G;
    const float *flux = p.fine_flux + (size_t)row * p.pitch;
    float *q = p.fine_source + (size_t)row * p.pitch;
    const int material = p.xs_index[i];
    const float *x = p.xs + (size_t)material * G * 3;
    const float *S = p.scatter + (size_t)material * G * G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) phi[g] = flux[g];
    __syncthreads();
    auto fis_term = [&](long long g) { return __fmul_rn(phi[g], x[3 * g]); };
    const float fission = __fmul_rn(pairwise_sum(fis_term, 0, G), inverse_k);
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const float *Srow = S + (size_t)g * G;
        auto sc_term = [&](long long g2) { return __fmul_rn(Srow[g2], phi[g2]); };
        const float scatter = pairwise_sum(sc_term, 0, G);
        const float chi = x[3 * g + 2];
        const float mix = __fadd_rn(__fmul_rn(fission, chi), scatter);
        const float fresh = (float)((double)mix / (4.0 * 3.14159265358979323846));
        const float old = q[g];
        const float d = __fsub_rn(fresh, old);
        res_g[g] = __fdiv_rn(__fmul_rn(d, d), __fmul_rn(old, old));
        q[g] = fresh;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ArrayReader rd{res_g};
        fine_residual[row] = pairwise_sum(rd, 0, G);
    }
}

// A lane's place in the pairwise_sum tree over n terms when 2^depth lanes share one sum: lane `sub` owns the
// leaf reached by descending with its bits (most significant first; a node of <= 16 terms is a leaf, as in
// utils.c:29-45); split_mask bit l = the node `sub` stands for at level l really has two children.
struct TreeSlot {
    int lo, sz;          // the leaf [lo, lo + sz) -- summed sequentially by its owner
    bool owner;          // this lane computes that leaf (the other lanes below an early leaf idle)
    unsigned split_mask;
};
__device__ __forceinline__ TreeSlot tree_slot(int n, int depth, int sub)
{
    TreeSlot t;
    t.lo = 0; t.sz = n; t.split_mask = 0;
    int level = 0;
    for (; level < depth; level++) {
        if (t.sz <= 16) break;
        t.split_mask |= 1u << level;
        const int half = t.sz / 2;
        if ((sub >> (depth - 1 - level)) & 1) { t.lo += half; t.sz -= half; }
        else t.sz = half;
    }
    t.owner = (sub & ((1 << (depth - level)) - 1)) == 0;
    return t;
}
// combine the leaves of 2^depth consecutive lanes in the recursion's own order: left + right at every node
__device__ __forceinline__ float tree_combine(float v, const TreeSlot &t, int depth, int sub)
{
    for (int level = depth - 1; level >= 0; level--) {
        const int span = 1 << (depth - level);
        const float right = __shfl_down_sync(0xffffffffu, v, span / 2, 1 << depth);
        // the node at `level` above this lane: split iff this lane descended through it
        if ((sub & (span - 1)) == 0 && ((t.split_mask >> level) & 1u)) v = __fadd_rn(v, right);
    }
    return v;
}
__host__ __device__ __forceinline__ int tree_depth(int n)
{
    int d = 0;
    while (n > 16) { n -= n / 2; d++; }
    return d;
}

// update_sources_kernel with every 104-term (G-term) sum spread over 2^depth lanes: each lane sums one leaf
// of <= 16 terms, the leaves are combined with shuffles in the order of the serial recursion -- the same
// additions in the same order (bit-identical), a sixth of the dependent chain.  One CTA of 256 threads per
// (region, fine interval); needs depth <= 5 (G <= 512).  dynamic shared memory: 2*G + 1 floats.
__global__ void update_sources_coop_kernel(const SourceParams p, float inverse_k, float *fine_residual, int depth)
{
    extern __shared__ float sh[];
    float *phi = sh;
    float *res_g = sh + p.G;
    float *fission_s = sh + 2 * p.G;
    const long long row = blockIdx.x;
    const long long i = row / p.fai;
    const int G = p.G;
    const float *flux = p.fine_flux + (size_t)row * p.pitch;
    float *q = p.fine_source + (size_t)row * p.pitch;
    const int material = p.xs_index[i];
    const float *x = p.xs + (size_t)material * G * 3;
    const float *S = p.scatter + (size_t)material * G * G;
    const int lanes = 1 << depth, sub = threadIdx.x & (lanes - 1), group = threadIdx.x >> depth;
    const int n_groups = blockDim.x >> depth;
    const TreeSlot t = tree_slot(G, depth, sub);
    for (int g = threadIdx.x; g < G; g += blockDim.x) phi[g] = flux[g];
    __syncthreads();
    {
        // every group forms the fission sum (whole warps take part in the shuffles), group 0 publishes it
        float v = 0.f;
        if (t.owner)
            for (int e = 0; e < t.sz; e++) v = __fadd_rn(v, __fmul_rn(phi[t.lo + e], x[3 * (t.lo + e)]));
        v = tree_combine(v, t, depth, sub);
        if (group == 0 && sub == 0) *fission_s = __fmul_rn(v, inverse_k);
    }
    __syncthreads();
    const float fission = *fission_s;
    for (int g0 = 0; g0 < G; g0 += n_groups) {       // uniform trip count: the shuffles need whole groups
        const int g = g0 + group;
        const bool live = g < G;
        float v = 0.f;
        if (live && t.owner) {
            const float *Srow = S + (size_t)g * G + t.lo;
            for (int e = 0; e < t.sz; e++) v = __fadd_rn(v, __fmul_rn(Srow[e], phi[t.lo + e]));
        }
        v = tree_combine(v, t, depth, sub);
        if (live && sub == 0) {
            const float chi = x[3 * g + 2];
            const float mix = __fadd_rn(__fmul_rn(fission, chi), v);
            const float fresh = (float)((double)mix / (4.0 * 3.14159265358979323846));
            const float old = q[g];
            const float d = __fsub_rn(fresh, old);
            res_g[g] = __fdiv_rn(__fmul_rn(d, d), __fmul_rn(old, old));
            q[g] = fresh;
        }
    }
    __syncthreads();
    {
        float v = 0.f;
        if (t.owner)
            for (int e = 0; e < t.sz; e++) v = __fadd_rn(v, res_g[t.lo + e]);
        v = tree_combine(v, t, depth, sub);
        if (group == 0 && sub == 0) fine_residual[row] = v;
    }
}

// per_region[i] = pairwise_sum(fine[i*fai .. +fai))
__global__ void region_fold_kernel(const float *fine, long long N, int fai, float *per_region)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    ArrayReader rd{fine + i * fai};
    per_region[i] = pairwise_sum(rd, 0, fai);
}

// ------------------------------------------------------------------ layout helpers

// host Track (40 B AoS, src/SimpleMOC_header.h:100-107) <-> flat device arrays
struct TrackImage {
    float p_weight;
    float z_height;
    long long rank_in, rank_out;
    float *f_psi, *b_psi;
};

__global__ void unpack_tracks_kernel(const TrackImage *img, long long n, float *p_weight, float *z_height)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    p_weight[t] = img[t].p_weight;
    z_height[t] = img[t].z_height;
}

__global__ void patch_tracks_kernel(TrackImage *img, long long n, const float *z_height)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    img[t].z_height = z_height[t];
}

// Largest total cross section of the slab (and whether every value is a finite, non-negative number): what the host
// needs to prove that no optical length sigT x ds of a sweep reaches the end of the exponential table, so that the
// attenuation may drop the reference's x > maxVal test (moc_attenuate.cuh, MODE 3 / 4).
// out[0]: bits of the maximum (non-negative floats order like their bit patterns), out[1]: number of bad values.
__global__ void sigt_range_kernel(const float *sigT, long long n_regions, int G, int pitch, unsigned int *out)
{
    const long long cells = n_regions * G;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned int mx = 0, bad = 0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < cells; e += stride) {
        const float v = sigT[(e / G) * pitch + (e % G)];
        if (!(v >= 0.0f) || v > 3.0e38f) bad++;
        else mx = max(mx, __float_as_uint(v));
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, mx);
        if (bad) atomicAdd(out + 1, bad);
    }
}

// ------------------------------------------------------------------ synthetic problem

// tracks.c:117-148: p_weight = urand() in (i, j, k) order; upward rays start at the bottom of their
// slot, downward rays at the top
__global__ void synth_tracks_kernel(float *p_weight, float *z_height, long long T3, int Z, int P, float z_sep,
                                    unsigned long long seed, unsigned long long first_draw)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < T3; t += stride) {
        const int k = (int)(t % Z), j = (int)((t / Z) % P);
        p_weight[t] = moc_urand(seed, first_draw + (unsigned long long)t);
        z_height[t] = (j < P / 2) ? __fmul_rn(z_sep, (float)k) : __fmul_rn(z_sep, (float)(k + 1));
    }
}

// dst[row * pitch + col] = urand(first_draw + row * cols + col): source.c:45-48, 83-86, 157-160, 170-172
__global__ void synth_rows_kernel(float *dst, long long rows, int cols, int pitch, unsigned long long seed,
                                  unsigned long long first_draw)
{
    const long long n = rows * cols, stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const long long row = e / cols;
        dst[row * pitch + (e - row * cols)] = moc_urand(seed, first_draw + (unsigned long long)e);
    }
}

// source.c:183-198: region 0 takes material 0 without a draw; region i > 0 draws (material, volume)
__global__ void synth_regions_kernel(int *xs_index, float *vol, long long N, long long X, unsigned long long seed,
                                     unsigned long long first_draw)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (i == 0) {
        xs_index[0] = 0;
        vol[0] = moc_urand(seed, first_draw);
    } else {
        xs_index[i] = (int)((long long)moc_rand31(seed, first_draw + 2ull * (unsigned long long)i - 1ull) % X);
        vol[i] = moc_urand(seed, first_draw + 2ull * (unsigned long long)i);
    }
}

// ------------------------------------------------------------------ K5: boundary exchange helpers

// sums[b] = pairwise_sum(slab[offsets[b] .. +n))  -- one CTA of 256 threads per border chunk
// (comms.c:120-121: the flux leaving through a face without a neighbour)
__global__ void border_chunk_sums_kernel(const float *slab, const long long *offsets, long long n, float *sums)
{
    __shared__ float slots[256];
    const float s = pairwise_sum_cta(slab + offsets[blockIdx.x], n, slots);
    if (threadIdx.x == 0) sums[blockIdx.x] = s;
}

// *leakage += sums[0], then sums[1], ...: the reference adds the chunk sums one by one in
// (round, direction) order (comms.c:118-121)  -- <<<1, 1>>>
__global__ void leakage_accumulate_kernel(const float *sums, int n, float *leakage)
{
    float l = *leakage;
    for (int b = 0; b < n; b++) l = __fadd_rn(l, sums[b]);
    *leakage = l;
}

// chunk blockIdx.y of the plan: slab[dst[y] .. +n4 float4) = stage[src[y] ..] or zeros when
// src[y] < 0 (comms.c:146-149,179-181).  n4 = chunk length in float4.
__global__ void exchange_scatter_kernel(float4 *slab, const float4 *stage, const long long *dst,
                                        const long long *src, long long n4)
{
    float4 *out = slab + dst[blockIdx.y];
    const long long from = src[blockIdx.y];
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (from < 0) {
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += stride) out[e] = zero;
    } else {
        const float4 *in = stage + from;
        for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += stride) out[e] = in[e];
    }
}


// ------------------------------------------------------------------ diagnostics: L2 gather probe

// What the L2 delivers for the access pattern of attenuate_kernel, without its arithmetic: groups
// of 8 lanes read 128 contiguous bytes per quad from 3 consecutive pseudo-random source rows and
// the region's sigT row, and (RED) reduce one float4 per lane into the flux slab.  bench.py
// reports the attenuation kernel's L2-level rate against this measured ceiling.
__device__ __forceinline__ uint32_t probe_hash(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <bool RED>
__global__ void __launch_bounds__(128) l2_gather_probe_kernel(const float4 *__restrict__ src, float4 *flux,
                                                              uint32_t n_regions, uint32_t fai, int pitch4, int quads,
                                                              int iters, float4 *sink)
{
    const int lane8 = threadIdx.x & 7;
    const uint32_t track = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    float4 acc = make_float4(0, 0, 0, 0);
    const float4 *sig = src + (size_t)2 * n_regions * fai * pitch4;
    for (int it = 0; it < iters; it++) {
        const uint32_t h = probe_hash(track * 0x9E3779B9u + it);
        const uint32_t region = h % n_regions;
        const uint32_t r0 = (h >> 24) % (fai - 2);
        const float4 *row = src + ((size_t)region * fai + r0) * pitch4;
        for (int v = 0; v < quads; v++) {
            float4 t = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const float4 y = __ldg(row + r * pitch4 + lane8 + 8 * v);
                t.x += y.x; t.y += y.y; t.z += y.z; t.w += y.w;
            }
            const float4 s4 = __ldg(sig + (size_t)region * pitch4 + lane8 + 8 * v);
            t.x += s4.x; t.y += s4.y; t.z += s4.z; t.w += s4.w;
            if (RED) red_add_v4(reinterpret_cast<float *>(flux + ((size_t)region * fai + r0 + 1) * pitch4 + lane8 + 8 * v), t);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
    }
    if (acc.x == 123.456f) sink[0] = acc;
}

}  // namespace moc
