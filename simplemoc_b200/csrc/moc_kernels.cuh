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
#define MOC_ATT_MIN_BLOCKS 1
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
    unsigned long long *pair_count;   // [T2*P] pass 1 output
    const unsigned long long *pair_base;  // [T2*P+1] exclusive scan of pair_count
    // records (pass 2)
    float *rec_ds;
    float *rec_zin;
    uint32_t *rec_code;
    uint32_t *track_off;              // [tracks in batch] first record, batch-relative
    unsigned long long *digest;       // [4] optional (nullptr = off)
    unsigned long long batch_first_record;   // pair_base[first pair of the batch]
    long long first_pair;             // first (i*P+j) of this launch
    int P, Z, fai, axial_exp;
    unsigned int n_regions;
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
    const uint32_t *track_off;        // batch-relative
    const uint32_t *seg_count;        // [T3]
    const float *p_weight;            // [T3]
    const float *az_weight;           // [T2]
    const float *mu;                  // [P] (float)cos(polar[j])
    float *psi;                       // [T3][2][G]
    const float *fine_source;         // [N][fai][G]
    float *fine_flux;                 // [N][fai][G]
    const float *sigT;                // [N][G]
    const float *table;               // [2*table_n]
    float table_dx, table_rdx, table_max, table_half_dx;
    int table_n;
    long long first_track, end_track; // tracks of this batch
    int P, Z, G, fai;
    float inv_2dz, inv_2dz2;          // 1/(2 dz), 1/(2 dz dz) of solver.c:75-76
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
            w.rec_ds[slot + made] = ds;
            w.rec_zin[slot + made] = zin;
            w.rec_code[slot + made] = pack_code(qsr, (uint32_t)r0, (uint32_t)which);
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
        // record layout is track-major: offsets = exclusive scan of the pass-1 counts
        unsigned long long mine = 0;
        uint32_t c[KPT];
#pragma unroll
        for (int r = 0; r < KPT; r++) {
            c[r] = (k0 + r < w.Z) ? w.seg_count[t0 + k0 + r] : 0u;
            mine += c[r];
        }
        unsigned long long tot;
        unsigned long long at = block_exclusive_scan(mine, scratch, tot);
        serial_at = w.pair_base[pair];
        at += serial_at - w.batch_first_record;
#pragma unroll
        for (int r = 0; r < KPT; r++) {
            if (k0 + r < w.Z) {
                w.track_off[(pair - w.first_pair) * w.Z + k0 + r] = (uint32_t)at;
                cursor[r] = at;
            }
            at += c[r];
        }
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
                        cursor[r] += cnt[r];
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
    }
}

// exclusive scan of pair_count[n] into pair_base[n+1]; single CTA (n ~ 2e5).
__global__ void pair_scan_kernel(const unsigned long long *in, unsigned long long *out, long long n)
{
    __shared__ unsigned long long scratch[34];
    const long long per = (n + blockDim.x - 1) / blockDim.x;
    const long long a = (long long)threadIdx.x * per;
    const long long b = (a + per < n) ? a + per : n;
    unsigned long long mine = 0;
    for (long long e = a; e < b; e++) mine += in[e];
    unsigned long long tot;
    unsigned long long run = block_exclusive_scan(mine, scratch, tot);
    for (long long e = a; e < b; e++) {
        out[e] = run;
        run += in[e];
    }
    if (threadIdx.x == 0) out[n] = tot;
}

// ------------------------------------------------------------------ K1: attenuation

// MUFU.RCP / MUFU.EX2 without the range fix-ups of the libdevice wrappers
__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// The table cell the reference picks for x (solver.c:1448): (int)(x / dx + 0.5f * dx), with an
// IEEE float division.  EXACT_DIV: the division instruction sequence of __fdiv_rn.
// !EXACT_DIV: quotient by one Newton step on x * fl(1/dx) (3 instructions); the host only
// selects this variant after table_cell_check_kernel has verified, for EVERY float in
// [0, maxVal], that it lands in the same cell.
template <bool EXACT_DIV>
__device__ __forceinline__ int table_cell(float x, float dx, float rdx, float half_dx)
{
    float q;
    if (EXACT_DIV) {
        q = __fdiv_rn(x, dx);
    } else {
        q = __fmul_rn(x, rdx);
        const float rem = __fmaf_rn(-q, dx, x);
        q = __fmaf_rn(rem, rdx, q);
    }
    return __float2int_rz(__fadd_rn(q, half_dx));
}

// exhaustive check of the fast cell selection: one thread per float bit pattern in [0, bits_max]
__global__ void table_cell_check_kernel(unsigned int bits_max, float dx, float rdx, float half_dx,
                                        unsigned long long *mismatches)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned int bad = 0;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= bits_max; b += stride) {
        const float x = __uint_as_float((unsigned int)b);
        bad += table_cell<true>(x, dx, rdx, half_dx) != table_cell<false>(x, dx, rdx, half_dx);
    }
    if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}

struct TableConsts {
    float dx, rdx, half_dx, x_max;
    int n;   // cells; s_tab[n] = (0, 1): the value of every x > x_max
};

// E = 1 - exp(-x) and D = exp(-x) = 1 - E.
//  MODE 0: the reference's linear table, cell chosen exactly as solver.c:1448 does (the slope sign
//          is the reference's), IEEE division.
//  MODE 1: the same with the verified fast division.
//  MODE 2: SFU: MUFU.EX2.
// All keep the reference's x > maxVal -> 1 rule (solver.c:1444-1445).
template <int MODE>
__device__ __forceinline__ void one_minus_exp(float x, const float2 *tab, const TableConsts &tc, float &E, float &D)
{
    if (MODE == 2) {
        const float d = ex2_approx(x * -1.4426950408889634f);
        const bool big = x > tc.x_max;
        D = big ? 0.0f : d;
        E = 1.0f - D;
    } else {
        int cell = table_cell<MODE == 0>(x, tc.dx, tc.rdx, tc.half_dx);
        cell = x > tc.x_max ? tc.n : cell;
        const float2 line = tab[cell];
        E = fmaf(line.x, x, line.y);
        D = 1.0f - E;
    }
}

// per-segment scalars of attenuate_fluxes, hoisted out of the group loop
struct SegmentScalars {
    float ds;
    float a1, a2;      // zin/(2dz), zin^2/(2dz^2)          : q0 = y2 + a1 (y1-y3) + a2 (y1-2y2+y3)
    float b1, b2;      // mu/(2dz), 2 mu zin/(2dz^2)        : q1 mu
    float b3, b3_3;    // mu^2/(2dz^2), the same / 3         : q2 mu^2
    float weight;
};

// one energy group of attenuate_fluxes (solver.c:66-82, 146-279).  Returns the tally.
// Same formulas, regrouped so that every factor that does not depend on the group is a
// per-segment scalar and every division is a multiplication by MUFU.RCP(sigT).
template <int MODE>
__device__ __forceinline__ float attenuate_quadratic(float y1, float y2, float y3, float sigT, float &psi,
                                                     const SegmentScalars &k, const float2 *tab,
                                                     const TableConsts &tc)
{
    const float d = y1 - y3;
    const float e = fmaf(-2.f, y2, y1 + y3);
    const float q0 = fmaf(k.a2, e, fmaf(k.a1, d, y2));
    const float q1m = fmaf(k.b2, e, k.b1 * d);          // q1 * mu
    const float q2m = k.b3 * e;                         // q2 * mu^2
    const float q2m3 = k.b3_3 * e;                      // q2 * mu^2 / 3
    const float tau = sigT * k.ds;
    float E, D;
    one_minus_exp<MODE>(tau, tab, tc, E, D);
    const float r1 = rcp_approx(sigT);
    const float r2 = r1 * r1;
    const float r3 = r2 * r1;
    const float r4 = r2 * r2;
    // solver.c:175-176 exactly as parenthesised there
    const float reuse = fmaf(2.f, E * r3, tau * (tau - 2.f));
    float in = fmaf(q0, tau, fmaf(sigT, psi, -q0) * E) * r2;
    in = fmaf(q1m, reuse, in);
    const float cubic = fmaf(-6.f, E, tau * fmaf(tau, tau - 3.f, 6.f));
    in = fmaf(q2m3, cubic * r4, in);
    float out = (q0 * E) * r1;
    out = fmaf(q1m * r2, tau - E, out);
    out = fmaf(q2m, reuse, out);
    psi = fmaf(psi, D, out);
    return k.weight * in;
}

// one energy group of attenuate_FSR_fluxes (solver.c:1104-1115)
template <int MODE>
__device__ __forceinline__ float attenuate_flat(float src, float sigT, float &psi, const SegmentScalars &k,
                                                const float2 *tab, const TableConsts &tc)
{
    const float tau = sigT * k.ds;
    float E, D;
    one_minus_exp<MODE>(tau, tab, tc, E, D);
    const float q = __fdiv_rn(src, sigT);   // flat source: the difference psi - q cancels, keep the division exact
    const float dpsi = (psi - q) * E;
    psi -= dpsi;
    return k.weight * dpsi;
}

// fine_flux is only ever reduced into by this kernel (never read), so the reductions carry no
// "memory" clobber: the compiler may hoist the next segment's loads above them.
__device__ __forceinline__ void red_add_v4(float *addr, float4 v)
{
    // sm_90+: one 16-byte reduction instead of four 4-byte ones
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ void red_add(float *addr, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v));
}

// L lanes cooperate on one 3D track (32/L tracks per warp).  Lane `lit` of a track
// owns NV4 float4 group-quads  g = 4*(lit + L*v) ..+3   and NS single groups
// g = 4*L*NV4 + lit + L*s  (so G=104 -> L=8, NV4=3, NS=1 uses every lane fully).
// The angular flux of the track lives in registers for the whole track.
template <int L, int NV4, int NS, int MODE, bool FLAT>
__global__ void __launch_bounds__(128, MOC_ATT_MIN_BLOCKS) attenuate_kernel(const AttenuateParams a)
{
    extern __shared__ float2 s_tab[];
    if (MODE != 2) {
        for (int e = threadIdx.x; e < a.table_n; e += blockDim.x)
            s_tab[e] = make_float2(a.table[2 * e], a.table[2 * e + 1]);
        if (threadIdx.x == 0) s_tab[a.table_n] = make_float2(0.f, 1.f);
        __syncthreads();
    }
    TableConsts tc;
    tc.dx = a.table_dx; tc.rdx = a.table_rdx; tc.half_dx = a.table_half_dx; tc.x_max = a.table_max;
    tc.n = a.table_n;
    constexpr int TPW = 32 / L;
    const int lane = threadIdx.x & 31;
    const int lit = lane % L;
    const int G = a.G;
    const long long t = a.first_track + ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * TPW + lane / L;
    const bool valid = t < a.end_track;

    uint32_t n_rec = 0, at = 0;
    SegmentScalars sc;
    sc.ds = 0.f; sc.a1 = sc.a2 = sc.b1 = sc.b2 = sc.b3 = sc.b3_3 = 0.f; sc.weight = 0.f;
    float mu = 0.f;
    if (valid) {
        n_rec = a.seg_count[t];
        at = a.track_off[t - a.first_track];
        const long long pair = t / a.Z;
        const int j = (int)(pair % a.P);
        const long long i = pair / a.P;
        mu = a.mu[j];
        float w0 = __fmul_rn(a.p_weight[t], a.az_weight[i]);   // solver.c:49
        if (FLAT) w0 = __fmul_rn(w0, mu);                      // solver.c:1064
        sc.weight = w0;
        sc.b1 = mu * a.inv_2dz;
        sc.b3 = mu * mu * a.inv_2dz2;
        sc.b3_3 = sc.b3 * (1.f / 3.f);
    }
    const float two_mu_c = 2.f * mu * a.inv_2dz2;
    unsigned int longest = n_rec;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned int o = __shfl_xor_sync(0xffffffffu, longest, d);
        longest = o > longest ? o : longest;
    }

    // group ownership
    const int g_tail = 4 * L * NV4;
    float4 psi4[NV4 > 0 ? NV4 : 1];
    float psi1[NS > 0 ? NS : 1];
    float *psi_row = a.psi + (size_t)2 * (size_t)(valid ? t : 0) * G;
#pragma unroll
    for (int v = 0; v < NV4; v++) {
        const int g = 4 * (lit + L * v);
        psi4[v] = (valid && g < G) ? *reinterpret_cast<const float4 *>(psi_row + g) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int g = g_tail + lit + L * s;
        psi1[s] = (valid && g < G) ? psi_row[g] : 0.f;
    }

    // the record of the next segment is fetched one iteration ahead
    const float *rec_ds = a.rec_ds + at;
    const float *rec_zin = a.rec_zin + at;
    const uint32_t *rec_code = a.rec_code + at;
    float next_ds = 0.f, next_zin = 0.f;
    uint32_t next_code = 0;
    if (n_rec > 0) {
        next_ds = __ldg(rec_ds);
        next_zin = __ldg(rec_zin);
        next_code = __ldg(rec_code);
    }

    for (unsigned int sgm = 0; sgm < longest; sgm++) {
        if (sgm < n_rec) {
            sc.ds = next_ds;
            const float zin = next_zin;
            const uint32_t code = next_code;
            if (sgm + 1 < n_rec) {
                next_ds = __ldg(rec_ds + sgm + 1);
                next_zin = __ldg(rec_zin + sgm + 1);
                next_code = __ldg(rec_code + sgm + 1);
            }
            sc.a1 = zin * a.inv_2dz;
            sc.a2 = zin * zin * a.inv_2dz2;
            sc.b2 = two_mu_c * zin;
            const uint32_t qsr = code & 0xffffffu;
            const uint32_t r0 = (code >> 24) & 63u;
            const uint32_t which = code >> 30;
            const size_t row0 = (size_t)qsr * a.fai + r0;
            const float *ya = a.fine_source + row0 * G;
            const float *st = a.sigT + (size_t)qsr * G;
            float *fl = a.fine_flux + (row0 + which) * G;
#pragma unroll
            for (int v = 0; v < NV4; v++) {
                const int g = 4 * (lit + L * v);
                if (g < G) {
                    const float4 s4 = __ldg(reinterpret_cast<const float4 *>(st + g));
                    float4 tally;
                    if (FLAT) {
                        const float4 y = __ldg(reinterpret_cast<const float4 *>(ya + g));
                        tally.x = attenuate_flat<MODE>(y.x, s4.x, psi4[v].x, sc, s_tab, tc);
                        tally.y = attenuate_flat<MODE>(y.y, s4.y, psi4[v].y, sc, s_tab, tc);
                        tally.z = attenuate_flat<MODE>(y.z, s4.z, psi4[v].z, sc, s_tab, tc);
                        tally.w = attenuate_flat<MODE>(y.w, s4.w, psi4[v].w, sc, s_tab, tc);
                    } else {
                        const float4 y1 = __ldg(reinterpret_cast<const float4 *>(ya + g));
                        const float4 y2 = __ldg(reinterpret_cast<const float4 *>(ya + G + g));
                        const float4 y3 = __ldg(reinterpret_cast<const float4 *>(ya + 2 * G + g));
                        tally.x = attenuate_quadratic<MODE>(y1.x, y2.x, y3.x, s4.x, psi4[v].x, sc, s_tab, tc);
                        tally.y = attenuate_quadratic<MODE>(y1.y, y2.y, y3.y, s4.y, psi4[v].y, sc, s_tab, tc);
                        tally.z = attenuate_quadratic<MODE>(y1.z, y2.z, y3.z, s4.z, psi4[v].z, sc, s_tab, tc);
                        tally.w = attenuate_quadratic<MODE>(y1.w, y2.w, y3.w, s4.w, psi4[v].w, sc, s_tab, tc);
                    }
                    red_add_v4(fl + g, tally);
                }
            }
#pragma unroll
            for (int s = 0; s < NS; s++) {
                const int g = g_tail + lit + L * s;
                if (g < G) {
                    const float s1 = __ldg(st + g);
                    float tally;
                    if (FLAT) {
                        tally = attenuate_flat<MODE>(__ldg(ya + g), s1, psi1[s], sc, s_tab, tc);
                    } else {
                        tally = attenuate_quadratic<MODE>(__ldg(ya + g), __ldg(ya + G + g), __ldg(ya + 2 * G + g),
                                                          s1, psi1[s], sc, s_tab, tc);
                    }
                    red_add(fl + g, tally);
                }
            }
        }
    }

    if (valid) {
#pragma unroll
        for (int v = 0; v < NV4; v++) {
            const int g = 4 * (lit + L * v);
            if (g < G) *reinterpret_cast<float4 *>(psi_row + g) = psi4[v];
        }
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int g = g_tail + lit + L * s;
            if (g < G) psi_row[g] = psi1[s];
        }
    }
}

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
    float *fine_source;     // [N][fai][G]
    float *fine_flux;       // [N][fai][G]
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
    const float *flux = p.fine_flux + (size_t)i * p.fai * p.G;
    const float *x = p.xs + (size_t)p.xs_index[i] * p.G * 3;
    const float vol = p.vol[i];
    const int G = p.G;
    auto per_fine = [&](long long jf) {
        auto per_group = [&](long long g) {
            return __fmul_rn(__fmul_rn(flux[jf * G + g], vol), x[3 * g]);
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
    const float *flux = p.fine_flux + (size_t)i * p.fai * p.G;
    const float *x = p.xs + (size_t)p.xs_index[i] * p.G * 3;
    const int G = p.G;
    for (int col = 0; col < 2; col++) {
        auto per_fine = [&](long long jf) {
            auto per_group = [&](long long g) { return __fmul_rn(x[3 * g + col], flux[jf * G + g]); };
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
    const long long per = (long long)p.fai * p.G;
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
    const float *flux = p.fine_flux + (size_t)row * G;
    float *q = p.fine_source + (size_t)row * G;
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

}  // namespace moc
