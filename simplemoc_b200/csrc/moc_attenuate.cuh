/* moc_attenuate.cuh -- K1, the multigroup attenuation + scalar-flux tally kernel
 * (reference src/solver.c:14-280 attenuate_fluxes, 1040-1138 attenuate_FSR_fluxes,
 * 1441-1464 interpolateTable).  Included by moc_kernels.cuh inside namespace moc.
 *
 * Blackwell specifics
 *   - the per-group arithmetic runs on PACKED FP32 pairs (FFMA2 / FMUL2 / FADD2, sm_100+):
 *     one issue slot per two energy groups, which turns the kernel from issue-bound into
 *     FMA-pipe/L2-bound (profiles/, DESIGN.md "K1");
 *   - every division is a multiplication by MUFU.RCP(sigT); the exponential is either the
 *     reference's table (shared memory, cell index bit-exact) or MUFU.EX2;
 *   - the quadratic fit of solver.c:74-76 depends on (source region, stencil, group) only, so it is
 *     done once per sweep (fit_coefficients_kernel); per segment the kernel gathers the three
 *     coefficient rows + sigT (4 x G floats) as 128-byte-per-8-lanes float4 loads that hit in the
 *     126 MB L2 (coefficients 31 MB + sigT 3.5 MB + scalar flux 17 MB on the default problem);
 *     tallies leave as 16-byte vector reductions (red.global.add.v4.f32);
 *   - a lane's odd group (G = 104: 96 + lane) runs on scalar FP32 instructions, which take one
 *     FMA-pipe cycle per warp instruction where a (half-empty) packed pair would take two.
 */
#pragma once
#include <type_traits>

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

// packed FP32x2 helpers (scalar operands broadcast for free: FFMA2 takes .F32 sources)
__device__ __forceinline__ float2 bc(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
// An FFMA2 whose three operands are all register PAIRS reads three even and three odd registers: alone it holds the
// sub-partition's issue port for 3.06 cycles where two scalar FFMA take 2.2, and 2.05 with at most two register pairs
// (tools/ubench/issue_slots.cu, profiles/r02_ubench_issue_slots.txt).  fma2r marks the seven such operations of a
// pair of groups; MOC_UNPACK_RRR = 1 spells them as two FFMA (same roundings).  Measured in the kernel that is 2 %
// SLOWER (356 vs 348 ms per launch on the default problem): mixed with the ALU-pipe instructions of the loop the
// packed form's extra cycle is hidden, the extra issue slots of the scalar form are not.  Kept at 0.
#ifndef MOC_UNPACK_RRR
#define MOC_UNPACK_RRR 0
#endif
__device__ __forceinline__ float2 fma2r(float2 a, float2 b, float2 c)
{
#if MOC_UNPACK_RRR
    return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y));
#else
    return __ffma2_rn(a, b, c);
#endif
}
__device__ __forceinline__ float fma2r(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, neg2(b)); }
// the same vocabulary on one float: the tail group of a lane (G = 104: 96 + lane) runs on scalar
// FFMA/FMUL/FADD -- one FMA-pipe cycle per warp instruction instead of the two a packed pair takes
__device__ __forceinline__ float neg2(float a) { return -a; }
__device__ __forceinline__ float fma2(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float mul2(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add2(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub2(float a, float b) { return __fadd_rn(a, -b); }
template <typename V> __device__ __forceinline__ V splat(float v);
template <> __device__ __forceinline__ float2 splat<float2>(float v) { return make_float2(v, v); }
template <> __device__ __forceinline__ float splat<float>(float v) { return v; }

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

// exhaustive check of the fast cell selection: every float bit pattern in [0, bits_max]
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
    int n;                 // cells; entry n of the shared table is (slope 0, intercept 1): every x > x_max
    const float *tab;      // shared memory, [n + 1] x (slope, intercept)
};

// per-segment scalars of attenuate_fluxes, hoisted out of the group loop
struct SegmentScalars {
    float ds;
    float a1, a2;      // zin, zin^2                 : q0 = c0 + a1 c1 + a2 c2   (solver.c:79)
    float b1, b2;      // mu, 2 mu zin               : q1 mu = b1 c1 + b2 c2     (solver.c:80)
    float b3;          // mu^2                       : q2 mu^2 = b3 c2           (solver.c:81)
    float weight;
};

// E = 1 - exp(-x), D = exp(-x) = 1 - E for a PAIR of groups.
//  MODE 0: the reference's linear table, cell chosen exactly as solver.c:1448 does (the slope sign
//          is the reference's), IEEE division.
//  MODE 1: the same with the verified fast division.
//  MODE 2: SFU: MUFU.EX2.
//  MODE 6: MODE 0 for the two-way sweep (moc_two_way.cuh), whose negative lengths select cells in front of the table:
//          answered from cell 0, as the oracle does where the reference reads its heap.
//  MODE 3 / 4: MODE 1 / 2 without the reference's x > maxVal -> 1 rule (solver.c:1444-1445): for segments so short
//  that no sigT of the slab can carry the optical length past maxVal (ds <= AttenuateParams::ds_noclamp, decided per
//  segment and warp) the rule cannot fire and two instructions per group go.  (Not per sweep: in the last 2D segment
//  of a track the reference measures the later 3D segments of a crossing ray from the ray's START height,
//  solver.c:514-523 -- lengths up to node height / cos(polar), 22 cm on the default problem, 0.08 % of all segments.)
template <int MODE>
__device__ __forceinline__ void one_minus_exp2(float2 x, const TableConsts &tc, float2 &E, float2 &D)
{
    if (MODE == 2 || MODE == 4) {
        const float2 arg = mul2(x, bc(-1.4426950408889634f));
        D.x = (MODE == 2 && x.x > tc.x_max) ? 0.0f : ex2_approx(arg.x);
        D.y = (MODE == 2 && x.y > tc.x_max) ? 0.0f : ex2_approx(arg.y);
        E = sub2(bc(1.0f), D);
    } else {
        float2 t;
        if (MODE == 0 || MODE == 6) {
            t.x = __fadd_rn(__fdiv_rn(x.x, tc.dx), tc.half_dx);
            t.y = __fadd_rn(__fdiv_rn(x.y, tc.dx), tc.half_dx);
        } else {
            float2 q = mul2(x, bc(tc.rdx));
            const float2 rem = fma2(neg2(q), bc(tc.dx), x);
            q = fma2(rem, bc(tc.rdx), q);
            t = add2(q, bc(tc.half_dx));
        }
        int c0 = __float2int_rz(t.x), c1 = __float2int_rz(t.y);
        if (MODE == 6) {
            c0 = max(c0, 0);
            c1 = max(c1, 0);
        }
        if (MODE != 3) {
            c0 = x.x > tc.x_max ? tc.n : c0;
            c1 = x.y > tc.x_max ? tc.n : c1;
        }
        // the halves of a packed operand come from different cells: one LDS.64 (slope, intercept) per group,
        // and the interpolation on two scalar FFMA (the same two FMA-pipe cycles as one FFMA2, without the
        // three register moves that re-pairing (slope0, slope1) / (intercept0, intercept1) costs)
        const float2 e0 = *reinterpret_cast<const float2 *>(tc.tab + 2 * c0);
        const float2 e1 = *reinterpret_cast<const float2 *>(tc.tab + 2 * c1);
        E.x = __fmaf_rn(e0.x, x.x, e0.y);
        E.y = __fmaf_rn(e1.x, x.y, e1.y);
        D = sub2(bc(1.0f), E);
    }
}

// one group on scalar instructions (same modes, same cell selection)
template <int MODE>
__device__ __forceinline__ void one_minus_exp2(float x, const TableConsts &tc, float &E, float &D)
{
    if (MODE == 2 || MODE == 4) {
        D = (MODE == 2 && x > tc.x_max) ? 0.0f : ex2_approx(__fmul_rn(x, -1.4426950408889634f));
        E = __fadd_rn(1.0f, -D);
    } else {
        float t;
        if (MODE == 0 || MODE == 6) {
            t = __fadd_rn(__fdiv_rn(x, tc.dx), tc.half_dx);
        } else {
            float q = __fmul_rn(x, tc.rdx);
            const float rem = __fmaf_rn(-q, tc.dx, x);
            q = __fmaf_rn(rem, tc.rdx, q);
            t = __fadd_rn(q, tc.half_dx);
        }
        int c = __float2int_rz(t);
        if (MODE == 6) c = max(c, 0);
        if (MODE != 3) c = x > tc.x_max ? tc.n : c;
        const float2 e = *reinterpret_cast<const float2 *>(tc.tab + 2 * c);
        E = __fmaf_rn(e.x, x, e.y);
        D = __fadd_rn(1.0f, -E);
    }
}

__device__ __forceinline__ float2 rcp_groups(float2 s) { return make_float2(rcp_approx(s.x), rcp_approx(s.y)); }
__device__ __forceinline__ float rcp_groups(float s) { return rcp_approx(s); }

// Energy groups of attenuate_fluxes (solver.c:66-82, 146-279): V = float2 is a PAIR of groups on packed
// FP32 instructions, V = float one group on scalar instructions.  Returns the tallies.
// (c0, d, e) are the fit coefficients (c0, c1, c2) of the segment's stencil, computed once per sweep by
// fit_coefficients_kernel exactly as solver.c:74-76 writes them, so that solver.c:79-81 is
//   q0 = c0 + a1 d + a2 e,  q1 mu = b1 d + b2 e,  q2 mu^2 = b3 e
// with the per-segment scalars a1 = zin, a2 = zin^2, b1 = mu, b2 = 2 mu zin, b3 = mu^2.
// The rest is the reference's formulas, regrouped so that every factor that does not depend on the
// group is a per-segment scalar and every division is a multiplication by MUFU.RCP(sigT):
//   in  = [ q0 tau + (sigT psi - q0) E + q2 mu^2 C3 / sigT^2 ] / sigT^2 + q1 mu R
//   out = q0 E / sigT + q1 mu (tau - E) / sigT^2 + q2 mu^2 R + psi (1 - E)
//   R   = tau (tau - 2) + 2 E / sigT^3          (solver.c:175-176 as parenthesised there)
//   C3  = [tau (tau (tau - 3) + 6) - 6 E] / 3
// COEF = false (source slabs larger than the L2, where three more rows per region would only add DRAM
// traffic): (c0, d, e) arrive as the three source rows (y1, y2, y3) of the stencil and the fit is done here,
// with the scalars a1 = zin/(2dz), a2 = zin^2/(2dz^2), b1 = mu/(2dz), b2 = 2 mu zin/(2dz^2), b3 = mu^2/(2dz^2).
template <int MODE, bool COEF, typename V>
__device__ __forceinline__ V attenuate_groups(V c0, V d, V e, V sigT, V &psi, const SegmentScalars &k,
                                              const TableConsts &tc)
{
    if (!COEF) {
        const V y1 = c0, y2 = d, y3 = e;
        c0 = y2;
        d = sub2(y1, y3);
        e = fma2(splat<V>(-2.f), y2, add2(y1, y3));
    }
    const V q0 = fma2(splat<V>(k.a2), e, fma2(splat<V>(k.a1), d, c0));
    const V q1m = fma2(splat<V>(k.b2), e, mul2(splat<V>(k.b1), d));   // q1 * mu
    const V q2m = mul2(splat<V>(k.b3), e);                            // q2 * mu^2
    const V tau = mul2(sigT, splat<V>(k.ds));
    V E, D;
    one_minus_exp2<MODE>(tau, tc, E, D);
    const V r1 = rcp_groups(sigT);
    const V r2 = mul2(r1, r1);
    const V Er1 = mul2(E, r1);
    const V reuse = fma2(splat<V>(2.f), mul2(Er1, r2), mul2(tau, add2(tau, splat<V>(-2.f))));
    const V A = fma2r(q0, tau, mul2(fma2r(sigT, psi, neg2(q0)), E));
    const V c3 = fma2(splat<V>(-2.f), E,
                      mul2(tau, fma2(tau, fma2(tau, splat<V>(1.f / 3.f), splat<V>(-1.f)), splat<V>(2.f))));
    const V X = fma2r(mul2(q2m, r2), c3, A);
    const V in = fma2r(X, r2, mul2(q1m, reuse));
    // psi (1 - E) with D = 1 - E formed first, as solver.c:264 does: folding it into (psi + out) - psi E
    // saves an instruction but moves the cancellation (measured: fewer elements within 1e-4 of the reference)
    V out = mul2(q0, Er1);
    out = fma2r(mul2(q1m, r2), sub2(tau, E), out);
    out = fma2r(q2m, reuse, out);
    psi = fma2r(psi, D, out);
    return mul2(splat<V>(k.weight), in);
}

// one energy group of attenuate_FSR_fluxes (solver.c:1104-1115); scalar: 11 FLOP, not worth packing
template <int MODE>
__device__ __forceinline__ float attenuate_flat(float src, float sigT, float &psi, const SegmentScalars &k,
                                                const TableConsts &tc)
{
    const float tau = sigT * k.ds;
    float E, D;
    one_minus_exp2<MODE>(tau, tc, E, D);
    const float q = __fdiv_rn(src, sigT);   // flat source: the difference psi - q cancels, keep the division exact
    const float dpsi = (psi - q) * E;
    psi -= dpsi;
    return k.weight * dpsi;
}

// The quadratic "fitting" of solver.c:74-76 for every stencil of every source region, written exactly as the
// reference writes it (IEEE divisions, same order of additions): stencil r0 of region i fits source rows
// r0, r0+1, r0+2 and is what fine intervals r0+1 (and, at the edges, 0 and fai-1: solver.c:55-112) use.
// coef[i][r0][0..2][pitch] = (c0, c1, c2).  N * (fai-2) * pitch threads of work, once per sweep.
__global__ void fit_coefficients_kernel(const float *__restrict__ fine_source, float *__restrict__ coef,
                                        long long n_regions, int fai, int pitch, float dz)
{
    const int S = fai - 2;
    const long long cells = n_regions * S * pitch;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= cells) return;
    const int g = (int)(e % pitch);
    const long long rs = e / pitch;
    const int r0 = (int)(rs % S);
    const long long i = rs / S;
    const float *y = fine_source + ((size_t)i * fai + r0) * pitch + g;
    const float y1 = y[0], y2 = y[pitch], y3 = y[2 * pitch];
    float *c = coef + (size_t)rs * 3 * pitch + g;
    const float two_dz = __fmul_rn(2.f, dz);
    c[0] = y2;
    c[pitch] = __fdiv_rn(__fadd_rn(y1, -y3), two_dz);
    c[2 * pitch] = __fdiv_rn(__fadd_rn(__fadd_rn(y1, -__fmul_rn(2.f, y2)), y3), __fmul_rn(two_dz, dz));
}

// The gathered rows are read-only for the whole launch: ld.global.nc.  They practically never hit in L1
// (0.8 %), but loading them with L1::no_allocate is 50 % SLOWER (measured: 521 vs 345 ms per launch).
__device__ __forceinline__ float4 gather4(const float4 *p) { return __ldg(p); }
__device__ __forceinline__ float gather1(const float *p) { return __ldg(p); }

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
// GC: the number of groups as a compile-time constant (row strides become immediates), 0 = a.G.
template <int L, int NV4, int NS, int MODE, bool FLAT, int GC, bool COEF>
__global__ void __launch_bounds__(128, MOC_ATT_MIN_BLOCKS) attenuate_kernel(const AttenuateParams a)
{
    extern __shared__ float s_tab[];   // [n+1] x (slope, intercept)
    TableConsts tc;
    tc.dx = a.table_dx; tc.rdx = a.table_rdx; tc.half_dx = a.table_half_dx; tc.x_max = a.table_max;
    tc.n = a.table_n;
    tc.tab = s_tab;
    if (MODE != 2) {
        for (int e = threadIdx.x; e < 2 * a.table_n; e += blockDim.x) s_tab[e] = a.table[e];
        if (threadIdx.x == 0) {
            s_tab[2 * a.table_n] = 0.f;
            s_tab[2 * a.table_n + 1] = 1.f;
        }
        __syncthreads();
    }
    constexpr int TPW = 32 / L;
    const int lane = threadIdx.x & 31;
    const int lit = lane % L;
    const int G = GC ? GC : a.G;
    const int W = GC ? (GC + 31) / 32 * 32 : a.pitch;   // row pitch of the source slab: 128-byte aligned rows
    const long long t = a.first_track + ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * TPW + lane / L;
    const bool valid = t < a.end_track;

    uint32_t n_rec = 0, at = 0;
    SegmentScalars sc;
    sc.ds = 0.f; sc.a1 = sc.a2 = sc.b1 = sc.b2 = sc.b3 = 0.f; sc.weight = 0.f;
    float mu = 0.f;
    if (valid) {
        n_rec = a.seg_count[t];
        const long long pair = t / a.Z;
        at = (uint32_t)(a.rec_base[pair] - a.batch_first_record) + (uint32_t)(t - pair * a.Z);
        const int j = (int)(pair % a.P);
        const long long i = pair / a.P;
        mu = a.mu[j];
        float w0 = __fmul_rn(a.p_weight[t], a.az_weight[i]);   // solver.c:49
        if (FLAT) w0 = __fmul_rn(w0, mu);                      // solver.c:1064
        sc.weight = w0;
        sc.b1 = COEF ? mu : mu * a.inv_2dz;
        sc.b3 = COEF ? mu * mu : mu * mu * a.inv_2dz2;
    }
    const float two_mu = COEF ? 2.f * mu : 2.f * mu * a.inv_2dz2;
    unsigned int longest = n_rec;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned int o = __shfl_xor_sync(0xffffffffu, longest, d);
        longest = o > longest ? o : longest;
    }

    // group ownership: quads at element offset 4*lit + 4*L*v, singles at g_tail + lit + L*s
    constexpr int g_tail = 4 * L * NV4;
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

    // Segment records: the L lanes of a track fetch its next L records (lane `lit` holds record
    // block*L + lit) one block ahead of use; each segment's record is then broadcast from its holder
    // with a width-L shuffle.  Records are segment-major inside a z-stack (record j of ray k at
    // base + j*Zs + k), so the tracks of a warp/CTA -- neighbouring rays -- share 32-byte sectors.
    const float *rec_ds = a.rec_ds + at;
    const float *rec_zin = a.rec_zin + at;
    const uint32_t *rec_code = a.rec_code + at;
    float cur_ds = 0.f, cur_zin = 0.f, nxt_ds = 0.f, nxt_zin = 0.f;
    uint32_t cur_code = 0, nxt_code = 0;
    const uint32_t Zs = (uint32_t)a.Zs;   // the j-th record of a track is Zs words after its (j-1)-th
    if ((uint32_t)lit < n_rec) {
        nxt_ds = __ldg(rec_ds + lit * Zs);
        nxt_zin = __ldg(rec_zin + lit * Zs);
        nxt_code = __ldg(rec_code + lit * Zs);
    }
    const float *const src_base = COEF ? a.coef : a.fine_source;   // fit coefficients, or the source rows themselves
    const float *const src_q = src_base + 4 * lit;               // this lane's quads
    const float *const sig_q = a.sigT + 4 * lit;
    float *const flx_q = a.fine_flux + 4 * lit;
    const float *const src_s = src_base + g_tail + lit;          // this lane's single groups
    const float *const sig_s = a.sigT + g_tail + lit;
    float *const flx_s = a.fine_flux + g_tail + lit;

    for (unsigned int sgm = 0; sgm < longest; sgm++) {
        const int slot = sgm % L;
        if (slot == 0) {
            cur_ds = nxt_ds; cur_zin = nxt_zin; cur_code = nxt_code;
            const uint32_t ahead = sgm + L + lit;
            if (ahead < n_rec) {
                nxt_ds = __ldg(rec_ds + ahead * Zs);
                nxt_zin = __ldg(rec_zin + ahead * Zs);
                nxt_code = __ldg(rec_code + ahead * Zs);
            }
        }
        const float seg_ds = __shfl_sync(0xffffffffu, cur_ds, slot, L);
        const float zin = __shfl_sync(0xffffffffu, cur_zin, slot, L);
        const uint32_t code = __shfl_sync(0xffffffffu, cur_code, slot, L);
        const bool need_clamp = __any_sync(0xffffffffu, sgm < n_rec && !(seg_ds <= a.ds_noclamp));
        if (sgm < n_rec) {
            sc.ds = seg_ds;
            sc.a1 = COEF ? zin : zin * a.inv_2dz;
            sc.a2 = COEF ? zin * zin : zin * zin * a.inv_2dz2;
            sc.b2 = two_mu * zin;
            const uint32_t qsr = code & 0xffffffu;
            const uint32_t r0 = (code >> 24) & 63u;
            const uint32_t which = code >> 30;
            // element offsets inside the source slab (< 2^32, checked by moc_create)
            // COEF: o_src addresses the (c0, c1, c2) rows of stencil r0 in the coefficient slab; otherwise the
            // first source row of the stencil (flat source: r0 = the fine interval itself)
            const uint32_t o_row = (qsr * a.fai + r0) * (uint32_t)W;
            const uint32_t o_src = COEF ? (qsr * a.coef_stencils + r0) * 3u * (uint32_t)W : o_row;
            const uint32_t o_sig = qsr * (uint32_t)W;
            const uint32_t o_flx = o_row + which * (uint32_t)W;
            auto groups = [&](auto mode) {
                constexpr int M = decltype(mode)::value;
    #pragma unroll
                for (int v = 0; v < NV4; v++) {
                    const int g = 4 * (lit + L * v);
                    if (g < G) {
                        const float4 s4 = gather4(reinterpret_cast<const float4 *>(sig_q + o_sig + 4 * L * v));
                        float4 tally;
                        if (FLAT) {
                            const float4 y = gather4(reinterpret_cast<const float4 *>(src_q + o_src + 4 * L * v));
                            tally.x = attenuate_flat<M>(y.x, s4.x, psi4[v].x, sc, tc);
                            tally.y = attenuate_flat<M>(y.y, s4.y, psi4[v].y, sc, tc);
                            tally.z = attenuate_flat<M>(y.z, s4.z, psi4[v].z, sc, tc);
                            tally.w = attenuate_flat<M>(y.w, s4.w, psi4[v].w, sc, tc);
                        } else {
                            const float4 k0 = gather4(reinterpret_cast<const float4 *>(src_q + o_src + 4 * L * v));
                            const float4 k1 = gather4(reinterpret_cast<const float4 *>(src_q + o_src + W + 4 * L * v));
                            const float4 k2 = gather4(reinterpret_cast<const float4 *>(src_q + o_src + 2 * W + 4 * L * v));
                            float2 plo = make_float2(psi4[v].x, psi4[v].y), phi = make_float2(psi4[v].z, psi4[v].w);
                            const float2 tlo = attenuate_groups<M, COEF>(
                                make_float2(k0.x, k0.y), make_float2(k1.x, k1.y), make_float2(k2.x, k2.y),
                                make_float2(s4.x, s4.y), plo, sc, tc);
                            const float2 thi = attenuate_groups<M, COEF>(
                                make_float2(k0.z, k0.w), make_float2(k1.z, k1.w), make_float2(k2.z, k2.w),
                                make_float2(s4.z, s4.w), phi, sc, tc);
                            psi4[v] = make_float4(plo.x, plo.y, phi.x, phi.y);
                            tally = make_float4(tlo.x, tlo.y, thi.x, thi.y);
                        }
                        red_add_v4(flx_q + o_flx + 4 * L * v, tally);
                    }
                }
    #pragma unroll
                for (int s = 0; s < NS; s++) {
                    const int g = g_tail + lit + L * s;
                    if (g < G) {
                        const float s1 = gather1(sig_s + o_sig + L * s);
                        float tally;
                        if (FLAT) {
                            tally = attenuate_flat<M>(gather1(src_s + o_src + L * s), s1, psi1[s], sc, tc);
                        } else {
                            // the tail group runs on scalar FFMA/FMUL/FADD, not on a half-empty pair
                            tally = attenuate_groups<M, COEF>(gather1(src_s + o_src + L * s), gather1(src_s + o_src + W + L * s),
                                                           gather1(src_s + o_src + 2 * W + L * s), s1, psi1[s], sc, tc);
                        }
                        red_add(flx_s + o_flx + L * s, tally);
                    }
                }
            };
            // the x > maxVal test of the table only where a segment is long enough to need it (warp-uniform)
            if constexpr (MODE == 0) {
                groups(std::integral_constant<int, 0>());
            } else {
                if (need_clamp) groups(std::integral_constant<int, MODE>());
                else groups(std::integral_constant<int, MODE + 2>());
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


#ifndef MOC_ABLATE
#define MOC_ABLATE 0   // experiments only: 1 = no operand copies, 2 = no reductions, 3 = neither (profiles/r02_K1_findings.md)
#endif

// ------------------------------------------------------------------ K1 with the gathers staged by the TMA unit
//
// The same attenuation, lane mapping and arithmetic as attenuate_kernel (8 lanes per track, 4 tracks per warp, the
// track's angular flux in registers); what changes is how the gathered rows reach the arithmetic.  There every lane
// issues its own LDG.128 for (c0, c1, c2, sigT) of the segment at the top of the iteration and needs them at once:
// ncu (profiles/r02_K1_baseline_ncu.txt) shows the FMA pipe 64 % busy and the L1TEX 66 % busy with 1.8 long-
// scoreboard stalls per issue -- neither unit saturated, the two phases of a warp simply do not overlap, and 96
// registers per thread leave only 5 warps per scheduler to cover for each other.  Here
//   - fit_coefficients4_kernel packs (c0, c1, c2, sigT) of every (region, stencil) into ONE contiguous block of
//     16 G bytes (G = 104: 1664 B = 13 full 128-byte lines, no padding): one segment of one track needs one block;
//   - every warp owns a two-stage ring in shared memory; the leader lane of each track issues ONE bulk copy
//     (cp.async.bulk global -> shared, SASS UBLKCP, completion counted on an mbarrier) for the segment two steps
//     ahead as soon as the stage is free, so a block has a whole segment's arithmetic (~2000 cycles) to arrive;
//   - the arithmetic reads its operands with LDS.128 at immediate offsets: no address arithmetic, no long
//     scoreboard, fewer live registers.
// L2 -> SM traffic per integration is unchanged (16 B + the 4-byte reduction); the copies bypass the LSU's tag
// stage and miss handling.
struct StagedParams {
    AttenuateParams a;
    const float *coef4;      // [N][stencils][4][G]  (c0 | c1 | c2 | sigT), rows of exactly G floats
};

// coef4[i][r0][0..2][G] = the fit of source rows r0 .. r0+2 as solver.c:74-76 writes it, [3][G] = sigT of region i
__global__ void fit_coefficients4_kernel(const float *__restrict__ fine_source, const float *__restrict__ sigT,
                                         float *__restrict__ coef4, long long n_regions, int fai, int pitch, int G, float dz)
{
    const int S = fai - 2;
    const long long cells = n_regions * S * G;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= cells) return;
    const int g = (int)(e % G);
    const long long rs = e / G;
    const int r0 = (int)(rs % S);
    const long long i = rs / S;
    const float *y = fine_source + ((size_t)i * fai + r0) * pitch + g;
    const float y1 = y[0], y2 = y[pitch], y3 = y[2 * pitch];
    float *c = coef4 + (size_t)rs * 4 * G + g;
    const float two_dz = __fmul_rn(2.f, dz);
    c[0] = y2;
    c[G] = __fdiv_rn(__fadd_rn(y1, -y3), two_dz);
    c[2 * G] = __fdiv_rn(__fadd_rn(__fadd_rn(y1, -__fmul_rn(2.f, y2)), y3), __fmul_rn(two_dz, dz));
    c[3 * G] = sigT[(size_t)i * pitch + g];
}

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
// global -> shared bulk copy by the TMA unit; `bytes` (a multiple of 16) are counted on the mbarrier when they land
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// shared-memory plan of attenuate_staged_kernel for G groups (host and device agree through these)
__host__ __device__ constexpr int staged_track_stride(int G, int L = 8)
{
    // 16 G bytes per block, padded against bank conflicts between the tracks of a warp.
    // L = 8 (four tracks per warp): a quarter-warp's LDS.128 reads one track's 128 contiguous bytes; the tail groups
    //   (8 lanes x 4 bytes per track) want consecutive tracks 32 bytes further along the banks: stride = 32 mod 128.
    // L = 4 (eight tracks per warp): a quarter-warp's LDS.128 reads 64 bytes of each of two tracks: stride = 64 mod 128.
    const int want = L == 8 ? 32 : 64;
    return 16 * G + ((want - (16 * G) % 128) + 128) % 128;
}
__host__ __device__ constexpr int staged_table_bytes(int table_n) { return (8 * (table_n + 1) + 127) / 128 * 128; }
__host__ __device__ constexpr int staged_smem_bytes(int G, int table_n, int warps, bool table, int L = 8)
{
    return (table ? staged_table_bytes(table_n) : 0) + 128 /* mbarriers */ + warps * 2 * (32 / L) * staged_track_stride(G, L);
}

// L lanes per track (8: four tracks per warp, four CTAs per SM; 4: eight tracks per warp -- the per-segment work that
// does not depend on the number of tracks is shared by twice as many -- two CTAs per SM with twice the registers).
template <int L, int NV4, int NS, int MODE, int GC>
__global__ void __launch_bounds__(128, L == 8 ? 4 : 2) attenuate_staged_kernel(const StagedParams sp)
{
    const AttenuateParams &a = sp.a;
    constexpr int TPW = 32 / L, G = GC;
    constexpr int STRIDE = staged_track_stride(G, L), STAGE = TPW * STRIDE, BLOCK_BYTES = 16 * G;
    extern __shared__ __align__(128) unsigned char s_raw[];
    constexpr bool TABLE = MODE != 2;
    const int table_bytes = TABLE ? staged_table_bytes(a.table_n) : 0;
    float *s_tab = reinterpret_cast<float *>(s_raw);
    TableConsts tc;
    tc.dx = a.table_dx; tc.rdx = a.table_rdx; tc.half_dx = a.table_half_dx; tc.x_max = a.table_max;
    tc.n = a.table_n;
    tc.tab = s_tab;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lit = lane % L, grp = lane / L;
    const uint32_t bars = smem_addr(s_raw + table_bytes) + 16u * warp;             // two mbarriers per warp
    unsigned char *const ring = s_raw + table_bytes + 128 + (size_t)warp * 2 * STAGE;
    if (TABLE) {
        for (int e = threadIdx.x; e < 2 * a.table_n; e += blockDim.x) s_tab[e] = a.table[e];
        if (threadIdx.x == 0) {
            s_tab[2 * a.table_n] = 0.f;
            s_tab[2 * a.table_n + 1] = 1.f;
        }
    }
#if MOC_ABLATE & 1
    for (int e = lane; e < 2 * STAGE / 4; e += 32) reinterpret_cast<float *>(ring)[e] = 0.5f;
#endif
    if (lane == 0) {
        mbar_init(bars, TPW);          // one arrival per track leader and use of the stage
        mbar_init(bars + 8, TPW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long t = a.first_track + ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * TPW + grp;
    const bool valid = t < a.end_track;
    uint32_t n_rec = 0, at = 0;
    SegmentScalars sc;
    sc.ds = 0.f; sc.a1 = sc.a2 = sc.b1 = sc.b2 = sc.b3 = 0.f; sc.weight = 0.f;
    float mu = 0.f;
    if (valid) {
        n_rec = a.seg_count[t];
        const long long pair = t / a.Z;
        at = (uint32_t)(a.rec_base[pair] - a.batch_first_record) + (uint32_t)(t - pair * a.Z);
        const int j = (int)(pair % a.P);
        const long long i = pair / a.P;
        mu = a.mu[j];
        sc.weight = __fmul_rn(a.p_weight[t], a.az_weight[i]);   // solver.c:49
        sc.b1 = mu;
        sc.b3 = mu * mu;
    }
    const float two_mu = 2.f * mu;
    unsigned int longest = n_rec;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned int o = __shfl_xor_sync(0xffffffffu, longest, d);
        longest = o > longest ? o : longest;
    }

    constexpr int g_tail = 4 * L * NV4;
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

    // segment records, eight at a time, one block ahead (as in attenuate_kernel)
    const float *rec_ds = a.rec_ds + at;
    const float *rec_zin = a.rec_zin + at;
    const uint32_t *rec_code = a.rec_code + at;
    float cur_ds = 0.f, cur_zin = 0.f, nxt_ds = 0.f, nxt_zin = 0.f;
    uint32_t cur_code = 0, nxt_code = 0;
    const uint32_t Zs = (uint32_t)a.Zs;
    if ((uint32_t)lit < n_rec) {
        nxt_ds = __ldg(rec_ds + lit * Zs);
        nxt_zin = __ldg(rec_zin + lit * Zs);
        nxt_code = __ldg(rec_code + lit * Zs);
    }
    const uint32_t stencils = (uint32_t)a.coef_stencils;
    // the leader lane of a track asks the TMA unit for the coefficient block of one of its segments
    auto request = [&](uint32_t code, bool active, int stage) {
#if MOC_ABLATE & 1      // experiment: no copies (operands stay whatever the ring was filled with)
        return;
#endif
        if (lit == 0) {
            const uint32_t bar = bars + 8u * stage;
            if (active) {
                const uint32_t qsr = code & 0xffffffu, r0 = (code >> 24) & 63u;
                const float *src = sp.coef4 + (size_t)(qsr * stencils + r0) * (size_t)(4 * G);
                mbar_arrive_expect_tx(bar, BLOCK_BYTES);
                bulk_copy_g2s(smem_addr(ring + stage * STAGE + grp * STRIDE), src, BLOCK_BYTES, bar);
            } else {
                mbar_arrive(bar);
            }
        }
    };
    // segments 0 and 1 (block 0 of the records still sits in nxt_*)
    {
        const uint32_t c0 = __shfl_sync(0xffffffffu, nxt_code, 0, L), c1 = __shfl_sync(0xffffffffu, nxt_code, 1, L);
        if (longest > 0) request(c0, 0 < n_rec, 0);
        if (longest > 1) request(c1, 1 < n_rec, 1);
    }
    float *const flx_q = a.fine_flux + 4 * lit;
    float *const flx_s = a.fine_flux + g_tail + lit;

    for (unsigned int sgm = 0; sgm < longest; sgm++) {
        const int slot = sgm % L, stage = sgm & 1;
        if (slot == 0) {
            cur_ds = nxt_ds; cur_zin = nxt_zin; cur_code = nxt_code;
            const uint32_t ahead = sgm + L + lit;
            if (ahead < n_rec) {
                nxt_ds = __ldg(rec_ds + ahead * Zs);
                nxt_zin = __ldg(rec_zin + ahead * Zs);
                nxt_code = __ldg(rec_code + ahead * Zs);
            }
        }
        const float seg_ds = __shfl_sync(0xffffffffu, cur_ds, slot, L);
        const float zin = __shfl_sync(0xffffffffu, cur_zin, slot, L);
        const uint32_t code = __shfl_sync(0xffffffffu, cur_code, slot, L);
        // the record of the segment two steps ahead: in this block of eight, or already in the next one
        const uint32_t code2 = __shfl_sync(0xffffffffu, slot < L - 2 ? cur_code : nxt_code, (slot + 2) % L, L);
        const bool need_clamp = __any_sync(0xffffffffu, sgm < n_rec && !(seg_ds <= a.ds_noclamp));
#if !(MOC_ABLATE & 1)
        while (!mbar_try_wait(bars + 8u * stage, (sgm >> 1) & 1u)) { }
#endif
        if (sgm < n_rec) {
            sc.ds = seg_ds;
            sc.a1 = zin;
            sc.a2 = zin * zin;
            sc.b2 = two_mu * zin;
            const uint32_t qsr = code & 0xffffffu;
            const uint32_t r0 = (code >> 24) & 63u;
            const uint32_t which = code >> 30;
            const uint32_t o_flx = (qsr * a.fai + r0 + which) * (uint32_t)a.pitch;
            const unsigned char *blk = ring + stage * STAGE + grp * STRIDE;
            const float *rows = reinterpret_cast<const float *>(blk);
            auto groups = [&](auto mode) {
                constexpr int M = decltype(mode)::value;
    #pragma unroll
                for (int v = 0; v < NV4; v++) {
                    const int g = 4 * (lit + L * v);
                    if (g < G) {
                        const float4 k0 = *reinterpret_cast<const float4 *>(rows + g);
                        const float4 k1 = *reinterpret_cast<const float4 *>(rows + G + g);
                        const float4 k2 = *reinterpret_cast<const float4 *>(rows + 2 * G + g);
                        const float4 s4 = *reinterpret_cast<const float4 *>(rows + 3 * G + g);
                        float2 plo = make_float2(psi4[v].x, psi4[v].y), phi = make_float2(psi4[v].z, psi4[v].w);
                        const float2 tlo = attenuate_groups<M, true>(
                            make_float2(k0.x, k0.y), make_float2(k1.x, k1.y), make_float2(k2.x, k2.y),
                            make_float2(s4.x, s4.y), plo, sc, tc);
                        const float2 thi = attenuate_groups<M, true>(
                            make_float2(k0.z, k0.w), make_float2(k1.z, k1.w), make_float2(k2.z, k2.w),
                            make_float2(s4.z, s4.w), phi, sc, tc);
                        psi4[v] = make_float4(plo.x, plo.y, phi.x, phi.y);
    #if MOC_ABLATE & 2      // experiment: no reductions (the tallies are folded into psi so the arithmetic stays live)
                        psi4[v].x += 1e-30f * (tlo.x + tlo.y + thi.x + thi.y);
    #else
                        red_add_v4(flx_q + o_flx + 4 * L * v, make_float4(tlo.x, tlo.y, thi.x, thi.y));
    #endif
                    }
                }
    #pragma unroll
                for (int s = 0; s < NS; s++) {
                    const int g = g_tail + lit + L * s;
                    if (g < G) {
                        const float tally = attenuate_groups<M, true>(rows[g], rows[G + g], rows[2 * G + g], rows[3 * G + g],
                                                                         psi1[s], sc, tc);
    #if MOC_ABLATE & 2
                        psi1[s] += 1e-30f * tally;
    #else
                        red_add(flx_s + o_flx + L * s, tally);
    #endif
                    }
                }
            };
            if constexpr (MODE == 0) {
                groups(std::integral_constant<int, 0>());
            } else {
                if (need_clamp) groups(std::integral_constant<int, MODE>());
                else groups(std::integral_constant<int, MODE + 2>());
            }
        }
        // Every lane has consumed its operands of this stage (the reductions above depend on them): hand the stage
        // to the copy of segment sgm + 2.  Write-after-read across the generic and the async proxy needs no proxy
        // fence (that is for generic WRITES the async proxy reads); the warp barrier orders the lanes' reads before
        // the leaders' requests, as the consumer-release / producer-acquire pair of a TMA pipeline does.
        __syncwarp();
        if (sgm + 2 < longest) request(code2, sgm + 2 < n_rec, stage);
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
