/* moc_walk_warp.cuh -- K0, the axial ray trace: one WARP per (2D track, polar angle) z-stack for
 * Z <= 32 * KPT (KPT <= 4: stack_walk_warp_kernel), ceil(Z / 128) warps of one CTA per stack for
 * 128 < Z <= 2048 (stack_walk_block_kernel); lane l of warp q owns the KPT consecutive rays
 * k = (32 q + l) * KPT .. +KPT-1.
 * Included by moc_kernels.cuh inside namespace moc.  Same contract as stack_walk_kernel
 * (reference src/solver.c:347-529; window semantics SURVEY A.3), different mapping:
 *
 *   - the per-step prefix over the stack (exits below a ray, segments below a ray) is ONE 32-bit
 *     warp shuffle scan of a packed (exits | segments) word, the step totals are two REDUX
 *     instructions; the one-warp form needs no __syncthreads (2 KB of shared memory per warp carry
 *     the crossing rays and the step's source regions between lanes, see below), the several-warp
 *     form passes the warps' totals through 66 words of shared memory;
 *   - one launch per ray direction (upward / downward), the direction is a template parameter:
 *     no per-ray direction branches and only one instantiation resident in the instruction caches;
 *   - rays that stay inside their fine axial interval for the whole 2D segment are handled
 *     branch-free, unrolled over the lane's rays; the interval-crossing walk exists ONCE, in a loop
 *     the lanes enter only for the rays that need it -- in the one-warp form the crossing rays of a
 *     step (a quarter of the stack) are first numbered across the warp and dealt out one per lane
 *     (COMPACT), and the hash + remainder of the step's consecutive rand() draws is spread evenly
 *     over the lanes the same way (staged): both loops ran with 10-13 lanes of 32 busy before;
 *   - fine intervals are computed without the IEEE-division sequence and without conversion
 *     instructions (XU pipe): a Newton quotient on FMA units and a directed-rounding add, used
 *     only after interval_check_kernel verified it against the division for every float of the
 *     domain's height range (moc_create); the double division by cos(polar) is a reciprocal
 *     product with one FMA correction (3 DP operations) that falls back to the IEEE division
 *     whenever its result lies within 4 ulp of a float rounding boundary (div_by_cos);
 *   - s_full = length / sin(polar) (a double division, uniform over the stack) is evaluated by
 *     lane l for step 32*b + l and broadcast, instead of by every lane for every step;
 *   - records are segment-major inside a stack (slot(ray k, segment j) = base + j * Zs + k): the
 *     rays of a stack write neighbouring words at every step instead of one word per 500 bytes;
 *   - the emitting pass walks ONCE: segment length, axial offset and stencil rows are written
 *     while walking (the slot of a segment inside its track is known), the source region -- which
 *     depends on the segment's serial index, i.e. on the scan -- is added afterwards from a byte
 *     per segment kept in registers;
 *   - x % n_regions and x % fai are multiplications by precomputed reciprocals (checked
 *     exhaustively at moc_create, otherwise the kernel keeps the hardware remainder).
 */
#pragma once

// Both on by default; 0 restores the per-lane forms (kept for A/B timing).
#ifndef MOC_WALK_COMPACT_CROSSING
#define MOC_WALK_COMPACT_CROSSING 1
#endif
#ifndef MOC_WALK_STAGED_HASH
#define MOC_WALK_STAGED_HASH 1
#endif
// per-warp shared-memory scratch of the walk: 128 crossing rays x 16 bytes, or 512 source regions
#define WALK_SCRATCH_BYTES 2048

// x % m.n for x < 2^31:  q = (x * magic) >> (32 + shift)   (Granlund-Montgomery, n = 31 bits)
__device__ __forceinline__ uint32_t fastmod31(uint32_t x, uint32_t n, uint32_t magic, uint32_t shift)
{
    const uint32_t q = __umulhi(x, magic) >> shift;
    return x - q * n;
}

// exhaustive check of fastmod31 against the hardware remainder: every x in [0, 2^31)
__global__ void fastmod_check_kernel(uint32_t n, uint32_t magic, uint32_t shift, unsigned long long *mismatches)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned int bad = 0;
    for (unsigned long long x = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; x < (1ull << 31); x += stride)
        bad += fastmod31((uint32_t)x, n, magic, shift) != (uint32_t)x % n;
    if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}

__device__ __forceinline__ uint32_t warp_inclusive_scan_u32(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, v, d);
        v += lane >= d ? up : 0u;
    }
    return v;
}

__device__ __forceinline__ unsigned long long warp_inclusive_scan_u64(unsigned long long v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, v, d);
        v += lane >= d ? up : 0ull;
    }
    return v;
}

// The fine axial interval of a height (solver.c:895-907: (int)(z/dz) for upward rays,
// (int)ceil(z/dz) for downward rays) without the division instruction sequence and without
// conversion instructions: quotient by one Newton step on z * fl(1/dz), rounding by adding
// 1.5 * 2^23 in round-toward-zero / round-up mode.  Used only where interval_check_kernel has
// verified, for EVERY float in [lo, hi], that it returns the same integer as the IEEE division.
template <bool UP>
__device__ __forceinline__ int interval_by_fma(float z, float dz, float rdz)
{
    float q = __fmul_rn(z, rdz);
    const float rem = __fmaf_rn(-q, dz, z);
    q = __fmaf_rn(rem, rdz, q);
    const float m = UP ? __fadd_rz(q, 12582912.0f) : __fadd_ru(q, 12582912.0f);
    return __float_as_int(m) - 0x4B400000;
}

// mode 0: (int)(z/dz), z >= 0;   mode 1: (int)ceilf(z/dz), any sign
__global__ void interval_check_kernel(unsigned int bits_lo, unsigned int bits_hi, float dz, float rdz, int mode,
                                      unsigned long long *mismatches)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned int bad = 0;
    for (unsigned long long b = (unsigned long long)bits_lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
         b <= bits_hi; b += stride) {
        const float z = __uint_as_float((unsigned int)b);
        if (mode == 0) bad += interval_by_fma<true>(z, dz, rdz) != axial_interval<true>(z, dz);
        else bad += interval_by_fma<false>(z, dz, rdz) != axial_interval<false>(z, dz);
    }
    if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}

// FAST: the handle verified the FMA intervals AND that no trial height can leave the verified
// range (ray heights inside the node -- checked by the kernel prologue -- and
// max(s_full * |cos|) <= node height -- checked at moc_create)
template <bool UP, bool FAST>
__device__ __forceinline__ int interval_of(const WalkParams &w, float z)
{
    return FAST ? interval_by_fma<UP>(z, w.dz_interval, w.iv_rdz) : axial_interval<UP>(z, w.dz_interval);
}

// (float)(num / cos_p) for a float-valued num, as solver.c:449 computes it: an IEEE double division,
// narrowed to float.  !FAST: exactly that instruction sequence.  FAST: q' = q + (num - cos_p q) rcos with
// q = num rcos and rcos = RN(1 / cos_p) -- three DP operations instead of the ~40 of the division.  q' is
// within 2 ulp(double) of the true quotient (q is within 2 ulp: one rounding of rcos, one of the product; the
// correction step with the exact FMA residual does not make it worse), and so is the IEEE quotient, hence
// the two differ by fewer than 4 ulp and narrow to the SAME float unless q' lies within 4 ulp of a
// float rounding boundary (the midpoint between two floats: low 29 mantissa bits = 0x10000000).  That case
// (9 / 2^29 of all quotients) takes the IEEE division; no theorem about q' being correctly rounded is needed.
template <bool FAST>
__device__ __forceinline__ float div_by_cos(float num, double cos_p, double rcos)
{
    const double a = (double)num;
    if (!FAST) return (float)__ddiv_rn(a, cos_p);
    const double q = __dmul_rn(a, rcos);
    const double r = __fma_rn(-cos_p, q, a);
    const double q2 = __fma_rn(r, rcos, q);
    const uint32_t lo29 = (uint32_t)__double2loint(q2) & 0x1fffffffu;
    if (lo29 - 0x0ffffffcu <= 8u) return (float)__ddiv_rn(a, cos_p);   // within 4 ulp of a tie: decide exactly
    return (float)q2;
}

// What attenuate_fluxes derives from the height a 3D segment starts at (solver.c:38-45, 55-58,
// 84-87): the offset inside the fine interval and the stencil rows.  Writes length and offset
// of the record (if `store`), returns the top byte of the record code (r0 | which << 6).
template <bool FAST>
__device__ __forceinline__ uint32_t emit_geometry(const WalkParams &w, float z_start, float ds, uint32_t slot, bool store)
{
    int iq;
    float iq_f;
    if (FAST) {
        // (int)(z / dz) by FMA quotient + round-toward-zero add (verified, see interval_by_fma)
        float q = __fmul_rn(z_start, w.fine_rdz);
        const float rem = __fmaf_rn(-q, w.dz_fine, z_start);
        q = __fmaf_rn(rem, w.fine_rdz, q);
        const float m = __fadd_rz(q, 12582912.0f);
        iq = __float_as_int(m) - 0x4B400000;
        iq_f = __fadd_rn(m, -12582912.0f);   // exact
    } else {
        iq = (int)__fdiv_rn(z_start, w.dz_fine);
        iq_f = (float)iq;
    }
    float zin = __fsub_rn(z_start, __fmul_rn(w.dz_fine, __fadd_rn(iq_f, 0.5f)));
    // iq % fai: fai_magic = floor(2^32 / fai) + 1 is exact for iq < 2^26 (2 <= fai <= 63); fai = 1 (a flat source
    // over a single fine interval, solver.c:1040-1138) has no 32-bit reciprocal: the remainder is 0
    const int fine = w.fai == 1 ? 0
                   : (unsigned)iq < (1u << 20) ? iq - (int)__umulhi((uint32_t)iq, w.fai_magic) * w.fai : iq % w.fai;
    int r0 = fine, which = 0;
    if (w.axial_exp == 2) {
        if (fine == 0) { r0 = 0; zin = __fsub_rn(zin, w.dz_fine); }
        else if (fine == w.fai - 1) { r0 = w.fai - 3; zin = __fadd_rn(zin, w.dz_fine); }
        else r0 = fine - 1;
        which = fine - r0;
    }
    if (store) {
        w.rec_ds[slot] = ds;
        w.rec_zin[slot] = zin;
    }
    return (uint32_t)r0 | ((uint32_t)which << 6);
}

// select element r of a small register array without dynamic indexing
template <int KPT, class T>
__device__ __forceinline__ T pick(const T (&v)[KPT], int r)
{
    T x = v[0];
#pragma unroll
    for (int q = 1; q < KPT; q++) x = (r == q) ? v[q] : x;
    return x;
}
template <int KPT, class T>
__device__ __forceinline__ void put(T (&v)[KPT], int r, T x)
{
#pragma unroll
    for (int q = 0; q < KPT; q++) v[q] = (r == q) ? x : v[q];
}

// BLOCK = false: the stack belongs to ONE warp (Z <= 32 KPT).  BLOCK = true: to the blockDim.x / 32 warps of a
// CTA (Z <= blockDim.x KPT): same per-lane work, the stack-wide prefixes and totals cross the warps through
// `xw` (2 x 33 u64 of shared memory) and two __syncthreads per 2D segment.
template <int KPT, bool FILL, bool UP, bool FAST, bool BLOCK>
__device__ __forceinline__ void walk_stack_warp(const WalkParams &w, const long long pair, const long long local,
                                                const long long i, const int j, const int lane,
                                                uint4 *scratch, unsigned long long *xw = nullptr)
{
    // `scratch` (one warp per stack only): this warp's WALK_SCRATCH_BYTES of shared memory -- crossing rays on their
    // way to the lane that walks them (16 bytes each, at most 32 KPT <= 128), then the source regions of the step
    uint32_t *const qbuf = reinterpret_cast<uint32_t *>(scratch);
    const int warp = BLOCK ? (int)(threadIdx.x >> 5) : 0;
    const int n_warps = BLOCK ? (int)(blockDim.x >> 5) : 1;
    const int Z = w.Z;
    const int n_seg = w.n_seg[i];
    const float *len = w.seg_len + w.seg_start[i];
    const double cos_p = w.cos_p[j], sin_p = w.sin_p[j];
    const double rcos = __ddiv_rn(1.0, cos_p);
    const long long t0 = pair * Z;
    const int k0 = (warp * 32 + lane) * KPT;

    float zh[KPT];
    uint32_t made_total[KPT];          // count pass: segments of the ray so far
    uint32_t cursor[KPT], end[KPT];    // emitting pass: next record slot of the ray / one past its last
    unsigned long long dg[4] = {0, 0, 0, 0};
    bool bad_height = false;
#pragma unroll
    for (int r = 0; r < KPT; r++) {
        zh[r] = (k0 + r < Z) ? w.z_height[t0 + k0 + r] : 0.f;
        bad_height = bad_height || !(zh[r] >= 0.0f && zh[r] <= w.node_dz_f);
        made_total[r] = 0;
        cursor[r] = end[r] = 0;
    }
    if (FAST && !FILL && bad_height) atomicOr(w.flags, 1u);   // the host repeats the sweep with IEEE divisions

    unsigned long long serial_at = 0;   // serial index of the first segment of the current step
    uint32_t base = 0;                  // record slot of (ray 0, segment 0) of this stack, batch-relative
    if (FILL) {
        // records of a stack are segment-major: slot(ray k, its j-th segment) = base + j * Zs + k, so
        // the rays of a stack write neighbouring words at every step (full 32-byte sectors);
        // cursor[r] = j, end[r] = the ray's segment count from pass 1 (rays cut from the window stop there)
        serial_at = w.pair_base[pair];
        base = (uint32_t)(w.rec_base[pair] - w.batch_first_record);
#pragma unroll
        for (int r = 0; r < KPT; r++) end[r] = (k0 + r < Z) ? w.seg_count[t0 + k0 + r] : 0u;
    }

    int lo = 0, hi = Z;
    float s_mine = 0.f;
    for (int n = 0; n < n_seg; n++) {
        // s_full = length / sin(p_angle): float / double -> double -> float  (solver.c:382-383)
        if ((n & 31) == 0) {
            const int nn = n + lane;
            s_mine = nn < n_seg ? (float)__ddiv_rn((double)len[nn], sin_p) : 0.f;
        }
        const float s_full = __shfl_sync(0xffffffffu, s_mine, n & 31);
        const double advance = __dmul_rn((double)s_full, cos_p);   // s * cos(p_angle), first trial of every ray
        const bool last = (n == n_seg - 1);

        // ---- every ray still in the window walks (tentatively: upward rays may turn out to be cut).
        // Rays that stay in their fine interval: one 3D segment of length s_full -- branch-free.
        uint32_t cnt[KPT], pcode[KPT];      // segments of this step; top bytes of their record codes (first 4)
        float z_after[KPT];
        uint32_t exits = 0, crossing = 0;   // bit r: ray r left the domain / crosses an interval boundary
#pragma unroll
        for (int r = 0; r < KPT; r++) {
            const int k = k0 + r;
            const bool in_window = (k >= lo && k < hi);
            const float z = (float)__dadd_rn((double)zh[r], advance);
            const bool same = interval_of<UP, FAST>(w, z) == interval_of<UP, FAST>(w, zh[r]);
            cnt[r] = in_window ? 1u : 0u;
            z_after[r] = in_window ? z : zh[r];
            crossing |= (in_window && !same) ? (1u << r) : 0u;
            pcode[r] = 0;
            if (FILL)
                pcode[r] = emit_geometry<FAST>(w, zh[r], s_full, base + cursor[r] * w.Zs + k, in_window && same && cursor[r] < end[r]);
        }
        if (last) {
#pragma unroll
            for (int r = 0; r < KPT; r++) {
                const int k = k0 + r;
                const float home = UP ? __fmul_rn(w.z_sep, (float)k) : __fmul_rn(w.z_sep, (float)(k + 1));
                z_after[r] = (k >= lo && k < hi) ? home : zh[r];
            }
        }
        // Rays that cross a boundary: the general loop of solver.c:409-525, one code copy.
        // (One warp per stack only: with several warps per stack -- tall stacks -- the per-lane form measured
        // 4 % faster, 2.5 against 2.6 ms on the small problem.)
        constexpr bool COMPACT = MOC_WALK_COMPACT_CROSSING && !BLOCK;
        // A quarter of the rays cross at a typical step, one to three per lane: walked where they live, the warp
        // needs max-over-lanes passes with a third of its lanes busy.  Instead the crossing rays are numbered
        // across the warp (one shuffle scan) and dealt out one per lane through shared memory: (height, ray,
        // first slot, last slot) go out, (segments | exit, height, stencil bytes) come back.
        if (COMPACT && __any_sync(0xffffffffu, crossing != 0)) {
            const uint32_t mine = __popc(crossing);
            const uint32_t incl = warp_inclusive_scan_u32(mine, lane);
            const uint32_t n_cross = __shfl_sync(0xffffffffu, incl, 31);
            {
                uint32_t at = incl - mine;
#pragma unroll
                for (int r = 0; r < KPT; r++)
                    if ((crossing >> r) & 1u)
                        scratch[at++] = make_uint4(__float_as_uint(zh[r]), (uint32_t)(k0 + r), cursor[r], end[r]);
            }
            __syncwarp();
            for (uint32_t x = lane; x < n_cross; x += 32) {
                const uint4 ray = scratch[x];
                const int k = (int)ray.y;
                const float home = UP ? __fmul_rn(w.z_sep, (float)k) : __fmul_rn(w.z_sep, (float)(k + 1));
                const uint32_t j0 = ray.z, j_end = ray.w;
                float s = s_full, z_cur = __uint_as_float(ray.x);
                int c = interval_of<UP, FAST>(w, z_cur);
                uint32_t made = 0, pc = 0, left = 0;
                bool finished = false;
                do {
                    bool out = false;
                    // float z = z_height + s * cos(p_angle)   -- double product, double sum, narrowed
                    float z = (float)__dadd_rn((double)z_cur, __dmul_rn((double)s, cos_p));
                    float ds;
                    if (interval_of<UP, FAST>(w, z) == c) {
                        finished = true;
                        ds = s;
                    } else {
                        c += UP ? 1 : -1;
                        z = (float)__dmul_rn(w.fine_dz, (double)c);   // (float)c is exact, so is (double)(float)c
                        ds = div_by_cos<FAST>(__fsub_rn(z, z_cur), cos_p, rcos);
                        s = __fsub_rn(s, ds);
                        if (s <= 0.0f) finished = true;
                        if (z <= 0.0f || z >= w.node_dz_f) {
                            finished = true;
                            out = true;
                            left = 1u << 31;
                        }
                    }
                    if (FILL) {
                        const uint32_t slot = base + (j0 + made) * w.Zs + k;
                        const bool store = j0 + made < j_end;
                        const uint32_t byte = emit_geometry<FAST>(w, z_cur, ds, slot, store);
                        if (made < 4) pc |= byte << (8 * made);
                        else if (store) w.rec_code[slot] = byte << 24;   // fifth and later: parked in the record
                    }
                    made++;
                    z_cur = (last || out) ? home : z;   // solver.c:514-523, after EVERY 3D segment
                } while (!finished);
                scratch[x] = make_uint4(made | left, __float_as_uint(z_cur), pc, 0u);
            }
            __syncwarp();
            {
                uint32_t at = incl - mine;
#pragma unroll
                for (int r = 0; r < KPT; r++)
                    if ((crossing >> r) & 1u) {
                        const uint4 res = scratch[at++];
                        cnt[r] = res.x & 0x7fffffffu;
                        exits |= (res.x >> 31) << r;
                        z_after[r] = __uint_as_float(res.y);
                        if (FILL) pcode[r] = res.z;
                    }
            }
            __syncwarp();   // the scratch is reused (source regions below, the next step's rays)
        }
        // the per-lane form: every lane walks its own crossing rays, one after the other
        while (!COMPACT && __any_sync(0xffffffffu, crossing != 0)) {
            if (crossing) {
                const int r = __ffs(crossing) - 1;
                crossing &= crossing - 1;
                const int k = k0 + r;
                const float home = UP ? __fmul_rn(w.z_sep, (float)k) : __fmul_rn(w.z_sep, (float)(k + 1));
                const uint32_t j0 = pick<KPT>(cursor, r), j_end = pick<KPT>(end, r);
                float s = s_full, z_cur = pick<KPT>(zh, r);
                int c = interval_of<UP, FAST>(w, z_cur);
                uint32_t made = 0, pc = 0;
                bool finished = false;
                do {
                    bool out = false;
                    // float z = z_height + s * cos(p_angle)   -- double product, double sum, narrowed
                    float z = (float)__dadd_rn((double)z_cur, __dmul_rn((double)s, cos_p));
                    float ds;
                    if (interval_of<UP, FAST>(w, z) == c) {
                        finished = true;
                        ds = s;
                    } else {
                        c += UP ? 1 : -1;
                        z = (float)__dmul_rn(w.fine_dz, (double)c);   // (float)c is exact, so is (double)(float)c
                        ds = div_by_cos<FAST>(__fsub_rn(z, z_cur), cos_p, rcos);
                        s = __fsub_rn(s, ds);
                        if (s <= 0.0f) finished = true;
                        if (z <= 0.0f || z >= w.node_dz_f) {
                            finished = true;
                            out = true;
                            exits |= 1u << r;
                        }
                    }
                    if (FILL) {
                        const uint32_t slot = base + (j0 + made) * w.Zs + k;
                        const bool store = j0 + made < j_end;
                        const uint32_t byte = emit_geometry<FAST>(w, z_cur, ds, slot, store);
                        if (made < 4) pc |= byte << (8 * made);
                        else if (store) w.rec_code[slot] = byte << 24;   // fifth and later: parked in the record
                    }
                    made++;
                    z_cur = (last || out) ? home : z;   // solver.c:514-523, after EVERY 3D segment
                } while (!finished);
                put<KPT>(cnt, r, made);
                put<KPT>(z_after, r, z_cur);
                if (FILL) put<KPT>(pcode, r, pc);
            }
        }

        // ---- one scan over the stack: segments and exits below every ray
        uint32_t lane_cnt = 0;
#pragma unroll
        for (int r = 0; r < KPT; r++) lane_cnt += cnt[r];
        const uint32_t lane_exits = __popc(exits);
        uint32_t cnt_before, exits_before;   // exclusive prefixes at this lane's first ray
        if (__any_sync(0xffffffffu, lane_cnt >= (1u << 18))) {
            const unsigned long long v = ((unsigned long long)lane_exits << 32) | lane_cnt;
            const unsigned long long ex = warp_inclusive_scan_u64(v, lane) - v;
            cnt_before = (uint32_t)ex;
            exits_before = (uint32_t)(ex >> 32);
        } else {
            const uint32_t v = (lane_exits << 24) | lane_cnt;
            const uint32_t ex = warp_inclusive_scan_u32(v, lane) - v;
            cnt_before = ex & 0xffffffu;
            exits_before = ex >> 24;
        }
        if (BLOCK) {
            // warp totals -> shared memory -> every thread adds the totals of the warps below its own
            if (lane == 31) xw[warp] = ((unsigned long long)(exits_before + lane_exits) << 32) | (cnt_before + lane_cnt);
            __syncthreads();
            unsigned long long below = 0;
            for (int q = 0; q < warp; q++) below += xw[q];
            cnt_before += (uint32_t)below;
            exits_before += (uint32_t)(below >> 32);
        }

        // ---- who is really processed: upward rays see end_stacked shrink as lower rays exit
        uint32_t taken_cnt = 0, taken_exits = 0, taken = 0;
        uint32_t first_serial[KPT];   // serial index (relative to serial_at) of ray r's first segment
        {
            uint32_t run_cnt = cnt_before, run_exits = exits_before;
#pragma unroll
            for (int r = 0; r < KPT; r++) {
                const int k = k0 + r;
                const bool in_window = (k >= lo && k < hi);
                const bool take = in_window && (!UP || (k + (int)run_exits < hi));
                first_serial[r] = run_cnt;
                if (take) {
                    taken |= 1u << r;
                    taken_cnt += cnt[r];
                    taken_exits += (exits >> r) & 1u;
                }
                run_cnt += cnt[r];
                run_exits += (exits >> r) & 1u;
            }
        }
        // what this warp really processed: its segments are draws serial_at + warp_first .. + warp_cnt - 1
        uint32_t done_exits = __reduce_add_sync(0xffffffffu, taken_exits);
        uint32_t done_cnt = __reduce_add_sync(0xffffffffu, taken_cnt);
        if (FILL) {
            // ---- the source region of every segment just written: rand() draw number `serial`
            // (solver.c:476-483).
            // The draws of a step are consecutive, so the hash + remainder (two thirds of the work here) is
            // dealt out evenly -- lane l takes draws l, l + 32, ... -- and parked in shared memory; the lanes
            // then only pick up the regions of their own rays' segments (a lane's rays hold 1 to ~6 segments,
            // walking them with the hash inside kept 13 lanes of 32 busy).
            const uint32_t warp_cnt = done_cnt;
            const uint32_t warp_first = __shfl_sync(0xffffffffu, cnt_before, 0);
            const bool staged = (MOC_WALK_STAGED_HASH && !BLOCK) && warp_cnt <= WALK_SCRATCH_BYTES / 4;
            if (staged) {
                for (uint32_t t = lane; t < warp_cnt; t += 32) {
                    const uint32_t draw = moc_rand31(w.seed, w.rand_base + serial_at + warp_first + t);
                    qbuf[t] = w.mod_fast ? fastmod31(draw, w.n_regions, w.mod_magic, w.mod_shift) : draw % w.n_regions;
                }
                __syncwarp();
            }
            auto tally_digest = [&](uint32_t qsr, uint32_t byte, unsigned long long serial) {
                const unsigned long long row = (unsigned long long)qsr * w.fai + (byte & 63u) + (byte >> 6);
                dg[0] += 1ull;
                dg[1] += row;
                dg[2] += (row + 1ull) * (2ull * serial + 1ull);
                dg[3] ^= mix64(serial * 0x100000001B3ULL + row);
            };
            // Staged: the first segment of every ray (for three rays in four the only one) is finished
            // branch-free, unrolled over the lane's rays; the loop below is left with the further segments of
            // the rays that crossed.  Otherwise the loop takes every segment.
            uint32_t todo = taken;
            const uint32_t m0 = staged ? 1u : 0u;
            if (staged) {
                todo = 0;
#pragma unroll
                for (int r = 0; r < KPT; r++) {
                    if ((taken >> r) & 1u) {
                        const uint32_t byte = pcode[r] & 0xffu;
                        const uint32_t qsr = qbuf[first_serial[r] - warp_first];
                        w.rec_code[base + cursor[r] * w.Zs + (k0 + r)] = qsr | (byte << 24);
                        if (w.digest) tally_digest(qsr, byte, serial_at + first_serial[r]);
                        todo |= cnt[r] > 1u ? (1u << r) : 0u;
                    }
                }
            }
            // One code copy; a lane runs through its own rays' segments.
            uint32_t m = 0, cnt_r = 0, cur = 0, pc = 0, rel = 0;
            for (;;) {
                if (m == cnt_r) {
                    if (!todo) break;
                    const int r = __ffs(todo) - 1;
                    todo &= todo - 1;
                    m = m0;
                    cnt_r = pick<KPT>(cnt, r);
                    pc = pick<KPT>(pcode, r);
                    rel = pick<KPT>(first_serial, r);
                    cur = base + pick<KPT>(cursor, r) * w.Zs + (k0 + r);
                }
                const uint32_t slot = cur + m * w.Zs;
                const uint32_t byte = m < 4 ? ((pc >> (8 * m)) & 0xffu) : (w.rec_code[slot] >> 24);
                const unsigned long long serial = serial_at + rel + m;
                uint32_t qsr;
                if (staged) qsr = qbuf[rel - warp_first + m];
                else {
                    const uint32_t draw = moc_rand31(w.seed, w.rand_base + serial);
                    qsr = w.mod_fast ? fastmod31(draw, w.n_regions, w.mod_magic, w.mod_shift) : draw % w.n_regions;
                }
                w.rec_code[slot] = qsr | (byte << 24);
                if (w.digest) tally_digest(qsr, byte, serial);
                m++;
            }
            if (staged) __syncwarp();   // every lane has read its regions before the scratch is written again
        }
#pragma unroll
        for (int r = 0; r < KPT; r++) {
            if ((taken >> r) & 1u) {
                zh[r] = z_after[r];
                if (FILL) cursor[r] += cnt[r];
                else made_total[r] += cnt[r];
            }
        }
        if (BLOCK) {
            // second exchange: what the warps really processed (the first one's slots are still being read)
            if (lane == 0) xw[33 + warp] = ((unsigned long long)done_exits << 32) | done_cnt;
            __syncthreads();
            unsigned long long all = 0;
            for (int q = 0; q < n_warps; q++) all += xw[33 + q];
            done_cnt = (uint32_t)all;
            done_exits = (uint32_t)(all >> 32);
        }
        if (UP) hi -= (int)done_exits;
        else lo += (int)done_exits;
        serial_at += done_cnt;
    }

    if (FILL) {
#pragma unroll
        for (int r = 0; r < KPT; r++)
            if (k0 + r < Z) w.z_height[t0 + k0 + r] = zh[r];
        if (w.digest) {
            // order-independent digest: three sums and one xor
#pragma unroll
            for (int q = 0; q < 4; q++) {
                unsigned long long v = dg[q];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d);
                    v = (q == 3) ? (v ^ o) : (v + o);
                }
                if (lane == 0) {
                    if (q == 3) atomicXor(w.digest + q, v);
                    else atomicAdd(w.digest + q, v);
                }
            }
        }
    } else {
        uint32_t mine = 0;
#pragma unroll
        for (int r = 0; r < KPT; r++) {
            if (k0 + r < Z) w.seg_count[t0 + k0 + r] = made_total[r];
            mine += made_total[r];
        }
        unsigned long long tot = mine;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
        uint32_t longest = 0;
#pragma unroll
        for (int r = 0; r < KPT; r++) longest = max(longest, made_total[r]);
        longest = __reduce_max_sync(0xffffffffu, longest);
        if (BLOCK) {
            __syncthreads();   // the last step's exchange has been read by everybody
            if (lane == 0) {
                xw[warp] = tot;
                xw[33 + warp] = longest;
            }
            __syncthreads();
            if (warp == 0 && lane == 0) {
                unsigned long long t = 0, l = 0;
                for (int q = 0; q < n_warps; q++) {
                    t += xw[q];
                    l = xw[33 + q] > l ? xw[33 + q] : l;
                }
                w.pair_count[pair] = t;
                w.pair_max[pair] = (uint32_t)l;
            }
        } else if (lane == 0) {
            w.pair_count[pair] = tot;
            w.pair_max[pair] = longest;   // the longest ray decides how many record rows the stack needs
        }
    }
}

// One launch handles the stacks of ONE direction (upward rays: polar index j < P/2; downward: the
// rest), so only one instantiation of the walk is resident in the instruction caches at a time.
// n_dir = number of stacks of that direction among pairs [first_pair, first_pair + n_pairs);
// dir_before = number of them among pairs [0, first_pair).  4 warps (stacks) per CTA.
// The grid may be smaller than the work (warps stride over the stacks): the emitting pass of batch b+1
// runs as a few resident CTAs per SM UNDER the attenuation of batch b (sweep_core), filling the issue
// slots that kernel leaves idle, instead of as a full grid in front of it.
// resident CTAs per SM the walk is compiled for (0: the compiler's choice: 96 registers emitting, 78 counting).
// Counting at 7 CTAs (72 registers, 32 bytes spilled): 11.0 instead of 11.5 ms per sweep; emitting at 6 (80 registers):
// no change (profiles/r02_K0_launch_bounds_ab.log).
#ifndef MOC_WALK_MIN_BLOCKS_FILL
#define MOC_WALK_MIN_BLOCKS_FILL 0
#endif
#ifndef MOC_WALK_MIN_BLOCKS_COUNT
#define MOC_WALK_MIN_BLOCKS_COUNT 7
#endif
template <int KPT, bool FILL, bool UP, bool FAST>
__global__ void __launch_bounds__(128, FILL ? MOC_WALK_MIN_BLOCKS_FILL : MOC_WALK_MIN_BLOCKS_COUNT) stack_walk_warp_kernel(const WalkParams w, long long dir_before, long long n_dir)
{
    __shared__ uint4 scratch[4][WALK_SCRATCH_BYTES / 16];
    const int lane = threadIdx.x & 31;
    const int H = w.P / 2, per_track = UP ? H : w.P - H;
    const long long step = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < n_dir; q += step) {
        const long long target = dir_before + q;   // whole warps stay or leave together
        const long long i = target / per_track;
        const int j = (int)(target - i * per_track) + (UP ? 0 : H);
        const long long pair = i * w.P + j;
        walk_stack_warp<KPT, FILL, UP, FAST, false>(w, pair, pair - w.first_pair, i, j, lane, scratch[threadIdx.x >> 5]);
    }
}

// Taller stacks (128 < Z <= 2048): one CTA of ceil(Z / 128) warps per stack, four rays per lane, the same walk.
template <bool FILL, bool UP, bool FAST>
__global__ void __launch_bounds__(512) stack_walk_block_kernel(const WalkParams w, long long dir_before, long long n_dir)
{
    __shared__ unsigned long long xw[66];
    const int lane = threadIdx.x & 31;
    const int H = w.P / 2, per_track = UP ? H : w.P - H;
    for (long long q = blockIdx.x; q < n_dir; q += gridDim.x) {
        const long long target = dir_before + q;
        const long long i = target / per_track;
        const int j = (int)(target - i * per_track) + (UP ? 0 : H);
        const long long pair = i * w.P + j;
        walk_stack_warp<4, FILL, UP, FAST, true>(w, pair, pair - w.first_pair, i, j, lane, nullptr, xw);
        __syncthreads();   // xw is reused by the next stack
    }
}
