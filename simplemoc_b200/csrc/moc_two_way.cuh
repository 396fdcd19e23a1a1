/* moc_two_way.cuh -- the reference's second sweep, two_way_transport_sweep (solver.c:556-891): compiled into the
 * reference, called from nowhere (main.c calls transport_sweep only).  SURVEY 8f row f3.
 *
 * What is different from the sweep the other kernels implement:
 *   - the ray trace: the fine interval comes from the ray height TRUNCATED TO AN INTEGER first,
 *     `(int) z_height / fine_delta_z` (solver.c:686), the distance left in the 2D segment is cut at the node
 *     boundary up front (solver.c:690-707), steps go to the next interval EDGE (solver.c:716-741) -- with the
 *     truncated interval the "next" edge can lie behind the ray, and the step length ds is then NEGATIVE;
 *   - the z-stack window [begin_stacked, end_stacked) moves once per 2D segment, after all of its rays
 *     (solver.c:790-797), and no ray height is reset inside the walk;
 *   - every ray then retraces its own segments last-to-first with -mu into the BACKWARD angular flux
 *     (solver.c:811-836) and all ray heights are reset (solver.c:839-845).
 *
 * A negative ds makes the reference's table lookup read in front of its table (undefined); oracle/moc_oracle.c
 * answers those from cell 0 and so does MODE 6 below.  This is a coverage path: plain kernels, one thread per ray
 * in the trace and one warp per track in the attenuation, no tuning.
 */
#pragma once
// included by moc_kernels.cuh inside namespace moc, after moc_attenuate.cuh

struct TwoWayParams {
    float *z_height;                    // [T3] end-of-walk heights in, reset heights out (solver.c:839-845)
    float z_sep, dz_fine;
    int axial_exp;
    unsigned long long *digest_back;    // [4] optional: (track * 4096 + segment, tally row) of the backward pass
    unsigned int *flags;                // bit 1: a walk hit the iteration guard, bit 2: a negative fine interval
};

constexpr int TWO_WAY_GUARD = 1 << 20;   // 3D segments of one ray inside one 2D segment: far beyond any real case

__device__ __forceinline__ void digest_add(unsigned long long *dg, unsigned long long key, unsigned long long row)
{
    dg[0] += 1ull;
    dg[1] += row;
    dg[2] += (row + 1ull) * (2ull * key + 1ull);
    dg[3] ^= mix64(key * 0x100000001B3ULL + row);
}

// attenuate_fluxes' view of a start height (solver.c:38-45, 55-58, 84-87; flat source: solver.c:1058-1060)
__device__ __forceinline__ void start_geometry(float zstart, float dz, int fai, int axial_exp, float &zin, int &fine,
                                               int &r0, int &which, unsigned int *flags)
{
    const int iq = (int)__fdiv_rn(zstart, dz);
    zin = __fsub_rn(zstart, __fmul_rn(dz, __fadd_rn((float)iq, 0.5f)));
    fine = iq % fai;
    if (fine < 0) {
        // a ray below the node: the reference indexes in front of its arrays.  Never seen; kept inside the region.
        fine += fai;
        if (flags) atomicOr(flags, 4u);
    }
    r0 = fine;
    which = 0;
    if (axial_exp == 2) {
        if (fine == 0) { r0 = 0; zin = __fsub_rn(zin, dz); }
        else if (fine == fai - 1) { r0 = fai - 3; zin = __fadd_rn(zin, dz); }
        else r0 = fine - 1;
        which = fine - r0;
    }
}

// One 2D segment of one ray: solver.c:683-788.  mu = (float)cos(polar).  EMIT=false only counts.
template <bool UP, bool EMIT>
__device__ __forceinline__ int two_way_segment(const WalkParams &w, float &zh, float s_full, float mu, int n_intervals,
                                               bool &completed, unsigned long long serial, unsigned long long slot,
                                               unsigned long long *dg)
{
    // (int) z_height / fine_delta_z: the cast binds first; int / double, truncated again on assignment
    int interval = (int)__ddiv_rn((double)(int)zh, w.fine_dz);
    // distance to the node boundary: (double - float) / float narrowed, resp. float / float
    const float bound = UP ? (float)__ddiv_rn(__dsub_rn(w.node_dz, (double)zh), (double)mu) : __fdiv_rn(-zh, mu);
    float s = s_full;
    completed = false;
    if (!(s_full < bound)) {
        s = bound;
        completed = true;
    }
    int made = 0;
    bool finished = false;
    while (!finished) {
        const float edge = (float)__dmul_rn((double)(UP ? interval + 1 : interval), w.fine_dz);
        const float dz_to_edge = __fsub_rn(edge, zh);
        const float s_to_edge = __fdiv_rn(dz_to_edge, mu);
        float ds, z;
        if (s_to_edge < s) {
            interval += UP ? 1 : -1;
            ds = s_to_edge;
            z = __fadd_rn(zh, dz_to_edge);
        } else {
            ds = s;
            z = __fadd_rn(zh, __fmul_rn(s, mu));
        }
        s = __fsub_rn(s, ds);
        if (s <= 0.f || interval < 0 || interval >= n_intervals) finished = true;
        if (made >= TWO_WAY_GUARD) {
            finished = true;
            atomicOr(w.flags, 2u);
        }
        if (EMIT) {
            float zin;
            int fine, r0, which;
            start_geometry(zh, w.dz_fine, w.fai, w.axial_exp, zin, fine, r0, which, w.flags);
            const unsigned long long m = serial + (unsigned)made;
            const unsigned int qsr = moc_rand31(w.seed, w.rand_base + m) % w.n_regions;
            const unsigned long long at = slot + (unsigned long long)made * w.Zs;
            w.rec_ds[at] = ds;
            w.rec_zin[at] = zin;
            w.rec_code[at] = pack_code(qsr, (uint32_t)r0, (uint32_t)which);
            if (dg) digest_add(dg, m, (unsigned long long)qsr * w.fai + fine);
        }
        made++;
        zh = z;
    }
    return made;
}

// CTA = one (2D track, polar angle) z-stack, thread = KPT consecutive rays; same two passes, record layout and
// serial numbering as stack_walk_kernel.  FILL leaves the END-of-walk heights in z_height: the backward pass of
// attenuate_two_way_kernel starts there and resets them.
template <int KPT, bool FILL>
__global__ void two_way_walk_kernel(const WalkParams w)
{
    __shared__ unsigned long long scratch[34];

    const long long pair = w.first_pair + blockIdx.x;
    const long long i = pair / w.P;
    const int j = (int)(pair % w.P);
    const bool up = j < w.P / 2;                    // solver.c:653-657
    const int n_seg = w.n_seg[i];
    const float *len = w.seg_len + w.seg_start[i];
    const double sin_p = w.sin_p[j];
    const float mu = (float)w.cos_p[j];             // solver.c:671: float mu = cos(p_angle)
    const int n_intervals = (int)lrint(w.node_dz / w.fine_dz);   // cai * fai
    const long long t0 = pair * w.Z;
    const int k0 = threadIdx.x * KPT;

    float zh[KPT];
    uint32_t made_total[KPT];
    unsigned long long cursor[KPT];
    unsigned long long dg[4] = {0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < KPT; r++) {
        made_total[r] = 0;
        cursor[r] = 0;
        zh[r] = (k0 + r < w.Z) ? w.z_height[t0 + k0 + r] : 0.f;
    }
    unsigned long long serial_at = 0;
    if (FILL) {
        serial_at = w.pair_base[pair];
        const unsigned long long base = w.rec_base[pair] - w.batch_first_record;
#pragma unroll
        for (int r = 0; r < KPT; r++) cursor[r] = base + k0 + r;
    }

    int lo = 0, hi = w.Z;
    for (int n = 0; n < n_seg; n++) {
        const float s_full = (float)__ddiv_rn((double)len[n], sin_p);   // solver.c:678
        uint32_t cnt[KPT], exits[KPT];
        float z_after[KPT];
        unsigned long long mine = 0;
#pragma unroll
        for (int r = 0; r < KPT; r++) {
            const int k = k0 + r;
            cnt[r] = exits[r] = 0;
            z_after[r] = zh[r];
            if (k >= lo && k < hi) {
                bool completed;
                float z = zh[r];
                cnt[r] = up ? two_way_segment<true, false>(w, z, s_full, mu, n_intervals, completed, 0, 0, nullptr)
                            : two_way_segment<false, false>(w, z, s_full, mu, n_intervals, completed, 0, 0, nullptr);
                exits[r] = completed ? 1u : 0u;
                z_after[r] = z;
                mine += ((unsigned long long)exits[r] << 32) | cnt[r];
            }
        }
        unsigned long long everything;
        unsigned long long run = block_exclusive_scan(mine, scratch, everything);
#pragma unroll
        for (int r = 0; r < KPT; r++) {
            if (cnt[r]) {
                if (FILL) {
                    bool completed;
                    float z = zh[r];
                    const unsigned long long serial = serial_at + (run & 0xffffffffull);
                    if (up) two_way_segment<true, true>(w, z, s_full, mu, n_intervals, completed, serial, cursor[r],
                                                        w.digest ? dg : nullptr);
                    else two_way_segment<false, true>(w, z, s_full, mu, n_intervals, completed, serial, cursor[r],
                                                      w.digest ? dg : nullptr);
                    cursor[r] += (unsigned long long)cnt[r] * w.Zs;
                }
                zh[r] = z_after[r];
                made_total[r] += cnt[r];
            }
            run += ((unsigned long long)exits[r] << 32) | cnt[r];
        }
        // the window moves after all rays of the 2D segment (solver.c:790-797)
        if (up) hi -= (int)(everything >> 32);
        else lo += (int)(everything >> 32);
        serial_at += (everything & 0xffffffffull);
    }

    if (FILL) {
#pragma unroll
        for (int r = 0; r < KPT; r++)
            if (k0 + r < w.Z) w.z_height[t0 + k0 + r] = zh[r];
        if (w.digest) {
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
        uint32_t longest = 0;
#pragma unroll
        for (int r = 0; r < KPT; r++) {
            if (k0 + r < w.Z) w.seg_count[t0 + k0 + r] = made_total[r];
            mine += made_total[r];
            longest = max(longest, made_total[r]);
        }
        unsigned long long tot;
        block_exclusive_scan(mine, scratch, tot);
        if (threadIdx.x == 0) w.pair_count[pair] = tot;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) longest = max(longest, __shfl_xor_sync(0xffffffffu, longest, d));
        if ((threadIdx.x & 31) == 0) atomicMax(w.pair_max + pair, longest);
    }
}

// Both attenuation passes of one track per warp; lane owns groups lane, lane + 32, ... (G <= 512).
//   forward  (solver.c:744-760): the records of the walk, f_psi, +mu
//   backward (solver.c:811-836): the same segments last to first, b_psi, -mu; the start height of each step is the
//             ray height, which moves back by ds * mu after it; geometry recomputed from it as attenuate_fluxes does
// MODE: 6 = the reference's table (IEEE division, cells in front of the table answered from cell 0), 2 = SFU.
template <int MODE, bool FLAT>
__global__ void __launch_bounds__(128) attenuate_two_way_kernel(const AttenuateParams a, const TwoWayParams tw)
{
    extern __shared__ float s_tab[];
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
    constexpr int NG = 16;
    const int lane = threadIdx.x & 31;
    const long long t = a.first_track + (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= a.end_track) return;
    const int G = a.G;
    const uint32_t W = (uint32_t)a.pitch;
    const uint32_t n_rec = a.seg_count[t];
    const long long pair = t / a.Z;
    const int k = (int)(t - pair * a.Z);
    const int j = (int)(pair % a.P);
    const long long i = pair / a.P;
    const uint32_t at = (uint32_t)(a.rec_base[pair] - a.batch_first_record) + (uint32_t)k;
    const uint32_t Zs = (uint32_t)a.Zs;
    const float mu_f = a.mu[j];
    const float w0 = __fmul_rn(a.p_weight[t], a.az_weight[i]);   // solver.c:49
    float *psi_row = a.psi + (size_t)2 * (size_t)t * G;
    unsigned long long dgb[4] = {0, 0, 0, 0};
    float zh = tw.z_height[t];

    for (int pass = 0; pass < 2; pass++) {
        const float mu = pass ? -mu_f : mu_f;
        float psi[NG];
#pragma unroll
        for (int q = 0; q < NG; q++) {
            const int g = lane + 32 * q;
            psi[q] = g < G ? psi_row[pass * G + g] : 0.f;
        }
        SegmentScalars sc;
        sc.weight = FLAT ? __fmul_rn(w0, mu) : w0;               // solver.c:1064
        sc.b1 = mu * a.inv_2dz;
        sc.b3 = mu * mu * a.inv_2dz2;
        for (uint32_t s = 0; s < n_rec; s++) {
            const uint32_t n = pass ? n_rec - 1 - s : s;
            const float ds = __ldg(a.rec_ds + at + n * Zs);
            const uint32_t code = __ldg(a.rec_code + at + n * Zs);
            const uint32_t qsr = code & 0xffffffu;
            float zin;
            uint32_t r0, which;
            if (pass == 0) {
                zin = __ldg(a.rec_zin + at + n * Zs);
                r0 = (code >> 24) & 63u;
                which = code >> 30;
            } else {
                int fine, r, wh;
                start_geometry(zh, tw.dz_fine, a.fai, tw.axial_exp, zin, fine, r, wh, lane == 0 ? tw.flags : nullptr);
                r0 = (uint32_t)r;
                which = (uint32_t)wh;
                if (tw.digest_back && lane == 0)
                    digest_add(dgb, (unsigned long long)t * 4096ull + n, (unsigned long long)qsr * a.fai + fine);
                zh = __fsub_rn(zh, __fmul_rn(ds, mu_f));         // solver.c:833
            }
            sc.ds = ds;
            sc.a1 = zin * a.inv_2dz;
            sc.a2 = zin * zin * a.inv_2dz2;
            sc.b2 = 2.f * mu * a.inv_2dz2 * zin;
            const uint32_t o_row = (qsr * a.fai + r0) * W;
            const uint32_t o_sig = qsr * W;
            const uint32_t o_flx = o_row + which * W;
#pragma unroll
            for (int q = 0; q < NG; q++) {
                const int g = lane + 32 * q;
                if (g < G) {
                    const float sg = __ldg(a.sigT + o_sig + g);
                    float tally;
                    if (FLAT)
                        tally = attenuate_flat<MODE>(__ldg(a.fine_source + o_row + g), sg, psi[q], sc, tc);
                    else
                        tally = attenuate_groups<MODE, false>(__ldg(a.fine_source + o_row + g),
                                                              __ldg(a.fine_source + o_row + W + g),
                                                              __ldg(a.fine_source + o_row + 2 * W + g), sg, psi[q], sc, tc);
                    atomicAdd(a.fine_flux + o_flx + g, tally);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NG; q++) {
            const int g = lane + 32 * q;
            if (g < G) psi_row[pass * G + g] = psi[q];
        }
    }
    if (lane == 0) {
        const bool up = j < a.P / 2;
        tw.z_height[t] = up ? __fmul_rn(tw.z_sep, (float)k) : __fmul_rn(tw.z_sep, (float)(k + 1));
        if (tw.digest_back) {
            atomicAdd(tw.digest_back + 0, dgb[0]);
            atomicAdd(tw.digest_back + 1, dgb[1]);
            atomicAdd(tw.digest_back + 2, dgb[2]);
            atomicXor(tw.digest_back + 3, dgb[3]);
        }
    }
}

