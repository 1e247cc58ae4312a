// k1h_fill.cuh -- K1h: the K1 inter-sequence fill with TWO pairs per lane group, packed 16 x 2.
//
// For short reads the scores fit 16 bits, so one 32-bit register carries the same DP cell of two
// different pairs (pair A in the low half, pair B in the high half) and every instruction of the
// recurrence does two cell updates:
//   * values are stored unsigned with a bias of 2^14 (HB_BIAS); "minus infinity" is HB_NEG.  No half
//     ever leaves [0, 65535], so a plain 32-bit IMAD adds a packed constant (b, a, the substitution
//     score pair) to both halves at once -- on the FMA pipe;
//   * the four maxima of the cell are VIMNMX.U16x2 with its two predicate outputs, and those
//     predicates ARE the reference's tie tests:
//         X  = max(xo, xe)   p = (xo >= xe)  <=>  x_trace == 'M'      (aligner.rs:444, open wins ties)
//         Y  = max(yo, ye)   p = (yo >= ye)  <=>  y_trace == 'M'      (aligner.rs:448)
//         m1 = max(X, d)     p = (X  >= d)   <=>  M == X when M != Y  (aligner.rs:458)
//         M  = max(Y, m1)    p = (Y  >= m1)  <=>  M == Y              (aligner.rs:455, Y tested first)
//     each predicate adds its bit to the trace word of its pair with a predicated IMAD / LEA;
//   * the substitution score pair comes from one LDS into a 256-entry table indexed by
//     (row residue A, row residue B, column residue A, column residue B), pre-biased by -a.
// 4 ALU-pipe + ~9 FMA/ALU instructions per TWO cells instead of 8 + 6 per cell in K1.
// Restrictions (the host falls back to K1 otherwise -- still GPU, never CPU): not local mode,
// <= 4 distinct residues per side, single band, (len1 + len2 + 2) * max|score| <= HB_RANGE.
#pragma once
#include "bg_common.cuh"
#include "k1_fill.cuh"

namespace bg {

constexpr int32_t HB_BIAS = 1 << 14;
constexpr int32_t HB_NEG = 1 << 10;      // biased "minus infinity": true value -(2^14 - 2^10)
constexpr int32_t HB_RANGE = 12000;      // max |true score| the host admits for this kernel
constexpr int32_t HB_MAXABS = 512;       // max |a|, |b|, |s|

__device__ __forceinline__ uint32_t hb_pack(int32_t v) { return (uint32_t)v * 65537u; }   // same value in both halves
__device__ __forceinline__ int32_t hb_half(uint32_t v, int h) { return (int32_t)((v >> (16 * h)) & 0xffffu); }

// v = max.u16x2(x, y); per half: if (x >= y) w += bit   (w_lo for pair A, w_hi for pair B)
__device__ __forceinline__ uint32_t hb_max_acc(uint32_t x, uint32_t y, uint32_t& w_lo, uint32_t& w_hi, uint32_t one, uint32_t bit) {
    uint32_t v;
    asm("{\n\t.reg .pred ph, pl;\n\t.reg .u16 r0, r1, r2, r3;\n\t"
        "max.u16x2 %0, %3, %4;\n\t"
        "mov.b32 {r0, r1}, %0;\n\tmov.b32 {r2, r3}, %3;\n\t"
        "setp.eq.u16 pl, r0, r2;\n\tsetp.eq.u16 ph, r1, r3;\n\t"
        "@pl mad.lo.u32 %1, %5, %6, %1;\n\t@ph mad.lo.u32 %2, %5, %6, %2;\n\t}"
        : "=r"(v), "+r"(w_lo), "+r"(w_hi) : "r"(x), "r"(y), "r"(one), "r"(bit));
    return v;
}
__device__ __forceinline__ uint32_t hb_add(uint32_t x, uint32_t one, uint32_t y) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(one), "r"(y));
    return d;
}

// Launch geometry: a warp holds 32/L lane groups, each with two pairs: slots (2g, 2g+1) of the warp's
// 2*(32/L) consecutive slots.  Trace word (t, k, half h, lane) at
//   trace_off + (((t * K) + k) * 2 + h) * 32 + lane          (single band).
template <int L, int C>
__global__ void __launch_bounds__(128, (C <= 19 ? 4 : 3)) k1h_fill(const FillArgs A) {
    constexpr int GP = 32 / L;
    constexpr int K = (C + 7) / 8;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint8_t s_row[256];
    __shared__ uint8_t s_col[256];
    __shared__ uint32_t s_pack[256];   // [(rA*4 + rB) * 16 + (cA*4 + cB)] = (sA - a) + (sB - a) * 65536

    for (int x = threadIdx.x; x < 256; x += blockDim.x) {
        s_row[x] = A.row_code[x]; s_col[x] = A.col_code[x];
        const int rA = (x >> 6) & 3, rB = (x >> 4) & 3, cA = (x >> 2) & 3, cB = x & 3;
        const int32_t sA = ((rA < A.n_rows && cA < A.n_cols) ? A.table[rA * A.n_cols + cA] : 0) - A.a;
        const int32_t sB = ((rB < A.n_rows && cB < A.n_cols) ? A.table[rB * A.n_cols + cB] : 0) - A.a;
        s_pack[x] = (uint32_t)(sA + sB * 65536);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int g = lane / L, p = lane % L;
    const uint32_t warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t slot0 = (warp_global * GP + g) * 2;

    PairDesc d[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        d[h].n = 0; d[h].m = 0; d[h].steps = 0; d[h].nbands = 0; d[h].pair_id = 0xFFFFFFFFu;
        d[h].a_off = d[h].b_off = d[h].trace_off = d[h].bnd_off = d[h].pad_off = 0;
        if (slot0 + h < A.n_slots) d[h] = A.desc[slot0 + h];
    }
    bool has[2]; uint32_t n[2], m[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) { has[h] = d[h].pair_id != 0xFFFFFFFFu; n[h] = has[h] ? d[h].n : 0; m[h] = has[h] ? d[h].m : 0; }
    const uint32_t n_max = max(n[0], n[1]), m_max = max(m[0], m[1]);
    const uint32_t steps_mine = has[0] ? d[0].steps : (has[1] ? d[1].steps : 0u);
    const uint32_t steps_w = __reduce_max_sync(FULL, steps_mine);
    const uint64_t trace_off = has[0] ? d[0].trace_off : d[1].trace_off;

    const int32_t a = A.a, b = A.b;
    const uint32_t one = (uint32_t)A.one;
    const uint32_t a2 = (uint32_t)(a * 65537), b2 = (uint32_t)(b * 65537);
    const int mode = A.mode;
    const bool row_gap = (mode == M_GLOBAL || mode == M_FITTING);
    const bool col_gap = (mode == M_GLOBAL);
    const bool track_col = (mode == M_SEMIGLOBAL || mode == M_FITTING);
    const bool track_row = (mode == M_SEMIGLOBAL || mode == M_OVERLAP);
    const uint8_t* sa[2] = {A.residues + d[0].a_off, A.residues + d[1].a_off};
    const uint8_t* sb[2] = {A.residues + d[0].b_off, A.residues + d[1].b_off};
    bool bad_residue = false;

    uint32_t p_m[2], c_m[2]; bool col_lane[2];
    int32_t rbest[2], cbest[2], corner[2]; uint32_t rj[2], ci[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t mcol0 = m[h] ? m[h] - 1 : 0;
        p_m[h] = mcol0 / C; c_m[h] = mcol0 % C;
        col_lane[h] = (m[h] > 0) && ((uint32_t)p == p_m[h]);
        rbest[h] = INT32_MIN; rj[h] = 0;
        cbest[h] = border_row(row_gap, a, b, m[h]); ci[h] = 0;
        corner[h] = border_col(col_gap, a, b, n[h]);
        if (p == 0) { rbest[h] = border_col(col_gap, a, b, n[h]); rj[h] = 0; }
    }

    const uint32_t jbase = (uint32_t)p * C;
    const bool lane_has_cols = jbase < m_max;
    uint32_t cc[C], MuA[C], Xu[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const uint32_t j0 = jbase + c;
        uint32_t code[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            code[h] = 0;
            if (has[h] && j0 < m[h]) {
                code[h] = s_col[sb[h][j0]];
                if (code[h] > 3u) { bad_residue = true; code[h] = 0; }
            }
        }
        cc[c] = (code[0] * 4u + code[1]) * 4u;
        MuA[c] = hb_pack(border_row(row_gap, a, b, j0 + 1) + a + HB_BIAS);
        Xu[c] = hb_pack(HB_NEG);
    }
    // row-n capture for a half: this lane's columns of row n_h are in MuA (biased by a + HB_BIAS)
    auto capture_row = [&](int h) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const uint32_t j = jbase + c + 1;
            const int32_t v = hb_half(MuA[c], h) - HB_BIAS - a;
            if (j <= m[h]) {
                if (track_row && v >= rbest[h]) { rbest[h] = v; rj[h] = j; }
                if (j == m[h]) corner[h] = v;
            }
        }
    };
    if (has[0] && n[0] == 0) capture_row(0);
    if (has[1] && n[1] == 0) capture_row(1);

    uint32_t MdiagA = hb_pack(border_row(row_gap, a, b, jbase) + a + HB_BIAS);
    uint32_t MlastA = hb_pack(a + HB_BIAS), Ylast = hb_pack(HB_NEG);
    uint32_t rcur = 0;
    auto load_rows = [&](uint32_t base) -> uint32_t {   // combined row code rA*4 + rB of row base + p
        const uint32_t idx = base + (uint32_t)p;
        uint32_t cd[2] = {0, 0};
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (has[h] && idx < n[h]) { cd[h] = s_row[sa[h][idx]]; if (cd[h] > 3u) { bad_residue = true; cd[h] = 0; } }
        return cd[0] * 4u + cd[1];
    };
    uint32_t cur_blk = 0, next_blk = load_rows(0);

    for (uint32_t t = 0; t < steps_w; ++t) {
        if ((t & (L - 1)) == 0) { cur_blk = next_blk; next_blk = load_rows(t + L); }
        const uint32_t r0 = __shfl_sync(FULL, cur_blk, (int)(t & (L - 1)), L);
        uint32_t MlA = __shfl_up_sync(FULL, MlastA, 1, L);
        uint32_t Yl = __shfl_up_sync(FULL, Ylast, 1, L);
        uint32_t r = __shfl_up_sync(FULL, rcur, 1, L);
        const uint32_t i0 = t - (uint32_t)p;
        const bool active = i0 < n_max;
        if (p == 0) {
            r = r0;
            MlA = hb_pack(border_col(col_gap, a, b, i0 + 1) + a + HB_BIAS);
            Yl = hb_pack(HB_NEG);
        }
        rcur = r;
        if (active) {
            uint32_t diagA = MdiagA, leftA = MlA, Y = Yl;
            uint32_t wA[K], wB[K];
#pragma unroll
            for (int k = 0; k < K; ++k) { wA[k] = 0; wB[k] = 0; }
            const unsigned char* rowp = reinterpret_cast<const unsigned char*>(s_pack) + r * 64u;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                uint32_t& wa_ = wA[c >> 3];
                uint32_t& wb_ = wB[c >> 3];
                const uint32_t sh = 4u * (c & 7);
                const uint32_t upA = MuA[c];
                const uint32_t X = hb_max_acc(upA, hb_add(Xu[c], one, b2), wa_, wb_, one, TR_XOPEN << sh);
                Y = hb_max_acc(leftA, hb_add(Y, one, b2), wa_, wb_, one, TR_YOPEN << sh);
                const uint32_t s2 = *reinterpret_cast<const uint32_t*>(rowp + cc[c]);
                const uint32_t m1 = hb_max_acc(X, hb_add(diagA, one, s2), wa_, wb_, one, TR_XEQ << sh);
                const uint32_t mx = hb_max_acc(Y, m1, wa_, wb_, one, TR_YEQ << sh);
                const uint32_t mxA = hb_add(mx, one, a2);
                diagA = upA; leftA = mxA;
                MuA[c] = mxA; Xu[c] = X;
            }
            MlastA = leftA; Ylast = Y; MdiagA = MlA;
            if (A.want_trace && lane_has_cols) {
                uint32_t* tp = A.trace + trace_off + (uint64_t)t * (uint64_t)(K * 64) + lane;
#pragma unroll
                for (int k = 0; k < K; ++k) { tp[k * 64] = wA[k]; tp[k * 64 + 32] = wB[k]; }
            }
            if (track_col) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v = MuA[0];
#pragma unroll
                    for (int c = 1; c < C; ++c) v = (c_m[h] == (uint32_t)c) ? MuA[c] : v;
                    const int32_t vv = hb_half(v, h) - HB_BIAS - a;
                    if (col_lane[h] && i0 < n[h] && vv > cbest[h]) { cbest[h] = vv; ci[h] = i0 + 1; }
                }
            }
            if (i0 + 1 == n[0]) capture_row(0);
            if (i0 + 1 == n[1]) capture_row(1);
        }
    }
    if (bad_residue) atomicOr(A.err_flag, 1u);

#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (track_row) {
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) {
                const int32_t ov = __shfl_xor_sync(FULL, rbest[h], o, L);
                const uint32_t oj = __shfl_xor_sync(FULL, rj[h], o, L);
                if (ov > rbest[h] || (ov == rbest[h] && oj > rj[h])) { rbest[h] = ov; rj[h] = oj; }
            }
        }
        const int src = g * L + (int)p_m[h];
        const int32_t cbest0 = __shfl_sync(FULL, cbest[h], src);
        const uint32_t ci0 = __shfl_sync(FULL, ci[h], src);
        const int32_t corner0 = __shfl_sync(FULL, corner[h], src);
        if (p == 0 && has[h]) {
            EndCell e; e.flags = 0;
            switch (mode) {
            case M_GLOBAL: e.score = corner0; e.k = n[h]; e.l = m[h]; break;
            case M_FITTING: e.score = cbest0; e.k = ci0; e.l = m[h]; break;
            case M_OVERLAP: e.score = rbest[h]; e.k = n[h]; e.l = rj[h]; break;
            default:
                if (cbest0 > rbest[h]) { e.score = cbest0; e.k = ci0; e.l = m[h]; e.flags = 1; }
                else { e.score = rbest[h]; e.k = n[h]; e.l = rj[h]; }
                break;
            }
            A.end[slot0 + h] = e;
        }
    }
}

}  // namespace bg
