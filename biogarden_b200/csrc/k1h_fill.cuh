// k1h_fill.cuh -- K1h: the K1 inter-sequence fill with TWO pairs per lane group, packed 16 x 2.
//
// For short reads the scores fit 16 bits, so one 32-bit register carries the same DP cell of two
// different pairs (pair A in the low half, pair B in the high half) and every instruction of the
// recurrence does two cell updates:
//   * values are stored unsigned with a bias of 2^15 (HB_BIAS); "minus infinity" is HB_NEG.  No half
//     ever leaves [0, 65535], so a plain 32-bit IMAD adds a packed constant to both halves at once
//     -- on the FMA pipe;
//   * every value lives in an ANTI-DIAGONAL frame: cell (i, j) stores H(i, j) - (i + j) * b.  A gap
//     extension moves one anti-diagonal on, so "X_up + b" and "Y_left + b" need no instruction at all:
//         X^(i,j) = max(S(i-1,j), X^(i-1,j))      S = M^ + (a - b)   (the state the registers carry)
//         Y^(i,j) = max(S(i,j-1), Y^(i,j-1))
//         M^(i,j) = max(Y^, X^, S(i-1,j-1) + (s - a - b))
//     which leaves two adds per cell pair (diagonal + score, M^ + (a - b)); ties are frame-independent
//     because all operands of one max belong to the same cell.  "Minus infinity" is never added to,
//     so it cannot drift.  The end-cell captures take the frame out again;
//   * the four maxima of the cell are VIMNMX.U16x2 with its two predicate outputs, and those
//     predicates ARE the reference's tie tests:
//         X  = max(xo, xe)   p = (xo >= xe)  <=>  x_trace == 'M'      (aligner.rs:444, open wins ties)
//         Y  = max(yo, ye)   p = (yo >= ye)  <=>  y_trace == 'M'      (aligner.rs:448)
//         m1 = max(X, d)     p = (X  >= d)   <=>  M == X when M != Y  (aligner.rs:458)
//         M  = max(Y, m1)    p = (Y  >= m1)  <=>  M == Y              (aligner.rs:455, Y tested first)
//     each predicate adds its bit to the trace word of its column with one predicated instruction; the 8
//     accumulations per cell pair are split between the ALU pipe (VIADD) and the FMA pipe (IMAD) by
//     HB_PIPES so that both pipes carry ~8 instructions per cell pair (ncu of the first version, all
//     on the FMA pipe: fmaheavy 78 % busy, ALU 41 %);
//   * the substitution score pair (s - a - b of pair A | of pair B):
//       PROF (every s - a - b fits a signed byte -- the usual case): once per ROW one LDS.64 fetches the row's two
//       byte profiles (scores of row residue A / B against the four column codes); per cell ONE PRMT whose selector
//       is the column's constant picks byte cA of profile A and byte cB of profile B and replicates their sign
//       bits into the upper bytes, i.e. builds the sign-extended 16 x 2 pair; VIADD.16x2 adds it to the diagonal
//       (a packed add: a negative low half must not borrow from the high half);
//       otherwise: one LDS into a 256-entry table indexed by (row residue A, row residue B, column residue A,
//       column residue B), pre-biased by -(a + b) with the borrow folded in, and an IMAD.
// PROF: 15 instructions per TWO cells (4 VIMNMX.U16x2, 8 accumulations, PRMT, VIADD.16x2, IMAD), no shared-memory
// access in the cell loop; table form: 16 (... 2 adds, 1 address, 1 LDS).
//
// Trace layout ("row blocks"): a lane keeps one accumulator per COLUMN and collects HB_TB = 4
// systolic steps in it -- pair A's nibbles in bits 0..15, pair B's in bits 16..31 -- then writes its
// CW = round_up(C, 4) words with 128-bit stores:
//     word (block tb, lane, column c) at trace_off + ((tb / 4) * 32 + lane) * 4 CW + (tb % 4) * CW + c
//     nibble of step t = 4 * tb + r: bits [4r, 4r + 4) (+16 for pair B).
// HB_TG = 4 consecutive row blocks of a lane are contiguous, so 16 steps x C columns of a lane group's
// pair form one 4 CW-word tile (192 B for C = 10) next to the tiles of the neighbouring lanes.  The
// traceback walk moves along a diagonal, i.e. to column c - 1 and step t - 1: consecutive cells of a
// path sit in adjacent words and a path stays ~10 steps inside one or two 128-byte lines (the first,
// step-major layout cost the walk a DRAM access per step: 20 KB per 150 bp pair, ncu).
//
// The two pairs of a lane group always have the same number of rows (the host leaves the second slot
// empty otherwise), so row n of both halves is complete exactly when the lane's last active step is.
// Restrictions (the host falls back to K1 otherwise -- still GPU, never CPU): not local mode,
// <= 4 distinct residues per side, single band, (len1 + len2 + 2) * max|score| <= HB_RANGE.
#pragma once
#include "bg_args.cuh"
#include "k1_fill.cuh"

namespace bg {

// HB_PIPES (template argument below): two bits per max (X, Y, m1, M): low / high half accumulates on the ALU pipe


__device__ __forceinline__ uint32_t hb_pack(int32_t v) { return (uint32_t)v * 65537u; }   // same value in both halves
__device__ __forceinline__ int32_t hb_half(uint32_t v, int h) { return (int32_t)((v >> (16 * h)) & 0xffffu); }

// v = max.u16x2(x, y); per half: if (x >= y) w |= bit (pair A) / bit << 16 (pair B).  PIPES bit 0 / bit 1:
// the low / high half's bit is added on the ALU pipe instead of the FMA pipe (predicated IMAD
// w = one * bit + w; the bits of one word never collide, so add == or).
template <int PIPES>
__device__ __forceinline__ uint32_t hb_max_acc(uint32_t x, uint32_t y, uint32_t& w, uint32_t one, uint32_t bit) {
    uint32_t v;
    const uint32_t lo = bit, hi = bit << 16;
    if (PIPES == 0)
        asm("{\n\t.reg .pred ph, pl;\n\t.reg .u16 r0, r1, r2, r3;\n\t"
            "max.u16x2 %0, %2, %3;\n\t"
            "mov.b32 {r0, r1}, %0;\n\tmov.b32 {r2, r3}, %2;\n\t"
            "setp.eq.u16 pl, r0, r2;\n\tsetp.eq.u16 ph, r1, r3;\n\t"
            "@pl mad.lo.u32 %1, %4, %5, %1;\n\t@ph mad.lo.u32 %1, %4, %6, %1;\n\t}"
            : "=r"(v), "+r"(w) : "r"(x), "r"(y), "r"(one), "r"(lo), "r"(hi));
    else if (PIPES == 1)
        asm("{\n\t.reg .pred ph, pl;\n\t.reg .u16 r0, r1, r2, r3;\n\t"
            "max.u16x2 %0, %2, %3;\n\t"
            "mov.b32 {r0, r1}, %0;\n\tmov.b32 {r2, r3}, %2;\n\t"
            "setp.eq.u16 pl, r0, r2;\n\tsetp.eq.u16 ph, r1, r3;\n\t"
            "@pl or.b32 %1, %1, %5;\n\t@ph mad.lo.u32 %1, %4, %6, %1;\n\t}"
            : "=r"(v), "+r"(w) : "r"(x), "r"(y), "r"(one), "r"(lo), "r"(hi));
    else if (PIPES == 2)
        asm("{\n\t.reg .pred ph, pl;\n\t.reg .u16 r0, r1, r2, r3;\n\t"
            "max.u16x2 %0, %2, %3;\n\t"
            "mov.b32 {r0, r1}, %0;\n\tmov.b32 {r2, r3}, %2;\n\t"
            "setp.eq.u16 pl, r0, r2;\n\tsetp.eq.u16 ph, r1, r3;\n\t"
            "@pl mad.lo.u32 %1, %4, %5, %1;\n\t@ph or.b32 %1, %1, %6;\n\t}"
            : "=r"(v), "+r"(w) : "r"(x), "r"(y), "r"(one), "r"(lo), "r"(hi));
    else
        asm("{\n\t.reg .pred ph, pl;\n\t.reg .u16 r0, r1, r2, r3;\n\t"
            "max.u16x2 %0, %2, %3;\n\t"
            "mov.b32 {r0, r1}, %0;\n\tmov.b32 {r2, r3}, %2;\n\t"
            "setp.eq.u16 pl, r0, r2;\n\tsetp.eq.u16 ph, r1, r3;\n\t"
            "@pl or.b32 %1, %1, %5;\n\t@ph or.b32 %1, %1, %6;\n\t}"
            : "=r"(v), "+r"(w) : "r"(x), "r"(y), "r"(one), "r"(lo), "r"(hi));
    return v;
}
__device__ __forceinline__ uint32_t hb_add(uint32_t x, uint32_t one, uint32_t y) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(one), "r"(y));
    return d;
}

__device__ __forceinline__ uint32_t hb_add16x2(uint32_t x, uint32_t y) {   // VIADD.16x2
    uint32_t d;
    asm("add.u16x2 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(y));
    return d;
}
__device__ __forceinline__ uint32_t hb_prmt(uint32_t lo, uint32_t hi, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel));
    return d;
}

// Launch geometry: a warp holds 32/L lane groups, each with two pairs: slots (2g, 2g+1) of the warp's
// 2*(32/L) consecutive slots.  TRACK: the last-column / last-row end-cell scans of semiglobal, fitting
// and overlap are compiled in (global only needs the corner cell).
template <int L, int C, bool TRACK, int HB_PIPES, int MINB, bool PROF>
__global__ void __launch_bounds__(128, MINB) k1h_fill(const FillArgs A) {
    constexpr int GP = 32 / L;
    constexpr int CW = (C + 3) & ~3;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint8_t s_row[256];
    __shared__ uint8_t s_col[256];
    // !PROF: [(rA*4 + rB) * 16 + (cA*4 + cB)] = (sA - a - b) + (sB - a - b) * 65536
    // PROF:  [(rA*4 + rB) * 2 + h] = the four bytes (s(r_h, c) - a - b) & 0xff, c = 0..3: the row's score profile
    __shared__ __align__(8) uint32_t s_pack[PROF ? 32 : 256];

    for (int x = threadIdx.x; x < 256; x += blockDim.x) {
        s_row[x] = A.row_code[x]; s_col[x] = A.col_code[x];
        if (!PROF) {
            const int rA = (x >> 6) & 3, rB = (x >> 4) & 3, cA = (x >> 2) & 3, cB = x & 3;
            const int32_t sA = ((rA < A.n_rows && cA < A.n_cols) ? A.table[rA * A.n_cols + cA] : 0) - A.a - A.b;
            const int32_t sB = ((rB < A.n_rows && cB < A.n_cols) ? A.table[rB * A.n_cols + cB] : 0) - A.a - A.b;
            s_pack[x] = (uint32_t)(sA + sB * 65536);
        } else if (x < 32) {
            const int r = (x & 1) ? ((x >> 1) & 3) : ((x >> 3) & 3);
            uint32_t v = 0;
            for (int c = 0; c < 4; ++c) {
                const int32_t sc = ((r < A.n_rows && c < A.n_cols) ? A.table[r * A.n_cols + c] : 0) - A.a - A.b;
                v |= ((uint32_t)sc & 0xffu) << (8 * c);
            }
            s_pack[x] = v;
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int g = lane / L, p = lane % L;
    const uint32_t warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t slot0 = (warp_global * GP + g) * 2;

    // descriptors: only what the loop needs stays live (lengths, sequence pointers)
    uint32_t n[2], m[2], steps_mine = 0;
    const uint8_t* sa[2]; const uint8_t* sb[2];
    uint64_t trace_off = 0;
    bool has[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        n[h] = 0; m[h] = 0; has[h] = false; sa[h] = A.residues; sb[h] = A.residues;
        if (slot0 + h < A.n_slots) {
            const PairDesc& d = A.desc[slot0 + h];
            if (d.pair_id != 0xFFFFFFFFu) {
                has[h] = true; n[h] = d.n; m[h] = d.m;
                sa[h] = A.residues + d.a_off; sb[h] = A.residues + d.b_off;
                steps_mine = d.steps; trace_off = d.trace_off;
            }
        }
    }
    const uint32_t n_max = max(n[0], n[1]), m_max = max(m[0], m[1]);   // host: n[0] == n[1] when both slots are used
    const uint32_t steps_w = __reduce_max_sync(FULL, steps_mine);       // multiple of HB_TB

    const int32_t a = A.a, b = A.b;
    const uint32_t one = (uint32_t)A.one;
    const uint32_t amb2 = (uint32_t)((a - b) * 65537);
    const int mode = A.mode;
    const bool row_gap = (mode == M_GLOBAL || mode == M_FITTING);
    const bool col_gap = (mode == M_GLOBAL);
    // state of cell (i, j) in the registers: S = M(i, j) - (i + j) * b + (a - b) + HB_BIAS
    auto to_state = [&](int32_t mval, uint32_t i, uint32_t j) -> uint32_t {
        return hb_pack(mval - (int32_t)(i + j) * b + (a - b) + HB_BIAS);
    };
    auto from_state = [&](uint32_t packed, int h, uint32_t i, uint32_t j) -> int32_t {
        return hb_half(packed, h) - HB_BIAS - (a - b) + (int32_t)(i + j) * b;
    };
    // column border as a packed linear function of the row: S(i, 0) = colS0 + i * colS1 for i >= 1
    // (global: a + (i-1) b - i b + (a - b) = 2 (a - b), constant; otherwise -i b + (a - b))
    const uint32_t colS1 = (uint32_t)(((col_gap ? b : 0) - b) * 65537);
    const uint32_t colS0 = hb_pack((col_gap ? a - b : 0) + (a - b) + HB_BIAS);
    uint32_t bad_residue = 0;

    const uint32_t jbase = (uint32_t)p * C;
    uint32_t cc[C], MuA[C], Xu[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const uint32_t j0 = jbase + c;
        uint32_t code[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            code[h] = 0;
            if (j0 < m[h]) {
                code[h] = s_col[sb[h][j0]];
                if (code[h] > 3u) { bad_residue = 1; code[h] = 0; }
            }
        }
        // PROF: PRMT selector -- byte 0 <- profile A [cA], byte 1 <- its sign, byte 2 <- profile B [cB], byte 3 <- its sign
        cc[c] = PROF ? (code[0] | ((code[0] | 8u) << 4) | ((code[1] | 4u) << 8) | ((code[1] | 12u) << 12))
                     : (code[0] * 4u + code[1]) * 4u;
        MuA[c] = to_state(border_row(row_gap, a, b, j0 + 1), 0, j0 + 1);
        Xu[c] = hb_pack(HB_NEG);
    }

    // last-column scan state (TRACK): first max over the rows of column m (aligner.rs:376-380, 247-251)
    uint32_t c_m[2]; bool col_lane[2]; int32_t cbest[2]; uint32_t ci[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t mcol0 = m[h] ? m[h] - 1 : 0;
        c_m[h] = mcol0 % C;
        col_lane[h] = (m[h] > 0) && ((uint32_t)p == mcol0 / C);
        cbest[h] = border_row(row_gap, a, b, m[h]); ci[h] = 0;
    }

    uint32_t MdiagA = to_state(border_row(row_gap, a, b, jbase), 0, jbase);
    uint32_t MlastA = hb_pack(HB_BIAS), Ylast = hb_pack(HB_NEG);
    uint32_t rcur = 0;
    auto load_rows = [&](uint32_t base) -> uint32_t {   // combined row code rA*4 + rB of row base + p
        const uint32_t idx = base + (uint32_t)p;
        uint32_t cd[2] = {0, 0};
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (idx < n[h]) { cd[h] = s_row[sa[h][idx]]; if (cd[h] > 3u) { bad_residue = 1; cd[h] = 0; } }
        return cd[0] * 4u + cd[1];
    };
    uint32_t cur_blk = 0, next_blk = load_rows(0);
    const uint32_t first_lane = (p == 0) ? 0xffffffffu : 0u;   // integer mask instead of a live predicate
    const uint32_t tgs = (uint32_t)A.tg_shift;
    uint32_t* const tbase = A.trace + trace_off + (((uint64_t)lane * CW) << tgs);

    for (uint32_t t0 = 0; t0 < steps_w; t0 += HB_TB) {
        uint32_t w[C];
#pragma unroll
        for (int c = 0; c < C; ++c) w[c] = 0;
#pragma unroll
        for (int rr = 0; rr < HB_TB; ++rr) {
            const uint32_t t = t0 + rr;
            if (rr == 0 && (t0 & (L - 1)) == 0) { cur_blk = next_blk; next_blk = load_rows(t0 + L); }
            const uint32_t r0 = __shfl_sync(FULL, cur_blk, (int)(t & (L - 1)), L);
            uint32_t MlA = __shfl_up_sync(FULL, MlastA, 1, L);
            uint32_t Yl = __shfl_up_sync(FULL, Ylast, 1, L);
            uint32_t r = __shfl_up_sync(FULL, rcur, 1, L);
            const uint32_t i0 = t - (uint32_t)p;
            // lane 0 of the group: row residue from the block, column border instead of a left neighbour
            r = (r & ~first_lane) | (r0 & first_lane);
            MlA = (MlA & ~first_lane) | ((colS0 + (i0 + 1) * colS1) & first_lane);
            Yl = (Yl & ~first_lane) | (hb_pack(HB_NEG) & first_lane);
            rcur = r;
            if (i0 < n_max) {
                uint32_t leftA = MlA, Y = Yl;
                const unsigned char* rowp = reinterpret_cast<const unsigned char*>(s_pack) + r * 64u;
                uint2 rprof = make_uint2(0u, 0u);
                if (PROF) rprof = *reinterpret_cast<const uint2*>(s_pack + 2u * r);
                // diagonal terms first, HB_DG columns at a time, while the previous row's M is still in MuA[]: the cell
                // loop can then overwrite MuA[c] in place (carrying the old value along as "diag of the next column"
                // made ptxas rotate the whole register array: one extra move per cell)
                uint32_t dg[C];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    if (PROF) {
                        // one PRMT builds the sign-extended score pair from the row's two byte profiles; the packed
                        // add keeps the halves apart (a negative low half would otherwise borrow from the high one)
                        dg[c] = hb_add16x2(c ? MuA[c - 1] : MdiagA, hb_prmt(rprof.x, rprof.y, cc[c]));
                    } else {
                        const uint32_t s2 = *reinterpret_cast<const uint32_t*>(rowp + cc[c]);
                        dg[c] = hb_add(c ? MuA[c - 1] : MdiagA, one, s2);
                    }
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const uint32_t sh = 4u * rr;   // this step's nibble inside the column word (compile-time after unrolling)
                    const uint32_t X = hb_max_acc<HB_PIPES & 3>(MuA[c], Xu[c], w[c], one, TR_XOPEN << sh);
                    Y = hb_max_acc<(HB_PIPES >> 2) & 3>(leftA, Y, w[c], one, TR_YOPEN << sh);
                    const uint32_t m1 = hb_max_acc<(HB_PIPES >> 4) & 3>(X, dg[c], w[c], one, TR_XEQ << sh);
                    const uint32_t mx = hb_max_acc<(HB_PIPES >> 6) & 3>(Y, m1, w[c], one, TR_YEQ << sh);
                    leftA = hb_add(mx, one, amb2);
                    MuA[c] = leftA; Xu[c] = X;
                }
                MlastA = leftA; Ylast = Y; MdiagA = MlA;
                if (TRACK) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t v = MuA[0];
#pragma unroll
                        for (int c = 1; c < C; ++c) v = (c_m[h] == (uint32_t)c) ? MuA[c] : v;
                        const int32_t vv = from_state(v, h, i0 + 1, m[h]);
                        if (col_lane[h] && i0 < n[h] && vv > cbest[h]) { cbest[h] = vv; ci[h] = i0 + 1; }
                    }
                }
            }
        }
        if (jbase < m_max) {
            const uint32_t tb = t0 / HB_TB;
            uint4* q = reinterpret_cast<uint4*>(tbase + (((uint64_t)(tb >> tgs) * (32u * CW)) << tgs) + (tb & ((1u << tgs) - 1u)) * CW);
#pragma unroll
            for (int c4 = 0; c4 < CW / 4; ++c4) {
                uint4 v;
                v.x = w[4 * c4];
                v.y = (4 * c4 + 1 < C) ? w[(4 * c4 + 1 < C) ? 4 * c4 + 1 : 0] : 0u;
                v.z = (4 * c4 + 2 < C) ? w[(4 * c4 + 2 < C) ? 4 * c4 + 2 : 0] : 0u;
                v.w = (4 * c4 + 3 < C) ? w[(4 * c4 + 3 < C) ? 4 * c4 + 3 : 0] : 0u;
                q[c4] = v;
            }
        }
    }
    if (bad_residue) atomicOr(A.err_flag, 1u);

    // ---- end cells: MuA[] holds row n of both halves (row 0 borders if n == 0) ------------------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t mcol0 = m[h] ? m[h] - 1 : 0;
        const uint32_t p_m = mcol0 / C;
        int32_t rbest = INT32_MIN; uint32_t rj = 0;
        int32_t corner = border_col(col_gap, a, b, n[h]);
        if (p == 0) { rbest = border_col(col_gap, a, b, n[h]); rj = 0; }   // row n, column 0 candidate
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const uint32_t j = jbase + c + 1;
            const int32_t v = from_state(MuA[c], h, n[h], j);
            if (j <= m[h]) {
                if (TRACK && v >= rbest) { rbest = v; rj = j; }   // last row, last max (aligner.rs:369-373)
                if (j == m[h]) corner = v;
            }
        }
        if (TRACK) {
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) {
                const int32_t ov = __shfl_xor_sync(FULL, rbest, o, L);
                const uint32_t oj = __shfl_xor_sync(FULL, rj, o, L);
                if (ov > rbest || (ov == rbest && oj > rj)) { rbest = ov; rj = oj; }
            }
        }
        const int src = g * L + (int)p_m;
        const int32_t cbest0 = __shfl_sync(FULL, cbest[h], src);
        const uint32_t ci0 = __shfl_sync(FULL, ci[h], src);
        const int32_t corner0 = __shfl_sync(FULL, corner, src);
        if (p == 0 && has[h]) {
            EndCell e; e.flags = 0;
            switch (mode) {
            case M_GLOBAL: e.score = corner0; e.k = n[h]; e.l = m[h]; break;
            case M_FITTING: e.score = cbest0; e.k = ci0; e.l = m[h]; break;
            case M_OVERLAP: e.score = rbest; e.k = n[h]; e.l = rj; break;
            default:
                if (cbest0 > rbest) { e.score = cbest0; e.k = ci0; e.l = m[h]; e.flags = 1; }
                else { e.score = rbest; e.k = n[h]; e.l = rj; }
                break;
            }
            A.end[slot0 + h] = e;
        }
    }
}

}  // namespace bg
