// k1_fill.cuh -- K1: inter-sequence affine-gap DP fill for sm_100a.
//
// Replaces compute_scores_global / compute_scores_local plus the mode-specific border
// initialisation and end-cell selection of the reference (aligner.rs:437-509, 98-104, 163,
// 233-237, 299, 360, 112, 173-176, 247-251, 308-312, 369-380).
//
// Work decomposition (one lane group of L lanes per pair, 32/L pairs per warp):
//   * the (len2) columns are cut into bands of L*C columns; inside a band lane p owns the C
//     consecutive columns p*C .. p*C+C-1 and keeps M[i-1][j], X[i-1][j] for them in registers;
//   * rows stream through the lanes systolically: at step t lane p computes row t-p, so the
//     only cross-lane traffic is (M, Y, row residue) of the strip's last column, handed to
//     lane p+1 with __shfl_up_sync -- no shared-memory round trip on the recurrence;
//   * per cell the recurrence is the DPX forms VIADDMNMX x2 / VIMNMX3 plus the four tie tests that
//     become the 4-bit direction code (bg_common.cuh).  The kernel is bound by the SM's ALU pipe
//     (ncu: 92 % ALU, 15 % FMA in the first version), so everything that can run on the FMA pipe
//     does: registers hold M + a instead of M (the "+ a" of both gap-open terms disappears and the
//     diagonal term uses a score table pre-biased by -a), the two remaining adds are IMADs, and every
//     tie bit is folded into the trace word by a predicated IMAD instead of SEL / IADD3 chains.
//     One 32-bit word of codes per lane-step and 8 columns is written to HBM with a fully coalesced
//     128-byte warp store (layout in bg_common.cuh);
//   * substitution scores: for <= 4-letter row alphabets with |s| <= 127 every column keeps a
//     packed byte profile in a register and a single PRMT (byte select + sign extend) yields
//     s(row residue, column residue); otherwise the dense table sits in shared memory.
//   * pairs longer than one band loop over bands; the band's last column (M, Y) per row goes
//     through a small global scratch column (written by lane L-1, prefetched by lane 0).
#pragma once
#include "bg_args.cuh"

namespace bg {


__device__ __forceinline__ int32_t prmt_sx(uint32_t packed, uint32_t sel) {
    int32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(packed), "r"(0u), "r"(sel));
    return d;
}

// x * one + y with one == 1 only known at run time: an integer add that ptxas has to issue as IMAD,
// i.e. on the FMA pipe, which this kernel leaves idle otherwise.
__device__ __forceinline__ int32_t fma_add(int32_t x, int32_t one, int32_t y) {
    int32_t d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(one), "r"(y));
    return d;
}
// if (x == y) w += bit   -- compare on the ALU pipe, accumulate as a predicated IMAD (w = one*bit + w)
__device__ __forceinline__ void acc_if_eq(uint32_t& w, int32_t x, int32_t y, int32_t one, uint32_t bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, %4, %0;\n\t}"
        : "+r"(w) : "r"(x), "r"(y), "r"(one), "r"(bit));
}
// local mode, bit 0: (mx == Y) ? (mx == 0) : (mx == X)
__device__ __forceinline__ void acc_local_bit0(uint32_t& w, int32_t mx, int32_t X, int32_t Y, int32_t one, uint32_t bit) {
    asm("{\n\t.reg .pred px, py, pz, p;\n\t"
        "setp.eq.s32 px, %1, %2;\n\tsetp.eq.s32 py, %1, %3;\n\tsetp.eq.s32 pz, %1, 0;\n\t"
        "and.pred p, py, pz;\n\t{\n\t.reg .pred q, nq;\n\tnot.pred nq, py;\n\tand.pred q, nq, px;\n\tor.pred p, p, q;\n\t}\n\t"
        "@p mad.lo.u32 %0, %4, %5, %0;\n\t}"
        : "+r"(w) : "r"(mx), "r"(X), "r"(Y), "r"(one), "r"(bit));
}

// if (x <= y) w += bit
__device__ __forceinline__ void acc_if_le(uint32_t& w, int32_t x, int32_t y, int32_t one, uint32_t bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.s32 p, %1, %2;\n\t@p mad.lo.u32 %0, %3, %4, %0;\n\t}"
        : "+r"(w) : "r"(x), "r"(y), "r"(one), "r"(bit));
}
// if (x == y) w += bit_y; else if (x == z) w += bit_z      (one compare feeds both)
__device__ __forceinline__ void acc_eq_else_eq(uint32_t& w, int32_t x, int32_t y, int32_t z, int32_t one, uint32_t bit_y, uint32_t bit_z) {
    asm("{\n\t.reg .pred py, pz;\n\tsetp.eq.s32 py, %1, %2;\n\tsetp.eq.and.s32 pz, %1, %3, !py;\n\t"
        "@py mad.lo.u32 %0, %4, %5, %0;\n\t@pz mad.lo.u32 %0, %4, %6, %0;\n\t}"
        : "+r"(w) : "r"(x), "r"(y), "r"(z), "r"(one), "r"(bit_y), "r"(bit_z));
}
// x * k + y on the FMA pipe (k only known at run time)
__device__ __forceinline__ uint32_t fma_mad_u32(uint32_t x, uint32_t k, uint32_t y) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(k), "r"(y));
    return d;
}

// One LOCAL-mode cell (aligner.rs:477-506), written so that the ALU pipe -- the binding one -- sees 9
// instructions: the gap states are VIADDMNMX.RELU (add + max + clamp fused; the clamp comes after the trace
// byte is decided, which reads the unclamped operands), their "opened here" tests compare the FMA-pipe sums
// x_up + b <= m_up + a, bit 0 is accumulated from two disjoint predicates -- M == 0 (STOP; then M == Y == 0
// because Y >= 0) and M == X != Y -- and the running first maximum is kept as one unsigned key
// (M << 5 | 31 - column), decoded once per row instead of compare + three selects per cell.
__device__ __forceinline__ void local_cell(uint32_t& w, uint32_t sh, int32_t upA, int32_t& Xc, int32_t& Y, int32_t leftA,
                                           int32_t diag_plus_s, int32_t b, int32_t one, uint32_t k32, uint32_t colkey,
                                           uint32_t& rowkey, int32_t& mx) {
    const int32_t tx = fma_add(Xc, one, b);
    acc_if_le(w, tx, upA, one, TR_XOPEN << sh);
    const int32_t X = __viaddmax_s32_relu(Xc, b, upA);
    const int32_t ty = fma_add(Y, one, b);
    acc_if_le(w, ty, leftA, one, TR_YOPEN << sh);
    Y = __viaddmax_s32_relu(Y, b, leftA);
    mx = __vimax3_s32(diag_plus_s, X, Y);
    acc_eq_else_eq(w, mx, Y, X, one, TR_YEQ << sh, TR_XEQ << sh);
    acc_if_eq(w, mx, 0, one, TR_XEQ << sh);
    rowkey = max(rowkey, fma_mad_u32((uint32_t)mx, k32, colkey));
    Xc = X;
}

// M[0][j] and M[i][0].
__device__ __forceinline__ int32_t border_row(bool row_gap, int32_t a, int32_t b, uint32_t j) {
    return (row_gap && j > 0) ? a + (int32_t)(j - 1) * b : 0;
}
__device__ __forceinline__ int32_t border_col(bool col_gap, int32_t a, int32_t b, uint32_t i) {
    return (col_gap && i > 0) ? a + (int32_t)(i - 1) * b : 0;
}

// Register budget: the recurrence is a dependent chain (Y -> M -> M+a -> next Y), so the kernel needs
// >= 4 warps per scheduler to keep the ALU pipe fed (ncu: 3 warps/scheduler left it 29 % idle).
template <int L, int C, bool IS_LOCAL, bool PROF4>
__global__ void __launch_bounds__(128, (C <= 20 ? 4 : 3)) k1_fill(const FillArgs A) {
    constexpr int G = 32 / L;
    constexpr int K = (C + 7) / 8;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t* s_row = smem_raw;            // [256]
    uint8_t* s_col = smem_raw + 256;      // [256]
    int32_t* s_tab = reinterpret_cast<int32_t*>(smem_raw + 512);   // n_rows x (n_cols + 1), last column = 0 (padding)

    // The table is stored pre-biased by -a: registers carry M + a, so diag + (s - a) = M[i-1][j-1] + s.
    const int ncol1 = A.n_cols + 1;
    for (int x = threadIdx.x; x < 256; x += blockDim.x) { s_row[x] = A.row_code[x]; s_col[x] = A.col_code[x]; }
    for (int x = threadIdx.x; x < A.n_rows * ncol1; x += blockDim.x) {
        const int r = x / ncol1, c = x - r * ncol1;
        s_tab[x] = ((c < A.n_cols) ? A.table[r * A.n_cols + c] : 0) - A.a;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int g = lane / L, p = lane % L;
    const uint32_t warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t slot = warp_global * G + g;

    PairDesc d;
    d.n = 0; d.m = 0; d.steps = 0; d.nbands = 0; d.pair_id = 0xFFFFFFFFu;
    d.a_off = d.b_off = d.trace_off = d.bnd_off = d.pad_off = 0;
    if (slot < A.n_slots) d = A.desc[slot];
    const bool has_pair = d.pair_id != 0xFFFFFFFFu;
    const uint32_t n = has_pair ? d.n : 0, m = has_pair ? d.m : 0;
    const uint32_t my_nbands = has_pair ? d.nbands : 0;
    const uint32_t steps_w = __reduce_max_sync(FULL, has_pair ? d.steps : 0u);
    const uint32_t nbands_w = __reduce_max_sync(FULL, my_nbands);

    const int32_t a = A.a, b = A.b, one = A.one;
    const uint32_t k32 = (uint32_t)A.one << 5;
    const int mode = A.mode;
    const bool row_gap = (mode == M_GLOBAL || mode == M_FITTING);
    const bool col_gap = (mode == M_GLOBAL);
    const bool track_col = (mode == M_SEMIGLOBAL || mode == M_FITTING);
    const bool track_row = (mode == M_SEMIGLOBAL || mode == M_OVERLAP);
    const uint8_t* sa = A.residues + d.a_off;
    const uint8_t* sb = A.residues + d.b_off;
    bool bad_residue = false;

    // column-m bookkeeping (last-column scan / global corner)
    const uint32_t band_cols = (uint32_t)(L * C);
    const uint32_t mcol0 = m ? m - 1 : 0;
    const uint32_t bd_m = mcol0 / band_cols;
    const uint32_t p_m = (mcol0 % band_cols) / C;
    const uint32_t c_m = mcol0 % C;
    const bool col_lane = (m > 0) && ((uint32_t)p == p_m);

    // running end-cell state
    int32_t best = 0; uint32_t bi = 0, bj = 0;                          // local: first max, row-major (aligner.rs:173-176)
    int32_t rbest = INT32_MIN; uint32_t rj = 0;                         // last row, last max (>=)
    int32_t cbest = border_row(row_gap, a, b, m); uint32_t ci = 0;      // last column, first max (>); row 0 candidate
    int32_t corner = border_col(col_gap, a, b, n);                      // M[n][m] when m == 0
    if (p == 0) { rbest = border_col(col_gap, a, b, n); rj = 0; }       // row n, column 0 candidate

    for (uint32_t bd = 0; bd < nbands_w; ++bd) {
        const bool band_on = has_pair && bd < my_nbands;
        const uint32_t jbase = bd * band_cols + (uint32_t)p * C;   // 0-based index of this lane's first column
        const bool lane_has_cols = band_on && jbase < m;
        const bool last_band_for_pair = (bd + 1 == my_nbands);

        // per-column constants: packed (biased) score profile (PROF4) or byte offset of the table column;
        // per-column state: MuA = M[i-1][j] + a, Xu = X[i-1][j]
        uint32_t cprof[C];
        int32_t MuA[C], Xu[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const uint32_t j0 = jbase + c;
            uint32_t code = (uint32_t)A.n_cols;   // padding column (score 0)
            if (band_on && j0 < m) {
                code = s_col[sb[j0]];
                if (code == 0xFFu) { bad_residue = true; code = 0; }
            }
            if (PROF4) {
                uint32_t pk = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int32_t sv = (r < A.n_rows) ? s_tab[r * ncol1 + code] : -a;
                    pk |= ((uint32_t)sv & 0xffu) << (8 * r);
                }
                cprof[c] = pk;
            } else {
                cprof[c] = code * 4u;
            }
            MuA[c] = border_row(row_gap, a, b, j0 + 1) + a;
            Xu[c] = NEG_INF;
        }
        int32_t MdiagA = border_row(row_gap, a, b, jbase) + a;
        int32_t MlastA = a, Ylast = NEG_INF;
        uint32_t rcur = 0;

        // row residues: each lane of the group fetches one every L steps, lane 0 consumes one per step
        auto load_rows = [&](uint32_t base) -> uint32_t {
            const uint32_t idx = base + (uint32_t)p;
            uint32_t cd = 0;
            if (band_on && idx < n) {
                cd = s_row[sa[idx]];
                if (cd == 0xFFu) { bad_residue = true; cd = 0; }
            }
            return cd;
        };
        uint32_t cur_blk = 0, next_blk = load_rows(0);
        int2 bnd_in = make_int2(a, NEG_INF);
        if (bd > 0 && p == 0 && band_on && n > 0) bnd_in = __ldcg(A.bnd + d.bnd_off);

        for (uint32_t t = 0; t < steps_w; ++t) {
            if ((t & (L - 1)) == 0) { cur_blk = next_blk; next_blk = load_rows(t + L); }
            const uint32_t r0 = __shfl_sync(FULL, cur_blk, (int)(t & (L - 1)), L);
            int32_t MlA = __shfl_up_sync(FULL, MlastA, 1, L);
            int32_t Yl = __shfl_up_sync(FULL, Ylast, 1, L);
            uint32_t r = __shfl_up_sync(FULL, rcur, 1, L);
            const uint32_t i0 = t - (uint32_t)p;          // 0-based row; wraps (inactive) while t < p
            const bool active = band_on && i0 < n;
            {
                // lane 0 of a group: row residue from the block, left border (band 0) or the previous band's last column instead
                // of a left neighbour -- as selects; a branch on p == 0 made the warp run lane 0's arm on its own every step
                const bool first = (p == 0);
                int32_t bM, bY;
                if (bd == 0) { bM = border_col(col_gap, a, b, i0 + 1) + a; bY = NEG_INF; }      // bd is uniform in the warp
                else {
                    bM = bnd_in.x; bY = bnd_in.y;
                    if (first && band_on && i0 + 1 < n) bnd_in = __ldcg(A.bnd + d.bnd_off + i0 + 1);
                }
                r = first ? r0 : r;
                MlA = first ? bM : MlA;
                Yl = first ? bY : Yl;
            }
            rcur = r;
            if (active) {
                int32_t leftA = MlA, Y = Yl;
                uint32_t w[K];
#pragma unroll
                for (int k = 0; k < K; ++k) w[k] = 0;
                uint32_t sel; const unsigned char* rowp;
                if (PROF4) sel = r * 0x1111u + 0x8880u;
                else rowp = reinterpret_cast<const unsigned char*>(s_tab) + r * (uint32_t)(ncol1 * 4);
                uint32_t rowkey = 0;
                // diagonal terms first, while the previous row's M is still in MuA[]: the cell loop then overwrites MuA[c]
                // in place (carrying the old value along as the next column's diagonal made ptxas rotate the register
                // array: one extra move per cell)
                int32_t dg[C];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    int32_t sb_;   // s - a
                    if (PROF4) sb_ = prmt_sx(cprof[c], sel);
                    else sb_ = *reinterpret_cast<const int32_t*>(rowp + cprof[c]);
                    dg[c] = fma_add(c ? MuA[c - 1] : MdiagA, one, sb_);
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    uint32_t& wk = w[c >> 3];
                    const uint32_t sh = 4u * (c & 7);
                    const int32_t upA = MuA[c];
                    if (IS_LOCAL) {
                        int32_t mx;
                        local_cell(wk, sh, upA, Xu[c], Y, leftA, dg[c], b, one, k32, 31u - c, rowkey, mx);
                        leftA = fma_add(mx, one, a);
                        MuA[c] = leftA;
                        continue;
                    }
                    // aligner.rs:443-444: xo = M[i-1][j] + a is the register itself
                    const int32_t X = __viaddmax_s32(Xu[c], b, upA);
                    acc_if_eq(wk, X, upA, one, TR_XOPEN << sh);
                    // aligner.rs:447-448: yo = M[i][j-1] + a likewise
                    Y = __viaddmax_s32(Y, b, leftA);
                    acc_if_eq(wk, Y, leftA, one, TR_YOPEN << sh);
                    // aligner.rs:451-466
                    const int32_t mx = __vimax3_s32(dg[c], X, Y);
                    acc_if_eq(wk, mx, Y, one, TR_YEQ << sh);
                    acc_if_eq(wk, mx, X, one, TR_XEQ << sh);
                    leftA = fma_add(mx, one, a);
                    MuA[c] = leftA; Xu[c] = X;
                }
                MlastA = leftA; Ylast = Y; MdiagA = MlA;
                if (IS_LOCAL) {   // first maximum of this row's cells; strictly greater than everything above
                    const int32_t v = (int32_t)(rowkey >> 5);
                    if (v > best) { best = v; bi = i0 + 1; bj = jbase + 32u - (rowkey & 31u); }
                }
                if (A.want_trace && lane_has_cols) {
                    uint32_t* tp = A.trace + d.trace_off + ((uint64_t)(bd * d.steps + t) * K) * 32u + lane;
#pragma unroll
                    for (int k = 0; k < K; ++k) tp[k * 32] = w[k];
                }
                if (track_col && bd == bd_m) {
                    int32_t v = MuA[0];
#pragma unroll
                    for (int c = 1; c < C; ++c) v = (c_m == (uint32_t)c) ? MuA[c] : v;
                    v -= a;
                    if (col_lane && v > cbest) { cbest = v; ci = i0 + 1; }
                }
                if (p == L - 1 && !last_band_for_pair) A.bnd[d.bnd_off + i0] = make_int2(MlastA, Ylast);
            }
        }
        // MuA[] now holds row n of this band, biased by a (row 0 borders if n == 0)
        if (band_on) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const uint32_t j = jbase + c + 1;
                const int32_t v = MuA[c] - a;
                if (j <= m) {
                    if (track_row && v >= rbest) { rbest = v; rj = j; }
                    if (j == m) corner = v;
                }
            }
        }
        __syncwarp();
    }

    if (bad_residue) atomicOr(A.err_flag, 1u);

    // ---- end-cell selection across the lane group ------------------------------------
    if (track_row) {
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) {
            const int32_t ov = __shfl_xor_sync(FULL, rbest, o, L);
            const uint32_t oj = __shfl_xor_sync(FULL, rj, o, L);
            if (ov > rbest || (ov == rbest && oj > rj)) { rbest = ov; rj = oj; }   // last max
        }
    }
    if (IS_LOCAL) {
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) {
            const int32_t ov = __shfl_xor_sync(FULL, best, o, L);
            const uint32_t oi = __shfl_xor_sync(FULL, bi, o, L);
            const uint32_t oj = __shfl_xor_sync(FULL, bj, o, L);
            if (ov > best || (ov == best && (oi < bi || (oi == bi && oj < bj)))) { best = ov; bi = oi; bj = oj; }
        }
    }
    // column-m owner -> lane 0 of the group
    const int src = g * L + (int)p_m;
    const int32_t cbest0 = __shfl_sync(FULL, cbest, src);
    const uint32_t ci0 = __shfl_sync(FULL, ci, src);
    const int32_t corner0 = __shfl_sync(FULL, corner, src);

    if (p == 0 && has_pair) {
        EndCell e; e.flags = 0;
        switch (mode) {
        case M_GLOBAL: e.score = corner0; e.k = n; e.l = m; break;
        case M_LOCAL: e.score = best; e.k = bi; e.l = bj; break;
        case M_FITTING: e.score = cbest0; e.k = ci0; e.l = m; break;
        case M_OVERLAP: e.score = rbest; e.k = n; e.l = rj; break;
        default:  // M_SEMIGLOBAL, aligner.rs:389
            if (cbest0 > rbest) { e.score = cbest0; e.k = ci0; e.l = m; e.flags = 1; }
            else { e.score = rbest; e.k = n; e.l = rj; }
            break;
        }
        A.end[slot] = e;
    }
}

}  // namespace bg
