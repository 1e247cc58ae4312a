// k2f_fine.cuh -- K2f: fine-grained wavefront fill for a launch with ONE (or very few) long pairs.
//
// K2 (k2_wave.cuh) gives every warp a band of 512 columns: right for a launch of dozens of 50-100 kbp pairs, but the
// reference's own example (examples/from_file.rs:14-32: one 9 559 x 8 457 pair) then runs on 17 warps of 2 SMs
// (7.2 ms, 0.7 % of the integer roofline) while 146 SMs idle.  Here a lane owns ONE column:
//   * column j belongs to lane (j - 1) % 32 of warp (j - 1) / 32: 265 warps for 8 457 columns, W per CTA;
//   * inside a warp the K1 systolic schedule (lane l computes row t - l at step t; M + a, Y and the row residue move
//     one lane per step with __shfl_up_sync);
//   * warp w + 1 consumes the last column of warp w through a ring -- in SHARED memory when both sit in the same CTA, in
//     global memory across CTAs.  A ring entry is ONE 64-bit word: M + a and Y in 30 bits each and a 4-bit generation
//     tag, so the entry is its own flag: no counters to publish, no fences (the first version published a row counter
//     behind __threadfence_block() / st.release every 4 rows: MEMBAR.SC waits for the warp's outstanding trace stores,
//     and a lone pair filled in 15 ms instead of K2's 6).  The consumer's lanes 0..7 poll 8 rows at a time, so
//     adjacent warps run ~40 rows apart instead of K2's 64;
//   * total time ~ (len1 + 38 * warps) steps of ONE cell each instead of (len1 + 64 * 17) steps of 16 cells;
//   * the direction codes leave in K2's trace layout (one word = 8 consecutive columns of one row, bg_common.cuh), so
//     the long-pair walker runs unchanged: every lane keeps its last 15 nibbles in a 64-bit shift register, and every
//     8 steps the 8 lanes of a column block transpose an 8 x 8 nibble tile with three butterfly shuffles and each
//     stores one finished word;
//   * end-cell candidates are merged by the last warp to finish, exactly as in K2 (SURVEY A.5 tie rules).
// The launch is cooperative (spin waits need co-residency) and walks the launch's pairs one after the other with a
// grid-wide barrier in between; bg_api.cu uses it for K2-class launches of at most bg_set_fine_pairs() pairs.
//
// STATUS (measured on B200, cfg1 = the 9 559 x 8 457 fixture pair): bit-exact in every mode (tests force it onto
// multi-pair launches and compare with K2 and the oracle), but NOT faster than K2 yet: fill 7.8 ms (4-8 warps per CTA)
// against K2's 6.5 ms.  Any wavefront needs ~len1 + len2 steps; what counts is the latency of one step, and a
// one-cell step here is ~150 instructions (hand-over, seven shuffles, trace history, transposes) issued by a warp
// that has nothing to hide its latencies behind: ~700 cycles per step against ~80 cycles per CELL in K2's 16-cell
// steps.  It is therefore opt-in (bg_set_fine_pairs / BG_FINE_PAIRS, default 0); getting a lone pair under 1 ms needs
// a step of a few dozen instructions (CTA-wide lock step instead of per-warp hand-over), which is future work.
#pragma once
#include <cooperative_groups.h>

#include "bg_args.cuh"
#include "k1_fill.cuh"

namespace bg {

__device__ __forceinline__ void st_release_gpu_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ring entry: bits 0-29 M + a, 30-59 Y (both as 30-bit two's complement: |values| < 2^29 by bg_api.cu's range check),
// 60-63 generation tag 1..15 of the row (0 = never written)
__device__ __forceinline__ unsigned long long fine_pack(int32_t M, int32_t Y, uint32_t row, uint32_t ring_rows) {
    const uint32_t tag = 1u + (row / ring_rows) % 15u;
    return (unsigned long long)((uint32_t)M & 0x3FFFFFFFu) | ((unsigned long long)((uint32_t)Y & 0x3FFFFFFFu) << 30) | ((unsigned long long)tag << 60);
}
__device__ __forceinline__ uint32_t fine_tag(uint32_t row, uint32_t ring_rows) { return 1u + (row / ring_rows) % 15u; }
__device__ __forceinline__ int32_t fine_sx30(uint32_t v) { return (int32_t)(v << 2) >> 2; }

template <bool IS_LOCAL, bool PROF4>
__global__ void __launch_bounds__(1024, 1) k2f_fine(const FineArgs W) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t D = FINE_RING, DG = FINE_RING_G;
    const FillArgs& A = W.f;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t* s_row = smem_raw;
    uint8_t* s_col = smem_raw + 256;
    int32_t* s_tab = reinterpret_cast<int32_t*>(smem_raw + 512);
    __shared__ unsigned long long s_ring[32][FINE_RING];
    __shared__ uint32_t s_cons[32];
    const int ncol1 = A.n_cols + 1;
    for (int x = threadIdx.x; x < 256; x += blockDim.x) { s_row[x] = A.row_code[x]; s_col[x] = A.col_code[x]; }
    for (int x = threadIdx.x; x < A.n_rows * ncol1; x += blockDim.x) {
        const int r = x / ncol1, c = x - r * ncol1;
        s_tab[x] = ((c < A.n_cols) ? A.table[r * A.n_cols + c] : 0) - A.a;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t gw = blockIdx.x * nw + w;            // this warp's band (32 columns)
    const uint32_t l0 = (uint32_t)lane & ~7u, q8 = (uint32_t)lane & 7u;
    bool bad_residue = false;

    for (uint32_t slot = 0; slot < A.n_slots; ++slot) {
        const PairDesc d = A.desc[slot];
        if (d.pair_id == 0xFFFFFFFFu) continue;          // (uniform over the grid)
        // hand-over state of the previous pair is dead: reset, then everybody starts together
        for (uint32_t x = threadIdx.x; x < 32u * D; x += blockDim.x) (&s_ring[0][0])[x] = 0ull;
        for (uint32_t x = threadIdx.x; x < DG; x += blockDim.x) W.ring_g[(uint64_t)blockIdx.x * DG + x] = 0ull;
        if (threadIdx.x < 32) s_cons[threadIdx.x] = 0;
        if (threadIdx.x == 0) W.cons_g[blockIdx.x] = 0;
        __threadfence();
        grid.sync();

        const uint32_t n = d.n, m = d.m;
        const uint32_t B = (m + 31u) / 32u;               // warps that have columns
        const int32_t a = A.a, b = A.b, one = A.one;
        const uint32_t k32 = (uint32_t)A.one << 5;
        const int mode = A.mode;
        const bool row_gap = (mode == M_GLOBAL || mode == M_FITTING);
        const bool col_gap = (mode == M_GLOBAL);
        const bool track_col = (mode == M_SEMIGLOBAL || mode == M_FITTING);
        const bool track_row = (mode == M_SEMIGLOBAL || mode == M_OVERLAP);
        const uint8_t* sa = A.residues + d.a_off;
        const uint8_t* sb = A.residues + d.b_off;

        if (gw < B || (B == 0 && gw == 0)) {
            const uint32_t j0 = gw * 32u + (uint32_t)lane;    // 0-based column of this lane
            const bool has_colm = j0 < m;
            const bool col_lane = (m > 0) && (j0 == m - 1);
            int32_t best = 0; uint32_t bi = 0, bj = 0;
            int32_t rbest = INT32_MIN; uint32_t rj = 0;
            int32_t cbest = border_row(row_gap, a, b, m); uint32_t ci = 0;
            int32_t corner = border_col(col_gap, a, b, n);
            uint32_t has_col = 0;
            if (gw == 0 && lane == 0) { rbest = border_col(col_gap, a, b, n); rj = 0; }   // row n, column 0 candidate

            uint32_t cprof; int32_t MuA, Xu = NEG_INF;
            {
                uint32_t code = (uint32_t)A.n_cols;
                if (has_colm) { code = s_col[sb[j0]]; if (code == 0xFFu) { bad_residue = true; code = 0; } }
                if (PROF4) {
                    uint32_t pk = 0;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int32_t sv = (r < A.n_rows) ? s_tab[r * ncol1 + code] : -a;
                        pk |= ((uint32_t)sv & 0xffu) << (8 * r);
                    }
                    cprof = pk;
                } else cprof = code * 4u;
                MuA = border_row(row_gap, a, b, j0 + 1) + a;
            }
            int32_t MdiagA = border_row(row_gap, a, b, j0) + a;
            int32_t MlastA = a, Ylast = NEG_INF;
            uint32_t rcur = 0, cur_blk = 0;
            unsigned long long hist = 0ull;                  // nibble of row (current - k) at bits [4k, 4k + 4)

            // where the boundary column comes from / goes to (generic pointers: shared or global memory)
            const bool has_in = gw > 0, has_next = (gw + 1 < B);
            const bool in_glob = (w == 0), out_glob = (w + 1 == nw);
            const uint32_t Din = in_glob ? DG : D, Dout = out_glob ? DG : D;
            const volatile unsigned long long* ring_in = in_glob ? W.ring_g + (uint64_t)(blockIdx.x ? blockIdx.x - 1 : 0) * DG : &s_ring[w ? w - 1 : 0][0];
            volatile unsigned long long* ring_out = out_glob ? W.ring_g + (uint64_t)blockIdx.x * DG : &s_ring[w][0];
            volatile uint32_t* cons_in = in_glob ? W.cons_g + (blockIdx.x ? blockIdx.x - 1 : 0) : &s_cons[w ? w - 1 : 0];
            const volatile uint32_t* cons_out = out_glob ? W.cons_g + blockIdx.x : &s_cons[w];
            int32_t blkM = a, blkY = NEG_INF;                 // lanes 0..7: boundary rows [t & ~7, +8) of the left neighbour

            // steps: rows stream for n + 31 steps; the trace tiles need up to 14 more (a tile of rows [R, R + 8) is complete
            // for all 8 lanes of a column block at step R + l0 + 14), rounded so that the last flush step (t % 8 == 6) is met
            const uint32_t t_end = (A.want_trace && n) ? (((n - 1u) & ~7u) + 24u + 14u + 1u) : (n + 31u);
            const uint32_t steps = max(n + 31u, t_end);
            for (uint32_t t = 0; t < steps; ++t) {
                const uint32_t tq = t & 31u;
                if (tq == 0) {
                    const uint32_t idx = t + (uint32_t)lane;
                    uint32_t cd = 0;
                    if (idx < n) { cd = s_row[sa[idx]]; if (cd == 0xFFu) { bad_residue = true; cd = 0; } }
                    cur_blk = cd;
                }
                // ---- hand-over in: every 8 steps lanes 0..7 fetch rows [t, t + 8) of the left neighbour's last column ----
                if (has_in && (t & 7u) == 0 && t < n) {
                    const uint32_t row = t + (uint32_t)lane;
                    if (lane < 8 && row < n) {
                        const uint32_t want = fine_tag(row, Din);
                        unsigned long long v;
                        do { v = ring_in[row & (Din - 1u)]; } while ((uint32_t)(v >> 60) != want);
                        blkM = fine_sx30((uint32_t)v & 0x3FFFFFFFu); blkY = fine_sx30((uint32_t)(v >> 30) & 0x3FFFFFFFu);
                    }
                    __syncwarp();
                    if (lane == 0) *cons_in = min(t + 8u, n);  // rows below are in registers now: their ring slots may be reused
                }
                const int32_t inM = __shfl_sync(FULL, blkM, (int)(t & 7u));
                const int32_t inY = __shfl_sync(FULL, blkY, (int)(t & 7u));
                const uint32_t r0 = __shfl_sync(FULL, cur_blk, (int)tq);
                int32_t MlA = __shfl_up_sync(FULL, MlastA, 1);
                int32_t Yl = __shfl_up_sync(FULL, Ylast, 1);
                uint32_t r = __shfl_up_sync(FULL, rcur, 1);
                const uint32_t i0 = t - (uint32_t)lane;
                const bool active = i0 < n;
                if (lane == 0) {
                    r = r0;
                    if (gw == 0) { MlA = border_col(col_gap, a, b, i0 + 1) + a; Yl = NEG_INF; }
                    else { MlA = inM; Yl = inY; }
                }
                rcur = r;
                uint32_t nib = 0;
                if (active) {
                    int32_t sb_;
                    if (PROF4) sb_ = prmt_sx(cprof, r * 0x1111u + 0x8880u);
                    else sb_ = *reinterpret_cast<const int32_t*>(reinterpret_cast<const unsigned char*>(s_tab) + r * (uint32_t)(ncol1 * 4) + cprof);
                    const int32_t dg = fma_add(MdiagA, one, sb_);
                    const int32_t upA = MuA;
                    int32_t Y = Yl;
                    if (IS_LOCAL) {
                        int32_t mx; uint32_t rowkey = 0;
                        local_cell(nib, 0u, upA, Xu, Y, MlA, dg, b, one, k32, 31u, rowkey, mx);
                        MuA = fma_add(mx, one, a);
                        if (mx > best) { best = mx; bi = i0 + 1; bj = j0 + 1; }
                    } else {
                        const int32_t X = __viaddmax_s32(Xu, b, upA);
                        acc_if_eq(nib, X, upA, one, TR_XOPEN);
                        Y = __viaddmax_s32(Y, b, MlA);
                        acc_if_eq(nib, Y, MlA, one, TR_YOPEN);
                        const int32_t mx = __vimax3_s32(dg, X, Y);
                        acc_if_eq(nib, mx, Y, one, TR_YEQ);
                        acc_if_eq(nib, mx, X, one, TR_XEQ);
                        MuA = fma_add(mx, one, a);
                        Xu = X;
                    }
                    MlastA = MuA; Ylast = Y; MdiagA = MlA;
                    if (track_col && col_lane) { const int32_t v = MuA - a; if (v > cbest) { cbest = v; ci = i0 + 1; } }
                }
                // ---- hand-over out: lane 31 finished row t - 31 ----
                const uint32_t i31 = t - 31u;
                if (has_next && i31 < n) {
                    if ((i31 & 7u) == 0) {                    // room in the ring for rows [i31, i31 + 8)?
                        if (lane == 31) { while (*cons_out + Dout < i31 + 8u) { } }
                        __syncwarp();
                    }
                    if (lane == 31) ring_out[i31 & (Dout - 1u)] = fine_pack(MlastA, Ylast, i31, Dout);
                }
                // ---- direction codes: 15-step history per lane, an 8 x 8 nibble transpose per column block every 8 steps ----
                if (A.want_trace) {
                    hist = (hist << 4) | (unsigned long long)nib;
                    if ((t & 7u) == 6u && t >= 14u) {           // (warp-uniform: the shuffles below need all lanes)
                        const bool tile_ok = t >= l0 + 14u;       // blocks further right complete their first tile later
                        const uint32_t R0 = t - l0 - 14u;         // rows [R0, R0 + 8) are complete in all 8 lanes of the block
                        // this lane's nibbles of those rows: row R0 + r at nibble 7 - r
                        uint32_t x = (uint32_t)(hist >> (4u * (7u - q8)));
                        // transpose across the 8 lanes (lane q8 holds column q8 -> lane q8 holds row R0 + 7 - q8, column c at nibble c)
                        uint32_t y = __shfl_xor_sync(FULL, x, 4);
                        x = (q8 & 4u) ? ((x & 0xFFFF0000u) | (y >> 16)) : ((x & 0x0000FFFFu) | (y << 16));
                        y = __shfl_xor_sync(FULL, x, 2);
                        x = (q8 & 2u) ? ((x & 0xFF00FF00u) | ((y >> 8) & 0x00FF00FFu)) : ((x & 0x00FF00FFu) | ((y << 8) & 0xFF00FF00u));
                        y = __shfl_xor_sync(FULL, x, 1);
                        x = (q8 & 1u) ? ((x & 0xF0F0F0F0u) | ((y >> 4) & 0x0F0F0F0Fu)) : ((x & 0x0F0F0F0Fu) | ((y << 4) & 0xF0F0F0F0u));
                        const uint32_t row0 = R0 + 7u - q8;        // 0-based DP row of this lane's word
                        const uint32_t jb = gw * 32u + l0;         // 0-based first column of the block
                        if (tile_ok && row0 < n && jb < m) {
                            // K2 layout (L = 32 lanes x 16 columns): word (band, t_std, k, lane16)
                            const uint32_t band = jb >> 9, lane16 = (jb & 511u) >> 4, k16 = (jb & 15u) >> 3;
                            const uint64_t idx = d.trace_off + (((uint64_t)band * d.steps + (row0 + lane16)) * 2u + k16) * 32u + lane16;
                            A.trace[idx] = x;
                        }
                    }
                }
            }
            // row n of this lane's column
            if (has_colm) {
                const int32_t v = MuA - a;
                if (track_row && v >= rbest) { rbest = v; rj = j0 + 1; }
                if (col_lane) { corner = v; has_col = 1; }
            }
            // ---- merge inside the warp, then across the warps of the pair (k2_wave.cuh's rules) ----
            if (track_row) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int32_t ov = __shfl_xor_sync(FULL, rbest, o);
                    const uint32_t oj = __shfl_xor_sync(FULL, rj, o);
                    if (ov > rbest || (ov == rbest && oj > rj)) { rbest = ov; rj = oj; }
                }
            }
            if (IS_LOCAL) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int32_t ov = __shfl_xor_sync(FULL, best, o);
                    const uint32_t oi = __shfl_xor_sync(FULL, bi, o);
                    const uint32_t oj = __shfl_xor_sync(FULL, bj, o);
                    if (ov > best || (ov == best && (oi < bi || (oi == bi && oj < bj)))) { best = ov; bi = oi; bj = oj; }
                }
            }
            const uint32_t p_m = m ? (m - 1u) & 31u : 0u;
            const int32_t cbest0 = __shfl_sync(FULL, cbest, (int)p_m);
            const uint32_t ci0 = __shfl_sync(FULL, ci, (int)p_m);
            const int32_t corner0 = __shfl_sync(FULL, corner, (int)p_m);
            const uint32_t hascol0 = __shfl_sync(FULL, has_col, (int)p_m);
            const uint32_t NW = max(B, 1u);
            bool merger = false;
            if (lane == 0) {
                WaveCand c;
                c.best = best; c.bi = bi; c.bj = bj; c.rbest = rbest; c.rj = rj;
                c.cbest = cbest0; c.ci = ci0; c.corner = corner0; c.has_col = hascol0;
                W.cand[(uint64_t)slot * W.cand_stride + gw] = c;
                __threadfence();
                merger = (atomicAdd(W.done + slot, 1u) == NW - 1);
                if (merger) __threadfence();
            }
            if (merger) {
                const WaveCand* cc = W.cand + (uint64_t)slot * W.cand_stride;
                int32_t fbest = 0; uint32_t fbi = 0, fbj = 0;
                int32_t frb = INT32_MIN; uint32_t frj = 0;
                int32_t fcb = border_row(row_gap, a, b, m); uint32_t fci = 0;
                int32_t fcorner = border_col(col_gap, a, b, n);
                for (uint32_t k = 0; k < NW; ++k) {
                    const volatile WaveCand* c = cc + k;
                    const int32_t vb = c->best; const uint32_t vbi = c->bi, vbj = c->bj;
                    if (vb > fbest || (vb == fbest && (vbi < fbi || (vbi == fbi && vbj < fbj)))) { fbest = vb; fbi = vbi; fbj = vbj; }
                    const int32_t vr = c->rbest; const uint32_t vrj = c->rj;
                    if (vr > frb || (vr == frb && vrj > frj)) { frb = vr; frj = vrj; }
                    if (c->has_col) { fcb = c->cbest; fci = c->ci; fcorner = c->corner; }
                }
                EndCell e; e.flags = 0;
                switch (mode) {
                case M_GLOBAL: e.score = fcorner; e.k = n; e.l = m; break;
                case M_LOCAL: e.score = fbest; e.k = fbi; e.l = fbj; break;
                case M_FITTING: e.score = fcb; e.k = fci; e.l = m; break;
                case M_OVERLAP: e.score = frb; e.k = n; e.l = frj; break;
                default:
                    if (fcb > frb) { e.score = fcb; e.k = fci; e.l = m; e.flags = 1; }
                    else { e.score = frb; e.k = n; e.l = frj; }
                    break;
                }
                A.end[slot] = e;
            }
        }
        grid.sync();      // the next pair reuses the rings and counters
    }
    if (bad_residue) atomicOr(A.err_flag, 1u);
}

}  // namespace bg
