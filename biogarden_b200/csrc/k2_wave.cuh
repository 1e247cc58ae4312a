// k2_wave.cuh -- K2: intra-sequence wavefront fill for long pairs (>= several kbp) on sm_100a.
//
// Same recurrence, same cell code and same trace layout as K1 (k1_fill.cuh), but ONE pair is spread
// over all warps of a GROUP of Q co-resident CTAs instead of living in one warp:
//   * the columns are cut into bands of 32*C; the group's workers -- a worker being one warp of one CTA
//     of the group, NW = Q * 16 of them -- claim bands in order from a per-pair counter;
//   * inside a band the 32 lanes run the K1 systolic schedule over all rows;
//   * band b+1 consumes the last column (M + a, Y) of band b: the producer's lane 31 values are
//     collected across 32 steps with shuffles and written as one coalesced 256-byte block to a ring
//     of NW+1 boundary columns in global memory, then a per-boundary progress counter is bumped;
//     the consumer waits on that counter once per 32 rows and fetches the block with one coalesced
//     load.  Adjacent bands therefore run 64 rows apart: the anti-diagonal wavefront of the north
//     star, at band granularity, with the diagonal tiles staged through registers / L2;
//   * the launch is cooperative, so all workers are co-resident and the spin waits cannot deadlock;
//     the last worker to finish a pair merges the workers' end-cell candidates;
//   * a launch with fewer bands than 16 warps per SM (a lone 10 kbp pair has 17) uses fewer warps per CTA, down to
//     one, so that the bands spread over the SMs: 16 warps on one SM share its four schedulers (19 instructions per
//     cell x 4 warps per scheduler = ~80 cycles per cell and warp), a warp alone on its scheduler runs at the
//     latency of the dependent chain;
//   * the grid is persistent (one CTA per SM) and every CTA walks its own list of (pair, group rank,
//     group size) assignments built by the host (WaveAssign): groups are sized in proportion to the
//     pairs' cell counts so that the pairs of a launch finish together, and a CTA moves on to its next
//     pair as soon as its own bands are done -- there is no barrier between pairs.
// The walk (K3) reads the trace exactly as for K1 with L = 32.
//
// Bounded-memory traceback (CKPT = true; pairs whose 0.5 B/cell trace does not fit the budget): the DP is run
// twice.  Pass 1 keeps no direction codes; it leaves row CHECKPOINTS -- the register state (M + a, X) of every
// column at the rows that are multiples of `ckpt_every` -- and the end cell.  Pass 2 goes through the row
// blocks bottom-up: one launch re-fills rows (row0, row0 + nrows] starting from the checkpoint at row0, this
// time storing the codes of just that block, and k3_walk_diag continues the walk through the block
// (WalkState carries it from launch to launch).  Trace memory is ckpt_every x len2 / 2 bytes instead of
// len1 x len2 / 2, checkpoints cost 8 B x len2 per block, and the cells are computed twice.
#pragma once
#include <cooperative_groups.h>

#include "bg_args.cuh"
#include "k1_fill.cuh"

namespace bg {






// Boundary hand-over flags: release / acquire at GPU scope on the counter itself.  The producer's 32
// lanes write their block, __syncwarp() orders those writes before lane 0's release store; the consumer's
// lane 0 acquires, __syncwarp() passes the ordering on to the other lanes.  (A full __threadfence() on
// all lanes would also wait for every lane's outstanding trace stores.)
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int C, bool IS_LOCAL, bool PROF4, bool CKPT = false>
__global__ void __launch_bounds__(K2_WARPS * 32, 1) k2_wave(const WaveArgs W) {
    namespace cg = cooperative_groups;
    constexpr int L = 32;
    constexpr int K = (C + 7) / 8;
    constexpr unsigned FULL = 0xffffffffu;
    const FillArgs& A = W.f;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t* s_row = smem_raw;
    uint8_t* s_col = smem_raw + 256;
    int32_t* s_tab = reinterpret_cast<int32_t*>(smem_raw + 512);
    const int ncol1 = A.n_cols + 1;
    for (int x = threadIdx.x; x < 256; x += blockDim.x) { s_row[x] = A.row_code[x]; s_col[x] = A.col_code[x]; }
    for (int x = threadIdx.x; x < A.n_rows * ncol1; x += blockDim.x) {
        const int r = x / ncol1, c = x - r * ncol1;
        s_tab[x] = ((c < A.n_cols) ? A.table[r * A.n_cols + c] : 0) - A.a;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, p = lane;
    bool bad_residue = false;

  for (uint32_t rd = 0; rd < W.n_rounds; ++rd) {
    const WaveAssign as = W.assign[(uint64_t)rd * gridDim.x + blockIdx.x];
    if (as.Q == 0) continue;
    const uint32_t slot = as.slot;
    const uint32_t wpc = blockDim.x >> 5;   // warps per CTA: K2_WARPS, fewer when the launch has fewer bands than the machine has warps
    const uint32_t Q = as.Q, NW = Q * wpc, R = NW + 1;
    const uint32_t wk = (uint32_t)as.rank * wpc + (threadIdx.x >> 5);
    const PairDesc d = A.desc[slot];
    CkptSlot cs; cs.ck_off = 0; cs.ck_stride = cs.row0 = cs.nrows = 0; cs.every = 1;
    if (CKPT) cs = W.cks[slot];
    const uint32_t row0 = cs.row0;
    // pass 2: nothing to do for empty blocks, blocks below the end cell, or after the walk has ended
    if (CKPT && !W.write_end && (cs.nrows == 0 || W.wstate[slot].done || A.end[slot].k <= row0)) continue;
    const uint32_t n = CKPT ? cs.nrows : d.n, m = d.m, nbands = d.nbands;
    const uint32_t tstride = d.steps;                       // steps per band in the trace layout
    const uint32_t steps = CKPT ? n + 31u : d.steps;        // steps this launch runs
    const uint32_t n_pad = (n + 31u) & ~31u;
    const int32_t a = A.a, b = A.b, one = A.one;
    const uint32_t k32 = (uint32_t)A.one << 5;
    const int mode = A.mode;
    const bool row_gap = (mode == M_GLOBAL || mode == M_FITTING);
    const bool col_gap = (mode == M_GLOBAL);
    const bool track_col = (mode == M_SEMIGLOBAL || mode == M_FITTING);
    const bool track_row = (mode == M_SEMIGLOBAL || mode == M_OVERLAP);
    const uint8_t* sa = A.residues + d.a_off + row0;
    const uint8_t* sb = A.residues + d.b_off;
    const int2* ck_in = nullptr;
    if (CKPT && row0 > 0) ck_in = W.ckpt + cs.ck_off + (uint64_t)(row0 / cs.every - 1u) * cs.ck_stride;

    const uint32_t band_cols = (uint32_t)(L * C);
    const uint32_t mcol0 = m ? m - 1 : 0;
    const uint32_t bd_m = mcol0 / band_cols;
    const uint32_t p_m = (mcol0 % band_cols) / C;
    const uint32_t c_m = mcol0 % C;
    const bool col_lane = (m > 0) && ((uint32_t)p == p_m);

    int32_t best = 0; uint32_t bi = 0, bj = 0;
    int32_t rbest = INT32_MIN; uint32_t rj = 0;
    int32_t cbest = border_row(row_gap, a, b, m); uint32_t ci = 0;
    int32_t corner = border_col(col_gap, a, b, n);
    uint32_t has_col = 0;
    if (wk == 0 && p == 0) { rbest = border_col(col_gap, a, b, n); rj = 0; }   // row n, column 0 candidate

    int2* const ring = A.bnd + d.bnd_off;
    unsigned long long* const prog = W.progress + (uint64_t)slot * W.prog_stride;

    // Bands are claimed dynamically (band b is always claimed before band b + 1, by a resident warp, and
    // depends on band b - 1 only): a static round-robin leaves most workers idle in a pair's last,
    // partly filled round (196 bands on 64 workers: 4 rounds for 3.06 rounds of work).
    for (;;) {
        uint32_t bd = 0;
        if (lane == 0) bd = atomicAdd(W.next_band + slot, 1u);
        bd = __shfl_sync(FULL, bd, 0);
        if (bd >= nbands) break;
        const uint32_t jbase = bd * band_cols + (uint32_t)p * C;
        const bool lane_has_cols = jbase < m;
        const bool has_next = (bd + 1 < nbands);
        const int2* bnd_rd = ring + (uint64_t)((bd + R - 1) % R) * n_pad;
        int2* bnd_wr = ring + (uint64_t)(bd % R) * n_pad;
        const unsigned long long rd_base = (unsigned long long)(bd - 1) * (n + 1ull);   // unused for bd == 0
        const unsigned long long wr_base = (unsigned long long)bd * (n + 1ull);

        uint32_t cprof[C];
        int32_t MuA[C], Xu[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const uint32_t j0 = jbase + c;
            uint32_t code = (uint32_t)A.n_cols;
            if (j0 < m) {
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
            if (CKPT && ck_in && j0 < m) { const int2 v = __ldcg(ck_in + j0); MuA[c] = v.x; Xu[c] = v.y; }
        }
        int32_t MdiagA = border_row(row_gap, a, b, jbase) + a;
        if (CKPT && ck_in) MdiagA = (jbase == 0) ? border_col(col_gap, a, b, row0) + a : ((jbase - 1 < m) ? __ldcg(ck_in + jbase - 1).x : a);
        int32_t MlastA = a, Ylast = NEG_INF;
        uint32_t rcur = 0;
        uint32_t cur_blk = 0;
        int2 in_blk = make_int2(a, NEG_INF);    // boundary rows [t0, t0+32) of the band to the left
        uint32_t next_ck = cs.every, ck_no = 0;  // CKPT pass 1: next checkpoint row of this lane, its index

        for (uint32_t t = 0; t < steps; ++t) {
            const uint32_t tq = t & 31u;
            if (tq == 0) {
                // row residues of the next 32 rows
                {
                    const uint32_t idx = t + (uint32_t)p;
                    uint32_t cd = 0;
                    if (idx < n) { cd = s_row[sa[idx]]; if (cd == 0xFFu) { bad_residue = true; cd = 0; } }
                    cur_blk = cd;
                }
                if (bd > 0) {
                    // wait until the left band has published rows [t, t+32)
                    const unsigned long long need = rd_base + (unsigned long long)min(t + 32u, n);
                    if (lane == 0 && t < n) {
                        unsigned ns = 32;
                        while (ld_acquire_gpu(prog + (bd + R - 1) % R) < need) { __nanosleep(ns); if (ns < 512) ns *= 2; }
                    }
                    __syncwarp();
                    const uint32_t row = t + (uint32_t)lane;
                    in_blk = (row < n) ? __ldcg(bnd_rd + row) : make_int2(a, NEG_INF);
                }
            }
            const uint32_t r0 = __shfl_sync(FULL, cur_blk, (int)tq);
            const int32_t inM = __shfl_sync(FULL, in_blk.x, (int)tq);
            const int32_t inY = __shfl_sync(FULL, in_blk.y, (int)tq);
            int32_t MlA = __shfl_up_sync(FULL, MlastA, 1);
            int32_t Yl = __shfl_up_sync(FULL, Ylast, 1);
            uint32_t r = __shfl_up_sync(FULL, rcur, 1);
            const uint32_t i0 = t - (uint32_t)p;
            const bool active = i0 < n;
            {
                // lane 0 takes the row residue from the block and the left border / the neighbour band's column instead of a
                // left neighbour -- as selects: a branch here made the warp run lane 0's arm on its own every step
                const bool first = (p == 0);
                const int32_t bM = (bd == 0) ? border_col(col_gap, a, b, row0 + i0 + 1) + a : inM;   // bd is uniform in the warp
                const int32_t bY = (bd == 0) ? NEG_INF : inY;
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
                    uint32_t& wk_ = w[c >> 3];
                    const uint32_t sh = 4u * (c & 7);
                    const int32_t upA = MuA[c];
                    if (IS_LOCAL) {
                        int32_t mx;
                        local_cell(wk_, sh, upA, Xu[c], Y, leftA, dg[c], b, one, k32, 31u - c, rowkey, mx);
                        leftA = fma_add(mx, one, a);
                        MuA[c] = leftA;
                        continue;
                    }
                    // aligner.rs:443-444: xo = M[i-1][j] + a is the register itself
                    const int32_t X = __viaddmax_s32(Xu[c], b, upA);
                    acc_if_eq(wk_, X, upA, one, TR_XOPEN << sh);
                    // aligner.rs:447-448: yo = M[i][j-1] + a likewise
                    Y = __viaddmax_s32(Y, b, leftA);
                    acc_if_eq(wk_, Y, leftA, one, TR_YOPEN << sh);
                    // aligner.rs:451-466
                    const int32_t mx = __vimax3_s32(dg[c], X, Y);
                    acc_if_eq(wk_, mx, Y, one, TR_YEQ << sh);
                    acc_if_eq(wk_, mx, X, one, TR_XEQ << sh);
                    leftA = fma_add(mx, one, a);
                    MuA[c] = leftA; Xu[c] = X;
                }
                MlastA = leftA; Ylast = Y; MdiagA = MlA;
                if (IS_LOCAL) {
                    const int32_t v = (int32_t)(rowkey >> 5);
                    if (v > best) { best = v; bi = row0 + i0 + 1; bj = jbase + 32u - (rowkey & 31u); }
                }
                if (A.want_trace && lane_has_cols) {
                    uint32_t* tp = A.trace + d.trace_off + ((uint64_t)bd * tstride + t) * (uint64_t)(K * 32) + lane;
#pragma unroll
                    for (int k = 0; k < K; ++k) tp[k * 32] = w[k];
                }
                if (CKPT && W.ckpt_write && i0 + 1 == next_ck) {
                    if (next_ck < n) {
                        int2* ck = W.ckpt + cs.ck_off + (uint64_t)(ck_no++) * cs.ck_stride + jbase;
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            if (jbase + c < m) __stcg(ck + c, make_int2(MuA[c], Xu[c]));
                    }
                    next_ck += cs.every;
                }
                if (track_col && bd == bd_m) {
                    int32_t v = MuA[0];
#pragma unroll
                    for (int c = 1; c < C; ++c) v = (c_m == (uint32_t)c) ? MuA[c] : v;
                    v -= a;
                    if (col_lane && v > cbest) { cbest = v; ci = i0 + 1; }
                }
            }
            if (has_next) {
                // lane 31 finishes row t - 31 in this step: its (M + a, Y) go straight to the ring (8 bytes per step; collecting 32
                // rows in the lanes for one coalesced store cost two shuffles and two selects per step); the counter is
                // published every 32 steps
                if (lane == 31 && active) bnd_wr[i0] = make_int2(MlastA, Ylast);
                if (tq == 31u || t + 1 == steps) {
                    __syncwarp();
                    if (lane == 0) {
                        long long done = (long long)t - 30;             // rows 0 .. t-31 are out
                        if (done < 0) done = 0;
                        if (done > (long long)n) done = (long long)n;
                        st_release_gpu(prog + bd % R, wr_base + (unsigned long long)done);
                    }
                }
            }
        }
        // MuA[] holds row n of this band (biased by a)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const uint32_t j = jbase + c + 1;
            const int32_t v = MuA[c] - a;
            if (j <= m) {
                if (track_row && v >= rbest) { rbest = v; rj = j; }
                if (j == m) { corner = v; }
            }
        }
        if (bd == bd_m && m > 0) has_col = 1;
        __syncwarp();
    }

    // ---- merge inside the warp, then across the workers of the group --------------------
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
    const int32_t cbest0 = __shfl_sync(FULL, cbest, (int)p_m);
    const uint32_t ci0 = __shfl_sync(FULL, ci, (int)p_m);
    const int32_t corner0 = __shfl_sync(FULL, corner, (int)p_m);
    bool merger = false;
    if (lane == 0) {
        WaveCand c;
        c.best = best; c.bi = bi; c.bj = bj; c.rbest = rbest; c.rj = rj;
        c.cbest = cbest0; c.ci = ci0; c.corner = corner0; c.has_col = has_col;
        W.cand[(uint64_t)slot * W.cand_stride + wk] = c;
        __threadfence();
        merger = (atomicAdd(W.done + slot, 1u) == NW - 1);   // last worker out merges
        if (merger) __threadfence();
    }
    if (merger) {
        const WaveCand* cc = W.cand + (uint64_t)slot * W.cand_stride;
        int32_t fbest = 0; uint32_t fbi = 0, fbj = 0;
        int32_t frb = INT32_MIN; uint32_t frj = 0;
        int32_t fcb = border_row(row_gap, a, b, m); uint32_t fci = 0;
        int32_t fcorner = border_col(col_gap, a, b, n);
        for (uint32_t q = 0; q < NW; ++q) {
            const volatile WaveCand* c = cc + q;
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
        if (!CKPT || W.write_end) A.end[slot] = e;
    }
  }   // persistent loop
    if (bad_residue) atomicOr(A.err_flag, 1u);
}

}  // namespace bg
