// k3_walk.cuh -- K3: on-device traceback walk + aligned-string assembly.
//
// Replaces SequenceAligner::backtrack (aligner.rs:511-592) and the semiglobal tail / prefix
// assembly (aligner.rs:382-432).  The walk is the reference's state machine verbatim
// (SURVEY A.4), including its late read of x_trace / y_trace: after an M->X move the walker
// consults the open/extend bit of the cell it ARRIVED at.  The emitted path is therefore
// defined by the stored direction codes alone and is reproduced bit for bit.
//
// One thread per pair: the walk is a serial chain of dependent 4-byte loads, so the way to
// keep the machine busy is many independent walks per SM, not lanes cooperating on one.
// Every walker records one 2-bit op per alignment column, back to front, into the pair's slot; from there either
// k_gather materialises the two strings on the device (device-resident results), or k_pack_ops packs the ops
// densely for the trip to the host, where host_expand.cpp turns them into strings (host-buffer entry points).
#pragma once
#include "bg_args.cuh"

namespace bg {

#ifndef K3_MINB
#define K3_MINB 16
#endif


// LT / CT != 0: the launch's geometry as compile-time constants (HALF: K1h's packed row-block layout).  The cell
// address is two divisions and a handful of multiplies by the geometry; with run-time operands that was most
// of the ~100 instructions a walk step cost (ncu: 29 k warp instructions per warp of 300-step walks, issue
// slots 57 % busy).  <0, 0, false> is the generic form.
template <int LT, int CT, bool HALF>
__global__ void __launch_bounds__(128, K3_MINB) k3_walk(const WalkArgs A) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= A.n_slots) return;
    const PairDesc d = A.desc[slot];
    if (d.pair_id == 0xFFFFFFFFu) return;
    const EndCell e = A.end[slot];
    const uint32_t n = d.n, m = d.m;
    const uint32_t L = LT ? (uint32_t)LT : (uint32_t)A.L, C = CT ? (uint32_t)CT : (uint32_t)A.C;
    const uint32_t K = (C + 7) / 8;
    const uint32_t band_cols = L * C;
    const uint32_t H = LT ? (HALF ? 2u : 1u) : (uint32_t)A.H;
    const uint32_t lane_base = ((slot % (H * (32u / L))) / H) * L;
    const uint32_t half = slot % H;
    // Output: the walk only records WHAT it did -- one 2-bit op per step (0 both residues, 1 seq1 residue
    // over a gap, 2 gap over seq2 residue), 16 per word, back to front -- plus where it ended.  k_gather turns
    // ops into the two aligned strings with coalesced residue reads; the walk itself is then a single chain
    // of dependent trace loads (fetching residues here cost two more memory round trips per step, ncu).
    // Slot: [first_a][first_b][ops words ...], capacity n + m ops (every op consumes a residue).
    uint32_t* slotw = reinterpret_cast<uint32_t*>(A.pad + d.pad_off);
    uint32_t* ops = slotw + 2;
    uint32_t pos = n + m;
    uint32_t wops = 0;
    const int mode = A.mode;
    auto push = [&](uint32_t op) {
        --pos;
        wops |= op << ((pos & 15u) * 2u);
        if ((pos & 15u) == 0) { ops[pos >> 4] = wops; wops = 0; }
    };
    const uint32_t CW = LT ? (HALF ? (uint32_t)((CT + 3) & ~3) : 0u) : (uint32_t)A.CW;
    uint64_t quad_idx = ~0ull; uint4 quad = make_uint4(0, 0, 0, 0);
    // INC (K1h layout with compile-time geometry, the short-read case): the cell's coordinates in the trace -- lane wp,
    // column wc of the lane, systolic step wt -- follow the walk incrementally, and the word index is kept as the sum of
    // a term that changes with the 4-step row block (every 4th step), one that changes with the lane (every C-th
    // column) and the column.  Recomputing the address from (k, l) every step (two divisions by the geometry, 64-bit
    // multiplies) was most of the ~160 thread instructions a step cost (ncu: issue slots 54 % busy in a kernel that
    // should only wait for memory).
    constexpr bool INC = (LT != 0) && HALF;
    uint32_t wc = 0, wt = 0, w_lane = 0, w_blk = 0, quad_i = 0xffffffffu;   // word offsets are relative to the pair's trace block (32 bits)
    const uint32_t tgs = (uint32_t)A.tg_shift;
    const uint4* const tquad = reinterpret_cast<const uint4*>(A.trace + (INC ? d.trace_off : 0));   // 16-byte aligned: the fill stores uint4 there
    auto set_blk = [&]() {
        const uint32_t tb = wt >> 2;
        w_blk = (((tb >> tgs) * (32u * CW)) << tgs) + (tb & ((1u << tgs) - 1u)) * CW;
    };
    auto nib_at = [&](uint32_t i, uint32_t j) -> uint32_t {   // i, j >= 1
        if (INC) {
            const uint32_t idx = w_lane + w_blk + wc;
            const uint32_t q = idx >> 2;
            if (q != quad_i) {   // the pair's half of every word is moved down once per quad, not once per step
                quad_i = q; quad = __ldg(tquad + q);
                const uint32_t h16 = (threadIdx.x & 1u) * 16u;    // == half * 16 (H == 2): not kept in a register across the loop
                quad.x >>= h16; quad.y >>= h16; quad.z >>= h16; quad.w >>= h16;
            }
            const uint32_t lo2 = (idx & 1u) ? quad.y : quad.x, hi2 = (idx & 1u) ? quad.w : quad.z;
            return (((idx & 2u) ? hi2 : lo2) >> ((wt & 3u) * 4u)) & 15u;
        }
        const uint32_t j0 = j - 1;
        if (CW) {   // K1h row blocks: a diagonal move goes to the previous word of the same 16-byte quad, which
                    // is kept in registers (2048 walks per SM touch 5 lines each: L1 cannot hold them, ncu 13 % hits)
            const uint32_t p = j0 / C, c = j0 - p * C;
            const uint32_t t = (i - 1) + p;
            const uint32_t tb = t >> 2;
            const uint64_t idx = d.trace_off + ((((uint64_t)(tb >> tgs) * 32u + lane_base + p) * CW) << tgs) + (tb & ((1u << tgs) - 1u)) * CW + c;
            const uint64_t q = idx >> 2;
            if (q != quad_idx) { quad_idx = q; quad = __ldg(reinterpret_cast<const uint4*>(A.trace) + q); }
            const uint32_t lo2 = (idx & 1u) ? quad.y : quad.x, hi2 = (idx & 1u) ? quad.w : quad.z;
            return (((idx & 2u) ? hi2 : lo2) >> ((t & 3u) * 4u + half * 16u)) & 15u;
        }
        const uint32_t bd = j0 / band_cols, rr = j0 - bd * band_cols;
        const uint32_t p = rr / C, c = rr - p * C;
        const uint32_t t = (i - 1) + p;
        const uint64_t idx = d.trace_off + (((uint64_t)(bd * d.steps + t) * K + (c >> 3)) * H + half) * 32u + lane_base + p;
        return (__ldg(A.trace + idx) >> ((c & 7u) * 4u)) & 15u;
    };

    uint32_t k = e.k, l = e.l, flags = 0;
    // one row up / one column left, with the trace coordinates following (they are only read while k, l >= 1)
    auto step_up = [&]() {
        --k;
        if (INC) { const uint32_t t0 = wt; --wt; if ((t0 & 3u) == 0u) set_blk(); }
    };
    auto step_left = [&]() {
        --l;
        if (INC) {
            if (wc) --wc;
            else { wc = C - 1u; w_lane -= CW << tgs; const uint32_t t0 = wt; --wt; if ((t0 & 3u) == 0u) set_blk(); }
        }
    };
    const bool colbr = (e.flags & 1u) != 0;
    if (mode == M_SEMIGLOBAL) {   // aligner.rs:389-404
        if (colbr) { for (uint32_t i = n; i > k; --i) push(1u); }
        else       { for (uint32_t i = m; i > l; --i) push(2u); }
    }
    if (INC && k != 0 && l != 0) {
        const uint32_t j0 = l - 1, wp = j0 / C;
        wc = j0 - wp * C; wt = (k - 1) + wp;
        w_lane = ((lane_base + wp) * CW) << tgs;
        set_blk();
    }
    uint32_t cur = 0;   // 0 = 'M', 1 = 'X', 2 = 'Y'
    const uint32_t bound = 2u * (n + m) + 8u;
    uint32_t it = 0;
    if (INC) {
        // Interior cells, branch-free: the 32 walks of a warp are in different states, and the state machine below, written
        // with branches, makes the warp run every arm in turn (ncu: 25 k warp instructions per warp of ~170-step walks,
        // issue slots 54 % busy).  Here every step is the same instruction stream: the transition as selects, the op push
        // and the coordinate updates predicated.  Borders (k == 0 or l == 0) and the end of the walk are left to the
        // general loop, which continues from whatever state this one stops in.
        const uint32_t lane_step = CW << tgs;
        // (K1h never runs local mode, so there is no stop test; every second iteration at least emits, so the loop ends
        //  after at most 2 (n + m) iterations without a hang counter)
        while (k != 0 && l != 0) {
            const uint32_t nib = nib_at(k, l);
            const bool s0 = cur == 0;
            const uint32_t op_m = (nib & TR_YEQ) ? 2u : (nib & TR_XEQ);           // aligner.rs:519-531: Y tested first
            const bool open = ((cur == 1 ? TR_XOPEN : TR_YOPEN) & nib) != 0;     // aligner.rs:541, 566: the gap was opened here
            const bool emit = s0 || !open;
            const uint32_t op = s0 ? op_m : cur;
            if (emit) push(op);
            const uint32_t dk = (emit && op != 2u) ? 1u : 0u, dl = (emit && op != 1u) ? 1u : 0u;
            const uint32_t wrap = (dl && wc == 0u) ? 1u : 0u;                     // leaving the lane: column C - 1 of the lane to the left, one step earlier
            k -= dk; l -= dl;
            wc = wrap ? C - 1u : wc - dl;
            w_lane -= wrap ? lane_step : 0u;
            wt -= dk + wrap;
            set_blk();
            cur = s0 ? op_m : (open ? 0u : cur);
        }
    }
    for (;; ++it) {
        if (it > bound) { flags |= WALK_HANG; break; }
        const bool interior = (k != 0 && l != 0);
        uint32_t nib = 0;
        if (interior) nib = nib_at(k, l);
        bool valid;
        switch (mode) {
            case M_GLOBAL: valid = (k != 0 || l != 0); break;                                   // aligner.rs:117
            case M_LOCAL: valid = interior && (nib & 3u) != 3u; break;                          // aligner.rs:181 (borders are 0)
            case M_SEMIGLOBAL: valid = interior; break;                                         // aligner.rs:409
            default: valid = (l != 0); break;                                                 // aligner.rs:256,317
        }
        if (!valid) break;
        if (cur == 0) {
            uint32_t t;   // 0 'R', 1 'X', 2 'Y'; m_trace borders: column 0 'X', then row 0 'Y' (aligner.rs:107-108)
            if (l == 0) t = 1; else if (k == 0) t = 2; else t = (nib & TR_YEQ) ? 2u : (nib & TR_XEQ);
            push(t);
            if (t == 0) { step_up(); step_left(); }
            else if (t == 1) { step_up(); cur = 1; }
            else { step_left(); cur = 2; }
        } else if (cur == 1) {
            if (interior && (nib & TR_XOPEN)) cur = 0;                 // x_trace borders stay 'I' (aligner.rs:52)
            else if (k == 0) { flags |= WALK_UNDERFLOW; break; }      // reference: seq1[usize::MAX] -> panic
            else { push(1u); step_up(); }
        } else {
            if (interior && (nib & TR_YOPEN)) cur = 0;
            else if (l == 0) { flags |= WALK_UNDERFLOW; break; }
            else { push(2u); step_left(); }
        }
    }
    if (mode == M_SEMIGLOBAL) {   // aligner.rs:417-428
        if (colbr) { for (uint32_t i = k; i > 0; --i) push(1u); k = 0; }
        else       { for (uint32_t i = l; i > 0; --i) push(2u); l = 0; }
    }
    if (pos & 15u) ops[pos >> 4] = wops;   // partial word (its low ops lie below the string and are never read)
    slotw[0] = k; slotw[1] = l;            // the strings start at seq1[k], seq2[l]
    const uint32_t len = n + m - pos;
    A.score[d.pair_id] = e.score;
    A.walk_flags[d.pair_id] = (uint8_t)ref_status(mode, n, m, e.score, flags);
    A.lens2[2ull * d.pair_id] = len;
    A.lens2[2ull * d.pair_id + 1] = len;
}

// K3, diagonal-window long-pair form: one warp per pair, 2-bit op output.
// A walk over a long pair is ~(n + m) dependent steps; what limits it is how many steps it gets out of one
// round of memory latency.  The path of two related sequences hugs a diagonal, so the window follows the
// diagonal: DIAG_ROWS rows tall, and in row k - r the 9 eight-column blocks around column l - r (+-32
// columns of drift).  The 32 lanes fetch the window with independent loads (one latency for up to
// DIAG_ROWS steps; a square 64 x 64 window, the first design, gave 64), then the reference's scalar state machine
// runs against shared memory until the path leaves the window (long gap runs, drift) and the window is
// re-anchored at the current cell.  Like k3_walk the walk only records 2-bit ops; k_gather turns them into
// strings.  Needs C % 8 == 0 (an eight-column block is one trace word), true for K2 and the wide K1 shapes.
constexpr int DIAG_ROWS = 256;
constexpr int DIAG_WB = 9;
constexpr int DIAG_HALF = 4;          // blocks left of the diagonal's block
constexpr int WALK_DIAG_WARPS = 4;

// CT = 16: the K2 geometry (32 lanes x 16 columns) as compile-time constants -- the window loader's address
// arithmetic (two divisions per word otherwise) is what a reload costs, not the memory round trip; CT = 0: any
// geometry with C % 8 == 0.  The window height adapts: a path that drifts out sideways after a few rows
// (unrelated sequences) gets short windows, a path that uses its window up gets tall ones.
template <int CT>
__global__ void __launch_bounds__(WALK_DIAG_WARPS * 32) k3_walk_diag(const WalkArgs A) {
    __shared__ uint32_t s_tile[WALK_DIAG_WARPS][DIAG_ROWS * DIAG_WB];
    // interior-cell state machine as a table: [state * 16 + code] -> bit 0 emit, bits 1-2 op, bit 3 k--, bit 4 l--,
    // bits 5-6 next state (borders and the local stop test are handled outside the table)
    __shared__ uint8_t s_lut[48];
    if (threadIdx.x < 48) {
        const uint32_t st = threadIdx.x >> 4, nb = threadIdx.x & 15u;
        uint32_t ent;
        if (st == 0) {
            const uint32_t t = (nb & TR_YEQ) ? 2u : (nb & TR_XEQ);
            ent = 1u | (t << 1) | ((t != 2u) << 3) | ((t != 1u) << 4) | (t << 5);
        } else if (st == 1) {
            ent = (nb & TR_XOPEN) ? 0u : (1u | (1u << 1) | (1u << 3) | (1u << 5));
        } else {
            ent = (nb & TR_YOPEN) ? 0u : (1u | (2u << 1) | (1u << 4) | (2u << 5));
        }
        s_lut[threadIdx.x] = (uint8_t)ent;
    }
    __syncthreads();
    const uint32_t wib = threadIdx.x >> 5;
    const uint32_t slot = blockIdx.x * WALK_DIAG_WARPS + wib;
    const uint32_t q = threadIdx.x & 31;
    constexpr unsigned FULL = 0xffffffffu;
    if (slot >= A.n_slots) return;
    const PairDesc d = A.desc[slot];
    if (d.pair_id == 0xFFFFFFFFu) return;
    const EndCell e = A.end[slot];
    const uint32_t n = d.n, m = d.m;
    const uint32_t L = CT ? 32u : (uint32_t)A.L, C = CT ? (uint32_t)CT : (uint32_t)A.C;
    const uint32_t K = (C + 7) / 8;
    const uint32_t band_cols = L * C;
    const uint32_t lane_base = (slot % (32u / L)) * L;
    const int mode = A.mode;
    uint32_t* tile = s_tile[wib];
    uint32_t* slotw = reinterpret_cast<uint32_t*>(A.pad + d.pad_off);
    uint32_t* ops = slotw + 2;
    uint32_t pos = n + m;
    uint32_t wops = 0;
    auto push = [&](uint32_t op) {     // every lane tracks pos / wops, lane 0 stores
        --pos;
        wops |= op << ((pos & 15u) * 2u);
        if ((pos & 15u) == 0) { if (q == 0) ops[pos >> 4] = wops; wops = 0; }
    };
    auto push_run = [&](uint32_t op, uint32_t count) {   // `count` equal ops, up to a word per iteration
        const uint32_t pattern = op * 0x55555555u;
        while (count) {
            const uint32_t hi = (pos & 15u) ? (pos & 15u) : 16u;      // free fields of the current word: [0, hi)
            const uint32_t take = min(count, hi), lo = hi - take;
            const uint32_t mask = (take == 16u) ? 0xffffffffu : (((1u << (2u * take)) - 1u) << (2u * lo));
            wops |= pattern & mask;
            pos -= take; count -= take;
            if ((pos & 15u) == 0) { if (q == 0) ops[pos >> 4] = wops; wops = 0; }
        }
    };

    uint32_t k = e.k, l = e.l, flags = 0;
    const bool colbr = (e.flags & 1u) != 0;
    uint32_t cur = 0;   // 0 = 'M', 1 = 'X', 2 = 'Y'
    uint64_t it = 0;
    uint32_t row0 = 0;
    if (A.cks) {
        const CkptSlot cs = A.cks[slot];
        if (cs.nrows == 0 && !A.last_launch) return;      // this launch holds no rows of the pair
        row0 = cs.row0;
    }
    WalkState* const ws_ = A.wstate ? A.wstate + slot : nullptr;
    if (ws_ && ws_->started) {
        if (ws_->done) return;
        k = ws_->k; l = ws_->l; cur = ws_->cur; pos = ws_->pos; wops = ws_->wops; flags = ws_->flags; it = ws_->it;
        __syncwarp();
    } else if (mode == M_SEMIGLOBAL) {   // aligner.rs:389-404
        if (colbr) push_run(1u, n - k); else push_run(2u, m - l);
    }
    const uint64_t bound = 3ull * ((uint64_t)n + m) + 64 + 2ull * (n / DIAG_ROWS);   // emits + state switches + one probe per window
    uint32_t probe_skip = 0;
    uint32_t win_rows = DIAG_ROWS;
    bool done = false, suspended = false;
    while (!done) {
        // ---- load the window anchored at (k, l): row r is DP row k - r, blocks cb_c(r) - 4 .. cb_c(r) + 4 ----
        const uint32_t k_hi = k, l_hi = l;
        uint32_t rows = 0;
        if (k > row0 && l >= 1) {
            rows = min(k - row0, win_rows);
            for (uint32_t x = q; x < rows * DIAG_WB; x += 32) {
                const uint32_t rr = x / DIAG_WB, bx = x - rr * DIAG_WB;
                const int32_t cb = (((int32_t)l_hi - (int32_t)rr - 1) >> 3) - DIAG_HALF + (int32_t)bx;
                uint32_t wv = 0;
                if (cb >= 0 && ((uint32_t)cb << 3) < m) {
                    const uint32_t i = k_hi - rr;
                    const uint32_t j0 = (uint32_t)cb << 3;
                    const uint32_t bd = j0 / band_cols, rem = j0 - bd * band_cols;
                    const uint32_t p = rem / C, c = rem - p * C;
                    const uint32_t t = (i - 1 - row0) + p;
                    const uint64_t idx = d.trace_off + ((uint64_t)bd * d.steps + t) * (uint64_t)(K * 32u) + (uint64_t)(c >> 3) * 32u + lane_base + p;
                    wv = __ldg(A.trace + idx);
                }
                tile[x] = wv;
            }
        }
        __syncwarp();
        // ---- walk inside the window (all lanes execute it redundantly; lane 0 stores) ----
        for (;;) {
            if (++it > bound) { flags |= WALK_HANG; done = true; break; }
            if (k != 0 && l != 0) {
                if (k <= row0) { suspended = true; done = true; break; }             // the codes of this row belong to the next launch
                const uint32_t rr = k_hi - k;
                const int32_t bx = (int32_t)((l - 1) >> 3) - ((((int32_t)l_hi - (int32_t)rr - 1) >> 3) - DIAG_HALF);
                if (rr >= rows) { win_rows = min(2u * win_rows, (uint32_t)DIAG_ROWS); break; }          // used the window up: re-anchor, taller
                if (bx < 0 || bx >= DIAG_WB) { if (rr < win_rows / 2u) win_rows = max(win_rows / 2u, 32u); break; }   // drifted out: re-anchor
                if (cur == 0 && probe_skip == 0) {
                    // vector probe: lane q looks at cell (k - q, l - q); a run of plain diagonal moves (neither tie
                    // bit set: 'R', and in local mode not the stop code) is emitted at once.  After a short run the
                    // next few steps go through the scalar path (unrelated sequences change state every 2-3 cells)
                    const uint32_t rq = rr + q;
                    bool good = (q < k - row0) && (q < l) && rq < rows;
                    if (good) {
                        const uint32_t lq = l - q;
                        const int32_t bq = (int32_t)((lq - 1) >> 3) - ((((int32_t)l_hi - (int32_t)rq - 1) >> 3) - DIAG_HALF);
                        good = bq >= 0 && bq < DIAG_WB;
                        if (good) good = ((tile[rq * DIAG_WB + (uint32_t)bq] >> (((lq - 1) & 7u) * 4u)) & 3u) == 0u;
                    }
                    const uint32_t mask = __ballot_sync(FULL, good);
                    const uint32_t run = (mask == FULL) ? 32u : (uint32_t)(__ffs((int)~mask) - 1);
                    if (run < 4u) probe_skip = 8u;
                    if (run) { push_run(0u, run); k -= run; l -= run; continue; }
                } else if (probe_skip) --probe_skip;
                const uint32_t nib = (tile[rr * DIAG_WB + (uint32_t)bx] >> (((l - 1) & 7u) * 4u)) & 15u;
                if (mode == M_LOCAL && (nib & 3u) == 3u) { done = true; break; }     // aligner.rs:181
                const uint32_t ent = s_lut[cur * 16u + nib];
                if (ent & 1u) push((ent >> 1) & 3u);
                k -= (ent >> 3) & 1u; l -= (ent >> 4) & 1u; cur = (ent >> 5) & 3u;
                continue;
            }
            // border cells (k == 0 or l == 0): the reference's border trace values (aligner.rs:107-108, 52)
            bool valid;
            switch (mode) {
                case M_GLOBAL: valid = (k != 0 || l != 0); break;
                case M_LOCAL: valid = false; break;
                case M_SEMIGLOBAL: valid = false; break;
                default: valid = (l != 0); break;
            }
            if (!valid) { done = true; break; }
            if (cur == 0) {
                if (l == 0) { push(1u); --k; cur = 1; }
                else { push(2u); --l; cur = 2; }
            } else if (cur == 1) {
                if (k == 0) { flags |= WALK_UNDERFLOW; done = true; break; }
                push(1u); --k;
            } else {
                if (l == 0) { flags |= WALK_UNDERFLOW; done = true; break; }
                push(2u); --l;
            }
        }
        __syncwarp();
    }
    if (suspended) {
        if (q == 0) {
            ws_->k = k; ws_->l = l; ws_->cur = cur; ws_->pos = pos; ws_->wops = wops; ws_->flags = flags; ws_->it = it;
            ws_->started = 1u; ws_->done = 0u;
        }
        return;
    }
    if (mode == M_SEMIGLOBAL) {   // aligner.rs:417-428
        if (colbr) { push_run(1u, k); k = 0; } else { push_run(2u, l); l = 0; }
    }
    if (q == 0) {
        if (ws_) { ws_->started = 1u; ws_->done = 1u; }
        if (pos & 15u) ops[pos >> 4] = wops;
        slotw[0] = k; slotw[1] = l;
        const uint32_t len = n + m - pos;
        A.score[d.pair_id] = e.score;
        A.walk_flags[d.pair_id] = (uint8_t)ref_status(mode, n, m, e.score, flags);
        A.lens2[2ull * d.pair_id] = len;
        A.lens2[2ull * d.pair_id + 1] = len;
    }
}

// K3, skewed-window long-pair form (the default for long pairs): k3_walk_diag's window, stored so that a walk
// step is ONE shared-memory byte load and a handful of ALU instructions.
//   * the window is kept as BYTES in diagonal coordinates: row r (DP row k_hi - r), offset d = l - (l_hi - r) + 36,
//     i.e. the anchor's diagonal is column 36 of every row and a path may drift 35 cells to either side;
//     a diagonal move is idx += 72, a move up (gap in seq2) idx += 73, a move left idx -= 1 -- no row / column /
//     word / shift arithmetic in the loop, and (k, l) are recovered from idx only when the window is left;
//   * the byte of a cell is the state machine's input pre-digested while the window is filled (in parallel, by
//     all lanes): bits 0-1 the move of state 'M' (0 diagonal, 1 up, 2 left; the reference tests M == Y first),
//     bit 2 "state 'X' keeps extending", bit 3 "state 'Y' keeps extending"; 0xFF = not in the window (borders,
//     outside the matrix, rows of another launch, the frame) and 0xFE = local-mode stop (aligner.rs:181), so one
//     compare ends the fast loop, and everything rare -- borders, re-anchoring, suspension at a row-block
//     boundary -- is handled between windows by the same code as in k3_walk_diag;
//   * ops are accumulated by shifting (first emitted = highest position), converted from / to the positional
//     format of WalkState / push_run at the window boundaries.
// The scalar step was ~75 SASS instructions with two dependent shared-memory reads (~300 cycles on the one warp a
// pair has); this one is ~15 with one.  Same op output, same WalkState: interchangeable with k3_walk_diag.
constexpr int SKEW_W = 72;
constexpr int SKEW_C0 = 36;
constexpr int SKEW_ROWS = 256;
constexpr int SKEW_WORDS = 10;        // 8-column trace words per window row (covers 70 + 7 columns at any alignment)
constexpr int WALK_SKEW_WARPS = 2;
constexpr int SKEW_BATCH = 16;        // independent trace-word loads in flight per lane while a window is filled

template <int CT>
__global__ void __launch_bounds__(WALK_SKEW_WARPS * 32) k3_walk_skew(const WalkArgs A) {
    __shared__ __align__(16) uint8_t s_tile[WALK_SKEW_WARPS][(SKEW_ROWS + 1) * SKEW_W + 8];
    const uint32_t wib = threadIdx.x >> 5;
    const uint32_t slot = blockIdx.x * WALK_SKEW_WARPS + wib;
    const uint32_t q = threadIdx.x & 31;
    constexpr unsigned FULL = 0xffffffffu;
    if (slot >= A.n_slots) return;
    const PairDesc d = A.desc[slot];
    if (d.pair_id == 0xFFFFFFFFu) return;
    const EndCell e = A.end[slot];
    const uint32_t n = d.n, m = d.m;
    const uint32_t L = CT ? 32u : (uint32_t)A.L, C = CT ? (uint32_t)CT : (uint32_t)A.C;
    const uint32_t K = (C + 7) / 8;
    const uint32_t band_cols = L * C;
    const uint32_t lane_base = (slot % (32u / L)) * L;
    const int mode = A.mode;
    uint8_t* tile = s_tile[wib];
    uint32_t* slotw = reinterpret_cast<uint32_t*>(A.pad + d.pad_off);
    uint32_t* ops = slotw + 2;
    uint32_t pos = n + m;
    uint32_t wops = 0;
    auto push = [&](uint32_t op) {     // every lane tracks pos / wops, lane 0 stores
        --pos;
        wops |= op << ((pos & 15u) * 2u);
        if ((pos & 15u) == 0) { if (q == 0) ops[pos >> 4] = wops; wops = 0; }
    };
    auto push_run = [&](uint32_t op, uint32_t count) {
        const uint32_t pattern = op * 0x55555555u;
        while (count) {
            const uint32_t hi = (pos & 15u) ? (pos & 15u) : 16u;
            const uint32_t take = min(count, hi), lo = hi - take;
            const uint32_t mask = (take == 16u) ? 0xffffffffu : (((1u << (2u * take)) - 1u) << (2u * lo));
            wops |= pattern & mask;
            pos -= take; count -= take;
            if ((pos & 15u) == 0) { if (q == 0) ops[pos >> 4] = wops; wops = 0; }
        }
    };

    uint32_t k = e.k, l = e.l, flags = 0;
    const bool colbr = (e.flags & 1u) != 0;
    uint32_t cur = 0;   // 0 = 'M', 1 = 'X', 2 = 'Y'
    uint64_t it = 0;
    uint32_t row0 = 0;
    if (A.cks) {
        const CkptSlot cs = A.cks[slot];
        if (cs.nrows == 0 && !A.last_launch) return;
        row0 = cs.row0;
    }
    WalkState* const ws_ = A.wstate ? A.wstate + slot : nullptr;
    if (ws_ && ws_->started) {
        if (ws_->done) return;
        k = ws_->k; l = ws_->l; cur = ws_->cur; pos = ws_->pos; wops = ws_->wops; flags = ws_->flags; it = ws_->it;
        __syncwarp();
    } else if (mode == M_SEMIGLOBAL) {   // aligner.rs:389-404
        if (colbr) push_run(1u, n - k); else push_run(2u, m - l);
    }
    const uint64_t bound = 3ull * ((uint64_t)n + m) + 64 + 4ull * (n / 32u);   // emits + state switches + windows
    uint32_t win_rows = SKEW_ROWS;
    uint32_t probe_skip = 0;
    bool done = false, suspended = false;
    while (!done) {
        if (++it > bound) { flags |= WALK_HANG; break; }
        if (k == 0 || l == 0) {
            // border cells: the reference's border trace values (aligner.rs:107-108, 52)
            bool valid;
            switch (mode) {
                case M_GLOBAL: valid = (k != 0 || l != 0); break;
                case M_LOCAL: valid = false; break;
                case M_SEMIGLOBAL: valid = false; break;
                default: valid = (l != 0); break;
            }
            if (!valid) break;
            if (cur == 0) {
                if (l == 0) { push(1u); --k; cur = 1; }
                else { push(2u); --l; cur = 2; }
            } else if (cur == 1) {
                if (k == 0) { flags |= WALK_UNDERFLOW; break; }
                push(1u); --k;
            } else {
                if (l == 0) { flags |= WALK_UNDERFLOW; break; }
                push(2u); --l;
            }
            continue;
        }
        if (k <= row0) { suspended = true; break; }              // the codes of this row belong to the next launch
        // ---- fill the window anchored at (k, l) ----
        const uint32_t k_hi = k, l_hi = l;
        const uint32_t rows = min(k - row0, win_rows);
        {
            uint4* t4 = reinterpret_cast<uint4*>(tile);
            const uint32_t n16 = ((rows + 1u) * SKEW_W + 15u) / 16u;
            for (uint32_t x = q; x < n16; x += 32) t4[x] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        }
        __syncwarp();
        // The words are fetched in batches of SKEW_BATCH independent loads per lane (one memory round trip per
        // batch; a load-use loop paid one per word: 80 x ~0.7 us per window, which WAS the walk time).
        auto word_of = [&](uint32_t x, uint32_t& rr, uint32_t& j0, int32_t& d0, uint64_t& idxw) -> bool {
            if (x >= rows * SKEW_WORDS) return false;
            rr = x / SKEW_WORDS;
            const uint32_t wx = x - rr * SKEW_WORDS;
            const int32_t ldiag = (int32_t)l_hi - (int32_t)rr;                   // column on the anchor's diagonal in this row
            const int32_t cb = ((ldiag - (SKEW_C0 - 1) - 1) >> 3) + (int32_t)wx;   // word holding column ldiag - 35 (offset 1), then the next ones
            if (cb < 0 || ((uint32_t)cb << 3) >= m) return false;
            const uint32_t i = k_hi - rr;
            j0 = (uint32_t)cb << 3;
            if (CT == 16) {
                idxw = d.trace_off + ((uint64_t)(j0 >> 9) * d.steps + (i - 1 - row0) + ((j0 >> 4) & 31u)) * 64u + ((j0 >> 3) & 1u) * 32u + ((j0 >> 4) & 31u);
            } else {
                const uint32_t bd = j0 / band_cols, rem = j0 - bd * band_cols;
                const uint32_t p = rem / C, c = rem - p * C;
                idxw = d.trace_off + ((uint64_t)bd * d.steps + (i - 1 - row0) + p) * (uint64_t)(K * 32u) + (uint64_t)(c >> 3) * 32u + lane_base + p;
            }
            d0 = (int32_t)j0 + 1 - ldiag + SKEW_C0;                               // offset of the word's first column
            return true;
        };
        for (uint32_t base = 0; base < rows * SKEW_WORDS; base += 32u * SKEW_BATCH) {
            uint32_t wv[SKEW_BATCH];
#pragma unroll
            for (int u = 0; u < SKEW_BATCH; ++u) {
                uint32_t rr, j0; int32_t d0; uint64_t idxw;
                wv[u] = word_of(base + (uint32_t)u * 32u + q, rr, j0, d0, idxw) ? __ldg(A.trace + idxw) : 0u;
            }
#pragma unroll
            for (int u = 0; u < SKEW_BATCH; ++u) {
                uint32_t rr, j0; int32_t d0; uint64_t idxw;
                if (!word_of(base + (uint32_t)u * 32u + q, rr, j0, d0, idxw)) continue;
                uint32_t w = wv[u];
                uint8_t* rowp = tile + rr * SKEW_W;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const uint32_t nib = w & 15u; w >>= 4;
                    const int32_t dd = d0 + c8;
                    if (dd >= 1 && dd <= SKEW_W - 2 && j0 + (uint32_t)c8 < m) {
                        uint32_t bb = ((nib & TR_YEQ) ? 2u : (nib & TR_XEQ)) | ((nib & TR_XOPEN) ? 0u : 4u) | ((nib & TR_YOPEN) ? 0u : 8u);
                        if (mode == M_LOCAL && (nib & 3u) == 3u) bb = 0xFEu;
                        rowp[dd] = (uint8_t)bb;
                    }
                }
            }
        }
        __syncwarp();
        // ---- fast walk inside the window (all lanes redundantly; lane 0 stores) ----
        uint32_t idx = SKEW_C0;
        uint32_t cnt = (pos & 15u) ? (pos & 15u) : 16u;            // free op fields of the current word
        uint32_t sw = (cnt == 16u) ? 0u : (wops >> (2u * cnt));    // its ops so far, shifted form
        uint32_t budget = 4u * (rows + SKEW_W) + 16u;
        uint32_t last = 0xFFu;
        for (;;) {
            if (--budget == 0) { flags |= WALK_HANG; done = true; break; }
            if (cur == 0 && probe_skip == 0) {
                // vector probe: lane q looks at the cell q diagonal steps ahead; a run of plain diagonal moves goes out at once
                const uint32_t pi = idx + (uint32_t)SKEW_W * q;
                const bool good = pi < (rows + 1u) * SKEW_W && (tile[pi] & 0xF3u) == 0u;     // in the window, not a sentinel, move 0
                const uint32_t mask = __ballot_sync(FULL, good);
                uint32_t run = (mask == FULL) ? 32u : (uint32_t)(__ffs((int)~mask) - 1);
                if (run < 4u) probe_skip = 8u;
                idx += SKEW_W * run;
                while (run) {
                    const uint32_t take = min(run, cnt);
                    sw = (take == 16u) ? 0u : (sw << (2u * take));
                    cnt -= take; pos -= take; run -= take;
                    if (cnt == 0) { if (q == 0) ops[pos >> 4] = sw; sw = 0; cnt = 16u; }
                }
            } else if (probe_skip) --probe_skip;
            const uint32_t b = tile[idx];
            if (b >= 0xFEu) { last = b; break; }
            const uint32_t op = cur ? cur : (b & 3u);
            const uint32_t emit = cur ? ((b >> (cur + 1u)) & 1u) : 1u;
            if (emit) {
                sw = (sw << 2) | op;
                --pos;
                if (--cnt == 0) { if (q == 0) ops[pos >> 4] = sw; sw = 0; cnt = 16u; }
                idx += (op == 0u) ? (uint32_t)SKEW_W : (op == 1u) ? (uint32_t)(SKEW_W + 1) : 0xffffffffu;
                cur = op;
            } else {
                cur = 0;
            }
        }
        wops = (cnt == 16u) ? 0u : (sw << (2u * cnt));
        {
            const uint32_t rr = idx / SKEW_W, dd = idx - rr * SKEW_W;
            k = k_hi - rr;
            l = (uint32_t)((int32_t)dd - SKEW_C0 + (int32_t)l_hi - (int32_t)rr);
            if (rr >= rows) win_rows = min(2u * win_rows, (uint32_t)SKEW_ROWS);
            else if (rr < win_rows / 2u) win_rows = max(win_rows / 2u, 32u);
        }
        if (last == 0xFEu) break;                                  // local: M == 0 here (aligner.rs:181)
        __syncwarp();
    }
    if (suspended) {
        if (q == 0) {
            ws_->k = k; ws_->l = l; ws_->cur = cur; ws_->pos = pos; ws_->wops = wops; ws_->flags = flags; ws_->it = it;
            ws_->started = 1u; ws_->done = 0u;
        }
        return;
    }
    if (mode == M_SEMIGLOBAL) {   // aligner.rs:417-428
        if (colbr) { push_run(1u, k); k = 0; } else { push_run(2u, l); l = 0; }
    }
    if (q == 0) {
        if (ws_) { ws_->started = 1u; ws_->done = 1u; }
        if (pos & 15u) ops[pos >> 4] = wops;
        slotw[0] = k; slotw[1] = l;
        const uint32_t len = n + m - pos;
        A.score[d.pair_id] = e.score;
        A.walk_flags[d.pair_id] = (uint8_t)ref_status(mode, n, m, e.score, flags);
        A.lens2[2ull * d.pair_id] = len;
        A.lens2[2ull * d.pair_id + 1] = len;
    }
}

// Score-only epilogue when no traceback is requested.
__global__ void k_scores_only(const PairDesc* desc, const EndCell* end, uint32_t n_slots, int32_t* score,
                              uint8_t* walk_flags, int mode) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const uint32_t id = desc[slot].pair_id;
    if (id == 0xFFFFFFFFu) return;
    score[id] = end[slot].score;
    walk_flags[id] = (uint8_t)ref_status(mode, desc[slot].n, desc[slot].m, end[slot].score, 0u);
}

// Host-buffer pipeline: chunk-relative string offsets -> offsets into the caller's arena.  `base` is a
// device scalar that runs along the chunks of one bg_align_batch call (the chunks' kernels are
// stream-ordered); k_bump publishes this chunk's byte count and advances it.
__global__ void k_rebase(uint64_t* off, uint64_t count, const uint64_t* base) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) off[i] += *base;
}
__global__ void k_bump(uint64_t* base, const uint64_t* chunk_total_entry, uint64_t* chunk_total_out) {
    const uint64_t t = *chunk_total_entry - *base;   // the entry was rebased as well
    *chunk_total_out = t;
    *base += t;
}

// Dense packing: one warp per slot writes a_align then b_align to arena[off[2p]..], arena[off[2p+1]..]: op q reads
// seq1[first_a + #(ops before q that consume seq1)] resp. seq2 likewise, the counts coming from warp ballots.

__global__ void __launch_bounds__(128) k_gather(const GatherArgs A) {
    const uint32_t slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (slot >= A.n_slots) return;
    const PairDesc d = A.desc[slot];
    if (d.pair_id == 0xFFFFFFFFu) return;
    const uint64_t o0 = A.off[2ull * d.pair_id], o1 = A.off[2ull * d.pair_id + 1];
    const uint32_t len = (uint32_t)(o1 - o0);
    {
        const uint32_t* slotw = reinterpret_cast<const uint32_t*>(A.pad + d.pad_off);
        const uint32_t* ops = slotw + 2;
        const uint8_t* sa = A.residues + d.a_off;
        const uint8_t* sb = A.residues + d.b_off;
        uint32_t ia = slotw[0], ib = slotw[1];
        const uint32_t pos0 = d.n + d.m - len;
        const uint32_t lt = (1u << lane) - 1u;
        // Two groups of 32 characters per iteration: the op words and then all four residue loads are issued
        // before anything is stored (a load-use-store loop of one group had one memory round trip per group
        // on its critical path: ncu long-scoreboard 18 stalls per issue).
        for (uint32_t base = 0; base < len; base += 64) {
            const uint32_t x0 = base + lane, x1 = x0 + 32;
            const bool v0 = x0 < len, v1 = x1 < len;
            uint32_t w0 = 0, w1 = 0;
            const uint32_t q0 = pos0 + x0, q1 = pos0 + x1;
            if (v0) w0 = __ldg(ops + (q0 >> 4));
            if (v1) w1 = __ldg(ops + (q1 >> 4));
            const uint32_t op0 = v0 ? (w0 >> ((q0 & 15u) * 2u)) & 3u : 3u, op1 = v1 ? (w1 >> ((q1 & 15u) * 2u)) & 3u : 3u;
            const bool a0 = v0 && op0 != 2u, b0 = v0 && op0 != 1u, a1 = v1 && op1 != 2u, b1 = v1 && op1 != 1u;
            const uint32_t bA0 = __ballot_sync(0xffffffffu, a0), bB0 = __ballot_sync(0xffffffffu, b0);
            const uint32_t bA1 = __ballot_sync(0xffffffffu, a1), bB1 = __ballot_sync(0xffffffffu, b1);
            const uint32_t ia1 = ia + __popc(bA0), ib1 = ib + __popc(bB0);
            uint8_t ca0 = '-', cb0 = '-', ca1 = '-', cb1 = '-';
            if (a0) ca0 = __ldg(sa + ia + __popc(bA0 & lt));
            if (b0) cb0 = __ldg(sb + ib + __popc(bB0 & lt));
            if (a1) ca1 = __ldg(sa + ia1 + __popc(bA1 & lt));
            if (b1) cb1 = __ldg(sb + ib1 + __popc(bB1 & lt));
            if (v0) { A.arena[o0 + x0] = ca0; A.arena[o1 + x0] = cb0; }
            if (v1) { A.arena[o0 + x1] = ca1; A.arena[o1 + x1] = cb1; }
            ia = ia1 + __popc(bA1); ib = ib1 + __popc(bB1);
        }
    }
}

// Compact results for the trip to the host (host-buffer entry points): instead of strings, every pair's ops are
// moved to the front of a dense, word-aligned run -- out word j of pair p holds columns 16 j .. 16 j + 15, column 0
// being the first alignment column -- next to its length and start cell.  About 0.3 B per column travels D2H
// instead of 2 B; host_expand.cpp rebuilds the strings from the caller's own residues.
//   k_ops_counts: counts[p] = {ceil(len / 16), len}; the caller scans them (one scan of the pair) and k_pack_ops places
//   the runs.  k_ops_sample: every `stride`-th entry of the scan plus the totals -- all the host needs to cut a chunk
//   into independent expansion tasks (a task re-derives the offsets inside its own sub-block from the lengths).
__global__ void k_ops_counts(const uint64_t* lens2, uint64_t n_pairs, ulonglong2* counts) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_pairs) { const uint64_t len = lens2[2 * p]; counts[p] = make_ulonglong2((len + 15ull) >> 4, len); }
    if (p == n_pairs) counts[p] = make_ulonglong2(0ull, 0ull);
}
__global__ void k_ops_sample(const ulonglong2* scan, uint64_t n_pairs, uint64_t stride, ulonglong2* samples) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nsub = (n_pairs + stride - 1) / stride;
    if (i < nsub) samples[i] = scan[i * stride];
    if (i == nsub) samples[i] = scan[n_pairs];        // totals: {op words, columns}
}

template <int G>   // lanes per slot
__global__ void __launch_bounds__(128) k_pack_ops(const PackOpsArgs A) {
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const uint32_t lane = threadIdx.x % G;
    if (slot >= A.n_slots) return;
    const PairDesc d = A.desc[slot];
    if (d.pair_id == 0xFFFFFFFFu) return;
    const uint32_t* slotw = reinterpret_cast<const uint32_t*>(A.pad + d.pad_off);
    const uint32_t* ops = slotw + 2;
    const uint32_t len = (uint32_t)A.lens2[2ull * d.pair_id];
    const uint32_t pos0 = d.n + d.m - len;
    const uint32_t cap_words = (d.n + d.m + 15u) >> 4;
    if (lane == 0) {
        A.len[d.pair_id] = len;
        A.first[2ull * d.pair_id] = slotw[0];
        A.first[2ull * d.pair_id + 1] = slotw[1];
    }
    const uint32_t nw = (len + 15u) >> 4, w0 = pos0 >> 4, sh = (pos0 & 15u) * 2u;
    uint32_t* dst = A.ops + A.woff[d.pair_id].x;
    for (uint32_t j = lane; j < nw; j += G) {
        const uint32_t lo = __ldg(ops + w0 + j);
        const uint32_t hi = (sh && w0 + j + 1 < cap_words) ? __ldg(ops + w0 + j + 1) : 0u;
        uint32_t v = sh ? __funnelshift_r(lo, hi, sh) : lo;
        const uint32_t left = len - 16u * j;                 // columns from this word on
        if (left < 16u) v &= (1u << (2u * left)) - 1u;       // deterministic tail bits
        dst[j] = v;
    }
}

}  // namespace bg
