// k4_edit.cuh -- K4: score-only Levenshtein distance (analysis::seq::edit_distance, seq.rs:105-130).
//
// Same systolic decomposition as K1 (lane p owns C consecutive columns, row t-p at step t,
// boundary handed over with __shfl_up_sync), with the unit-cost min-plus cell
//   E[i][j] = min(E[i-1][j-1] + (s1[i-1] != s2[j-1]), E[i][j-1] + 1, E[i-1][j] + 1)
// on raw bytes (the reference compares bytes, any alphabet).  The reference's u128 table is an
// artefact of its container type: distances are bounded by max(len1, len2) < 2^32.
#pragma once
#include "bg_args.cuh"

namespace bg {


template <int L, int C>
__global__ void __launch_bounds__(128) k4_edit(const EditArgs A) {
    constexpr int G = 32 / L;
    constexpr unsigned FULL = 0xffffffffu;
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
    const uint8_t* sa = A.residues + d.a_off;
    const uint8_t* sb = A.residues + d.b_off;
    const uint32_t band_cols = (uint32_t)(L * C);
    const uint32_t mcol0 = m ? m - 1 : 0;
    const uint32_t p_m = (mcol0 % band_cols) / C;
    int32_t corner = (int32_t)n;   // E[n][0]

    for (uint32_t bd = 0; bd < nbands_w; ++bd) {
        const bool band_on = has_pair && bd < my_nbands;
        const uint32_t jbase = bd * band_cols + (uint32_t)p * C;
        const bool last_band_for_pair = (bd + 1 == my_nbands);
        uint32_t cb[C];
        int32_t Eu[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const uint32_t j0 = jbase + c;
            cb[c] = (band_on && j0 < m) ? (uint32_t)sb[j0] : 0x100u;   // 0x100 never equals a byte
            Eu[c] = (int32_t)(j0 + 1);                                  // E[0][j]
        }
        int32_t Ediag = (int32_t)jbase;   // E[0][jbase]
        int32_t Elast = 0;
        uint32_t rcur = 0;
        auto load_rows = [&](uint32_t base) -> uint32_t {
            const uint32_t idx = base + (uint32_t)p;
            return (band_on && idx < n) ? (uint32_t)sa[idx] : 0u;
        };
        uint32_t cur_blk = 0, next_blk = load_rows(0);
        int32_t bnd_in = 0;
        if (bd > 0 && p == 0 && band_on && n > 0) bnd_in = __ldcg(A.bnd + d.bnd_off);

        for (uint32_t t = 0; t < steps_w; ++t) {
            if ((t & (L - 1)) == 0) { cur_blk = next_blk; next_blk = load_rows(t + L); }
            const uint32_t r0 = __shfl_sync(FULL, cur_blk, (int)(t & (L - 1)), L);
            int32_t El = __shfl_up_sync(FULL, Elast, 1, L);
            uint32_t r = __shfl_up_sync(FULL, rcur, 1, L);
            const uint32_t i0 = t - (uint32_t)p;
            const bool active = band_on && i0 < n;
            if (p == 0) {
                r = r0;
                if (bd == 0) El = (int32_t)(i0 + 1);   // E[i][0] = i
                else {
                    El = bnd_in;
                    if (band_on && i0 + 1 < n) bnd_in = __ldcg(A.bnd + d.bnd_off + i0 + 1);
                }
            }
            rcur = r;
            if (active) {
                int32_t diag = Ediag, left = El;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int32_t up = Eu[c];
                    const int32_t sub = diag + (r != cb[c] ? 1 : 0);
                    const int32_t v = min(sub, min(left, up) + 1);
                    diag = up; left = v; Eu[c] = v;
                }
                Elast = left; Ediag = El;
                if (p == L - 1 && !last_band_for_pair) A.bnd[d.bnd_off + i0] = Elast;
            }
        }
        if (band_on) {
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (jbase + c + 1 == m) corner = Eu[c];
        }
        __syncwarp();
    }
    const int32_t corner0 = __shfl_sync(FULL, corner, g * L + (int)p_m);
    if (p == 0 && has_pair) A.out[d.pair_id] = (uint64_t)(uint32_t)corner0;
}

}  // namespace bg

// ---------------------------------------------------------------------------------------------
// K4b: bit-parallel Levenshtein distance (Myers 1999, block formulation as in Hyyro / Edlib) for
// batches whose residues take at most 4 distinct byte values (DNA reads).  One thread per pair; the
// pattern (seq2) lives in W 32-bit blocks of vertical delta vectors Pv / Mv plus 4 match masks per
// block, all in registers; every text residue costs ~20 integer ops per BLOCK OF 32 CELLS instead of
// ~7 per cell.  Exact: the same D[n][m] as the reference's table (seq.rs:105-130).  Pairs with
// len2 > 32 * W or richer alphabets take the systolic kernel above.
namespace bg {



template <int W>
__global__ void __launch_bounds__(128) k4_myers(const MyersArgs A) {
    __shared__ uint8_t s_lut[256];
    for (int x = threadIdx.x; x < 256; x += blockDim.x) s_lut[x] = A.lut[x];
    __syncthreads();
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (A.cls_count) {   // device-side plan: this class's slot range comes from the histogram
        uint32_t begin = 0;
        for (uint32_t k = 0; k < A.cls; ++k) begin += A.cls_count[k];
        if (slot >= A.cls_count[A.cls]) return;
        slot += begin;
    } else if (slot >= A.n_slots) return;
    struct { uint64_t a_off, b_off; uint32_t n, m, pair_id; } d;
    if (A.cdesc) {
        const MyersSlot c = A.cdesc[slot];
        d.a_off = ((uint64_t)c.a_off_hi << 32) | c.a_off_lo; d.n = c.n; d.m = c.m; d.pair_id = c.pair_id; d.b_off = d.a_off + c.n;
    } else {
        const PairDesc f = A.desc[slot];
        d.a_off = f.a_off; d.b_off = f.b_off; d.n = f.n; d.m = f.m; d.pair_id = f.pair_id;
    }
    if (d.pair_id == 0xFFFFFFFFu) return;
    const uint32_t n = d.n, m = d.m;
    if (m == 0 || n == 0) { A.out[d.pair_id] = (uint64_t)(m == 0 ? n : m); return; }
    const uint8_t* text = A.residues + d.a_off;
    const uint8_t* pat = A.residues + d.b_off;

    // aligned 32-bit fetches of the byte streams
    auto byte_at = [&](const uint8_t* base, uint32_t idx, uintptr_t& wp, uint32_t& cw) -> uint32_t {
        const uintptr_t q = reinterpret_cast<uintptr_t>(base + idx), w = q & ~(uintptr_t)3;
        if (w != wp) { wp = w; cw = __ldg(reinterpret_cast<const uint32_t*>(w)); }
        return (cw >> ((q & 3u) * 8u)) & 0xffu;
    };

    uint32_t peq0[W], peq1[W], peq2[W], peq3[W], Pv[W], Mv[W];
    uintptr_t wp = 0; uint32_t cw = 0;
    bool bad = false;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
        if ((uint32_t)w * 32u < m) {
            const uint32_t cnt = min(32u, m - (uint32_t)w * 32u);
            for (uint32_t bpos = 0; bpos < cnt; ++bpos) {
                const uint32_t c = s_lut[byte_at(pat, (uint32_t)w * 32u + bpos, wp, cw)];
                bad |= (c > 3u);
                const uint32_t bit = 1u << bpos;
                e0 |= (c == 0) ? bit : 0u; e1 |= (c == 1) ? bit : 0u; e2 |= (c == 2) ? bit : 0u; e3 |= (c == 3) ? bit : 0u;
            }
        }
        peq0[w] = e0; peq1[w] = e1; peq2[w] = e2; peq3[w] = e3;
        Pv[w] = 0xffffffffu; Mv[w] = 0u;
    }
    const uint32_t wl = (m - 1) >> 5, lastbit = (m - 1) & 31u;
    int32_t score = (int32_t)m;
    wp = 0; cw = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t c = s_lut[byte_at(text, i, wp, cw)];
        bad |= (c > 3u);
        const bool is0 = (c == 0), is1 = (c == 1), is2 = (c == 2);
        uint32_t hp = 1u, hm = 0u;           // horizontal delta entering block 0: D[i][0] - D[i-1][0] = +1
#pragma unroll
        for (int w = 0; w < W; ++w) {
            if ((uint32_t)w <= wl) {
                uint32_t Eq = is0 ? peq0[w] : (is1 ? peq1[w] : (is2 ? peq2[w] : peq3[w]));
                const uint32_t pv = Pv[w], mv = Mv[w];
                const uint32_t Xv = Eq | mv;
                Eq |= hm;
                const uint32_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
                uint32_t Ph = mv | ~(Xh | pv);
                uint32_t Mh = pv & Xh;
                if ((uint32_t)w == wl) score += (int32_t)((Ph >> lastbit) & 1u) - (int32_t)((Mh >> lastbit) & 1u);
                const uint32_t hp_out = Ph >> 31, hm_out = Mh >> 31;
                Ph = (Ph << 1) | hp;
                Mh = (Mh << 1) | hm;
                Pv[w] = Mh | ~(Xv | Ph);
                Mv[w] = Ph & Xv;
                hp = hp_out; hm = hm_out;
            }
        }
    }
    A.out[d.pair_id] = (uint64_t)(uint32_t)score;
    if (bad) atomicOr(A.err_flag, 2u);
}

// 256-bin byte histogram of a residue arena (which symbols occur at all)
__global__ void k_byte_hist(const uint8_t* data, uint64_t n, unsigned int* hist) {
    __shared__ unsigned int s[256];
    for (int x = threadIdx.x; x < 256; x += blockDim.x) s[x] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 16ull;
    for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16ull; base < n; base += stride) {
        if (base + 16 <= n && ((reinterpret_cast<uintptr_t>(data) + base) & 15) == 0) {
            const uint4 v = *reinterpret_cast<const uint4*>(data + base);
            const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int b = 0; b < 4; ++b) atomicAdd(&s[(ws[k] >> (8 * b)) & 0xffu], 1u);
        } else {
            for (uint64_t x = base; x < n && x < base + 16; ++x) atomicAdd(&s[data[x]], 1u);
        }
    }
    __syncthreads();
    for (int x = threadIdx.x; x < 256; x += blockDim.x) if (s[x]) atomicAdd(&hist[x], 1u);
}

}  // namespace bg
