// k4_edit.cuh -- K4: score-only Levenshtein distance (analysis::seq::edit_distance, seq.rs:105-130).
//
// Same systolic decomposition as K1 (lane p owns C consecutive columns, row t-p at step t,
// boundary handed over with __shfl_up_sync), with the unit-cost min-plus cell
//   E[i][j] = min(E[i-1][j-1] + (s1[i-1] != s2[j-1]), E[i][j-1] + 1, E[i-1][j] + 1)
// on raw bytes (the reference compares bytes, any alphabet).  The reference's u128 table is an
// artefact of its container type: distances are bounded by max(len1, len2) < 2^32.
#pragma once
#include "bg_common.cuh"

namespace bg {

struct EditArgs {
    const PairDesc* desc;
    uint32_t n_slots;
    const uint8_t* residues;
    int32_t* bnd;        // band-boundary column scratch (multi-band pairs only), int32 per row
    uint64_t* out;       // [pair]
};

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
