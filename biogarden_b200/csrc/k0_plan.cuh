// k0_plan.cuh -- K0: the launch plan of a pipeline chunk, built on the GPU.
//
// What build_plan() (bg_api.cu) does on the host for every chunk -- order the pairs of a length class by len1
// (longest first, so that the lane groups of a warp carry similar work and K1h finds partners with the same row
// count), insert K1h's holes, and lay out trace blocks, output slots and band scratch -- costs ~25 ns per pair and
// core, more than the GPU needs to ALIGN a 150 bp pair, and ships 64 B per pair over the host link.  Here the host
// sends the chunk's 16 B/pair offsets (which the caller already holds) and per-class totals; a handful of small,
// fully parallel kernels do the rest and produce exactly the PairDesc array the host planner would (same slot order,
// same offsets: the sort is stable, and bg_debug_plan_compare checks it descriptor by descriptor):
//   k_plan_keys        one thread per pair: class (pick_shape_m, the host's function) and sort key
//   (cub radix sort of (key, pair id); skipped for a chunk with one class and one len1)
//   k_plan_clear       every slot starts out empty (the host sized the ranges from upper bounds)
//   (cub exclusive scan of {holes, output bytes, band-scratch elements} per pair, computed on the fly)
//   k_plan_write       one thread per pair: its slot and descriptor
//   k_plan_warp_words  one thread per warp of the launch: step count and trace words
//   (cub exclusive sum)   k_plan_warp_write: steps / trace offset into the warp's slots
// first version: one CTA per class doing all of it in a tile loop -- 2 ms for a 262 144-pair chunk on its single SM,
// which made the e2e path 4 ms SLOWER than planning on the host; this one takes ~0.1 ms.
#pragma once
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "bg_args.cuh"
#include "launch.h"

namespace bg {

__global__ void k_plan_keys(const PlanArgs A) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.n_pairs) return;
    const uint64_t n = A.off[2ull * p + 1] - A.off[2ull * p], m = A.off[2ull * p + 2] - A.off[2ull * p + 1];
    int si = A.force_si;
    if (si < 0) {
        const bool is_long = n + m > LONG_WALK_LEN;
        si = shape_index(pick_shape_m((uint32_t)m, A.half_ok && !is_long, is_long));
    }
    const uint64_t rank = (uint64_t)(uint8_t)A.rank_of_shape[si];
    A.keys[p] = (rank << 32) | (uint64_t)(0x7fffffffu - (uint32_t)n);
    A.ids[p] = p;
}

__global__ void k_plan_clear(PairDesc* desc, uint32_t n_slots) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    PairDesc d;
    d.a_off = d.b_off = d.trace_off = d.bnd_off = d.pad_off = 0; d.n = d.m = d.steps = d.nbands = 0; d.pair_id = 0xFFFFFFFFu; d.pad_ = 1u;
    desc[s] = d;
}

struct PlanSums { uint64_t pad, bnd; uint32_t holes, pad_; };
struct PlanSumsAdd {
    __host__ __device__ PlanSums operator()(const PlanSums& a, const PlanSums& b) const { return PlanSums{a.pad + b.pad, a.bnd + b.bnd, a.holes + b.holes, 0u}; }
};

// What the k-th pair (sorted order) adds to the running sums: its output slot, its band scratch, and -- K1h classes --
// a hole after it when it ends a run of equal len1 of odd length (runs start at even slots, so that is exactly when
// the lane group's second slot has no partner).  The run's start is found by binary search in the sorted keys.
struct PlanItem {
    PlanArgs A; const uint64_t* keys; const uint32_t* ids;
    __device__ PlanSums operator()(uint32_t k) const {
        const uint64_t key = keys[k];
        const PlanCls& c = A.cls[key >> 32];
        const uint64_t id = ids[k];
        const uint64_t o0 = A.off[2 * id], o1 = A.off[2 * id + 1], o2 = A.off[2 * id + 2];
        const uint64_t n = o1 - o0, m = o2 - o1;
        PlanSums s;
        s.pad = 2ull * ((n + m + 3ull) & ~3ull) + 16ull;
        s.bnd = (m > (uint64_t)(c.L * c.C)) ? n : 0ull;
        s.holes = 0; s.pad_ = 0;
        if (c.half && (k + 1 == c.sorted_begin + c.count || keys[k + 1] != key)) {
            uint32_t lo = c.sorted_begin, hi = k;          // first index in [sorted_begin, k] whose key equals `key` (keys ascend)
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
            s.holes = (k - lo + 1u) & 1u;
        }
        return s;
    }
};

__global__ void k_plan_write(const PlanArgs A, const uint64_t* keys, const uint32_t* ids, const PlanSums* sums) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.n_pairs) return;
    const uint64_t key = keys[k];
    const PlanCls& c = A.cls[key >> 32];
    const uint32_t id = ids[k];
    const uint64_t o0 = A.off[2ull * id], o1 = A.off[2ull * id + 1], o2 = A.off[2ull * id + 2];
    const PlanSums s = sums[k];
    const uint32_t holes0 = sums[c.sorted_begin].holes;       // holes of earlier classes (exclusive scan)
    const uint32_t band_cols = (uint32_t)(c.L * c.C);
    PairDesc d;
    d.a_off = o0 - A.base; d.b_off = o1 - A.base;
    d.trace_off = 0; d.steps = 0;
    d.n = (uint32_t)(o1 - o0); d.m = (uint32_t)(o2 - o1);
    d.nbands = d.m ? (d.m + band_cols - 1) / band_cols : 0u;
    d.pair_id = id; d.pad_ = 1u;
    d.pad_off = s.pad;                                          // == the class's pad base + the sum inside the class
    d.bnd_off = (d.nbands > 1u) ? s.bnd : 0ull;
    A.desc[c.slot_begin + (k - c.sorted_begin) + (s.holes - holes0)] = d;
}

// Uniform item (one class, one len1, one len2 -- a read set): no sort, no holes except behind the last pair, and every
// running sum is index x constant.  One launch writes what k_plan_clear .. k_plan_warp_write produce (descriptor by
// descriptor the same: bg_debug_plan_compare).  The ten launches of the general planner each wait for SM slots between
// the blocks of the fill that runs next to them: 0.4-0.9 ms per item against ~0.1 ms of kernel time.
__global__ void k_plan_uniform(const PlanArgs A) {
    const PlanCls& c = A.cls[0];
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= c.slot_cap) return;
    PairDesc d;
    d.a_off = d.b_off = d.trace_off = d.bnd_off = d.pad_off = 0; d.n = d.m = d.steps = d.nbands = 0; d.pair_id = 0xFFFFFFFFu; d.pad_ = 1u;
    const uint32_t L = (uint32_t)c.L, C = (uint32_t)c.C, K = (C + 7u) / 8u, band_cols = L * C;
    // lengths from pair 0 (all pairs are alike)
    const uint64_t n0 = A.off[1] - A.off[0], m0 = A.off[2] - A.off[1];
    const uint32_t nb0 = m0 ? (uint32_t)((m0 + band_cols - 1) / band_cols) : 0u;
    if (s < c.count) {
        const uint64_t o0 = A.off[2ull * s], o1 = A.off[2ull * s + 1];
        d.a_off = o0 - A.base; d.b_off = o1 - A.base;
        d.n = (uint32_t)n0; d.m = (uint32_t)m0; d.nbands = nb0;
        d.pair_id = s;
        d.pad_off = (uint64_t)s * (2ull * ((n0 + m0 + 3ull) & ~3ull) + 16ull);
        d.bnd_off = (nb0 > 1u) ? (uint64_t)s * n0 : 0ull;
    }
    const uint32_t w = s / c.G2;
    if (w * c.G2 < c.count) {        // the warp has at least one pair: every slot of it carries the warp's step count and trace offset
        uint32_t steps; unsigned long long ww;
        if (c.half) {
            steps = (((uint32_t)n0 + L - 1u + HB_TB - 1u) / HB_TB) * HB_TB;
            ww = (unsigned long long)((steps / HB_TB + HB_TG_MAX - 1u) / HB_TG_MAX) * HB_TG_MAX * 32ull * hb_words_per_lane_block((int)C);
        } else {
            steps = (uint32_t)n0 + L - 1u;
            ww = (unsigned long long)nb0 * steps * K * 32ull;
        }
        if (steps) { d.steps = steps; d.trace_off = (unsigned long long)w * ww; }
    }
    A.desc[c.slot_begin + s] = d;
}

// warp w of the item (classes back to back, slot_cap / G2 warps each) -> (class, first slot)
__device__ __forceinline__ bool plan_warp_locate(const PlanArgs& A, uint32_t w, uint32_t& cls, uint32_t& slot0) {
    uint32_t first = 0;
    for (uint32_t k = 0; k < A.n_cls; ++k) {
        const uint32_t nw = A.cls[k].slot_cap / A.cls[k].G2;
        if (w < first + nw) { cls = k; slot0 = A.cls[k].slot_begin + (w - first) * A.cls[k].G2; return true; }
        first += nw;
    }
    return false;
}

__global__ void k_plan_warp_words(const PlanArgs A, unsigned long long* words, uint32_t* steps_out, uint32_t n_warps) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > n_warps) return;
    if (w == n_warps) { words[w] = 0; return; }
    uint32_t k = 0, slot0 = 0;
    uint32_t steps = 0; unsigned long long ww = 0;
    if (plan_warp_locate(A, w, k, slot0)) {
        const PlanCls& c = A.cls[k];
        const uint32_t L = (uint32_t)c.L, C = (uint32_t)c.C, K = (C + 7u) / 8u;
        uint32_t maxn = 0, maxb = 0; bool any = false;
        for (uint32_t g = 0; g < c.G2; ++g) {
            const PairDesc& d = A.desc[slot0 + g];
            if (d.pair_id == 0xFFFFFFFFu) continue;
            any = true; maxn = max(maxn, d.n); maxb = max(maxb, d.nbands);
        }
        if (any) {
            if (c.half) {
                steps = ((maxn + L - 1u + HB_TB - 1u) / HB_TB) * HB_TB;
                ww = (unsigned long long)((steps / HB_TB + HB_TG_MAX - 1u) / HB_TG_MAX) * HB_TG_MAX * 32ull * hb_words_per_lane_block((int)C);
            } else {
                steps = maxn + L - 1u;
                ww = (unsigned long long)maxb * steps * K * 32ull;
            }
        }
    }
    words[w] = ww; steps_out[w] = steps;
}

__global__ void k_plan_warp_write(const PlanArgs A, const unsigned long long* woff, const uint32_t* steps, uint32_t n_warps) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_warps || !steps[w]) return;
    uint32_t k = 0, slot0 = 0;
    if (!plan_warp_locate(A, w, k, slot0)) return;
    const PlanCls& c = A.cls[k];
    uint32_t first = 0;                                  // first warp of the class: trace offsets restart per launch
    for (uint32_t q = 0; q < k; ++q) first += A.cls[q].slot_cap / A.cls[q].G2;
    const unsigned long long off = woff[w] - woff[first];
    for (uint32_t g = 0; g < c.G2; ++g) {
        PairDesc& d = A.desc[slot0 + g];
        d.steps = steps[w]; d.trace_off = off;
    }
}

}  // namespace bg
