// k0_plan.cuh -- K0: the launch plan of a pipeline chunk, built on the GPU.
//
// What build_plan() (bg_api.cu) does on the host for every chunk -- order the pairs of a length class by len1
// (longest first, so that the lane groups of a warp carry similar work and K1h finds partners with the same row
// count), insert K1h's holes, and lay out trace blocks, output slots and band scratch -- costs ~25 ns per pair and
// core, more than the GPU needs to ALIGN a 150 bp pair, and ships 64 B per pair over the host link.  Here the host
// sends the chunk's 16 B/pair offsets (which the caller already holds) and per-class totals; three kernels do the
// rest and produce exactly the PairDesc array the host planner would (same slot order, same offsets: the sort is
// stable, and bg_debug_plan_compare checks it descriptor by descriptor):
//   k_plan_keys   one thread per pair: class (pick_shape_m, the host's function) and sort key
//   (cub radix sort of (key, pair id); skipped for a chunk with one class and one len1)
//   k_plan_build  one CTA per class: runs of equal len1 -> holes, slot numbers, exclusive sums of the slot sizes,
//                 then per warp the step count and the trace offset
#pragma once
#include <cub/block/block_scan.cuh>

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

constexpr int PLAN_TPB = 1024;
constexpr int PLAN_IPT = 4;

struct PlanSums { uint32_t holes; uint64_t pad, bnd; };
struct PlanSumsAdd {
    __device__ PlanSums operator()(const PlanSums& a, const PlanSums& b) const { return PlanSums{a.holes + b.holes, a.pad + b.pad, a.bnd + b.bnd}; }
};

// keys_sorted / ids_sorted: the sorted order (nullptr: identity order, one run).
__global__ void __launch_bounds__(PLAN_TPB) k_plan_build(const PlanArgs A, const uint64_t* keys_sorted, const uint32_t* ids_sorted) {
    using ScanMax = cub::BlockScan<uint32_t, PLAN_TPB>;
    using ScanSum = cub::BlockScan<PlanSums, PLAN_TPB>;
    using ScanU64 = cub::BlockScan<unsigned long long, PLAN_TPB>;
    __shared__ union { typename ScanMax::TempStorage mx; typename ScanSum::TempStorage sm; typename ScanU64::TempStorage u; } tmp;
    __shared__ uint32_t s_run_start;     // carries between tiles
    __shared__ PlanSums s_carry;
    __shared__ unsigned long long s_tcarry;
    const PlanCls c = A.cls[blockIdx.x];
    const uint32_t tid = threadIdx.x;
    const uint32_t L = (uint32_t)c.L, C = (uint32_t)c.C, band_cols = L * C, K = (C + 7u) / 8u;

    // every slot of the class's range starts out empty (the host sized the range from upper bounds)
    for (uint32_t s = tid; s < c.slot_cap; s += PLAN_TPB) {
        PairDesc d;
        d.a_off = d.b_off = d.trace_off = d.bnd_off = d.pad_off = 0; d.n = d.m = d.steps = d.nbands = 0; d.pair_id = 0xFFFFFFFFu; d.pad_ = 1u;
        A.desc[c.slot_begin + s] = d;
    }
    if (tid == 0) { s_run_start = 0; s_carry = PlanSums{0u, 0ull, 0ull}; s_tcarry = 0ull; }
    __syncthreads();

    auto len1_at = [&](uint32_t k) -> uint32_t {    // len1 of the k-th pair of the class in sorted order
        if (keys_sorted) return 0x7fffffffu - (uint32_t)(keys_sorted[c.sorted_begin + k] & 0xffffffffull);
        const uint64_t id = ids_sorted ? ids_sorted[c.sorted_begin + k] : (uint64_t)(c.sorted_begin + k);
        return (uint32_t)(A.off[2 * id + 1] - A.off[2 * id]);
    };

    for (uint32_t tile = 0; tile < c.count; tile += PLAN_TPB * PLAN_IPT) {
        uint32_t id[PLAN_IPT], n[PLAN_IPT], m[PLAN_IPT], rs[PLAN_IPT], nb[PLAN_IPT];
        bool valid[PLAN_IPT], is_end[PLAN_IPT];
        uint64_t o0[PLAN_IPT], o1[PLAN_IPT];
        uint32_t local_max = 0;
#pragma unroll
        for (int j = 0; j < PLAN_IPT; ++j) {
            const uint32_t k = tile + tid * PLAN_IPT + j;
            valid[j] = k < c.count;
            id[j] = 0; n[j] = m[j] = 0; rs[j] = 0; nb[j] = 0; is_end[j] = false; o0[j] = o1[j] = 0;
            if (valid[j]) {
                id[j] = ids_sorted ? ids_sorted[c.sorted_begin + k] : c.sorted_begin + k;
                o0[j] = A.off[2ull * id[j]]; o1[j] = A.off[2ull * id[j] + 1];
                n[j] = (uint32_t)(o1[j] - o0[j]); m[j] = (uint32_t)(A.off[2ull * id[j] + 2] - o1[j]);
                const bool is_start = (k == 0) || len1_at(k - 1) != n[j];
                is_end[j] = (k + 1 == c.count) || len1_at(k + 1) != n[j];
                if (is_start) local_max = k + 1;          // (k + 1 so that 0 means "no run starts here")
                rs[j] = local_max;
                nb[j] = m[j] ? (m[j] + band_cols - 1) / band_cols : 0u;
            }
        }
        // start of the run every element belongs to: inclusive max-scan of the run starts
        uint32_t excl_max = 0;
        ScanMax(tmp.mx).ExclusiveScan(local_max, excl_max, 0u, cub::Max());
        __syncthreads();
        const uint32_t carry_rs = s_run_start;
        PlanSums local{0u, 0ull, 0ull}, item[PLAN_IPT];
#pragma unroll
        for (int j = 0; j < PLAN_IPT; ++j) {
            item[j] = local;       // exclusive within the thread
            if (valid[j]) {
                const uint32_t k = tile + tid * PLAN_IPT + j;
                uint32_t r = rs[j] ? rs[j] : excl_max;
                if (!r) r = carry_rs;
                rs[j] = r - 1;                            // first element of the run
                const uint32_t hole = (c.half && is_end[j] && ((k - rs[j] + 1u) & 1u)) ? 1u : 0u;
                local.holes += hole;
                local.pad += 2ull * (((uint64_t)n[j] + m[j] + 3ull) & ~3ull) + 16ull;
                local.bnd += (nb[j] > 1u) ? n[j] : 0u;
            }
        }
        PlanSums excl, total;
        ScanSum(tmp.sm).ExclusiveScan(local, excl, PlanSums{0u, 0ull, 0ull}, PlanSumsAdd(), total);
        const PlanSums carry = s_carry;
        __syncthreads();
        // carries for the next tile: the last run start seen so far, the running sums
        if (tid == PLAN_TPB - 1) {
            const uint32_t last = local_max ? local_max : (excl_max ? excl_max : carry_rs);
            s_run_start = last;
            s_carry = PlanSums{carry.holes + total.holes, carry.pad + total.pad, carry.bnd + total.bnd};
        }
#pragma unroll
        for (int j = 0; j < PLAN_IPT; ++j) {
            if (!valid[j]) continue;
            const uint32_t k = tile + tid * PLAN_IPT + j;
            const uint32_t slot = c.slot_begin + k + carry.holes + excl.holes + item[j].holes;
            PairDesc d;
            d.a_off = o0[j] - A.base; d.b_off = o1[j] - A.base;
            d.trace_off = 0; d.steps = 0;
            d.n = n[j]; d.m = m[j]; d.nbands = nb[j]; d.pair_id = id[j]; d.pad_ = 1u;
            d.pad_off = c.pad_base + carry.pad + excl.pad + item[j].pad;
            d.bnd_off = (nb[j] > 1u) ? c.bnd_base + carry.bnd + excl.bnd + item[j].bnd : 0ull;
            A.desc[slot] = d;
        }
        __syncthreads();
    }
    __syncthreads();

    // per warp: systolic steps (longest pair of the warp) and the trace block
    const uint32_t nwarps = c.slot_cap / c.G2;
    for (uint32_t wt = 0; wt < nwarps; wt += PLAN_TPB) {
        const uint32_t w = wt + tid;
        uint32_t steps = 0; unsigned long long words = 0;
        if (w < nwarps) {
            uint32_t maxn = 0, maxb = 0; bool any = false;
            for (uint32_t g = 0; g < c.G2; ++g) {
                const PairDesc& d = A.desc[c.slot_begin + w * c.G2 + g];
                if (d.pair_id == 0xFFFFFFFFu) continue;
                any = true; maxn = max(maxn, d.n); maxb = max(maxb, d.nbands);
            }
            if (any) {
                if (c.half) {
                    steps = ((maxn + L - 1u + HB_TB - 1u) / HB_TB) * HB_TB;
                    words = (unsigned long long)((steps / HB_TB + HB_TG_MAX - 1u) / HB_TG_MAX) * HB_TG_MAX * 32ull * hb_words_per_lane_block((int)C);
                } else {
                    steps = maxn + L - 1u;
                    words = (unsigned long long)maxb * steps * K * 32ull;
                }
            }
        }
        unsigned long long excl = 0, total = 0;
        ScanU64(tmp.u).ExclusiveSum(words, excl, total);
        const unsigned long long tc = s_tcarry;
        __syncthreads();
        if (tid == 0) s_tcarry = tc + total;
        if (w < nwarps && steps) {
            for (uint32_t g = 0; g < c.G2; ++g) {
                PairDesc& d = A.desc[c.slot_begin + w * c.G2 + g];
                d.steps = steps; d.trace_off = tc + excl;
            }
        }
        __syncthreads();
    }
}

}  // namespace bg
