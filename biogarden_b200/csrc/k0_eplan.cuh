// k0_eplan.cuh -- K0e: launch slots of the bit-parallel edit-distance kernel (K4b, k4_edit.cuh) built on the GPU.
//
// K4b runs one thread per pair; what it needs from a plan is (a) the pairs of one block count W = 4 / 8 / 10
// (len2 <= 128 / 256 / 320) next to each other, because W is a template argument, and (b) neighbouring threads with
// similar len1, because a warp runs as long as its longest text.  The host planner (build_plan) spends ~35 ns per pair
// and core on that.  Here the chunk's offsets (16 B per pair, which the caller holds anyway) go to the device and two
// small kernels bucket the pairs -- a counting sort over 3 classes x 64 len1 ranges, longest range first:
//   k_eplan_count     one thread per pair: histogram of the buckets (shared-memory atomics, one global add per block and bucket)
//   k_eplan_scatter   every block scans the 192 counters (an exclusive sum is a few hundred adds), then one thread per
//                     pair takes the next free slot of its bucket and writes the 16-byte MyersSlot
// The order inside a bucket is whatever the atomics give: it only affects speed, never results.  (First version: 16-bit
// keys + cub radix sort; its five launches with decoupled look-back took 0.4-0.9 ms per chunk next to running K4b
// grids -- longer than the chunk's K4b itself.)  The kernels of the three classes are launched over the whole chunk
// and take their slot range from cls_count.
#pragma once
#include "bg_args.cuh"

namespace bg {

constexpr int EPLAN_NB = 64;                       // len1 ranges per class
constexpr int EPLAN_BUCKETS = 3 * EPLAN_NB + 1;    // + one bucket for pairs that do not fit K4b / the 16-byte slot

__device__ __forceinline__ uint32_t eplan_bucket(const EditPlanArgs& A, uint64_t o0, uint64_t o1, uint64_t o2) {
    if (o1 < o0 || o2 < o1) return 3u * EPLAN_NB;          // not monotone (k_eplan_count reports it)
    const uint64_t n = o1 - o0, m = o2 - o1;
    if (m > 320 || n > 0xFFFFFFFFull || (o0 - A.base) >= (1ull << 48)) return 3u * EPLAN_NB;
    const uint32_t cls = m <= 128 ? 0u : m <= 256 ? 1u : 2u;
    const uint32_t r = (uint32_t)min((uint64_t)(EPLAN_NB - 1), n >> A.n_shift);
    return cls * EPLAN_NB + (EPLAN_NB - 1u - r);   // longest first inside a class
}

// (a kernel, not cudaMemsetAsync: memsets and small copies may be queued on a copy engine behind the bulk uploads of the next chunks)
__global__ void k_eplan_zero(uint32_t* scratch, uint32_t words, uint32_t* err_flag) {
    for (uint32_t x = threadIdx.x; x < words; x += blockDim.x) scratch[x] = 0;
    if (threadIdx.x == 0 && err_flag) *err_flag = 0;
}

__global__ void __launch_bounds__(256) k_eplan_count(const EditPlanArgs A) {
    __shared__ uint32_t s_cnt[EPLAN_BUCKETS];
    for (int x = threadIdx.x; x < EPLAN_BUCKETS; x += blockDim.x) s_cnt[x] = 0;
    __syncthreads();
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    // the host has not looked at the offsets (that pass cost more than the GPU needs for the distances): they are
    // validated here, and the chunk's cell count is summed for bg_last_timing
    unsigned long long cells = 0;
    bool bad = false;
    if (p < A.n_pairs) {
        const uint64_t o0 = A.off[2ull * p], o1 = A.off[2ull * p + 1], o2 = A.off[2ull * p + 2];
        bad = o1 < o0 || o2 < o1;
        const uint32_t b = eplan_bucket(A, o0, o1, o2);
        if (b < 3u * EPLAN_NB) cells = (o1 - o0) * (o2 - o1);
        atomicAdd(&s_cnt[b], 1u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
    if ((threadIdx.x & 31) == 0 && cells) atomicAdd(A.cells, cells);
    if (bad) atomicOr(A.err_flag, 4u);
    __syncthreads();
    for (int x = threadIdx.x; x < EPLAN_BUCKETS; x += blockDim.x)
        if (s_cnt[x]) atomicAdd(A.hist + x, s_cnt[x]);
}

__global__ void __launch_bounds__(256) k_eplan_scatter(const EditPlanArgs A) {
    __shared__ uint32_t s_base[EPLAN_BUCKETS];
    if (threadIdx.x < 32) {   // exclusive sum of the histogram by one warp
        uint32_t run = 0;
        for (int x0 = 0; x0 < EPLAN_BUCKETS; x0 += 32) {
            const int x = x0 + (int)threadIdx.x;
            const uint32_t v = x < EPLAN_BUCKETS ? A.hist[x] : 0u;
            uint32_t inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if ((int)threadIdx.x >= o) inc += u; }
            if (x < EPLAN_BUCKETS) s_base[x] = run + inc - v;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < 4) {
        // slots per class; [3]: pairs that do not fit (the host then redoes the batch on the general path)
        const uint32_t lo = s_base[threadIdx.x * EPLAN_NB];
        const uint32_t hi = threadIdx.x < 3 ? s_base[(threadIdx.x + 1) * EPLAN_NB] : lo + A.hist[3 * EPLAN_NB];
        A.cls_count[threadIdx.x] = hi - lo;
    }
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.n_pairs) return;
    const uint64_t o0 = A.off[2ull * p], o1 = A.off[2ull * p + 1], o2 = A.off[2ull * p + 2];
    const uint32_t b = eplan_bucket(A, o0, o1, o2);
    const uint32_t s = s_base[b] + atomicAdd(A.cursor + b, 1u);
    const uint64_t a_off = o0 - A.base;
    MyersSlot ms;
    ms.a_off_lo = (uint32_t)a_off; ms.a_off_hi = (uint16_t)(a_off >> 32); ms.pair_id = p;
    ms.n = (uint32_t)(o1 - o0); ms.m = (uint16_t)(o2 - o1);
    A.slots[s] = ms;
}

}  // namespace bg
