// k5_distance.cuh -- K5: the "compare two sequences position by position" primitives next to the alignment path
// (SURVEY 8f, rank 4): analysis::seq::hamming_distance (seq.rs:74-83) batched, and analysis::stat::
// p_distance_matrix (stat.rs:138-152).  Both are pure byte compares: HBM-bound (2 bytes read per position), so
// the kernels are about coalesced 16-byte loads and enough bytes in flight per SM -- no shared memory, no
// tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bg_args.cuh"

namespace bg {

__host__ __device__ __forceinline__ uint64_t umin64(uint64_t x, uint64_t y) { return x < y ? x : y; }

// differing bytes of two 32-bit words: __vcmpne4 gives 0xff per differing byte lane
__device__ __forceinline__ uint32_t diff4(uint32_t x, uint32_t y) { return __popc(__vcmpne4(x, y)) >> 3; }

// count of positions x in [0, len) with a[x] != b[x]; a group of G lanes (G = 8, 16, 32; groups aligned in the
// warp, ALL 32 lanes must call) cooperates, every lane of the group returns the group's total.
// Arbitrary alignment of a and b: a scalar head up to a's 16-byte boundary, then 16-byte loads of a, with b
// loaded 16-byte aligned as well when it has the same misalignment, word-wise when it is word-aligned relative
// to a, and assembled from aligned words with a funnel shift otherwise.
template <int G>
__device__ __forceinline__ uint32_t group_mismatches(const uint8_t* a, const uint8_t* b, uint64_t len, uint32_t lane) {
    uint32_t cnt = 0;
    const uint64_t head = umin64(len, (uint64_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(a) & 15u)) & 15u));
    for (uint64_t x = lane; x < head; x += G) cnt += a[x] != b[x];
    const uint8_t* a2 = a + head; const uint8_t* b2 = b + head;
    const uint64_t body = (len - head) / 16;
    if ((reinterpret_cast<uintptr_t>(b2) & 15u) == 0) {
        const uint4* pa = reinterpret_cast<const uint4*>(a2); const uint4* pb = reinterpret_cast<const uint4*>(b2);
        uint64_t x = lane;
        for (; x + G < body; x += 2 * G) {          // two independent 16-byte pairs in flight per lane
            const uint4 va = __ldg(pa + x), vb = __ldg(pb + x), wa = __ldg(pa + x + G), wb = __ldg(pb + x + G);
            cnt += diff4(va.x, vb.x) + diff4(va.y, vb.y) + diff4(va.z, vb.z) + diff4(va.w, vb.w);
            cnt += diff4(wa.x, wb.x) + diff4(wa.y, wb.y) + diff4(wa.z, wb.z) + diff4(wa.w, wb.w);
        }
        for (; x < body; x += G) {
            const uint4 va = __ldg(pa + x), vb = __ldg(pb + x);
            cnt += diff4(va.x, vb.x) + diff4(va.y, vb.y) + diff4(va.z, vb.z) + diff4(va.w, vb.w);
        }
    } else if ((reinterpret_cast<uintptr_t>(b2) & 3u) == 0) {
        const uint4* pa = reinterpret_cast<const uint4*>(a2); const uint32_t* pb = reinterpret_cast<const uint32_t*>(b2);
        for (uint64_t x = lane; x < body; x += G) {
            const uint4 va = __ldg(pa + x);
            cnt += diff4(va.x, __ldg(pb + 4 * x)) + diff4(va.y, __ldg(pb + 4 * x + 1)) + diff4(va.z, __ldg(pb + 4 * x + 2)) + diff4(va.w, __ldg(pb + 4 * x + 3));
        }
    } else {
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(b2) & 3u) * 8u;
        const uint4* pa = reinterpret_cast<const uint4*>(a2);
        const uint32_t* pb = reinterpret_cast<const uint32_t*>(b2 - (sh >> 3));
        for (uint64_t x = lane; x < body; x += G) {
            const uint4 va = __ldg(pa + x);
            const uint32_t w0 = __ldg(pb + 4 * x), w1 = __ldg(pb + 4 * x + 1), w2 = __ldg(pb + 4 * x + 2), w3 = __ldg(pb + 4 * x + 3), w4 = __ldg(pb + 4 * x + 4);
            cnt += diff4(va.x, __funnelshift_r(w0, w1, sh)) + diff4(va.y, __funnelshift_r(w1, w2, sh)) +
                   diff4(va.z, __funnelshift_r(w2, w3, sh)) + diff4(va.w, __funnelshift_r(w3, w4, sh));
        }
    }
    for (uint64_t x = head + body * 16 + lane; x < len; x += G) cnt += a[x] != b[x];     // tail < 16 bytes
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    return cnt;
}
__device__ __forceinline__ uint32_t warp_mismatches(const uint8_t* a, const uint8_t* b, uint64_t len, uint32_t lane) {
    return group_mismatches<32>(a, b, len, lane);
}



__global__ void __launch_bounds__(256) k5_hamming(const HammingArgs A, const uint64_t* piece_first /* [n_pairs + 1] */, uint64_t n_pieces) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp0 = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t piece = warp0; piece < n_pieces; piece += nwarps) {
        // pair of this piece: last p with piece_first[p] <= piece
        uint64_t lo = 0, hi = A.n_pairs;
        while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (__ldg(piece_first + mid) <= piece) lo = mid; else hi = mid; }
        const uint64_t p = lo;
        const uint64_t o0 = A.seq_off[2 * p], o1 = A.seq_off[2 * p + 1], o2 = A.seq_off[2 * p + 2];
        const uint64_t n = o1 - o0, m = o2 - o1;
        if (n != m) { if (lane == 0) atomicOr(A.err_flag, 2u); continue; }
        const uint64_t start = (piece - __ldg(piece_first + p)) * HAM_SPLIT;
        const uint64_t len = umin64(HAM_SPLIT, n - start);
        const uint32_t c = n ? warp_mismatches(A.residues + o0 + start, A.residues + o1 + start, len, lane) : 0u;
        if (lane == 0 && c) atomicAdd(reinterpret_cast<unsigned long long*>(A.out + p), (unsigned long long)c);
    }
}

// Every pair is a single piece (no pair longer than HAM_SPLIT): G lanes per pair, no search, no atomics, no
// grid-stride loop -- one group per pair, so that the block scheduler keeps every SM full of independent groups
// (the generic kernel above found a piece's pair by binary search: 22 dependent loads per warp for 4 M pairs).
template <int G>
__global__ void __launch_bounds__(256) k5_hamming_direct(const HammingArgs A) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t p = tid / G;
    const uint32_t lane = (uint32_t)(tid % G);
    uint64_t o0 = 0, n = 0, m = 0, o1 = 0;
    if (p < A.n_pairs) {
        o0 = A.seq_off[2 * p]; o1 = A.seq_off[2 * p + 1];
        n = o1 - o0; m = A.seq_off[2 * p + 2] - o1;
        if (n != m) { if (lane == 0) atomicOr(A.err_flag, 2u); n = 0; }
    }
    const uint32_t c = group_mismatches<G>(A.residues + o0, A.residues + o1, n, lane);
    if (p < A.n_pairs && lane == 0) A.out[p] = c;
}


// One warp per unordered pair (i, j), i < j: count over the zip of the two rows (zip stops at the shorter one,
// stat.rs:145), divide in f32 exactly as the reference does (IEEE division, round to nearest), write both
// (i, j) and (j, i).  The diagonal is 0 / columns.
__global__ void __launch_bounds__(256) k5_pdist(const PDistArgs A) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t R = A.rows;
    const uint64_t n_items = R * (R + 1) / 2;
    const uint64_t warp0 = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t it = warp0; it < n_items; it += nwarps) {
        // item -> (i, j) with i <= j: row i holds R - i items
        uint64_t i = (uint64_t)((2.0 * (double)R + 1.0 - sqrt((2.0 * (double)R + 1.0) * (2.0 * (double)R + 1.0) - 8.0 * (double)it)) * 0.5);
        while (i > 0 && i * R - i * (i - 1) / 2 > it) --i;
        while ((i + 1) * R - (i + 1) * i / 2 <= it) ++i;
        const uint64_t j = i + (it - (i * R - i * (i - 1) / 2));
        float v = 0.0f;
        if (i != j) {
            const uint64_t a0 = A.seq_off[i], a1 = A.seq_off[i + 1], b0 = A.seq_off[j], b1 = A.seq_off[j + 1];
            const uint64_t len = umin64(a1 - a0, b1 - b0);
            uint64_t cnt = 0;
            for (uint64_t s = 0; s < len; s += (1ull << 30)) cnt += warp_mismatches(A.residues + a0 + s, A.residues + b0 + s, umin64(1ull << 30, len - s), lane);
            v = (float)cnt;
        }
        if (lane == 0) {
            const float q = __fdiv_rn(v, A.columns);
            A.out[i * R + j] = q;
            A.out[j * R + i] = q;
        }
    }
}

}  // namespace bg
