// k0_unpack.cuh -- packed residues (bg_batch::packing) -> one byte per residue in HBM, right after the H2D copy.
//
// The fill kernels translate residue BYTES through 256-entry code maps in shared memory (they read every residue
// once per band: < 2 % of their traffic), so packed input is unpacked once, on the device, instead of giving every
// kernel a second residue path: 2-bit DNA is read with 128-bit loads (64 residues per load) and written as 4 x 128 bit;
// 0.06 ms per 10^6 pairs of 150 bp against 10 ms of alignment.  What packing buys is the host link: 75 instead of
// 300 bytes per pair H2D.
#pragma once
#include "bg_args.cuh"
#include "pack_util.h"

namespace bg {

// 64 residues per thread.  packed16: the packed bytes, 16-byte aligned; bit0: bit offset of residue 0 in it (even).
__global__ void __launch_bounds__(256) k_unpack2(const UnpackArgs A) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t r = t * 64;
    if (r >= A.count) return;
    const uint64_t q = A.bit0 + 2 * r;
    const uint32_t sh = (uint32_t)(q & 31);
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(A.packed) + (q >> 5);
    uint32_t w[5];
    if ((reinterpret_cast<uintptr_t>(w32) & 15) == 0) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(w32));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = __ldg(w32 + k);
    }
    w[4] = sh ? __ldg(w32 + 4) : 0u;            // (the staging buffer is padded by 16 bytes)
    const uint32_t alpha = (uint32_t)A.alphabet[0] | ((uint32_t)A.alphabet[1] << 8) | ((uint32_t)A.alphabet[2] << 16) | ((uint32_t)A.alphabet[3] << 24);
    uint8_t* out = A.out + r;
    const bool full = r + 64 <= A.count && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t x = __funnelshift_r(w[k], w[k + 1], sh);      // 16 codes
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t b = (x >> (8 * j)) & 0xffu;
            const uint32_t sel = (b & 3u) | (((b >> 2) & 3u) << 4) | (((b >> 4) & 3u) << 8) | (((b >> 6) & 3u) << 12);
            o[j] = __byte_perm(alpha, 0u, sel);
        }
        if (full) reinterpret_cast<uint4*>(out)[k] = make_uint4(o[0], o[1], o[2], o[3]);
        else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (r + 16 * k + j < A.count) out[16 * k + j] = (uint8_t)(o[j >> 2] >> (8 * (j & 3)));
        }
    }
}

// any packing, one residue per thread
__global__ void __launch_bounds__(256) k_unpack_any(const UnpackArgs A) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= A.count) return;
    const uint64_t q = A.bit0 + (uint64_t)A.bits * r;
    const uint32_t sh = (uint32_t)(q & 7);
    uint32_t w = A.packed[q >> 3];
    if (sh + A.bits > 8) w |= (uint32_t)A.packed[(q >> 3) + 1] << 8;
    A.out[r] = A.alphabet[(w >> sh) & ((1u << A.bits) - 1u)];
}

}  // namespace bg
