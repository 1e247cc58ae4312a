// pack_util.h -- reading packed residues (bg_batch::packing), host and device.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define BG_HD __host__ __device__
#else
#define BG_HD
#endif

namespace bg {

// code of residue i of a packed arena (bits = 2 or 5); never reads a byte that holds no bit of the residue
BG_HD inline uint32_t packed_code(const uint8_t* packed, uint32_t bits, uint64_t i) {
    const uint64_t bit = (uint64_t)bits * i;
    const uint32_t sh = (uint32_t)(bit & 7);
    uint32_t w = packed[bit >> 3];
    if (sh + bits > 8) w |= (uint32_t)packed[(bit >> 3) + 1] << 8;
    return (w >> sh) & ((1u << bits) - 1u);
}
// bytes [b0, b1) of the arena hold residues [r0, r1)
BG_HD inline void packed_byte_range(uint32_t bits, uint64_t r0, uint64_t r1, uint64_t& b0, uint64_t& b1) {
    b0 = (bits * r0) >> 3;
    b1 = (bits * r1 + 7) >> 3;
    if (b1 < b0) b1 = b0;
}

}  // namespace bg
