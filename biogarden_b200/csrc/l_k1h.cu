// l_k1h.cu -- instantiations of K1h (k1h_fill.cuh), the packed 16 x 2 fill, score pairs from the 256-entry table
// in shared memory (any |score| up to HB_MAXABS).  The byte-profile form lives in l_k1hp.cu.
#include "launch.h"
#include "k1h_fill.cuh"

namespace bg {

bool dispatch_k1hp(Shape sh, bool track, dim3 grid, cudaStream_t st, const FillArgs& a);

// Blocks per SM by columns per lane (registers ~ 4 C + 60; measured on cfg2: (16,10) with 6 blocks/SM and
// 4 of the 8 accumulations on the ALU pipe fills 11 % faster than (8,19) with 3 blocks/SM).
constexpr int k1h_minb(int C) { return C <= 10 ? 6 : C <= 12 ? 4 : C <= 24 ? 3 : 2; }

bool dispatch_k1h(Shape sh, bool track, bool prof8, dim3 grid, cudaStream_t st, const FillArgs& a) {
    if (prof8) return dispatch_k1hp(sh, track, grid, st, a);
#define X(L_, C_) if (sh.L == L_ && sh.C == C_) { \
        if (track) k1h_fill<L_, C_, true, 0x55, k1h_minb(C_), false><<<grid, 128, 0, st>>>(a); \
        else k1h_fill<L_, C_, false, 0x55, k1h_minb(C_), false><<<grid, 128, 0, st>>>(a); \
        return true; }
    BG_HALF_SHAPES(X)
#undef X
    return false;
}

}  // namespace bg
