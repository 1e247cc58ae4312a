// l_k1hp.cu -- instantiations of K1h (k1h_fill.cuh) with the score pair built by one PRMT from the row's byte
// profiles (every s - a - b fits a signed byte): no shared-memory load and no address add per cell pair.
#include <cstdlib>

#include "launch.h"
#include "k1h_fill.cuh"

namespace bg {

// blocks per SM: without the end-cell scans 19 / 20 columns per lane fit 128 registers (4 blocks).  (More than 16 warps per SM
// would need <= 96 registers whatever the block size: each SM sub-partition holds 16 384 registers, i.e. four warps of 116.)
constexpr int k1hp_minb(int C, bool track) { return C <= 10 ? 6 : C <= 12 ? 4 : (C <= 20 && !track) ? 4 : C <= 24 ? 3 : 2; }
// Pipe split of the 8 trace-bit accumulations per cell pair (HB_PIPES, k1h_fill.cuh).  The ALU pipe already carries
// 4 VIMNMX + PRMT + VIADD.16x2, the FMA pipe one IMAD.  Measured on cfg2 (10^6 pairs of 150 bp, fill time; build with
// -DBG_HBP_SWEEP and set BG_HBP_PIPES to repeat it):
//   (8, 19) at 4 blocks/SM: 0x00 7.87 ms, 0x05 7.49, 0x15 7.43, 0x55 7.41, 0x44 7.40, 0x45 7.35, 0xD5 7.57, 0xFF 8.34
//   (16, 10) at 6 blocks/SM: 0x00 9.20, 0x05 8.56, 0x15 8.34, 0x55 8.31, 0x57 8.45, 0x5F 8.72, 0xFF 9.49
// Tried and dropped on (8, 19): the diagonal add as an IMAD on the FMA pipe (valid when every s - a - b >= 0): best split
// 7.46 ms; M^ + (a - b) as VIADD.16x2 on the ALU pipe: best 7.78 ms -- it sits on the cell-to-cell dependency chain.
constexpr int k1hp_pipes(int C) { return C >= 16 ? 0x45 : 0x55; }

bool dispatch_k1hp(Shape sh, bool track, dim3 grid, cudaStream_t st, const FillArgs& a) {
#ifdef BG_HBP_SWEEP
    static const int sweep = [] { const char* e = getenv("BG_HBP_PIPES"); return e ? (int)strtol(e, nullptr, 0) : -1; }();
    if (!track && sweep >= 0) {
#define V(P_) if (sweep == P_) { \
        if (sh.L == 16 && sh.C == 10) { k1h_fill<16, 10, false, P_, 6, true><<<grid, 128, 0, st>>>(a); return true; } \
        if (sh.L == 8 && sh.C == 19) { k1h_fill<8, 19, false, P_, 4, true><<<grid, 128, 0, st>>>(a); return true; } \
        }
        V(0x00) V(0x05) V(0x15) V(0x44) V(0x45) V(0x55) V(0x57) V(0x5F) V(0xD5) V(0xFF)
#undef V
    }
#endif
#define X(L_, C_) if (sh.L == L_ && sh.C == C_) { \
        if (track) k1h_fill<L_, C_, true, k1hp_pipes(C_), k1hp_minb(C_, true), true><<<grid, 128, 0, st>>>(a); \
        else k1h_fill<L_, C_, false, k1hp_pipes(C_), k1hp_minb(C_, false), true><<<grid, 128, 0, st>>>(a); \
        return true; }
    BG_HALF_SHAPES(X)
#undef X
    return false;
}

}  // namespace bg
