// l_k1l.cu -- instantiations of K1 (k1_fill.cuh), local-flavour recurrence (aligner.rs:471-509).
#include "launch.h"
#include "k1_fill.cuh"

namespace bg {

void dispatch_k1_local(Shape sh, bool prof4, dim3 grid, size_t smem, cudaStream_t st, const FillArgs& a) {
#define X(L_, C_) if (sh.L == L_ && sh.C == C_) { \
        if (prof4) k1_fill<L_, C_, true, true><<<grid, 128, smem, st>>>(a); \
        else k1_fill<L_, C_, true, false><<<grid, 128, smem, st>>>(a); \
        return; }
    BG_SHAPES(X)
#undef X
}

}  // namespace bg
