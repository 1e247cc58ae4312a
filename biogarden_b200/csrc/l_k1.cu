// l_k1.cu -- instantiations of K1 (k1_fill.cuh), global-flavour recurrence.
#include "launch.h"
#include "k1_fill.cuh"

namespace bg {

void dispatch_k1_local(Shape sh, bool prof4, dim3 grid, size_t smem, cudaStream_t st, const FillArgs& a);

void dispatch_k1(Shape sh, bool local, bool prof4, dim3 grid, size_t smem, cudaStream_t st, const FillArgs& a) {
    if (local) { dispatch_k1_local(sh, prof4, grid, smem, st, a); return; }
#define X(L_, C_) if (sh.L == L_ && sh.C == C_) { \
        if (prof4) k1_fill<L_, C_, false, true><<<grid, 128, smem, st>>>(a); \
        else k1_fill<L_, C_, false, false><<<grid, 128, smem, st>>>(a); \
        return; }
    BG_SHAPES(X)
#undef X
}

}  // namespace bg
