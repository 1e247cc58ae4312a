// l_k3.cu -- instantiations of K3 (k3_walk.cuh): traceback walks, string assembly, offset scan.
#include <cub/device/device_scan.cuh>

#include "launch.h"
#include "k3_walk.cuh"

namespace bg {

// One-thread-per-pair walker with the launch's geometry compiled in.
void dispatch_walk(Shape sh, bool half, uint32_t ns, cudaStream_t st, const WalkArgs& a) {
    const dim3 grid((ns + 127) / 128);
    if (half) {
#define X(L_, C_) if (sh.L == L_ && sh.C == C_) { k3_walk<L_, C_, true><<<grid, 128, 0, st>>>(a); return; }
        BG_HALF_SHAPES(X)
#undef X
    } else {
#define X(L_, C_) if (sh.L == L_ && sh.C == C_) { k3_walk<L_, C_, false><<<grid, 128, 0, st>>>(a); return; }
        BG_SHAPES(X)
#undef X
    }
    k3_walk<0, 0, false><<<grid, 128, 0, st>>>(a);
}

// Long pairs: one warp per pair.  k2_C: WAVE_C / WAVE_C_NARROW -> the (32 lanes x k2_C columns) geometry as compile-time
// constants; 0 -> any geometry with C % 8 == 0.
void launch_long_walk(LongWalk kind, int k2_C, uint32_t ns, cudaStream_t st, const WalkArgs& a) {
    const unsigned gs = (ns + WALK_SKEW_WARPS - 1) / WALK_SKEW_WARPS, gd = (ns + WALK_DIAG_WARPS - 1) / WALK_DIAG_WARPS;
    switch (kind) {
    case LW_SKEW:
        if (k2_C == WAVE_C) k3_walk_skew<WAVE_C><<<gs, WALK_SKEW_WARPS * 32, 0, st>>>(a);
        else if (k2_C == WAVE_C_NARROW) k3_walk_skew<WAVE_C_NARROW><<<gs, WALK_SKEW_WARPS * 32, 0, st>>>(a);
        else k3_walk_skew<0><<<gs, WALK_SKEW_WARPS * 32, 0, st>>>(a);
        break;
    case LW_DIAG:
        if (k2_C == WAVE_C) k3_walk_diag<WAVE_C><<<gd, WALK_DIAG_WARPS * 32, 0, st>>>(a);
        else k3_walk_diag<0><<<gd, WALK_DIAG_WARPS * 32, 0, st>>>(a);
        break;
    }
}

void launch_scores_only(const PairDesc* desc, const EndCell* end, uint32_t ns, int32_t* score, uint8_t* flags, int mode, cudaStream_t st) {
    k_scores_only<<<(ns + 127) / 128, 128, 0, st>>>(desc, end, ns, score, flags, mode);
}
void launch_gather(const GatherArgs& a, cudaStream_t st) { k_gather<<<(unsigned)((a.n_slots + 3) / 4), 128, 0, st>>>(a); }
void launch_ops_counts(const uint64_t* lens2, uint64_t n_pairs, ulonglong2* counts, cudaStream_t st) {
    k_ops_counts<<<(unsigned)((n_pairs + 1 + 255) / 256), 256, 0, st>>>(lens2, n_pairs, counts);
}
struct AddU2 {
    __host__ __device__ ulonglong2 operator()(const ulonglong2& a, const ulonglong2& b) const { return make_ulonglong2(a.x + b.x, a.y + b.y); }
};
cudaError_t scan_counts(void* tmp, size_t& tmp_bytes, const ulonglong2* counts, ulonglong2* out, int count, cudaStream_t st) {
    return cub::DeviceScan::ExclusiveScan(tmp, tmp_bytes, counts, out, AddU2(), make_ulonglong2(0ull, 0ull), count, st);
}
void launch_ops_sample(const ulonglong2* scan, uint64_t n_pairs, uint64_t stride, ulonglong2* samples, cudaStream_t st) {
    const uint64_t nsub = (n_pairs + stride - 1) / stride;
    k_ops_sample<<<(unsigned)((nsub + 1 + 255) / 256), 256, 0, st>>>(scan, n_pairs, stride, samples);
}
void launch_pack_ops(const PackOpsArgs& a, bool long_pairs, cudaStream_t st) {
    if (long_pairs) k_pack_ops<32><<<(unsigned)(((uint64_t)a.n_slots * 32 + 127) / 128), 128, 0, st>>>(a);
    else k_pack_ops<8><<<(unsigned)(((uint64_t)a.n_slots * 8 + 127) / 128), 128, 0, st>>>(a);
}
void launch_rebase(uint64_t* off, uint64_t count, const uint64_t* base, cudaStream_t st) {
    k_rebase<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(off, count, base);
}
void launch_bump(uint64_t* base, const uint64_t* chunk_total_entry, uint64_t* chunk_total_out, cudaStream_t st) {
    k_bump<<<1, 1, 0, st>>>(base, chunk_total_entry, chunk_total_out);
}
// Exclusive scan of the aligned lengths (tmp == nullptr: size query)
cudaError_t scan_lengths(void* tmp, size_t& tmp_bytes, const uint64_t* lens, uint64_t* off, int count, cudaStream_t st) {
    return cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, lens, off, count, st);
}

}  // namespace bg
