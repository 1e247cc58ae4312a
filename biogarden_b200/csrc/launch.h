// launch.h -- host-callable launchers of the kernel families.  Each family is instantiated in its own
// translation unit (l_*.cu) so that the library builds in parallel; bg_api.cu only sees these declarations.
#pragma once
#include <cuda_runtime.h>

#include "bg_args.cuh"

namespace bg {

struct Shape { int L, C; };

// Kernel shapes compiled in: (lanes per pair, columns per lane).  A band is L*C columns.
#define BG_SHAPES(X) \
    X(8, 8) X(8, 12) X(8, 16) X(8, 19) X(8, 24) X(16, 10) X(16, 16) X(32, 5) X(32, 8) X(32, 12) X(32, 16) X(32, 20) X(32, 24) X(32, 32)
// ... of which have a packed 16 x 2 instantiation (K1h)
#define BG_HALF_SHAPES(X) X(8, 8) X(8, 12) X(8, 16) X(8, 19) X(8, 24) X(16, 10) X(16, 16) X(32, 12) X(32, 16) X(32, 20) X(32, 24) X(32, 32)
constexpr int WAVE_C = 16;   // columns per lane of the K2 wavefront kernel (bands of 32 * WAVE_C columns)
constexpr int WAVE_C_NARROW = 8;   // ... for launches with so few bands that most SMs would idle (a lone 10 kbp pair)
constexpr int BG_N_SHAPES = 14;
constexpr uint64_t LONG_WALK_LEN = 16384;   // len1 + len2 above which a pair is walked by a warp (k3_walk_skew) instead of a thread

// Shape number si (its position in BG_SHAPES).
__host__ __device__ inline Shape shape_at(int si) {
    int i = 0;
#define X(L_, C_) if (si == i) return Shape{L_, C_}; ++i;
    BG_SHAPES(X)
#undef X
    return Shape{0, 0};
}
__host__ __device__ inline int shape_index(Shape s) {
    int i = 0;
#define X(L_, C_) if (s.L == L_ && s.C == C_) return i; ++i;
    BG_SHAPES(X)
#undef X
    return -1;
}
__host__ __device__ inline bool shape_has_half(Shape s) {
#define X(L_, C_) if (s.L == L_ && s.C == C_) return true;
    BG_HALF_SHAPES(X)
#undef X
    return false;
}

// Length class -> kernel shape (the same function on the host and in the device-side planner, k0_plan.cuh).
// Short pairs use few lanes per pair (the systolic pipeline costs L-1 fill/drain steps per pair) and many columns
// per lane; wide pairs use a full warp, and pairs wider than 1024 columns loop over bands of the L=32 shape that
// wastes the fewest padded columns.  half_ok: the packed 16 x 2 kernel is available for the call (16 lanes x 10
// columns at 6 blocks/SM then beats 8 x 24 at 3 for 153-160 columns; up to 152 columns 8 x 19 at 4 blocks/SM is fastest).  c8: only shapes with C % 8 == 0 (an eight-column block is one trace
// word, which the long-pair walker's window loader relies on) -- for pairs with len1 + len2 > LONG_WALK_LEN.
__host__ __device__ inline Shape pick_shape_m(uint32_t m, bool half_ok, bool c8) {
    if (c8) {
        if (m <= 64) return Shape{8, 8};
        if (m <= 128) return Shape{8, 16};
        if (m <= 192) return Shape{8, 24};
        if (m <= 256) return Shape{16, 16};
        if (m <= 512) return Shape{32, 16};
        if (m <= 768) return Shape{32, 24};
        if (m <= 1024) return Shape{32, 32};
    } else {
        if (half_ok && m > 152 && m <= 160) return Shape{16, 10};
        if (m <= 64) return Shape{8, 8};
        if (m <= 96) return Shape{8, 12};
        if (m <= 128) return Shape{8, 16};
        if (m <= 152) return Shape{8, 19};
        if (m <= 192) return Shape{8, 24};
        if (m <= 256) return Shape{16, 16};
        if (m <= 384) return Shape{32, 12};
        if (m <= 512) return Shape{32, 16};
        if (m <= 640) return Shape{32, 20};
        if (m <= 768) return Shape{32, 24};
        if (m <= 1024) return Shape{32, 32};
    }
    Shape best{32, 32};
    uint64_t best_cols = ~0ull;
    const int cs[4] = {32, 24, 20, 16};
    for (int k = 0; k < 4; ++k) {
        if (c8 && (cs[k] & 7)) continue;
        const uint64_t band = 32ull * cs[k], cols = (m + band - 1) / band * band;
        if (cols < best_cols) { best_cols = cols; best = Shape{32, cs[k]}; }
    }
    return best;
}

// K0: launch descriptors built on the device (k0_plan.cuh)
size_t plan_scratch_bytes(uint32_t n_pairs, uint32_t n_slots);
cudaError_t launch_plan(PlanArgs a, bool sort, uint32_t n_slots, void* scratch, size_t scratch_bytes, cudaStream_t st);

size_t edit_plan_scratch_bytes(uint32_t n_pairs);
cudaError_t launch_edit_plan(EditPlanArgs a, void* scratch, size_t scratch_bytes, uint32_t* err_flag, cudaStream_t st);   // also clears *err_flag   // K0e (k0_eplan.cuh)

void launch_unpack(const UnpackArgs& a, cudaStream_t st);

// K1 / K1h / K2 fills
void dispatch_k1(Shape sh, bool local, bool prof4, dim3 grid, size_t smem, cudaStream_t st, const FillArgs& a);
bool dispatch_k1h(Shape sh, bool track, bool prof8, dim3 grid, cudaStream_t st, const FillArgs& a);
cudaError_t launch_k2(bool local, bool prof4, int C, int n_cta, int warps_per_cta, size_t smem, cudaStream_t st, const WaveArgs& a, bool ckpt = false);

cudaError_t launch_k2f(bool local, bool prof4, int n_cta, int warps_per_cta, size_t smem, cudaStream_t st, const FineArgs& a);

// K3 walks + string assembly
void dispatch_walk(Shape sh, bool half, uint32_t ns, cudaStream_t st, const WalkArgs& a);
enum LongWalk { LW_SKEW = 0, LW_DIAG = 1 };
void launch_long_walk(LongWalk kind, int k2_C, uint32_t ns, cudaStream_t st, const WalkArgs& a);   // k2_C: 16 / 8 = K2 geometry compiled in, 0 = generic
void launch_scores_only(const PairDesc* desc, const EndCell* end, uint32_t ns, int32_t* score, uint8_t* flags, int mode, cudaStream_t st);
void launch_gather(const GatherArgs& a, cudaStream_t st);
void launch_ops_counts(const uint64_t* lens2, uint64_t n_pairs, ulonglong2* counts, cudaStream_t st);
cudaError_t scan_counts(void* tmp, size_t& tmp_bytes, const ulonglong2* counts, ulonglong2* out, int count, cudaStream_t st);
void launch_ops_sample(const ulonglong2* scan, uint64_t n_pairs, uint64_t stride, ulonglong2* samples, cudaStream_t st);
void launch_pack_ops(const PackOpsArgs& a, bool long_pairs, cudaStream_t st);
void launch_rebase(uint64_t* off, uint64_t count, const uint64_t* base, cudaStream_t st);
void launch_bump(uint64_t* base, const uint64_t* chunk_total_entry, uint64_t* chunk_total_out, cudaStream_t st);
cudaError_t scan_lengths(void* tmp, size_t& tmp_bytes, const uint64_t* lens, uint64_t* off, int count, cudaStream_t st);

// K4 edit distance, K5 position-wise compares
void dispatch_k4(Shape sh, dim3 grid, cudaStream_t st, const EditArgs& a);
void launch_myers(int W, uint32_t ns, cudaStream_t st, const MyersArgs& a);
void launch_byte_hist(const uint8_t* data, uint64_t n, unsigned int* hist, int blocks, cudaStream_t st);
void launch_hamming_direct(int group, uint64_t n_pairs, cudaStream_t st, const HammingArgs& a);
void launch_hamming_pieces(unsigned blocks, cudaStream_t st, const HammingArgs& a, const uint64_t* piece_first, uint64_t pieces);
void launch_pdist(unsigned blocks, cudaStream_t st, const PDistArgs& a);

}  // namespace bg
