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

// K1 / K1h / K2 fills
void dispatch_k1(Shape sh, bool local, bool prof4, dim3 grid, size_t smem, cudaStream_t st, const FillArgs& a);
bool dispatch_k1h(Shape sh, bool track, dim3 grid, cudaStream_t st, const FillArgs& a);
cudaError_t launch_k2(bool local, bool prof4, int n_cta, size_t smem, cudaStream_t st, const WaveArgs& a, bool ckpt = false);

// K3 walks + string assembly
void dispatch_walk(Shape sh, bool half, uint32_t ns, cudaStream_t st, const WalkArgs& a);
enum LongWalk { LW_SKEW = 0, LW_DIAG = 1, LW_TILE = 2, LW_WARP = 3 };
void launch_long_walk(LongWalk kind, bool k2_geometry, uint32_t ns, cudaStream_t st, const WalkArgs& a);
void launch_scores_only(const PairDesc* desc, const EndCell* end, uint32_t ns, int32_t* score, uint8_t* flags, int mode, cudaStream_t st);
void launch_gather(const GatherArgs& a, cudaStream_t st);
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
