// l_k45.cu -- instantiations of K4 / K4b (k4_edit.cuh: edit distance) and K5 (k5_distance.cuh: hamming, p-distance).
#include "launch.h"
#include "k4_edit.cuh"
#include "k5_distance.cuh"

namespace bg {

void dispatch_k4(Shape sh, dim3 grid, cudaStream_t st, const EditArgs& a) {
#define X(L_, C_) if (sh.L == L_ && sh.C == C_) { k4_edit<L_, C_><<<grid, 128, 0, st>>>(a); return; }
    BG_SHAPES(X)
#undef X
}
void launch_myers(int W, uint32_t ns, cudaStream_t st, const MyersArgs& a) {
    const unsigned grid = (ns + 127) / 128;
    if (W == 4) k4_myers<4><<<grid, 128, 0, st>>>(a);
    else if (W == 8) k4_myers<8><<<grid, 128, 0, st>>>(a);
    else k4_myers<10><<<grid, 128, 0, st>>>(a);
}
void launch_byte_hist(const uint8_t* data, uint64_t n, unsigned int* hist, int blocks, cudaStream_t st) {
    k_byte_hist<<<blocks, 256, 0, st>>>(data, n, hist);
}
void launch_hamming_direct(int group, uint64_t n_pairs, cudaStream_t st, const HammingArgs& a) {
    if (group == 8) k5_hamming_direct<8><<<(unsigned)((n_pairs * 8 + 255) / 256), 256, 0, st>>>(a);
    else if (group == 16) k5_hamming_direct<16><<<(unsigned)((n_pairs * 16 + 255) / 256), 256, 0, st>>>(a);
    else k5_hamming_direct<32><<<(unsigned)((n_pairs * 32 + 255) / 256), 256, 0, st>>>(a);
}
void launch_hamming_pieces(unsigned blocks, cudaStream_t st, const HammingArgs& a, const uint64_t* piece_first, uint64_t pieces) {
    k5_hamming<<<blocks, 256, 0, st>>>(a, piece_first, pieces);
}
void launch_pdist(unsigned blocks, cudaStream_t st, const PDistArgs& a) { k5_pdist<<<blocks, 256, 0, st>>>(a); }

}  // namespace bg
