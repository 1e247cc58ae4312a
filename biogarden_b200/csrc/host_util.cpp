// host_util.cpp -- host-side helpers of the C ABI that need no CUDA (include/bgalign.h "helpers for host mirrors").
#include "../../include/bgalign.h"

#include <algorithm>
#include <thread>
#include <vector>

extern "C" int bg_residue_histogram(const bg_batch* in, uint64_t* hist_a, uint64_t* hist_b) {
    if (!in || !hist_a || !hist_b || (in->n_pairs && (!in->seq_off || !in->residues))) return BG_EINVAL_ARG;
    for (int i = 0; i < 256; ++i) hist_a[i] = hist_b[i] = 0;
    const uint64_t N = in->n_pairs;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    nt = (unsigned)std::min<uint64_t>(nt, std::max<uint64_t>(1, N / 8192));
    std::vector<std::vector<uint64_t>> part(nt, std::vector<uint64_t>(512, 0));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) {
        const uint64_t lo = N * t / nt, hi = N * (t + 1) / nt;
        th.emplace_back([&, t, lo, hi] {
            uint64_t* h = part[t].data();
            for (uint64_t p = lo; p < hi; ++p) {
                for (uint64_t x = in->seq_off[2 * p]; x < in->seq_off[2 * p + 1]; ++x) h[in->residues[x]]++;
                for (uint64_t x = in->seq_off[2 * p + 1]; x < in->seq_off[2 * p + 2]; ++x) h[256 + in->residues[x]]++;
            }
        });
    }
    for (auto& x : th) x.join();
    for (unsigned t = 0; t < nt; ++t)
        for (int i = 0; i < 256; ++i) { hist_a[i] += part[t][i]; hist_b[i] += part[t][256 + i]; }
    return BG_OK;
}
