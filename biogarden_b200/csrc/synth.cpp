// synth.cpp -- deterministic synthetic workloads (SURVEY 8d generator).  Host code, no CUDA.
//
// Every pair has its own splitmix64 stream derived from (seed, pair index), so any rank can
// generate any sub-range of a workload and get the same bytes as a single-process run.
#include "../../include/bgsynth.h"

#include <algorithm>
#include <thread>
#include <vector>

namespace {

struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    inline uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    inline uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    inline double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

inline uint64_t pair_seed(uint64_t seed, uint64_t pair) {
    SplitMix64 k(seed * 0xD1B54A32D192ED03ull + 0x8CB92BA72F3D8DD7ull);
    return k.next() ^ (pair * 0x9E3779B97F4A7C15ull + 0x2545F4914F6CDD1Dull);
}

// Generates one pair; writes into a / b when non-null.  Returns lengths.
inline void gen_pair(uint64_t seed, uint64_t pair, const char* alpha, uint32_t na, uint32_t lo, uint32_t hi, int resize_b,
                     std::vector<uint8_t>& a, std::vector<uint8_t>& b) {
    SplitMix64 r(pair_seed(seed, pair));
    const uint32_t la = lo + r.below(hi - lo + 1);
    const uint32_t lb_target = lo + r.below(hi - lo + 1);
    const bool related = r.below(10) != 0;   // 90 % mutated copies, 10 % independent
    a.resize(la);
    for (uint32_t i = 0; i < la; ++i) a[i] = (uint8_t)alpha[r.below(na)];
    b.clear();
    if (related) {
        for (uint32_t i = 0; i < la; ++i) {
            const double u = r.unit();
            if (u < 0.05) b.push_back((uint8_t)alpha[r.below(na)]);                       // substitution
            else if (u < 0.06) { b.push_back((uint8_t)alpha[r.below(na)]); b.push_back(a[i]); }   // insertion
            else if (u < 0.07) { /* deletion */ }
            else b.push_back(a[i]);
        }
        if (resize_b) {
            if (b.size() > lb_target) b.resize(lb_target);
            while (b.size() < lb_target) b.push_back((uint8_t)alpha[r.below(na)]);
        }
    } else {
        b.resize(lb_target);
        for (uint32_t i = 0; i < lb_target; ++i) b[i] = (uint8_t)alpha[r.below(na)];
    }
}

}  // namespace

extern "C" int bg_synth_pairs(uint64_t seed, uint64_t first_pair, uint64_t n_pairs, const char* alphabet, int alphabet_len,
                              uint32_t len_lo, uint32_t len_hi, int resize_b,
                              uint8_t* residues, uint64_t* seq_off, uint64_t* n_residues) {
    if (!alphabet || alphabet_len <= 0 || len_hi < len_lo || !seq_off) return 5;   /* BG_EINVAL_ARG */
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    nt = (unsigned)std::min<uint64_t>(nt, std::max<uint64_t>(1, n_pairs / 4096));
    // pass 1: lengths
    std::vector<uint32_t> la(n_pairs), lb(n_pairs);
    auto pass = [&](bool fill) {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) {
            const uint64_t lo = n_pairs * t / nt, hi = n_pairs * (t + 1) / nt;
            th.emplace_back([&, lo, hi, fill] {
                std::vector<uint8_t> a, b;
                for (uint64_t p = lo; p < hi; ++p) {
                    gen_pair(seed, first_pair + p, alphabet, (uint32_t)alphabet_len, len_lo, len_hi, resize_b, a, b);
                    if (!fill) { la[p] = (uint32_t)a.size(); lb[p] = (uint32_t)b.size(); }
                    else {
                        std::copy(a.begin(), a.end(), residues + seq_off[2 * p]);
                        std::copy(b.begin(), b.end(), residues + seq_off[2 * p + 1]);
                    }
                }
            });
        }
        for (auto& x : th) x.join();
    };
    pass(false);
    seq_off[0] = 0;
    for (uint64_t p = 0; p < n_pairs; ++p) {
        seq_off[2 * p + 1] = seq_off[2 * p] + la[p];
        seq_off[2 * p + 2] = seq_off[2 * p + 1] + lb[p];
    }
    if (n_residues) *n_residues = seq_off[2 * n_pairs];
    if (residues) pass(true);
    return 0;
}
