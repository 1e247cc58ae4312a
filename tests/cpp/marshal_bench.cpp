// marshal_bench.cpp -- what a host shim pays around the C ABI: the C++ mirror's
// Tile -> align_batch -> Vec<(score, Sequence, Sequence)> on config #2's shape, split into
//   marshal in  : Tile (one heap vector per Sequence) -> residue arena + offsets (what bg_batch wants)
//   engine      : bg_align_batch_ops (host buffers in, compact results out)
//   marshal out : one Sequence pair per alignment, expanded straight from the ops (bg_expand_ops), all host threads
// The Rust shim (rust/biogarden-gpu) does the same three steps; it cannot be built in this image.
//   g++ -std=c++17 -O2 -pthread -Iinclude tests/cpp/marshal_bench.cpp -Lbiogarden_b200 -lbgalign -Wl,-rpath,$PWD/biogarden_b200 -o tests/cpp/marshal_bench
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "biogarden.hpp"

using namespace biogarden;
using clk = std::chrono::steady_clock;
static double ms(clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); }

int main(int argc, char** argv) {
    const size_t n_pairs = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1000000;
    std::mt19937_64 rng(2);
    ds::Tile tile;
    tile.data.reserve(2 * n_pairs);
    const char alpha[] = "ACGT";
    for (size_t p = 0; p < n_pairs; ++p) {
        std::vector<uint8_t> a(150), b(150);
        for (auto& c : a) c = (uint8_t)alpha[rng() & 3];
        for (size_t i = 0; i < 150; ++i) b[i] = (rng() % 100 < 6) ? (uint8_t)alpha[rng() & 3] : a[i];
        tile.push(ds::Sequence(std::move(a))); tile.push(ds::Sequence(std::move(b)));
    }
    alignment::aligner::SequenceAligner al;
    double best_total = 1e30, in_ms = 0, eng_ms = 0, out_ms = 0;
    for (int rep = 0; rep < 5; ++rep) {
        const auto t0 = clk::now();
        std::vector<uint8_t> res; std::vector<uint64_t> off{0};
        res.reserve(300 * n_pairs); off.reserve(2 * n_pairs + 1);
        for (const auto& s : tile.data) { res.insert(res.end(), s.chain.begin(), s.chain.end()); off.push_back(res.size()); }
        const auto t1 = clk::now();
        (void)t1;
        // (the mirror's align_batch repeats the marshalling; time the whole call and subtract)
        const auto t2 = clk::now();
        auto out = al.align_batch(tile, BG_GLOBAL, alignment::score::unit, -2, -1);
        const auto t3 = clk::now();
        // engine alone on the same buffers
        bg_batch batch{n_pairs, res.data(), off.data()};
        std::array<uint8_t, 256> rc, cc; rc.fill(0xFF); cc.fill(0xFF);
        int32_t table[16]; int k = 0;
        for (char x : {'A', 'C', 'G', 'T'}) { rc[(uint8_t)x] = cc[(uint8_t)x] = (uint8_t)k++; }
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) table[i * 4 + j] = i == j ? 1 : -1;
        bg_params prm{BG_GLOBAL, -2, -1, 0u, table, 4, 4, rc.data(), cc.data()};
        bg_ops_result r{};
        const auto t4 = clk::now();
        if (bg_align_batch_ops(al.context(), &batch, &prm, &r) != BG_OK) { fprintf(stderr, "engine error\n"); return 1; }
        const auto t5 = clk::now();
        bg_ops_result_free(&r);
        const double total = ms(t2, t3);
        if (total < best_total) { best_total = total; in_ms = ms(t0, t1); eng_ms = ms(t4, t5); out_ms = total - in_ms - eng_ms; }
        if (std::get<1>(out[0]).len() < 150) return 2;
    }
    printf("{\"pairs\": %zu, \"tile_to_results_ms\": %.2f, \"marshal_in_ms\": %.2f, \"engine_ops_call_ms\": %.2f, \"expand_into_sequences_ms\": %.2f, "
           "\"note\": \"C++ mirror (include/biogarden.hpp): pageable std::vector arena, one heap vector per Sequence, expansion on all host threads; expand = total - in - engine\"}\n",
           n_pairs, best_total, in_ms, eng_ms, out_ms);
    return 0;
}
