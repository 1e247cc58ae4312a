// biogarden.hpp -- header-only C++ mirror of the reference's alignment interface over the C ABI
// (include/bgalign.h).  Same names, argument meaning and error behaviour as the Rust crate:
//
//   biogarden::ds::Sequence                      src/ds/sequence.rs:10-13
//   biogarden::ds::Tile                          src/ds/tile.rs:9-11
//   biogarden::BioError                          src/error.rs:8-14
//   biogarden::alignment::score::{blosum62,pam250,unit}   src/alignment/score.rs:38,78,114
//   biogarden::alignment::aligner::SequenceAligner        src/alignment/aligner.rs:28-609
//       global_alignment / local_alignment / fitting_alignment / overlap_alignment /
//       semiglobal_alignment(seq1, seq2, score, a, b) -> (score, a_align, b_align)
//       + align_batch(tile, mode, score, a, b)              (new: pair p = (tile[2p], tile[2p+1]))
//   biogarden::analysis::seq::edit_distance / edit_distance_batch   src/analysis/seq.rs:105-130
//
// Where the reference returns Err(BioError::X) this mirror throws BioError{X}; where the reference
// panics or never returns it throws BioError{ReferenceUndefined}.  Link with
// -lbgalign (biogarden_b200/libbgalign.so).  No CPU fallback exists.
#ifndef BIOGARDEN_HPP
#define BIOGARDEN_HPP

#include <algorithm>
#include <array>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#include "bgalign.h"

namespace biogarden {

struct BioError : std::runtime_error {
    enum Kind { InvalidInputSize, InvalidArgumentRange, ReferenceUndefined, Engine };
    Kind kind;
    BioError(Kind k, const std::string& what) : std::runtime_error(what), kind(k) {}
};

namespace ds {

struct Sequence {
    std::vector<uint8_t> chain;
    std::string id;   // empty = None
    Sequence() = default;
    Sequence(const char* s) : chain(s, s + std::char_traits<char>::length(s)) {}
    Sequence(const std::string& s) : chain(s.begin(), s.end()) {}
    Sequence(std::vector<uint8_t> v) : chain(std::move(v)) {}
    void push(uint8_t x) { chain.push_back(x); }
    size_t len() const { return chain.size(); }
    bool is_empty() const { return chain.empty(); }
    void reverse() { std::reverse(chain.begin(), chain.end()); }
    uint8_t operator[](size_t i) const { return chain.at(i); }
    bool operator==(const Sequence& o) const { return chain == o.chain; }   // id ignored, sequence.rs:113-117
    std::string str() const { return std::string(chain.begin(), chain.end()); }
};

struct Tile {
    std::vector<Sequence> data;
    void push(Sequence s) { data.push_back(std::move(s)); }
    size_t len() const { return data.size(); }
    bool is_empty() const { return data.empty(); }
    const Sequence& operator[](size_t i) const { return data.at(i); }
};

}  // namespace ds

namespace alignment {

using ScoreFn = std::function<int32_t(uint8_t, uint8_t)>;

namespace score {
inline int32_t table26(const char* name, uint8_t a, uint8_t b) {
    const int ia = (int)a - 65, ib = (int)b - 65;
    if (ia < 0 || ia >= 26 || ib < 0 || ib >= 26)   // the reference panics here (score.rs:40)
        throw BioError(BioError::ReferenceUndefined, std::string("score::") + name + " indexes out of bounds");
    return bg_score_table26(name)[ia * 26 + ib];
}
inline int32_t blosum62(uint8_t a, uint8_t b) { return table26("blosum62", a, b); }
inline int32_t pam250(uint8_t a, uint8_t b) { return table26("pam250", a, b); }
inline int32_t unit(uint8_t a, uint8_t b) { return table26("unit", a, b); }
}  // namespace score

namespace aligner {

using Aligned = std::tuple<int32_t, ds::Sequence, ds::Sequence>;

class SequenceAligner {
  public:
    explicit SequenceAligner(const std::vector<int>& devices = {}) {   // SequenceAligner::new, aligner.rs:44
        const int rc = bg_create(devices.empty() ? nullptr : devices.data(), (int)devices.size(), &ctx_);
        if (rc != BG_OK) throw BioError(BioError::Engine, std::string("bg_create: ") + bg_strerror(rc));
    }
    ~SequenceAligner() { bg_destroy(ctx_); }
    SequenceAligner(const SequenceAligner&) = delete;
    SequenceAligner& operator=(const SequenceAligner&) = delete;

    Aligned global_alignment(const ds::Sequence& s1, const ds::Sequence& s2, const ScoreFn& sc, int32_t a, int32_t b) { return single(BG_GLOBAL, s1, s2, sc, a, b); }
    Aligned local_alignment(const ds::Sequence& s1, const ds::Sequence& s2, const ScoreFn& sc, int32_t a, int32_t b) { return single(BG_LOCAL, s1, s2, sc, a, b); }
    Aligned fitting_alignment(const ds::Sequence& s1, const ds::Sequence& s2, const ScoreFn& sc, int32_t a, int32_t b) { return single(BG_FITTING, s1, s2, sc, a, b); }
    Aligned overlap_alignment(const ds::Sequence& s1, const ds::Sequence& s2, const ScoreFn& sc, int32_t a, int32_t b) { return single(BG_OVERLAP, s1, s2, sc, a, b); }
    Aligned semiglobal_alignment(const ds::Sequence& s1, const ds::Sequence& s2, const ScoreFn& sc, int32_t a, int32_t b) { return single(BG_SEMIGLOBAL, s1, s2, sc, a, b); }

    // New batched entry point: pair p = (tile[2p], tile[2p+1]); status[p] != 0 marks pairs on which the
    // reference itself is undefined (the engine's extension is returned for them).
    std::vector<Aligned> align_batch(const ds::Tile& pairs, bg_mode mode, const ScoreFn& sc, int32_t a, int32_t b,
                                     std::vector<uint8_t>* status = nullptr) {
        if (pairs.len() % 2) throw BioError(BioError::InvalidInputSize, "Provided inputs have invalid size!");
        std::vector<uint8_t> res; std::vector<uint64_t> off{0};
        for (const auto& s : pairs.data) { res.insert(res.end(), s.chain.begin(), s.chain.end()); off.push_back(res.size()); }
        bg_batch batch{pairs.len() / 2, res.data(), off.data()};
        // callback -> dense table over the residues present; sign check first (aligner.rs:87-89)
        if ((mode == BG_GLOBAL || mode == BG_LOCAL || mode == BG_FITTING) && (a > 0 || b > 0))
            throw BioError(BioError::InvalidArgumentRange, "The provided has is within an unsupported range!");
        if (mode == BG_FITTING)
            for (size_t p = 0; p + 1 < pairs.len(); p += 2)
                if (pairs[p].len() < pairs[p + 1].len()) throw BioError(BioError::InvalidInputSize, "Provided inputs have invalid size!");
        std::array<uint64_t, 256> ha{}, hb{};
        bg_residue_histogram(&batch, ha.data(), hb.data());
        std::array<uint8_t, 256> rc, cc; rc.fill(0xFF); cc.fill(0xFF);
        std::vector<int> rows, cols;
        for (int x = 0; x < 256; ++x) { if (ha[x]) { rc[x] = (uint8_t)rows.size(); rows.push_back(x); } if (hb[x]) { cc[x] = (uint8_t)cols.size(); cols.push_back(x); } }
        const int nr = std::max<int>(1, (int)rows.size()), nc = std::max<int>(1, (int)cols.size());
        std::vector<int32_t> table((size_t)nr * nc, 0);
        for (size_t i = 0; i < rows.size(); ++i)
            for (size_t j = 0; j < cols.size(); ++j) table[i * nc + j] = sc((uint8_t)rows[i], (uint8_t)cols[j]);
        bg_params prm{(int32_t)mode, a, b, 0u, table.data(), nr, nc, rc.data(), cc.data()};
        // compact results + expansion straight into the Sequences this function has to allocate anyway (no arena in between):
        // what the Rust shim does too (rust/biogarden-gpu/src/lib.rs)
        bg_ops_result r{};
        const int err = bg_align_batch_ops(ctx_, &batch, &prm, &r);
        if (err == BG_EINVAL_RANGE) throw BioError(BioError::InvalidArgumentRange, "The provided has is within an unsupported range!");
        if (err == BG_EINVAL_SIZE) throw BioError(BioError::InvalidInputSize, "Provided inputs have invalid size!");
        if (err != BG_OK) throw BioError(BioError::Engine, std::string(bg_strerror(err)) + ": " + bg_last_error(ctx_));
        std::vector<Aligned> out(r.n_pairs);
        auto expand = [&](uint64_t lo, uint64_t hi) {
            for (uint64_t p = lo; p < hi; ++p) {
                const uint64_t len = r.len[p];
                std::vector<uint8_t> x(len), y(len);
                bg_expand_ops(res.data() + off[2 * p] + r.first[2 * p], res.data() + off[2 * p + 1] + r.first[2 * p + 1],
                              r.ops + r.ops_off[p], len, x.data(), y.data());
                out[p] = Aligned(r.score[p], ds::Sequence(std::move(x)), ds::Sequence(std::move(y)));
            }
        };
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const uint64_t nt = std::max<uint64_t>(1, std::min<uint64_t>(hw, r.n_pairs / 4096));
        if (nt == 1) expand(0, r.n_pairs);
        else {
            std::vector<std::thread> th;
            for (uint64_t t = 0; t < nt; ++t) th.emplace_back(expand, r.n_pairs * t / nt, r.n_pairs * (t + 1) / nt);
            for (auto& t : th) t.join();
        }
        if (status) status->assign(r.status, r.status + r.n_pairs);
        else
            for (uint64_t p = 0; p < r.n_pairs; ++p)
                if (r.status[p] != BG_ST_OK) { bg_ops_result_free(&r); throw BioError(BioError::ReferenceUndefined, "the reference panics or never returns on this input"); }
        bg_ops_result_free(&r);
        return out;
    }

    bg_ctx* context() { return ctx_; }

  private:
    Aligned single(bg_mode mode, const ds::Sequence& s1, const ds::Sequence& s2, const ScoreFn& sc, int32_t a, int32_t b) {
        ds::Tile t; t.push(s1); t.push(s2);
        return align_batch(t, mode, sc, a, b)[0];
    }
    bg_ctx* ctx_ = nullptr;
};

}  // namespace aligner
}  // namespace alignment

namespace analysis { namespace seq {

inline std::vector<uint64_t> edit_distance_batch(alignment::aligner::SequenceAligner& al, const ds::Tile& pairs) {
    if (pairs.len() % 2) throw BioError(BioError::InvalidInputSize, "Provided inputs have invalid size!");
    std::vector<uint8_t> res; std::vector<uint64_t> off{0};
    for (const auto& s : pairs.data) { res.insert(res.end(), s.chain.begin(), s.chain.end()); off.push_back(res.size()); }
    bg_batch batch{pairs.len() / 2, res.data(), off.data()};
    std::vector<uint64_t> out(batch.n_pairs);
    const int err = bg_edit_distance_batch(al.context(), &batch, out.data());
    if (err != BG_OK) throw BioError(BioError::Engine, bg_strerror(err));
    return out;
}
inline uint64_t edit_distance(alignment::aligner::SequenceAligner& al, const ds::Sequence& a, const ds::Sequence& b) {
    ds::Tile t; t.push(a); t.push(b);
    return edit_distance_batch(al, t)[0];
}

// analysis::seq::hamming_distance (seq.rs:74-83): Err(InvalidInputSize) when the lengths differ.
inline std::vector<uint64_t> hamming_distance_batch(alignment::aligner::SequenceAligner& al, const ds::Tile& pairs) {
    if (pairs.len() % 2) throw BioError(BioError::InvalidInputSize, "Provided inputs have invalid size!");
    std::vector<uint8_t> res; std::vector<uint64_t> off{0};
    for (const auto& s : pairs.data) { res.insert(res.end(), s.chain.begin(), s.chain.end()); off.push_back(res.size()); }
    bg_batch batch{pairs.len() / 2, res.data(), off.data()};
    std::vector<uint64_t> out(batch.n_pairs);
    const int err = bg_hamming_distance_batch(al.context(), &batch, out.data());
    if (err == BG_EINVAL_SIZE) throw BioError(BioError::InvalidInputSize, "Provided inputs have invalid size!");
    if (err != BG_OK) throw BioError(BioError::Engine, bg_strerror(err));
    return out;
}
inline uint64_t hamming_distance(alignment::aligner::SequenceAligner& al, const ds::Sequence& a, const ds::Sequence& b) {
    ds::Tile t; t.push(a); t.push(b);
    return hamming_distance_batch(al, t)[0];
}

}}  // namespace analysis::seq

namespace analysis { namespace stat {

// analysis::stat::p_distance_matrix (stat.rs:138-152): rows x rows, row-major f32.
inline std::vector<float> p_distance_matrix(alignment::aligner::SequenceAligner& al, const ds::Tile& matrix) {
    if (matrix.len() == 0) throw std::out_of_range("p_distance_matrix of an empty Tile (the reference panics on data[0])");
    std::vector<uint8_t> res; std::vector<uint64_t> off{0};
    for (const auto& s : matrix.data) { res.insert(res.end(), s.chain.begin(), s.chain.end()); off.push_back(res.size()); }
    std::vector<float> out(matrix.len() * matrix.len());
    const int err = bg_p_distance_matrix(al.context(), res.data(), off.data(), matrix.len(), out.data());
    if (err != BG_OK) throw BioError(BioError::Engine, bg_strerror(err));
    return out;
}

}}  // namespace analysis::stat

namespace io { namespace fasta {

// io::fasta::Reader::read_all over a text in memory (fasta.rs:95-136) by the native parser; ids are kept.
inline ds::Tile read_all(const std::string& text, std::vector<std::string>* ids = nullptr) {
    bg_fasta f{};
    const int err = bg_fasta_parse(reinterpret_cast<const uint8_t*>(text.data()), text.size(), 0, &f);
    if (err == BG_EINVAL_FASTA) throw std::runtime_error("Expected > at record start.");
    if (err != BG_OK) throw BioError(BioError::Engine, bg_strerror(err));
    ds::Tile t;
    for (uint64_t r = 0; r < f.n_records; ++r) {
        t.push(ds::Sequence(std::vector<uint8_t>(f.residues + f.seq_off[r], f.residues + f.seq_off[r + 1])));
        if (ids) ids->emplace_back(reinterpret_cast<const char*>(f.ids) + f.id_off[r], f.ids + f.id_off[r + 1] - (f.ids + f.id_off[r]));
    }
    bg_fasta_free(&f);
    return t;
}

}}  // namespace io::fasta
}  // namespace biogarden
#endif
