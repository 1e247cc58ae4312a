// fasta_ingest.cpp -- FASTA text -> the batch layout of bgalign.h (one residue arena + offsets), the step before
// the alignment path.  Record grammar of the reference reader (src/io/fasta.rs:95-136):
//   * lines end at '\n' (read_line); a record starts at a line whose first byte is '>';
//   * id = the header after '>' with trailing whitespace trimmed, up to its first whitespace character (may be
//     empty), the rest is the description (dropped here: Sequence::from(Record) keeps id and seq only);
//   * sequence = the following lines, each with trailing whitespace trimmed (leading / inner whitespace stays),
//     concatenated, up to the next record start or the end of the text;
//   * read_all stops at the first EMPTY record (empty id, no description, empty sequence; fasta.rs:232-234), and
//     a first line that does not start with '>' is an error ("Expected > at record start.", fasta.rs:104-109).
// Whitespace is ASCII whitespace (space, \t, \n, \v, \f, \r): the reference trims Unicode whitespace, so a
// line ending in a multi-byte space (U+0085, U+00A0, ...) differs -- FASTA payloads are ASCII.
// Host code only; two parallel passes over the text (sizes, then payload straight into the final arena), so a
// multi-GB read set is ingested at memory speed instead of one core's.
#include "../../include/bgalign.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

inline bool is_ws(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

struct Seg { uint64_t begin, end; uint64_t n_rec = 0, n_res = 0, n_id = 0; int64_t first_empty = -1; };

// Visits the records whose header line starts inside [begin, end) (begin is a record start).
// on_record(header_lo, header_hi /* trimmed */, k) then on_line(lo, hi /* trimmed */) per sequence line.
template <class R, class L>
void walk(const uint8_t* t, uint64_t len, uint64_t begin, uint64_t end, R&& on_record, L&& on_line) {
    uint64_t pos = begin;
    while (pos < end) {
        // header line [pos, eol)
        const uint8_t* nl = (const uint8_t*)memchr(t + pos, '\n', len - pos);
        uint64_t eol = nl ? (uint64_t)(nl - t) : len;
        uint64_t hi = eol;
        while (hi > pos + 1 && is_ws(t[hi - 1])) --hi;
        if (!on_record(pos + 1, hi)) return;
        pos = nl ? eol + 1 : len;
        while (pos < len && t[pos] != '>') {
            nl = (const uint8_t*)memchr(t + pos, '\n', len - pos);
            eol = nl ? (uint64_t)(nl - t) : len;
            hi = eol;
            while (hi > pos && is_ws(t[hi - 1])) --hi;
            on_line(pos, hi);
            pos = nl ? eol + 1 : len;
        }
    }
}

inline bool t_is_start(const uint8_t* t, uint64_t e) { return e == 0 || (t[e - 1] == '\n' && t[e] == '>'); }

inline void split_header(const uint8_t* t, uint64_t lo, uint64_t hi, uint64_t& id_hi, bool& has_desc) {
    id_hi = lo;
    while (id_hi < hi && !is_ws(t[id_hi])) ++id_hi;
    has_desc = id_hi < hi;          // splitn(2, whitespace): a second field exists as soon as one whitespace is there
}

}  // namespace

extern "C" {

void bg_fasta_free(bg_fasta* f) {
    if (!f) return;
    free(f->residues); free(f->seq_off); free(f->ids); free(f->id_off);
    memset(f, 0, sizeof *f);
}

int bg_fasta_parse(const uint8_t* text, uint64_t len, int n_threads, bg_fasta* out) {
    if (!out || (!text && len)) return BG_EINVAL_ARG;
    memset(out, 0, sizeof *out);
    if (len == 0) {
        out->seq_off = (uint64_t*)calloc(1, 8); out->id_off = (uint64_t*)calloc(1, 8);
        out->residues = (uint8_t*)malloc(1); out->ids = (uint8_t*)malloc(1);
        return (out->seq_off && out->id_off && out->residues && out->ids) ? BG_OK : BG_ENOMEM;
    }
    if (text[0] != '>') return BG_EINVAL_FASTA;      // fasta.rs:104-109 (only the first line can fail this test)
    unsigned hw = std::thread::hardware_concurrency();
    uint64_t T = n_threads > 0 ? (uint64_t)n_threads : (hw ? hw : 1);
    T = std::max<uint64_t>(1, std::min<uint64_t>(T, len / (1u << 20) + 1));
    // cut at record starts ("\n>")
    std::vector<Seg> seg;
    uint64_t b = 0;
    for (uint64_t k = 1; k <= T && b < len; ++k) {
        uint64_t e = (k == T) ? len : std::max<uint64_t>(b + 1, len * k / T);
        while (e < len && !(t_is_start(text, e))) {
            const uint8_t* nl = (const uint8_t*)memchr(text + e, '\n', len - e);
            e = nl ? (uint64_t)(nl - text) + 1 : len;
        }
        Seg s; s.begin = b; s.end = e; seg.push_back(s);
        b = e;
    }
    // pass 1: sizes
    auto run = [&](auto&& fn) {
        if (seg.size() == 1) { fn(0); return; }
        std::vector<std::thread> th;
        for (size_t k = 0; k < seg.size(); ++k) th.emplace_back(fn, k);
        for (auto& x : th) x.join();
    };
    run([&](size_t k) {
        Seg& s = seg[k];
        uint64_t rec_res = 0; bool open = false, rec_hdr_empty = false;
        auto close = [&] { if (open && rec_hdr_empty && rec_res == 0 && s.first_empty < 0) s.first_empty = (int64_t)s.n_rec - 1; };
        walk(text, len, s.begin, s.end,
             [&](uint64_t lo, uint64_t hi) {
                 close();
                 uint64_t id_hi; bool has_desc; split_header(text, lo, hi, id_hi, has_desc);
                 s.n_rec++; s.n_id += id_hi - lo; rec_res = 0; open = true; rec_hdr_empty = (id_hi == lo) && !has_desc;
                 return true;
             },
             [&](uint64_t lo, uint64_t hi) { s.n_res += hi - lo; rec_res += hi - lo; });
        close();
    });
    // read_all ends at the first empty record: drop it and everything after it
    uint64_t n_rec = 0; bool cut = false; size_t cut_seg = 0; uint64_t cut_local = 0;
    for (size_t k = 0; k < seg.size() && !cut; ++k) {
        if (seg[k].first_empty >= 0) { cut = true; cut_seg = k; cut_local = (uint64_t)seg[k].first_empty; n_rec += cut_local; }
        else n_rec += seg[k].n_rec;
    }
    const size_t n_seg = cut ? cut_seg + 1 : seg.size();
    // totals (over-allocated when cut: the cut segment's full size is an upper bound)
    uint64_t tot_res = 0, tot_id = 0;
    std::vector<uint64_t> rec0(n_seg + 1, 0), res0(n_seg + 1, 0), id0(n_seg + 1, 0);
    for (size_t k = 0; k < n_seg; ++k) {
        rec0[k + 1] = rec0[k] + seg[k].n_rec; res0[k + 1] = res0[k] + seg[k].n_res; id0[k + 1] = id0[k] + seg[k].n_id;
    }
    tot_res = res0[n_seg]; tot_id = id0[n_seg];
    out->residues = (uint8_t*)malloc(tot_res + 16); out->ids = (uint8_t*)malloc(tot_id + 16);
    out->seq_off = (uint64_t*)malloc((rec0[n_seg] + 1) * 8); out->id_off = (uint64_t*)malloc((rec0[n_seg] + 1) * 8);
    if (!out->residues || !out->ids || !out->seq_off || !out->id_off) { bg_fasta_free(out); return BG_ENOMEM; }
    // pass 2: payload
    seg.resize(n_seg);
    run([&](size_t k) {
        const Seg& s = seg[k];
        uint64_t r = rec0[k], ro = res0[k], io = id0[k];
        const uint64_t r_stop = (cut && k == cut_seg) ? rec0[k] + cut_local : ~0ull;
        walk(text, len, s.begin, s.end,
             [&](uint64_t lo, uint64_t hi) {
                 if (r >= r_stop) return false;
                 uint64_t id_hi; bool has_desc; split_header(text, lo, hi, id_hi, has_desc);
                 out->seq_off[r] = ro; out->id_off[r] = io;
                 memcpy(out->ids + io, text + lo, id_hi - lo); io += id_hi - lo;
                 ++r;
                 return true;
             },
             [&](uint64_t lo, uint64_t hi) { memcpy(out->residues + ro, text + lo, hi - lo); ro += hi - lo; });
        if (k + 1 == n_seg) { out->seq_off[r] = ro; out->id_off[r] = io; }
    });
    out->n_records = n_rec;
    // when cut inside the last kept segment, the terminating offsets were written at index r == n_rec by that segment
    return BG_OK;
}

int bg_fasta_parse_packed(const uint8_t* text, uint64_t len, int n_threads, int bits, bg_fasta* out) {
    if (bits != 2 && bits != 5) return BG_EINVAL_ARG;
    int rc = bg_fasta_parse(text, len, n_threads, out);
    if (rc) return rc;
    // second pass over the payload only (1 B read, 0.25 / 0.625 B written per residue), in parallel
    const uint64_t n = out->n_records ? out->seq_off[out->n_records] : 0;
    uint8_t* packed = (uint8_t*)malloc(bg_packed_bytes(n, bits));
    if (!packed) { bg_fasta_free(out); return BG_ENOMEM; }
    rc = bg_pack_residues(out->residues, n, bits, n_threads, packed, out->alphabet);
    if (rc) { free(packed); bg_fasta_free(out); return rc; }
    free(out->residues);
    out->residues = packed;
    out->packing = (uint32_t)bits;
    return BG_OK;
}

}  // extern "C"
