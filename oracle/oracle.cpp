// oracle.cpp -- CPU restatement of biogarden's SequenceAligner + edit_distance.
// TEST INFRASTRUCTURE ONLY (see oracle.h).  Citations are file:line in the
// reference checkout (robsndr/biogarden, crate version 0.1.0).
//
// The literal form keeps the reference's data layout and control flow so that it
// (a) pins the semantics against the reference's golden files and (b) is an honest
// stand-in for the reference's CPU cost.  Rust's run-time checks are restated as
// explicit checks that throw: Panic (slice / ndarray index out of bounds, usize
// underflow feeding an index) and Hang (the `_ => {}` arm at aligner.rs:549).
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "score_tables.inc"

namespace {

struct Panic {};
struct Hang {};

// ndarray::Array2 stand-in: storage + bounds-checked indexing only (Cargo.toml:11;
// no arithmetic of ndarray's is on the path).
template <class T>
struct Mat {
    size_t R = 0, C = 0;
    std::vector<T> d;
    void assign(size_t r, size_t c, T v) { R = r; C = c; d.assign(r * c, v); }
    inline T& at(size_t i, size_t j) {
        if (i >= R || j >= C) throw Panic();
        return d[i * C + j];
    }
    void fill(T v) { std::fill(d.begin(), d.end(), v); }
    void fill_col0(T v) { for (size_t i = 0; i < R; ++i) d[i * C] = v; }   // column_mut(0).fill
    void fill_row0(T v) { for (size_t j = 0; j < C; ++j) d[j] = v; }       // row_mut(0).fill
};

// Sequence stand-in: Vec<u8> with checked Index (sequence.rs:119-139).
struct Seq {
    const uint8_t* p; size_t n;
    inline uint8_t at(size_t i) const { if (i >= n) throw Panic(); return p[i]; }
    size_t len() const { return n; }
};

typedef int32_t (*ScoreFn)(const void* ctx, uint8_t a, uint8_t b);

// score.rs:38-41 / 78-80 / 114-116: TABLE[(a - 65, b - 65)] on a 26x26 array.
template <const int8_t (*T)[26]>
int32_t score26(const void*, uint8_t a, uint8_t b) {
    size_t ia = (size_t)a - 65, ib = (size_t)b - 65;  // usize underflow wraps -> OOB -> panic
    if (ia >= 26 || ib >= 26) throw Panic();
    return T[ia][ib];
}
int32_t score_table(const void* ctx, uint8_t a, uint8_t b) {
    return ((const int32_t*)ctx)[(size_t)a * 256 + b];
}
ScoreFn pick_scorer(int scorer) {
    switch (scorer) {
        case ORC_SCORE_BLOSUM62: return &score26<ORC_BLOSUM62>;
        case ORC_SCORE_PAM250: return &score26<ORC_PAM250>;
        case ORC_SCORE_UNIT: return &score26<ORC_UNIT>;
        default: return &score_table;
    }
}

inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }  // release-mode `+`
inline int32_t sat_add(int32_t a, int32_t b) {                                               // i32::saturating_add
    int64_t s = (int64_t)a + b;
    if (s > INT32_MAX) return INT32_MAX;
    if (s < INT32_MIN) return INT32_MIN;
    return (int32_t)s;
}

}  // namespace

// ------------------------------------------------------------------ literal ---
struct orc_aligner {
    size_t R, C;                       // buffer_size (aligner.rs:30)
    Mat<int32_t> m, x, y;              // aligner.rs:32-34
    Mat<uint8_t> mt, xt, yt;           // aligner.rs:36-38

    orc_aligner() { alloc(1024, 1024); }   // aligner.rs:44-55
    void alloc(size_t r, size_t c) {       // also resize_buffers, aligner.rs:594-602
        R = r; C = c;
        m.assign(r, c, 0);
        x.assign(r, c, INT32_MIN);
        y.assign(r, c, INT32_MIN);
        mt.assign(r, c, 0);
        xt.assign(r, c, 'I');
        yt.assign(r, c, 'I');
    }
    void maybe_resize(size_t n, size_t mm) {   // aligner.rs:92-94 (note `>`, not `>=`)
        if (n > R || mm > C) alloc(n + 1, mm + 1);
    }

    // aligner.rs:437-469
    void compute_scores_global(const Seq& s1, const Seq& s2, ScoreFn score, const void* ctx, int32_t a, int32_t b) {
        for (size_t i = 1; i < s1.len() + 1; ++i) {
            for (size_t j = 1; j < s2.len() + 1; ++j) {
                x.at(i, j) = std::max(wadd(m.at(i - 1, j), a), sat_add(x.at(i - 1, j), b));
                xt.at(i, j) = (x.at(i, j) == wadd(m.at(i - 1, j), a)) ? 'M' : 'I';
                y.at(i, j) = std::max(wadd(m.at(i, j - 1), a), sat_add(y.at(i, j - 1), b));
                yt.at(i, j) = (y.at(i, j) == wadd(m.at(i, j - 1), a)) ? 'M' : 'I';
                int32_t maximum = std::max(wadd(m.at(i - 1, j - 1), score(ctx, s1.at(i - 1), s2.at(j - 1))),
                                           std::max(x.at(i, j), y.at(i, j)));
                if (maximum == y.at(i, j)) mt.at(i, j) = 'Y';
                else if (maximum == x.at(i, j)) mt.at(i, j) = 'X';
                else mt.at(i, j) = 'R';
                m.at(i, j) = maximum;
            }
        }
    }

    // aligner.rs:471-509
    void compute_scores_local(const Seq& s1, const Seq& s2, ScoreFn score, const void* ctx, int32_t a, int32_t b) {
        for (size_t i = 1; i < s1.len() + 1; ++i) {
            for (size_t j = 1; j < s2.len() + 1; ++j) {
                x.at(i, j) = std::max(wadd(m.at(i - 1, j), a), sat_add(x.at(i - 1, j), b));
                xt.at(i, j) = (x.at(i, j) == wadd(m.at(i - 1, j), a)) ? 'M' : 'I';
                x.at(i, j) = x.at(i, j) < 0 ? 0 : x.at(i, j);
                y.at(i, j) = std::max(wadd(m.at(i, j - 1), a), sat_add(y.at(i, j - 1), b));
                yt.at(i, j) = (y.at(i, j) == wadd(m.at(i, j - 1), a)) ? 'M' : 'I';
                y.at(i, j) = y.at(i, j) < 0 ? 0 : y.at(i, j);
                int32_t maximum = std::max(wadd(m.at(i - 1, j - 1), score(ctx, s1.at(i - 1), s2.at(j - 1))),
                                           std::max(x.at(i, j), y.at(i, j)));
                if (maximum == y.at(i, j)) mt.at(i, j) = 'Y';
                else if (maximum == x.at(i, j)) mt.at(i, j) = 'X';
                else mt.at(i, j) = 'R';
                m.at(i, j) = maximum < 0 ? 0 : maximum;
            }
        }
    }

    // aligner.rs:511-592.  Returns the two strings in *push* order (i.e. reversed);
    // callers apply the reversals the reference applies.
    void backtrack(const Seq& s1, const Seq& s2, size_t& k, size_t& l,
                   const std::function<bool(size_t, size_t)>& valid,
                   std::vector<uint8_t>& r1, std::vector<uint8_t>& r2) {
        uint8_t cur = 'M';
        while (valid(k, l)) {
            if (cur == 'M') {
                uint8_t t = mt.at(k, l);
                if (t == 'R') {
                    r1.push_back(s1.at(k - 1)); r2.push_back(s2.at(l - 1)); k -= 1; l -= 1;
                } else if (t == 'X') {
                    cur = 'X'; r1.push_back(s1.at(k - 1)); r2.push_back('-'); k -= 1;
                } else if (t == 'Y') {
                    cur = 'Y'; r1.push_back('-'); r2.push_back(s2.at(l - 1)); l -= 1;
                } else {
                    throw Hang();   // aligner.rs:549: `_ => {}` leaves (k,l,cur) unchanged forever
                }
            } else if (cur == 'X') {
                if (xt.at(k, l) == 'M') cur = 'M';
                else { r1.push_back(s1.at(k - 1)); r2.push_back('-'); k -= 1; }
            } else {
                if (yt.at(k, l) == 'M') cur = 'M';
                else { r1.push_back('-'); r2.push_back(s2.at(l - 1)); l -= 1; }
            }
        }
    }

    int run(int mode, const Seq& s1, const Seq& s2, ScoreFn score, const void* ctx, int32_t a, int32_t b,
            int32_t* out_score, std::vector<uint8_t>& o1, std::vector<uint8_t>& o2) {
        const size_t n = s1.len(), mm = s2.len();
        std::vector<uint8_t> r1, r2;   // reversed (push-order) buffers
        size_t k, l;
        switch (mode) {
        case ORC_GLOBAL: {   // aligner.rs:84-121
            if (a > 0 || b > 0) return ORC_ERR_RANGE;
            maybe_resize(n, mm);
            m.at(0, 1) = a;
            for (size_t j = 2; j < mm + 1; ++j) m.at(0, j) = wadd(m.at(0, j - 1), b);
            m.at(1, 0) = a;
            for (size_t i = 2; i < n + 1; ++i) m.at(i, 0) = wadd(m.at(i - 1, 0), b);
            mt.fill_col0('X'); mt.fill_row0('Y');
            compute_scores_global(s1, s2, score, ctx, a, b);
            *out_score = m.at(n, mm);
            k = n; l = mm;
            backtrack(s1, s2, k, l, [](size_t p, size_t q) { return p != 0 || q != 0; }, r1, r2);
            break;
        }
        case ORC_LOCAL: {    // aligner.rs:150-185
            if (a > 0 || b > 0) return ORC_ERR_RANGE;
            maybe_resize(n, mm);
            m.fill(0);
            mt.fill_col0('X'); mt.fill_row0('Y');
            compute_scores_local(s1, s2, score, ctx, a, b);
            size_t bi = 0, bj = 0; int32_t best = INT32_MIN;   // indexed_iter fold, strict >
            for (size_t i = 0; i < R; ++i)
                for (size_t j = 0; j < C; ++j)
                    if (m.d[i * C + j] > best) { best = m.d[i * C + j]; bi = i; bj = j; }
            *out_score = best;
            k = bi; l = bj;
            backtrack(s1, s2, k, l, [this](size_t p, size_t q) { return (p != 0 || q != 0) && m.at(p, q) > 0; }, r1, r2);
            break;
        }
        case ORC_FITTING: {  // aligner.rs:216-260
            if (a > 0 || b > 0) return ORC_ERR_RANGE;
            if (n < mm) return ORC_ERR_SIZE;
            maybe_resize(n, mm);
            m.fill(0);
            m.at(0, 1) = a;
            for (size_t j = 2; j < mm + 1; ++j) m.at(0, j) = wadd(m.at(0, j - 1), b);
            mt.fill_col0('X'); mt.fill_row0('Y');
            compute_scores_global(s1, s2, score, ctx, a, b);
            if (mm >= C) throw Panic();   // .column(seq2.len())
            size_t bi = 0; int32_t best = INT32_MIN;
            for (size_t i = 0; i < R; ++i) if (m.d[i * C + mm] > best) { best = m.d[i * C + mm]; bi = i; }
            *out_score = best;
            k = bi; l = mm;
            backtrack(s1, s2, k, l, [](size_t, size_t q) { return q != 0; }, r1, r2);
            break;
        }
        case ORC_OVERLAP: {  // aligner.rs:290-321 (no sign check)
            maybe_resize(n, mm);
            m.fill(0);
            mt.fill_col0('X'); mt.fill_row0('Y');
            compute_scores_global(s1, s2, score, ctx, a, b);
            if (n >= R) throw Panic();    // .row(seq1.len())
            size_t bj = 0; int32_t best = INT32_MIN;
            for (size_t j = 0; j < C; ++j) if (m.d[n * C + j] >= best) { best = m.d[n * C + j]; bj = j; }
            *out_score = best;
            k = n; l = bj;
            backtrack(s1, s2, k, l, [](size_t, size_t q) { return q != 0; }, r1, r2);
            break;
        }
        case ORC_SEMIGLOBAL: {  // aligner.rs:351-435 (no sign check)
            maybe_resize(n, mm);
            m.fill(0);
            mt.fill_col0('X'); mt.fill_row0('Y');
            compute_scores_global(s1, s2, score, ctx, a, b);
            if (n >= R) throw Panic();
            size_t rj = 0; int32_t rbest = INT32_MIN;      // last max (>=), aligner.rs:369-373
            for (size_t j = 0; j < C; ++j) if (m.d[n * C + j] >= rbest) { rbest = m.d[n * C + j]; rj = j; }
            if (mm >= C) throw Panic();
            size_t ci = 0; int32_t cbest = INT32_MIN;      // first max (>), aligner.rs:376-380
            for (size_t i = 0; i < R; ++i) if (m.d[i * C + mm] > cbest) { cbest = m.d[i * C + mm]; ci = i; }
            const bool col = cbest > rbest;                // aligner.rs:389
            if (col) {
                k = ci; l = mm; *out_score = cbest;
                for (size_t i = n + 1; i-- > ci + 1;) { r1.push_back(s1.at(i - 1)); r2.push_back('-'); }
            } else {
                k = n; l = rj; *out_score = rbest;
                for (size_t i = mm + 1; i-- > rj + 1;) { r1.push_back('-'); r2.push_back(s2.at(i - 1)); }
            }
            // aligner.rs:410-414: backtrack reverses, the caller reverses back -> push order
            backtrack(s1, s2, k, l, [](size_t p, size_t q) { return p * q != 0; }, r1, r2);
            if (col) { for (size_t i = k; i-- > 0;) { r1.push_back(s1.at(i)); r2.push_back('-'); } }
            else     { for (size_t i = l; i-- > 0;) { r1.push_back('-'); r2.push_back(s2.at(i)); } }
            break;
        }
        default: return ORC_ERR_RANGE;
        }
        o1.assign(r1.rbegin(), r1.rend());
        o2.assign(r2.rbegin(), r2.rend());
        return ORC_OK;
    }
};

// --------------------------------------------------------------------- lean ---
namespace {

// Fresh-aligner buffer dims (aligner.rs:45, 92-94, 594-595).
inline void fresh_dims(size_t n, size_t m, size_t& R, size_t& C) {
    if (n > 1024 || m > 1024) { R = n + 1; C = m + 1; } else { R = 1024; C = 1024; }
}

// trace nibble: bits 1:0 = M-trace {0 R, 1 X, 2 Y, 3 STOP (local: M == 0)}, bit 2 = x_trace=='M', bit 3 = y_trace=='M'
struct Lean {
    size_t n, m;
    std::vector<uint8_t> tr;   // (n*m+1)/2 bytes, interior cells only
    inline void put(size_t i, size_t j, unsigned v) {
        size_t c = (i - 1) * m + (j - 1);
        uint8_t& b = tr[c >> 1];
        b = (c & 1) ? (uint8_t)((b & 0x0f) | (v << 4)) : (uint8_t)((b & 0xf0) | v);
    }
    inline unsigned get(size_t i, size_t j) const {
        size_t c = (i - 1) * m + (j - 1);
        return (tr[c >> 1] >> ((c & 1) * 4)) & 15;
    }
};

int lean_run(int mode, const Seq& s1, const Seq& s2, ScoreFn score, const void* ctx, int32_t a, int32_t b,
             int32_t* out_score, std::vector<uint8_t>& o1, std::vector<uint8_t>& o2) {
    const size_t n = s1.len(), m = s2.len();
    if (mode == ORC_GLOBAL || mode == ORC_LOCAL || mode == ORC_FITTING)
        if (a > 0 || b > 0) return ORC_ERR_RANGE;
    if (mode == ORC_FITTING && n < m) return ORC_ERR_SIZE;
    size_t R, C; fresh_dims(n, m, R, C);
    // Index n / m must exist in the buffers whenever the fill or a border init touches them.
    const bool row_border = (mode == ORC_GLOBAL || mode == ORC_FITTING);  // writes row0[1..=m]
    const bool col_border = (mode == ORC_GLOBAL);                          // writes col0[1..=n]
    if (row_border && (C < 2 || m >= C)) return ORC_PANIC;
    if (col_border && (R < 2 || n >= R)) return ORC_PANIC;
    if (n >= 1 && m >= 1 && (n >= R || m >= C)) return ORC_PANIC;
    if ((mode == ORC_OVERLAP || mode == ORC_SEMIGLOBAL) && n >= R) return ORC_PANIC;
    if ((mode == ORC_FITTING || mode == ORC_SEMIGLOBAL) && m >= C) return ORC_PANIC;

    Lean t; t.n = n; t.m = m; t.tr.assign((n * m + 1) / 2 + 1, 0);
    std::vector<int32_t> Mp(m + 1), Mc(m + 1), Xp(m + 1, INT32_MIN), colM(n + 1);
    const bool local = (mode == ORC_LOCAL);
    // row 0 of M (A.1)
    Mp[0] = 0;
    for (size_t j = 1; j <= m; ++j) Mp[j] = row_border ? (j == 1 ? a : wadd(Mp[j - 1], b)) : 0;
    colM[0] = Mp[m];
    int32_t best = INT32_MIN; size_t bi = 0, bj = 0;   // local first-max, row-major, strict >
    if (local) { for (size_t j = 0; j <= m; ++j) if (Mp[j] > best) { best = Mp[j]; bi = 0; bj = j; } }
    int32_t col0 = 0;
    for (size_t i = 1; i <= n; ++i) {
        col0 = col_border ? (i == 1 ? a : wadd(col0, b)) : 0;
        Mc[0] = col0;
        if (local && Mc[0] > best) { best = Mc[0]; bi = i; bj = 0; }
        int32_t Y = INT32_MIN;
        const uint8_t c1 = s1.at(i - 1);
        for (size_t j = 1; j <= m; ++j) {
            int32_t xo = wadd(Mp[j], a), X = std::max(xo, sat_add(Xp[j], b));
            unsigned nib = (X == xo) ? 4u : 0u;
            if (local && X < 0) X = 0;
            int32_t yo = wadd(Mc[j - 1], a); Y = std::max(yo, sat_add(Y, b));
            nib |= (Y == yo) ? 8u : 0u;
            if (local && Y < 0) Y = 0;
            int32_t mx = std::max(wadd(Mp[j - 1], score(ctx, c1, s2.at(j - 1))), std::max(X, Y));
            nib |= (mx == Y) ? 2u : (mx == X) ? 1u : 0u;
            if (local) { if (mx < 0) mx = 0; if (mx == 0) nib |= 3u; }
            t.put(i, j, nib);
            Xp[j] = X; Mc[j] = mx;
            if (local && mx > best) { best = mx; bi = i; bj = j; }
        }
        colM[i] = Mc[m];
        std::swap(Mp, Mc);
    }
    // Mp now holds row n.
    size_t k = 0, l = 0; bool col = false; int32_t sc = 0;
    std::vector<uint8_t> r1, r2;
    auto row_last_max = [&](size_t& rj, int32_t& rb) {  // over the whole buffer row: cells beyond m are 0
        rb = INT32_MIN; rj = 0;
        for (size_t j = 0; j <= m && j < C; ++j) if (Mp[j] >= rb) { rb = Mp[j]; rj = j; }
        if (C > m + 1 && 0 >= rb) { rb = 0; rj = C - 1; }
    };
    auto col_first_max = [&](size_t& ci, int32_t& cb) {
        cb = INT32_MIN; ci = 0;
        for (size_t i = 0; i <= n && i < R; ++i) if (colM[i] > cb) { cb = colM[i]; ci = i; }
        if (R > n + 1 && 0 > cb) { cb = 0; ci = n + 1; }
    };
    switch (mode) {
    case ORC_GLOBAL: k = n; l = m; sc = Mp[m]; break;
    case ORC_LOCAL:
        // whole-buffer scan: cells outside the rectangle are 0 after fill (only matters if best < 0: impossible)
        k = bi; l = bj; sc = best; break;
    case ORC_FITTING: { col_first_max(k, sc); l = m; break; }
    case ORC_OVERLAP: { row_last_max(l, sc); k = n; break; }
    case ORC_SEMIGLOBAL: {
        size_t rj, ci; int32_t rb, cb;
        row_last_max(rj, rb); col_first_max(ci, cb);
        col = cb > rb;
        if (col) { k = ci; l = m; sc = cb; for (size_t i = n + 1; i-- > ci + 1;) { r1.push_back(s1.at(i - 1)); r2.push_back('-'); } }
        else     { k = n; l = rj; sc = rb; for (size_t i = m + 1; i-- > rj + 1;) { r1.push_back('-'); r2.push_back(s2.at(i - 1)); } }
        break;
    }
    }
    *out_score = sc;
    // start cell outside the filled rectangle: the reference reads stale / zero trace bytes
    // there -> hang (m_trace 0) or panic; either way undefined.
    if (k > n || l > m) {
        bool enters = false;
        switch (mode) {
            case ORC_SEMIGLOBAL: enters = (k * l != 0); break;
            default: enters = (l != 0); break;
        }
        if (enters) return ORC_HANG;   // fresh buffers: m_trace outside the rectangle is 0
    }
    // walk (A.4)
    uint8_t cur = 'M';
    auto valid = [&](size_t p, size_t q) -> bool {
        switch (mode) {
            case ORC_GLOBAL: return p != 0 || q != 0;
            case ORC_LOCAL: {
                if (!(p != 0 || q != 0)) return false;
                if (p == 0 || q == 0) return false;               // borders are 0 in local mode
                return (t.get(p, q) & 3u) != 3u;
            }
            case ORC_SEMIGLOBAL: return p * q != 0;
            default: return q != 0;
        }
    };
    auto MT = [&](size_t p, size_t q) -> uint8_t {
        if (q == 0) return p == 0 ? 'Y' : 'X';
        if (p == 0) return 'Y';
        unsigned c = t.get(p, q) & 3u;
        return c == 0 ? 'R' : c == 1 ? 'X' : 'Y';
    };
    auto XTm = [&](size_t p, size_t q) { return p != 0 && q != 0 && (t.get(p, q) & 4u); };
    auto YTm = [&](size_t p, size_t q) { return p != 0 && q != 0 && (t.get(p, q) & 8u); };
    while (valid(k, l)) {
        if (cur == 'M') {
            uint8_t c = MT(k, l);
            if (c == 'R') { r1.push_back(s1.at(k - 1)); r2.push_back(s2.at(l - 1)); --k; --l; }
            else if (c == 'X') { cur = 'X'; r1.push_back(s1.at(k - 1)); r2.push_back('-'); --k; }
            else { cur = 'Y'; r1.push_back('-'); r2.push_back(s2.at(l - 1)); --l; }
        } else if (cur == 'X') {
            if (XTm(k, l)) cur = 'M'; else { r1.push_back(s1.at(k - 1)); r2.push_back('-'); --k; }
        } else {
            if (YTm(k, l)) cur = 'M'; else { r1.push_back('-'); r2.push_back(s2.at(l - 1)); --l; }
        }
    }
    if (mode == ORC_SEMIGLOBAL) {
        if (col) { for (size_t i = k; i-- > 0;) { r1.push_back(s1.at(i)); r2.push_back('-'); } }
        else     { for (size_t i = l; i-- > 0;) { r1.push_back('-'); r2.push_back(s2.at(i)); } }
    }
    o1.assign(r1.rbegin(), r1.rend());
    o2.assign(r2.rbegin(), r2.rend());
    return ORC_OK;
}

template <class F>
int guarded(F&& f) {
    try { return f(); }
    catch (const Panic&) { return ORC_PANIC; }
    catch (const Hang&) { return ORC_HANG; }
    catch (const std::bad_alloc&) { return ORC_PANIC; }
}

int emit(int st, const std::vector<uint8_t>& o1, const std::vector<uint8_t>& o2,
         uint8_t* a_out, uint8_t* b_out, size_t cap, size_t* out_len) {
    if (st != ORC_OK) { if (out_len) *out_len = 0; return st; }
    if (out_len) *out_len = o1.size();
    if (o1.size() != o2.size()) return ORC_PANIC;
    if (a_out && b_out) {
        if (o1.size() > cap) return ORC_PANIC;
        if (!o1.empty()) { memcpy(a_out, o1.data(), o1.size()); memcpy(b_out, o2.data(), o2.size()); }
    }
    return ORC_OK;
}

inline uint64_t fnv1a(uint64_t h, const std::vector<uint8_t>& v) {
    for (uint8_t c : v) { h ^= c; h *= 1099511628211ull; }
    return h;
}

template <class F>
double run_threads(uint64_t n_items, int n_threads, F&& body) {
    if (n_threads < 1) n_threads = 1;
    auto t0 = std::chrono::steady_clock::now();
    if (n_threads == 1) { body(0, 0, n_items); }
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) {
            uint64_t lo = n_items * t / n_threads, hi = n_items * (t + 1) / n_threads;
            th.emplace_back([&, t, lo, hi] { body(t, lo, hi); });
        }
        for (auto& x : th) x.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

extern "C" {

orc_aligner* orc_aligner_new(void) { return new orc_aligner(); }
void orc_aligner_free(orc_aligner* a) { delete a; }

int orc_align(orc_aligner* al, int mode, const uint8_t* s1, size_t n, const uint8_t* s2, size_t m,
              int scorer, const int32_t* table, int32_t a, int32_t b,
              int32_t* score, uint8_t* a_out, uint8_t* b_out, size_t cap, size_t* out_len) {
    std::vector<uint8_t> o1, o2;
    int32_t sc = 0;
    int st = guarded([&] { return al->run(mode, Seq{s1, n}, Seq{s2, m}, pick_scorer(scorer), table, a, b, &sc, o1, o2); });
    if (score) *score = sc;
    return emit(st, o1, o2, a_out, b_out, cap, out_len);
}

int orc_align_lean(int mode, const uint8_t* s1, size_t n, const uint8_t* s2, size_t m,
                   int scorer, const int32_t* table, int32_t a, int32_t b,
                   int32_t* score, uint8_t* a_out, uint8_t* b_out, size_t cap, size_t* out_len) {
    std::vector<uint8_t> o1, o2;
    int32_t sc = 0;
    int st = guarded([&] { return lean_run(mode, Seq{s1, n}, Seq{s2, m}, pick_scorer(scorer), table, a, b, &sc, o1, o2); });
    if (score) *score = sc;
    return emit(st, o1, o2, a_out, b_out, cap, out_len);
}

// seq.rs:105-130, literal: full (n+1)x(m+1) table of u128.
int orc_edit_distance(const uint8_t* s1, size_t n, const uint8_t* s2, size_t m, uint64_t* out) {
    typedef unsigned __int128 u128;
    std::vector<std::vector<u128>> memo(n + 1, std::vector<u128>(m + 1, 0));
    for (size_t i = 0; i < n + 1; ++i) memo[i][0] = (u128)i;
    for (size_t j = 0; j < m + 1; ++j) memo[0][j] = (u128)j;
    for (size_t i = 1; i < n + 1; ++i)
        for (size_t j = 1; j < m + 1; ++j) {
            u128 minimum = std::min(memo[i - 1][j - 1] + (u128)(s1[i - 1] != s2[j - 1]),
                                    std::min(memo[i][j - 1] + 1, memo[i - 1][j] + 1));
            memo[i][j] = minimum;
        }
    *out = (uint64_t)memo[n][m];
    return ORC_OK;
}

int orc_edit_distance_lean(const uint8_t* s1, size_t n, const uint8_t* s2, size_t m, uint64_t* out) {
    std::vector<uint64_t> prev(m + 1), cur(m + 1);
    for (size_t j = 0; j <= m; ++j) prev[j] = j;
    for (size_t i = 1; i <= n; ++i) {
        cur[0] = i;
        for (size_t j = 1; j <= m; ++j)
            cur[j] = std::min(prev[j - 1] + (uint64_t)(s1[i - 1] != s2[j - 1]), std::min(cur[j - 1] + 1, prev[j] + 1));
        std::swap(prev, cur);
    }
    *out = prev[m];
    return ORC_OK;
}

double orc_align_batch(int mode, const uint8_t* residues, const uint64_t* seq_off, uint64_t n_pairs,
                       int scorer, const int32_t* table, int32_t a, int32_t b, int n_threads, int lean,
                       int32_t* score, uint8_t* status, uint64_t* len, uint64_t* hash,
                       uint8_t* arena, const uint64_t* out_off) {
    ScoreFn fn = pick_scorer(scorer);
    return run_threads(n_pairs, n_threads, [&](int, uint64_t lo, uint64_t hi) {
        const bool fresh = (lean & 2) != 0; const bool lean_form = (lean & 1) != 0;
        orc_aligner* al = lean_form ? nullptr : new orc_aligner();
        std::vector<uint8_t> o1, o2;
        for (uint64_t p = lo; p < hi; ++p) {
            if (!lean_form && fresh && p != lo) { delete al; al = new orc_aligner(); }
            Seq s1{residues + seq_off[2 * p], (size_t)(seq_off[2 * p + 1] - seq_off[2 * p])};
            Seq s2{residues + seq_off[2 * p + 1], (size_t)(seq_off[2 * p + 2] - seq_off[2 * p + 1])};
            int32_t sc = 0; o1.clear(); o2.clear();
            int st = guarded([&] {
                return lean_form ? lean_run(mode, s1, s2, fn, table, a, b, &sc, o1, o2)
                            : al->run(mode, s1, s2, fn, table, a, b, &sc, o1, o2);
            });
            if (st == ORC_PANIC || st == ORC_HANG) {
                // a panicking aligner is gone in Rust; start over with a fresh one
                if (!lean_form) { delete al; al = new orc_aligner(); }
                o1.clear(); o2.clear();
            } else if (!lean_form && (al->R != 1024 || al->C != 1024)) {
                // keep "fresh aligner" semantics across pairs (A.6: dims persist across calls)
                delete al; al = new orc_aligner();
            }
            if (score) score[p] = sc;
            if (status) status[p] = (uint8_t)st;
            if (len) len[p] = o1.size();
            if (hash) hash[p] = fnv1a(fnv1a(14695981039346656037ull, o1), o2);
            if (arena && out_off && st == ORC_OK) {
                size_t cap = s1.n + s2.n;
                if (!o1.empty()) {
                    memcpy(arena + out_off[p], o1.data(), o1.size());
                    memcpy(arena + out_off[p] + cap, o2.data(), o2.size());
                }
            }
        }
        delete al;
    });
}

// analysis::seq::hamming_distance, src/analysis/seq.rs:74-83: length check, then zip + filter(a != b) + count.
int orc_hamming_distance(const uint8_t* s1, size_t n, const uint8_t* s2, size_t m, uint64_t* out) {
    if (n != m) return ORC_ERR_SIZE;                    // seq.rs:81
    uint64_t c = 0;
    for (size_t x = 0; x < n; ++x) if (s1[x] != s2[x]) ++c;   // seq.rs:76-80
    *out = c;
    return ORC_OK;
}

// analysis::stat::p_distance_matrix, src/analysis/stat.rs:138-152.  The reference visits every ordered pair
// (i, j), counts mismatches over the zip of the two rows (zip ends with the shorter row), and stores
// count as f32 / columns as f32 into both [i][j] and [j][i]; `columns` is the length of row 0 (tile.rs:31-33).
int orc_p_distance_matrix(const uint8_t* residues, const uint64_t* seq_off, uint64_t rows, float* out) {
    if (rows == 0) return ORC_PANIC;                    // self.data[0] (tile.rs:32)
    const float columns = (float)(seq_off[1] - seq_off[0]);
    for (uint64_t x = 0; x < rows * rows; ++x) out[x] = 0.0f;   // Array2::zeros (stat.rs:140)
    for (uint64_t i = 0; i < rows; ++i) {
        const uint8_t* a = residues + seq_off[i]; const uint64_t la = seq_off[i + 1] - seq_off[i];
        for (uint64_t j = 0; j < rows; ++j) {
            const uint8_t* b = residues + seq_off[j]; const uint64_t lb = seq_off[j + 1] - seq_off[j];
            float p_dist = 0.0f;                                   // stat.rs:143
            if (i != j) {
                uint64_t c = 0;
                const uint64_t len = la < lb ? la : lb;
                for (uint64_t x = 0; x < len; ++x) if (a[x] != b[x]) ++c;
                p_dist = (float)c;                                 // `.count() as f32` (stat.rs:145)
            }
            out[i * rows + j] = p_dist / columns;                  // stat.rs:147
            out[j * rows + i] = p_dist / columns;                  // stat.rs:148
        }
    }
    return ORC_OK;
}

double orc_edit_distance_batch(const uint8_t* residues, const uint64_t* seq_off, uint64_t n_pairs,
                               int n_threads, int lean, uint64_t* out) {
    return run_threads(n_pairs, n_threads, [&](int, uint64_t lo, uint64_t hi) {
        for (uint64_t p = lo; p < hi; ++p) {
            const uint8_t* s1 = residues + seq_off[2 * p]; size_t n = seq_off[2 * p + 1] - seq_off[2 * p];
            const uint8_t* s2 = residues + seq_off[2 * p + 1]; size_t m = seq_off[2 * p + 2] - seq_off[2 * p + 1];
            uint64_t d = 0;
            if (lean) orc_edit_distance_lean(s1, n, s2, m, &d); else orc_edit_distance(s1, n, s2, m, &d);
            out[p] = d;
        }
    });
}

int orc_hw_threads(void) {
    unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}

}  // extern "C"
