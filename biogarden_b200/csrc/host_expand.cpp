// host_expand.cpp -- host side of the compact result path: 2-bit alignment ops -> the two aligned strings.
//
// The traceback walk (k3_walk.cuh) records WHAT the reference's backtrack() emitted, one 2-bit op per alignment
// column: 0 = both residues (aligner.rs:531-536), 1 = seq1 residue over a gap (:537-541, :561-565), 2 = gap over
// seq2 residue (:542-547, :578-582).  Only ops, lengths and start cells travel device -> host (about 0.3 B per
// column instead of 2 B); the strings are materialised here, next to the caller's residues, straight into the
// result arena.  D2H is the scarce direction on the 8-GPU box (profiles/pcie_roof_r02.json: 93 GB/s for all
// eight GPUs together), which is why this exists.
//
// Two implementations, chosen once at run time:
//   * AVX-512 VBMI2: 64 columns per iteration -- the op planes become two 64-bit masks (BMI2 pext) and each
//     string is one byte-expand (vpexpandb) of the next residues into the columns that consume one, '-' elsewhere;
//   * portable scalar code (any x86-64 / any host).
// Both are exact: a column is a residue copy or a '-', nothing is computed.
#include "../../include/bgalign.h"

#include <cstdlib>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#define BG_X86 1
#endif

namespace bg {

// ops: word w holds ops 16w .. 16w+15, op q at bits [2(q & 15), +2); q = 0 is the first alignment column.
// s1 / s2 point at the first residue the alignment consumes (seq1[first_a], seq2[first_b]).
static void expand_scalar(const uint8_t* s1, const uint8_t* s2, const uint32_t* ops, uint64_t len, uint8_t* a_out, uint8_t* b_out) {
    uint64_t ia = 0, ib = 0;
    for (uint64_t q = 0; q < len; ++q) {
        const uint32_t op = (ops[q >> 4] >> ((q & 15u) * 2u)) & 3u;
        a_out[q] = (op != 2u) ? s1[ia++] : (uint8_t)'-';
        b_out[q] = (op != 1u) ? s2[ib++] : (uint8_t)'-';
    }
}

#ifdef BG_X86
__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi2,bmi2,popcnt")))
static void expand_vbmi2(const uint8_t* s1, const uint8_t* s2, const uint32_t* ops, uint64_t len, uint8_t* a_out, uint8_t* b_out) {
    const __m512i dash = _mm512_set1_epi8('-');
    const uint64_t EVEN = 0x5555555555555555ull;
    uint64_t q = 0;
    const uint64_t nwords = (len + 15) >> 4;
    for (; q < len; q += 64) {
        const uint64_t w = q >> 4;      // 4 ops words = 64 columns; the tail may have fewer
        uint64_t lo = 0, hi = 0;
        if (w + 4 <= nwords) {
            memcpy(&lo, ops + w, 8); memcpy(&hi, ops + w + 2, 8);
        } else {
            uint32_t t[4] = {0, 0, 0, 0};
            for (uint64_t x = 0; w + x < nwords; ++x) t[x] = ops[w + x];
            lo = (uint64_t)t[0] | ((uint64_t)t[1] << 32); hi = (uint64_t)t[2] | ((uint64_t)t[3] << 32);
        }
        const uint64_t a_only = _pext_u64(lo, EVEN) | (_pext_u64(hi, EVEN) << 32);        // op == 1
        const uint64_t b_only = _pext_u64(lo, EVEN << 1) | (_pext_u64(hi, EVEN << 1) << 32);   // op == 2
        const uint64_t left = len - q;
        const __mmask64 valid = left >= 64 ? ~0ull : ((1ull << left) - 1ull);
        const __mmask64 ma = ~b_only & valid, mb = ~a_only & valid;   // columns that consume a seq1 / seq2 residue
        const unsigned ca = (unsigned)_mm_popcnt_u64(ma), cb = (unsigned)_mm_popcnt_u64(mb);
        const __mmask64 la = ca >= 64 ? ~0ull : ((1ull << ca) - 1ull), lb = cb >= 64 ? ~0ull : ((1ull << cb) - 1ull);
        const __m512i ra = _mm512_maskz_loadu_epi8(la, s1), rb = _mm512_maskz_loadu_epi8(lb, s2);   // masked: never reads past the sequence
        _mm512_mask_storeu_epi8(a_out + q, valid, _mm512_mask_expand_epi8(dash, ma, ra));
        _mm512_mask_storeu_epi8(b_out + q, valid, _mm512_mask_expand_epi8(dash, mb, rb));
        s1 += ca; s2 += cb;
    }
}
#endif

typedef void (*expand_fn)(const uint8_t*, const uint8_t*, const uint32_t*, uint64_t, uint8_t*, uint8_t*);

static expand_fn pick_expand(int* kind_out) {
    int kind = 0;
    expand_fn f = expand_scalar;
#ifdef BG_X86
    const char* e = getenv("BG_EXPAND");   // "scalar" forces the portable code (tests)
    if (!(e && !strcmp(e, "scalar")) && __builtin_cpu_supports("avx512vbmi2") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("bmi2")) {
        f = expand_vbmi2; kind = 1;
    }
#endif
    if (kind_out) *kind_out = kind;
    return f;
}

// dst <- src without reading dst's cache lines first: whole 64-byte lines leave with non-temporal stores (a plain store
// makes the core fetch the line it is about to overwrite -- "read for ownership" -- which doubles the DRAM traffic of
// writing 300 MB of strings per 10^6 pairs; with eight ranks on one host that traffic is what bounds the call).
#ifdef BG_X86
__attribute__((target("avx512f")))
static void stream_copy_avx512(uint8_t* dst, const uint8_t* src, size_t n) {
    size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;
    if (head > n) head = n;
    memcpy(dst, src, head); dst += head; src += head; n -= head;
    for (; n >= 64; n -= 64, dst += 64, src += 64) _mm512_stream_si512(reinterpret_cast<__m512i*>(dst), _mm512_loadu_si512(src));
    memcpy(dst, src, n);
    _mm_sfence();
}
#endif
void stream_copy(uint8_t* dst, const uint8_t* src, size_t n) {
#ifdef BG_X86
    static const bool ok = __builtin_cpu_supports("avx512f");
    if (ok && n >= 256) { stream_copy_avx512(dst, src, n); return; }
#endif
    memcpy(dst, src, n);
}

void expand_ops(const uint8_t* s1, const uint8_t* s2, const uint32_t* ops, uint64_t len, uint8_t* a_out, uint8_t* b_out) {
    static const expand_fn f = pick_expand(nullptr);
    f(s1, s2, ops, len, a_out, b_out);
}

}  // namespace bg

extern "C" {

int bg_expand_ops(const uint8_t* seq1_from, const uint8_t* seq2_from, const uint32_t* ops, uint64_t len, uint8_t* a_out, uint8_t* b_out) {
    if (len && (!ops || !a_out || !b_out)) return BG_EINVAL_ARG;
    bg::expand_ops(seq1_from, seq2_from, ops, len, a_out, b_out);
    return BG_OK;
}

int bg_expand_ops_impl(int which, const uint8_t* seq1_from, const uint8_t* seq2_from, const uint32_t* ops, uint64_t len, uint8_t* a_out, uint8_t* b_out) {
    if (len && (!ops || !a_out || !b_out)) return BG_EINVAL_ARG;
    if (which == 0) { bg::expand_scalar(seq1_from, seq2_from, ops, len, a_out, b_out); return BG_OK; }
#ifdef BG_X86
    if (which == 1 && __builtin_cpu_supports("avx512vbmi2") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("bmi2")) {
        bg::expand_vbmi2(seq1_from, seq2_from, ops, len, a_out, b_out); return BG_OK;
    }
#endif
    return BG_EUNSUPPORTED;
}

const char* bg_expand_kind(void) {
    int kind = 0;
    bg::pick_expand(&kind);
    return kind == 1 ? "avx512-vbmi2" : "scalar";
}

}  // extern "C"
