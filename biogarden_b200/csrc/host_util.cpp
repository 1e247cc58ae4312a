// host_util.cpp -- host-side helpers of the C ABI that need no CUDA (include/bgalign.h "helpers for host mirrors").
#include "../../include/bgalign.h"
#include "pack_util.h"

#include <algorithm>
#include <thread>
#include <vector>

extern "C" int bg_residue_histogram(const bg_batch* in, uint64_t* hist_a, uint64_t* hist_b) {
    if (!in || !hist_a || !hist_b || (in->n_pairs && (!in->seq_off || !in->residues))) return BG_EINVAL_ARG;
    for (int i = 0; i < 256; ++i) hist_a[i] = hist_b[i] = 0;
    const uint64_t N = in->n_pairs;
    if (in->packing != BG_PACK_NONE) {   // packed batches: count codes, report them under the alphabet's bytes
        if ((in->packing != BG_PACK_2BIT && in->packing != BG_PACK_5BIT) || !in->alphabet) return BG_EINVAL_ARG;
        const uint32_t bits = in->packing;
        uint64_t ca[32] = {0}, cb[32] = {0};
        for (uint64_t p = 0; p < N; ++p)
            for (int side = 0; side < 2; ++side) {
                uint64_t* h = side ? cb : ca;
                for (uint64_t i = in->seq_off[2 * p + side]; i < in->seq_off[2 * p + side + 1]; ++i) {
                    h[bg::packed_code(in->residues, bits, i)]++;
                }
            }
        for (uint32_t c = 0; c < (1u << bits); ++c) { hist_a[in->alphabet[c]] += ca[c]; hist_b[in->alphabet[c]] += cb[c]; }
        return BG_OK;
    }
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

// ---- packed residues (bg_batch::packing) -------------------------------------------------------------------------
#include <cstring>

namespace bg {

// one packed byte -> its four residue bytes (2-bit packing)
void make_unpack_lut2(const uint8_t* alphabet, uint32_t* lut) {
    for (int b = 0; b < 256; ++b)
        lut[b] = (uint32_t)alphabet[b & 3] | ((uint32_t)alphabet[(b >> 2) & 3] << 8) | ((uint32_t)alphabet[(b >> 4) & 3] << 16) | ((uint32_t)alphabet[(b >> 6) & 3] << 24);
}

// residues [first, first + count) of a packed arena -> bytes; lut2: make_unpack_lut2's table (2-bit; nullptr: built here)
void unpack_residues(const uint8_t* packed, uint32_t bits, const uint8_t* alphabet, uint64_t first, uint64_t count, uint8_t* out, const uint32_t* lut2) {
    if (bits == BG_PACK_2BIT) {
        uint64_t i = first, end = first + count;
        for (; i < end && (i & 3); ++i) *out++ = alphabet[(packed[i >> 2] >> ((i & 3) * 2)) & 3];
        if (i + 4 <= end) {
            uint32_t local[256];
            if (!lut2) { make_unpack_lut2(alphabet, local); lut2 = local; }
            for (; i + 4 <= end; i += 4, out += 4) { const uint32_t v = lut2[packed[i >> 2]]; memcpy(out, &v, 4); }
        }
        for (; i < end; ++i) *out++ = alphabet[(packed[i >> 2] >> ((i & 3) * 2)) & 3];
        return;
    }
    for (uint64_t i = first; i < first + count; ++i) *out++ = alphabet[packed_code(packed, 5, i)];
}

}  // namespace bg

extern "C" {

uint64_t bg_packed_bytes(uint64_t n, int bits) {
    return (bits == 2 ? (n + 3) / 4 : bits == 5 ? (5 * n + 7) / 8 : n) + 16;
}

int bg_pack_residues(const uint8_t* residues, uint64_t n, int bits, int n_threads, uint8_t* packed, uint8_t* alphabet) {
    if ((n && (!residues || !packed)) || !alphabet || (bits != 2 && bits != 5)) return BG_EINVAL_ARG;
    unsigned hw = std::thread::hardware_concurrency();
    uint64_t T = n_threads > 0 ? (uint64_t)n_threads : (hw ? hw : 1);
    T = std::max<uint64_t>(1, std::min<uint64_t>(T, n / (1u << 20) + 1));
    auto run = [&](auto&& fn) {
        if (T == 1) { fn(0); return; }
        std::vector<std::thread> th;
        for (uint64_t k = 0; k < T; ++k) th.emplace_back(fn, k);
        for (auto& x : th) x.join();
    };
    uint8_t code[256];
    memset(code, 0xFF, sizeof code);
    memset(alphabet, 0, 32);
    if (bits == 2) {
        std::vector<std::vector<uint8_t>> seen(T, std::vector<uint8_t>(256, 0));
        run([&](uint64_t k) {
            const uint64_t lo = n * k / T, hi = n * (k + 1) / T;
            uint8_t* s = seen[k].data();
            for (uint64_t i = lo; i < hi; ++i) s[residues[i]] = 1;
        });
        int na = 0;
        for (int b = 0; b < 256; ++b) {
            bool any = false;
            for (uint64_t k = 0; k < T; ++k) any = any || seen[k][b];
            if (!any) continue;
            if (na == 4) return BG_EINVAL_RESIDUE;
            code[b] = (uint8_t)na; alphabet[na++] = (uint8_t)b;
        }
    } else {
        for (int c = 0; c < 26; ++c) { code['A' + c] = (uint8_t)c; alphabet[c] = (uint8_t)('A' + c); }
    }
    std::vector<int> bad(T, 0);
    const uint64_t unit = bits == 2 ? 4 : 8;      // residues per whole number of packed bytes
    memset(packed + bg_packed_bytes(n, bits) - 16, 0, 16);
    run([&](uint64_t k) {
        const uint64_t lo = (n * k / T) / unit * unit, hi = (k + 1 == T) ? n : (n * (k + 1) / T) / unit * unit;
        if (bits == 2) {
            for (uint64_t i = lo; i < hi; i += 4) {
                uint32_t v = 0;
                for (uint64_t j = 0; j < 4 && i + j < hi; ++j) { const uint8_t c = code[residues[i + j]]; if (c == 0xFF) { bad[k] = 1; return; } v |= (uint32_t)c << (2 * j); }
                packed[i >> 2] = (uint8_t)v;
            }
        } else {
            for (uint64_t i = lo; i < hi; i += 8) {
                uint64_t v = 0;
                for (uint64_t j = 0; j < 8 && i + j < hi; ++j) { const uint8_t c = code[residues[i + j]]; if (c == 0xFF) { bad[k] = 1; return; } v |= (uint64_t)c << (5 * j); }
                const uint64_t nb = std::min<uint64_t>(5, (5 * (hi - i) + 7) / 8);
                memcpy(packed + 5 * (i >> 3), &v, nb);
            }
        }
    });
    for (int b : bad) if (b) return BG_EINVAL_RESIDUE;
    return BG_OK;
}

int bg_unpack_residues(const uint8_t* packed, int bits, const uint8_t* alphabet, uint64_t first, uint64_t count, uint8_t* out) {
    if ((count && (!packed || !out)) || !alphabet || (bits != 2 && bits != 5)) return BG_EINVAL_ARG;
    bg::unpack_residues(packed, (uint32_t)bits, alphabet, first, count, out, nullptr);
    return BG_OK;
}

}  // extern "C"
