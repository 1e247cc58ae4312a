/*
 * bgsynth.h -- deterministic synthetic workloads for tests and bench.py (libbgsynth.so).
 *
 * Benchmark / test tooling, NOT part of the product library: libbgalign.so neither exports nor uses it, so the
 * reference arm of bench.py (CPU oracle) maps only oracle/ and this generator.
 */
#ifndef BGSYNTH_H
#define BGSYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Deterministic synthetic workloads (SURVEY 8d generator: splitmix64 stream; a uniform over
 * the alphabet, b = a mutated with P(sub)=.05, P(ins)=.01, P(del)=.01 for 90% of pairs,
 * independent uniform for the rest).  Lengths uniform in [len_lo, len_hi]; if fix_b_len != 0,
 * b is truncated / padded with random residues to exactly the drawn length of its own.
 * Call with residues == NULL to size: returns total residues in *n_residues. */
int bg_synth_pairs(uint64_t seed, uint64_t first_pair, uint64_t n_pairs, const char* alphabet, int alphabet_len,
                   uint32_t len_lo, uint32_t len_hi, int resize_b,
                   uint8_t* residues, uint64_t* seq_off, uint64_t* n_residues);

#ifdef __cplusplus
}
#endif
#endif /* BGSYNTH_H */
