/*
 * oracle.h -- CPU restatement of biogarden's alignment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker / the timed CPU baseline.
 *
 * Parity status: PINNED.  The restatement reproduces every known-answer test the
 * reference holds for this path (tests/golden/kat.json: 5 integration goldens with
 * full aligned strings, 5 aligner doctests, edit_distance 299 and its doctest);
 * see tests/test_oracle_golden.py.  The Rust crate itself cannot be built in this
 * image (no cargo/rustc), so there is no oracle/_ref.
 *
 * Two forms of every function:
 *   literal : same data layout and per-call passes as the reference (six full
 *             matrices owned by a reusable aligner, whole-buffer fill/argmax,
 *             scoring through a function pointer, u128 edit-distance table).
 *             This is also what bench.py times as the CPU baseline.
 *   lean    : rolling rows + 4 bit/cell trace; must agree with the literal form on
 *             the reference-defined domain; used where 15 B/cell does not fit.
 */
#ifndef BG_ORACLE_H
#define BG_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_GLOBAL = 0, ORC_LOCAL = 1, ORC_SEMIGLOBAL = 2, ORC_FITTING = 3, ORC_OVERLAP = 4 };

/* Outcome of one reference call. PANIC / HANG = the Rust code would panic (index
 * out of bounds, usize underflow) or spin forever (aligner.rs:549). */
enum { ORC_OK = 0, ORC_ERR_RANGE = 1, ORC_ERR_SIZE = 2, ORC_PANIC = 3, ORC_HANG = 4 };

/* scorer ids; ORC_SCORE_TABLE takes a dense 256x256 int32 table [s1 byte][s2 byte] */
enum { ORC_SCORE_BLOSUM62 = 0, ORC_SCORE_PAM250 = 1, ORC_SCORE_UNIT = 2, ORC_SCORE_TABLE = 3 };

typedef struct orc_aligner orc_aligner;

/* SequenceAligner::new (aligner.rs:44-55) */
orc_aligner* orc_aligner_new(void);
void orc_aligner_free(orc_aligner*);

/* The five public methods (aligner.rs:84,150,216,290,351), literal form.
 * a_out/b_out need room for n+m bytes each (cap); *out_len = aligned length. */
int orc_align(orc_aligner* al, int mode, const uint8_t* s1, size_t n, const uint8_t* s2, size_t m,
              int scorer, const int32_t* table, int32_t a, int32_t b,
              int32_t* score, uint8_t* a_out, uint8_t* b_out, size_t cap, size_t* out_len);

/* Same contract, lean form, "fresh aligner" semantics. */
int orc_align_lean(int mode, const uint8_t* s1, size_t n, const uint8_t* s2, size_t m,
                   int scorer, const int32_t* table, int32_t a, int32_t b,
                   int32_t* score, uint8_t* a_out, uint8_t* b_out, size_t cap, size_t* out_len);

/* analysis::seq::edit_distance (seq.rs:105-130): literal (full u128 table) and lean. */
int orc_edit_distance(const uint8_t* s1, size_t n, const uint8_t* s2, size_t m, uint64_t* out);
int orc_edit_distance_lean(const uint8_t* s1, size_t n, const uint8_t* s2, size_t m, uint64_t* out);

/* Batch drivers used for parity sweeps and as the timed CPU baseline.
 * Pairs follow the Tile convention: pair p = (seq 2p, seq 2p+1), sequence s =
 * residues[seq_off[s] .. seq_off[s+1]).  One literal aligner per thread, pairs
 * dealt statically.  Outputs: score[p], status[p], aligned strings in a padded
 * arena: a_align of pair p at arena[out_off[p] ..), b_align right after it at
 * arena[out_off[p] + cap_p ..) with cap_p = n_p + m_p; len[p] = aligned length.
 * lean: bit 0 = lean form, bit 1 = (literal form only) a brand-new aligner for every pair, i.e.
 * "fresh aligner" semantics even where stale buffer contents would otherwise leak (A.6).
 * arena / out_off may be NULL (then strings are hashed only): hash[p] = FNV-1a-64
 * over a_align, then b_align.  Returns wall seconds of the compute section. */
double orc_align_batch(int mode, const uint8_t* residues, const uint64_t* seq_off, uint64_t n_pairs,
                       int scorer, const int32_t* table, int32_t a, int32_t b, int n_threads, int lean,
                       int32_t* score, uint8_t* status, uint64_t* len, uint64_t* hash,
                       uint8_t* arena, const uint64_t* out_off);
/* analysis::seq::hamming_distance (seq.rs:74-83): 0 and *out = #positions with s1[x] != s2[x]; ORC_ERR_SIZE when
 * the lengths differ (Err(BioError::InvalidInputSize)). */
int orc_hamming_distance(const uint8_t* s1, size_t n, const uint8_t* s2, size_t m, uint64_t* out);
/* analysis::stat::p_distance_matrix (stat.rs:138-152) over `rows` sequences: the reference's double loop, f32
 * arithmetic and all (out: rows x rows).  Returns nonzero for rows == 0 (the reference panics on data[0]). */
int orc_p_distance_matrix(const uint8_t* residues, const uint64_t* seq_off, uint64_t rows, float* out);

double orc_edit_distance_batch(const uint8_t* residues, const uint64_t* seq_off, uint64_t n_pairs,
                               int n_threads, int lean, uint64_t* out);

int orc_hw_threads(void);

#ifdef __cplusplus
}
#endif
#endif
