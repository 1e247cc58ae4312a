/*
 * bgalign.h -- C ABI of the B200 batched pairwise-alignment engine (libbgalign.so).
 *
 * This is the drop-in boundary for biogarden's alignment hot path.  The reference
 * (robsndr/biogarden, Rust) has no FFI today; each entry point below names the
 * reference interface it stands behind (file:line in the reference checkout).  A Rust
 * caller binds these with a plain `extern "C"` block (INTEGRATION.md shows the stub and
 * the patched SequenceAligner methods); this repository drives the same symbols from C++
 * (include/biogarden.hpp) and Python ctypes (biogarden_b200/native.py).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types cross this boundary;
 *   - inputs are borrowed for the duration of a call and never written;
 *   - outputs are library-owned pinned host memory until bg_result_free();
 *   - every function returns a bg_err code (0 = ok); nothing panics or hangs across the ABI;
 *   - a bg_ctx is used by one caller thread at a time (mirrors `&mut self`, aligner.rs:84);
 *   - all arithmetic runs in CUDA kernels built for sm_100a.  There is no CPU path: with no
 *     usable device bg_create() fails with BG_ENODEVICE.
 */
#ifndef BGALIGN_H
#define BGALIGN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BG_API_VERSION 2

typedef struct bg_ctx bg_ctx;

/* Alignment flavour = which public SequenceAligner method is being replaced:
 *   BG_GLOBAL     global_alignment      aligner.rs:84-121
 *   BG_LOCAL      local_alignment       aligner.rs:150-185
 *   BG_SEMIGLOBAL semiglobal_alignment  aligner.rs:351-435
 *   BG_FITTING    fitting_alignment     aligner.rs:216-260
 *   BG_OVERLAP    overlap_alignment     aligner.rs:290-321 */
typedef enum { BG_GLOBAL = 0, BG_LOCAL = 1, BG_SEMIGLOBAL = 2, BG_FITTING = 3, BG_OVERLAP = 4 } bg_mode;

/* Call-level return codes.  BG_EINVAL_RANGE / BG_EINVAL_SIZE map 1:1 onto
 * BioError::InvalidArgumentRange / InvalidInputSize (error.rs:9-10) and are returned under
 * exactly the reference's conditions (aligner.rs:87-89,153-155,219-225; semiglobal and overlap
 * do NOT check the sign of the penalties, aligner.rs:290-296,351-357). */
typedef enum {
    BG_OK = 0,
    BG_EINVAL_RANGE = 1,   /* a > 0 || b > 0 in global / local / fitting */
    BG_EINVAL_SIZE = 2,    /* fitting with len1 < len2 in some pair; odd Tile length */
    BG_ECUDA = 3,          /* CUDA runtime / launch failure (bg_last_error has the text) */
    BG_ENOMEM = 4,         /* host or device allocation failed */
    BG_EINVAL_ARG = 5,     /* NULL pointer, unknown mode, malformed offsets ... */
    BG_EINVAL_RESIDUE = 6, /* a residue byte has no row/column in the score table (the shipped
                              scorers panic on such bytes, score.rs:40,79,115) */
    BG_ENODEVICE = 7,      /* no CUDA device / library built without the requested device */
    BG_EUNSUPPORTED = 8,   /* penalties or scores outside the engine's 32-bit-safe range */
    BG_EINVAL_FASTA = 9    /* "Expected > at record start." (io/fasta.rs:104-109) */
} bg_err;

/* Per-pair status.  BG_ST_REF_UNDEFINED: the reference itself panics or never returns on
 * this input; score / strings are then the engine's documented
 * extension (the recurrence restricted to the (len1+1)x(len2+1) rectangle). */
typedef enum { BG_ST_OK = 0, BG_ST_REF_UNDEFINED = 1 } bg_status;

/* A batch of sequence pairs in Tile order (ds/tile.rs:9-11): sequence s occupies
 * residues[seq_off[s] .. seq_off[s+1]); pair p = (sequence 2p, sequence 2p+1), the layout of
 * every alignment fixture (tests/integration.rs:236-242) and of examples/from_file.rs:26-27.
 * Raw bytes, exactly what Sequence::chain holds (ds/sequence.rs:10-13). */
typedef struct {
    uint64_t n_pairs;
    const uint8_t* residues;
    const uint64_t* seq_off; /* 2*n_pairs + 1 entries, non-decreasing; residue indices in every packing */
    /* Packed residues (what bg_fasta_parse_packed / bg_pack_residues produce): 4x / 1.6x fewer bytes over the host
     * link and in host memory; the device unpacks them with 128-bit loads before the fill kernels read them.
     *   BG_PACK_NONE  one byte per residue (Sequence::chain, ds/sequence.rs:10-13)
     *   BG_PACK_2BIT  residue i = code at bits [2 (i & 3), +2) of residues[i >> 2]               (<= 4 letters: DNA)
     *   BG_PACK_5BIT  residue i = code at bits [5 i, 5 i + 5) of the little-endian bit stream     (<= 32 letters: protein)
     * alphabet[code] is the residue byte the code stands for (4 resp. 32 entries); results are identical to the
     * unpacked batch's.  Zero-initialised trailing fields mean BG_PACK_NONE (API version 1 callers). */
    uint32_t packing;
    uint32_t reserved_;
    const uint8_t* alphabet;
} bg_batch;
#define BG_PACK_NONE 0u
#define BG_PACK_2BIT 2u
#define BG_PACK_5BIT 5u

#define BG_F_SCORE_ONLY 1u /* skip traceback: result.arena / off stay empty */

/* Parameters of one alignment call.  The score callback of the reference
 * (`&dyn Fn(&u8,&u8)->i32`, aligner.rs:85) cannot cross the ABI; the caller materialises it
 * once, on its own thread, over the residues present in the batch:
 *   table[row_code[x] * n_cols + col_code[y]] = score(x, y),  x from a seq1, y from a seq2.
 * Argument order is (seq1 residue, seq2 residue) as in aligner.rs:451; no symmetry is assumed.
 * Codes 0xFF mark bytes that must not occur. */
typedef struct {
    int32_t mode;       /* bg_mode */
    int32_t gap_open;   /* `a` */
    int32_t gap_extend; /* `b` */
    uint32_t flags;
    const int32_t* table;
    int32_t n_rows, n_cols;
    const uint8_t* row_code; /* [256] */
    const uint8_t* col_code; /* [256] */
} bg_params;

/* Result of bg_align_batch: what the reference returns per call, `(i32, Sequence, Sequence)`
 * (aligner.rs:85), for every pair.  a_align of pair p = arena[off[2p] .. off[2p+1]),
 * b_align = arena[off[2p+1] .. off[2p+2]); both have the same length. */
typedef struct {
    uint64_t n_pairs;
    int32_t* score;  /* [n_pairs] */
    uint8_t* status; /* [n_pairs] bg_status */
    uint8_t* arena;
    uint64_t* off;   /* [2*n_pairs + 1] */
    void* owner_;    /* private */
} bg_result;

/* Phase timings of the most recent device pass, measured with CUDA events on the engine's
 * own stream (milliseconds; kernels only, no host<->device copies). */
typedef struct {
    double encode_ms;   /* residue validation + recoding */
    double fill_ms;     /* DP fill kernels (K1/K2) incl. end-cell selection */
    double walk_ms;     /* traceback walk (K3) */
    double compact_ms;  /* aligned-string compaction */
    double total_ms;    /* first launch -> last kernel end */
    uint64_t cells;     /* sum over pairs of len1*len2 */
    uint64_t launches;  /* kernels launched */
    uint64_t trace_bytes; /* bytes of packed direction codes written */
    uint64_t h2d_bytes, d2h_bytes; /* of the last host-buffer call */
    uint64_t cells_packed16;    /* of `cells`: filled by the packed 16 x 2 kernel (two cells per lane instruction) */
    uint64_t cells_bitparallel; /* of `cells`: edit distance by the bit-parallel kernel (32 cells per word column) */
    uint64_t fill_launches;     /* DP fill kernel launches behind fill_ms */
    uint64_t cells_refilled;    /* of `cells`: pairs aligned with bounded-memory traceback (their cells are filled twice) */
} bg_timing;

/* ---- lifetime ---------------------------------------------------------------------- */
/* SequenceAligner::new (aligner.rs:44-55).  devices = CUDA ordinals to shard batches over
 * (NULL / n_dev == 0 -> the current device only). */
int bg_create(const int* devices, int n_dev, bg_ctx** out);
void bg_destroy(bg_ctx* ctx);
const char* bg_strerror(int err);
const char* bg_last_error(const bg_ctx* ctx);
int bg_version(void);

/* ---- the hot path, host buffers (what a drop-in caller uses) ---------------------- */
/* global/local/semiglobal/fitting/overlap_alignment for every pair of the batch
 * (aligner.rs:84,150,216,290,351).  H2D of the batch, all kernels, D2H of the results. */
int bg_align_batch(bg_ctx* ctx, const bg_batch* in, const bg_params* p, bg_result* out);
void bg_result_free(bg_result* r);

/* ---- the hot path, compact results ---------------------------------------------------- */
/* The same alignments as bg_align_batch, but as WHAT backtrack() emitted instead of the emitted characters: one
 * 2-bit op per alignment column -- 0: both residues (aligner.rs:531-536), 1: seq1 residue over '-' (:537-541,
 * :561-565), 2: '-' over seq2 residue (:542-547, :578-582) -- 16 ops per little-endian word, column q of pair p at
 * bits [2 (q & 15), +2) of ops[ops_off[p] + (q >> 4)], column 0 first.  The strings are
 *     a_align = expand(seq1 from first[2p], ops), b_align = expand(seq2 from first[2p+1], ops), both len[p] long,
 * which bg_expand_ops does (AVX-512 VBMI2 when the host has it).  A shim that owns its containers expands every
 * pair straight into them; bg_align_batch does exactly that into its arena.  D2H traffic is ~0.3 B per column
 * instead of 2 B.  ops_off leaves gaps between pipeline chunks; ops_off[n_pairs] = size of ops in words. */
typedef struct {
    uint64_t n_pairs;
    int32_t* score;    /* [n_pairs] */
    uint8_t* status;   /* [n_pairs] bg_status */
    uint32_t* len;     /* [n_pairs] aligned length in columns */
    uint32_t* first;   /* [2 * n_pairs] index of the first residue of seq1 / seq2 inside the alignment */
    uint32_t* ops;
    uint64_t* ops_off; /* [n_pairs + 1] word offsets into ops */
    void* owner_;      /* private */
} bg_ops_result;
int bg_align_batch_ops(bg_ctx* ctx, const bg_batch* in, const bg_params* p, bg_ops_result* out);
void bg_ops_result_free(bg_ops_result* r);
/* Expands `len` ops into a_out / b_out (len bytes each).  seq1_from / seq2_from point at the first residue the
 * alignment consumes, i.e. residues + seq_off[2p] + first[2p] resp. residues + seq_off[2p+1] + first[2p+1]. */
int bg_expand_ops(const uint8_t* seq1_from, const uint8_t* seq2_from, const uint32_t* ops, uint64_t len,
                  uint8_t* a_out, uint8_t* b_out);
/* "avx512-vbmi2" or "scalar": which implementation bg_expand_ops uses on this host. */
const char* bg_expand_kind(void);

/* analysis::seq::edit_distance for every pair (seq.rs:105-130): out[p] = distance.  Never
 * fails on any byte content (the reference never errs).  Batches of <= 4 distinct bytes whose second sequences are
 * <= 320 long (read sets, raw or 2-bit packed) take the bit-parallel kernel with launch slots built on the device: the
 * host then does not read seq_off at all; offsets are validated on the device, and a batch it cannot take (a longer
 * pair, a fifth byte value, offsets that are not monotone -> BG_EINVAL_ARG as always) is redone by the general path. */
int bg_edit_distance_batch(bg_ctx* ctx, const bg_batch* in, uint64_t* out);

/* ---- before the hot path: FASTA text -> batch layout -------------------- */
/* io::fasta::Reader::read_all (fasta.rs:95-136) over a text held in memory, straight into the layout bg_batch
 * takes: record r has residues[seq_off[r] .. seq_off[r+1]) and id ids[id_off[r] .. id_off[r+1]).  An alignment
 * batch of pairs (2p, 2p+1) is then {n_records / 2, residues, seq_off}.  Library-allocated (malloc), freed by
 * bg_fasta_free.  n_threads <= 0: all host cores.  Host code; needs no GPU. */
typedef struct bg_fasta {
    uint64_t n_records;
    uint8_t* residues;
    uint64_t* seq_off;    /* [n_records + 1] */
    uint8_t* ids;
    uint64_t* id_off;     /* [n_records + 1] */
    uint32_t packing;     /* BG_PACK_*: how `residues` is stored (bg_fasta_parse: BG_PACK_NONE) */
    uint32_t reserved_;
    uint8_t alphabet[32]; /* packing != BG_PACK_NONE: code -> residue byte */
} bg_fasta;
int bg_fasta_parse(const uint8_t* text, uint64_t len, int n_threads, bg_fasta* out);
/* The same records with the residues packed (bits = 2 or 5) -- the batch {n_records / 2, residues, seq_off, bits,
 * alphabet} goes straight to bg_align_batch / bg_edit_distance_batch / bg_batch_upload.  bits = 2: the text may use
 * at most 4 distinct residue bytes (codes in ascending byte order); bits = 5: 'A'..'Z' (code = byte - 'A', the
 * index the shipped scorers use, score.rs:40).  Anything else: BG_EINVAL_RESIDUE. */
int bg_fasta_parse_packed(const uint8_t* text, uint64_t len, int n_threads, int bits, bg_fasta* out);
void bg_fasta_free(bg_fasta* f);

/* Packing helpers for callers that hold bytes (a shim packs while it marshals its Tile).  bg_pack_residues: n residue
 * bytes -> packed (capacity bg_packed_bytes(n, bits)); alphabet[32] is an output for bits = 2 (chosen from the data)
 * and for bits = 5 (always 'A' + code).  bg_unpack_residues: residues [first, first + count) of a packed arena -> bytes. */
uint64_t bg_packed_bytes(uint64_t n_residues, int bits);
int bg_pack_residues(const uint8_t* residues, uint64_t n, int bits, int n_threads, uint8_t* packed, uint8_t* alphabet);
int bg_unpack_residues(const uint8_t* packed, int bits, const uint8_t* alphabet, uint64_t first, uint64_t count, uint8_t* out);

/* ---- next to the hot path: position-wise compares ----------------------- */
/* analysis::seq::hamming_distance for every pair (seq.rs:74-83): out[p] = #positions where the two sequences
 * differ.  BG_EINVAL_SIZE if any pair has len1 != len2 (the reference's Err(InvalidInputSize)). */
int bg_hamming_distance_batch(bg_ctx* ctx, const bg_batch* in, uint64_t* out);
/* analysis::stat::p_distance_matrix (stat.rs:138-152) of `rows` sequences (row r = residues[seq_off[r] ..
 * seq_off[r+1])): out is rows x rows f32, row-major; out[i][j] = mismatches over the zip of rows i and j, as
 * f32, divided by (len(row 0) as f32); 0 on the diagonal.  rows == 0 -> BG_EINVAL_SIZE (the reference panics). */
int bg_p_distance_matrix(bg_ctx* ctx, const uint8_t* residues, const uint64_t* seq_off, uint64_t rows, float* out);

/* ---- the hot path, device-resident (benchmarks: inputs already in HBM) ------------- */
typedef struct bg_dbatch bg_dbatch;   /* batch resident on one device of the context */
typedef struct bg_dresult bg_dresult; /* results resident on that device */

int bg_batch_upload(bg_ctx* ctx, int dev_index, const bg_batch* in, bg_dbatch** out);
void bg_dbatch_free(bg_dbatch* b);
/* Runs encode + fill + walk + compaction on the batch's device, asynchronously on the
 * context's stream for that device; *out stays on the device. */
int bg_align_device(bg_ctx* ctx, const bg_dbatch* in, const bg_params* p, bg_dresult** out);
int bg_edit_distance_device(bg_ctx* ctx, const bg_dbatch* in, bg_dresult** out);
int bg_dresult_download(bg_ctx* ctx, bg_dresult* r, bg_result* out); /* syncs */
int bg_dresult_download_u64(bg_ctx* ctx, bg_dresult* r, uint64_t* out);
void bg_dresult_free(bg_dresult* r);
int bg_sync(bg_ctx* ctx);
/* cudaStream_t of device dev_index (so a harness can bracket calls with its own events) */
void* bg_stream(bg_ctx* ctx, int dev_index);
int bg_device_ordinal(bg_ctx* ctx, int dev_index);
int bg_last_timing(const bg_ctx* ctx, bg_timing* out);
/* Build the launch plan (length classes, trace layout) of an uploaded batch now instead of
 * inside the first align / edit-distance call. */
int bg_batch_prepare(bg_ctx* ctx, bg_dbatch* b, int for_edit);
/* Tuning knobs (tests, experiments): force one kernel shape (lanes per pair, columns per lane;
 * 0,0 = automatic by length class) and bound the per-launch trace buffer. */
int bg_set_shape(bg_ctx* ctx, int lanes_per_pair, int cols_per_lane);
int bg_set_trace_budget(bg_ctx* ctx, uint64_t bytes);
/* Trace memory one launch of the long-pair path (pairs wider than 4096 columns) may use; 0 = automatic (80 % of
 * the device).  A pair whose 0.5 B/cell trace does not fit is aligned with BOUNDED-MEMORY traceback: row
 * checkpoints in a first pass, then re-fill + walk block by block (results are identical; the cells are
 * computed twice).  This replaces the reference's "six full matrices or nothing" (aligner.rs:594-602). */
int bg_set_long_trace_budget(bg_ctx* ctx, uint64_t bytes);

/* Launches of the long-pair path (pairs wider than 4096 columns) with at most `max_pairs` pairs use the fine-grained
 * wavefront kernel -- one column per lane, a lone 10 kbp pair spread over ~270 warps -- instead of 512-column bands
 * (0 = never, the default: the kernel is bit-exact but not yet faster than the banded one, see k2f_fine.cuh).
 * Results are identical either way (tests force both). */
int bg_set_fine_pairs(bg_ctx* ctx, int max_pairs);

/* 1: build every launch plan on the host; 0 (default): the chunks of bg_align_batch are planned on the device from
 * their sequence offsets (16 B per pair H2D instead of 64 B of descriptors, no per-pair host work). */
int bg_set_host_plan(bg_ctx* ctx, int on);

/* Page-lock / unlock caller memory (cudaHostRegister).  The host-buffer entry points read the caller's residue
 * arena directly: pinned, that is an asynchronous DMA overlapped with the kernels; pageable, the driver stages it
 * synchronously.  Results are always returned in library-owned pinned memory. */
int bg_pin_host(const void* ptr, uint64_t bytes);
int bg_unpin_host(const void* ptr);

/* ---- helpers for host mirrors ------------------------------------------------------ */
/* The shipped scorers' 26x26 tables, indexed [a-'A'][b-'A'] (score.rs:5-35,45-75,82-111).
 * name = "blosum62" | "pam250" | "unit"; returns NULL for anything else. */
const int8_t* bg_score_table26(const char* name);

/* Byte histograms of the seq1 / seq2 sides of a batch (which residues a score callback has to
 * be evaluated on).  hist_a / hist_b: [256]. */
int bg_residue_histogram(const bg_batch* in, uint64_t* hist_a, uint64_t* hist_b);

/* Whether the REFERENCE defines a result for (mode, len1, len2) given the outcome the engine
 * computed; used by host mirrors that want reference-identical failure behaviour. Returns a
 * bg_status. (fresh-aligner semantics.) */
int bg_ref_status(int mode, uint64_t len1, uint64_t len2, int32_t score, int walk_flags);

#ifdef __cplusplus
}
#endif
#endif /* BGALIGN_H */
