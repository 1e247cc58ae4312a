// bg_common.cuh -- shared device/host definitions of the alignment engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bg {

// "minus infinity" for the X / Y borders.  The reference uses i32::MIN with saturating_add
// (aligner.rs:49-50,443,447); all that matters is that the border never wins a max and never
// ties with a real value.  bg_api.cu refuses (BG_EUNSUPPORTED) parameter ranges for which a
// real score could come within 2^28 of it.
constexpr int32_t NEG_INF = -(1 << 29);

// Direction nibble, one per DP cell: everything backtrack() reads, as four
// independent tie bits so that the fill kernel never needs a select chain:
//   bit 0  M == X   (aligner.rs:458)      decode of m_trace: bit 1 -> 'Y', else bit 0 -> 'X', else 'R'
//   bit 1  M == Y   (aligner.rs:455)      (the reference tests Y first, so both bits set means 'Y')
//   bit 2  x_trace == 'M'  (gap in seq2 was opened here; aligner.rs:444)
//   bit 3  y_trace == 'M'  (aligner.rs:448)
// Local mode only: bit 0 is overloaded when bit 1 is set -- it then says M == 0 (STOP), which is all
// the walk's `m[k][l] > 0` test needs because M == 0 implies m_trace == 'Y' (A.3).
constexpr uint32_t TR_XEQ = 1, TR_YEQ = 2, TR_XOPEN = 4, TR_YOPEN = 8;

enum Mode : int { M_GLOBAL = 0, M_LOCAL = 1, M_SEMIGLOBAL = 2, M_FITTING = 3, M_OVERLAP = 4 };

// One launch slot = one sequence pair placed in a warp's lane group.  Built by the host in
// launch order (pairs of one length class, longest first); 64 bytes.
struct __align__(16) PairDesc {
    uint64_t a_off;      // seq1 bytes at residues + a_off
    uint64_t b_off;      // seq2 bytes
    uint64_t trace_off;  // uint32-word offset of the WARP's trace block
    uint64_t bnd_off;    // int2 offset of this pair's band-boundary column scratch (multi-band only)
    uint64_t pad_off;    // byte offset (multiple of 4) of this pair's padded output slot, 2 * round_up(n + m, 4) bytes
    uint32_t n, m;       // len1, len2
    uint32_t steps;      // systolic steps per band for this warp = max n in the warp + L - 1
    uint32_t pair_id;    // index in the caller's batch; 0xFFFFFFFF = empty slot
    uint32_t nbands;     // column bands of L*C columns
    uint32_t pad_;
};
static_assert(sizeof(PairDesc) == 64, "PairDesc layout");

// Where the walk starts and what it returns, per slot.
struct __align__(16) EndCell {
    int32_t score;
    uint32_t k, l;   // start cell
    uint32_t flags;  // bit 0: semiglobal column branch (aligner.rs:389)
};

// A traceback walk suspended at a row-block boundary (bounded-memory traceback, k2_wave.cuh): everything
// k3_walk_diag needs to go on in the next launch.
struct __align__(16) WalkState {
    uint32_t k, l, cur, pos, wops, flags;
    uint32_t started;   // 0: the first launch initialises the walk from the end cell
    uint32_t done;      // the walk has ended and its results are written
    uint64_t it;
    uint64_t pad_;
};

// Bounded-memory traceback: what one launch does for one slot (built by the host per launch).
struct __align__(16) CkptSlot {
    uint64_t ck_off;      // int2 offset of the slot's checkpoint rows
    uint32_t ck_stride;   // elements per checkpoint row (len2 rounded up to 32)
    uint32_t row0;        // the launch fills / walks DP rows row0 + 1 .. row0 + nrows (nrows == 0: nothing)
    uint32_t nrows;
    uint32_t every;       // checkpoint spacing of the slot, rows (multiple of 32)
    uint32_t pad_[2];
};
static_assert(sizeof(CkptSlot) == 32, "CkptSlot layout");

constexpr uint32_t WALK_UNDERFLOW = 1;  // reference would index seq[usize::MAX] (A.6)
constexpr uint32_t WALK_HANG = 2;       // step bound exceeded (cannot happen with well-formed traces)

// Does the REFERENCE define a result for this pair?  (restated for a fresh aligner whose
// buffers are 1024 x 1024 unless a length exceeds 1024: aligner.rs:45, 92-94, 594-595.)  0 = yes,
// 1 = the reference panics / hangs / reads outside the rectangle (BG_ST_REF_UNDEFINED); the engine's
// result is then its documented extension.  Evaluated on the device by the walk kernels.
__host__ __device__ inline int ref_status(int mode, uint64_t n, uint64_t m, int32_t score, uint32_t walk_flags) {
    uint64_t R, C;
    if (n > 1024 || m > 1024) { R = n + 1; C = m + 1; } else { R = 1024; C = 1024; }
    if (walk_flags & (WALK_UNDERFLOW | WALK_HANG)) return 1;
    const bool row_border = (mode == M_GLOBAL || mode == M_FITTING);   // writes row0[1..=m]
    const bool col_border = (mode == M_GLOBAL);                          // writes col0[1..=n]
    if (row_border && (C < 2 || m >= C)) return 1;
    if (col_border && (R < 2 || n >= R)) return 1;
    if (n >= 1 && m >= 1 && (n >= R || m >= C)) return 1;                          // fill indexes [n][m]
    if ((mode == M_OVERLAP || mode == M_SEMIGLOBAL) && n >= R) return 1;            // .row(len1)
    if ((mode == M_FITTING || mode == M_SEMIGLOBAL) && m >= C) return 1;            // .column(len2)
    // whole-buffer scans see the zeroed cells outside the rectangle (aligner.rs:247,308,369,376)
    if ((mode == M_SEMIGLOBAL || mode == M_OVERLAP) && score == 0 && C > m + 1) return 1;
    if (mode == M_FITTING && score < 0 && R > n + 1) return 1;
    return 0;
}

// Geometry of the trace block of a warp: word (t, k, lane) of band bd lives at
//   trace_off + ((bd * steps + t) * K + k) * 32 + lane,   K = ceil(C / 8) words per lane-step.
// Every warp-wide store of one k is a fully coalesced 128-byte line.
__host__ __device__ inline uint32_t words_per_lane_step(int C) { return (uint32_t)((C + 7) / 8); }

}  // namespace bg
