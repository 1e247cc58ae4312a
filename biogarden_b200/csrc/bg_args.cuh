// bg_args.cuh -- kernel argument structs and launch constants shared by the host side (bg_api.cu) and the
// kernel translation units (l_*.cu); the kernels themselves live in k*.cuh, each included by exactly one l_*.cu.
#pragma once
#include "bg_common.cuh"

namespace bg {

struct FillArgs {
    const PairDesc* desc;
    uint32_t n_slots;
    const uint8_t* residues;
    const int32_t* table;      // n_rows x n_cols (device)
    int32_t n_rows, n_cols;
    const uint8_t* row_code;   // [256] (device)
    const uint8_t* col_code;   // [256]
    int32_t a, b;              // gap open / extend
    int32_t mode;
    int32_t want_trace;
    uint32_t* trace;
    int2* bnd;
    EndCell* end;
    uint32_t* err_flag;        // bit 0: residue without a table row/column
    int32_t one;               // == 1 at run time; keeps the FMA-pipe adds below as IMADs (see k1_fill)
    int32_t tg_shift;          // K1h trace tiling (k1h_fill.cuh)
};

// ---- K0 (k0_plan.cuh): launch descriptors built on the device ----
// The host only classifies (it reads the offsets once anyway to validate them) and sizes; which pair sits in which
// launch slot, the K1h holes, the trace / output / scratch offsets are computed on the GPU from the offsets alone.
constexpr int PLAN_MAX_CLS = 16;
struct PlanCls {
    int32_t L, C;                    // kernel shape of the class
    uint32_t half;                   // 1: K1h -- two pairs per lane group with equal row counts, so runs of equal
                                     //    len1 are padded to even length with empty slots ("holes")
    uint32_t G2;                     // slots per warp
    uint32_t sorted_begin, count;    // the class's range in the sorted pair order (class, then len1 descending)
    uint32_t slot_begin, slot_cap;   // its descriptor range; cap = upper bound incl. holes, a multiple of G2
    uint64_t pad_base;               // bytes: exact (the host sums the slot sizes per class)
    uint64_t bnd_base;               // int2 elements: exact
};
struct PlanArgs {
    const uint64_t* off;             // [2n + 1] sequence offsets of the item (device copy of the caller's slice)
    uint64_t base;                   // off[0]: the item's residues start at device offset 0
    uint32_t n_pairs, n_cls;
    uint32_t half_ok;                // as in pick_shape_m
    int32_t force_si;                // >= 0: every pair uses this shape (bg_set_shape)
    uint32_t uniform;                // 1: one class and every pair has the same len1 and the same len2 (read sets): descriptors are a
                                     //    closed form of the pair index, written by ONE kernel (k_plan_uniform) instead of ten launches
    uint64_t* keys;                  // [2][n] sort keys: class rank << 32 | (0x7fffffff - len1)  (set by launch_plan)
    uint32_t* ids;                   // [2][n] pair ids
    PairDesc* desc;
    int8_t rank_of_shape[16];        // shape number -> class rank in cls[] (-1: not present)
    PlanCls cls[PLAN_MAX_CLS];
};

struct UnpackArgs {
    const uint8_t* packed;    // device copy of the packed bytes, 16-byte aligned
    uint64_t bit0;            // bit offset of the first residue to unpack
    uint64_t count;           // residues
    uint32_t bits;            // 2 or 5
    uint8_t alphabet[32];     // code -> residue byte
    uint8_t* out;             // [count]
};

// ---- K1h (k1h_fill.cuh) ----
constexpr int32_t HB_BIAS = 1 << 15;
constexpr int32_t HB_NEG = 1 << 8;       // biased "minus infinity": below every value a cell can take (>= 2^15 - 2*HB_RANGE - 3*HB_MAXABS)
constexpr int32_t HB_RANGE = 15000;      // max (len1 + len2 + 2) * max|score| the host admits: |H| and the frame shift each stay below it
constexpr int32_t HB_MAXABS = 512;       // max |a|, |b|, |s|
constexpr int HB_TB = 4;                 // systolic steps per trace row block
constexpr int HB_TG_MAX = 4;             // FillArgs::tg_shift <= 2: 2^tg_shift row blocks of one lane are stored back to back
__host__ __device__ inline uint32_t hb_words_per_lane_block(int C) { return (uint32_t)((C + 3) & ~3); }

// ---- K2 (k2_wave.cuh) ----
struct WaveCand {          // one worker's end-cell candidates (SURVEY A.5 tie rules applied when merged)
    int32_t best; uint32_t bi, bj;        // local
    int32_t rbest; uint32_t rj;           // last row, last max
    int32_t cbest; uint32_t ci;           // last column, first max (valid iff has_col)
    int32_t corner;                       // M[n][m] (valid iff has_col)
    uint32_t has_col;
    uint32_t pad_[7];
};
static_assert(sizeof(WaveCand) == 64, "WaveCand layout");
// Which pair a CTA works on in round r of a launch, as which member of the pair's CTA group.  The host
// sizes the groups in proportion to the pairs' cell counts (a launch holds only as many pairs as their
// traces fit in memory -- about as many as the machine has SMs / 4 -- so equal groups would leave the SMs
// of the small pairs idle until the largest pair is done) and hands every CTA its list.
struct WaveAssign { uint32_t slot; uint16_t rank; uint16_t Q; };   // Q == 0: idle in this round
struct WaveArgs {
    FillArgs f;
    unsigned long long* progress;   // [slot][prog_stride], zeroed before launch
    WaveCand* cand;                 // [slot][cand_stride]
    uint32_t* done;                 // [slot] workers that have finished the pair (zeroed before launch)
    uint32_t* next_band;            // [slot] next unclaimed band of the pair (zeroed before launch)
    const WaveAssign* assign;       // [n_rounds][gridDim.x]
    uint32_t n_rounds;
    uint32_t prog_stride;           // >= max(Q) * K2_WARPS + 1
    uint32_t cand_stride;           // >= max(Q) * K2_WARPS
    // ---- bounded-memory traceback (k2_wave<.., CKPT = true> only) ----
    const CkptSlot* cks;            // [slot]: the row block this launch fills, where the slot's checkpoints live
    uint32_t ckpt_write;            // pass 1: store the checkpoints; pass 2: 0
    uint32_t write_end;             // pass 1: publish the end cell; pass 2: 0
    int2* ckpt;                     // per slot [row block][ck_stride]: (M + a, X) of row (block + 1) * every, by 0-based column
    const WalkState* wstate;        // pass 2: skip pairs whose walk has already ended
};
constexpr int K2_MAX_Q = 13;        // CTAs per pair (196 bands of a 100 kbp pair / 16 warps)
constexpr int K2_WARPS = 16;        // warps per CTA

// ---- K2f (k2f_fine.cuh): one column per lane, for launches of very few long pairs ----
constexpr uint32_t FINE_RING = 64;        // rows of a boundary column in flight between two warps of a CTA (shared memory)
constexpr uint32_t FINE_RING_G = 256;     // ... between the last warp of a CTA and the first of the next (global memory)
struct FineArgs {
    FillArgs f;
    WaveCand* cand;          // [slot][cand_stride]: one record per warp of the pair
    uint32_t* done;          // [slot] warps that have finished the pair (zeroed before launch)
    uint32_t cand_stride;
    unsigned long long* ring_g;   // [CTAs][FINE_RING_G] boundary column of the CTA's last warp: (M + a, Y, generation tag) per row
    uint32_t* cons_g;        // [CTAs] rows the next CTA has consumed
};

// ---- K3 (k3_walk.cuh) ----
struct WalkArgs {
    const PairDesc* desc;
    const EndCell* end;
    uint32_t n_slots;
    const uint8_t* residues;
    const uint32_t* trace;
    int32_t mode;
    int32_t L, C;           // geometry K1 used for this launch
    int32_t H;              // pairs per lane group: 1 (K1 / K2) or 2 (K1h, packed 16 x 2)
    int32_t tg_shift;       // K1h trace tiling
    int32_t CW;             // 0: step-major trace words (bg_common.cuh); > 0: K1h row blocks, CW words per lane and block (k1h_fill.cuh)
    uint8_t* pad;           // padded output slots
    int32_t* score;         // [pair]
    uint8_t* walk_flags;    // [pair]
    uint64_t* lens2;        // [2*pairs + 1]: lens2[2p] = lens2[2p+1] = aligned length
    // bounded-memory traceback (k3_walk_diag only): the trace holds DP rows row0 + 1 .. of the pair; the walk is
    // resumed from / suspended into wstate[slot] when it reaches row row0 > 0
    const CkptSlot* cks = nullptr;
    WalkState* wstate = nullptr;
    uint32_t last_launch = 0;   // row block 0: every walk that is still open ends here
};
struct GatherArgs {
    const PairDesc* desc;
    uint32_t n_slots;
    const uint8_t* pad;
    const uint64_t* off;   // exclusive scan of lens2
    uint8_t* arena;
    const uint8_t* residues;
};
struct PackOpsArgs {
    const PairDesc* desc;
    uint32_t n_slots;
    const uint8_t* pad;      // op slots written by the walkers
    const uint64_t* lens2;   // [2 * pairs + 1]: aligned length of pair p at [2p]
    const ulonglong2* woff;  // [pairs + 1]: exclusive scan of {ceil(len / 16), len}
    uint32_t* len;           // [pairs]
    uint32_t* first;         // [2 * pairs]: (first_a, first_b) -- the strings start at seq1[first_a], seq2[first_b]
    uint32_t* ops;           // dense op words
};

// ---- K4 (k4_edit.cuh) ----
struct EditArgs {
    const PairDesc* desc;
    uint32_t n_slots;
    const uint8_t* residues;
    int32_t* bnd;        // band-boundary column scratch (multi-band pairs only), int32 per row
    uint64_t* out;       // [pair]
};
// Compact launch slot of the host pipeline (16 bytes instead of PairDesc's 64: the pipeline is bound by the H2D
// copy): seq2 follows seq1 in the arena, so b_off = a_off + n; m <= 320 in the bit-parallel classes.
struct __align__(16) MyersSlot {
    uint32_t a_off_lo; uint32_t pair_id;   // pair_id 0xFFFFFFFF = empty slot
    uint32_t n; uint16_t m; uint16_t a_off_hi;
};
static_assert(sizeof(MyersSlot) == 16, "MyersSlot layout");
struct MyersArgs {
    const PairDesc* desc;
    const MyersSlot* cdesc;  // != nullptr: compact slots instead of desc
    uint32_t n_slots;
    const uint32_t* cls_count;   // != nullptr (device-side plan, k0_eplan.cuh): [3] slots per class; this launch handles
    uint32_t cls;                //   class `cls`, i.e. slots [sum(cls_count[0 .. cls)), + cls_count[cls]) of cdesc
    const uint8_t* residues;
    const uint8_t* lut;      // [256] byte -> code 0..3, 0xFF = not in the 4-symbol alphabet (device)
    uint64_t* out;           // [pair]
    uint32_t* err_flag;      // bit 1: a byte outside the alphabet was met (the caller then reruns with K4)
};

// K0e (k0_eplan.cuh): MyersSlot array of a pipeline chunk built on the device
struct EditPlanArgs {
    const uint64_t* off;     // [2n + 1] sequence offsets of the chunk (device copy of the caller's slice)
    uint64_t base;           // off[0]: the chunk's residues start at device offset 0
    uint32_t n_pairs;
    uint32_t n_shift;        // len1 >> n_shift < 64 for (nearly) every pair: the len1 ranges of the counting sort
    uint32_t* hist;          // [193] zeroed by the launcher: pairs per bucket (class x len1 range; [192]: pairs that do not fit K4b)
    uint32_t* cursor;        // [193] zeroed by the launcher
    MyersSlot* slots;        // [n] out
    uint32_t* cls_count;     // [4] out: slots per class (W = 4, 8, 10); [3]: pairs that do not fit
    unsigned long long* cells;   // out (zeroed by the launcher): sum of len1 * len2 over the pairs that fit
    uint32_t* err_flag;      // bit 2: an offset pair that is not monotone
};

// ---- K5 (k5_distance.cuh) ----
struct HammingArgs {
    const uint8_t* residues;
    const uint64_t* seq_off;   // [2 * n_pairs + 1], relative to residues
    uint64_t n_pairs;
    uint64_t* out;             // [n_pairs]
    uint32_t* err_flag;        // bit 1: a pair with len1 != len2 (the reference returns Err(InvalidInputSize))
};
// One warp per pair; pairs longer than HAM_SPLIT bytes are cut into pieces handled by different warps of a grid-
// stride loop and accumulated with one atomic per piece (out zeroed before launch).
constexpr uint64_t HAM_SPLIT = 1u << 16;
struct PDistArgs {
    const uint8_t* residues;
    const uint64_t* seq_off;   // [rows + 1]
    uint64_t rows;
    float columns;             // (columns as f32) = len of row 0 (stat.rs:139,147)
    float* out;                // rows x rows, row-major
};

}  // namespace bg
