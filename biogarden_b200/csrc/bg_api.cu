// bg_api.cu -- C ABI of libbgalign.so (include/bgalign.h): contexts, batch planning, launches.
//
// Host side is deliberately thin: it cuts a batch into length classes (one kernel shape per
// class), lays out the trace / output slots, launches K1 (fill) -> K3 (walk) -> scan -> gather
// on the context's stream and moves bytes.  All arithmetic is in the kernels; there is no CPU
// path (bg_create fails without a device).
#include "../../include/bgalign.h"
#include "../../include/bg_score_tables.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "bg_common.cuh"
#include "k1_fill.cuh"
#include "k3_walk.cuh"
#include "k4_edit.cuh"

using namespace bg;

// ----------------------------------------------------------------------------- utilities
#define CU_TRY(ctx, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));                \
            return e__ == cudaErrorMemoryAllocation ? BG_ENOMEM : BG_ECUDA;                       \
        }                                                                                         \
    } while (0)

namespace {

struct Shape { int L, C; };

// Kernel shapes compiled in: (lanes per pair, columns per lane).  A band is L*C columns.
#define BG_SHAPES(X) X(32, 2) X(32, 4) X(32, 5) X(32, 8) X(32, 12) X(32, 16) X(16, 10) X(8, 19)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want; else { p = nullptr; (void)cudaGetLastError(); }
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Chunk { uint32_t slot_begin, slot_end; uint64_t trace_words; };
struct LaunchClass { Shape sh; std::vector<Chunk> chunks; };

struct Plan {
    std::vector<PairDesc> desc;      // all slots, class after class
    std::vector<LaunchClass> classes;
    uint64_t max_trace_words = 0, bnd_elems = 0, pad_bytes = 0, cells = 0, total_trace_words = 0;
    uint32_t max_n = 0, max_m = 0;
    bool built = false;
};

struct PhaseEv { cudaEvent_t a, b; int phase; };   // phase: 0 encode, 1 fill, 2 walk, 3 compact

struct Device {
    int ordinal = 0;
    cudaStream_t stream = nullptr;
    DevBuf trace, end, bnd, pad, table, codes, err, cubtmp;
    std::vector<PhaseEv> evs;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    size_t total_mem = 0;
    cudaEvent_t get_event() {
        if (ev_used == ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); ev_pool.push_back(e); }
        return ev_pool[ev_used++];
    }
};

}  // namespace

struct bg_ctx {
    std::vector<Device> devs;
    std::string last_error;
    bg_timing timing{};
    uint64_t trace_budget_words = 0;
    int force_L = 0, force_C = 0;
    void set_error(const std::string& s) { last_error = s; }
};

struct bg_dbatch {
    bg_ctx* ctx; int dev_index;
    uint64_t n_pairs = 0, n_residues = 0;
    DevBuf residues, desc_align, desc_edit;
    std::vector<uint64_t> seq_off;   // host copy, rebased to 0
    Plan plan_align, plan_edit;
};

struct bg_dresult {
    bg_ctx* ctx; int dev_index;
    uint64_t n_pairs = 0;
    int kind = 0;             // 0 align, 1 edit distance
    int mode = 0; bool score_only = false;
    DevBuf score, flags, lens2, off, arena, out64;
    std::vector<uint32_t> n, m;   // for status rules
};

namespace {

struct HostResultOwner { std::vector<void*> pinned; };

int pinned_alloc(void** p, size_t bytes) {
    if (bytes == 0) bytes = 1;
    cudaError_t e = cudaHostAlloc(p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { (void)cudaGetLastError(); *p = nullptr; return BG_ENOMEM; }
    return BG_OK;
}

// ------------------------------------------------------------------------------ planning
Shape pick_shape(const bg_ctx* ctx, uint32_t m) {
    if (ctx->force_L) return Shape{ctx->force_L, ctx->force_C};
    if (m <= 64) return Shape{32, 2};
    if (m <= 128) return Shape{32, 4};
    if (m <= 160) return Shape{32, 5};
    if (m <= 256) return Shape{32, 8};
    if (m <= 384) return Shape{32, 12};
    return Shape{32, 16};
}

int shape_index(Shape s) {
    int i = 0;
#define X(L_, C_) if (s.L == L_ && s.C == C_) return i; ++i;
    BG_SHAPES(X)
#undef X
    return -1;
}

// Builds launch classes.  with_trace: trace blocks are laid out and chunked by the budget.
int build_plan(bg_ctx* ctx, const std::vector<uint64_t>& off, uint64_t n_pairs, bool with_trace, Plan& P) {
    P = Plan();
    if (n_pairs >= 0xFFFFFFF0ull) { ctx->set_error("too many pairs in one device batch"); return BG_EINVAL_ARG; }
    constexpr int NS = 16;
    std::vector<uint32_t> per_class[NS];
    Shape shapes[NS];
    int nshape = 0;
#define X(L_, C_) shapes[nshape++] = Shape{L_, C_};
    BG_SHAPES(X)
#undef X
    for (uint64_t p = 0; p < n_pairs; ++p) {
        const uint64_t n = off[2 * p + 1] - off[2 * p], m = off[2 * p + 2] - off[2 * p + 1];
        if (n > 0x7FFFFFF0ull || m > 0x7FFFFFF0ull) { ctx->set_error("sequence longer than 2^31"); return BG_EUNSUPPORTED; }
        const int si = shape_index(pick_shape(ctx, (uint32_t)m));
        if (si < 0) { ctx->set_error("forced kernel shape is not compiled in"); return BG_EINVAL_ARG; }
        per_class[si].push_back((uint32_t)p);
        P.cells += n * m;
        P.max_n = std::max<uint32_t>(P.max_n, (uint32_t)n);
        P.max_m = std::max<uint32_t>(P.max_m, (uint32_t)m);
    }
    uint64_t pad_off = 0, bnd_off = 0;
    for (int si = 0; si < nshape; ++si) {
        auto& ids = per_class[si];
        if (ids.empty()) continue;
        const Shape sh = shapes[si];
        const uint32_t G = 32 / sh.L, K = words_per_lane_step(sh.C), band_cols = sh.L * sh.C;
        // longest first (by rows, then columns) so that the lane groups of a warp and the warps of
        // a wave carry similar work; skipped when the class is uniform.
        bool uniform = true;
        {
            const uint64_t n0 = off[2 * (uint64_t)ids[0] + 1] - off[2 * (uint64_t)ids[0]];
            for (uint32_t id : ids) if (off[2 * (uint64_t)id + 1] - off[2 * (uint64_t)id] != n0) { uniform = false; break; }
        }
        if (!uniform) {
            std::stable_sort(ids.begin(), ids.end(), [&](uint32_t x, uint32_t y) {
                const uint64_t nx = off[2 * (uint64_t)x + 1] - off[2 * (uint64_t)x], ny = off[2 * (uint64_t)y + 1] - off[2 * (uint64_t)y];
                return nx > ny;
            });
        }
        LaunchClass lc; lc.sh = sh;
        Chunk ch; ch.slot_begin = (uint32_t)P.desc.size(); ch.trace_words = 0;
        const size_t nwarps = (ids.size() + G - 1) / G;
        for (size_t w = 0; w < nwarps; ++w) {
            uint32_t maxn = 0, maxb = 0;
            for (uint32_t gidx = 0; gidx < G; ++gidx) {
                const size_t k = w * G + gidx;
                if (k >= ids.size()) break;
                const uint64_t id = ids[k];
                const uint32_t n = (uint32_t)(off[2 * id + 1] - off[2 * id]), m = (uint32_t)(off[2 * id + 2] - off[2 * id + 1]);
                maxn = std::max(maxn, n);
                maxb = std::max(maxb, (m + band_cols - 1) / band_cols);
            }
            const uint32_t steps = maxn + sh.L - 1;
            const uint64_t warp_words = with_trace ? (uint64_t)maxb * steps * K * 32ull : 0;
            if (with_trace && ch.trace_words > 0 && ch.trace_words + warp_words > ctx->trace_budget_words) {
                ch.slot_end = (uint32_t)P.desc.size();
                lc.chunks.push_back(ch);
                P.max_trace_words = std::max(P.max_trace_words, ch.trace_words);
                ch.slot_begin = ch.slot_end; ch.trace_words = 0;
            }
            for (uint32_t gidx = 0; gidx < G; ++gidx) {
                const size_t k = w * G + gidx;
                PairDesc d; memset(&d, 0, sizeof d);
                d.pair_id = 0xFFFFFFFFu; d.steps = steps; d.trace_off = ch.trace_words;
                if (k < ids.size()) {
                    const uint64_t id = ids[k];
                    d.a_off = off[2 * id]; d.b_off = off[2 * id + 1];
                    d.n = (uint32_t)(off[2 * id + 1] - off[2 * id]); d.m = (uint32_t)(off[2 * id + 2] - off[2 * id + 1]);
                    d.nbands = (d.m + band_cols - 1) / band_cols;
                    d.pair_id = (uint32_t)id;
                    d.pad_off = pad_off; pad_off += 2ull * ((uint64_t)d.n + d.m);
                    if (d.nbands > 1) { d.bnd_off = bnd_off; bnd_off += d.n; }
                }
                P.desc.push_back(d);
            }
            ch.trace_words += warp_words;
            P.total_trace_words += warp_words;
        }
        ch.slot_end = (uint32_t)P.desc.size();
        lc.chunks.push_back(ch);
        P.max_trace_words = std::max(P.max_trace_words, ch.trace_words);
        P.classes.push_back(lc);
    }
    P.pad_bytes = pad_off; P.bnd_elems = bnd_off;
    P.built = true;
    return BG_OK;
}

// ------------------------------------------------------------------------------ launches
template <int L, int C>
void launch_k1(bool local, bool prof4, dim3 grid, size_t smem, cudaStream_t st, const FillArgs& a) {
    if (local) {
        if (prof4) k1_fill<L, C, true, true><<<grid, 128, smem, st>>>(a);
        else k1_fill<L, C, true, false><<<grid, 128, smem, st>>>(a);
    } else {
        if (prof4) k1_fill<L, C, false, true><<<grid, 128, smem, st>>>(a);
        else k1_fill<L, C, false, false><<<grid, 128, smem, st>>>(a);
    }
}
void dispatch_k1(Shape sh, bool local, bool prof4, dim3 grid, size_t smem, cudaStream_t st, const FillArgs& a) {
#define X(L_, C_) if (sh.L == L_ && sh.C == C_) { launch_k1<L_, C_>(local, prof4, grid, smem, st, a); return; }
    BG_SHAPES(X)
#undef X
}
void dispatch_k4(Shape sh, dim3 grid, cudaStream_t st, const EditArgs& a) {
#define X(L_, C_) if (sh.L == L_ && sh.C == C_) { k4_edit<L_, C_><<<grid, 128, 0, st>>>(a); return; }
    BG_SHAPES(X)
#undef X
}

struct Phase {
    Device& dv; int phase; cudaEvent_t a;
    Phase(Device& d, int ph) : dv(d), phase(ph) { a = dv.get_event(); cudaEventRecord(a, dv.stream); }
    ~Phase() { cudaEvent_t b = dv.get_event(); cudaEventRecord(b, dv.stream); dv.evs.push_back(PhaseEv{a, b, phase}); }
};

int check_batch(bg_ctx* ctx, const bg_batch* in) {
    if (!in || (in->n_pairs && (!in->seq_off || (!in->residues && in->seq_off[2 * in->n_pairs] != 0)))) {
        ctx->set_error("null batch pointers"); return BG_EINVAL_ARG;
    }
    for (uint64_t s = 0; s < 2 * in->n_pairs; ++s)
        if (in->seq_off[s + 1] < in->seq_off[s]) { ctx->set_error("seq_off not monotone"); return BG_EINVAL_ARG; }
    return BG_OK;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

int bg_version(void) { return BG_API_VERSION; }

const char* bg_strerror(int err) {
    switch (err) {
        case BG_OK: return "ok";
        case BG_EINVAL_RANGE: return "gap penalties outside the supported range (BioError::InvalidArgumentRange)";
        case BG_EINVAL_SIZE: return "inputs have invalid size (BioError::InvalidInputSize)";
        case BG_ECUDA: return "CUDA error";
        case BG_ENOMEM: return "out of memory";
        case BG_EINVAL_ARG: return "invalid argument";
        case BG_EINVAL_RESIDUE: return "residue without an entry in the score table";
        case BG_ENODEVICE: return "no usable CUDA device";
        case BG_EUNSUPPORTED: return "parameters outside the engine's supported range";
        default: return "unknown error";
    }
}

const char* bg_last_error(const bg_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int bg_create(const int* devices, int n_dev, bg_ctx** out) {
    if (!out) return BG_EINVAL_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { (void)cudaGetLastError(); return BG_ENODEVICE; }
    bg_ctx* ctx = new bg_ctx();
    std::vector<int> ords;
    if (!devices || n_dev <= 0) { int cur = 0; cudaGetDevice(&cur); ords.push_back(cur); }
    else ords.assign(devices, devices + n_dev);
    for (int o : ords) {
        if (o < 0 || o >= count) { delete ctx; return BG_ENODEVICE; }
        Device dv; dv.ordinal = o;
        if (cudaSetDevice(o) != cudaSuccess) { delete ctx; return BG_ENODEVICE; }
        if (cudaStreamCreateWithFlags(&dv.stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return BG_ECUDA; }
        size_t fr = 0, tot = 0; cudaMemGetInfo(&fr, &tot); dv.total_mem = tot;
        ctx->devs.push_back(dv);
    }
    uint64_t budget_mb = 8192;
    if (const char* e = getenv("BG_TRACE_BUDGET_MB")) budget_mb = strtoull(e, nullptr, 10);
    ctx->trace_budget_words = budget_mb * (1024ull * 1024ull / 4ull);
    if (const char* e = getenv("BG_FORCE_SHAPE")) {   // "L,C" -- experiments / tests
        int l = 0, c = 0;
        if (sscanf(e, "%d,%d", &l, &c) == 2) { ctx->force_L = l; ctx->force_C = c; }
    }
    *out = ctx;
    return BG_OK;
}

void bg_destroy(bg_ctx* ctx) {
    if (!ctx) return;
    for (auto& dv : ctx->devs) {
        cudaSetDevice(dv.ordinal);
        cudaStreamSynchronize(dv.stream);
        for (DevBuf* b : {&dv.trace, &dv.end, &dv.bnd, &dv.pad, &dv.table, &dv.codes, &dv.err, &dv.cubtmp}) b->release();
        for (auto e : dv.ev_pool) cudaEventDestroy(e);
        cudaStreamDestroy(dv.stream);
    }
    delete ctx;
}

void* bg_stream(bg_ctx* ctx, int dev_index) {
    if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size()) return nullptr;
    return (void*)ctx->devs[dev_index].stream;
}
int bg_device_ordinal(bg_ctx* ctx, int dev_index) {
    if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size()) return -1;
    return ctx->devs[dev_index].ordinal;
}
int bg_set_shape(bg_ctx* ctx, int L, int C) {   // 0,0 = automatic
    if (!ctx) return BG_EINVAL_ARG;
    if (L && shape_index(Shape{L, C}) < 0) return BG_EINVAL_ARG;
    ctx->force_L = L; ctx->force_C = C;
    return BG_OK;
}
int bg_set_trace_budget(bg_ctx* ctx, uint64_t bytes) {
    if (!ctx || bytes < 4096) return BG_EINVAL_ARG;
    ctx->trace_budget_words = bytes / 4;
    return BG_OK;
}

int bg_sync(bg_ctx* ctx) {
    if (!ctx) return BG_EINVAL_ARG;
    for (auto& dv : ctx->devs) {
        cudaSetDevice(dv.ordinal);
        CU_TRY(ctx, cudaStreamSynchronize(dv.stream));
    }
    return BG_OK;
}

int bg_last_timing(const bg_ctx* cctx, bg_timing* out) {
    bg_ctx* ctx = const_cast<bg_ctx*>(cctx);
    if (!ctx || !out) return BG_EINVAL_ARG;
    bg_timing t = ctx->timing;
    t.encode_ms = t.fill_ms = t.walk_ms = t.compact_ms = t.total_ms = 0;
    for (auto& dv : ctx->devs) {
        cudaSetDevice(dv.ordinal);
        CU_TRY(ctx, cudaStreamSynchronize(dv.stream));
        double ph[4] = {0, 0, 0, 0};
        for (auto& ev : dv.evs) {
            float ms = 0; cudaEventElapsedTime(&ms, ev.a, ev.b);
            ph[ev.phase] += ms;
        }
        double tot = 0;
        if (!dv.evs.empty()) { float ms = 0; cudaEventElapsedTime(&ms, dv.evs.front().a, dv.evs.back().b); tot = ms; }
        // devices run concurrently: report the slowest
        t.encode_ms = std::max(t.encode_ms, ph[0]); t.fill_ms = std::max(t.fill_ms, ph[1]);
        t.walk_ms = std::max(t.walk_ms, ph[2]); t.compact_ms = std::max(t.compact_ms, ph[3]);
        t.total_ms = std::max(t.total_ms, tot);
    }
    *out = t;
    return BG_OK;
}

// ------------------------------------------------------------------------- device batches
int bg_batch_upload(bg_ctx* ctx, int dev_index, const bg_batch* in, bg_dbatch** out) {
    if (!ctx || !out || dev_index < 0 || dev_index >= (int)ctx->devs.size()) return BG_EINVAL_ARG;
    *out = nullptr;
    int rc = check_batch(ctx, in);
    if (rc) return rc;
    Device& dv = ctx->devs[dev_index];
    CU_TRY(ctx, cudaSetDevice(dv.ordinal));
    bg_dbatch* B = new bg_dbatch();
    B->ctx = ctx; B->dev_index = dev_index; B->n_pairs = in->n_pairs;
    const uint64_t base = in->n_pairs ? in->seq_off[0] : 0;
    B->seq_off.resize(2 * in->n_pairs + 1);
    for (uint64_t s = 0; s <= 2 * in->n_pairs; ++s) B->seq_off[s] = in->n_pairs ? in->seq_off[s] - base : 0;
    B->n_residues = B->seq_off.back();
    cudaError_t e = B->residues.ensure(B->n_residues + 16);
    if (e != cudaSuccess) { delete B; ctx->set_error("device allocation for residues failed"); return BG_ENOMEM; }
    if (B->n_residues) {
        e = cudaMemcpyAsync(B->residues.p, in->residues + base, B->n_residues, cudaMemcpyHostToDevice, dv.stream);
        if (e != cudaSuccess) { B->residues.release(); delete B; ctx->set_error(cudaGetErrorString(e)); return BG_ECUDA; }
    }
    ctx->timing.h2d_bytes = B->n_residues;
    *out = B;
    return BG_OK;
}

void bg_dbatch_free(bg_dbatch* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->devs[b->dev_index].ordinal);
    cudaStreamSynchronize(b->ctx->devs[b->dev_index].stream);
    b->residues.release(); b->desc_align.release(); b->desc_edit.release();
    delete b;
}

void bg_dresult_free(bg_dresult* r) {
    if (!r) return;
    cudaSetDevice(r->ctx->devs[r->dev_index].ordinal);
    cudaStreamSynchronize(r->ctx->devs[r->dev_index].stream);
    for (DevBuf* b : {&r->score, &r->flags, &r->lens2, &r->off, &r->arena, &r->out64}) b->release();
    delete r;
}

static int ensure_plan(bg_ctx* ctx, bg_dbatch* B, bool edit) {
    Device& dv = ctx->devs[B->dev_index];
    Plan& P = edit ? B->plan_edit : B->plan_align;
    DevBuf& D = edit ? B->desc_edit : B->desc_align;
    if (P.built) return BG_OK;
    int rc = build_plan(ctx, B->seq_off, B->n_pairs, !edit, P);
    if (rc) return rc;
    if (!P.desc.empty()) {
        if (D.ensure(P.desc.size() * sizeof(PairDesc)) != cudaSuccess) { ctx->set_error("device allocation for descriptors failed"); return BG_ENOMEM; }
        CU_TRY(ctx, cudaMemcpyAsync(D.p, P.desc.data(), P.desc.size() * sizeof(PairDesc), cudaMemcpyHostToDevice, dv.stream));
        CU_TRY(ctx, cudaStreamSynchronize(dv.stream));   // P.desc is pageable host memory
        ctx->timing.h2d_bytes += P.desc.size() * sizeof(PairDesc);
    }
    return BG_OK;
}

// Builds the launch plan now (otherwise it is built by the first align / edit call).
int bg_batch_prepare(bg_ctx* ctx, bg_dbatch* b, int for_edit) {
    if (!ctx || !b) return BG_EINVAL_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->devs[b->dev_index].ordinal));
    return ensure_plan(ctx, b, for_edit != 0);
}

int bg_align_device(bg_ctx* ctx, const bg_dbatch* cin, const bg_params* p, bg_dresult** out) {
    if (!ctx || !cin || !p || !out) return BG_EINVAL_ARG;
    *out = nullptr;
    bg_dbatch* B = const_cast<bg_dbatch*>(cin);
    const int mode = p->mode;
    if (mode < BG_GLOBAL || mode > BG_OVERLAP) { ctx->set_error("unknown mode"); return BG_EINVAL_ARG; }
    // aligner.rs:87-89,153-155,219-221: sign check in global / local / fitting only
    if ((mode == BG_GLOBAL || mode == BG_LOCAL || mode == BG_FITTING) && (p->gap_open > 0 || p->gap_extend > 0)) return BG_EINVAL_RANGE;
    if (!p->table || !p->row_code || !p->col_code || p->n_rows <= 0 || p->n_cols <= 0 || p->n_rows > 255 || p->n_cols > 255) {
        ctx->set_error("score table missing or malformed"); return BG_EINVAL_ARG;
    }
    const uint64_t N = B->n_pairs;
    if (mode == BG_FITTING)   // aligner.rs:223-225
        for (uint64_t q = 0; q < N; ++q)
            if (B->seq_off[2 * q + 1] - B->seq_off[2 * q] < B->seq_off[2 * q + 2] - B->seq_off[2 * q + 1]) return BG_EINVAL_SIZE;

    Device& dv = ctx->devs[B->dev_index];
    CU_TRY(ctx, cudaSetDevice(dv.ordinal));
    int rc = ensure_plan(ctx, B, false);
    if (rc) return rc;
    Plan& P = B->plan_align;

    // 32-bit safety of the recurrence (bg_common.cuh NEG_INF)
    int64_t maxabs = std::max<int64_t>(llabs((long long)p->gap_open), llabs((long long)p->gap_extend));
    bool fits8 = true;
    for (int i = 0; i < p->n_rows * p->n_cols; ++i) {
        maxabs = std::max<int64_t>(maxabs, llabs((long long)p->table[i]));
        if (p->table[i] < -128 || p->table[i] > 127) fits8 = false;
    }
    if (maxabs > (1 << 20) || (int64_t)((uint64_t)P.max_n + P.max_m + 2) * maxabs >= (1ll << 28)) {
        ctx->set_error("scores * length exceed the 32-bit-safe range"); return BG_EUNSUPPORTED;
    }
    const size_t smem = 512 + (size_t)p->n_rows * (p->n_cols + 1) * 4;
    if (smem > 48 * 1024) { ctx->set_error("score table too large for shared memory"); return BG_EUNSUPPORTED; }
    const bool prof4 = fits8 && p->n_rows <= 4;
    const bool local = (mode == BG_LOCAL);
    const bool score_only = (p->flags & BG_F_SCORE_ONLY) != 0;

    bg_dresult* R = new bg_dresult();
    R->ctx = ctx; R->dev_index = B->dev_index; R->n_pairs = N; R->kind = 0; R->mode = mode; R->score_only = score_only;
    R->n.resize(N); R->m.resize(N);
    for (uint64_t q = 0; q < N; ++q) {
        R->n[q] = (uint32_t)(B->seq_off[2 * q + 1] - B->seq_off[2 * q]);
        R->m[q] = (uint32_t)(B->seq_off[2 * q + 2] - B->seq_off[2 * q + 1]);
    }
    auto fail = [&](int code, const char* what) { ctx->set_error(what); bg_dresult_free(R); return code; };

    const size_t n_slots = P.desc.size();
    bool ok = true;
    ok &= dv.table.ensure((size_t)p->n_rows * p->n_cols * 4) == cudaSuccess;
    ok &= dv.codes.ensure(512) == cudaSuccess;
    ok &= dv.err.ensure(4) == cudaSuccess;
    ok &= dv.end.ensure(std::max<size_t>(1, n_slots) * sizeof(EndCell)) == cudaSuccess;
    ok &= dv.bnd.ensure(std::max<uint64_t>(1, P.bnd_elems) * sizeof(int2)) == cudaSuccess;
    ok &= R->score.ensure(std::max<uint64_t>(1, N) * 4) == cudaSuccess;
    ok &= R->flags.ensure(std::max<uint64_t>(1, N)) == cudaSuccess;
    if (!score_only) {
        ok &= dv.trace.ensure(std::max<uint64_t>(1, P.max_trace_words) * 4) == cudaSuccess;
        ok &= dv.pad.ensure(std::max<uint64_t>(1, P.pad_bytes)) == cudaSuccess;
        ok &= R->lens2.ensure((2 * N + 1) * 8) == cudaSuccess;
        ok &= R->off.ensure((2 * N + 1) * 8) == cudaSuccess;
        ok &= R->arena.ensure(std::max<uint64_t>(1, P.pad_bytes)) == cudaSuccess;
    }
    if (!ok) return fail(BG_ENOMEM, "device allocation failed (trace / output buffers)");

    cudaStream_t st = dv.stream;
    dv.evs.clear(); dv.ev_used = 0;
    ctx->timing.cells = P.cells; ctx->timing.launches = 0;
    ctx->timing.trace_bytes = score_only ? 0 : P.total_trace_words * 4;

    // small parameter uploads (pageable -> staged synchronously by the runtime)
    uint8_t codes[512];
    memcpy(codes, p->row_code, 256); memcpy(codes + 256, p->col_code, 256);
    for (int i = 0; i < 256; ++i) {
        if (codes[i] != 0xFF && codes[i] >= p->n_rows) return fail(BG_EINVAL_ARG, "row_code entry out of range");
        if (codes[256 + i] != 0xFF && codes[256 + i] >= p->n_cols) return fail(BG_EINVAL_ARG, "col_code entry out of range");
    }
    CU_TRY(ctx, cudaMemcpyAsync(dv.table.p, p->table, (size_t)p->n_rows * p->n_cols * 4, cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaMemcpyAsync(dv.codes.p, codes, 512, cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaMemsetAsync(dv.err.p, 0, 4, st));
    if (!score_only) CU_TRY(ctx, cudaMemsetAsync(R->lens2.p, 0, (2 * N + 1) * 8, st));

    FillArgs fa;
    fa.desc = nullptr; fa.n_slots = 0;
    fa.residues = B->residues.as<uint8_t>();
    fa.table = dv.table.as<int32_t>(); fa.n_rows = p->n_rows; fa.n_cols = p->n_cols;
    fa.row_code = dv.codes.as<uint8_t>(); fa.col_code = dv.codes.as<uint8_t>() + 256;
    fa.a = p->gap_open; fa.b = p->gap_extend; fa.mode = mode; fa.want_trace = score_only ? 0 : 1;
    fa.trace = dv.trace.as<uint32_t>(); fa.bnd = dv.bnd.as<int2>(); fa.end = nullptr; fa.err_flag = dv.err.as<uint32_t>();

    for (const LaunchClass& lc : P.classes) {
        const uint32_t G = 32 / lc.sh.L;
        for (const Chunk& ch : lc.chunks) {
            const uint32_t ns = ch.slot_end - ch.slot_begin;
            if (!ns) continue;
            fa.desc = B->desc_align.as<PairDesc>() + ch.slot_begin;
            fa.end = dv.end.as<EndCell>() + ch.slot_begin;
            fa.n_slots = ns;
            const uint32_t nwarps = (ns + G - 1) / G;
            {
                Phase ph(dv, 1);
                dispatch_k1(lc.sh, local, prof4, dim3((nwarps + 3) / 4), smem, st, fa);
                ctx->timing.launches++;
            }
            CU_TRY(ctx, cudaGetLastError());
            {
                Phase ph(dv, 2);
                if (score_only) {
                    k_scores_only<<<(ns + 127) / 128, 128, 0, st>>>(fa.desc, fa.end, ns, R->score.as<int32_t>(), R->flags.as<uint8_t>());
                } else {
                    WalkArgs wa;
                    wa.desc = fa.desc; wa.end = fa.end; wa.n_slots = ns; wa.residues = fa.residues;
                    wa.trace = fa.trace; wa.mode = mode; wa.L = lc.sh.L; wa.C = lc.sh.C;
                    wa.pad = dv.pad.as<uint8_t>(); wa.score = R->score.as<int32_t>(); wa.walk_flags = R->flags.as<uint8_t>();
                    wa.lens2 = R->lens2.as<uint64_t>();
                    k3_walk<<<(ns + 127) / 128, 128, 0, st>>>(wa);
                }
                ctx->timing.launches++;
            }
            CU_TRY(ctx, cudaGetLastError());
        }
    }
    if (!score_only) {
        Phase ph(dv, 3);
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, R->lens2.as<uint64_t>(), R->off.as<uint64_t>(), (int)(2 * N + 1), st);
        if (dv.cubtmp.ensure(tmp + 16) != cudaSuccess) return fail(BG_ENOMEM, "device allocation failed (scan)");
        cub::DeviceScan::ExclusiveSum(dv.cubtmp.p, tmp, R->lens2.as<uint64_t>(), R->off.as<uint64_t>(), (int)(2 * N + 1), st);
        ctx->timing.launches += 2;
        if (n_slots) {
            GatherArgs ga;
            ga.desc = B->desc_align.as<PairDesc>(); ga.n_slots = (uint32_t)n_slots; ga.pad = dv.pad.as<uint8_t>();
            ga.off = R->off.as<uint64_t>(); ga.arena = R->arena.as<uint8_t>();
            k_gather<<<(unsigned)((n_slots + 3) / 4), 128, 0, st>>>(ga);
            ctx->timing.launches++;
        }
        CU_TRY(ctx, cudaGetLastError());
    }
    *out = R;
    return BG_OK;
}

int bg_edit_distance_device(bg_ctx* ctx, const bg_dbatch* cin, bg_dresult** out) {
    if (!ctx || !cin || !out) return BG_EINVAL_ARG;
    *out = nullptr;
    bg_dbatch* B = const_cast<bg_dbatch*>(cin);
    Device& dv = ctx->devs[B->dev_index];
    CU_TRY(ctx, cudaSetDevice(dv.ordinal));
    int rc = ensure_plan(ctx, B, true);
    if (rc) return rc;
    Plan& P = B->plan_edit;
    const uint64_t N = B->n_pairs;
    bg_dresult* R = new bg_dresult();
    R->ctx = ctx; R->dev_index = B->dev_index; R->n_pairs = N; R->kind = 1;
    bool ok = R->out64.ensure(std::max<uint64_t>(1, N) * 8) == cudaSuccess;
    ok &= dv.bnd.ensure(std::max<uint64_t>(1, P.bnd_elems) * sizeof(int2)) == cudaSuccess;
    if (!ok) { ctx->set_error("device allocation failed"); bg_dresult_free(R); return BG_ENOMEM; }
    dv.evs.clear(); dv.ev_used = 0;
    ctx->timing.cells = P.cells; ctx->timing.launches = 0; ctx->timing.trace_bytes = 0;
    EditArgs ea;
    ea.residues = B->residues.as<uint8_t>(); ea.bnd = dv.bnd.as<int32_t>(); ea.out = R->out64.as<uint64_t>();
    for (const LaunchClass& lc : P.classes) {
        const uint32_t G = 32 / lc.sh.L;
        for (const Chunk& ch : lc.chunks) {
            const uint32_t ns = ch.slot_end - ch.slot_begin;
            if (!ns) continue;
            ea.desc = B->desc_edit.as<PairDesc>() + ch.slot_begin; ea.n_slots = ns;
            const uint32_t nwarps = (ns + G - 1) / G;
            Phase ph(dv, 1);
            dispatch_k4(lc.sh, dim3((nwarps + 3) / 4), dv.stream, ea);
            ctx->timing.launches++;
        }
    }
    CU_TRY(ctx, cudaGetLastError());
    *out = R;
    return BG_OK;
}

int bg_ref_status(int mode, uint64_t n, uint64_t m, int32_t score, int walk_flags) {
    // Fresh-aligner buffer dims (aligner.rs:45, 92-94, 594-595)
    uint64_t R, C;
    if (n > 1024 || m > 1024) { R = n + 1; C = m + 1; } else { R = 1024; C = 1024; }
    if (walk_flags & (WALK_UNDERFLOW | WALK_HANG)) return BG_ST_REF_UNDEFINED;
    const bool row_border = (mode == BG_GLOBAL || mode == BG_FITTING);   // writes row0[1..=m]
    const bool col_border = (mode == BG_GLOBAL);                          // writes col0[1..=n]
    if (row_border && (C < 2 || m >= C)) return BG_ST_REF_UNDEFINED;
    if (col_border && (R < 2 || n >= R)) return BG_ST_REF_UNDEFINED;
    if (n >= 1 && m >= 1 && (n >= R || m >= C)) return BG_ST_REF_UNDEFINED;   // fill indexes [n][m]
    if ((mode == BG_OVERLAP || mode == BG_SEMIGLOBAL) && n >= R) return BG_ST_REF_UNDEFINED;   // .row(len1)
    if ((mode == BG_FITTING || mode == BG_SEMIGLOBAL) && m >= C) return BG_ST_REF_UNDEFINED;   // .column(len2)
    // whole-buffer scans see the zeroed cells outside the rectangle (aligner.rs:247,308,369,376)
    if ((mode == BG_SEMIGLOBAL || mode == BG_OVERLAP) && score == 0 && C > m + 1) return BG_ST_REF_UNDEFINED;
    if (mode == BG_FITTING && score < 0 && R > n + 1) return BG_ST_REF_UNDEFINED;
    return BG_ST_OK;
}

int bg_dresult_download(bg_ctx* ctx, bg_dresult* r, bg_result* out) {
    if (!ctx || !r || !out || r->kind != 0) return BG_EINVAL_ARG;
    memset(out, 0, sizeof *out);
    Device& dv = ctx->devs[r->dev_index];
    CU_TRY(ctx, cudaSetDevice(dv.ordinal));
    const uint64_t N = r->n_pairs;
    uint32_t err = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&err, dv.err.p, 4, cudaMemcpyDeviceToHost, dv.stream));
    uint64_t total = 0;
    if (!r->score_only)
        CU_TRY(ctx, cudaMemcpyAsync(&total, r->off.as<uint64_t>() + 2 * N, 8, cudaMemcpyDeviceToHost, dv.stream));
    CU_TRY(ctx, cudaStreamSynchronize(dv.stream));
    if (err & 1u) { ctx->set_error("a residue byte has no row/column in the score table"); return BG_EINVAL_RESIDUE; }

    HostResultOwner* own = new HostResultOwner();
    auto grab = [&](size_t bytes) -> void* { void* q = nullptr; if (pinned_alloc(&q, bytes) == BG_OK) own->pinned.push_back(q); return q; };
    out->n_pairs = N;
    out->score = (int32_t*)grab(N * 4);
    out->status = (uint8_t*)grab(N);
    out->off = (uint64_t*)grab((2 * N + 1) * 8);
    out->arena = (uint8_t*)grab(total);
    out->owner_ = own;
    if (!out->score || !out->status || !out->off || !out->arena) { bg_result_free(out); ctx->set_error("pinned host allocation failed"); return BG_ENOMEM; }
    if (N) {
        CU_TRY(ctx, cudaMemcpyAsync(out->score, r->score.p, N * 4, cudaMemcpyDeviceToHost, dv.stream));
        CU_TRY(ctx, cudaMemcpyAsync(out->status, r->flags.p, N, cudaMemcpyDeviceToHost, dv.stream));
    }
    if (!r->score_only) {
        CU_TRY(ctx, cudaMemcpyAsync(out->off, r->off.p, (2 * N + 1) * 8, cudaMemcpyDeviceToHost, dv.stream));
        if (total) CU_TRY(ctx, cudaMemcpyAsync(out->arena, r->arena.p, total, cudaMemcpyDeviceToHost, dv.stream));
    } else {
        memset(out->off, 0, (2 * N + 1) * 8);
    }
    CU_TRY(ctx, cudaStreamSynchronize(dv.stream));
    ctx->timing.d2h_bytes = N * 5 + (r->score_only ? 0 : (2 * N + 1) * 8 + total);
    for (uint64_t q = 0; q < N; ++q)
        out->status[q] = (uint8_t)bg_ref_status(r->mode, r->n[q], r->m[q], out->score[q], out->status[q]);
    return BG_OK;
}

int bg_dresult_download_u64(bg_ctx* ctx, bg_dresult* r, uint64_t* out) {
    if (!ctx || !r || r->kind != 1 || (!out && r->n_pairs)) return BG_EINVAL_ARG;
    Device& dv = ctx->devs[r->dev_index];
    CU_TRY(ctx, cudaSetDevice(dv.ordinal));
    if (r->n_pairs) CU_TRY(ctx, cudaMemcpyAsync(out, r->out64.p, r->n_pairs * 8, cudaMemcpyDeviceToHost, dv.stream));
    CU_TRY(ctx, cudaStreamSynchronize(dv.stream));
    ctx->timing.d2h_bytes = r->n_pairs * 8;
    return BG_OK;
}

void bg_result_free(bg_result* r) {
    if (!r) return;
    if (r->owner_) {
        HostResultOwner* own = (HostResultOwner*)r->owner_;
        for (void* q : own->pinned) cudaFreeHost(q);
        delete own;
    }
    memset(r, 0, sizeof *r);
}

// ------------------------------------------------------------------ host-buffer entry points
namespace {
// contiguous shards with ~equal cell counts
std::vector<uint64_t> shard_bounds(const bg_batch* in, int nd) {
    std::vector<uint64_t> b(nd + 1, 0);
    const uint64_t N = in->n_pairs;
    if (nd == 1) { b[1] = N; return b; }
    std::vector<double> pre(N + 1, 0.0);
    for (uint64_t q = 0; q < N; ++q) {
        const double n = (double)(in->seq_off[2 * q + 1] - in->seq_off[2 * q]), m = (double)(in->seq_off[2 * q + 2] - in->seq_off[2 * q + 1]);
        pre[q + 1] = pre[q] + n * m + 64.0;
    }
    for (int d = 1; d < nd; ++d) {
        const double target = pre[N] * d / nd;
        b[d] = std::lower_bound(pre.begin(), pre.end(), target) - pre.begin();
        if (b[d] > N) b[d] = N;
        if (b[d] < b[d - 1]) b[d] = b[d - 1];
    }
    b[nd] = N;
    return b;
}
}  // namespace

int bg_align_batch(bg_ctx* ctx, const bg_batch* in, const bg_params* p, bg_result* out) {
    if (!ctx || !in || !p || !out) return BG_EINVAL_ARG;
    memset(out, 0, sizeof *out);
    int rc = check_batch(ctx, in);
    if (rc) return rc;
    const int nd = (int)ctx->devs.size();
    const std::vector<uint64_t> bounds = shard_bounds(in, nd);
    std::vector<bg_result> parts(nd);
    std::vector<int> rcs(nd, BG_OK);
    std::vector<std::string> errs(nd);
    uint64_t h2d = 0, d2h = 0;
    std::mutex mu;
    auto work = [&](int d) {
        bg_batch sub; sub.n_pairs = bounds[d + 1] - bounds[d]; sub.residues = in->residues; sub.seq_off = in->seq_off + 2 * bounds[d];
        memset(&parts[d], 0, sizeof(bg_result));
        bg_dbatch* B = nullptr; bg_dresult* R = nullptr;
        int r = bg_batch_upload(ctx, d, &sub, &B);
        if (!r) r = bg_align_device(ctx, B, p, &R);
        if (!r) r = bg_dresult_download(ctx, R, &parts[d]);
        {
            std::lock_guard<std::mutex> lk(mu);
            if (r) errs[d] = ctx->last_error;
            if (B) h2d += B->n_residues + B->plan_align.desc.size() * sizeof(PairDesc);
            d2h += ctx->timing.d2h_bytes;
        }
        if (R) bg_dresult_free(R);
        if (B) bg_dbatch_free(B);
        rcs[d] = r;
    };
    if (nd == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int d = 0; d < nd; ++d) th.emplace_back(work, d);
        for (auto& t : th) t.join();
    }
    for (int d = 0; d < nd; ++d)
        if (rcs[d]) {
            ctx->set_error(errs[d]);
            for (auto& q : parts) bg_result_free(&q);
            return rcs[d];
        }
    ctx->timing.h2d_bytes = h2d; ctx->timing.d2h_bytes = d2h;
    if (nd == 1) { *out = parts[0]; return BG_OK; }
    // stitch the shards back into input order (pairs were dealt in contiguous ranges)
    const uint64_t N = in->n_pairs;
    uint64_t total = 0;
    for (auto& q : parts) total += q.off ? q.off[2 * q.n_pairs] : 0;
    HostResultOwner* own = new HostResultOwner();
    auto grab = [&](size_t bytes) -> void* { void* q = nullptr; if (pinned_alloc(&q, bytes) == BG_OK) own->pinned.push_back(q); return q; };
    out->n_pairs = N; out->owner_ = own;
    out->score = (int32_t*)grab(N * 4); out->status = (uint8_t*)grab(N);
    out->off = (uint64_t*)grab((2 * N + 1) * 8); out->arena = (uint8_t*)grab(total);
    if (!out->score || !out->status || !out->off || !out->arena) {
        bg_result_free(out); for (auto& q : parts) bg_result_free(&q);
        ctx->set_error("pinned host allocation failed"); return BG_ENOMEM;
    }
    uint64_t base = 0;
    for (int d = 0; d < nd; ++d) {
        bg_result& q = parts[d];
        const uint64_t lo = bounds[d], cnt = q.n_pairs;
        if (cnt) {
            memcpy(out->score + lo, q.score, cnt * 4);
            memcpy(out->status + lo, q.status, cnt);
            for (uint64_t s = 0; s < 2 * cnt; ++s) out->off[2 * lo + s] = base + q.off[s];
            memcpy(out->arena + base, q.arena, q.off[2 * cnt]);
            base += q.off[2 * cnt];
        }
        bg_result_free(&q);
    }
    out->off[2 * N] = base;
    return BG_OK;
}

int bg_edit_distance_batch(bg_ctx* ctx, const bg_batch* in, uint64_t* out) {
    if (!ctx || !in || (!out && in->n_pairs)) return BG_EINVAL_ARG;
    int rc = check_batch(ctx, in);
    if (rc) return rc;
    const int nd = (int)ctx->devs.size();
    const std::vector<uint64_t> bounds = shard_bounds(in, nd);
    std::vector<int> rcs(nd, BG_OK);
    auto work = [&](int d) {
        bg_batch sub; sub.n_pairs = bounds[d + 1] - bounds[d]; sub.residues = in->residues; sub.seq_off = in->seq_off + 2 * bounds[d];
        bg_dbatch* B = nullptr; bg_dresult* R = nullptr;
        int r = bg_batch_upload(ctx, d, &sub, &B);
        if (!r) r = bg_edit_distance_device(ctx, B, &R);
        if (!r) r = bg_dresult_download_u64(ctx, R, out + bounds[d]);
        if (R) bg_dresult_free(R);
        if (B) bg_dbatch_free(B);
        rcs[d] = r;
    };
    if (nd == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int d = 0; d < nd; ++d) th.emplace_back(work, d);
        for (auto& t : th) t.join();
    }
    for (int r : rcs) if (r) return r;
    return BG_OK;
}

const int8_t* bg_score_table26(const char* name) {
    if (!name) return nullptr;
    if (!strcmp(name, "blosum62")) return &BG_TABLE_BLOSUM62[0][0];
    if (!strcmp(name, "pam250")) return &BG_TABLE_PAM250[0][0];
    if (!strcmp(name, "unit")) return &BG_TABLE_UNIT[0][0];
    return nullptr;
}

}  // extern "C"
