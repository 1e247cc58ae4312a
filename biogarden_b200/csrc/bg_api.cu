// bg_api.cu -- C ABI of libbgalign.so (include/bgalign.h): contexts, batch planning, launches.
//
// Host side is deliberately thin: it cuts a batch into length classes (one kernel shape per
// class), lays out the trace / output slots, launches K1 (fill) -> K3 (walk) -> scan -> gather
// and moves bytes.  All arithmetic is in the kernels; there is no CPU path (bg_create fails
// without a device).
//
// Two ways in:
//   * device-resident (bg_batch_upload / bg_align_device / bg_dresult_download): one pass over a
//     batch that already sits in HBM, everything asynchronous on work set 0's stream;
//   * host buffers (bg_align_batch / bg_edit_distance_batch): the batch is cut into chunks that
//     flow through a 3-deep software pipeline per device (plan on the host | H2D | kernels | D2H),
//     each stage of chunk c overlapping other stages of chunks c+1, c+2 on separate streams.
#include "../../include/bgalign.h"
#include "../../include/bg_score_tables.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bg_args.cuh"
#include "launch.h"
#include "pack_util.h"

namespace bg {
// host_util.cpp: residues [first, first + count) of a packed arena -> bytes
void make_unpack_lut2(const uint8_t* alphabet, uint32_t* lut);
void unpack_residues(const uint8_t* packed, uint32_t bits, const uint8_t* alphabet, uint64_t first, uint64_t count, uint8_t* out, const uint32_t* lut2);
// host_expand.cpp: 2-bit alignment ops -> the two aligned strings (AVX-512 VBMI2 or portable code)
void expand_ops(const uint8_t* s1, const uint8_t* s2, const uint32_t* ops, uint64_t len, uint8_t* a_out, uint8_t* b_out);
void stream_copy(uint8_t* dst, const uint8_t* src, size_t n);
}
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

constexpr int MAX_SHAPES = 20;
constexpr int PIPE_DEPTH = 6;
// Pairs wider than this run on K2 (one pair per thread-block cluster, bands of 32 * WAVE_C columns).
constexpr uint32_t WAVE_MIN_COLS = 4096;

// Size-keyed free lists so that steady-state calls never hit cudaMalloc / cudaFree / cudaHostAlloc
// (each of which synchronises the device or pins pages: milliseconds to 100s of milliseconds).
struct BlockCache {
    bool pinned_host = false;
    std::mutex mu;
    std::multimap<size_t, void*> free_;
    void* get(size_t bytes, size_t* got) {
        if (bytes == 0) bytes = 1;
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = free_.lower_bound(bytes);
            if (it != free_.end() && it->first <= bytes * 2 + (1 << 20)) {
                void* p = it->second; *got = it->first; free_.erase(it); return p;
            }
        }
        void* p = nullptr;
        size_t want = bytes + bytes / 16 + 256;
        cudaError_t e = pinned_host ? cudaHostAlloc(&p, want, cudaHostAllocDefault) : cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            trim();   // give cached blocks back and retry with the exact size
            want = bytes;
            e = pinned_host ? cudaHostAlloc(&p, want, cudaHostAllocDefault) : cudaMalloc(&p, want);
            if (e != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        }
        *got = want;
        return p;
    }
    void put(void* p, size_t bytes) {
        if (!p) return;
        std::lock_guard<std::mutex> lk(mu);
        free_.emplace(bytes, p);
    }
    void trim() {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& kv : free_) { if (pinned_host) cudaFreeHost(kv.second); else cudaFree(kv.second); }
        free_.clear();
    }
};

BlockCache& pinned_cache() { static BlockCache c; c.pinned_host = true; return c; }

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    BlockCache* cache = nullptr;
    bool ensure(size_t bytes) {
        if (bytes <= cap && p) return true;
        release();
        p = cache->get(bytes, &cap);
        if (!p) { cap = 0; return false; }
        return true;
    }
    void release() {
        if (p && cache) cache->put(p, cap);
        p = nullptr; cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {   // pinned host block from the global cache
    void* p = nullptr; size_t cap = 0;
    bool ensure(size_t bytes) {
        if (bytes <= cap && p) return true;
        release();
        p = pinned_cache().get(bytes, &cap);
        if (!p) { cap = 0; return false; }
        return true;
    }
    void release() { if (p) pinned_cache().put(p, cap); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Chunk {
    uint32_t slot_begin, slot_end; uint64_t trace_words;
    std::vector<WaveAssign> assign; uint32_t n_rounds = 0, max_Q = 1;   // K2 launches only
    uint32_t wpc = K2_WARPS;              // K2: warps per CTA of this launch (fewer when it has fewer bands than the machine has warps)
    uint32_t fine_bands = 0;              // > 0: K2f launch (k2f_fine.cuh, one column per lane) -- the widest pair's warps
    // bounded-memory traceback (k2_wave.cuh): pairs whose whole traces do not fit the budget; trace_words then
    // covers one row block of every pair
    uint32_t ckpt_nb = 0;                 // > 0: bounded-memory chunk, every pair cut into this many row blocks
    std::vector<CkptSlot> ck_table;       // [ckpt_nb + 1][slots]: pass 1, then the blocks bottom-up
    uint64_t ckpt_elems = 0;
};
struct LaunchClass { Shape sh; std::vector<Chunk> chunks; bool wave = false; bool half = false; int myers_W = 0; bool long_walk = false; };

struct Plan {
    std::vector<LaunchClass> classes;
    size_t n_slots = 0;
    uint64_t max_trace_words = 0, bnd_elems = 0, pad_bytes = 0, cells = 0, total_trace_words = 0;
    uint64_t cells_half = 0, cells_myers = 0;   // of `cells`: in K1h / K4b classes
    uint64_t cells_ckpt = 0;                    // of `cells`: pairs on the bounded-memory path (filled twice)
    uint64_t max_wave_slots = 0;      // largest K2 launch (slots), for the progress / candidate scratch
    uint64_t ckpt_elems = 0;          // bounded-memory traceback: int2 elements of the largest checkpoint array
    uint64_t max_nw = K2_WARPS;   // K2: most workers (warps) any pair of any launch has
    uint32_t max_n = 0, max_m = 0;
    int32_t half_maxabs = 0;          // > 0: short classes were laid out for K1h (packed 16 x 2) with this max |score|
    bool wave_overlap = false;        // K2 launches were cut for TWO trace buffers: the walk of launch c runs next to the fill of launch c + 1
    bool myers = false;               // edit distance: pairs with len2 <= 320 laid out one per thread for K4b
    bool compact = false;             // host pipeline, all classes K4b: the staged descriptors are MyersSlot (16 B), not PairDesc
    bool built = false;
};

struct PhaseEv { cudaEvent_t a, b; int phase; };   // phase: 0 encode, 1 fill, 2 walk, 3 compact

// Everything one in-flight unit of work needs on one device.
struct WorkSet {
    int ordinal = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t walk_stream = nullptr;   // traceback walks of chunk c run here, next to the fill of chunk c+1
    cudaStream_t post_stream = nullptr;   // host-buffer pipeline: walk + scan + gather of this set's chunk go here (next to the next chunk's fill)
    uint64_t split_cap_words = (6ull << 30) / 4;   // all launches of a plan get their own trace region up to this total
    int launch_parity = -1;               // host-buffer pipeline: >= 0 -> this item's first fill goes to the main (0) or the second (1) fill stream, so that
                                          // the fills of consecutive items alternate and the tail of one overlaps the head of the next
    cudaStream_t aux[2] = {nullptr, nullptr};   // edit-distance pipeline: the K4b launches of a chunk's three classes run side by side (created on first use)
    cudaStream_t fill2_stream = nullptr;  // host-buffer pipeline: every other fill launch of a chunk goes here, so that the tail of
                                          // one length class's launch overlaps the head of the next (their traces are disjoint)
    BlockCache* cache = nullptr;
    DevBuf trace2;                        // second trace buffer (chunks alternate)
    DevBuf trace, end, bnd, pad, table, codes, err, cubtmp, progress, cand;   // scratch + parameters
    DevBuf residues, desc, score, flags, lens2, off, arena, out64;     // pipeline mode: chunk in / out
    DevBuf assign;                                                     // K2: CTA assignment table of the launch
    DevBuf ckpt, wstate, ckslots;                                      // K2 bounded-memory traceback: row checkpoints, suspended walks, launch table
    DevBuf run;                                                        // (unused by the alignment pipeline since results travel as ops)
    DevBuf len, first, ops;                                            // pipeline mode: compact results of the chunk (k_pack_ops)
    DevBuf packed;                                                     // packed input: the chunk's packed bytes before k_unpack
    DevBuf samples;                                                    // pipeline mode: sampled offset scans of the chunk
    PinBuf samples_h;
    DevBuf poff, pkeys, pids, psort;                                   // pipeline mode: device-side planner (offsets of the chunk, sort keys / ids / scratch)
    PinBuf stage;                                                     // descriptor staging
    PinBuf scalars;                                                   // [0] total bytes (u64), [1] err flag
    cudaEvent_t ev_scan = nullptr;
    std::vector<PhaseEv> evs;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    cudaEvent_t get_event() {
        if (ev_used == ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); ev_pool.push_back(e); }
        return ev_pool[ev_used++];
    }
    void reset_events() { evs.clear(); ev_used = 0; }
    std::vector<DevBuf*> all_bufs() {
        return {&trace, &trace2, &end, &bnd, &pad, &table, &codes, &err, &cubtmp, &progress, &cand, &residues, &desc, &score, &flags, &lens2, &off, &arena, &out64, &run, &assign, &ckpt, &wstate, &ckslots, &len, &first, &ops, &poff, &pkeys, &pids, &psort, &samples, &packed};
    }
};

struct Device {
    int ordinal = 0;
    size_t total_mem = 0;
    BlockCache* cache = nullptr;   // device blocks of this ordinal
    WorkSet ws[PIPE_DEPTH];
};

// Validated, device-independent view of bg_params.
struct Prepared {
    int mode = 0; int32_t a = 0, b = 0;
    bool local = false, prof4 = false, score_only = false;
    size_t smem = 0;
    int n_rows = 0, n_cols = 0;
    std::vector<int32_t> table;
    uint8_t codes[512];
    int64_t maxabs = 0;
    int32_t half_maxabs = 0;   // > 0: K1h (packed 16 x 2) may be used for short classes
    bool half_prof8 = false;   // ... in its byte-profile form: every s - a - b (and -a - b) fits a signed byte
};

}  // namespace

// A persistent host thread that runs one job at a time (the per-device drivers / finishers of multi-device and
// pipelined calls: spawning std::threads per call cost more than a small batch takes).
struct Worker {
    std::mutex mu; std::condition_variable cv;
    std::function<void()> job; bool has = false, done = true, quit = false;
    std::thread th;
    Worker() : th([this] { loop(); }) {}
    ~Worker() { { std::lock_guard<std::mutex> lk(mu); quit = true; } cv.notify_all(); th.join(); }
    void loop() {
        for (;;) {
            std::function<void()> j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return has || quit; });
                if (!has && quit) return;
                j = std::move(job); has = false;
            }
            j();
            { std::lock_guard<std::mutex> lk(mu); done = true; }
            cv.notify_all();
        }
    }
    void run(std::function<void()> f) {
        { std::lock_guard<std::mutex> lk(mu); job = std::move(f); has = true; done = false; }
        cv.notify_all();
    }
    void wait() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return done; }); }
};

struct bg_ctx {
    std::vector<std::unique_ptr<Worker>> workers;   // [2 d] driver of device d (d >= 1), [2 d + 1] its finisher; created on first use
    Worker& worker(size_t i) {
        if (workers.size() <= i) workers.resize(i + 1);
        if (!workers[i]) workers[i].reset(new Worker());
        return *workers[i];
    }
    Worker& dev_worker(int d) { return worker(2 * (size_t)d); }
    Worker& fin_worker(int d) { return worker(2 * (size_t)d + 1); }
    std::vector<Device> devs;
    std::string last_error;
    std::mutex err_mu;
    bg_timing timing{};
    std::atomic<uint64_t> h2d{0}, d2h{0}, launches{0};
    uint64_t trace_budget_words = 0;
    uint64_t long_budget_words = 0;   // K2 pairs; 0 = automatic (most of the device)
    int force_L = 0, force_C = 0;
    int fine_max_pairs = 0;           // K2-class launches of at most this many pairs run on K2f (one column per lane); 0 (default): never --
                                      // K2f is bit-exact but, measured on cfg1, no faster than K2 yet (7.8 vs 6.5 ms; see k2f_fine.cuh)
    bool host_plan = false;           // build every launch plan on the host (default: pipeline chunks are planned on the device, k0_plan.cuh)
    int num_sms = 148;
    void set_error(const std::string& s) { std::lock_guard<std::mutex> lk(err_mu); last_error = s; }
};

struct bg_dbatch {
    bg_ctx* ctx; int dev_index;
    uint64_t n_pairs = 0, n_residues = 0;
    DevBuf residues, desc_align, desc_edit;
    std::vector<uint64_t> seq_off;   // host copy, rebased to 0
    Plan plan_align, plan_edit;
    bool edit_lut_ok = false;
    uint8_t edit_lut[256];
};

struct bg_dresult {
    bg_ctx* ctx; int dev_index;
    uint64_t n_pairs = 0;
    int kind = 0;             // 0 align, 1 edit distance
    int mode = 0; bool score_only = false;
    DevBuf score, flags, lens2, off, arena, out64;
};

namespace {

struct HostResultOwner {
    std::vector<std::pair<void*, size_t>> pinned;
    void* grab(size_t bytes) {
        size_t got = 0;
        void* q = pinned_cache().get(bytes, &got);
        if (q) pinned.emplace_back(q, got);
        return q;
    }
    void release_all() { for (auto& q : pinned) pinned_cache().put(q.first, q.second); pinned.clear(); }
};

// ------------------------------------------------------------------------------ planning
Shape pick_shape(const bg_ctx* ctx, uint32_t m, bool half_ok = false, bool c8 = false) {
    if (ctx->force_L) return Shape{ctx->force_L, ctx->force_C};
    return pick_shape_m(m, half_ok, c8);
}

// Scratch vectors of build_plan, recycled across calls and threads.  A fresh 0.5 MB std::vector is an mmap
// whose pages fault in one by one (~3 us each in this VM, serialised on the process's mmap lock when a dozen
// plan threads do it at once): that was a third of the plan time.
struct PlanScratch { std::vector<uint8_t> cls; std::vector<uint32_t> ids, tmp, cnt, slots; };
struct PlanScratchPool {
    std::mutex mu; std::vector<PlanScratch*> free_;
    PlanScratch* get() {
        { std::lock_guard<std::mutex> lk(mu); if (!free_.empty()) { PlanScratch* p = free_.back(); free_.pop_back(); return p; } }
        return new PlanScratch();
    }
    void put(PlanScratch* p) { std::lock_guard<std::mutex> lk(mu); if (free_.size() < 64) free_.push_back(p); else delete p; }
};
PlanScratchPool& plan_scratch_pool() { static PlanScratchPool* p = new PlanScratchPool(); return *p; }
struct PlanScratchLease {
    PlanScratch* s; PlanScratchLease() : s(plan_scratch_pool().get()) {}
    ~PlanScratchLease() { plan_scratch_pool().put(s); }
};

// Is this caller memory pageable (neither allocated pinned nor registered)?  Copies out of pageable memory are staged
// synchronously by the driver (32 ms instead of 14 ms per 10^6-pair call), so the pipelines stage such chunks
// themselves: the chunk's plan task also copies its residues into pinned staging, in parallel with the other chunks.
bool host_is_pageable(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

// Host worker pool: the batch scan and the per-chunk launch plans of every call run here.  Creating a dozen
// std::threads per bg_align_batch call cost ~0.4 ms before the first chunk could be issued.
struct TaskHandle {
    std::shared_ptr<std::atomic<int>> done;
    bool joinable() const { return (bool)done; }
    void join();
};
class HostPool {
  public:
    HostPool() {
        // One worker per core this process can count on: under a one-process-per-GPU launcher (torchrun / MPI
        // export the number of local ranks) the cores are shared, and 8 ranks x 14 concurrent plan tasks on 32 cores
        // made every rank's FIRST chunks late (N = 8: host-to-host step 15 -> 47 ms).  Tasks are taken in FIFO
        // order, so a smaller pool finishes the early chunks' plans first.  BG_HOST_THREADS overrides.
        unsigned hw = std::thread::hardware_concurrency();
        unsigned share = 1;
        for (const char* v : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS", "SLURM_NTASKS_PER_NODE"})
            if (const char* e = getenv(v)) { const int x = atoi(e); if (x > 1) { share = (unsigned)x; break; } }
        // (two cores are left to the threads that drive the GPU -- issuer and finisher of the pipeline: with every core busy
        //  expanding strings their wake-ups came up to 0.7 ms late, measured at the end of a cfg2 call)
        const unsigned avail = std::min(hw ? hw : 8u, 32u) / share;
        unsigned nt = std::max(3u, avail > 4 ? avail - 2 : avail);
        if (const char* e = getenv("BG_HOST_THREADS")) nt = (unsigned)std::max(1, atoi(e));
        for (unsigned t = 0; t < nt; ++t) std::thread([this] { run(); }).detach();
    }
    TaskHandle submit(std::function<void()> fn) {
        TaskHandle h; h.done = std::make_shared<std::atomic<int>>(0);
        { std::lock_guard<std::mutex> lk(mu_); q_.emplace_back(std::move(fn), h.done); }
        cv_.notify_one();
        return h;
    }
    void wait(const std::shared_ptr<std::atomic<int>>& d) {
        std::unique_lock<std::mutex> lk(mu_done_);
        cv_done_.wait(lk, [&] { return d->load(std::memory_order_acquire) != 0; });
    }
  private:
    void run() {
        for (;;) {
            std::pair<std::function<void()>, std::shared_ptr<std::atomic<int>>> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return !q_.empty(); });
                job = std::move(q_.front()); q_.pop_front();
            }
            job.first();
            { std::lock_guard<std::mutex> lk(mu_done_); job.second->store(1, std::memory_order_release); }
            cv_done_.notify_all();
        }
    }
    std::mutex mu_, mu_done_; std::condition_variable cv_, cv_done_;
    std::deque<std::pair<std::function<void()>, std::shared_ptr<std::atomic<int>>>> q_;
};
HostPool& host_pool() { static HostPool* p = new HostPool(); return *p; }   // never destroyed: its threads end with the process
void TaskHandle::join() { if (done) { host_pool().wait(done); done.reset(); } }

// K1h leaves the second slot of a lane group empty when the next pair has a different row count, so a
// plan can hold up to two slots per pair.
size_t plan_desc_capacity(uint64_t n_pairs) { return 2 * (size_t)n_pairs + 8 * MAX_SHAPES; }

// Builds launch classes for pairs [0, n_pairs) whose sequence offsets are off[0 .. 2n] (absolute;
// `base` is subtracted, i.e. the residues of this batch start at device offset 0).  Descriptors go to
// `dst` (capacity plan_desc_capacity(n_pairs)).  with_trace: trace blocks are laid out and the
// launches are cut into chunks that fit the trace budget.
int build_plan(bg_ctx* ctx, const uint64_t* off, uint64_t base, uint64_t n_pairs, bool with_trace,
               uint64_t budget_words, uint64_t wave_budget_words, int32_t half_maxabs, Plan& P, PairDesc* dst, bool myers = false) {
    P = Plan();
    P.half_maxabs = half_maxabs;
    P.myers = myers;
    if (n_pairs >= 0xFFFFFFF0ull) { ctx->set_error("too many pairs in one device batch"); return BG_EINVAL_ARG; }
    Shape shapes[MAX_SHAPES];
    int nshape = 0;
#define X(L_, C_) shapes[nshape++] = Shape{L_, C_};
    BG_SHAPES(X)
#undef X
    const int wave_si = nshape;               // pseudo class: K2 wavefront (alignment only)
    shapes[nshape++] = Shape{32, WAVE_C};
    // pseudo classes: K4b bit-parallel edit distance, one thread per pair, W = 4 / 8 / 10 blocks of 32 columns
    static_assert(MAX_SHAPES >= 15 + 1 + 3, "room for the pseudo classes");
    const int myers_si = nshape;
    shapes[nshape++] = Shape{32, 4}; shapes[nshape++] = Shape{32, 8}; shapes[nshape++] = Shape{32, 10};
    static const bool plan_prof = getenv("BG_PLAN_PROF") != nullptr;
    auto tp0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!plan_prof) return;
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[plan] %s: %.3f ms\n", what, std::chrono::duration<double, std::milli>(t - tp0).count());
        tp0 = t;
    };
    // pass 1: class of every pair
    PlanScratchLease scratch;
    std::vector<uint8_t>& cls = scratch.s->cls;
    cls.resize(n_pairs);
    size_t count[MAX_SHAPES] = {0};
    uint32_t cls_min_n[MAX_SHAPES], cls_max_n[MAX_SHAPES] = {0}, cls_max_m[MAX_SHAPES] = {0};
    for (int s = 0; s < MAX_SHAPES; ++s) cls_min_n[s] = 0xFFFFFFFFu;
    uint32_t last_m = 0xFFFFFFFFu; int last_si = -1; bool last_long = false;
    uint64_t wave_bands16 = 0, wave_cells = 0;    // K2 class: bands at WAVE_C columns per lane, cells
    for (uint64_t p = 0; p < n_pairs; ++p) {
        const uint64_t n = off[2 * p + 1] - off[2 * p], m = off[2 * p + 2] - off[2 * p + 1];
        if (n > 0x7FFFFFF0ull || m > 0x7FFFFFF0ull) { ctx->set_error("sequence longer than 2^31"); return BG_EUNSUPPORTED; }
        const bool is_long = with_trace && n + m > LONG_WALK_LEN;
        if ((uint32_t)m != last_m || is_long != last_long) {
            last_m = (uint32_t)m; last_long = is_long;
            if (with_trace && m > WAVE_MIN_COLS && !ctx->force_L) last_si = wave_si;
            else if (myers && m <= 320) last_si = myers_si + (m <= 128 ? 0 : m <= 256 ? 1 : 2);
            else last_si = shape_index(pick_shape(ctx, last_m, with_trace && half_maxabs > 0 && !is_long, is_long));
            if (last_si < 0) { ctx->set_error("forced kernel shape is not compiled in"); return BG_EINVAL_ARG; }
        }
        cls[p] = (uint8_t)last_si; count[last_si]++;
        if (last_si == wave_si) { wave_bands16 += (m + 32ull * WAVE_C - 1) / (32ull * WAVE_C); wave_cells += n * m; }
        P.cells += n * m;
        cls_min_n[last_si] = std::min<uint32_t>(cls_min_n[last_si], (uint32_t)n);
        cls_max_n[last_si] = std::max<uint32_t>(cls_max_n[last_si], (uint32_t)n);
        cls_max_m[last_si] = std::max<uint32_t>(cls_max_m[last_si], (uint32_t)m);
    }
    for (int s = 0; s < nshape; ++s) { P.max_n = std::max(P.max_n, cls_max_n[s]); P.max_m = std::max(P.max_m, cls_max_m[s]); }
    // K2 with few long pairs (config #1: ONE pair of 17 bands): bands of 8 columns per lane instead of 16 -- twice the warps,
    // half the work per step of each (a lone warp per SM runs at the latency of its cell-to-cell chain, so the step time is
    // what counts).  Only when all of it fits one launch without the bounded-memory path, which is compiled for WAVE_C alone.
    if (count[wave_si]) {
        static const int narrow_env = [] { const char* e = getenv("BG_K2_NARROW"); return e ? atoi(e) : -1; }();
        const bool fits = wave_cells / 8 + (1ull << 22) < wave_budget_words / 2;
        const bool narrow = narrow_env >= 0 ? narrow_env != 0 : (wave_bands16 * 2 <= (uint64_t)std::max(1, ctx->num_sms));
        if (narrow && fits && (int)count[wave_si] > ctx->fine_max_pairs) shapes[wave_si] = Shape{32, WAVE_C_NARROW};
    }
    lap("pass 1");
    // pass 2: bucket pair ids per class
    std::vector<uint32_t>& ids = scratch.s->ids;
    ids.resize(n_pairs);
    size_t start[MAX_SHAPES + 1]; start[0] = 0;
    for (int s = 0; s < nshape; ++s) start[s + 1] = start[s] + count[s];
    {
        size_t cur[MAX_SHAPES];
        for (int s = 0; s < nshape; ++s) cur[s] = start[s];
        for (uint64_t p = 0; p < n_pairs; ++p) ids[cur[cls[p]]++] = (uint32_t)p;
    }
    uint64_t pad_off = 0, bnd_off = 0;
    size_t nd = 0;
    lap("pass 2");
    for (int si = 0; si < nshape; ++si) {
        if (!count[si]) continue;
        uint32_t* cid = ids.data() + start[si];
        const size_t cn = count[si];
        const Shape sh = shapes[si];
        const bool wave = (si == wave_si);
        const uint32_t K = words_per_lane_step(sh.C), band_cols = sh.L * sh.C;
        // K1h (two pairs per lane group, 16-bit halves): short single-band classes whose scores provably fit
        bool half = false;
        // (every single-band shape except (32,5) / (32,8) has a packed instantiation: BG_HALF_SHAPES)
        if (with_trace && half_maxabs > 0 && !wave && (sh.L == 8 || (sh.L == 16 && (sh.C == 10 || sh.C == 16)) || (sh.L == 32 && sh.C >= 12))) {
            const uint32_t cmn = cls_max_n[si], cmm = cls_max_m[si];
            half = cmm <= band_cols && ((int64_t)cmn + cmm + 2) * half_maxabs <= HB_RANGE;
        }
        const uint32_t G = (half ? 2u : 1u) * (32 / sh.L);
        const uint64_t class_budget = wave ? wave_budget_words : budget_words;
        auto len_n = [&](uint32_t id) { return (uint32_t)(off[2 * (uint64_t)id + 1] - off[2 * (uint64_t)id]); };
        auto len_m = [&](uint32_t id) { return (uint32_t)(off[2 * (uint64_t)id + 2] - off[2 * (uint64_t)id + 1]); };
        // longest first so that the lane groups of a warp and the warps of a wave carry similar work;
        // skipped when the class is uniform.
        const bool uniform = cls_min_n[si] == cls_max_n[si];
        const uint32_t n0 = cls_min_n[si];
        if (wave) {
            // K2: largest pairs (cells) first; cluster c then owns pairs c, c + NC, ... of this list
            std::stable_sort(cid, cid + cn, [&](uint32_t x, uint32_t y) {
                return (uint64_t)len_n(x) * len_m(x) > (uint64_t)len_n(y) * len_m(y);
            });
        } else if (!uniform) {
            const uint32_t nmin = n0, nmax = cls_max_n[si];
            const uint64_t range = (uint64_t)nmax - nmin + 1;
            if (range <= (1u << 22) && range <= 4 * cn + 1024) {
                // stable counting sort, descending by row count (a comparison sort of 10^6 ids costs more
                // than the GPU needs for the whole chunk)
                std::vector<uint32_t>& cnt = scratch.s->cnt; std::vector<uint32_t>& tmp = scratch.s->tmp;
                cnt.assign(range + 1, 0); tmp.assign(cid, cid + cn);
                for (size_t k = 0; k < cn; ++k) cnt[nmax - len_n(tmp[k]) + 1]++;
                for (uint64_t r = 0; r < range; ++r) cnt[r + 1] += cnt[r];
                for (size_t k = 0; k < cn; ++k) cid[cnt[nmax - len_n(tmp[k])]++] = tmp[k];
            } else {
                std::stable_sort(cid, cid + cn, [&](uint32_t x, uint32_t y) { return len_n(x) > len_n(y); });
            }
        }
        // launch slots of the class in order; K1h: the two pairs of a lane group must have the same row
        // count (k1h_fill.cuh), so a hole (HOLE) follows a pair whose successor differs
        constexpr uint32_t HOLE = 0xFFFFFFFFu;
        std::vector<uint32_t>& slots_h = scratch.s->slots;
        slots_h.clear();
        if (half && !uniform) {
            slots_h.reserve(cn + cn / 8 + 16);
            for (size_t k = 0; k < cn; ++k) {
                slots_h.push_back(cid[k]);
                if ((slots_h.size() & 1) && (k + 1 == cn || len_n(cid[k + 1]) != len_n(cid[k]))) slots_h.push_back(HOLE);
            }
        }
        const uint32_t* sl = slots_h.empty() ? cid : slots_h.data();
        const size_t sn = slots_h.empty() ? cn : slots_h.size();
        LaunchClass lc; lc.sh = sh; lc.wave = wave; lc.half = half;
        if (si >= myers_si) lc.myers_W = sh.C;
        if (half || si >= myers_si) {
            uint64_t cc = 0;
            for (size_t k = 0; k < cn; ++k) cc += (uint64_t)len_n(cid[k]) * len_m(cid[k]);
            (half ? P.cells_half : P.cells_myers) += cc;
        }
        const size_t wave_clusters = std::max<size_t>(1, (size_t)ctx->num_sms / 4);   // pairs a K2 launch keeps busy at once
        const uint64_t bnd_off_class = bnd_off;
        // K2, bounded-memory traceback: when the whole traces of the next pairs do not fit the budget -- one pair
        // alone, or so few together that most SMs would idle (a pair keeps at most 12 CTAs busy) -- a group of up
        // to `wave_clusters` pairs is cut into nb row blocks each (same nb for all, so they stay in step) such
        // that one block of every pair fits.  ck_group[k] != 0 marks the members (k = position in the sorted list).
        std::vector<uint32_t> ck_group, ck_rb, ck_nb;
        if (wave && with_trace) {
            auto full_words = [&](size_t k) {
                const uint64_t nb_ = (len_m(cid[k]) + band_cols - 1) / band_cols;
                return nb_ * ((uint64_t)len_n(cid[k]) + sh.L - 1) * K * 32ull;
            };
            static const size_t min_fit = [] { const char* e = getenv("BG_CKPT_MIN_FIT"); return e ? (size_t)std::max(1, atoi(e)) : (size_t)6; }();
            uint32_t gid = 0;
            auto fits_from = [&](size_t k) {       // how many pairs from position k on fit together, whole traces
                size_t fit = 0; uint64_t words = 0;
                while (k + fit < cn && fit < min_fit && words + full_words(k + fit) <= class_budget) { words += full_words(k + fit); ++fit; }
                return fit;
            };
            for (size_t k = 0; k < cn;) {
                if (fits_from(k) >= std::min(min_fit, cn - k)) break;    // the list is sorted: the rest fits as well
                if (ck_group.empty()) { ck_group.assign(cn, 0); ck_rb.assign(cn, 0); ck_nb.assign(cn, 0); }
                size_t S = 1;
                while (k + S < cn && S < wave_clusters && fits_from(k + S) < std::min(min_fit, cn - k - S)) ++S;
                uint32_t nb = 0;
                for (;; S = (S + 1) / 2) {
                    uint32_t maxn_g = 0;
                    for (size_t x = k; x < k + S; ++x) maxn_g = std::max(maxn_g, len_n(cid[x]));
                    const uint32_t t_max = std::max(2u, (maxn_g + 63u) / 64u);      // blocks are at least 64 rows
                    for (uint32_t t = 2;; t = std::min(t_max, t < 16 ? t + 1 : t + t / 8)) {
                        uint64_t w2 = 0;
                        for (size_t x = k; x < k + S; ++x) {
                            const uint32_t rb = std::max(64u, ((len_n(cid[x]) + t - 1) / t + 31u) & ~31u);
                            w2 += (uint64_t)((len_m(cid[x]) + band_cols - 1) / band_cols) * ((uint64_t)rb + sh.L - 1) * K * 32ull;
                        }
                        if (w2 <= class_budget) { nb = t; break; }
                        if (t == t_max) break;
                    }
                    if (nb || S == 1) break;
                }
                if (!nb) { ctx->set_error("trace budget too small for 64-row blocks of the longest pair"); return BG_ENOMEM; }
                ++gid;
                for (size_t x = k; x < k + S; ++x) {
                    ck_group[x] = gid; ck_nb[x] = nb;
                    ck_rb[x] = std::max(64u, ((len_n(cid[x]) + nb - 1) / nb + 31u) & ~31u);
                }
                k += S;
            }
        }
        // Opt-in (BG_WAVE_OVERLAP=1), measured and NOT a win: several K2 launches and no bounded-memory group -> cut the
        // launches for HALF the budget, so that two launches' traces are resident at once and the walk of launch c runs
        // next to the fill of launch c + 1 instead of after it.  cfg5, 125 pairs: 5 launches instead of 3, fill 616 -> 715 ms
        // (more launch tails, and the walk's warps compete with the persistent cooperative grid), walks 61 -> 104 ms:
        // 1 060 -> 964 GCUPS.
        uint64_t class_budget_eff = class_budget;
        if (wave && with_trace && ck_group.empty() && cn > 1 && getenv("BG_WAVE_OVERLAP")) {
            uint64_t total = 0, top = 0;
            for (size_t k = 0; k < cn; ++k) {
                const uint64_t nb_ = (len_m(cid[k]) + band_cols - 1) / band_cols;
                const uint64_t wds = nb_ * ((uint64_t)len_n(cid[k]) + sh.L - 1) * K * 32ull;
                total += wds; if (k < 8) top += wds;
            }
            if (total > class_budget && top <= class_budget / 2) { class_budget_eff = class_budget / 2; P.wave_overlap = true; }
        }
        lap("class prologue");
        // (a 32-bit division per pair is a third of the plan's cost: most classes have a single band)
        const bool single_band = cls_max_m[si] <= band_cols;
        auto bands_of = [&](uint32_t m_) -> uint32_t { return single_band ? (m_ ? 1u : 0u) : (m_ + band_cols - 1) / band_cols; };
        // which walker the class gets (run_align): k3_walk (one thread per pair) or, when the class holds long pairs,
        // k3_walk_skew (one warp per pair; needs C % 8 == 0, which pick_shape guarantees for long pairs unless a shape
        // is forced).  Both leave 2-bit ops in the pair's slot.
        lc.long_walk = !half && (sh.C & 7) == 0 && (wave || (uint64_t)cls_max_n[si] + cls_max_m[si] > LONG_WALK_LEN);
        Chunk ch; ch.slot_begin = (uint32_t)nd; ch.trace_words = 0;
        const size_t nwarps = (sn + G - 1) / G;
        for (size_t w = 0; w < nwarps; ++w) {
            uint32_t maxn = 0, maxb = 0;
            for (uint32_t gidx = 0; gidx < G; ++gidx) {
                const size_t k = w * G + gidx;
                if (k >= sn) break;
                if (sl[k] == HOLE) continue;
                maxn = std::max(maxn, len_n(sl[k]));
                maxb = std::max(maxb, bands_of(len_m(sl[k])));
            }
            // K1h: row-block trace layout, HB_TB steps per block, CW words per lane and block
            const uint32_t steps = half ? ((maxn + sh.L - 1 + HB_TB - 1) / HB_TB) * HB_TB : maxn + sh.L - 1;
            uint64_t warp_words = !with_trace ? 0 :
                half ? (uint64_t)((steps / HB_TB + HB_TG_MAX - 1) / HB_TG_MAX) * HB_TG_MAX * 32ull * hb_words_per_lane_block(sh.C) : (uint64_t)maxb * steps * K * 32ull;
            // K2 launches run one pair per resident cluster at a time: close a chunk at a multiple of the
            // cluster count once memory is nearly used up, so that the (length-sorted) pairs of a launch finish together
            const bool wave_round = wave && ch.trace_words > 0 && ((nd - ch.slot_begin) % wave_clusters) == 0 &&
                                    ch.trace_words + warp_words * wave_clusters > class_budget_eff;
            // bounded-memory groups (decided above): a chunk of their own, block-sized traces
            const uint32_t grp = (wave && !ck_group.empty()) ? ck_group[w] : 0u;
            const bool grp_first = grp && (w == 0 || ck_group[w - 1] != grp);
            const bool grp_last = grp && (w + 1 == nwarps || ck_group[w + 1] != grp);
            if (grp) warp_words = (uint64_t)maxb * (ck_rb[w] + sh.L - 1) * K * 32ull;
            if (grp_first && nd > ch.slot_begin) {
                ch.slot_end = (uint32_t)nd;
                lc.chunks.push_back(ch);
                P.max_trace_words = std::max(P.max_trace_words, ch.trace_words);
                P.max_wave_slots = std::max<uint64_t>(P.max_wave_slots, ch.slot_end - ch.slot_begin);
                ch = Chunk(); ch.slot_begin = ch.slot_end = (uint32_t)nd; ch.trace_words = 0;
            }
            if (!grp)
            if (with_trace && ch.trace_words > 0 && (wave_round || ch.trace_words + warp_words > class_budget_eff)) {
                ch.slot_end = (uint32_t)nd;
                lc.chunks.push_back(ch);
                P.max_trace_words = std::max(P.max_trace_words, ch.trace_words);
                if (wave) P.max_wave_slots = std::max<uint64_t>(P.max_wave_slots, ch.slot_end - ch.slot_begin);
                ch.slot_begin = ch.slot_end; ch.trace_words = 0;
            }
            for (uint32_t gidx = 0; gidx < G; ++gidx) {
                const size_t k = w * G + gidx;
                PairDesc d;     // built in registers, stored once (field-wise stores into dst cannot be combined: dst may alias off)
                d.pair_id = 0xFFFFFFFFu; d.steps = steps; d.trace_off = ch.trace_words;
                d.a_off = d.b_off = d.bnd_off = d.pad_off = 0; d.n = d.m = d.nbands = 0; d.pad_ = 1u;
                if (k < sn && sl[k] != HOLE) {
                    const uint64_t id = sl[k];
                    const uint64_t o0 = off[2 * id], o1 = off[2 * id + 1], o2 = off[2 * id + 2];
                    d.a_off = o0 - base; d.b_off = o1 - base;
                    d.n = (uint32_t)(o1 - o0); d.m = (uint32_t)(o2 - o1);
                    d.nbands = bands_of(d.m);
                    d.pair_id = (uint32_t)id;
                    d.pad_off = pad_off; pad_off += 2ull * (((uint64_t)d.n + d.m + 3ull) & ~3ull) + 16ull;   // + header of an op slot (k3_walk.cuh)
                    if (!wave && d.nbands > 1) { d.bnd_off = bnd_off; bnd_off += d.n; }   // K2: assigned below, once the group sizes are known
                }
                dst[nd++] = d;
            }
            if (grp) {
                PairDesc& d = dst[nd - 1];
                d.steps = ck_rb[w] + sh.L - 1;
                P.cells_ckpt += (uint64_t)d.n * d.m;
                ch.trace_words += warp_words;
                P.total_trace_words += warp_words;
                if (grp_last) {
                    // the launch table: entry [0] = pass 1 (whole pair), [1 + x] = row block nb - 1 - x
                    const uint32_t nb = ck_nb[w], nsl = (uint32_t)nd - ch.slot_begin;
                    ch.ckpt_nb = nb;
                    ch.ck_table.assign((size_t)(nb + 1) * nsl, CkptSlot{});
                    uint64_t ck_off = 0;
                    for (uint32_t x = 0; x < nsl; ++x) {
                        const PairDesc& dx = dst[ch.slot_begin + x];
                        const uint32_t rb = dx.steps - (sh.L - 1), stride = (dx.m + 31u) & ~31u;
                        for (uint32_t l = 0; l <= nb; ++l) {
                            CkptSlot& e = ch.ck_table[(size_t)l * nsl + x];
                            e.ck_off = ck_off; e.ck_stride = stride; e.every = rb;
                            if (l == 0) { e.row0 = 0; e.nrows = dx.n; }
                            else {
                                const uint64_t r0 = (uint64_t)(nb - l) * rb;
                                e.row0 = (uint32_t)std::min<uint64_t>(r0, dx.n);
                                e.nrows = r0 < dx.n ? (uint32_t)std::min<uint64_t>(rb, dx.n - r0) : 0u;
                            }
                        }
                        ck_off += (uint64_t)(nb - 1) * stride;
                    }
                    ch.ckpt_elems = ck_off;
                    P.ckpt_elems = std::max<uint64_t>(P.ckpt_elems, ck_off);
                    ch.slot_end = (uint32_t)nd;
                    lc.chunks.push_back(ch);
                    P.max_trace_words = std::max(P.max_trace_words, ch.trace_words);
                    P.max_wave_slots = std::max<uint64_t>(P.max_wave_slots, nsl);
                    ch = Chunk(); ch.slot_begin = ch.slot_end = (uint32_t)nd; ch.trace_words = 0;
                }
                continue;
            }
            ch.trace_words += warp_words;
            P.total_trace_words += warp_words;
        }
        ch.slot_end = (uint32_t)nd;
        lc.chunks.push_back(ch);
        P.max_trace_words = std::max(P.max_trace_words, ch.trace_words);
        if (wave) P.max_wave_slots = std::max<uint64_t>(P.max_wave_slots, ch.slot_end - ch.slot_begin);
        if (wave) {
            // CTA groups of every K2 launch: sizes proportional to the pairs' cell counts (bounded by the number
            // of bands a pair has), CTAs handed out to the pairs -- largest first -- by earliest availability
            bnd_off = bnd_off_class;
            const uint32_t n_cta = (uint32_t)std::max(1, ctx->num_sms);
            for (Chunk& wc : lc.chunks) {
                const uint32_t ns = wc.slot_end - wc.slot_begin;
                // very few pairs in the launch: one column per lane (K2f) spreads a lone pair over hundreds of warps
                if (ns && !wc.ckpt_nb && (int)ns <= ctx->fine_max_pairs) {
                    uint32_t maxb = 0;
                    for (uint32_t x = 0; x < ns; ++x) maxb = std::max(maxb, (dst[wc.slot_begin + x].m + 31u) / 32u);
                    if (maxb <= (uint32_t)std::max(1, ctx->num_sms) * 32u) wc.fine_bands = std::max(1u, maxb);
                }
                double total = 0;
                uint64_t total_bands = 0;
                for (uint32_t x = 0; x < ns; ++x) {
                    total += (double)dst[wc.slot_begin + x].n * (double)dst[wc.slot_begin + x].m;
                    total_bands += dst[wc.slot_begin + x].nbands;
                }
                // warps per CTA: the launch's bands spread over all SMs (a lone 10 kbp pair: 17 bands -> 17 CTAs of one warp)
                static const int wpc_env = [] { const char* e = getenv("BG_K2_WPC"); return e ? std::max(1, std::min((int)K2_WARPS, atoi(e))) : 0; }();
                wc.wpc = wpc_env ? (uint32_t)wpc_env : (uint32_t)std::min<uint64_t>(K2_WARPS, std::max<uint64_t>(1, (total_bands + n_cta - 1) / n_cta));
                const double ideal = std::max(1.0, total / n_cta);                 // cells per CTA if the launch were perfectly balanced
                std::vector<double> avail(n_cta, 0.0);
                std::vector<std::vector<WaveAssign>> lists(n_cta);
                std::vector<uint32_t> order(n_cta);
                for (uint32_t x = 0; x < ns; ++x) {                                  // slots are in largest-first order
                    PairDesc& d = dst[wc.slot_begin + x];
                    const double cells = (double)d.n * (double)d.m;
                    // one warp per band: the whole pair is in flight at once (n + 64 * bands steps), and a warp whose pair
                    // has no band left moves on to its CTA's next pair.  Measured on cfg5 (31 pairs per launch): groups
                    // sized in proportion to the cells (one pair per group, ~5 CTAs) fill in 772 ms, 12 CTAs per pair in 682 ms.
                    uint32_t q = std::max<uint32_t>(1, (d.nbands + wc.wpc - 1) / wc.wpc);
                    (void)ideal;
                    uint32_t q_cap = (wc.wpc == K2_WARPS) ? K2_MAX_Q : n_cta;
                    if (wc.ckpt_nb && ns == 1) {   // a single huge pair may own the machine; its boundary ring (q * 16 + 1 columns of n rows) is kept below 2 GiB
                        const uint64_t ring_q = ((2ull << 30) / sizeof(int2)) / ((((uint64_t)d.n + 31ull) & ~31ull) * wc.wpc);
                        q_cap = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_cta, ring_q));
                    }
                    q = std::min<uint32_t>(std::min<uint32_t>(q, q_cap), n_cta);
                    if (const char* e = getenv("BG_K2_Q")) q = std::min<uint32_t>(std::max(1, atoi(e)), std::min<uint32_t>(q_cap, n_cta));
                    for (uint32_t c = 0; c < n_cta; ++c) order[c] = c;
                    std::partial_sort(order.begin(), order.begin() + q, order.end(),
                                      [&](uint32_t u, uint32_t v) { return avail[u] < avail[v] || (avail[u] == avail[v] && u < v); });
                    double start = 0;
                    for (uint32_t r = 0; r < q; ++r) start = std::max(start, avail[order[r]]);
                    for (uint32_t r = 0; r < q; ++r) {
                        lists[order[r]].push_back(WaveAssign{x, (uint16_t)r, (uint16_t)q});
                        avail[order[r]] = start + cells / q;
                    }
                    wc.max_Q = std::max(wc.max_Q, q);
                    d.bnd_off = bnd_off; bnd_off += ((uint64_t)q * wc.wpc + 1) * (((uint64_t)d.n + 31ull) & ~31ull);
                }
                wc.n_rounds = 0;
                for (auto& l : lists) wc.n_rounds = std::max<uint32_t>(wc.n_rounds, (uint32_t)l.size());
                wc.assign.assign((size_t)wc.n_rounds * n_cta, WaveAssign{0, 0, 0});
                for (uint32_t c = 0; c < n_cta; ++c)
                    for (size_t r = 0; r < lists[c].size(); ++r) wc.assign[r * n_cta + c] = lists[c][r];
                P.max_nw = std::max<uint64_t>(P.max_nw, (uint64_t)wc.max_Q * wc.wpc);
            }
        }
        lap("class slots");
        P.classes.push_back(lc);
    }
    P.n_slots = nd;
    P.pad_bytes = pad_off; P.bnd_elems = bnd_off;
    lap("epilogue");
    P.built = true;
    return BG_OK;
}

// ------------------------------------------------------------------------------ launches (launch.h, l_*.cu)
struct Phase {
    WorkSet& ws; int phase; cudaEvent_t a;
    cudaStream_t st;
    Phase(WorkSet& w, int ph, cudaStream_t s = nullptr) : ws(w), phase(ph), st(s ? s : w.stream) { a = ws.get_event(); cudaEventRecord(a, st); }
    ~Phase() { cudaEvent_t b = ws.get_event(); cudaEventRecord(b, st); ws.evs.push_back(PhaseEv{a, b, phase}); }
};

// One multi-threaded pass over the offsets of a host batch: validity, length statistics and the cost
// (cells) of every block of SCAN_BLOCK pairs -- everything the host-buffer entry points need before
// they can start cutting chunks.  (Three separate serial passes cost ~8 ms per million pairs.)
constexpr uint64_t SCAN_BLOCK = 4096;
// What the device-side planner (k0_plan.cuh) needs to know about the pairs of one kernel shape, per SCAN_BLOCK pairs:
// counted in the same pass that validates the offsets.
struct ClassStat {
    uint64_t count = 0, pad_bytes = 0, bnd_elems = 0, cells = 0;
    uint32_t min_n = 0xFFFFFFFFu, max_n = 0, max_m = 0;
    void add(const ClassStat& o) {
        count += o.count; pad_bytes += o.pad_bytes; bnd_elems += o.bnd_elems; cells += o.cells;
        min_n = std::min(min_n, o.min_n); max_n = std::max(max_n, o.max_n); max_m = std::max(max_m, o.max_m);
    }
};
struct BatchScan {
    bool monotone = true, fitting_violation = false, has_wide = false;
    uint32_t class_mask = 0;          // length classes present (by len2)
    uint64_t max_len_sum = 0, max_m = 0;
    std::vector<double> block_cost;   // per SCAN_BLOCK pairs
    bool with_stats = false;          // block_cls filled (alignment calls with traceback)
    bool half_ok = false; int force_si = -1;
    std::vector<ClassStat> block_cls; // [block][BG_N_SHAPES]
};
// Kernel shape and class-mask bit by len2, tabulated once per process: the scan of a batch with varying lengths asks for
// them once per pair, and pick_shape_m + shape_index (a chain of ~30 compares) was most of its ~35 ns per pair.
struct ScanTables {
    uint8_t si[2][2][WAVE_MIN_COLS + 1];       // [half kernel available][pair long enough for the warp walker][len2] -> shape number
    uint16_t band[BG_N_SHAPES];                // columns of one band of the shape
    uint8_t mask_bit[WAVE_MIN_COLS + 1];       // BatchScan::class_mask bit of len2
    ScanTables() {
        for (int h = 0; h < 2; ++h)
            for (int lg = 0; lg < 2; ++lg)
                for (uint32_t m = 0; m <= WAVE_MIN_COLS; ++m)
                    si[h][lg][m] = (uint8_t)shape_index(pick_shape_m(m, h && !lg, lg != 0));
        for (int k = 0; k < BG_N_SHAPES; ++k) { const Shape sh = shape_at(k); band[k] = (uint16_t)(sh.L * sh.C); }
        for (uint32_t m = 0; m <= WAVE_MIN_COLS; ++m)
            mask_bit[m] = (uint8_t)(m <= 64 ? 0 : m <= 96 ? 1 : m <= 128 ? 2 : m <= 160 ? 3 : m <= 192 ? 4 : m <= 256 ? 5 : m <= 384 ? 6 :
                                    m <= 512 ? 7 : m <= 640 ? 8 : m <= 768 ? 9 : m <= 1024 ? 10 : 11);
    }
};
const ScanTables& scan_tables() { static const ScanTables* t = new ScanTables(); return *t; }

void scan_batch(const bg_batch* in, BatchScan& S) {
    const ScanTables& T = scan_tables();
    const uint64_t N = in->n_pairs;
    const uint64_t nblocks = (N + SCAN_BLOCK - 1) / SCAN_BLOCK;
    S.block_cost.assign(nblocks, 0.0);
    if (S.with_stats) S.block_cls.assign(nblocks * BG_N_SHAPES, ClassStat());
    unsigned nt = std::thread::hardware_concurrency();
    nt = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(std::min<unsigned>(nt ? nt : 1, 16), nblocks / 8));
    std::vector<BatchScan> part(nt);
    auto work = [&](unsigned t) {
        BatchScan& P = part[t];
        const uint64_t b_lo = nblocks * t / nt, b_hi = nblocks * (t + 1) / nt;
        const uint64_t* off = in->seq_off;
        uint64_t last_m = ~0ull;
        uint64_t st_m = ~0ull; bool st_long = false; int st_si = 0; uint32_t st_band = 1;
        for (uint64_t blk = b_lo; blk < b_hi; ++blk) {
            const uint64_t q_hi = std::min(N, (blk + 1) * SCAN_BLOCK);
            double cost = 0;
            ClassStat* cs = S.with_stats ? &S.block_cls[blk * BG_N_SHAPES] : nullptr;
            // runs of pairs with the same two lengths (read sets are mostly uniform) are accounted once
            uint64_t run_n = ~0ull, run_m = ~0ull, run_len = 0;
            auto flush = [&] {
                if (!run_len) return;
                const uint64_t n = run_n, m = run_m, k = run_len;
                run_len = 0;
                if (n < m) P.fitting_violation = true;
                if (m > WAVE_MIN_COLS) P.has_wide = true;
                if (cs && m <= WAVE_MIN_COLS && n < 0x7FFFFFF0ull) {
                    const bool is_long = n + m > LONG_WALK_LEN;
                    if (m != st_m || is_long != st_long) {
                        st_m = m; st_long = is_long;
                        st_si = S.force_si >= 0 ? S.force_si : (int)T.si[S.half_ok ? 1 : 0][is_long ? 1 : 0][m];
                        st_band = T.band[st_si];
                    }
                    ClassStat& c = cs[st_si];
                    c.count += k; c.cells += k * n * m;
                    c.pad_bytes += k * (2ull * ((n + m + 3ull) & ~3ull) + 16ull);
                    if (m > st_band) c.bnd_elems += k * n;
                    c.min_n = std::min<uint32_t>(c.min_n, (uint32_t)n); c.max_n = std::max<uint32_t>(c.max_n, (uint32_t)n);
                    c.max_m = std::max<uint32_t>(c.max_m, (uint32_t)m);
                }
                if (m != last_m) {
                    last_m = m;
                    P.class_mask |= 1u << (m <= WAVE_MIN_COLS ? T.mask_bit[m] : 11);
                }
                P.max_len_sum = std::max(P.max_len_sum, n + m);
                P.max_m = std::max(P.max_m, m);
                cost += (double)k * ((double)n * (double)m + 64.0);
            };
            for (uint64_t q = blk * SCAN_BLOCK; q < q_hi; ++q) {
                const uint64_t o0 = off[2 * q], o1 = off[2 * q + 1], o2 = off[2 * q + 2];
                if (o1 < o0 || o2 < o1) { P.monotone = false; continue; }
                const uint64_t n = o1 - o0, m = o2 - o1;
                if (n == run_n && m == run_m) { ++run_len; continue; }
                flush();
                run_n = n; run_m = m; run_len = 1;
            }
            flush();
            S.block_cost[blk] = cost;
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<TaskHandle> th;
        for (unsigned t = 1; t < nt; ++t) th.push_back(host_pool().submit([&work, t] { work(t); }));
        work(0);
        for (auto& x : th) x.join();
    }
    for (auto& P : part) {
        S.monotone = S.monotone && P.monotone;
        S.fitting_violation = S.fitting_violation || P.fitting_violation;
        S.has_wide = S.has_wide || P.has_wide;
        S.class_mask |= P.class_mask;
        S.max_len_sum = std::max(S.max_len_sum, P.max_len_sum);
        S.max_m = std::max(S.max_m, P.max_m);
    }
}

// Pipeline chunks for pairs [lo, hi) at SCAN_BLOCK granularity, roughly equal cell counts.
// nd > 1: the chunks feed a queue that nd devices share -- nd times as many chunks, and the ramp goes by rounds of nd.
// ramp_down: the chunks shrink towards the end (what follows the GPU's last kernel -- the string expansion of the last
// chunk on the host, formerly its D2H -- is proportional to the last chunk's size); compact-result calls have no such tail.
std::vector<uint64_t> chunk_bounds_from_scan(const BatchScan& S, uint64_t lo, uint64_t hi, double nchunk_override = 0.0, int nd = 1, bool ramp_down = true) {
    std::vector<uint64_t> b{lo};
    if (hi == lo) { b.push_back(hi); return b; }
    const uint64_t blk_lo = lo / SCAN_BLOCK, blk_hi = (hi + SCAN_BLOCK - 1) / SCAN_BLOCK;
    double total = 0;
    for (uint64_t k = blk_lo; k < blk_hi; ++k) total += S.block_cost[k];
    static const double nchunk_target = [] { const char* e = getenv("BG_PIPE_CHUNKS"); return e ? std::max(1.0, atof(e)) : 4.0; }();   // (cfg2, compact results: 3 / 4 / 5 / 6 -> 12.5 / 12.5 / 12.9 / 13.6 ms per call)
    const double target = nchunk_override > 0 ? std::max(total / nchunk_override, 1.0e8) : std::max(total / (nchunk_target * nd), 1.0e9);
    const uint64_t max_pairs = 262144;
    // every length class of a chunk becomes its own launch: keep >= ~4 waves of warps per launch
    const uint64_t min_pairs = 8192ull * (uint64_t)__builtin_popcount(S.class_mask ? S.class_mask : 1u);
    double acc = 0, done_cost = 0; uint64_t start = lo;
    for (uint64_t k = blk_lo; k < blk_hi; ++k) {
        acc += S.block_cost[k];
        const uint64_t end = std::min(hi, (k + 1) * SCAN_BLOCK);
        if (end <= start) continue;
        // ramp: the first two chunks are a quarter / a half of the rest, so planning + H2D of chunk 0 is short
        // ... and ramp down: the last chunk's strings travel D2H after the GPU has gone idle, so the chunks
        // shrink towards the end (half of what is left, but not below a quarter of the regular size)
        const size_t nb = (b.size() - 1) / (size_t)nd + 1;
        // (measured, cfg2: a plan costs ~28 ns per pair on one host thread, the GPU aligns a pair in ~12 ns; all plans
        //  start together, so chunk c's plan is ready in time only if it is < ~0.4 of everything before it)
        static const double ramp[] = {0.125, 0.1875, 0.25, 0.375, 0.5, 0.75};
        const double scale = nb <= 6 ? ramp[nb - 1] : 1.0;
        const double left = total - done_cost;
        const double want = ramp_down ? std::min(target * scale, std::max(left * 0.5, target * 0.25)) : target * scale;
        if ((acc >= want && (double)(end - start) >= (double)min_pairs * scale) || (double)(end - start) >= (double)max_pairs * scale) {
            b.push_back(end); start = end; done_cost += acc; acc = 0;
        }
    }
    if (b.back() != hi) b.push_back(hi);
    return b;
}

// Launch plan of pairs [lo, hi) (lo a multiple of SCAN_BLOCK; hi a multiple or the batch end) for the DEVICE-side
// planner: classes, descriptor ranges and trace regions sized from the scan's per-class statistics (upper bounds
// where the exact value depends on the order the device will establish: K1h holes, per-warp step counts).  Returns
// false when the item needs the host planner: a forced split into several launches per class (trace budget), too
// many pairs, no statistics.
bool plan_from_stats(const bg_ctx* ctx, const BatchScan& S, uint64_t lo, uint64_t hi, uint64_t budget_words, int32_t half_maxabs,
                     Plan& P, PlanArgs& A, bool& need_sort) {
    if (!S.with_stats || hi <= lo || (lo % SCAN_BLOCK) != 0 || hi - lo >= 0x7FFFFFF0ull) return false;
    ClassStat cls[BG_N_SHAPES];
    for (uint64_t blk = lo / SCAN_BLOCK; blk * SCAN_BLOCK < hi; ++blk)
        for (int s = 0; s < BG_N_SHAPES; ++s) cls[s].add(S.block_cls[blk * BG_N_SHAPES + s]);
    uint64_t counted = 0;
    for (int s = 0; s < BG_N_SHAPES; ++s) counted += cls[s].count;
    if (counted != hi - lo) return false;          // pairs outside the K1 / K1h classes (wide pairs, oversize)
    P = Plan();
    P.half_maxabs = half_maxabs;
    A = PlanArgs();
    A.n_pairs = (uint32_t)(hi - lo); A.half_ok = S.half_ok ? 1u : 0u; A.force_si = S.force_si; A.n_cls = 0;
    for (int s = 0; s < 16; ++s) A.rank_of_shape[s] = -1;
    uint64_t nd = 0, pad_off = 0, bnd_off = 0, sorted = 0;
    int n_present = 0; bool uniform_n = true;
    for (int si = 0; si < BG_N_SHAPES; ++si) {
        const ClassStat& c = cls[si];
        if (!c.count) continue;
        if (A.n_cls >= PLAN_MAX_CLS) return false;
        ++n_present;
        if (c.min_n != c.max_n) uniform_n = false;
        const Shape sh = shape_at(si);
        const uint32_t K = words_per_lane_step(sh.C), band_cols = (uint32_t)(sh.L * sh.C);
        bool half = false;
        if (half_maxabs > 0 && shape_has_half(sh))
            half = c.max_m <= band_cols && ((int64_t)c.max_n + c.max_m + 2) * half_maxabs <= HB_RANGE;
        const uint32_t G2 = (half ? 2u : 1u) * (32u / (uint32_t)sh.L);
        const uint64_t holes_ub = half ? std::min<uint64_t>(c.count, (uint64_t)c.max_n - c.min_n + 1) : 0;
        const uint64_t slot_cap = (c.count + holes_ub + G2 - 1) / G2 * G2;
        const uint64_t nwarps = slot_cap / G2;
        uint64_t warp_words;
        if (half) {
            const uint32_t steps = ((c.max_n + sh.L - 1 + HB_TB - 1) / HB_TB) * HB_TB;
            warp_words = (uint64_t)((steps / HB_TB + HB_TG_MAX - 1) / HB_TG_MAX) * HB_TG_MAX * 32ull * hb_words_per_lane_block(sh.C);
        } else {
            const uint64_t maxb = c.max_m ? (c.max_m + band_cols - 1) / band_cols : 0;
            warp_words = maxb * ((uint64_t)c.max_n + sh.L - 1) * K * 32ull;
        }
        const uint64_t trace_words = nwarps * warp_words;
        if (trace_words > budget_words || nd + slot_cap >= 0xFFFFFFF0ull) return false;
        LaunchClass lc; lc.sh = sh; lc.half = half;
        lc.long_walk = !half && (sh.C & 7) == 0 && (uint64_t)c.max_n + c.max_m > LONG_WALK_LEN;
        Chunk ch; ch.slot_begin = (uint32_t)nd; ch.slot_end = (uint32_t)(nd + slot_cap); ch.trace_words = trace_words;
        lc.chunks.push_back(ch);
        P.classes.push_back(lc);
        PlanCls& pc = A.cls[A.n_cls];
        pc.L = sh.L; pc.C = sh.C; pc.half = half ? 1u : 0u; pc.G2 = G2;
        pc.sorted_begin = (uint32_t)sorted; pc.count = (uint32_t)c.count;
        pc.slot_begin = (uint32_t)nd; pc.slot_cap = (uint32_t)slot_cap;
        pc.pad_base = pad_off; pc.bnd_base = bnd_off;
        A.rank_of_shape[si] = (int8_t)A.n_cls;
        ++A.n_cls;
        nd += slot_cap; sorted += c.count; pad_off += c.pad_bytes; bnd_off += c.bnd_elems;
        P.cells += c.cells; if (half) P.cells_half += c.cells;
        P.total_trace_words += trace_words; P.max_trace_words = std::max(P.max_trace_words, trace_words);
        P.max_n = std::max(P.max_n, c.max_n); P.max_m = std::max(P.max_m, c.max_m);
    }
    P.n_slots = nd; P.pad_bytes = pad_off; P.bnd_elems = bnd_off;
    P.built = true;
    (void)ctx;
    need_sort = !(n_present == 1 && uniform_n);
    // one class, one len1, and -- every len2 <= max_m, so equal sums mean equal terms -- one len2
    static const bool no_uniform = getenv("BG_NO_UNIFORM_PLAN") != nullptr;
    A.uniform = 0;
    if (!need_sort && !no_uniform && A.n_cls == 1)
        for (int si = 0; si < BG_N_SHAPES; ++si)
            if (cls[si].count && cls[si].max_n > 0 && cls[si].cells == cls[si].count * (uint64_t)cls[si].max_n * (uint64_t)cls[si].max_m) A.uniform = 1;
    return true;
}

// scan != nullptr: the host-buffer entry points validate and measure the batch in one multi-threaded pass.
int check_batch(bg_ctx* ctx, const bg_batch* in, BatchScan* scan = nullptr) {
    if (!in || (in->n_pairs && (!in->seq_off || (!in->residues && in->seq_off[2 * in->n_pairs] != in->seq_off[0])))) {
        ctx->set_error("null batch pointers"); return BG_EINVAL_ARG;
    }
    if (in->packing != BG_PACK_NONE && ((in->packing != BG_PACK_2BIT && in->packing != BG_PACK_5BIT) || !in->alphabet)) {
        ctx->set_error("unknown residue packing, or packed batch without an alphabet"); return BG_EINVAL_ARG;
    }
    if (scan) {
        scan_batch(in, *scan);
        if (!scan->monotone) { ctx->set_error("seq_off not monotone"); return BG_EINVAL_ARG; }
        return BG_OK;
    }
    for (uint64_t s = 0; s < 2 * in->n_pairs; ++s)
        if (in->seq_off[s + 1] < in->seq_off[s]) { ctx->set_error("seq_off not monotone"); return BG_EINVAL_ARG; }
    return BG_OK;
}

// > 0: the packed 16 x 2 kernel (K1h) may be used with these parameters, with this bound on |score| (0: not).
int32_t half_maxabs_of(const bg_params* p) {
    if (!p || !p->table || p->n_rows <= 0 || p->n_cols <= 0 || p->n_rows > 4 || p->n_cols > 4) return 0;
    if (p->mode == BG_LOCAL || (p->flags & BG_F_SCORE_ONLY) || getenv("BG_NO_HALF")) return 0;
    int64_t maxabs = std::max<int64_t>(llabs((long long)p->gap_open), llabs((long long)p->gap_extend));
    for (int i = 0; i < p->n_rows * p->n_cols; ++i) maxabs = std::max<int64_t>(maxabs, llabs((long long)p->table[i]));
    return maxabs <= HB_MAXABS ? (int32_t)std::max<int64_t>(1, maxabs) : 0;
}

// Validates bg_params against the reference's rules and the engine's numeric range.
int prepare_params(bg_ctx* ctx, const bg_params* p, const uint64_t* off, uint64_t N, Prepared& pp, const BatchScan* scan = nullptr) {
    const int mode = p->mode;
    if (mode < BG_GLOBAL || mode > BG_OVERLAP) { ctx->set_error("unknown mode"); return BG_EINVAL_ARG; }
    // aligner.rs:87-89,153-155,219-221: sign check in global / local / fitting only
    if ((mode == BG_GLOBAL || mode == BG_LOCAL || mode == BG_FITTING) && (p->gap_open > 0 || p->gap_extend > 0)) return BG_EINVAL_RANGE;
    if (!p->table || !p->row_code || !p->col_code || p->n_rows <= 0 || p->n_cols <= 0 || p->n_rows > 255 || p->n_cols > 255) {
        ctx->set_error("score table missing or malformed"); return BG_EINVAL_ARG;
    }
    uint64_t max_len_sum = 0;
    if (scan) {
        if (mode == BG_FITTING && scan->fitting_violation) return BG_EINVAL_SIZE;   // aligner.rs:223-225
        max_len_sum = scan->max_len_sum;
    } else {
        for (uint64_t q = 0; q < N; ++q) {
            const uint64_t n = off[2 * q + 1] - off[2 * q], m = off[2 * q + 2] - off[2 * q + 1];
            if (mode == BG_FITTING && n < m) return BG_EINVAL_SIZE;   // aligner.rs:223-225
            max_len_sum = std::max(max_len_sum, n + m);
        }
    }
    pp.mode = mode; pp.a = p->gap_open; pp.b = p->gap_extend;
    pp.local = (mode == BG_LOCAL); pp.score_only = (p->flags & BG_F_SCORE_ONLY) != 0;
    pp.n_rows = p->n_rows; pp.n_cols = p->n_cols;
    pp.table.assign(p->table, p->table + (size_t)p->n_rows * p->n_cols);
    int64_t maxabs = std::max<int64_t>(llabs((long long)p->gap_open), llabs((long long)p->gap_extend));
    bool fits8 = true;   // the packed byte profile holds s - a (k1_fill.cuh)
    for (int32_t v : pp.table) {
        maxabs = std::max<int64_t>(maxabs, llabs((long long)v));
        const int64_t biased = (int64_t)v - p->gap_open;
        if (biased < -128 || biased > 127) fits8 = false;
    }
    if (-(int64_t)p->gap_open < -128 || -(int64_t)p->gap_open > 127) fits8 = false;
    // 32-bit safety of the recurrence (bg_common.cuh NEG_INF)
    if (maxabs > (1 << 20) || (int64_t)(max_len_sum + 2) * maxabs >= (1ll << 28)) {
        ctx->set_error("scores * length exceed the 32-bit-safe range"); return BG_EUNSUPPORTED;
    }
    pp.maxabs = maxabs;
    pp.smem = 512 + (size_t)p->n_rows * (p->n_cols + 1) * 4;
    if (pp.smem > 48 * 1024) { ctx->set_error("score table too large for shared memory"); return BG_EUNSUPPORTED; }
    pp.prof4 = fits8 && p->n_rows <= 4;
    pp.half_maxabs = half_maxabs_of(p);
    pp.half_prof8 = !getenv("BG_NO_HALF_PROF");
    {
        const int64_t ab = (int64_t)p->gap_open + p->gap_extend;
        if (-ab < -128 || -ab > 127) pp.half_prof8 = false;
        for (int32_t v : pp.table) if (v - ab < -128 || v - ab > 127) pp.half_prof8 = false;
    }
    memcpy(pp.codes, p->row_code, 256); memcpy(pp.codes + 256, p->col_code, 256);
    for (int i = 0; i < 256; ++i) {
        if (pp.codes[i] != 0xFF && pp.codes[i] >= p->n_rows) { ctx->set_error("row_code entry out of range"); return BG_EINVAL_ARG; }
        if (pp.codes[256 + i] != 0xFF && pp.codes[256 + i] >= p->n_cols) { ctx->set_error("col_code entry out of range"); return BG_EINVAL_ARG; }
    }
    return BG_OK;
}

struct AlignIO {
    const uint8_t* residues; const PairDesc* desc; const Plan* plan; uint64_t N;
    int32_t* score; uint8_t* flags; uint64_t* lens2; uint64_t* off; uint8_t* arena;
    // compact results (host-buffer entry points): ops != nullptr -> instead of strings in `arena`, the pairs' 2-bit
    // ops are packed densely into `ops` at word offsets `off` ([N + 1], exclusive scan of ceil(len / 16)), next to
    // len / first; `lens2` is still the walkers' output
    uint32_t* len = nullptr; uint32_t* first = nullptr; uint32_t* ops = nullptr;
    ulonglong2* samples = nullptr; uint64_t sample_stride = 0;   // every sample_stride-th entry of the {op words, columns} scan + totals
};

// Host bytes [b0, b1) of a batch's residue arena that hold residues [r0, r1).
void host_byte_range(uint32_t packing, uint64_t r0, uint64_t r1, uint64_t& b0, uint64_t& b1) {
    if (packing == BG_PACK_NONE) { b0 = r0; b1 = r1; }
    else packed_byte_range(packing, r0, r1, b0, b1);
}

// Residues [r0, r1) of a host batch -> dst[0 .. r1 - r0) on the device, one byte per residue.  src points at host byte
// b0 of host_byte_range(packing, r0, r1) (the caller's arena, or a pinned staging copy of exactly that range).
// Packed batches: the packed bytes go to `staging` and k_unpack writes dst (same stream).
int upload_residues(bg_ctx* ctx, const uint8_t* src, uint32_t packing, const uint8_t* alphabet, uint64_t r0, uint64_t r1,
                    DevBuf& staging, uint8_t* dst, cudaStream_t st, uint64_t* h2d_bytes) {
    if (r1 <= r0) return BG_OK;
    uint64_t b0, b1;
    host_byte_range(packing, r0, r1, b0, b1);
    if (h2d_bytes) *h2d_bytes += b1 - b0;
    if (packing == BG_PACK_NONE) {
        CU_TRY(ctx, cudaMemcpyAsync(dst, src, b1 - b0, cudaMemcpyHostToDevice, st));
        return BG_OK;
    }
    const uint64_t lead = b0 & 15;                      // keep the bytes' alignment mod 16 (128-bit loads in k_unpack2)
    if (!staging.ensure(lead + (b1 - b0) + 64)) { ctx->set_error("device allocation failed (packed residues)"); return BG_ENOMEM; }
    CU_TRY(ctx, cudaMemcpyAsync(staging.as<uint8_t>() + lead, src, b1 - b0, cudaMemcpyHostToDevice, st));
    UnpackArgs ua;
    ua.packed = staging.as<uint8_t>(); ua.bit0 = (uint64_t)packing * r0 - 8 * (b0 - lead); ua.count = r1 - r0; ua.bits = packing; ua.out = dst;
    memcpy(ua.alphabet, alphabet, packing == BG_PACK_2BIT ? 4 : 32);
    launch_unpack(ua, st);
    CU_TRY(ctx, cudaGetLastError());
    return BG_OK;
}

// Uploads the score table / code maps into the work set and clears its error flag.
int upload_params(bg_ctx* ctx, WorkSet& ws, const Prepared& pp) {
    if (!ws.table.ensure(pp.table.size() * 4) || !ws.codes.ensure(512) || !ws.err.ensure(4)) {
        ctx->set_error("device allocation failed (parameters)"); return BG_ENOMEM;
    }
    CU_TRY(ctx, cudaMemcpyAsync(ws.table.p, pp.table.data(), pp.table.size() * 4, cudaMemcpyHostToDevice, ws.stream));
    CU_TRY(ctx, cudaMemcpyAsync(ws.codes.p, pp.codes, 512, cudaMemcpyHostToDevice, ws.stream));
    CU_TRY(ctx, cudaMemsetAsync(ws.err.p, 0, 4, ws.stream));
    return BG_OK;
}

// Enqueues fill -> walk -> scan -> gather for one planned batch on ws.stream.  Asynchronous.
int run_align(bg_ctx* ctx, WorkSet& ws, const AlignIO& io, const Prepared& pp) {
    const Plan& P = *io.plan;
    const uint64_t N = io.N;
    bool ok = ws.end.ensure(std::max<size_t>(1, P.n_slots) * sizeof(EndCell));
    ok = ok && ws.bnd.ensure(std::max<uint64_t>(1, P.bnd_elems) * sizeof(int2));
    size_t n_chunks = 0;
    for (const LaunchClass& lc : P.classes) n_chunks += lc.chunks.size();
    // Host-buffer pipeline: everything after a fill runs on the work set's post stream, so that the next fill
    // (integer-pipe bound) overlaps the walk / scan / gather of the previous launch (latency bound, small
    // grids).  Every launch then needs its own trace region, which pipeline chunks can afford.
    const bool split = ws.post_stream != nullptr && !pp.score_only && P.max_wave_slots == 0 &&
                       (n_chunks == 1 || P.total_trace_words <= ws.split_cap_words);
    if (!pp.score_only) {
        ok = ok && ws.trace.ensure(std::max<uint64_t>(1, split ? P.total_trace_words : P.max_trace_words) * 4);
        ok = ok && ws.pad.ensure(std::max<uint64_t>(1, P.pad_bytes));
    }
    if (P.max_wave_slots) {
        const uint64_t nw = P.max_nw;
        // progress counters, then one "workers done" counter per pair
        ok = ok && ws.progress.ensure(P.max_wave_slots * (nw + 1) * 8 + (2 * P.max_wave_slots + 4) * 4) && ws.cand.ensure(P.max_wave_slots * nw * sizeof(WaveCand));
    }
    if (!ok) { ctx->set_error("device allocation failed (trace / scratch buffers)"); return BG_ENOMEM; }
    cudaStream_t st = ws.stream;
    if (!pp.score_only) CU_TRY(ctx, cudaMemsetAsync(io.lens2, 0, (2 * N + 1) * 8, st));

    FillArgs fa;
    fa.desc = nullptr; fa.n_slots = 0;
    fa.residues = io.residues;
    fa.table = ws.table.as<int32_t>(); fa.n_rows = pp.n_rows; fa.n_cols = pp.n_cols;
    fa.row_code = ws.codes.as<uint8_t>(); fa.col_code = ws.codes.as<uint8_t>() + 256;
    fa.a = pp.a; fa.b = pp.b; fa.mode = pp.mode; fa.want_trace = pp.score_only ? 0 : 1; fa.one = 1;
    static const int tg_shift = [] { const char* e = getenv("BG_K1H_TG"); const int v = e ? atoi(e) : 1; return v < 0 ? 0 : v > 2 ? 2 : v; }();   // measured on cfg2: 1 (pairs of row blocks) is best
    fa.tg_shift = tg_shift;
    fa.trace = ws.trace.as<uint32_t>(); fa.bnd = ws.bnd.as<int2>(); fa.end = nullptr; fa.err_flag = ws.err.as<uint32_t>();

    // Chunks alternate between two trace buffers; the walk of chunk c runs on its own stream next to the
    // fill of chunk c+1 (the fill is bound by the integer pipes, the walk by memory requests, so they
    // overlap well on the same SMs).  Events order: fill(c) -> walk(c) -> fill(c+2) (buffer reuse).
    // Measured on B200 (cfg2): overlapping buys 2 % (17.9 vs 18.3 ms/step) -- both kernels just time-share the
    // SMs -- and it blurs the per-kernel event timings the roofline is computed from, so it is opt-in.
    static const bool want_overlap = [] { const char* e = getenv("BG_OVERLAP"); return e && atoi(e) != 0; }();
    const bool overlap = !split && !pp.score_only && n_chunks >= 2 && ((want_overlap && P.max_wave_slots == 0) || P.wave_overlap);
    if (overlap && !ws.trace2.ensure(std::max<uint64_t>(1, P.max_trace_words) * 4)) { ctx->set_error("device allocation failed (second trace buffer)"); return BG_ENOMEM; }
    cudaStream_t wst = overlap ? ws.walk_stream : split ? ws.post_stream : st;
    cudaStream_t pst = split ? ws.post_stream : st;
    cudaStream_t st2 = (split && ws.fill2_stream && (n_chunks >= 2 || ws.launch_parity >= 0)) ? ws.fill2_stream : nullptr;
    const size_t parity = ws.launch_parity > 0 ? 1 : 0;
    if (st2) {
        cudaEvent_t ev0 = ws.get_event();
        CU_TRY(ctx, cudaEventRecord(ev0, st));
        CU_TRY(ctx, cudaStreamWaitEvent(st2, ev0, 0));       // inputs, parameters and memsets issued on the main stream so far
    }
    cudaEvent_t ev_walk_done[2] = {nullptr, nullptr};
    if (overlap) {
        cudaEvent_t ev0 = ws.get_event();
        CU_TRY(ctx, cudaEventRecord(ev0, st));
        CU_TRY(ctx, cudaStreamWaitEvent(wst, ev0, 0));     // parameters / memsets issued on the main stream so far
    }
    size_t chunk_no = 0;
    uint64_t trace_base = 0;   // split: launches do not share trace memory
    for (const LaunchClass& lc : P.classes) {
        const uint32_t G = 32 / lc.sh.L;
        for (const Chunk& ch : lc.chunks) {
            const uint32_t ns = ch.slot_end - ch.slot_begin;
            if (!ns) continue;
            const int buf = overlap ? (int)(chunk_no & 1) : 0;
            ++chunk_no;
            fa.trace = (buf ? ws.trace2 : ws.trace).as<uint32_t>() + (split ? trace_base : 0);
            trace_base += ch.trace_words;
            if (overlap && ev_walk_done[buf]) CU_TRY(ctx, cudaStreamWaitEvent(st, ev_walk_done[buf], 0));
            fa.desc = io.desc + ch.slot_begin;
            fa.end = ws.end.as<EndCell>() + ch.slot_begin;
            fa.n_slots = ns;
            const uint32_t nwarps = (ns + G - 1) / G;
            cudaStream_t fst = st;          // stream of this launch's fill
            if (lc.wave && ch.ckpt_nb) {
                // ---- bounded-memory traceback: pass 1 (checkpoints + end cells), then row block by row block,
                //      bottom-up, re-fill with direction codes + resume the walks (k2_wave.cuh) ----
                const uint32_t NB = ch.ckpt_nb;
                if (!ws.ckpt.ensure(std::max<uint64_t>(1, ch.ckpt_elems) * sizeof(int2)) || !ws.wstate.ensure(ns * sizeof(WalkState)) ||
                    !ws.ckslots.ensure(ch.ck_table.size() * sizeof(CkptSlot)) ||
                    !ws.assign.ensure(std::max<size_t>(1, ch.assign.size()) * sizeof(WaveAssign))) {
                    ctx->set_error("device allocation failed (row checkpoints)"); return BG_ENOMEM;
                }
                CU_TRY(ctx, cudaMemsetAsync(ws.wstate.p, 0, ns * sizeof(WalkState), st));
                CU_TRY(ctx, cudaMemcpyAsync(ws.assign.p, ch.assign.data(), ch.assign.size() * sizeof(WaveAssign), cudaMemcpyHostToDevice, st));
                CU_TRY(ctx, cudaMemcpyAsync(ws.ckslots.p, ch.ck_table.data(), ch.ck_table.size() * sizeof(CkptSlot), cudaMemcpyHostToDevice, st));
                const uint64_t nw = (uint64_t)ch.max_Q * ch.wpc;
                const uint64_t prog_bytes = (uint64_t)ns * (nw + 1) * 8;
                WaveArgs wa; wa.f = fa; wa.progress = ws.progress.as<unsigned long long>(); wa.cand = ws.cand.as<WaveCand>();
                wa.assign = ws.assign.as<WaveAssign>(); wa.n_rounds = ch.n_rounds;
                wa.prog_stride = (uint32_t)(nw + 1); wa.cand_stride = (uint32_t)nw;
                wa.done = reinterpret_cast<uint32_t*>(ws.progress.as<unsigned char>() + prog_bytes);
                wa.next_band = wa.done + ns + 2;
                wa.ckpt = ws.ckpt.as<int2>();
                wa.wstate = ws.wstate.as<WalkState>();
                WalkArgs wk;
                wk.desc = fa.desc; wk.end = fa.end; wk.n_slots = ns; wk.residues = fa.residues;
                wk.trace = fa.trace; wk.mode = pp.mode; wk.L = lc.sh.L; wk.C = lc.sh.C; wk.H = 1; wk.tg_shift = fa.tg_shift; wk.CW = 0;
                wk.pad = ws.pad.as<uint8_t>(); wk.score = io.score; wk.walk_flags = io.flags; wk.lens2 = io.lens2;
                wk.wstate = ws.wstate.as<WalkState>();
                for (uint32_t l = 0; l <= NB; ++l) {
                    {
                        Phase ph(ws, 1);
                        CU_TRY(ctx, cudaMemsetAsync(ws.progress.p, 0, prog_bytes + (2 * (uint64_t)ns + 4) * 4, st));
                        wa.cks = ws.ckslots.as<CkptSlot>() + (size_t)l * ns;
                        wa.f.want_trace = l ? 1 : 0; wa.ckpt_write = l ? 0 : 1; wa.write_end = l ? 0 : 1;
                        CU_TRY(ctx, launch_k2(pp.local, pp.prof4, lc.sh.C, ctx->num_sms, (int)ch.wpc, pp.smem, st, wa, true));
                    }
                    if (l) {
                        Phase ph(ws, 2);
                        wk.cks = wa.cks; wk.last_launch = (l == NB) ? 1u : 0u;
                        static const bool old_diag = [] { const char* e = getenv("BG_LONG_WALK"); return e && !strcmp(e, "diag"); }();
                        launch_long_walk(old_diag ? LW_DIAG : LW_SKEW, WAVE_C, ns, st, wk);
                    }
                    CU_TRY(ctx, cudaGetLastError());
                }
                ctx->launches += 1 + 2 * (uint64_t)NB;
                continue;
            }
            if (lc.wave && ch.fine_bands) {
                // ---- K2f: one column per lane; W warps per CTA so that the pair's warps cover as many SMs as they can ----
                static const int fine_w = [] { const char* e = getenv("BG_FINE_WARPS"); return e ? std::max(1, std::min(32, atoi(e))) : 0; }();
                const uint32_t B = ch.fine_bands;
                const int Wc = fine_w ? fine_w : (int)std::min<uint32_t>(32u, std::max<uint32_t>(8u, (B + (uint32_t)ctx->num_sms - 1) / (uint32_t)ctx->num_sms));
                const int n_cta = (int)((B + Wc - 1) / Wc);
                const size_t cand_bytes = (size_t)ns * B * sizeof(WaveCand);
                const size_t ctr_bytes = ((size_t)ns + 2 * (size_t)n_cta + 8) * 4;
                if (!ws.cand.ensure(cand_bytes) || !ws.progress.ensure(ctr_bytes) || !ws.assign.ensure((size_t)n_cta * FINE_RING_G * sizeof(int2))) {
                    ctx->set_error("device allocation failed (K2f scratch)"); return BG_ENOMEM;
                }
                CU_TRY(ctx, cudaMemsetAsync(ws.progress.p, 0, ctr_bytes, st));
                FineArgs fx; fx.f = fa; fx.cand = ws.cand.as<WaveCand>(); fx.cand_stride = B;
                fx.done = ws.progress.as<uint32_t>(); fx.cons_g = fx.done + ns + 2;
                fx.ring_g = ws.assign.as<unsigned long long>();
                Phase ph(ws, 1);
                CU_TRY(ctx, launch_k2f(pp.local, pp.prof4, n_cta, Wc, pp.smem, st, fx));
            } else if (lc.wave) {
                const uint64_t nw = (uint64_t)ch.max_Q * ch.wpc;
                const uint64_t prog_bytes = (uint64_t)ns * (nw + 1) * 8;
                CU_TRY(ctx, cudaMemsetAsync(ws.progress.p, 0, prog_bytes + (2 * (uint64_t)ns + 4) * 4, st));
                if (!ws.assign.ensure(std::max<size_t>(1, ch.assign.size()) * sizeof(WaveAssign))) { ctx->set_error("device allocation failed (K2 assignment)"); return BG_ENOMEM; }
                CU_TRY(ctx, cudaMemcpyAsync(ws.assign.p, ch.assign.data(), ch.assign.size() * sizeof(WaveAssign), cudaMemcpyHostToDevice, st));
                WaveArgs wa; wa.f = fa; wa.progress = ws.progress.as<unsigned long long>(); wa.cand = ws.cand.as<WaveCand>();
                wa.assign = ws.assign.as<WaveAssign>(); wa.n_rounds = ch.n_rounds;
                wa.prog_stride = (uint32_t)(nw + 1); wa.cand_stride = (uint32_t)nw;
                wa.done = reinterpret_cast<uint32_t*>(ws.progress.as<unsigned char>() + prog_bytes);
                wa.next_band = wa.done + ns + 2;
                Phase ph(ws, 1);
                CU_TRY(ctx, launch_k2(pp.local, pp.prof4, lc.sh.C, ctx->num_sms, (int)ch.wpc, pp.smem, st, wa));
            } else if (lc.half) {
                fst = (st2 && ((chunk_no + parity) & 1)) ? st2 : st;
                Phase ph(ws, 1, fst);
                const uint32_t nw2 = (ns + 2 * G - 1) / (2 * G);
                if (!dispatch_k1h(lc.sh, pp.mode != BG_GLOBAL, pp.half_prof8, dim3((nw2 + 3) / 4), fst, fa)) { ctx->set_error("internal: K1h shape not compiled"); return BG_ECUDA; }
            } else {
                fst = (st2 && ((chunk_no + parity) & 1)) ? st2 : st;
                Phase ph(ws, 1, fst);
                dispatch_k1(lc.sh, pp.local, pp.prof4, dim3((nwarps + 3) / 4), pp.smem, fst, fa);
            }
            CU_TRY(ctx, cudaGetLastError());
            if (overlap || split) {
                cudaEvent_t evf = ws.get_event();
                CU_TRY(ctx, cudaEventRecord(evf, fst));
                CU_TRY(ctx, cudaStreamWaitEvent(wst, evf, 0));
            }
            {
                cudaEvent_t pa = ws.get_event();
                cudaEventRecord(pa, wst);
                if (pp.score_only) {
                    launch_scores_only(fa.desc, fa.end, ns, io.score, io.flags, pp.mode, wst);
                } else {
                    WalkArgs wa;
                    wa.desc = fa.desc; wa.end = fa.end; wa.n_slots = ns; wa.residues = fa.residues;
                    wa.trace = fa.trace; wa.mode = pp.mode; wa.L = lc.sh.L; wa.C = lc.sh.C; wa.H = lc.half ? 2 : 1;
                    wa.tg_shift = fa.tg_shift;
                    wa.CW = lc.half ? (int32_t)hb_words_per_lane_block(lc.sh.C) : 0;
                    wa.pad = ws.pad.as<uint8_t>(); wa.score = io.score; wa.walk_flags = io.flags; wa.lens2 = io.lens2;
                    // long pairs: one warp per pair with a trace window in shared memory; short pairs: one thread per pair
                    if (lc.long_walk) {
                        static const bool old_diag = [] { const char* e = getenv("BG_LONG_WALK"); return e && !strcmp(e, "diag"); }();
                        launch_long_walk(old_diag ? LW_DIAG : LW_SKEW, (lc.sh.L == 32 && (lc.sh.C == WAVE_C || lc.sh.C == WAVE_C_NARROW)) ? lc.sh.C : 0, ns, wst, wa);
                    }
                    else dispatch_walk(lc.sh, lc.half, ns, wst, wa);
                }
                cudaEvent_t pb = ws.get_event();
                cudaEventRecord(pb, wst);
                ws.evs.push_back(PhaseEv{pa, pb, 2});
                if (overlap) ev_walk_done[buf] = pb;
            }
            CU_TRY(ctx, cudaGetLastError());
            ctx->launches += 3;
        }
    }
    if (overlap)
        for (int b2 = 0; b2 < 2; ++b2)
            if (ev_walk_done[b2]) CU_TRY(ctx, cudaStreamWaitEvent(st, ev_walk_done[b2], 0));   // scan / gather need every walk
    if (!pp.score_only && io.ops) {
        // counts[p] = {ceil(len / 16), len} -> exclusive scan (into io.off, [N + 1] x 16 B) -> the pairs' op runs, packed
        // densely; the counts live behind the scan's temporary storage in ws.cubtmp
        Phase ph(ws, 3, pst);
        ulonglong2* scan_out = reinterpret_cast<ulonglong2*>(io.off);
        size_t tmp = 0;
        scan_counts(nullptr, tmp, scan_out, scan_out, (int)(N + 1), pst);
        const size_t tmp_al = (tmp + 255) & ~(size_t)255;
        if (!ws.cubtmp.ensure(tmp_al + (N + 1) * 16)) { ctx->set_error("device allocation failed (scan)"); return BG_ENOMEM; }
        ulonglong2* counts = reinterpret_cast<ulonglong2*>(ws.cubtmp.as<unsigned char>() + tmp_al);
        launch_ops_counts(io.lens2, N, counts, pst);
        scan_counts(ws.cubtmp.p, tmp, counts, scan_out, (int)(N + 1), pst);
        ctx->launches += 2;
        if (P.n_slots) {
            PackOpsArgs pa;
            pa.desc = io.desc; pa.n_slots = (uint32_t)P.n_slots; pa.pad = ws.pad.as<uint8_t>(); pa.lens2 = io.lens2;
            pa.woff = scan_out; pa.len = io.len; pa.first = io.first; pa.ops = io.ops;
            launch_pack_ops(pa, (uint64_t)P.max_n + P.max_m > 4096, pst);
            ctx->launches++;
        }
        launch_ops_sample(scan_out, N, io.sample_stride, io.samples, pst);
        ctx->launches++;
        CU_TRY(ctx, cudaGetLastError());
    } else if (!pp.score_only) {
        Phase ph(ws, 3, pst);
        size_t tmp = 0;
        scan_lengths(nullptr, tmp, io.lens2, io.off, (int)(2 * N + 1), pst);
        if (!ws.cubtmp.ensure(tmp + 16)) { ctx->set_error("device allocation failed (scan)"); return BG_ENOMEM; }
        scan_lengths(ws.cubtmp.p, tmp, io.lens2, io.off, (int)(2 * N + 1), pst);
        ctx->launches += 2;
        if (P.n_slots) {
            GatherArgs ga;
            ga.desc = io.desc; ga.n_slots = (uint32_t)P.n_slots; ga.pad = ws.pad.as<uint8_t>();
            ga.off = io.off; ga.arena = io.arena; ga.residues = io.residues;
            launch_gather(ga, pst);
            ctx->launches++;
        }
        CU_TRY(ctx, cudaGetLastError());
    }
    return BG_OK;
}

// lut: host [256] byte -> code 0..3 / 0xFF, only needed when the plan has K4b classes.
int run_edit(bg_ctx* ctx, WorkSet& ws, const uint8_t* residues, const PairDesc* desc, const Plan& P, uint64_t* out,
             const uint8_t* lut) {
    if (!ws.bnd.ensure(std::max<uint64_t>(1, P.bnd_elems) * sizeof(int2)) || !ws.err.ensure(4) || !ws.codes.ensure(512)) {
        ctx->set_error("device allocation failed"); return BG_ENOMEM;
    }
    CU_TRY(ctx, cudaMemsetAsync(ws.err.p, 0, 4, ws.stream));
    if (P.myers && lut) CU_TRY(ctx, cudaMemcpyAsync(ws.codes.p, lut, 256, cudaMemcpyHostToDevice, ws.stream));
    EditArgs ea;
    ea.residues = residues; ea.bnd = ws.bnd.as<int32_t>(); ea.out = out;
    for (const LaunchClass& lc : P.classes) {
        const uint32_t G = 32 / lc.sh.L;
        for (const Chunk& ch : lc.chunks) {
            const uint32_t ns = ch.slot_end - ch.slot_begin;
            if (!ns) continue;
            Phase ph(ws, 1);
            if (lc.myers_W) {
                MyersArgs ma;
                ma.desc = desc + ch.slot_begin; ma.n_slots = ns; ma.residues = residues; ma.lut = ws.codes.as<uint8_t>();
                ma.cdesc = P.compact ? reinterpret_cast<const MyersSlot*>(desc) + ch.slot_begin : nullptr;
                ma.out = out; ma.err_flag = ws.err.as<uint32_t>(); ma.cls_count = nullptr; ma.cls = 0;
                launch_myers(lc.myers_W, ns, ws.stream, ma);
            } else {
                ea.desc = desc + ch.slot_begin; ea.n_slots = ns;
                const uint32_t nwarps = (ns + G - 1) / G;
                dispatch_k4(lc.sh, dim3((nwarps + 3) / 4), ws.stream, ea);
            }
            ctx->launches++;
        }
    }
    CU_TRY(ctx, cudaGetLastError());
    return BG_OK;
}

// <= 4 distinct byte values -> lut (byte -> 0..3, 0xFF elsewhere); false if the alphabet is richer
bool make_edit_lut(const uint64_t* hist, uint8_t* lut) {
    int k = 0;
    for (int x = 0; x < 256; ++x) lut[x] = 0xFF;
    for (int x = 0; x < 256; ++x)
        if (hist[x]) { if (k == 4) return false; lut[x] = (uint8_t)k++; }
    return true;
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
        case BG_EINVAL_FASTA: return "Expected > at record start.";
        default: return "unknown error";
    }
}

const char* bg_last_error(const bg_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

void bg_destroy(bg_ctx* ctx) {
    if (!ctx) return;
    for (auto& dv : ctx->devs) {
        cudaSetDevice(dv.ordinal);
        for (WorkSet& ws : dv.ws) {
            if (ws.stream) cudaStreamSynchronize(ws.stream);
            for (DevBuf* b : ws.all_bufs()) b->release();
            ws.stage.release(); ws.scalars.release(); ws.samples_h.release();
            for (auto e : ws.ev_pool) cudaEventDestroy(e);
            if (ws.ev_scan) cudaEventDestroy(ws.ev_scan);
            if (ws.walk_stream) cudaStreamDestroy(ws.walk_stream);
            for (cudaStream_t a : ws.aux) if (a) cudaStreamDestroy(a);
            if (ws.stream) cudaStreamDestroy(ws.stream);
        }
        if (dv.cache) { dv.cache->trim(); delete dv.cache; dv.cache = nullptr; }
    }
    delete ctx;
}

int bg_create(const int* devices, int n_dev, bg_ctx** out) {
    if (!out) return BG_EINVAL_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { (void)cudaGetLastError(); return BG_ENODEVICE; }
    bg_ctx* ctx = new bg_ctx();
    std::vector<int> ords;
    if (!devices || n_dev <= 0) { int cur = 0; cudaGetDevice(&cur); ords.push_back(cur); }
    else ords.assign(devices, devices + n_dev);
    ctx->devs.resize(ords.size());
    for (size_t i = 0; i < ords.size(); ++i) {
        const int o = ords[i];
        Device& dv = ctx->devs[i];
        if (o < 0 || o >= count || cudaSetDevice(o) != cudaSuccess) { bg_destroy(ctx); return BG_ENODEVICE; }
        dv.ordinal = o;
        size_t fr = 0, tot = 0; cudaMemGetInfo(&fr, &tot); dv.total_mem = tot;
        { int sms = 0; if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, o) == cudaSuccess && sms > 0) ctx->num_sms = sms; }
        dv.cache = new BlockCache();
        // the walk / post streams get the highest priority: their small, latency-bound grids must slip in
        // between the blocks of a concurrently running fill instead of queueing behind all of them
        int prio_least = 0, prio_greatest = 0;
        cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
        for (WorkSet& ws : dv.ws) {
            ws.ordinal = o; ws.cache = dv.cache;
            for (DevBuf* b : ws.all_bufs()) b->cache = dv.cache;
            if (cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking) != cudaSuccess ||
                cudaStreamCreateWithPriority(&ws.walk_stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
                cudaEventCreateWithFlags(&ws.ev_scan, cudaEventDisableTiming) != cudaSuccess) { bg_destroy(ctx); return BG_ECUDA; }
        }
    }
    uint64_t budget_mb = 8192;   // (device-resident cfg2: 1 / 2 / 4 / 8 GiB per launch -> 2215 / 2229 / 2229 / 2189 GCUPS: more launches hide more of the
                                 //  walks behind fills, +1.8 % overall, but the fills then share the SMs with walks for longer and the fill phase --
                                 //  what the roofline fraction is computed from -- stretches from 8.53 to 8.97 ms: kept at 8 GiB)
    if (const char* e = getenv("BG_TRACE_BUDGET_MB")) budget_mb = strtoull(e, nullptr, 10);
    ctx->trace_budget_words = budget_mb * (1024ull * 1024ull / 4ull);
    if (const char* e = getenv("BG_LONG_TRACE_BUDGET_MB")) ctx->long_budget_words = strtoull(e, nullptr, 10) * (1024ull * 1024ull / 4ull);
    if (const char* e = getenv("BG_HOST_PLAN")) ctx->host_plan = atoi(e) != 0;
    if (const char* e = getenv("BG_FINE_PAIRS")) ctx->fine_max_pairs = std::max(0, atoi(e));
    if (const char* e = getenv("BG_FORCE_SHAPE")) {   // "L,C" -- experiments / tests
        int l = 0, c = 0;
        if (sscanf(e, "%d,%d", &l, &c) == 2) { ctx->force_L = l; ctx->force_C = c; }
    }
    *out = ctx;
    return BG_OK;
}

void* bg_stream(bg_ctx* ctx, int dev_index) {
    if (!ctx || dev_index < 0 || dev_index >= (int)ctx->devs.size()) return nullptr;
    return (void*)ctx->devs[dev_index].ws[0].stream;
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

int bg_set_fine_pairs(bg_ctx* ctx, int max_pairs) {
    if (!ctx || max_pairs < 0) return BG_EINVAL_ARG;
    ctx->fine_max_pairs = max_pairs;
    return BG_OK;
}

int bg_set_host_plan(bg_ctx* ctx, int on) {
    if (!ctx) return BG_EINVAL_ARG;
    ctx->host_plan = on != 0;
    return BG_OK;
}

int bg_set_long_trace_budget(bg_ctx* ctx, uint64_t bytes) {   // 0 = automatic
    if (!ctx || (bytes && bytes < 4096)) return BG_EINVAL_ARG;
    ctx->long_budget_words = bytes / 4;
    return BG_OK;
}

int bg_sync(bg_ctx* ctx) {
    if (!ctx) return BG_EINVAL_ARG;
    for (auto& dv : ctx->devs) {
        cudaSetDevice(dv.ordinal);
        for (WorkSet& ws : dv.ws) { CU_TRY(ctx, cudaStreamSynchronize(ws.stream)); CU_TRY(ctx, cudaStreamSynchronize(ws.walk_stream)); }
    }
    return BG_OK;
}

int bg_last_timing(const bg_ctx* cctx, bg_timing* out) {
    bg_ctx* ctx = const_cast<bg_ctx*>(cctx);
    if (!ctx || !out) return BG_EINVAL_ARG;
    bg_timing t = ctx->timing;
    t.encode_ms = t.fill_ms = t.walk_ms = t.compact_ms = t.total_ms = 0;
    t.h2d_bytes = ctx->h2d; t.d2h_bytes = ctx->d2h; t.launches = ctx->launches; t.fill_launches = 0;
    for (auto& dv : ctx->devs) {
        cudaSetDevice(dv.ordinal);
        double ph[6] = {0, 0, 0, 0, 0, 0}, tot = 0;   // 0 encode, 1 fill, 2 walk, 3 compact, 4 H2D, 5 device-side plan
        for (WorkSet& ws : dv.ws) {
            CU_TRY(ctx, cudaStreamSynchronize(ws.stream));
            // A phase's time is the UNION of its kernels' intervals: launches of one phase can run side by side (fills
            // alternate between two streams, walks run next to the following fill), and a sum would count that time twice.
            std::vector<std::pair<double, double>> iv[6];
            double t_min = 0, t_max = 0; bool any = false;
            for (auto& ev : ws.evs) {
                float s0 = 0, dur = 0;
                cudaEventElapsedTime(&s0, ws.evs.front().a, ev.a); cudaEventElapsedTime(&dur, ev.a, ev.b);
                iv[ev.phase].emplace_back((double)s0, (double)s0 + dur);
                if (!any || s0 < t_min) t_min = s0;
                if (!any || s0 + dur > t_max) t_max = (double)s0 + dur;
                any = true;
                if (ev.phase == 1) t.fill_launches++;
            }
            for (int k = 0; k < 6; ++k) {
                std::sort(iv[k].begin(), iv[k].end());
                double cur_a = 0, cur_b = -1;
                for (auto& x : iv[k]) {
                    if (cur_b < cur_a || x.first > cur_b) { if (cur_b > cur_a) ph[k] += cur_b - cur_a; cur_a = x.first; cur_b = x.second; }
                    else cur_b = std::max(cur_b, x.second);
                }
                if (cur_b > cur_a) ph[k] += cur_b - cur_a;
            }
            if (any) tot = std::max<double>(tot, t_max - t_min);
        }
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
    WorkSet& ws = dv.ws[0];
    CU_TRY(ctx, cudaSetDevice(dv.ordinal));
    bg_dbatch* B = new bg_dbatch();
    B->ctx = ctx; B->dev_index = dev_index; B->n_pairs = in->n_pairs;
    for (DevBuf* b : {&B->residues, &B->desc_align, &B->desc_edit}) b->cache = dv.cache;
    const uint64_t base = in->n_pairs ? in->seq_off[0] : 0;
    B->seq_off.resize(2 * in->n_pairs + 1);
    for (uint64_t s = 0; s <= 2 * in->n_pairs; ++s) B->seq_off[s] = in->n_pairs ? in->seq_off[s] - base : 0;
    B->n_residues = B->seq_off.back();
    if (!B->residues.ensure(B->n_residues + 16)) { delete B; ctx->set_error("device allocation for residues failed"); return BG_ENOMEM; }
    ctx->h2d = 0;
    if (B->n_residues) {
        uint64_t b0, b1, moved = 0;
        host_byte_range(in->packing, base, base + B->n_residues, b0, b1);
        rc = upload_residues(ctx, in->residues + b0, in->packing, in->alphabet, base, base + B->n_residues, ws.packed, B->residues.as<uint8_t>(), ws.stream, &moved);
        if (rc) { B->residues.release(); delete B; return rc; }
        ctx->h2d = moved;
    }
    *out = B;
    return BG_OK;
}

void bg_dbatch_free(bg_dbatch* b) {
    if (!b) return;
    Device& dv = b->ctx->devs[b->dev_index];
    cudaSetDevice(dv.ordinal);
    cudaStreamSynchronize(dv.ws[0].stream);
    b->residues.release(); b->desc_align.release(); b->desc_edit.release();
    delete b;
}

// Does not synchronise: the blocks go back to the device cache and are only ever handed out again to
// work that is enqueued behind the current work on work set 0's stream (device-resident calls) or after
// a full sync (the host-buffer entry points sync all streams first), so stream order protects them.
void bg_dresult_free(bg_dresult* r) {
    if (!r) return;
    for (DevBuf* b : {&r->score, &r->flags, &r->lens2, &r->off, &r->arena, &r->out64}) b->release();
    delete r;
}

static int ensure_plan(bg_ctx* ctx, bg_dbatch* B, bool edit, int32_t half_maxabs = 0) {
    Device& dv = ctx->devs[B->dev_index];
    WorkSet& ws = dv.ws[0];
    Plan& P = edit ? B->plan_edit : B->plan_align;
    DevBuf& D = edit ? B->desc_edit : B->desc_align;
    if (P.built && (edit || P.half_maxabs == half_maxabs)) return BG_OK;
    // descriptors are built straight into cached pinned memory and copied once
    const size_t cap = plan_desc_capacity(B->n_pairs);
    PinBuf stage;
    if (!stage.ensure(cap * sizeof(PairDesc))) { ctx->set_error("pinned staging allocation failed"); return BG_ENOMEM; }
    // long pairs (K2) need whole traces of several GB each: let them use most of the device
    const uint64_t wave_budget = ctx->long_budget_words ? ctx->long_budget_words : std::max<uint64_t>(ctx->trace_budget_words, (uint64_t)(0.8 * (double)dv.total_mem) / 4);
    // with BG_OVERLAP: chunks of <= 2 GiB of trace so that walk(c) overlaps fill(c+1) (run_align); otherwise large
    // chunks -- the walk is latency-bound and wants as many pairs per launch as possible
    static const bool want_overlap = [] { const char* e = getenv("BG_OVERLAP"); return e && atoi(e) != 0; }();
    const uint64_t chunk_budget = want_overlap ? std::min<uint64_t>(ctx->trace_budget_words, (2ull << 30) / 4) : ctx->trace_budget_words;
    bool myers = false;
    if (edit && B->n_residues && !getenv("BG_NO_MYERS")) {
        // which byte values occur at all?  (<= 4 -> bit-parallel K4b)
        if (!ws.cubtmp.ensure(1024) || !ws.scalars.ensure(1024)) { stage.release(); ctx->set_error("allocation failed"); return BG_ENOMEM; }
        CU_TRY(ctx, cudaMemsetAsync(ws.cubtmp.p, 0, 1024, ws.stream));
        launch_byte_hist(B->residues.as<uint8_t>(), B->n_residues, ws.cubtmp.as<unsigned int>(), ctx->num_sms * 4, ws.stream);
        CU_TRY(ctx, cudaMemcpyAsync(ws.scalars.p, ws.cubtmp.p, 1024, cudaMemcpyDeviceToHost, ws.stream));
        CU_TRY(ctx, cudaStreamSynchronize(ws.stream));
        uint64_t hist[256];
        for (int x = 0; x < 256; ++x) hist[x] = ws.scalars.as<unsigned int>()[x];
        myers = B->edit_lut_ok = make_edit_lut(hist, B->edit_lut);
    }
    int rc = build_plan(ctx, B->seq_off.data(), 0, B->n_pairs, !edit, chunk_budget, wave_budget, edit ? 0 : half_maxabs, P, stage.as<PairDesc>(), myers);
    if (rc) { stage.release(); return rc; }
    if (P.n_slots) {
        if (!D.ensure(P.n_slots * sizeof(PairDesc))) { stage.release(); ctx->set_error("device allocation for descriptors failed"); return BG_ENOMEM; }
        cudaError_t ce = cudaMemcpyAsync(D.p, stage.p, P.n_slots * sizeof(PairDesc), cudaMemcpyHostToDevice, ws.stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(ws.stream);
        stage.release();
        CU_TRY(ctx, ce);
        ctx->h2d += P.n_slots * sizeof(PairDesc);
    } else stage.release();
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
    const uint64_t N = B->n_pairs;
    Prepared pp;
    int rc = prepare_params(ctx, p, B->seq_off.data(), N, pp);
    if (rc) return rc;
    Device& dv = ctx->devs[B->dev_index];
    WorkSet& ws = dv.ws[0];
    CU_TRY(ctx, cudaSetDevice(dv.ordinal));
    rc = ensure_plan(ctx, B, false, pp.half_maxabs);
    if (rc) return rc;
    Plan& P = B->plan_align;

    bg_dresult* R = new bg_dresult();
    R->ctx = ctx; R->dev_index = B->dev_index; R->n_pairs = N; R->kind = 0; R->mode = pp.mode; R->score_only = pp.score_only;
    for (DevBuf* b : {&R->score, &R->flags, &R->lens2, &R->off, &R->arena, &R->out64}) b->cache = dv.cache;
    bool ok = R->score.ensure(std::max<uint64_t>(1, N) * 4) && R->flags.ensure(std::max<uint64_t>(1, N));
    if (!pp.score_only)
        ok = ok && R->lens2.ensure((2 * N + 1) * 8) && R->off.ensure((2 * N + 1) * 8) && R->arena.ensure(std::max<uint64_t>(1, P.pad_bytes));
    if (!ok) { bg_dresult_free(R); ctx->set_error("device allocation failed (output buffers)"); return BG_ENOMEM; }

    ws.reset_events();
    ctx->launches = 0;
    ctx->timing.cells = P.cells; ctx->timing.cells_packed16 = P.cells_half; ctx->timing.cells_bitparallel = 0; ctx->timing.cells_refilled = P.cells_ckpt;
    ctx->timing.trace_bytes = pp.score_only ? 0 : P.total_trace_words * 4;
    rc = upload_params(ctx, ws, pp);
    AlignIO io{B->residues.as<uint8_t>(), B->desc_align.as<PairDesc>(), &P, N,
               R->score.as<int32_t>(), R->flags.as<uint8_t>(), R->lens2.as<uint64_t>(), R->off.as<uint64_t>(), R->arena.as<uint8_t>()};
    // Device-resident batches: when the traces of ALL launches fit next to each other (45 % of the device), every
    // launch keeps its own trace region, so the walk / gather of launch c run on the high-priority stream next to
    // the fill of launch c + 1 (the walk is a latency-bound chain of dependent loads that leaves the integer pipes
    // idle) and the fills alternate between two streams.  The main stream then waits for the post stream, so that
    // everything ordered after this call on the main stream -- the next call's fills included -- sees the results.
    static const bool no_dev_split = getenv("BG_NO_DEV_SPLIT") != nullptr;
    const uint64_t cap_saved = ws.split_cap_words;
    const bool dev_split = !no_dev_split && !pp.score_only && P.max_wave_slots == 0 &&
                           (double)P.total_trace_words * 4.0 <= 0.45 * (double)dv.total_mem;
    if (dev_split) {
        ws.post_stream = ws.walk_stream; ws.fill2_stream = dv.ws[1].stream;
        ws.split_cap_words = P.total_trace_words;
    }
    if (!rc) rc = run_align(ctx, ws, io, pp);
    if (dev_split) {
        cudaEvent_t evj = ws.get_event();
        if (cudaEventRecord(evj, ws.post_stream) != cudaSuccess || cudaStreamWaitEvent(ws.stream, evj, 0) != cudaSuccess) { if (!rc) { ctx->set_error("stream join failed"); rc = BG_ECUDA; } }
        ws.post_stream = nullptr; ws.fill2_stream = nullptr; ws.split_cap_words = cap_saved;
    }
    if (rc) { bg_dresult_free(R); return rc; }
    *out = R;
    return BG_OK;
}

int bg_edit_distance_device(bg_ctx* ctx, const bg_dbatch* cin, bg_dresult** out) {
    if (!ctx || !cin || !out) return BG_EINVAL_ARG;
    *out = nullptr;
    bg_dbatch* B = const_cast<bg_dbatch*>(cin);
    Device& dv = ctx->devs[B->dev_index];
    WorkSet& ws = dv.ws[0];
    CU_TRY(ctx, cudaSetDevice(dv.ordinal));
    int rc = ensure_plan(ctx, B, true);
    if (rc) return rc;
    Plan& P = B->plan_edit;
    const uint64_t N = B->n_pairs;
    bg_dresult* R = new bg_dresult();
    R->ctx = ctx; R->dev_index = B->dev_index; R->n_pairs = N; R->kind = 1;
    for (DevBuf* b : {&R->score, &R->flags, &R->lens2, &R->off, &R->arena, &R->out64}) b->cache = dv.cache;
    if (!R->out64.ensure(std::max<uint64_t>(1, N) * 8)) { bg_dresult_free(R); ctx->set_error("device allocation failed"); return BG_ENOMEM; }
    ws.reset_events();
    ctx->launches = 0;
    ctx->timing.cells = P.cells; ctx->timing.trace_bytes = 0; ctx->timing.cells_packed16 = 0; ctx->timing.cells_bitparallel = P.cells_myers; ctx->timing.cells_refilled = 0;
    rc = run_edit(ctx, ws, B->residues.as<uint8_t>(), B->desc_edit.as<PairDesc>(), P, R->out64.as<uint64_t>(), B->edit_lut_ok ? B->edit_lut : nullptr);
    if (rc) { bg_dresult_free(R); return rc; }
    *out = R;
    return BG_OK;
}

int bg_ref_status(int mode, uint64_t n, uint64_t m, int32_t score, int walk_flags) {
    return ref_status(mode, n, m, score, (uint32_t)walk_flags) ? BG_ST_REF_UNDEFINED : BG_ST_OK;
}

int bg_dresult_download(bg_ctx* ctx, bg_dresult* r, bg_result* out) {
    if (!ctx || !r || !out || r->kind != 0) return BG_EINVAL_ARG;
    memset(out, 0, sizeof *out);
    Device& dv = ctx->devs[r->dev_index];
    WorkSet& ws = dv.ws[0];
    CU_TRY(ctx, cudaSetDevice(dv.ordinal));
    const uint64_t N = r->n_pairs;
    if (!ws.scalars.ensure(16)) { ctx->set_error("pinned allocation failed"); return BG_ENOMEM; }
    uint64_t* h_total = ws.scalars.as<uint64_t>();
    uint32_t* h_err = reinterpret_cast<uint32_t*>(h_total + 1);
    *h_total = 0; *h_err = 0;
    CU_TRY(ctx, cudaMemcpyAsync(h_err, ws.err.p, 4, cudaMemcpyDeviceToHost, ws.stream));
    if (!r->score_only)
        CU_TRY(ctx, cudaMemcpyAsync(h_total, r->off.as<uint64_t>() + 2 * N, 8, cudaMemcpyDeviceToHost, ws.stream));
    CU_TRY(ctx, cudaStreamSynchronize(ws.stream));
    if (*h_err & 1u) { ctx->set_error("a residue byte has no row/column in the score table"); return BG_EINVAL_RESIDUE; }
    const uint64_t total = *h_total;

    HostResultOwner* own = new HostResultOwner();
    out->n_pairs = N;
    out->score = (int32_t*)own->grab(N * 4);
    out->status = (uint8_t*)own->grab(N);
    out->off = (uint64_t*)own->grab((2 * N + 1) * 8);
    out->arena = (uint8_t*)own->grab(total);
    out->owner_ = own;
    if (!out->score || !out->status || !out->off || !out->arena) { bg_result_free(out); ctx->set_error("pinned host allocation failed"); return BG_ENOMEM; }
    if (N) {
        CU_TRY(ctx, cudaMemcpyAsync(out->score, r->score.p, N * 4, cudaMemcpyDeviceToHost, ws.stream));
        CU_TRY(ctx, cudaMemcpyAsync(out->status, r->flags.p, N, cudaMemcpyDeviceToHost, ws.stream));
    }
    if (!r->score_only) {
        CU_TRY(ctx, cudaMemcpyAsync(out->off, r->off.p, (2 * N + 1) * 8, cudaMemcpyDeviceToHost, ws.stream));
        if (total) CU_TRY(ctx, cudaMemcpyAsync(out->arena, r->arena.p, total, cudaMemcpyDeviceToHost, ws.stream));
    } else {
        memset(out->off, 0, (2 * N + 1) * 8);
    }
    CU_TRY(ctx, cudaStreamSynchronize(ws.stream));
    ctx->d2h = N * 5 + (r->score_only ? 0 : (2 * N + 1) * 8 + total);
    return BG_OK;   // status: evaluated by the walk kernels (bg_common.cuh ref_status)
}

int bg_dresult_download_u64(bg_ctx* ctx, bg_dresult* r, uint64_t* out) {
    if (!ctx || !r || r->kind != 1 || (!out && r->n_pairs)) return BG_EINVAL_ARG;
    Device& dv = ctx->devs[r->dev_index];
    WorkSet& ws = dv.ws[0];
    CU_TRY(ctx, cudaSetDevice(dv.ordinal));
    if (r->n_pairs) CU_TRY(ctx, cudaMemcpyAsync(out, r->out64.p, r->n_pairs * 8, cudaMemcpyDeviceToHost, ws.stream));
    CU_TRY(ctx, cudaStreamSynchronize(ws.stream));
    ctx->d2h = r->n_pairs * 8;
    return BG_OK;
}

void bg_result_free(bg_result* r) {
    if (!r) return;
    if (r->owner_) {
        HostResultOwner* own = (HostResultOwner*)r->owner_;
        own->release_all();
        delete own;
    }
    memset(r, 0, sizeof *r);
}

}  // extern "C"

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

// contiguous shards with ~equal cell counts at SCAN_BLOCK granularity
std::vector<uint64_t> shard_bounds_from_scan(const BatchScan& S, uint64_t N, int nd) {
    std::vector<uint64_t> b(nd + 1, 0);
    b[nd] = N;
    if (nd == 1) return b;
    double total = 0;
    for (double c : S.block_cost) total += c;
    double acc = 0; int d = 1;
    for (size_t k = 0; k < S.block_cost.size() && d < nd; ++k) {
        acc += S.block_cost[k];
        while (d < nd && acc >= total * d / nd) { b[d] = std::min<uint64_t>(N, (k + 1) * SCAN_BLOCK); ++d; }
    }
    for (; d < nd; ++d) b[d] = N;
    for (int q = 1; q <= nd; ++q) if (b[q] < b[q - 1]) b[q] = b[q - 1];
    return b;
}

// ---- host-buffer alignment -------------------------------------------------------------------------------
// Results leave the device in compact form -- score, status, aligned length, start cell and 2-bit ops per pair
// (k_pack_ops) -- and the strings are rebuilt on the host from the caller's own residues (host_expand.cpp).  D2H
// is the scarce direction of the 8-GPU box (profiles/pcie_roof_r02.json), and the host pass also writes every
// string at its final, dense position whatever device or chunk it came from, so there is nothing to stitch.
struct OpsOut {        // compact result arrays of the whole batch (pinned, caller order)
    int32_t* score = nullptr; uint8_t* status = nullptr; uint32_t* len = nullptr; uint32_t* first = nullptr;
    uint32_t* ops = nullptr; uint64_t* ops_off = nullptr; uint64_t ops_cap_words = 0;
};

// One unit of device work: pairs [lo, lo + n) of the caller's batch (pipeline mode), or a set of pairs gathered
// into staging (long pairs dealt to devices by size; map[q] = caller index of local pair q).
struct WorkItem {
    uint64_t lo = 0, n = 0;
    const uint8_t* res = nullptr;       // host residues of the item: byte b0 of host_byte_range(packing, off[0], off[2n]) onwards
    uint32_t packing = BG_PACK_NONE;    // as in bg_batch (gathered long-mode items are unpacked while they are gathered)
    const uint8_t* alphabet = nullptr;
    const uint64_t* off = nullptr;      // [2n + 1]
    const uint32_t* map = nullptr;
    uint64_t ops_base = 0, ops_cap = 0; // the item's region of OpsOut::ops (upper-bound layout), words
    uint64_t words = 0;                 // op words the item really produced
    uint64_t cols = 0;                  // sum of aligned lengths
    uint64_t stride = 4096;             // pairs per expansion task; the device samples its offset scans at this stride
    std::vector<ulonglong2> samples;    // [nsub + 1]: {op words, columns} before sub-block i of the item; [nsub] = totals
    PinBuf gather;                      // long mode: gathered residues | offsets | map | per-pair outputs
    int32_t* t_score = nullptr; uint8_t* t_status = nullptr; uint32_t* t_len = nullptr; uint32_t* t_first = nullptr;   // long mode
};

struct Prebuilt { Plan plan; PinBuf stage, res; int rc = BG_OK; TaskHandle th; bool dev_plan = false, need_sort = false; PlanArgs pa; };

struct AlignJob {
    std::vector<Prebuilt> pre;                  // per item
    bg_ctx* ctx = nullptr; const bg_batch* in = nullptr; const Prepared* pp = nullptr; const BatchScan* scan = nullptr;
    OpsOut oo;
    std::vector<WorkItem> items;
    bool long_mode = false;
    std::vector<std::vector<int>> dev_items;   // long mode: items per device; pipeline mode: empty (shared queue)
    std::atomic<int> next_item{0};
    std::atomic<int> rc{BG_OK};
    // string expansion (bg_align_batch only): items are expanded in caller order as soon as every earlier item's
    // column total is known
    bool want_strings = false;
    uint64_t* off = nullptr; uint8_t* arena = nullptr; uint64_t arena_cap = 0;
    std::mutex mu; std::condition_variable cv;
    std::vector<char> arrived; int frontier = 0; uint64_t arena_base = 0;
    int pending = 0;                            // expansion tasks in flight
    uint32_t lut2[256];                         // 2-bit packed input: packed byte -> four residue bytes
    void fail(int code) { int expect = BG_OK; rc.compare_exchange_strong(expect, code); }
};

// The residues the expansion of pair p reads: seq1 from first_a on, seq2 from first_b on -- the caller's bytes, or
// (packed batches) unpacked into a buffer of the calling thread.
struct PairResidues {
    const uint8_t* s1; const uint8_t* s2;
    PairResidues(const bg_batch* in, const uint32_t* lut2, uint64_t p, uint32_t first_a, uint32_t first_b) {
        const uint64_t o0 = in->seq_off[2 * p], o1 = in->seq_off[2 * p + 1], o2 = in->seq_off[2 * p + 2];
        if (in->packing == BG_PACK_NONE) { s1 = in->residues + o0 + first_a; s2 = in->residues + o1 + first_b; return; }
        static thread_local std::vector<uint8_t> buf;
        const uint64_t n1 = o1 - o0 - first_a, n2 = o2 - o1 - first_b;
        if (buf.size() < n1 + n2 + 128) buf.resize(n1 + n2 + 128);
        unpack_residues(in->residues, in->packing, in->alphabet, o0 + first_a, n1, buf.data(), lut2);
        unpack_residues(in->residues, in->packing, in->alphabet, o1 + first_b, n2, buf.data() + n1 + 64, lut2);
        s1 = buf.data(); s2 = buf.data() + n1 + 64;
    }
};

// Expands pairs [p_lo, p_hi) (caller order): off[2p] holds the pair's a_align offset relative to `base`.
void expand_pairs(AlignJob& J, uint64_t p_lo, uint64_t p_hi, uint64_t base) {
    const bg_batch* in = J.in;
    for (uint64_t p = p_lo; p < p_hi; ++p) {
        const uint64_t len = J.oo.len[p], o = J.off[2 * p] + base;
        J.off[2 * p] = o; J.off[2 * p + 1] = o + len;
        if (!len) continue;
        if (o + 2 * len > J.arena_cap) { J.ctx->set_error("internal: arena bound exceeded"); J.fail(BG_ECUDA); return; }
        const PairResidues pr(in, J.lut2, p, J.oo.first[2 * p], J.oo.first[2 * p + 1]);
        expand_ops(pr.s1, pr.s2, J.oo.ops + J.oo.ops_off[p], len, J.arena + o, J.arena + o + len);
    }
}

// Hands pairs [p_lo, p_hi) to the host pool in pieces of similar column counts.  Caller holds J.mu.
void submit_expand_locked(AlignJob& J, uint64_t p_lo, uint64_t p_hi, uint64_t base, uint64_t cols) {
    const uint64_t per_task = std::max<uint64_t>(1ull << 20, cols / 64 + 1);
    uint64_t start = p_lo, acc = 0;
    for (uint64_t p = p_lo; p < p_hi; ++p) {
        acc += J.oo.len[p];
        if (acc >= per_task || p + 1 == p_hi) {
            const uint64_t a = start, b = p + 1;
            ++J.pending;
            host_pool().submit([&J, a, b, base] {
                expand_pairs(J, a, b, base);
                { std::lock_guard<std::mutex> lk(J.mu); --J.pending; }
                J.cv.notify_all();
            });
            start = p + 1; acc = 0;
        }
    }
}

// Sub-block i of a pipeline item (pairs [lo + i * stride, ...)): op offsets of its pairs and, when strings are wanted,
// their final arena offsets + the strings themselves.  The sub-block's starting offsets come from the device's scans
// (WorkItem::samples), so sub-blocks are independent of each other and no serial pass over the item exists.
void expand_sub(AlignJob& J, int c, uint64_t i, uint64_t arena_base, bool strings) {
    const WorkItem& it = J.items[c];
    const bg_batch* in = J.in;
    OpsOut& oo = J.oo;
    const uint64_t p_lo = it.lo + i * it.stride, p_hi = std::min(it.lo + it.n, p_lo + it.stride);
    const uint64_t nsub = (it.n + it.stride - 1) / it.stride;
    const uint64_t o_begin = arena_base + 2 * it.samples[i].y;
    const uint64_t o_end = arena_base + 2 * (i + 1 < nsub ? it.samples[i + 1].y : it.cols);
    if (strings && o_end > J.arena_cap) { J.ctx->set_error("internal: arena bound exceeded"); J.fail(BG_ECUDA); return; }
    // the sub-block's strings are one contiguous run of the arena: they are built in a cache-resident buffer of this
    // thread and leave it with non-temporal stores (stream_copy)
    static thread_local std::vector<uint8_t> stage;
    uint8_t* out = nullptr;
    if (strings) { if (stage.size() < o_end - o_begin + 64) stage.resize(o_end - o_begin + 64); out = stage.data(); }
    uint64_t w = it.ops_base + it.samples[i].x, o = o_begin;
    for (uint64_t p = p_lo; p < p_hi; ++p) {
        const uint64_t len = oo.len[p];
        oo.ops_off[p] = w;
        if (strings) {
            J.off[2 * p] = o; J.off[2 * p + 1] = o + len;
            if (len) {
                const PairResidues pr(in, J.lut2, p, oo.first[2 * p], oo.first[2 * p + 1]);
                expand_ops(pr.s1, pr.s2, oo.ops + w, len, out + (o - o_begin), out + (o - o_begin) + len);
            }
        }
        w += (len + 15) >> 4; o += 2 * len;
    }
    if (strings) {
        if (o != o_end) { J.ctx->set_error("internal: sampled column offsets disagree with the lengths"); J.fail(BG_ECUDA); return; }
        stream_copy(J.arena + o_begin, out, o_end - o_begin);
    }
}

// Caller holds J.mu.
void submit_item_locked(AlignJob& J, int c, uint64_t arena_base, bool strings) {
    const WorkItem& it = J.items[c];
    const uint64_t nsub = (it.n + it.stride - 1) / it.stride;
    for (uint64_t i = 0; i < nsub; ++i) {
        ++J.pending;
        host_pool().submit([&J, c, i, arena_base, strings] {
            expand_sub(J, c, i, arena_base, strings);
            { std::lock_guard<std::mutex> lk(J.mu); --J.pending; }
            J.cv.notify_all();
        });
    }
}

// The item's results are on the host.  Pipeline items: hand their sub-blocks to the host pool -- at once when only the
// op offsets are wanted, in caller order when strings are (an item's arena base is the column total of everything
// before it).  Long-mode items hold scattered pairs: scatter their per-pair outputs to caller order here; the strings
// are laid out after the last item (run_align_job).
void finish_item(AlignJob& J, int c) {
    WorkItem& it = J.items[c];
    OpsOut& oo = J.oo;
    if (it.map) {
        uint64_t w = it.ops_base;
        for (uint64_t q = 0; q < it.n; ++q) {
            const uint64_t p = it.map[q];
            oo.score[p] = it.t_score[q]; oo.status[p] = it.t_status[q];
            if (oo.len) {
                oo.len[p] = it.t_len[q]; oo.first[2 * p] = it.t_first[2 * q]; oo.first[2 * p + 1] = it.t_first[2 * q + 1];
                oo.ops_off[p] = w; w += ((uint64_t)it.t_len[q] + 15) >> 4;
            }
        }
        if (oo.len && w - it.ops_base != it.words) { J.ctx->set_error("internal: op word count mismatch"); J.fail(BG_ECUDA); }
        return;
    }
    if (!oo.len) return;                               // score-only
    std::lock_guard<std::mutex> lk(J.mu);
    if (J.long_mode) {                                  // one device, one item, caller order: op offsets now, strings at the end
        if (J.rc.load() == BG_OK) submit_item_locked(J, c, 0, false);
        return;
    }
    if (!J.want_strings) { if (J.rc.load() == BG_OK) submit_item_locked(J, c, 0, false); return; }
    J.arrived[c] = 1;
    while (J.frontier < (int)J.items.size() && J.arrived[J.frontier]) {
        if (J.rc.load() == BG_OK) submit_item_locked(J, J.frontier, J.arena_base, true);
        J.arena_base += 2 * J.items[J.frontier].cols;
        ++J.frontier;
    }
}

// One device's share of a job: its items flow through the work sets as a pipeline on dedicated streams --
//     H2D stream:      residues + launch descriptors of item c
//     compute stream:  all kernels, strictly item after item (each item gets the whole GPU; with one stream per
//                      item the kernels of three items time-share the SMs, all three finish together and the copy
//                      engines then sit idle: measured 17 ms instead of 11 for cfg2)
//     D2H streams:     per-pair results, then the packed ops once their size is known
// driven by two host threads: the issuer takes the next item (pipeline mode: from the queue all devices share, so a
// faster or less loaded device simply takes more chunks) as a work set becomes free, a finisher waits for each
// item's results, issues the ops copy and hands the item to finish_item.
int device_pipeline(AlignJob& J, int d) {
    bg_ctx* ctx = J.ctx;
    const Prepared& pp = *J.pp;
    Device& dv = ctx->devs[d];
    if (cudaSetDevice(dv.ordinal) != cudaSuccess) { ctx->set_error("cudaSetDevice failed"); return BG_ECUDA; }
    const bool long_mode = J.long_mode;
    const bool ops_mode = J.oo.len != nullptr;

    std::vector<Prebuilt>& pre = J.pre;   // launch plans, built by the host pool from the start of the job on
    const int nitems = (int)J.items.size();

    // streams: the work sets' own streams are borrowed for the stages; kernels of every work set go to st_comp
    cudaStream_t st_comp = dv.ws[0].stream, st_h2d = dv.ws[1].stream, st_small = dv.ws[2].stream, st_arena = dv.ws[0].walk_stream;
    cudaStream_t st_plan = dv.ws[2].walk_stream;     // high priority, like the post stream
    static const bool no_split = getenv("BG_NO_SPLIT") != nullptr;
    cudaStream_t st_post = (long_mode || no_split) ? st_comp : dv.ws[1].walk_stream;
    cudaStream_t saved[PIPE_DEPTH];
    // every launch of an item gets its own trace region (so that its walk runs next to the next fill) as long as the items
    // in flight together stay within 45 % of the device: cfg4's chunks need ~12 GB each (six length classes, 0.5 MB per
    // 1000 x 1000 pair); with the old fixed 6 GB cap they ran fill -> walk -> fill on one stream, 52 instead of 43 ms per call
    const uint64_t cap_saved = dv.ws[0].split_cap_words;
    const uint64_t split_cap = (uint64_t)(0.45 * (double)dv.total_mem) / 4 / (uint64_t)std::max(1, std::min(PIPE_DEPTH, nitems));
    for (int s = 0; s < PIPE_DEPTH; ++s) {
        saved[s] = dv.ws[s].stream; dv.ws[s].stream = st_comp; dv.ws[s].reset_events();
        dv.ws[s].post_stream = (st_post == st_comp) ? nullptr : st_post;
        dv.ws[s].split_cap_words = std::max(cap_saved, split_cap);
    }
    static const bool no_fill2 = getenv("BG_NO_FILL2") != nullptr;
    cudaStream_t st_fill2 = (PIPE_DEPTH > 3 && !no_fill2 && st_post != st_comp) ? saved[3] : nullptr;   // an otherwise idle work-set stream
    for (int s = 0; s < PIPE_DEPTH; ++s) dv.ws[s].fill2_stream = st_fill2;
    cudaEvent_t ev_h2d[PIPE_DEPTH], ev_comp[PIPE_DEPTH], ev_arena[PIPE_DEPTH], ev_plan[PIPE_DEPTH];
    for (int s = 0; s < PIPE_DEPTH; ++s) {
        cudaEventCreateWithFlags(&ev_h2d[s], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_comp[s], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_arena[s], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_plan[s], cudaEventDisableTiming);
    }
    static const bool prof = getenv("BG_PROFILE_HOST") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };

    // issuer -> finisher hand-over (k-th item this device took -> work set k % PIPE_DEPTH)
    std::mutex mu; std::condition_variable cv;
    std::vector<int> taken;               // item index of the k-th item issued here
    int issued = 0, finished = 0; bool issuer_done = false;

    Worker& finisher = ctx->fin_worker(d);
    finisher.run([&] {
        cudaSetDevice(dv.ordinal);
        for (int k = 0;; ++k) {
            int c;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return issued > k || issuer_done; });
                if (issued <= k) break;
                c = taken[k];
            }
            const int s = k % PIPE_DEPTH;
            WorkSet& ws = dv.ws[s];
            WorkItem& it = J.items[c];
            uint64_t* h_total = ws.scalars.as<uint64_t>();
            uint32_t* h_err = reinterpret_cast<uint32_t*>(h_total + 1);
            int rc = BG_OK;
            if (cudaEventSynchronize(ws.ev_scan) != cudaSuccess) { ctx->set_error("cudaEventSynchronize failed"); rc = BG_ECUDA; }
            if (!rc && (*h_err & 1u)) { ctx->set_error("a residue byte has no row/column in the score table"); rc = BG_EINVAL_RESIDUE; }
            if (!rc && ops_mode) {
                const uint64_t nsub = (it.n + it.stride - 1) / it.stride;
                const ulonglong2* sh = ws.samples_h.as<ulonglong2>();
                it.samples.assign(sh, sh + nsub + 1);
                it.words = sh[nsub].x; it.cols = sh[nsub].y;
                if (it.words > it.ops_cap) { ctx->set_error("internal: op arena bound exceeded"); rc = BG_ECUDA; }
                else if (it.words &&
                         (cudaMemcpyAsync(J.oo.ops + it.ops_base, ws.ops.p, it.words * 4, cudaMemcpyDeviceToHost, st_arena) != cudaSuccess ||
                          cudaEventRecord(ev_arena[s], st_arena) != cudaSuccess || cudaEventSynchronize(ev_arena[s]) != cudaSuccess)) {
                    ctx->set_error("ops copy failed"); rc = BG_ECUDA;
                }
            }
            if (prof) fprintf(stderr, "[bgalign]   dev %d item %d results on the host at %.2f ms\n", d, c, since());
            if (!rc) {
                ctx->d2h += it.n * 5 + (ops_mode ? it.n * 12 + it.words * 4 + (it.samples.size()) * 16 : 0);
                finish_item(J, c);
                if (prof) fprintf(stderr, "[bgalign]   dev %d item %d finished at %.2f ms\n", d, c, since());
            } else {
                J.fail(rc);
            }
            { std::lock_guard<std::mutex> lk(mu); finished = k + 1; }
            cv.notify_all();
        }
    });

    static const bool no_alt = getenv("BG_NO_ALT_FILL") != nullptr;
    auto issue = [&](int s, int c, int k) -> int {
        WorkSet& ws = dv.ws[s];
        ws.launch_parity = no_alt ? -1 : ((k + 1) & 1);   // (chunk_no is 1 at the first launch of a plan: item 0 starts on the main stream)
        WorkItem& it = J.items[c];
        const uint64_t n = it.n;
        const uint64_t base = it.off[0], nres = it.off[2 * n] - base;
        if (!ws.scalars.ensure(16)) { ctx->set_error("pinned staging allocation failed"); return BG_ENOMEM; }
        Prebuilt& pb = pre[c];
        if (pb.th.joinable()) pb.th.join();
        int rc = pb.rc;
        if (rc) return rc;
        const Plan& P = pb.plan;
        bool ok = ws.residues.ensure(nres + 16) && ws.desc.ensure(std::max<size_t>(1, P.n_slots) * sizeof(PairDesc)) &&
                  ws.score.ensure(std::max<uint64_t>(1, n) * 4) && ws.flags.ensure(std::max<uint64_t>(1, n));
        if (ops_mode)
            ok = ok && ws.lens2.ensure((2 * n + 1) * 8) && ws.off.ensure((n + 1) * 16) && ws.len.ensure(std::max<uint64_t>(1, n) * 4) &&
                 ws.first.ensure(std::max<uint64_t>(1, n) * 8) && ws.ops.ensure(std::max<uint64_t>(1, it.ops_cap) * 4);
        if (!ok) { ctx->set_error("device allocation failed (pipeline buffers)"); return BG_ENOMEM; }
        const uint8_t* src = pb.res.p ? (const uint8_t*)pb.res.p : it.res;
        cudaEvent_t pe_a = nullptr;
        if (prof) { pe_a = ws.get_event(); cudaEventRecord(pe_a, st_h2d); }
        uint64_t res_bytes = 0;
        rc = upload_residues(ctx, src, it.packing, it.alphabet, base, base + nres, ws.packed, ws.residues.as<uint8_t>(), st_h2d, &res_bytes);
        if (rc) return rc;
        if (pb.dev_plan) {
            // the chunk's 16 B/pair offsets instead of 64 B/pair descriptors; the planner runs on its own high-priority
            // stream next to the previous chunk's fill and hands the descriptors to the compute stream
            const size_t scratch_bytes = plan_scratch_bytes((uint32_t)n, (uint32_t)P.n_slots);
            if (!ws.poff.ensure((2 * n + 1) * 8) || !ws.psort.ensure(scratch_bytes)) { ctx->set_error("device allocation failed (planner buffers)"); return BG_ENOMEM; }
            CU_TRY(ctx, cudaMemcpyAsync(ws.poff.p, pb.stage.p ? pb.stage.p : (const void*)it.off, (2 * n + 1) * 8, cudaMemcpyHostToDevice, st_h2d));
            CU_TRY(ctx, cudaEventRecord(ev_h2d[s], st_h2d));
            if (prof) { cudaEvent_t pe_b = ws.get_event(); cudaEventRecord(pe_b, st_h2d); ws.evs.push_back(PhaseEv{pe_a, pe_b, 4}); }
            CU_TRY(ctx, cudaStreamWaitEvent(st_plan, ev_h2d[s], 0));
            cudaEvent_t pp_a = nullptr;
            if (prof) { pp_a = ws.get_event(); cudaEventRecord(pp_a, st_plan); }
            PlanArgs pa = pb.pa;
            pa.off = ws.poff.as<uint64_t>(); pa.base = base; pa.desc = ws.desc.as<PairDesc>();
            CU_TRY(ctx, launch_plan(pa, pb.need_sort, (uint32_t)P.n_slots, ws.psort.p, scratch_bytes, st_plan));
            if (prof) { cudaEvent_t pp_b = ws.get_event(); cudaEventRecord(pp_b, st_plan); ws.evs.push_back(PhaseEv{pp_a, pp_b, 5}); }
            CU_TRY(ctx, cudaEventRecord(ev_plan[s], st_plan));
            CU_TRY(ctx, cudaStreamWaitEvent(st_comp, ev_plan[s], 0));
            ctx->h2d += res_bytes + (2 * n + 1) * 8;
            ctx->launches += pb.need_sort ? 16 : 10;
        } else {
            if (P.n_slots) CU_TRY(ctx, cudaMemcpyAsync(ws.desc.p, pb.stage.p, P.n_slots * sizeof(PairDesc), cudaMemcpyHostToDevice, st_h2d));
            CU_TRY(ctx, cudaEventRecord(ev_h2d[s], st_h2d));
            ctx->h2d += res_bytes + P.n_slots * sizeof(PairDesc);
        }
        CU_TRY(ctx, cudaStreamWaitEvent(st_comp, ev_h2d[s], 0));
        rc = upload_params(ctx, ws, pp);
        if (rc) return rc;
        AlignIO io{ws.residues.as<uint8_t>(), ws.desc.as<PairDesc>(), &P, n,
                   ws.score.as<int32_t>(), ws.flags.as<uint8_t>(), ws.lens2.as<uint64_t>(), ws.off.as<uint64_t>(), nullptr};
        uint64_t nsub = 0;
        if (ops_mode) {
            // expansion tasks of ~128k columns (256 KB of strings, cache resident): pairs per task from the item's mean length, a power of two
            const uint64_t mean = std::max<uint64_t>(1, nres / (2 * std::max<uint64_t>(1, n)));
            uint64_t stride = 64;
            while (stride < 8192 && stride * mean < (128u << 10)) stride *= 2;
            it.stride = stride;
            nsub = (n + stride - 1) / stride;
            if (!ws.samples.ensure((nsub + 1) * 16) || !ws.samples_h.ensure((nsub + 1) * 16)) { ctx->set_error("allocation failed (offset samples)"); return BG_ENOMEM; }
            io.len = ws.len.as<uint32_t>(); io.first = ws.first.as<uint32_t>(); io.ops = ws.ops.as<uint32_t>();
            io.samples = ws.samples.as<ulonglong2>(); io.sample_stride = stride;
        }
        rc = run_align(ctx, ws, io, pp);
        if (rc) return rc;
        // which stream the item's last kernels went to (run_align: post stream unless the plan has several launches)
        size_t n_launch = 0;
        for (const LaunchClass& lc : P.classes) n_launch += lc.chunks.size();
        cudaStream_t st_last = (ws.post_stream && !pp.score_only && P.max_wave_slots == 0 &&
                                (n_launch == 1 || P.total_trace_words <= ws.split_cap_words)) ? ws.post_stream : st_comp;
        CU_TRY(ctx, cudaEventRecord(ev_comp[s], st_last));
        CU_TRY(ctx, cudaStreamWaitEvent(st_small, ev_comp[s], 0));
        uint64_t* h_total = ws.scalars.as<uint64_t>();
        uint32_t* h_err = reinterpret_cast<uint32_t*>(h_total + 1);
        *h_total = 0;
        CU_TRY(ctx, cudaMemcpyAsync(h_err, ws.err.p, 4, cudaMemcpyDeviceToHost, st_small));
        int32_t* o_score = it.map ? it.t_score : J.oo.score + it.lo;
        uint8_t* o_status = it.map ? it.t_status : J.oo.status + it.lo;
        if (ops_mode) {
            CU_TRY(ctx, cudaMemcpyAsync(ws.samples_h.p, ws.samples.p, (nsub + 1) * 16, cudaMemcpyDeviceToHost, st_small));
            if (n) {
                CU_TRY(ctx, cudaMemcpyAsync(it.map ? it.t_len : J.oo.len + it.lo, ws.len.p, n * 4, cudaMemcpyDeviceToHost, st_small));
                CU_TRY(ctx, cudaMemcpyAsync(it.map ? it.t_first : J.oo.first + 2 * it.lo, ws.first.p, n * 8, cudaMemcpyDeviceToHost, st_small));
            }
        }
        if (n) {
            CU_TRY(ctx, cudaMemcpyAsync(o_score, ws.score.p, n * 4, cudaMemcpyDeviceToHost, st_small));
            CU_TRY(ctx, cudaMemcpyAsync(o_status, ws.flags.p, n, cudaMemcpyDeviceToHost, st_small));
        }
        CU_TRY(ctx, cudaEventRecord(ws.ev_scan, st_small));
        {
            std::lock_guard<std::mutex> lk(ctx->err_mu);   // (devices share the context's counters)
            ctx->timing.cells += P.cells; ctx->timing.cells_packed16 += P.cells_half; ctx->timing.cells_refilled += P.cells_ckpt;
            ctx->timing.trace_bytes += pp.score_only ? 0 : P.total_trace_words * 4;
        }
        return BG_OK;
    };

    int rc_all = BG_OK;
    for (int k = 0; rc_all == BG_OK; ++k) {
        int c;
        if (long_mode) { if (k >= (int)J.dev_items[d].size()) break; c = J.dev_items[d][k]; }
        else { c = J.next_item.fetch_add(1); if (c >= nitems) break; }
        const int s = k % PIPE_DEPTH;
        {   // work set s is free once the item that used it last has left it
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return finished >= k - PIPE_DEPTH + 1; });
        }
        rc_all = J.rc.load();
        if (rc_all) break;
        const double t0 = prof ? since() : 0;
        rc_all = issue(s, c, k);
        if (prof) fprintf(stderr, "[bgalign]   dev %d item %d (%llu pairs) issued %.2f .. %.2f ms\n", d, c, (unsigned long long)J.items[c].n, t0, since());
        if (rc_all) break;
        { std::lock_guard<std::mutex> lk(mu); taken.push_back(c); issued = k + 1; }
        cv.notify_all();
    }
    if (rc_all) J.fail(rc_all);
    { std::lock_guard<std::mutex> lk(mu); issuer_done = true; }
    cv.notify_all();
    finisher.wait();
    if (prof) fprintf(stderr, "[bgalign] dev %d finisher done at %.2f ms\n", d, since());
    cudaStreamSynchronize(st_h2d); cudaStreamSynchronize(st_plan); cudaStreamSynchronize(st_comp); cudaStreamSynchronize(st_small); cudaStreamSynchronize(st_arena);
    if (prof) {
        fprintf(stderr, "[bgalign] dev %d drained at %.2f ms\n", d, since());
        // GPU-side timeline of the kernels (events on the compute stream), relative to the first one
        cudaEvent_t e0 = nullptr;
        for (int s = 0; s < PIPE_DEPTH && !e0; ++s) if (!dv.ws[s].evs.empty()) e0 = dv.ws[s].evs.front().a;
        if (!dv.ws[0].evs.empty()) e0 = dv.ws[0].evs.front().a;
        for (int s = 0; s < PIPE_DEPTH; ++s)
            for (auto& ev : dv.ws[s].evs) {
                float t_a = 0, dur = 0;
                cudaEventElapsedTime(&t_a, e0, ev.a); cudaEventElapsedTime(&dur, ev.a, ev.b);
                fprintf(stderr, "[bgalign]   gpu ws%d phase %d: start %.3f ms, %.3f ms\n", s, ev.phase, t_a, dur);
            }
    }
    cudaStreamSynchronize(st_post);
    if (st_fill2) cudaStreamSynchronize(st_fill2);
    for (int s = 0; s < PIPE_DEPTH; ++s) {
        dv.ws[s].stream = saved[s]; dv.ws[s].post_stream = nullptr; dv.ws[s].fill2_stream = nullptr; dv.ws[s].launch_parity = -1;
        dv.ws[s].split_cap_words = cap_saved;
        cudaEventDestroy(ev_h2d[s]); cudaEventDestroy(ev_comp[s]); cudaEventDestroy(ev_arena[s]); cudaEventDestroy(ev_plan[s]);
    }
    return J.rc.load();
}

constexpr int EDIT_RETRY_GENERAL = -77;   // internal: a byte outside the sampled 4-symbol alphabet turned up


// bg_edit_distance_batch when every pair fits the bit-parallel kernel (<= 4 distinct bytes, len2 <= 320): the chunk's
// launch slots are built on the device from its offsets (k0_eplan.cuh), so the host does nothing per pair after the
// scan.  Per chunk: H2D (residues, packed or not, + 16 B/pair offsets) | K0e + the three K4b launches on the work
// set's own stream (chunks overlap on the GPU) | D2H of 8 B/pair.  With the host planner (edit_pipeline below, still
// used for richer alphabets and longer pairs) cfg3 was bound by ~35 ns per pair and core of planning.
int edit_pipeline_dev(bg_ctx* ctx, int d, const bg_batch* in, uint64_t lo, uint64_t hi, uint64_t* out, const uint8_t* lut) {
    Device& dv = ctx->devs[d];
    if (cudaSetDevice(dv.ordinal) != cudaSuccess) { ctx->set_error("cudaSetDevice failed"); return BG_ECUDA; }
    const uint64_t* off = in->seq_off;
    // Chunks: equal pair counts (the host does not read the offsets on this path).  A chunk's K4b launches take ~0.7-0.9 ms
    // whatever its size below one wave of threads (one thread per pair, 300 k thread slots): few, large chunks -- the first one
    // half size so that the GPU starts early.
    static const int edit_chunks = [] { const char* e = getenv("BG_EDIT_CHUNKS_DEV"); return e ? std::max(1, atoi(e)) : 12; }();
    std::vector<uint64_t> cb{lo};
    {
        const uint64_t n_all = hi - lo;
        const uint64_t k = std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)edit_chunks, (n_all + 32767) / 32768));
        const uint64_t per = std::max<uint64_t>(1, (2 * n_all + 2 * k - 2) / (2 * k - 1));   // first chunk per / 2, then k - 1 chunks of per
        uint64_t at = lo + std::min(n_all, std::max<uint64_t>(1, per / 2));
        while (at < hi) { cb.push_back(at); at += per; }
        cb.push_back(hi);
    }
    // sanity of the chunk boundaries (the pairs in between are validated on the device): monotone, and no chunk so large
    // that it cannot be a batch of short pairs -- anything else goes to the general path, which scans the offsets
    for (size_t c = 0; c + 1 < cb.size(); ++c) {
        const uint64_t o_lo = off[2 * cb[c]], o_hi = off[2 * cb[c + 1]];
        if (o_hi < o_lo || o_hi - o_lo > (cb[c + 1] - cb[c]) * 65536ull + (1ull << 20)) return EDIT_RETRY_GENERAL;
    }
    const int nchunks = (int)cb.size() - 1;
    static const bool prof = getenv("BG_PROFILE_HOST") != nullptr;
    static const bool no_stage = getenv("BG_NO_STAGE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    // pageable caller memory is staged into pinned buffers by pool tasks, chunk by chunk (see host_is_pageable)
    const bool stage_res = !no_stage && hi > lo && host_is_pageable(in->residues);
    // len1 ranges of the device-side counting sort: 64 ranges up to ~1.25 x the longest len1 of a sample (longer ones share the last range)
    uint32_t n_shift = 0;
    {
        uint64_t max_n = 1;
        const uint64_t n_all = hi - lo, stride = std::max<uint64_t>(1, n_all / 512);
        for (uint64_t q = lo; q < hi; q += stride) max_n = std::max(max_n, off[2 * q + 1] - off[2 * q]);
        max_n += max_n / 4;
        while (n_shift < 32 && (max_n >> n_shift) >= 64) ++n_shift;
    }
    const bool stage_off = !no_stage && hi > lo && host_is_pageable(off);
    struct Staged { PinBuf off, res; int rc = BG_OK; TaskHandle th; };
    std::vector<Staged> pre(nchunks);
    if (stage_res || stage_off)
        for (int c = 0; c < nchunks; ++c)
            pre[c].th = host_pool().submit([&, c] {
                const uint64_t lo2 = cb[c], n = cb[c + 1] - cb[c];
                if (stage_off) {
                    if (!pre[c].off.ensure((2 * n + 1) * 8)) { ctx->set_error("pinned staging allocation failed"); pre[c].rc = BG_ENOMEM; return; }
                    memcpy(pre[c].off.p, off + 2 * lo2, (2 * n + 1) * 8);
                }
                if (stage_res) {
                    uint64_t b0, b1; host_byte_range(in->packing, off[2 * lo2], off[2 * (lo2 + n)], b0, b1);
                    if (!pre[c].res.ensure(b1 - b0 + 16)) { ctx->set_error("pinned staging allocation failed"); pre[c].rc = BG_ENOMEM; return; }
                    memcpy(pre[c].res.p, in->residues + b0, b1 - b0);
                }
            });
    cudaStream_t st_h2d = dv.ws[0].walk_stream, st_d2h = dv.ws[1].walk_stream;
    for (int s = 0; s < PIPE_DEPTH; ++s) dv.ws[s].reset_events();
    cudaEvent_t ev_h2d[PIPE_DEPTH], ev_comp[PIPE_DEPTH];
    for (int s = 0; s < PIPE_DEPTH; ++s) {
        cudaEventCreateWithFlags(&ev_h2d[s], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_comp[s], cudaEventDisableTiming);
    }
    PinBuf host_out[PIPE_DEPTH];      // `out` is caller memory of unknown kind: results are staged in pinned memory per work set
    PinBuf lut_pin;
    if (!lut_pin.ensure(256)) { ctx->set_error("pinned staging allocation failed"); return BG_ENOMEM; }
    memcpy(lut_pin.p, lut, 256);

    std::mutex mu; std::condition_variable cv;
    int issued = 0, finished = 0; bool issuer_done = false;
    std::atomic<int> rc_shared{BG_OK};
    std::thread finisher([&] {
        cudaSetDevice(dv.ordinal);
        for (int c = 0;; ++c) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return issued > c || issuer_done; });
                if (issued <= c) break;
            }
            const int s = c % PIPE_DEPTH;
            WorkSet& ws = dv.ws[s];
            int rc = BG_OK;
            if (cudaEventSynchronize(ws.ev_scan) != cudaSuccess) { ctx->set_error("cudaEventSynchronize failed"); rc = BG_ECUDA; }
            const uint32_t* sc = reinterpret_cast<const uint32_t*>(ws.scalars.as<uint64_t>() + 1);
            // a byte outside the alphabet, offsets that are not monotone, or a pair K4b / the 16-byte slot cannot take:
            // the general path (which scans the offsets on the host and reports errors) redoes the batch
            if (!rc && ((sc[0] & 6u) || sc[1])) rc = EDIT_RETRY_GENERAL;
            if (!rc) {
                const uint64_t c_lo = cb[c], c_n = cb[c + 1] - cb[c];
                memcpy(out + c_lo, host_out[s].p, c_n * 8);
                ctx->d2h += c_n * 8;
                const uint64_t cells = ws.scalars.as<uint64_t>()[2];
                static std::mutex timing_mu;   // several devices' finishers add to the same counters
                std::lock_guard<std::mutex> lk(timing_mu);
                ctx->timing.cells += cells; ctx->timing.cells_bitparallel += cells;
            } else {
                int expect = BG_OK; rc_shared.compare_exchange_strong(expect, rc);
            }
            if (prof) fprintf(stderr, "[bgalign]   edit chunk %d results on the host at %.2f ms\n", c, since());
            { std::lock_guard<std::mutex> lk(mu); finished = c + 1; }
            cv.notify_all();
        }
    });

    auto issue = [&](int s, int c) -> int {
        WorkSet& ws = dv.ws[s];
        const uint64_t c_lo = cb[c], n = cb[c + 1] - cb[c];
        const uint64_t base = off[2 * c_lo], nres = off[2 * (c_lo + n)] - base;
        if (n >= 0x7FFFFFF0ull) { ctx->set_error("too many pairs in one pipeline chunk"); return BG_EINVAL_ARG; }
        const size_t scratch_bytes = edit_plan_scratch_bytes((uint32_t)n);
        if (!host_out[s].ensure(std::max<uint64_t>(1, n) * 8) || !ws.scalars.ensure(32)) { ctx->set_error("pinned staging allocation failed"); return BG_ENOMEM; }
        if (pre[c].th.joinable()) pre[c].th.join();
        if (pre[c].rc) return pre[c].rc;
        if (!ws.residues.ensure(nres + 16) || !ws.desc.ensure(std::max<uint64_t>(1, n) * sizeof(MyersSlot)) || !ws.out64.ensure(std::max<uint64_t>(1, n) * 8) ||
            !ws.poff.ensure((2 * n + 1) * 8) || !ws.psort.ensure(scratch_bytes) || !ws.err.ensure(32) || !ws.codes.ensure(512)) {
            ctx->set_error("device allocation failed (pipeline buffers)"); return BG_ENOMEM;
        }
        if (c < PIPE_DEPTH) CU_TRY(ctx, cudaMemcpyAsync(ws.codes.p, lut_pin.p, 256, cudaMemcpyHostToDevice, st_h2d));   // once per work set and call
        uint64_t res_bytes = 0;
        {
            uint64_t hb0, hb1; host_byte_range(in->packing, base, base + nres, hb0, hb1);
            int urc = upload_residues(ctx, stage_res ? (const uint8_t*)pre[c].res.p : in->residues + hb0, in->packing, in->alphabet, base, base + nres,
                                      ws.packed, ws.residues.as<uint8_t>(), st_h2d, &res_bytes);
            if (urc) return urc;
        }
        CU_TRY(ctx, cudaMemcpyAsync(ws.poff.p, stage_off ? (const void*)pre[c].off.p : (const void*)(off + 2 * c_lo), (2 * n + 1) * 8, cudaMemcpyHostToDevice, st_h2d));
        CU_TRY(ctx, cudaEventRecord(ev_h2d[s], st_h2d));
        ctx->h2d += res_bytes + (2 * n + 1) * 8;
        cudaStream_t st = ws.stream;
        CU_TRY(ctx, cudaStreamWaitEvent(st, ev_h2d[s], 0));
        // err words: [0] K4b's flags, [1..4] the planner's class histogram ([4] = pairs that do not fit)
        // (no memsets / small copies on the compute streams: they may be queued on a copy engine behind the next chunks' uploads)
        uint32_t* errw = ws.err.as<uint32_t>();
        {
            Phase ph(ws, 5);
            EditPlanArgs pa;
            pa.off = ws.poff.as<uint64_t>(); pa.base = base; pa.n_pairs = (uint32_t)n; pa.hist = nullptr; pa.cursor = nullptr;
            pa.n_shift = n_shift; pa.slots = ws.desc.as<MyersSlot>(); pa.cls_count = errw + 1;
            CU_TRY(ctx, launch_edit_plan(pa, ws.psort.p, scratch_bytes, errw, st));
            ctx->launches += 3;
        }
        {
            // the three classes are independent and each is a partial wave of a latency-bound kernel (one thread per pair):
            // they run side by side on the work set's stream and its two auxiliary streams
            for (cudaStream_t& a : ws.aux)
                if (!a && cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking) != cudaSuccess) { ctx->set_error("cudaStreamCreate failed"); return BG_ECUDA; }
            MyersArgs ma;
            ma.desc = nullptr; ma.cdesc = ws.desc.as<MyersSlot>(); ma.n_slots = (uint32_t)n; ma.residues = ws.residues.as<uint8_t>();
            ma.lut = ws.codes.as<uint8_t>(); ma.out = ws.out64.as<uint64_t>(); ma.err_flag = errw; ma.cls_count = errw + 1;
            static const int Ws[3] = {4, 8, 10};
            cudaEvent_t ev_plan = ws.get_event();
            CU_TRY(ctx, cudaEventRecord(ev_plan, st));
            for (uint32_t k = 0; k < 3; ++k) {
                cudaStream_t sk = k ? ws.aux[k - 1] : st;
                if (k) CU_TRY(ctx, cudaStreamWaitEvent(sk, ev_plan, 0));
                {
                    Phase ph(ws, 1, sk);
                    ma.cls = k;
                    launch_myers(Ws[k], (uint32_t)n, sk, ma);
                }
                ctx->launches++;
                if (k) {
                    cudaEvent_t ev_k = ws.get_event();
                    CU_TRY(ctx, cudaEventRecord(ev_k, sk));
                    CU_TRY(ctx, cudaStreamWaitEvent(st, ev_k, 0));
                }
            }
            CU_TRY(ctx, cudaGetLastError());
        }
        CU_TRY(ctx, cudaEventRecord(ev_comp[s], st));
        CU_TRY(ctx, cudaStreamWaitEvent(st_d2h, ev_comp[s], 0));
        if (n) CU_TRY(ctx, cudaMemcpyAsync(host_out[s].p, ws.out64.p, n * 8, cudaMemcpyDeviceToHost, st_d2h));
        uint32_t* sc = reinterpret_cast<uint32_t*>(ws.scalars.as<uint64_t>() + 1);
        CU_TRY(ctx, cudaMemcpyAsync(sc, errw, 4, cudaMemcpyDeviceToHost, st_d2h));
        CU_TRY(ctx, cudaMemcpyAsync(sc + 1, errw + 4, 4, cudaMemcpyDeviceToHost, st_d2h));
        CU_TRY(ctx, cudaMemcpyAsync(ws.scalars.as<uint64_t>() + 2, ws.psort.as<uint32_t>() + 254, 8, cudaMemcpyDeviceToHost, st_d2h));   // the chunk's cells (k_eplan_count)
        CU_TRY(ctx, cudaEventRecord(ws.ev_scan, st_d2h));
        return BG_OK;
    };

    int rc_all = BG_OK;
    for (int c = 0; c < nchunks && rc_all == BG_OK; ++c) {
        const int s = c % PIPE_DEPTH;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return finished >= c - PIPE_DEPTH + 1; });
        }
        rc_all = rc_shared.load();
        if (rc_all) break;
        const double t0 = prof ? since() : 0;
        rc_all = issue(s, c);
        if (prof) fprintf(stderr, "[bgalign]   edit chunk %d (%llu pairs, device plan) issued %.2f .. %.2f ms\n", c, (unsigned long long)(cb[c + 1] - cb[c]), t0, since());
        if (rc_all) break;
        { std::lock_guard<std::mutex> lk(mu); issued = c + 1; }
        cv.notify_all();
    }
    { std::lock_guard<std::mutex> lk(mu); issuer_done = true; }
    cv.notify_all();
    finisher.join();
    if (rc_all == BG_OK) rc_all = rc_shared.load();
    cudaStreamSynchronize(st_h2d); cudaStreamSynchronize(st_d2h);
    for (int s = 0; s < PIPE_DEPTH; ++s) cudaStreamSynchronize(dv.ws[s].stream);
    if (prof && !dv.ws[0].evs.empty()) {   // GPU-side timeline of the kernels, relative to the first one
        cudaEvent_t e0 = dv.ws[0].evs.front().a;
        for (int s = 0; s < PIPE_DEPTH; ++s)
            for (auto& ev : dv.ws[s].evs) {
                float t_a = 0, dur = 0;
                cudaEventElapsedTime(&t_a, e0, ev.a); cudaEventElapsedTime(&dur, ev.a, ev.b);
                fprintf(stderr, "[bgalign]   gpu ws%d phase %d: start %.3f ms, %.3f ms\n", s, ev.phase, t_a, dur);
            }
    }
    for (auto& h : host_out) h.release();
    lut_pin.release();
    for (auto& pb : pre) { if (pb.th.joinable()) pb.th.join(); pb.off.release(); pb.res.release(); }
    for (int s = 0; s < PIPE_DEPTH; ++s) { cudaEventDestroy(ev_h2d[s]); cudaEventDestroy(ev_comp[s]); }
    return rc_all;
}

int edit_pipeline(bg_ctx* ctx, int d, const bg_batch* in, uint64_t lo, uint64_t hi, const BatchScan& scan, uint64_t* out, const uint8_t* lut) {
    // Same three-stage structure as align_pipeline (H2D stream | compute stream | D2H stream, issuer + finisher
    // threads); the path is bound by the H2D copy of the residues (4 B of input per ~45 cells).
    Device& dv = ctx->devs[d];
    if (cudaSetDevice(dv.ordinal) != cudaSuccess) { ctx->set_error("cudaSetDevice failed"); return BG_ECUDA; }
    const uint64_t* off = in->seq_off;
    // Edit distance is bound by the host plan (a pair costs the GPU ~3 ns and one host thread ~30-40 ns): many
    // chunks, so that all cores plan at once and the first chunks are ready early.
    static const double edit_chunks = [] { const char* e = getenv("BG_EDIT_CHUNKS"); return e ? std::max(1.0, atof(e)) : 24.0; }();
    const std::vector<uint64_t> cb = chunk_bounds_from_scan(scan, lo, hi, edit_chunks);
    const int nchunks = (int)cb.size() - 1;
    static const bool no_compact = getenv("BG_NO_COMPACT") != nullptr;
    struct Prebuilt { Plan plan; PinBuf stage, res; int rc = BG_OK; TaskHandle th; bool dev_plan = false, need_sort = false; PlanArgs pa; };
    std::vector<Prebuilt> pre(nchunks);   // all chunk plans, one pool task per chunk, consumed as they finish
    static const bool no_stage = getenv("BG_NO_STAGE") != nullptr;
    const bool stage_res = !no_stage && hi > lo && host_is_pageable(in->residues);
    for (int c = 0; c < nchunks; ++c)
        pre[c].th = host_pool().submit([&, c] {
            const uint64_t lo2 = cb[c], n = cb[c + 1] - cb[c];
            if (!pre[c].stage.ensure(plan_desc_capacity(n) * sizeof(PairDesc))) { ctx->set_error("pinned staging allocation failed"); pre[c].rc = BG_ENOMEM; return; }
            if (stage_res) {
                uint64_t b0, b1; host_byte_range(in->packing, off[2 * lo2], off[2 * (lo2 + n)], b0, b1);
                const uint64_t nb = b1 - b0;
                if (!pre[c].res.ensure(nb + 16)) { ctx->set_error("pinned staging allocation failed"); pre[c].rc = BG_ENOMEM; return; }
                memcpy(pre[c].res.p, in->residues + b0, nb);
            }
            const auto t0 = std::chrono::steady_clock::now();
            pre[c].rc = build_plan(ctx, off + 2 * lo2, off[2 * lo2], n, false, 0, 0, 0, pre[c].plan, pre[c].stage.as<PairDesc>(), lut != nullptr);
            if (pre[c].rc == BG_OK && !no_compact) {
                // all classes bit-parallel and the arena offsets fit 48 bits: shrink the slots in place (slot i is read
                // at byte 64 i before anything at or beyond byte 16 i is written)
                Plan& P = pre[c].plan;
                bool all = P.myers && !P.classes.empty();
                for (const LaunchClass& lc : P.classes) all = all && lc.myers_W > 0;
                const PairDesc* src = pre[c].stage.as<PairDesc>();
                for (size_t x = 0; all && x < P.n_slots; ++x)
                    all = src[x].pair_id == 0xFFFFFFFFu || (src[x].a_off < (1ull << 48) && src[x].m <= 0xFFFFu && src[x].b_off == src[x].a_off + src[x].n);
                if (all) {
                    MyersSlot* dst = pre[c].stage.as<MyersSlot>();
                    for (size_t x = 0; x < P.n_slots; ++x) {
                        const PairDesc f = src[x];
                        MyersSlot ms; ms.a_off_lo = (uint32_t)f.a_off; ms.a_off_hi = (uint16_t)(f.a_off >> 32); ms.pair_id = f.pair_id; ms.n = f.n; ms.m = (uint16_t)f.m;
                        dst[x] = ms;
                    }
                    P.compact = true;
                }
            }
            if (getenv("BG_PROFILE_HOST"))
                fprintf(stderr, "[bgalign]   edit plan of chunk %d (%llu pairs): %.2f ms\n", c, (unsigned long long)n,
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        });
    static const bool prof = getenv("BG_PROFILE_HOST") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    // Every work set computes on its OWN stream: a chunk is at most one partial wave of the bit-parallel kernel
    // (one thread per pair, ~1.1 ms whatever the chunk size), so chunks must overlap on the GPU -- on one shared
    // compute stream the pipeline ran at 1.15 ms per chunk regardless of its size (measured, cfg3).
    cudaStream_t st_h2d = dv.ws[0].walk_stream, st_d2h = dv.ws[1].walk_stream;
    cudaStream_t saved[PIPE_DEPTH];
    for (int s = 0; s < PIPE_DEPTH; ++s) { saved[s] = dv.ws[s].stream; dv.ws[s].reset_events(); }
    cudaEvent_t ev_h2d[PIPE_DEPTH], ev_comp[PIPE_DEPTH];
    for (int s = 0; s < PIPE_DEPTH; ++s) {
        cudaEventCreateWithFlags(&ev_h2d[s], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_comp[s], cudaEventDisableTiming);
    }
    // `out` is caller memory of unknown kind: results are staged in pinned memory per work set
    PinBuf host_out[PIPE_DEPTH];

    std::mutex mu; std::condition_variable cv;
    int issued = 0, finished = 0; bool issuer_done = false;
    std::atomic<int> rc_shared{BG_OK};
    std::thread finisher([&] {
        cudaSetDevice(dv.ordinal);
        for (int c = 0;; ++c) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return issued > c || issuer_done; });
                if (issued <= c) break;
            }
            const int s = c % PIPE_DEPTH;
            WorkSet& ws = dv.ws[s];
            int rc = BG_OK;
            if (cudaEventSynchronize(ws.ev_scan) != cudaSuccess) { ctx->set_error("cudaEventSynchronize failed"); rc = BG_ECUDA; }
            if (!rc && lut && (*reinterpret_cast<uint32_t*>(ws.scalars.as<uint64_t>() + 1) & 2u)) rc = EDIT_RETRY_GENERAL;
            if (!rc) {
                const uint64_t c_lo = cb[c], c_n = cb[c + 1] - cb[c];
                memcpy(out + c_lo, host_out[s].p, c_n * 8);
                ctx->d2h += c_n * 8;
            } else {
                int expect = BG_OK; rc_shared.compare_exchange_strong(expect, rc);
            }
            if (prof) fprintf(stderr, "[bgalign]   edit chunk %d results on the host at %.2f ms\n", c, since());
            { std::lock_guard<std::mutex> lk(mu); finished = c + 1; }
            cv.notify_all();
        }
    });

    auto issue = [&](int s, int c) -> int {
        WorkSet& ws = dv.ws[s];
        const uint64_t c_lo = cb[c], n = cb[c + 1] - cb[c];
        const uint64_t base = off[2 * c_lo], nres = off[2 * (c_lo + n)] - base;
        if (!host_out[s].ensure(std::max<uint64_t>(1, n) * 8) || !ws.scalars.ensure(16)) { ctx->set_error("pinned staging allocation failed"); return BG_ENOMEM; }
        if (pre[c].th.joinable()) pre[c].th.join();
        if (pre[c].rc) return pre[c].rc;
        const Plan& P = pre[c].plan;
        if (!ws.residues.ensure(nres + 16) || !ws.desc.ensure(std::max<size_t>(1, P.n_slots) * sizeof(PairDesc)) || !ws.out64.ensure(std::max<uint64_t>(1, n) * 8)) {
            ctx->set_error("device allocation failed (pipeline buffers)"); return BG_ENOMEM;
        }
        uint64_t res_bytes = 0;
        {
            uint64_t hb0, hb1; host_byte_range(in->packing, base, base + nres, hb0, hb1);
            int urc = upload_residues(ctx, stage_res ? (const uint8_t*)pre[c].res.p : in->residues + hb0, in->packing, in->alphabet, base, base + nres,
                                      ws.packed, ws.residues.as<uint8_t>(), st_h2d, &res_bytes);
            if (urc) return urc;
        }
        const size_t slot_bytes = P.compact ? sizeof(MyersSlot) : sizeof(PairDesc);
        if (P.n_slots) CU_TRY(ctx, cudaMemcpyAsync(ws.desc.p, pre[c].stage.p, P.n_slots * slot_bytes, cudaMemcpyHostToDevice, st_h2d));
        CU_TRY(ctx, cudaEventRecord(ev_h2d[s], st_h2d));
        ctx->h2d += res_bytes + P.n_slots * slot_bytes;
        cudaStream_t st_comp = ws.stream;
        CU_TRY(ctx, cudaStreamWaitEvent(st_comp, ev_h2d[s], 0));
        int rc = run_edit(ctx, ws, ws.residues.as<uint8_t>(), ws.desc.as<PairDesc>(), P, ws.out64.as<uint64_t>(), lut);
        if (rc) return rc;
        CU_TRY(ctx, cudaEventRecord(ev_comp[s], st_comp));
        CU_TRY(ctx, cudaStreamWaitEvent(st_d2h, ev_comp[s], 0));
        if (n) CU_TRY(ctx, cudaMemcpyAsync(host_out[s].p, ws.out64.p, n * 8, cudaMemcpyDeviceToHost, st_d2h));
        CU_TRY(ctx, cudaMemcpyAsync(ws.scalars.as<uint64_t>() + 1, ws.err.p, 4, cudaMemcpyDeviceToHost, st_d2h));
        CU_TRY(ctx, cudaEventRecord(ws.ev_scan, st_d2h));
        ctx->timing.cells += P.cells; ctx->timing.cells_bitparallel += P.cells_myers;
        return BG_OK;
    };

    int rc_all = BG_OK;
    for (int c = 0; c < nchunks && rc_all == BG_OK; ++c) {
        const int s = c % PIPE_DEPTH;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return finished >= c - PIPE_DEPTH + 1; });
        }
        rc_all = rc_shared.load();
        if (rc_all) break;
        const double t0 = prof ? since() : 0;
        rc_all = issue(s, c);
        if (prof) fprintf(stderr, "[bgalign]   edit chunk %d (%llu pairs) issued %.2f .. %.2f ms\n", c, (unsigned long long)(cb[c + 1] - cb[c]), t0, since());
        if (rc_all) break;
        { std::lock_guard<std::mutex> lk(mu); issued = c + 1; }
        cv.notify_all();
    }
    { std::lock_guard<std::mutex> lk(mu); issuer_done = true; }
    cv.notify_all();
    finisher.join();
    if (rc_all == BG_OK) rc_all = rc_shared.load();
    cudaStreamSynchronize(st_h2d); cudaStreamSynchronize(st_d2h);
    for (int s = 0; s < PIPE_DEPTH; ++s) cudaStreamSynchronize(dv.ws[s].stream);
    for (auto& h : host_out) h.release();
    for (auto& pb : pre) { if (pb.th.joinable()) pb.th.join(); pb.stage.release(); pb.res.release(); }
    for (int s = 0; s < PIPE_DEPTH; ++s) { dv.ws[s].stream = saved[s]; cudaEventDestroy(ev_h2d[s]); cudaEventDestroy(ev_comp[s]); }
    return rc_all;
}

// Builds the job's items, starts the plan tasks, runs the devices and (strings wanted) the host expansion.
// oo's arrays and (want_strings) off / arena are allocated by the caller.
int run_align_job(AlignJob& J) {
    bg_ctx* ctx = J.ctx;
    const bg_batch* in = J.in;
    const Prepared& pp = *J.pp;
    const BatchScan& scan = *J.scan;
    const uint64_t N = in->n_pairs;
    const uint64_t* off = in->seq_off;
    const int nd = (int)ctx->devs.size();
    const bool ops_mode = J.oo.len != nullptr;
    static const bool prof = getenv("BG_PROFILE_HOST") != nullptr;
    // Batches with long pairs (K2 class) are not cut into pipeline chunks: their traces take GBs per pair, the
    // copies are negligible next to the fill, and every launch should see as many pairs as memory allows.  With
    // several devices the pairs are dealt by size instead -- largest first, each to the device with the least
    // work so far (1 000 pairs of 2.5e9 .. 1e10 cells: a contiguous split leaves the tail to chance).
    J.long_mode = scan.has_wide;
    if (in->packing == BG_PACK_2BIT) make_unpack_lut2(in->alphabet, J.lut2);
    if (N == 0) { if (J.want_strings) J.off[0] = 0; return BG_OK; }
    if (!J.long_mode) {
        const std::vector<uint64_t> cb = chunk_bounds_from_scan(scan, 0, N, 0.0, nd, J.want_strings || !J.oo.len);
        J.items.resize(cb.size() - 1);
        for (size_t c = 0; c + 1 < cb.size(); ++c) {
            WorkItem& it = J.items[c];
            it.lo = cb[c]; it.n = cb[c + 1] - cb[c];
            it.off = off + 2 * it.lo; it.packing = in->packing; it.alphabet = in->alphabet;
            { uint64_t b0, b1; host_byte_range(in->packing, off[2 * it.lo], off[2 * cb[c + 1]], b0, b1); it.res = in->residues + b0; }
            // upper-bound layout of the op arena: B(p) = residues before pair p / 16 + p  (ceil((n + m) / 16) words per pair at most)
            it.ops_base = ((off[2 * it.lo] - off[0]) >> 4) + it.lo;
            it.ops_cap = ((off[2 * cb[c + 1]] - off[0]) >> 4) + cb[c + 1] - it.ops_base;
        }
    } else if (nd == 1) {
        J.items.resize(1);
        WorkItem& it = J.items[0];
        it.lo = 0; it.n = N; it.off = off; it.packing = in->packing; it.alphabet = in->alphabet;
        { uint64_t b0, b1; host_byte_range(in->packing, off[0], off[2 * N], b0, b1); it.res = in->residues + b0; }
        it.ops_base = 0; it.ops_cap = ((off[2 * N] - off[0]) >> 4) + N;
        J.dev_items.assign(1, std::vector<int>{0});
    } else {
        std::vector<uint32_t> order(N);
        for (uint64_t p = 0; p < N; ++p) order[p] = (uint32_t)p;
        auto cells = [&](uint32_t p) { return (double)(off[2ull * p + 1] - off[2ull * p]) * (double)(off[2ull * p + 2] - off[2ull * p + 1]) + 64.0; };
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return cells(x) > cells(y); });
        std::vector<double> load(nd, 0.0);
        std::vector<std::vector<uint32_t>> sets(nd);
        for (uint32_t p : order) {
            int best = 0;
            for (int d = 1; d < nd; ++d) if (load[d] < load[best]) best = d;
            sets[best].push_back(p); load[best] += cells(p);
        }
        J.dev_items.assign(nd, std::vector<int>());
        uint64_t ops_base = 0;
        for (int d = 0; d < nd; ++d) {
            if (sets[d].empty()) continue;
            std::sort(sets[d].begin(), sets[d].end());
            J.items.emplace_back();
            J.dev_items[d].push_back((int)J.items.size() - 1);
        }
        int k = 0;
        for (int d = 0; d < nd; ++d) {
            if (sets[d].empty()) continue;
            WorkItem& it = J.items[k++];
            const uint64_t n = sets[d].size();
            uint64_t nres = 0;
            for (uint32_t p : sets[d]) nres += off[2ull * p + 2] - off[2ull * p];
            const uint64_t res_al = (nres + 63) & ~63ull;
            const uint64_t bytes = res_al + (2 * n + 1) * 8 + n * 4 /*map*/ + n * 4 /*score*/ + n * 4 /*len*/ + n * 8 /*first*/ + n + 64;
            if (!it.gather.ensure(bytes)) { ctx->set_error("pinned staging allocation failed"); return BG_ENOMEM; }
            uint8_t* base = it.gather.as<uint8_t>();
            uint64_t* g_off = reinterpret_cast<uint64_t*>(base + res_al);
            uint32_t* g_map = reinterpret_cast<uint32_t*>(g_off + 2 * n + 1);
            it.t_score = reinterpret_cast<int32_t*>(g_map + n);
            it.t_len = reinterpret_cast<uint32_t*>(it.t_score + n);
            it.t_first = it.t_len + n;
            it.t_status = reinterpret_cast<uint8_t*>(it.t_first + 2 * n);
            uint64_t w = 0;
            for (uint64_t q = 0; q < n; ++q) {
                const uint64_t p = sets[d][q];
                const uint64_t l1 = off[2 * p + 1] - off[2 * p], l2 = off[2 * p + 2] - off[2 * p + 1];
                if (in->packing == BG_PACK_NONE) memcpy(base + w, in->residues + off[2 * p], l1 + l2);
                else unpack_residues(in->residues, in->packing, in->alphabet, off[2 * p], l1 + l2, base + w, in->packing == BG_PACK_2BIT ? J.lut2 : nullptr);
                g_off[2 * q] = w; g_off[2 * q + 1] = w + l1; w += l1 + l2;
                g_map[q] = (uint32_t)p;
            }
            g_off[2 * n] = w;
            it.lo = 0; it.n = n; it.res = base; it.off = g_off; it.map = g_map;
            it.ops_base = ops_base; it.ops_cap = (nres >> 4) + n + 1; ops_base += it.ops_cap;
        }
    }
    const int nitems = (int)J.items.size();
    J.arrived.assign(nitems, 0);
    if (ops_mode && nitems && J.items.back().ops_base + J.items.back().ops_cap > J.oo.ops_cap_words) { ctx->set_error("internal: op arena too small"); return BG_ECUDA; }

    // launch plans of all items, one pool task each, in item order (the first items' plans finish first)
    const uint64_t ws_budget = ctx->trace_budget_words;   // per work set; several fit the B200's 180 GB many times over
    const uint64_t wave_budget = ctx->long_budget_words ? ctx->long_budget_words : std::max<uint64_t>(ctx->trace_budget_words, (uint64_t)(0.8 * (double)ctx->devs[0].total_mem) / 4);
    static const bool no_stage = getenv("BG_NO_STAGE") != nullptr;
    const bool stage_res = !no_stage && !J.long_mode && host_is_pageable(J.items[0].res);
    J.pre = std::vector<Prebuilt>(nitems);
    for (int c = 0; c < nitems; ++c)
        J.pre[c].th = host_pool().submit([&J, c, ctx, ws_budget, wave_budget, stage_res] {
            WorkItem& it = J.items[c];
            Prebuilt& pb = J.pre[c];
            // pipeline chunks are planned on the device from their offsets (k0_plan.cuh); the host only sizes the launches
            pb.dev_plan = !J.long_mode && !ctx->host_plan && !J.pp->score_only && !it.map &&
                          plan_from_stats(ctx, *J.scan, it.lo, it.lo + it.n, ws_budget, J.pp->half_maxabs, pb.plan, pb.pa, pb.need_sort);
            if (pb.dev_plan) {
                if (stage_res) {     // pageable caller memory: stage offsets and residues (see host_is_pageable)
                    uint64_t hb0, hb1; host_byte_range(it.packing, it.off[0], it.off[2 * it.n], hb0, hb1);
                    const uint64_t nb = hb1 - hb0;
                    if (!pb.stage.ensure((2 * it.n + 1) * 8) || !pb.res.ensure(nb + 16)) { ctx->set_error("pinned staging allocation failed"); pb.rc = BG_ENOMEM; return; }
                    memcpy(pb.stage.p, it.off, (2 * it.n + 1) * 8);
                    memcpy(pb.res.p, it.res, nb);
                }
                return;
            }
            if (!pb.stage.ensure(plan_desc_capacity(it.n) * sizeof(PairDesc))) { ctx->set_error("pinned staging allocation failed"); pb.rc = BG_ENOMEM; return; }
            if (stage_res) {
                uint64_t hb0, hb1; host_byte_range(it.packing, it.off[0], it.off[2 * it.n], hb0, hb1);
                const uint64_t nb = hb1 - hb0;
                if (!pb.res.ensure(nb + 16)) { ctx->set_error("pinned staging allocation failed"); pb.rc = BG_ENOMEM; return; }
                memcpy(pb.res.p, it.res, nb);
            }
            const auto t0 = std::chrono::steady_clock::now();
            pb.rc = build_plan(ctx, it.off, it.off[0], it.n, !J.pp->score_only, ws_budget, wave_budget, J.pp->half_maxabs, pb.plan, pb.stage.as<PairDesc>());
            if (getenv("BG_PROFILE_HOST"))
                fprintf(stderr, "[bgalign]   plan of item %d (%llu pairs): %.2f ms\n", c, (unsigned long long)it.n,
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        });

    // one driver per device: the caller's thread takes device 0, the context's persistent workers the others
    std::vector<int> rcs(nd, BG_OK);
    for (int d = 0; d < nd; ++d) { ctx->dev_worker(d); ctx->fin_worker(d); }   // create them here: worker() grows a vector and is not thread safe
    for (int d = 1; d < nd; ++d) ctx->dev_worker(d).run([&J, &rcs, d] { rcs[d] = device_pipeline(J, d); });
    rcs[0] = device_pipeline(J, 0);
    for (int d = 1; d < nd; ++d) ctx->dev_worker(d).wait();
    for (auto& pb : J.pre) { if (pb.th.joinable()) pb.th.join(); pb.stage.release(); pb.res.release(); }
    int rc = J.rc.load();
    for (int d = 0; d < nd && !rc; ++d) rc = rcs[d];

    if (!rc && J.want_strings && ops_mode) {
        std::unique_lock<std::mutex> lk(J.mu);
        J.cv.wait(lk, [&] { return J.pending == 0; });
        if (J.long_mode) {   // items hold scattered pairs: lay the strings out in caller order now
            uint64_t cols = 0;
            for (uint64_t p = 0; p < N; ++p) { J.off[2 * p] = 2 * cols; cols += J.oo.len[p]; }
            submit_expand_locked(J, 0, N, 0, cols);
            J.arena_base = 2 * cols;
        } else if (J.frontier != nitems) { ctx->set_error("internal: items missing at the end of the job"); rc = BG_ECUDA; }
        J.cv.wait(lk, [&] { return J.pending == 0; });
        J.off[2 * N] = J.arena_base;
        if (!rc) rc = J.rc.load();
    } else {
        std::unique_lock<std::mutex> lk(J.mu);
        J.cv.wait(lk, [&] { return J.pending == 0; });
    }
    for (WorkItem& it : J.items) it.gather.release();
    if (prof) fprintf(stderr, "[bgalign] job done, rc %d\n", rc);
    return rc;
}

// Validates, prepares and sizes: the part bg_align_batch and bg_align_batch_ops share.
int begin_align_call(bg_ctx* ctx, const bg_batch* in, const bg_params* p, BatchScan& scan, Prepared& pp) {
    // the scan also counts what the device-side planner needs per kernel shape, which depends on whether the packed
    // kernel is available for these parameters (pick_shape_m's half_ok)
    scan.with_stats = !ctx->host_plan && !(p->flags & BG_F_SCORE_ONLY);
    scan.half_ok = half_maxabs_of(p) > 0;
    scan.force_si = ctx->force_L ? shape_index(Shape{ctx->force_L, ctx->force_C}) : -1;
    int rc = check_batch(ctx, in, &scan);
    if (rc) return rc;
    static const uint64_t zero_off[1] = {0};
    rc = prepare_params(ctx, p, in->n_pairs ? in->seq_off : zero_off, in->n_pairs, pp, &scan);
    if (rc) return rc;
    rc = bg_sync(ctx);   // cached blocks may still be in use by device-resident work (see bg_dresult_free)
    if (rc) return rc;
    ctx->h2d = 0; ctx->d2h = 0; ctx->launches = 0;
    ctx->timing.cells = 0; ctx->timing.trace_bytes = 0; ctx->timing.cells_packed16 = 0; ctx->timing.cells_bitparallel = 0; ctx->timing.cells_refilled = 0;
    return BG_OK;
}

// op words the compact result arrays need for this batch (upper-bound layout, see run_align_job)
uint64_t ops_capacity_words(const bg_batch* in, int nd) {
    const uint64_t N = in->n_pairs;
    if (!N) return 1;
    return ((in->seq_off[2 * N] - in->seq_off[0]) >> 4) + N + (uint64_t)nd + 1;
}

}  // namespace

extern "C" {

int bg_align_batch(bg_ctx* ctx, const bg_batch* in, const bg_params* p, bg_result* out) {
    if (!ctx || !in || !p || !out) return BG_EINVAL_ARG;
    memset(out, 0, sizeof *out);
    static const bool prof = getenv("BG_PROFILE_HOST") != nullptr;
    const auto t_enter = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enter).count(); };
    BatchScan scan;
    Prepared pp;
    int rc = begin_align_call(ctx, in, p, scan, pp);
    if (rc) return rc;
    if (prof) fprintf(stderr, "[bgalign] scan + parameters done at %.2f ms\n", since());
    const uint64_t N = in->n_pairs;
    const uint64_t nres = N ? in->seq_off[2 * N] - in->seq_off[0] : 0;

    HostResultOwner* own = new HostResultOwner();
    HostResultOwner tmp;   // the compact arrays: only needed until the strings are written
    out->n_pairs = N; out->owner_ = own;
    out->score = (int32_t*)own->grab(N * 4); out->status = (uint8_t*)own->grab(N);
    out->off = (uint64_t*)own->grab((2 * N + 1) * 8);
    out->arena = (uint8_t*)own->grab(pp.score_only ? 1 : 2 * nres);   // a pair's strings are at most len1 + len2 long each
    AlignJob J;
    J.ctx = ctx; J.in = in; J.pp = &pp; J.scan = &scan;
    J.oo.score = out->score; J.oo.status = out->status;
    bool ok = out->score && out->status && out->off && out->arena;
    if (!pp.score_only) {
        J.oo.ops_cap_words = ops_capacity_words(in, (int)ctx->devs.size());
        J.oo.len = (uint32_t*)tmp.grab(std::max<uint64_t>(1, N) * 4); J.oo.first = (uint32_t*)tmp.grab(std::max<uint64_t>(1, N) * 8);
        J.oo.ops = (uint32_t*)tmp.grab(J.oo.ops_cap_words * 4 + 64); J.oo.ops_off = (uint64_t*)tmp.grab((N + 1) * 8);
        ok = ok && J.oo.len && J.oo.first && J.oo.ops && J.oo.ops_off;
        J.want_strings = true; J.off = out->off; J.arena = out->arena; J.arena_cap = 2 * nres;
    }
    if (!ok) { tmp.release_all(); bg_result_free(out); ctx->set_error("pinned host allocation failed"); return BG_ENOMEM; }
    if (prof) fprintf(stderr, "[bgalign] pipelines start at %.2f ms\n", since());
    rc = run_align_job(J);
    tmp.release_all();
    if (rc) { bg_result_free(out); return rc; }
    if (pp.score_only) memset(out->off, 0, (2 * N + 1) * 8);
    if (prof) fprintf(stderr, "[bgalign] bg_align_batch returns at %.2f ms\n", since());
    return BG_OK;
}

// Compact form of bg_align_batch: the same alignments as 2-bit ops (see bgalign.h).  A shim that builds its own
// containers (Vec<u8> per Sequence, ds/sequence.rs:10-13) expands each pair straight into them with bg_expand_ops.
int bg_align_batch_ops(bg_ctx* ctx, const bg_batch* in, const bg_params* p, bg_ops_result* out) {
    if (!ctx || !in || !p || !out) return BG_EINVAL_ARG;
    memset(out, 0, sizeof *out);
    if (p->flags & BG_F_SCORE_ONLY) { ctx->set_error("bg_align_batch_ops: use bg_align_batch for score-only calls"); return BG_EINVAL_ARG; }
    BatchScan scan;
    Prepared pp;
    int rc = begin_align_call(ctx, in, p, scan, pp);
    if (rc) return rc;
    const uint64_t N = in->n_pairs;
    HostResultOwner* own = new HostResultOwner();
    AlignJob J;
    J.ctx = ctx; J.in = in; J.pp = &pp; J.scan = &scan;
    J.oo.ops_cap_words = ops_capacity_words(in, (int)ctx->devs.size());
    J.oo.score = (int32_t*)own->grab(N * 4); J.oo.status = (uint8_t*)own->grab(N);
    J.oo.len = (uint32_t*)own->grab(std::max<uint64_t>(1, N) * 4); J.oo.first = (uint32_t*)own->grab(std::max<uint64_t>(1, N) * 8);
    J.oo.ops = (uint32_t*)own->grab(J.oo.ops_cap_words * 4 + 64); J.oo.ops_off = (uint64_t*)own->grab((N + 1) * 8);
    out->n_pairs = N; out->owner_ = own;
    out->score = J.oo.score; out->status = J.oo.status; out->len = J.oo.len; out->first = J.oo.first; out->ops = J.oo.ops; out->ops_off = J.oo.ops_off;
    if (!out->score || !out->status || !out->len || !out->first || !out->ops || !out->ops_off) {
        bg_ops_result_free(out); ctx->set_error("pinned host allocation failed"); return BG_ENOMEM;
    }
    rc = run_align_job(J);
    if (rc) { bg_ops_result_free(out); return rc; }
    out->ops_off[N] = J.oo.ops_cap_words;
    return BG_OK;
}

void bg_ops_result_free(bg_ops_result* r) {
    if (!r) return;
    if (r->owner_) {
        HostResultOwner* own = (HostResultOwner*)r->owner_;
        own->release_all();
        delete own;
    }
    memset(r, 0, sizeof *r);
}

int bg_edit_distance_batch(bg_ctx* ctx, const bg_batch* in, uint64_t* out) {
    if (!ctx || !in || (!out && in->n_pairs)) return BG_EINVAL_ARG;
    static const bool prof = getenv("BG_PROFILE_HOST") != nullptr;
    const auto t_call = std::chrono::steady_clock::now();
    auto since_call = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count(); };
    const int nd = (int)ctx->devs.size();
    int rc = bg_sync(ctx);
    if (rc) return rc;
    ctx->h2d = 0; ctx->d2h = 0; ctx->launches = 0;
    ctx->timing.cells = 0; ctx->timing.trace_bytes = 0; ctx->timing.cells_packed16 = 0; ctx->timing.cells_bitparallel = 0; ctx->timing.cells_refilled = 0;
    // Alphabet from a sample of the batch: <= 4 byte values -> bit-parallel K4b.  The kernel flags any
    // byte outside the sampled alphabet, in which case the batch is redone with the general kernel.
    uint8_t lut[256];
    bool use_lut = false;
    if (in->n_pairs && !getenv("BG_NO_MYERS") && in->packing == BG_PACK_2BIT && in->alphabet) {
        uint64_t hist[256] = {0};
        for (int c = 0; c < 4; ++c) hist[in->alphabet[c]] = 1;      // a 2-bit batch has at most four letters by construction
        use_lut = make_edit_lut(hist, lut);
    } else if (in->n_pairs && !getenv("BG_NO_MYERS") && in->packing == BG_PACK_NONE && in->seq_off && in->residues &&
               in->seq_off[2 * in->n_pairs] >= in->seq_off[0]) {
        const uint64_t b0 = in->seq_off[0], b1 = in->seq_off[2 * in->n_pairs];
        uint64_t hist[256] = {0};
        const uint64_t span = b1 - b0, take = std::min<uint64_t>(span, 1u << 16);   // (2 x 1 MiB of byte increments cost 2 ms)
        for (uint64_t x = 0; x < take; ++x) hist[in->residues[b0 + x]]++;
        for (uint64_t x = 0; x < take; ++x) hist[in->residues[b1 - 1 - x]]++;
        use_lut = span > 0 && make_edit_lut(hist, lut);
    }
    // Fast path (<= 4 letters, every len2 <= 320): the host does not even read the offsets -- the device validates them,
    // buckets the pairs and counts the cells (k0_eplan.cuh); anything it cannot take comes back as a retry and the batch
    // goes through the general path below, which scans on the host and reports errors.
    bool fast = use_lut && !ctx->host_plan && in->seq_off && in->residues;
    if (fast) {   // a look at 512 pairs: batches with longer second sequences go straight to the general path
        const uint64_t stride = std::max<uint64_t>(1, in->n_pairs / 512);
        for (uint64_t q = 0; q < in->n_pairs && fast; q += stride) {
            const uint64_t o0 = in->seq_off[2 * q], o1 = in->seq_off[2 * q + 1], o2 = in->seq_off[2 * q + 2];
            fast = o0 <= o1 && o1 <= o2 && o2 - o1 <= 320;
        }
    }
    if (fast) {
        std::vector<uint64_t> eq(nd + 1);
        for (int d = 0; d <= nd; ++d) eq[d] = in->n_pairs * (uint64_t)d / (uint64_t)nd;
        std::vector<int> rcs(nd, BG_OK);
        auto work = [&](int d) { rcs[d] = eq[d + 1] > eq[d] ? edit_pipeline_dev(ctx, d, in, eq[d], eq[d + 1], out, lut) : BG_OK; };
        if (nd == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (int d = 0; d < nd; ++d) th.emplace_back(work, d);
            for (auto& t : th) t.join();
        }
        if (prof) fprintf(stderr, "[bgalign] edit: device-planned pipelines done at %.2f ms\n", since_call());
        bool retry = false;
        for (int r : rcs) {
            if (r == EDIT_RETRY_GENERAL) retry = true;
            else if (r) return r;
        }
        if (!retry) return BG_OK;
        rc = bg_sync(ctx);
        if (rc) return rc;
        ctx->h2d = 0; ctx->d2h = 0; ctx->launches = 0; ctx->timing.cells = 0; ctx->timing.cells_bitparallel = 0;
    }
    BatchScan scan;
    rc = check_batch(ctx, in, &scan);
    if (rc) return rc;
    if (prof) fprintf(stderr, "[bgalign] edit: scan done at %.2f ms\n", since_call());
    const std::vector<uint64_t> bounds = (in->n_pairs < 8ull * SCAN_BLOCK * nd) ? shard_bounds(in, nd) : shard_bounds_from_scan(scan, in->n_pairs, nd);
    for (int attempt = 0; attempt < 2; ++attempt) {
        std::vector<int> rcs(nd, BG_OK);
        auto work = [&](int d) { rcs[d] = edit_pipeline(ctx, d, in, bounds[d], bounds[d + 1], scan, out, use_lut ? lut : nullptr); };
        if (nd == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (int d = 0; d < nd; ++d) th.emplace_back(work, d);
            for (auto& t : th) t.join();
        }
        if (prof) fprintf(stderr, "[bgalign] edit: pipelines done at %.2f ms\n", since_call());
        bool retry = false;
        for (int r : rcs) {
            if (r == EDIT_RETRY_GENERAL) retry = true;
            else if (r) return r;
        }
        if (!retry) return BG_OK;
        use_lut = false;
        rc = bg_sync(ctx);
        if (rc) return rc;
    }
    return BG_ECUDA;
}

// ---- K5: hamming_distance batched / p_distance_matrix -----------------------------------
// analysis::seq::hamming_distance for every pair (seq.rs:74-83).  Any pair with len1 != len2 makes the call
// return BG_EINVAL_SIZE (the reference returns Err(InvalidInputSize) for that pair), out is then unspecified.
int bg_hamming_distance_batch(bg_ctx* ctx, const bg_batch* in, uint64_t* out) {
    if (!ctx || !in || (!out && in->n_pairs)) return BG_EINVAL_ARG;
    int rc = check_batch(ctx, in);
    if (rc) return rc;
    const uint64_t N = in->n_pairs;
    if (!N) return BG_OK;
    for (uint64_t q = 0; q < N; ++q)
        if (in->seq_off[2 * q + 1] - in->seq_off[2 * q] != in->seq_off[2 * q + 2] - in->seq_off[2 * q + 1]) return BG_EINVAL_SIZE;
    rc = bg_sync(ctx);
    if (rc) return rc;
    const int nd = (int)ctx->devs.size();
    const std::vector<uint64_t> bounds = shard_bounds(in, nd);
    std::vector<int> rcs(nd, BG_OK);
    auto work = [&](int d) {
        const uint64_t lo = bounds[d], hi = bounds[d + 1], n = hi - lo;
        if (!n) return;
        Device& dv = ctx->devs[d];
        WorkSet& ws = dv.ws[0];
        auto fail = [&](int code, const char* what) { ctx->set_error(what); rcs[d] = code; };
        if (cudaSetDevice(dv.ordinal) != cudaSuccess) return fail(BG_ECUDA, "cudaSetDevice failed");
        const uint64_t base = in->seq_off[2 * lo], nres = in->seq_off[2 * hi] - base;
        // pieces of at most HAM_SPLIT bytes; offsets rebased to the shard
        PinBuf stage;
        if (!stage.ensure((2 * n + 1 + n + 1 + n) * 8)) return fail(BG_ENOMEM, "pinned staging allocation failed");
        uint64_t* h_off = stage.as<uint64_t>(); uint64_t* h_first = h_off + 2 * n + 1; uint64_t* h_out = h_first + n + 1;
        uint64_t pieces = 0, max_len = 0;
        for (uint64_t q = 0; q < n; ++q) {
            h_off[2 * q] = in->seq_off[2 * (lo + q)] - base; h_off[2 * q + 1] = in->seq_off[2 * (lo + q) + 1] - base;
            h_first[q] = pieces;
            const uint64_t len = in->seq_off[2 * (lo + q) + 1] - in->seq_off[2 * (lo + q)];
            max_len = std::max(max_len, len);
            pieces += std::max<uint64_t>(1, (len + HAM_SPLIT - 1) / HAM_SPLIT);
        }
        h_off[2 * n] = nres; h_first[n] = pieces;
        if (!ws.residues.ensure(nres + 64) || !ws.off.ensure((2 * n + 1) * 8) || !ws.lens2.ensure((n + 1) * 8) || !ws.out64.ensure(n * 8) || !ws.err.ensure(4)) {
            stage.release(); return fail(BG_ENOMEM, "device allocation failed (hamming)");
        }
        cudaStream_t st = ws.stream;
        uint64_t hb0, hb1, res_bytes = 0; host_byte_range(in->packing, base, base + nres, hb0, hb1);
        if (upload_residues(ctx, in->residues + hb0, in->packing, in->alphabet, base, base + nres, ws.packed, ws.residues.as<uint8_t>(), st, &res_bytes)) { stage.release(); rcs[d] = BG_ECUDA; return; }
        cudaError_t e = cudaSuccess;
        if (e == cudaSuccess) e = cudaMemcpyAsync(ws.off.p, h_off, (2 * n + 1) * 8, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ws.lens2.p, h_first, (n + 1) * 8, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(ws.out64.p, 0, n * 8, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(ws.err.p, 0, 4, st);
        if (e == cudaSuccess) {
            HammingArgs ha{ws.residues.as<uint8_t>(), ws.off.as<uint64_t>(), n, ws.out64.as<uint64_t>(), ws.err.as<uint32_t>()};
            const uint64_t blocks = std::min<uint64_t>((pieces + 7) / 8, (uint64_t)ctx->num_sms * 8);
            {
                Phase ph(ws, 1);     // bg_last_timing().fill_ms = the compare kernel alone
                if (pieces == n) {   // no pair longer than one piece: one lane group per pair
                    launch_hamming_direct(max_len <= 256 ? 8 : max_len <= 1024 ? 16 : 32, n, st, ha);
                } else {
                    launch_hamming_pieces((unsigned)std::max<uint64_t>(1, blocks), st, ha, ws.lens2.as<uint64_t>(), pieces);
                }
            }
            e = cudaGetLastError();
            ctx->launches++;
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, ws.out64.p, n * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { stage.release(); ctx->set_error(std::string("hamming: ") + cudaGetErrorString(e)); rcs[d] = BG_ECUDA; return; }
        memcpy(out + lo, h_out, n * 8);
        ctx->h2d += res_bytes + (3 * n + 2) * 8; ctx->d2h += n * 8;
        stage.release();
    };
    ctx->h2d = 0; ctx->d2h = 0; ctx->launches = 0;
    for (auto& dv : ctx->devs) for (WorkSet& w : dv.ws) w.reset_events();
    ctx->timing = bg_timing{};
    ctx->timing.cells = in->seq_off[2 * N] - in->seq_off[0];      // bytes compared x 2 (both sequences): HBM bytes the kernel reads
    if (nd == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int d = 0; d < nd; ++d) th.emplace_back(work, d);
        for (auto& t : th) t.join();
    }
    for (int r : rcs) if (r) return r;
    return BG_OK;
}

// analysis::stat::p_distance_matrix (stat.rs:138-152): out[i * rows + j] = (#positions where row i and row j
// differ, over the zip of the two rows) as f32 / (len(row 0) as f32); 0 on the diagonal.  rows == 0 is
// BG_EINVAL_SIZE (the reference indexes data[0], tile.rs:32, and panics).  Runs on the context's first device.
int bg_p_distance_matrix(bg_ctx* ctx, const uint8_t* residues, const uint64_t* seq_off, uint64_t rows, float* out) {
    if (!ctx || !seq_off || !out) return BG_EINVAL_ARG;
    if (rows == 0) return BG_EINVAL_SIZE;
    for (uint64_t r = 0; r < rows; ++r) if (seq_off[r + 1] < seq_off[r]) { ctx->set_error("seq_off not monotone"); return BG_EINVAL_ARG; }
    if (rows > (1ull << 20)) { ctx->set_error("p_distance_matrix: too many rows"); return BG_EUNSUPPORTED; }
    int rc = bg_sync(ctx);
    if (rc) return rc;
    Device& dv = ctx->devs[0];
    WorkSet& ws = dv.ws[0];
    CU_TRY(ctx, cudaSetDevice(dv.ordinal));
    const uint64_t base = seq_off[0], nres = seq_off[rows] - base;
    if (nres && !residues) return BG_EINVAL_ARG;
    PinBuf stage;
    if (!stage.ensure((rows + 1) * 8 + rows * rows * 4)) { ctx->set_error("pinned staging allocation failed"); return BG_ENOMEM; }
    uint64_t* h_off = stage.as<uint64_t>();
    float* h_out = reinterpret_cast<float*>(h_off + rows + 1);
    for (uint64_t r = 0; r <= rows; ++r) h_off[r] = seq_off[r] - base;
    if (!ws.residues.ensure(nres + 64) || !ws.off.ensure((rows + 1) * 8) || !ws.arena.ensure(rows * rows * 4)) {
        stage.release(); ctx->set_error("device allocation failed (p_distance_matrix)"); return BG_ENOMEM;
    }
    cudaStream_t st = ws.stream;
    cudaError_t e = nres ? cudaMemcpyAsync(ws.residues.p, residues + base, nres, cudaMemcpyHostToDevice, st) : cudaSuccess;
    if (e == cudaSuccess) e = cudaMemcpyAsync(ws.off.p, h_off, (rows + 1) * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        PDistArgs pa{ws.residues.as<uint8_t>(), ws.off.as<uint64_t>(), rows, (float)(seq_off[1] - seq_off[0]), ws.arena.as<float>()};
        const uint64_t items = rows * (rows + 1) / 2;
        const uint64_t blocks = std::min<uint64_t>((items + 7) / 8, (uint64_t)ctx->num_sms * 8);
        launch_pdist((unsigned)std::max<uint64_t>(1, blocks), st, pa);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, ws.arena.p, rows * rows * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { stage.release(); ctx->set_error(std::string("p_distance_matrix: ") + cudaGetErrorString(e)); return BG_ECUDA; }
    memcpy(out, h_out, rows * rows * 4);
    ctx->h2d = nres + (rows + 1) * 8; ctx->d2h = rows * rows * 4; ctx->launches = 1;
    stage.release();
    return BG_OK;
}

// Host-only diagnostic (no CUDA call): plans a batch of long pairs (lens = n0, m0, n1, m1, ...) under a long-pair
// trace budget and checks the bounded-memory layout it produces.  out[0] = launches (chunks with slots), out[1] =
// bounded-memory chunks, out[2] = largest block count, out[3] = cells on the bounded path, out[4] = violations:
// a chunk whose trace exceeds the budget, a block below 64 rows, a launch table whose blocks do not tile the pair's
// rows exactly once bottom-up, or checkpoint regions that overlap.
int bg_debug_plan_long(const uint64_t* lens, uint64_t n_pairs, uint64_t budget_bytes, uint64_t* out) {
    if (!lens || !out) return BG_EINVAL_ARG;
    bg_ctx ctx;
    std::vector<uint64_t> off(2 * n_pairs + 1, 0);
    for (uint64_t i = 0; i < 2 * n_pairs; ++i) off[i + 1] = off[i] + lens[i];
    std::vector<PairDesc> dst(plan_desc_capacity(n_pairs));
    Plan P;
    const int rc = build_plan(&ctx, off.data(), 0, n_pairs, true, 8192ull << 18, budget_bytes / 4, 0, P, dst.data());
    if (rc) return rc;
    uint64_t launches = 0, ck_chunks = 0, max_nb = 0, viol = 0;
    for (const LaunchClass& lc : P.classes)
        for (const Chunk& ch : lc.chunks) {
            const uint32_t ns = ch.slot_end - ch.slot_begin;
            if (!ns) continue;
            ++launches;
            if (lc.wave && ch.trace_words * 4 > budget_bytes && !(ns == 1 && !ch.ckpt_nb)) ++viol;   // (one whole pair that fits nowhere else is checkpointed, never oversized)
            if (!ch.ckpt_nb) continue;
            ++ck_chunks; max_nb = std::max<uint64_t>(max_nb, ch.ckpt_nb);
            if (ch.trace_words * 4 > budget_bytes) ++viol;
            uint64_t next_off = 0;
            for (uint32_t x = 0; x < ns; ++x) {
                const PairDesc& d = dst[ch.slot_begin + x];
                uint64_t covered = 0, expect_hi = d.n;
                const CkptSlot& p1 = ch.ck_table[x];
                if (p1.row0 != 0 || p1.nrows != d.n || p1.every < 64 || (p1.every & 31u)) ++viol;
                if (p1.ck_off != next_off) ++viol;
                next_off += (uint64_t)(ch.ckpt_nb - 1) * p1.ck_stride;
                if (p1.ck_stride < d.m) ++viol;
                if (d.steps != p1.every + 31u) ++viol;
                for (uint32_t l = 1; l <= ch.ckpt_nb; ++l) {          // bottom-up
                    const CkptSlot& e = ch.ck_table[(size_t)l * ns + x];
                    if (e.nrows == 0) continue;
                    if ((uint64_t)e.row0 + e.nrows != expect_hi || e.nrows > e.every || e.row0 % e.every) ++viol;
                    expect_hi = e.row0; covered += e.nrows;
                }
                if (covered != d.n || expect_hi != 0) ++viol;
            }
            if (next_off != ch.ckpt_elems) ++viol;
        }
    out[0] = launches; out[1] = ck_chunks; out[2] = max_nb; out[3] = P.cells_ckpt; out[4] = viol;
    return BG_OK;
}

// Diagnostic: plans the whole batch as ONE pipeline item twice -- with the host planner (build_plan) and with the
// device-side planner (k0_plan.cuh) -- and compares the two descriptor arrays field by field, class by class (the
// device's descriptor ranges are sized from upper bounds, so its classes start at other slot numbers and end in
// empty slots).  out[0] = 1 if the item is eligible for the device planner, out[1] = descriptors compared,
// out[2] = descriptors that differ, out[3] = device slots beyond the host's count that are not empty.
// The host's one pass over a batch's offsets (scan_batch), without a device: out = {monotone, class mask, max len1 + len2,
// max len2, has pairs wider than the K1 classes, sum of the block costs, nanoseconds the scan took}.  Host-logic tests
// and timing on machines without a GPU.
int bg_debug_scan(const bg_batch* in, int with_stats, int half_ok, uint64_t* out) {
    if (!in || !out || (in->n_pairs && !in->seq_off)) return BG_EINVAL_ARG;
    BatchScan S;
    S.with_stats = with_stats != 0; S.half_ok = half_ok != 0;
    const auto t0 = std::chrono::steady_clock::now();
    scan_batch(in, S);
    const auto t1 = std::chrono::steady_clock::now();
    double cost = 0;
    for (double c : S.block_cost) cost += c;
    out[0] = S.monotone ? 1 : 0; out[1] = S.class_mask; out[2] = S.max_len_sum; out[3] = S.max_m; out[4] = S.has_wide ? 1 : 0;
    out[5] = (uint64_t)cost; out[6] = (uint64_t)std::chrono::duration<double, std::nano>(t1 - t0).count();
    return BG_OK;
}

int bg_debug_plan_compare(bg_ctx* ctx, const bg_batch* in, const bg_params* p, uint64_t* out) {
    if (!ctx || !in || !p || !out) return BG_EINVAL_ARG;
    out[0] = out[1] = out[2] = out[3] = 0;
    BatchScan scan;
    Prepared pp;
    const bool saved = ctx->host_plan;
    ctx->host_plan = false;
    int rc = begin_align_call(ctx, in, p, scan, pp);
    ctx->host_plan = saved;
    if (rc) return rc;
    const uint64_t N = in->n_pairs;
    if (!N || scan.has_wide) return BG_OK;
    const uint64_t budget = 1ull << 40;
    Plan PD; PlanArgs A; bool need_sort = false;
    if (!plan_from_stats(ctx, scan, 0, N, budget, pp.half_maxabs, PD, A, need_sort)) return BG_OK;
    out[0] = 1;
    std::vector<PairDesc> host_desc(plan_desc_capacity(N));
    Plan PH;
    rc = build_plan(ctx, in->seq_off, in->seq_off[0], N, true, budget, budget, pp.half_maxabs, PH, host_desc.data());
    if (rc) return rc;
    Device& dv = ctx->devs[0];
    WorkSet& ws = dv.ws[0];
    CU_TRY(ctx, cudaSetDevice(dv.ordinal));
    const size_t scratch_bytes = plan_scratch_bytes((uint32_t)N, (uint32_t)PD.n_slots);
    if (!ws.poff.ensure((2 * N + 1) * 8) || !ws.desc.ensure(std::max<size_t>(1, PD.n_slots) * sizeof(PairDesc)) || !ws.psort.ensure(scratch_bytes)) {
        ctx->set_error("device allocation failed (planner buffers)"); return BG_ENOMEM;
    }
    CU_TRY(ctx, cudaMemcpyAsync(ws.poff.p, in->seq_off, (2 * N + 1) * 8, cudaMemcpyHostToDevice, ws.stream));
    A.off = ws.poff.as<uint64_t>(); A.base = in->seq_off[0]; A.desc = ws.desc.as<PairDesc>();
    CU_TRY(ctx, launch_plan(A, need_sort, (uint32_t)PD.n_slots, ws.psort.p, scratch_bytes, ws.stream));
    std::vector<PairDesc> dev_desc(PD.n_slots);
    CU_TRY(ctx, cudaMemcpyAsync(dev_desc.data(), ws.desc.p, PD.n_slots * sizeof(PairDesc), cudaMemcpyDeviceToHost, ws.stream));
    CU_TRY(ctx, cudaStreamSynchronize(ws.stream));
    if (PH.classes.size() != PD.classes.size()) { out[2] = ~0ull; return BG_OK; }
    for (size_t k = 0; k < PH.classes.size(); ++k) {
        const LaunchClass& h = PH.classes[k]; const LaunchClass& d = PD.classes[k];
        if (h.chunks.size() != 1 || d.chunks.size() != 1 || h.sh.L != d.sh.L || h.sh.C != d.sh.C || h.half != d.half || h.long_walk != d.long_walk) { out[2] = ~0ull; return BG_OK; }
        const uint32_t hb = h.chunks[0].slot_begin, hn = h.chunks[0].slot_end - hb, db = d.chunks[0].slot_begin, dn = d.chunks[0].slot_end - db;
        if (hn > dn || h.chunks[0].trace_words > d.chunks[0].trace_words) { out[2] = ~0ull; return BG_OK; }
        for (uint32_t j = 0; j < dn; ++j) {
            const PairDesc& y = dev_desc[db + j];
            if (j >= hn) { if (y.pair_id != 0xFFFFFFFFu || y.steps != 0) ++out[3]; continue; }
            const PairDesc& x = host_desc[hb + j];
            ++out[1];
            bool same = x.pair_id == y.pair_id && x.steps == y.steps && x.trace_off == y.trace_off && x.pad_ == y.pad_;
            if (same && x.pair_id != 0xFFFFFFFFu)
                same = x.a_off == y.a_off && x.b_off == y.b_off && x.n == y.n && x.m == y.m && x.nbands == y.nbands &&
                       x.pad_off - host_desc[hb].pad_off == y.pad_off - dev_desc[db].pad_off && (x.nbands <= 1 || x.bnd_off == y.bnd_off);
            if (!same) ++out[2];
        }
    }
    return BG_OK;
}

// Page-lock / unlock caller memory (cudaHostRegister): the host-buffer entry points copy straight out of the
// caller's residue arena, which is a true asynchronous DMA only when that memory is pinned; from pageable memory
// the driver stages every copy synchronously (measured: see INTEGRATION.md).  A shim pins its arena once.
int bg_pin_host(const void* ptr, uint64_t bytes) {
    if (!ptr || !bytes) return BG_EINVAL_ARG;
    const cudaError_t e = cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { (void)cudaGetLastError(); return BG_OK; }
    if (e != cudaSuccess) { (void)cudaGetLastError(); return e == cudaErrorMemoryAllocation ? BG_ENOMEM : BG_ECUDA; }
    return BG_OK;
}
int bg_unpin_host(const void* ptr) {
    if (!ptr) return BG_EINVAL_ARG;
    const cudaError_t e = cudaHostUnregister(const_cast<void*>(ptr));
    if (e != cudaSuccess) { (void)cudaGetLastError(); return BG_ECUDA; }
    return BG_OK;
}

// Host-only diagnostic (no CUDA call): time of build_plan for n_pairs uniform pairs of len x len with traceback,
// the way the host pipeline calls it.  Returns milliseconds (best of `reps`).
double bg_debug_plan_ms(uint64_t n_pairs, uint32_t len, int half, int reps) {
    bg_ctx ctx;
    std::vector<uint64_t> off(2 * n_pairs + 1);
    for (uint64_t i = 0; i <= 2 * n_pairs; ++i) off[i] = i * len;
    std::vector<PairDesc> dst(plan_desc_capacity(n_pairs));
    double best = 1e30;
    for (int r = 0; r < std::max(1, reps); ++r) {
        Plan P;
        const auto t0 = std::chrono::steady_clock::now();
        build_plan(&ctx, off.data(), 0, n_pairs, true, 8192ull << 18, 8192ull << 18, half ? 1 : 0, P, dst.data());
        best = std::min(best, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    return best;
}

const int8_t* bg_score_table26(const char* name) {
    if (!name) return nullptr;
    if (!strcmp(name, "blosum62")) return &BG_TABLE_BLOSUM62[0][0];
    if (!strcmp(name, "pam250")) return &BG_TABLE_PAM250[0][0];
    if (!strcmp(name, "unit")) return &BG_TABLE_UNIT[0][0];
    return nullptr;
}

}  // extern "C"
