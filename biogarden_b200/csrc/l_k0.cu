// l_k0.cu -- K0 (k0_plan.cuh): the device-side launch planner, its radix sort and scans.
#include <cub/device/device_radix_sort.cuh>

#include "launch.h"
#include "k0_plan.cuh"
#include "k0_unpack.cuh"
#include "k0_eplan.cuh"

namespace bg {

namespace {
size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
using PlanIter = cub::TransformInputIterator<PlanSums, PlanItem, cub::CountingInputIterator<uint32_t>>;
}

// Scratch layout (bytes) for an item of n pairs / n_slots descriptor slots:
//   [keys 2 n u64][ids 2 n u32][sums n PlanSums][words (W + 1) u64][woff (W + 1) u64][steps W u32][cub temp]
size_t plan_scratch_bytes(uint32_t n, uint32_t n_slots) {
    size_t sort_b = 0, scan1_b = 0, scan2_b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 40, 0);
    PlanItem fn{};
    PlanIter it(cub::CountingInputIterator<uint32_t>(0), fn);
    cub::DeviceScan::ExclusiveScan(nullptr, scan1_b, it, (PlanSums*)nullptr, PlanSumsAdd(), PlanSums{0, 0, 0, 0}, (int)n, 0);
    cub::DeviceScan::ExclusiveSum(nullptr, scan2_b, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)n_slots + 1, 0);
    const size_t W = n_slots;
    return al256(2ull * n * 8) + al256(2ull * n * 4) + al256((size_t)n * sizeof(PlanSums)) + 2 * al256((W + 1) * 8) + al256(W * 4) +
           al256(std::max(sort_b, std::max(scan1_b, scan2_b))) + 256;
}

// sort == false: one class, one len1 -- the input order is the slot order.
cudaError_t launch_plan(PlanArgs a, bool sort, uint32_t n_slots, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (!a.n_pairs || !a.n_cls) return cudaSuccess;
    const uint32_t n = a.n_pairs;
    unsigned char* p = reinterpret_cast<unsigned char*>(scratch);
    uint64_t* keys = reinterpret_cast<uint64_t*>(p); p += al256(2ull * n * 8);
    uint32_t* ids = reinterpret_cast<uint32_t*>(p); p += al256(2ull * n * 4);
    PlanSums* sums = reinterpret_cast<PlanSums*>(p); p += al256((size_t)n * sizeof(PlanSums));
    uint32_t n_warps = 0;
    for (uint32_t k = 0; k < a.n_cls; ++k) n_warps += a.cls[k].slot_cap / a.cls[k].G2;
    unsigned long long* words = reinterpret_cast<unsigned long long*>(p); p += al256(((size_t)n_slots + 1) * 8);
    unsigned long long* woff = reinterpret_cast<unsigned long long*>(p); p += al256(((size_t)n_slots + 1) * 8);
    uint32_t* steps = reinterpret_cast<uint32_t*>(p); p += al256((size_t)n_slots * 4);
    void* tmp = p;
    size_t tmp_bytes = scratch_bytes - (size_t)(p - reinterpret_cast<unsigned char*>(scratch));
    if (a.uniform && a.n_cls == 1 && !sort) {
        k_plan_uniform<<<(a.cls[0].slot_cap + 255) / 256, 256, 0, st>>>(a);
        return cudaGetLastError();
    }
    a.keys = keys; a.ids = ids;
    k_plan_keys<<<(n + 255) / 256, 256, 0, st>>>(a);
    const uint64_t* ks = keys; const uint32_t* is = ids;
    cudaError_t e;
    if (sort) {
        e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const uint64_t*)keys, keys + n, (const uint32_t*)ids, ids + n, (int)n, 0, 40, st);
        if (e != cudaSuccess) return e;
        ks = keys + n; is = ids + n;
    }
    k_plan_clear<<<(n_slots + 255) / 256, 256, 0, st>>>(a.desc, n_slots);
    PlanItem fn{a, ks, is};
    PlanIter it(cub::CountingInputIterator<uint32_t>(0), fn);
    e = cub::DeviceScan::ExclusiveScan(tmp, tmp_bytes, it, sums, PlanSumsAdd(), PlanSums{0, 0, 0, 0}, (int)n, st);
    if (e != cudaSuccess) return e;
    k_plan_write<<<(n + 255) / 256, 256, 0, st>>>(a, ks, is, sums);
    k_plan_warp_words<<<(n_warps + 1 + 255) / 256, 256, 0, st>>>(a, words, steps, n_warps);
    e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (const unsigned long long*)words, woff, (int)n_warps + 1, st);
    if (e != cudaSuccess) return e;
    k_plan_warp_write<<<(n_warps + 255) / 256, 256, 0, st>>>(a, woff, steps, n_warps);
    return cudaGetLastError();
}

// K0e scratch: [hist 193 u32][cursor 193 u32], padded to 256 each; the cell counter lives in the last 8 bytes of the hist block
size_t edit_plan_scratch_bytes(uint32_t) { return 2048; }
cudaError_t launch_edit_plan(EditPlanArgs a, void* scratch, size_t scratch_bytes, uint32_t* err_flag, cudaStream_t st) {
    if (!a.n_pairs) { k_eplan_zero<<<1, 32, 0, st>>>(a.cls_count, 4, err_flag); return cudaGetLastError(); }
    static_assert(EPLAN_BUCKETS <= 254, "scratch layout");
    if (scratch_bytes < 2048) return cudaErrorInvalidValue;
    a.hist = reinterpret_cast<uint32_t*>(scratch);
    a.cursor = a.hist + 256;
    a.cells = reinterpret_cast<unsigned long long*>(a.hist + 254);
    a.err_flag = err_flag;
    k_eplan_zero<<<1, 256, 0, st>>>(a.hist, 512, err_flag);
    const unsigned grid = (a.n_pairs + 255) / 256;
    k_eplan_count<<<grid, 256, 0, st>>>(a);
    k_eplan_scatter<<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

void launch_unpack(const UnpackArgs& a, cudaStream_t st) {
    if (!a.count) return;
    if (a.bits == 2) k_unpack2<<<(unsigned)((a.count + 64 * 256 - 1) / (64 * 256)), 256, 0, st>>>(a);
    else k_unpack_any<<<(unsigned)((a.count + 255) / 256), 256, 0, st>>>(a);
}

}  // namespace bg
