// l_k0.cu -- K0 (k0_plan.cuh): the device-side launch planner and its radix sort.
#include <cub/device/device_radix_sort.cuh>

#include "launch.h"
#include "k0_plan.cuh"

namespace bg {

size_t plan_sort_tmp_bytes(uint32_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 40, 0);
    return bytes;
}

// keys / ids: [2][n] each (in | sorted); tmp: plan_sort_tmp_bytes(n).  sort == false: one class, one len1 -- identity order.
cudaError_t launch_plan(const PlanArgs& a, bool sort, void* tmp, size_t tmp_bytes, cudaStream_t st) {
    if (!a.n_pairs || !a.n_cls) return cudaSuccess;
    const uint64_t* ks = nullptr; const uint32_t* is = nullptr;
    if (sort) {
        k_plan_keys<<<(a.n_pairs + 255) / 256, 256, 0, st>>>(a);
        uint64_t* keys_out = a.keys + a.n_pairs; uint32_t* ids_out = a.ids + a.n_pairs;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const uint64_t*)a.keys, keys_out, (const uint32_t*)a.ids, ids_out, (int)a.n_pairs, 0, 40, st);
        if (e != cudaSuccess) return e;
        ks = keys_out; is = ids_out;
    }
    k_plan_build<<<a.n_cls, PLAN_TPB, 0, st>>>(a, ks, is);
    return cudaGetLastError();
}

}  // namespace bg
