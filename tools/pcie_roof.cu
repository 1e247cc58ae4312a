// pcie_roof.cu -- the box's host<->device copy roof with N GPUs copying at once (VERDICT r01 next #2d).
//
// Plain CUDA, pinned host memory, ONE cudaMemcpyAsync per copy and GPU, all GPUs started together from one
// thread; a round's time is first-start -> last-end on the host clock after a device sync on both sides.
//   mode h2d / d2h : every GPU copies `mb` MiB in one direction
//   mode both      : every GPU copies `mb` MiB each way on two streams (what a pipelined align step does)
// Output: one JSON object, aggregate GB/s (all GPUs, both directions summed for `both`) per (N, mode, size).
// The e2e legs of bench.py cannot beat bytes / these rates.
//   nvcc -O2 -o tools/pcie_roof tools/pcie_roof.cu && tools/pcie_roof [max_gpus]
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    int max_gpus = argc > 1 ? atoi(argv[1]) : count;
    if (max_gpus > count) max_gpus = count;
    const size_t cap = 512ull << 20;
    std::vector<void*> h_in(max_gpus), h_out(max_gpus), d_in(max_gpus), d_out(max_gpus);
    std::vector<cudaStream_t> s_in(max_gpus), s_out(max_gpus);
    for (int g = 0; g < max_gpus; ++g) {
        CK(cudaSetDevice(g));
        CK(cudaHostAlloc(&h_in[g], cap, cudaHostAllocPortable));
        CK(cudaHostAlloc(&h_out[g], cap, cudaHostAllocPortable));
        CK(cudaMalloc(&d_in[g], cap));
        CK(cudaMalloc(&d_out[g], cap));
        CK(cudaStreamCreateWithFlags(&s_in[g], cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s_out[g], cudaStreamNonBlocking));
        // touch the host pages
        for (size_t x = 0; x < cap; x += 4096) { ((volatile char*)h_in[g])[x] = 1; ((volatile char*)h_out[g])[x] = 1; }
    }
    auto sync_all = [&](int n) { for (int g = 0; g < n; ++g) { cudaSetDevice(g); cudaDeviceSynchronize(); } };
    printf("{\"gpus_visible\": %d, \"results\": [", count);
    bool first = true;
    const int ns[] = {1, 2, 4, 8};
    const size_t mbs[] = {32, 256};
    for (int n : ns) {
        if (n > max_gpus) break;
        for (int mode = 0; mode < 3; ++mode) {
            for (size_t mb : mbs) {
                const size_t bytes = mb << 20;
                const int reps = 6;
                double best = 1e30;
                for (int r = 0; r < reps; ++r) {
                    sync_all(n);
                    const auto t0 = std::chrono::steady_clock::now();
                    for (int g = 0; g < n; ++g) {
                        cudaSetDevice(g);
                        if (mode == 0 || mode == 2) cudaMemcpyAsync(d_in[g], h_in[g], bytes, cudaMemcpyHostToDevice, s_in[g]);
                        if (mode == 1 || mode == 2) cudaMemcpyAsync(h_out[g], d_out[g], bytes, cudaMemcpyDeviceToHost, s_out[g]);
                    }
                    sync_all(n);
                    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    if (r > 0 && dt < best) best = dt;
                }
                const double total = (double)bytes * n * (mode == 2 ? 2 : 1);
                printf("%s\n  {\"n_gpus\": %d, \"mode\": \"%s\", \"mib_per_copy\": %zu, \"aggregate_gbs\": %.1f, \"per_gpu_per_dir_gbs\": %.1f, \"ms\": %.3f}",
                       first ? "" : ",", n, mode == 0 ? "h2d" : mode == 1 ? "d2h" : "both", mb, total / best / 1e9,
                       (double)bytes / best / 1e9, best * 1e3);
                first = false;
            }
        }
    }
    printf("\n]}\n");
    return 0;
}
