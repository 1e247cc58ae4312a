// int32_peak.cu -- measures the SM integer-pipe issue rates the alignment roofline is quoted against
// ("INT32 peak must be MEASURED").  For each instruction class: 8 independent dependency
// chains per thread, 1024 threads per CTA, 2 CTAs per SM on all SMs; reports lane-ops / clk / SM from
// the SM cycle counter and Tops/s from CUDA events.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o int32_peak int32_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ILP 8
#define ITERS 4096

struct OpIadd   { static constexpr const char* name = "iadd3";            __device__ static int f(int x, int y, int a, int b) { return x + y + b; } };
struct OpViaddmax { static constexpr const char* name = "viaddmnmx_s32";  __device__ static int f(int x, int y, int a, int b) { return __viaddmax_s32(x, a, y); } };
struct OpVimax3 { static constexpr const char* name = "lop3 + vimnmx3_s32 (2 instr)";      __device__ static int f(int x, int y, int a, int b) { return __vimax3_s32(x ^ 1, y, b); } };
struct OpViaddmax16 { static constexpr const char* name = "viaddmnmx_s16x2"; __device__ static int f(int x, int y, int a, int b) { return (int)__viaddmax_s16x2((unsigned)x, (unsigned)a, (unsigned)y); } };
struct OpVimax316 { static constexpr const char* name = "viadd + vimnmx3_s16x2 (2 instr)";  __device__ static int f(int x, int y, int a, int b) { return (int)__vimax3_s16x2((unsigned)x, (unsigned)y, (unsigned)b) + 1; } };
struct OpPrmt   { static constexpr const char* name = "prmt";             __device__ static int f(int x, int y, int a, int b) { int r; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(b)); return r; } };
struct OpLop3   { static constexpr const char* name = "lop3";             __device__ static int f(int x, int y, int a, int b) { int r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(x), "r"(y), "r"(b)); return r; } };
struct OpShf    { static constexpr const char* name = "shf";              __device__ static int f(int x, int y, int a, int b) { int r; asm volatile("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(5)); return r; } };
struct OpSetpSel { static constexpr const char* name = "setp+selp (2 instr)"; __device__ static int f(int x, int y, int a, int b) { int r; asm volatile("{.reg .pred p; setp.ne.s32 p, %1, %2; selp.s32 %0, %3, %1, p;}" : "=r"(r) : "r"(x), "r"(y), "r"(b)); return r; } };
struct OpImad   { static constexpr const char* name = "imad (fma pipe)";  __device__ static int f(int x, int y, int a, int b) { int r; asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(y)); return r; } };
struct OpMixAluFma { static constexpr const char* name = "viaddmnmx + imad interleaved (2 instr)"; __device__ static int f(int x, int y, int a, int b) { int r; asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(y)); return __viaddmax_s32(r, a, y); } };
struct OpShfl   { static constexpr const char* name = "shfl.up + iadd (2 instr)"; __device__ static int f(int x, int y, int a, int b) { return __shfl_up_sync(0xffffffffu, x, 1) + y; } };

template <class Op>
__global__ void __launch_bounds__(1024) bench(int* out, long long* cyc, int a, int b) {
    int x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x + k * a;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = Op::f(x[k], x[(k + 1) % ILP], a, b);
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class Op>
void run(int sms, int* out, long long* cyc, int instr_per_op, FILE* js, bool first) {
    const int blocks = sms * 2;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<Op><<<blocks, 1024>>>(out, cyc, 3, 7);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    const int reps = 5;
    for (int r = 0; r < reps; ++r) bench<Op><<<blocks, 1024>>>(out, cyc, 3, 7);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = (long long*)malloc(blocks * sizeof(long long));
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += (double)h[i]; avg /= blocks;
    free(h);
    const double ops_per_cta = 1024.0 * ILP * ITERS * instr_per_op;
    const double per_clk_sm = 2.0 * ops_per_cta / avg;          // 2 CTAs share an SM
    const double tops = (double)reps * blocks * ops_per_cta / (ms * 1e-3) / 1e12;
    printf("%-42s %8.1f lane-instr/clk/SM   %7.2f T lane-instr/s   (%.3f ms/launch)\n", Op::name, per_clk_sm, tops, ms / reps);
    fprintf(js, "%s\n  {\"op\": \"%s\", \"lane_instr_per_clk_per_sm\": %.2f, \"tera_lane_instr_per_s\": %.3f}", first ? "" : ",", Op::name, per_clk_sm, tops);
}

int main(int argc, char** argv) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s: %d SMs, max clock %d MHz\n", p.name, sms, clk / 1000);
    int* out; long long* cyc;
    cudaMalloc(&out, sms * 2 * 1024 * sizeof(int)); cudaMalloc(&cyc, sms * 2 * sizeof(long long));
    FILE* js = fopen(argc > 1 ? argv[1] : "int32_peak.json", "w");
    fprintf(js, "{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %d, \"ops\": [", p.name, sms, clk / 1000);
    run<OpIadd>(sms, out, cyc, 1, js, true);
    run<OpViaddmax>(sms, out, cyc, 1, js, false);
    run<OpVimax3>(sms, out, cyc, 2, js, false);
    run<OpViaddmax16>(sms, out, cyc, 1, js, false);
    run<OpVimax316>(sms, out, cyc, 2, js, false);
    run<OpPrmt>(sms, out, cyc, 1, js, false);
    run<OpLop3>(sms, out, cyc, 1, js, false);
    run<OpShf>(sms, out, cyc, 1, js, false);
    run<OpSetpSel>(sms, out, cyc, 2, js, false);
    run<OpImad>(sms, out, cyc, 1, js, false);
    run<OpMixAluFma>(sms, out, cyc, 2, js, false);
    run<OpShfl>(sms, out, cyc, 2, js, false);
    fprintf(js, "\n]}\n"); fclose(js);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
