// FP32 issue-rate microbenchmark for sm_100a: scalar FFMA / FADD / FMNMX vs packed FFMA2 / FADD2, alone and mixed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fp32 tools/microbench_fp32.cu ; run on a B200.
// Reports achieved TFLOP/s (FMA = 2 flop) and warp-instructions per clock per SM, to ground the FP32 roofline
// denominator and the design choice between scalar and f32x2 arithmetic.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmin_(float a, float b) { float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

constexpr int ILP = 8, INNER = 64;

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, int iters, float seed) {
    float a[ILP], m[ILP];
    u64 p[ILP];
    const float b = seed * 1.0001f, c = seed * 0.5f;
    const u64 b2 = pk(b, b), c2 = pk(c, c);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = seed + i + threadIdx.x; m[i] = a[i]; p[i] = pk(a[i], a[i] + 1.f); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < INNER; ++j) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) a[i] = ffma(a[i], b, c);                                   // FFMA
                if (MODE == 1) p[i] = fma2(p[i], b2, c2);                                 // FFMA2
                if (MODE == 2) a[i] = fadd(a[i], b);                                      // FADD
                if (MODE == 3) p[i] = add2(p[i], b2);                                     // FADD2
                if (MODE == 4) a[i] = fmin_(a[i], b);                                     // FMNMX (alu pipe)
                if (MODE == 5) { a[i] = ffma(a[i], b, c); m[i] = fmin_(m[i], a[i]); }     // FFMA + FMNMX
                if (MODE == 6) { p[i] = fma2(p[i], b2, c2); m[i] = fmin_(m[i], b); }      // FFMA2 + FMNMX
                if (MODE == 7) { a[i] = ffma(a[i], b, c); m[i] = fadd(m[i], b); }         // FFMA + FADD
                if (MODE == 8) { p[i] = fma2(p[i], b2, c2); m[i] = fadd(m[i], b); }       // FFMA2 + FADD
                if (MODE == 9) { p[i] = fma2(p[i], b2, c2); a[i] = ffma(a[i], b, c); }    // FFMA2 + FFMA
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float x, y; upk(p[i], x, y); s += a[i] + m[i] + x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double flop_per_group, double inst_per_group, float* out, int sms) {
    const int iters = 2000, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<blocks, 256>>>(out, 50, 1.f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        bench<MODE><<<blocks, 256>>>(out, iters, 1.f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double groups = (double)blocks * 256 * iters * INNER * ILP;
    const double tflops = groups * flop_per_group / (best * 1e-3) / 1e12;
    const double winst_per_s = groups / 32 * inst_per_group / (best * 1e-3);
    printf("{\"mode\": \"%s\", \"ms\": %.3f, \"tflops\": %.2f, \"warp_inst_per_clk_per_sm_at_1965MHz\": %.3f}\n", name, best, tflops,
           winst_per_s / sms / 1.965e9);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, (size_t)sms * 8 * 256 * sizeof(float));
    run<0>("FFMA", 2, 1, out, sms);
    run<1>("FFMA2", 4, 1, out, sms);
    run<2>("FADD", 1, 1, out, sms);
    run<3>("FADD2", 2, 1, out, sms);
    run<4>("FMNMX", 1, 1, out, sms);
    run<5>("FFMA+FMNMX", 3, 2, out, sms);
    run<6>("FFMA2+FMNMX", 5, 2, out, sms);
    run<7>("FFMA+FADD", 3, 2, out, sms);
    run<8>("FFMA2+FADD", 5, 2, out, sms);
    run<9>("FFMA2+FFMA", 6, 2, out, sms);
    return 0;
}
