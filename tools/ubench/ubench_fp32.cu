// Micro-benchmark (B200): issue rate of the FP32 instructions the blend kernels are made of, per SM sub-partition:
// FFMA (3-register), FFMA2 (fma.rn.f32x2), FADD/FADD2, FMUL2, FMNMX (alu pipe), a FFMA+FMNMX mix, MUFU.EX2.
// Each thread runs 8 independent dependency chains, so the pipes -- not latency -- bound the rate.
// Prints warp-instructions per clock per SM (4 sub-partitions).  nvcc -arch=sm_100a -O3 -o ubench_fp32 ubench_fp32.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int ITERS = 2048, CH = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float a, float b) {
    float v[CH];
    u64 w[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { v[i] = a + i + threadIdx.x; w[i] = ((u64)__float_as_uint(v[i]) << 32) | __float_as_uint(b + i); }
    const u64 pa = ((u64)__float_as_uint(a) << 32) | __float_as_uint(a), pb = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) v[i] = fmaf(v[i], a, b);                       // FFMA
            if (MODE == 1) w[i] = ffma2(w[i], pa, pb);                    // FFMA2
            if (MODE == 2) v[i] = v[i] + b;                               // FADD
            if (MODE == 3) w[i] = fadd2(w[i], pb);                        // FADD2
            if (MODE == 4) w[i] = fmul2(w[i], pa);                        // FMUL2
            if (MODE == 5) v[i] = fminf(v[i], b + (float)it);             // FMNMX (+ conversion hoisted? keep b varying)
            if (MODE == 6) { v[i] = fmaf(v[i], a, b); v[i] = fminf(v[i], 1e30f); }   // FFMA + FMNMX (two pipes)
            if (MODE == 7) v[i] = ex2(v[i]);                              // MUFU.EX2
            if (MODE == 8) v[i] = v[i] * a;                               // FMUL
            if (MODE == 9) { w[i] = ffma2(w[i], pa, pb); v[i] = fminf(v[i], 1e30f); } // FFMA2 + FMNMX
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += v[i] + __uint_as_float((unsigned)(w[i] >> 32)) + __uint_as_float((unsigned)w[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int instr_per_iter, float *out, int sms, float clk_ghz) {
    const int grid = sms * 8;
    k<MODE><<<grid, 256>>>(out, 1.0001f, 0.5f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)grid * 8 /*warps*/ * ITERS * CH * instr_per_iter;
    const double clocks = ms * 1e-3 * clk_ghz * 1e9;
    printf("%-16s %8.3f ms   %6.2f warp-instr/clk/SM (%.2f per sub-partition)\n", name, ms, warp_instr / clocks / sms,
           warp_instr / clocks / sms / 4);
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    const float ghz = khz * 1e-6f;
    printf("SMs %d, clock %.3f GHz (nominal max; rates assume it)\n", sms, ghz);
    run<0>("FFMA", 1, out, sms, ghz);
    run<1>("FFMA2", 1, out, sms, ghz);
    run<2>("FADD", 1, out, sms, ghz);
    run<3>("FADD2", 1, out, sms, ghz);
    run<8>("FMUL", 1, out, sms, ghz);
    run<4>("FMUL2", 1, out, sms, ghz);
    run<5>("FMNMX", 1, out, sms, ghz);
    run<6>("FFMA+FMNMX", 2, out, sms, ghz);
    run<9>("FFMA2+FMNMX", 2, out, sms, ghz);
    run<7>("MUFU.EX2", 1, out, sms, ghz);
    return 0;
}
