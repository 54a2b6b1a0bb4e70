// Do the DMMA (mma.sync.m8n8k4.f64) and DFMA pipes of sm_100a run concurrently?  Register-resident loops:
// (a) all warps DMMA, (b) all warps DFMA, (c) even warps DMMA / odd warps DFMA, (d) both interleaved in every warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/fp64_mix.cu -o tools/fp64_mix
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// mode 0: DMMA only, 1: DFMA only, 2: split by warp parity, 3: interleaved in every warp (NM dmma + NF dfma per iteration)
template <int MODE, int NM, int NF>
__global__ void __launch_bounds__(512) mix(double* out, int iters, double a0, double b0) {
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    double c[NM][2];
    double f[NF];
#pragma unroll
    for (int i = 0; i < NM; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = i;
    const int warp = threadIdx.x >> 5;
    const bool do_m = MODE == 0 || MODE == 3 || (MODE == 2 && (warp & 1) == 0);
    const bool do_f = MODE == 1 || MODE == 3 || (MODE == 2 && (warp & 1) == 1);
    if (MODE == 3) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < (NM > NF ? NM : NF); ++i) {
                if (i < NM)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
                if (i < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;\n" : "+d"(f[i]) : "d"(a), "d"(b));
            }
        }
    } else if (do_m) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NM; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    } else if (do_f) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;\n" : "+d"(f[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 512));
    const int iters = 20000, threads = 512;     // 16 warps per SM, 4 per scheduler
    const double warps = (double)sms * threads / 32;
    const double fm = 2.0 * 8 * 8 * 4, ff = 2.0 * 32;        // flop per warp instruction
    {
        double ms = time_ms([&] { mix<0, 8, 8><<<sms, threads>>>(out, iters, 1.0, 1e-3); });
        printf("{\"mode\": \"dmma only\", \"ms\": %.4f, \"dmma_tflops\": %.2f}\n", ms, fm * 8 * iters * warps / ms * 1e-9);
    }
    {
        double ms = time_ms([&] { mix<1, 8, 8><<<sms, threads>>>(out, iters, 0.999, 1e-3); });
        printf("{\"mode\": \"dfma only\", \"ms\": %.4f, \"dfma_tflops\": %.2f}\n", ms, ff * 8 * iters * warps / ms * 1e-9);
    }
    {
        // even warps: 8 DMMA per iteration (2048 FMA-lane-ops... = 8*256 FMA), odd warps: 64 DFMA per iteration (64*32 = 2048 FMA): equal work
        double ms = time_ms([&] { mix<2, 8, 64><<<sms, threads>>>(out, iters, 0.999, 1e-3); });
        const double m = fm * 8 * iters * warps / 2, f = ff * 64 * iters * warps / 2;
        printf("{\"mode\": \"split by warp (equal flops)\", \"ms\": %.4f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"total_tflops\": %.2f}\n", ms, m / ms * 1e-9, f / ms * 1e-9, (m + f) / ms * 1e-9);
    }
    {
        double ms = time_ms([&] { mix<3, 8, 8><<<sms, threads>>>(out, iters, 0.999, 1e-3); });
        const double m = fm * 8 * iters * warps, f = ff * 8 * iters * warps;
        printf("{\"mode\": \"interleaved 8 dmma + 8 dfma per warp\", \"ms\": %.4f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"total_tflops\": %.2f}\n", ms, m / ms * 1e-9, f / ms * 1e-9, (m + f) / ms * 1e-9);
    }
    {
        double ms = time_ms([&] { mix<3, 8, 32><<<sms, threads>>>(out, iters, 0.999, 1e-3); });
        const double m = fm * 8 * iters * warps, f = ff * 32 * iters * warps;
        printf("{\"mode\": \"interleaved 8 dmma + 32 dfma per warp\", \"ms\": %.4f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"total_tflops\": %.2f}\n", ms, m / ms * 1e-9, f / ms * 1e-9, (m + f) / ms * 1e-9);
    }
    {
        double ms = time_ms([&] { mix<3, 8, 64><<<sms, threads>>>(out, iters, 0.999, 1e-3); });
        const double m = fm * 8 * iters * warps, f = ff * 64 * iters * warps;
        printf("{\"mode\": \"interleaved 8 dmma + 64 dfma per warp\", \"ms\": %.4f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"total_tflops\": %.2f}\n", ms, m / ms * 1e-9, f / ms * 1e-9, (m + f) / ms * 1e-9);
    }
    return 0;
}
