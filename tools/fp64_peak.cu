// FP64 peak microbenchmark for B200 (sm_100a): register-resident DMMA and DFMA loops,
// plus a cuBLAS DGEMM yardstick. The DMMA figure is the denominator of every
// "tensor" roofline fraction this repo reports (MEASURED_PEAKS.json has no FP64 entry).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/fp64_peak.cu -lcublas -o tools/fp64_peak
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dmma_m8n8k4(double* out, int iters, double a0, double b0) {
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_m16n8k8(double* out, int iters, double a0, double b0) {
    double a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = a0 + (threadIdx.x + i) * 1e-9;
    b[0] = b0; b[1] = b0 * 0.5;
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double a0, double b0) {
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"}\n", prop.name, sms, prop.major, prop.minor);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256 * 4));
    const int iters = 20000;
    for (int wps = 4; wps <= 16; wps *= 2) {       // warps per SM via blocks of 128 threads
        for (int cta_per_sm = 1; cta_per_sm <= 2; ++cta_per_sm) {
            int threads = wps * 32 / cta_per_sm;
            if (threads < 32 || threads > 256) continue;
            int grid = sms * cta_per_sm;
            {
                double ms = time_ms([&] { dmma_m8n8k4<8><<<grid, threads>>>(out, iters, 1.0, 1e-3); }, 5);
                double flops = 2.0 * 8 * 8 * 4 * 8.0 * iters * (double)(grid * threads / 32);
                printf("{\"kernel\": \"dmma_m8n8k4\", \"warps_per_sm\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", wps, cta_per_sm, ms, flops / ms * 1e-9);
            }
            {
                double ms = time_ms([&] { dmma_m16n8k8<4><<<grid, threads>>>(out, iters, 1.0, 1e-3); }, 5);
                double flops = 2.0 * 16 * 8 * 8 * 4.0 * iters * (double)(grid * threads / 32);
                printf("{\"kernel\": \"dmma_m16n8k8\", \"warps_per_sm\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", wps, cta_per_sm, ms, flops / ms * 1e-9);
            }
            {
                double ms = time_ms([&] { dfma_loop<16><<<grid, threads>>>(out, iters, 0.999, 1e-3); }, 5);
                double flops = 2.0 * 16.0 * iters * (double)(grid * threads);
                printf("{\"kernel\": \"dfma\", \"warps_per_sm\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", wps, cta_per_sm, ms, flops / ms * 1e-9);
            }
        }
    }
    // sustained DMMA: ~2 s back to back
    {
        int grid = sms * 2, threads = 256;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int launches = 60;
        for (int i = 0; i < launches; ++i) dmma_m8n8k4<8><<<grid, threads>>>(out, iters * 4, 1.0, 1e-3);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 8 * 8 * 4 * 8.0 * iters * 4 * (double)(grid * threads / 32) * launches;
        printf("{\"kernel\": \"dmma_m8n8k4_sustained\", \"seconds\": %.3f, \"tflops\": %.3f}\n", ms * 1e-3, flops / ms * 1e-9);
    }
    // cuBLAS DGEMM yardstick
    {
        cublasHandle_t h; cublasCreate(&h);
        for (int n : {1024, 2048, 4096, 8192}) {
            double *A, *B, *C;
            CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n)); CK(cudaMalloc(&C, sizeof(double) * n * n));
            CK(cudaMemset(A, 0, sizeof(double) * n * n)); CK(cudaMemset(B, 0, sizeof(double) * n * n));
            double al = 1.0, be = 0.0;
            double ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n); }, 5);
            printf("{\"kernel\": \"cublasDgemm\", \"n\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", n, ms, 2.0 * n * (double)n * n / ms * 1e-9);
            cudaFree(A); cudaFree(B); cudaFree(C);
        }
        cublasDestroy(h);
    }
    return 0;
}
