// FP64 tensor-core MMA (mma.sync.m8n8k4.f64) against the DFMA pipe on B200: does the tensor path carry FP64 contractions
// any faster than the FP64 ALU, and can the two run side by side?  (VERDICT r01 asked for this micro-benchmark before
// considering the rank-4 cell build of k_score's four-term columns on the tensor pipe.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/mb/dmma.cu -o /tmp/dmma && /tmp/dmma
#include <cstdio>
#include <cuda_runtime.h>

#define NACC 8

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// MODE 0: DMMA only; 1: DFMA only (same number of independent chains); 2: both interleaved
template <int mode>
__global__ void k(int iters, double seed, double *out)
{
    double c0[NACC], c1[NACC], f[NACC];
    for (int i = 0; i < NACC; i++) { c0[i] = seed + i; c1[i] = seed - i; f[i] = seed * 0.5 + i; }
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x, m = 0.999999, c = 1e-7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (mode != 1) dmma(c0[i], c1[i], a, b);
            if (mode != 0) f[i] = fma(f[i], m, c);
        }
    }
    double s = 0;
    for (int i = 0; i < NACC; i++) s += c0[i] + c1[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 4, threads = 256, iters = 20000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; mode++) {
        auto launch = [&](int n) {
            if (mode == 0) k<0><<<blocks, threads>>>(n, 1.0, out);
            else if (mode == 1) k<1><<<blocks, threads>>>(n, 1.0, out);
            else k<2><<<blocks, threads>>>(n, 1.0, out);
        };
        launch(100);
        cudaEventRecord(e0);
        launch(iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double warps = (double)blocks * threads / 32, n = (double)iters * NACC;
        const double mma_flop = (mode != 1) ? warps * n * 512.0 : 0.0, fma_flop = (mode != 0) ? warps * 32 * n * 2.0 : 0.0;
        printf("mode %d (%s): %.3f ms  DMMA %.2f TFLOP/s  DFMA %.2f TFLOP/s  (%s)\n", mode,
               mode == 0 ? "DMMA only" : mode == 1 ? "DFMA only" : "DMMA + DFMA interleaved", ms, mma_flop / ms / 1e9, fma_flop / ms / 1e9,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
