// DFMA throughput on B200 against operand pattern: the same two sources for every chain (operand reuse) vs
// three distinct register pairs per instruction (what real code issues).  64 warps/SM, ILP 8.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(int iters, double *out, const double *in)
{
    double a[8], b[8], c[8];
    for (int i = 0; i < 8; i++) { a[i] = 1.0 + threadIdx.x * 1e-3 + i; b[i] = in[i] ; c[i] = in[8 + i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = fma(a[i], b[0], c[0]);          // two shared sources
            else if (MODE == 1) a[i] = fma(a[i], b[i], c[i]);     // three distinct register pairs
            else a[i] = fma(b[i], c[(i + 1) & 7], a[i]);          // distinct, accumulator last
        }
    }
    double s = 0; for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const double *in, int threads = 1024)
{
    double *out; int blocks = 148 * 2, iters = 4096;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    k<MODE><<<blocks, threads>>>(iters, out, in); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(iters, out, in); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d, %2d warps/SM: %6.2f TFLOP/s\n", MODE, threads / 32 * 2, 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main()
{
    double h[16]; for (int i = 0; i < 16; i++) h[i] = 0.999 + 1e-4 * i;
    double *in; cudaMalloc(&in, sizeof h); cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    run<0>(in); run<1>(in); run<2>(in);
    run<0>(in, 256); run<1>(in, 256); run<2>(in, 256);
    run<0>(in, 128); run<1>(in, 128); run<2>(in, 128);
    return 0;
}
