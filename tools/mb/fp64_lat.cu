// FP64 pipe micro-benchmark for B200: dependent-issue latency and throughput vs ILP and warps/SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void k(int iters, double *out, long long *cyc)
{
    double a[ILP];
    for (int i = 0; i < ILP; i++) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
    const double m = 0.999999, c = 1e-7;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int warps_per_sm)
{
    double *out; long long *cyc, h;
    int threads = 32 * warps_per_sm, blocks = 148, iters = 4096;
    if (threads > 1024) { blocks = 148 * (threads / 1024); threads = 1024; }
    cudaMalloc(&out, sizeof(double) * blocks * threads); cudaMalloc(&cyc, 8);
    k<ILP><<<blocks, threads>>>(iters, out, cyc); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<ILP><<<blocks, threads>>>(iters, out, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double tf = 2.0 * ILP * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    printf("ILP %d warps/SM %2d : %6.2f cycles per dependent DFMA step (warp 0), %6.2f TFLOP/s\n", ILP, warps_per_sm,
           (double)h / iters, tf);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w : {4, 8, 16, 32, 64}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
