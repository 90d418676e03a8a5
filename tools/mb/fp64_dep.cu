// Dependent-issue latency of FP64 instructions on B200, one warp per scheduler, one chain per thread:
//   0: a = fma(a, m, c)            (destination = first source)
//   1: a = fma(m, c, a)            (destination = accumulator)
//   2: b = fma(a, m, c); a = fma(b, m, c)   (ping-pong between two register pairs)
//   3: a -> a*a -> fma -> a+c ...  (DFMA / DMUL / DADD mix, as in the log's polynomial)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(int iters, double *out, long long *cyc, double m, double c)
{
    double a = 1.0 + threadIdx.x * 1e-3, b = 0.5;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll 8
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) a = fma(a, m, c);
            else if (MODE == 1) a = fma(m, c, a);
            else if (MODE == 2) { b = fma(a, m, c); a = fma(b, m, c); }
            else { b = a * a; b = fma(b, m, a); a = b + c; a = fma(a, m, c); }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(int warps_per_sm, const char *what, int ops)
{
    double *out; long long *cyc, h;
    const int threads = 32 * warps_per_sm, blocks = 148, iters = 2048;
    cudaMalloc(&out, sizeof(double) * blocks * threads); cudaMalloc(&cyc, 8);
    k<MODE><<<blocks, threads>>>(iters, out, cyc, 0.999999, 1e-7); cudaDeviceSynchronize();
    k<MODE><<<blocks, threads>>>(iters, out, cyc, 0.999999, 1e-7); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s %2d warps/SM: %6.2f cycles per dependent FP64 instruction\n", what, warps_per_sm, (double)h / (iters * 8.0 * ops));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w : {4, 16}) {
        run<0>(w, "a = fma(a, m, c)", 1);
        run<1>(w, "a = fma(m, c, a)", 1);
        run<2>(w, "ping-pong a <-> b", 2);
        run<3>(w, "DMUL, DFMA, DADD, DFMA chain", 4);
    }
    return 0;
}
