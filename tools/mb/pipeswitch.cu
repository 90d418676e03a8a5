// Does alternating FP64 and integer instructions cost issue cycles on B200?  8 DFMA chains + 8 integer
// chains per thread, 16 warps/SM; layout A interleaves them one by one, layout B issues 8 DFMA then 8 integer.
#include <cstdio>
#include <cuda_runtime.h>
#ifndef INTOP
#define INTOP "mad.lo.s32 %0, %0, %1, %0;"
#endif
template <int LAYOUT, int NI> __global__ void k(int iters, double *out, int *iout, double m, double c, int im)
{
    double a[8];
    int v[8];
    for (int i = 0; i < 8; i++) { a[i] = 1.0 + threadIdx.x * 1e-3 + i; v[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; it++) {
        if (LAYOUT == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(c));
                if (i < NI) asm volatile(INTOP : "+r"(v[i]) : "r"(im));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(c));
#pragma unroll
            for (int i = 0; i < NI; i++) asm volatile(INTOP : "+r"(v[i]) : "r"(im));
        }
    }
    double s = 0; int t = 0;
    for (int i = 0; i < 8; i++) { s += a[i]; t += v[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int LAYOUT, int NI> void run()
{
    double *out; int *iout; const int threads = 256, blocks = 148 * 2, iters = 8192;
    cudaMalloc(&out, sizeof(double) * blocks * threads); cudaMalloc(&iout, sizeof(int) * blocks * threads);
    k<LAYOUT, NI><<<blocks, threads>>>(iters, out, iout, 0.999999, 1e-7, 3); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<LAYOUT, NI><<<blocks, threads>>>(iters, out, iout, 0.999999, 1e-7, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // per scheduler: 4 warps x iters x 8 DFMA
    printf("layout %s, %d integer ops per 8 DFMA: %5.2f cycles per DFMA per scheduler\n", LAYOUT ? "blocked    " : "interleaved", NI,
           ms * 1e-3 * 1.965e9 / (4.0 * iters * 8));
    cudaFree(out); cudaFree(iout);
}
int main()
{
    run<0, 0>(); run<0, 4>(); run<1, 4>(); run<0, 8>(); run<1, 8>();
    return 0;
}
