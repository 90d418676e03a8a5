// Which producers make an FP64 source operand expensive on B200?  8 DFMA chains per thread, 16 warps/SM; every
// step multiplies by an operand x produced in a different way (the chain value itself stays in the FP64 pipe).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(int iters, double *out, double m, double c, int z)
{
    __shared__ double sh[256];
    sh[threadIdx.x] = 0.999999 + 1e-9 * threadIdx.x;
    __syncthreads();
    double a[8];
    int v[8];
    for (int i = 0; i < 8; i++) { a[i] = 1.0 + threadIdx.x * 1e-3 + i; v[i] = (threadIdx.x * 7 + i) & 255; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            double x;
            if (MODE == 0) x = m;                                                        // loop-invariant register
            else if (MODE == 1) { x = sh[v[i]]; v[i] = (v[i] + z) & 255; }               // shared-memory gather
            else if (MODE == 2) { x = (double)(v[i] + 1); }                              // I2F.F64 (v fixed)
            else if (MODE == 3) { x = __hiloint2double(0x3ff00000 + (v[i] << 4), 0); }   // hi word built by the ALU, lo = 0
            else if (MODE == 4) { x = __hiloint2double(0x3ff00000 + (v[i] << 4), v[i]); } // both words built by the ALU
            else { x = (double)(v[i] + 1); v[i] += z; }                                  // I2F.F64, new value every step
            a[i] = fma(a[i], x, c);
            if (MODE >= 2) a[i] = fma(a[i], 1e-3, c); // keep magnitudes bounded (same extra DFMA for modes 2..5)
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *what, int nf)
{
    double *out; const int threads = 256, blocks = 148 * 2, iters = 8192;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    k<MODE><<<blocks, threads>>>(iters, out, 0.999999, 1e-7, 0); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(iters, out, 0.999999, 1e-7, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-52s %5.2f cycles per FP64 instruction per scheduler\n", what, ms * 1e-3 * 1.965e9 / (4.0 * iters * 8 * nf));
    cudaFree(out);
}
int main()
{
    run<0>("multiplier: loop-invariant register", 1);
    run<1>("multiplier: shared-memory gather every step", 1);
    run<2>("multiplier: I2F.F64 of a fixed int (hoisted?)", 2);
    run<3>("multiplier: hi word from the ALU, lo = 0", 2);
    run<4>("multiplier: both words from the ALU", 2);
    run<5>("multiplier: I2F.F64 every step", 2);
    return 0;
}
