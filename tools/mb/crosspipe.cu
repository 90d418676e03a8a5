// FP64 -> integer -> FP64 dependencies (what the table-driven log does with the exponent and mantissa words):
// does routing each chain through an ALU instruction cost FP64 throughput?  8 chains per thread, 16 warps/SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(int iters, double *out, double m, double c, int z)
{
    double a[8];
    for (int i = 0; i < 8; i++) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            a[i] = fma(a[i], m, c);
            if (MODE == 1) a[i] = __hiloint2double(__double2hiint(a[i]) + z, __double2loint(a[i]));           // hi word through the ALU
            if (MODE == 2) a[i] = __hiloint2double(__double2hiint(a[i]) + z, __double2loint(a[i]) ^ z);       // both words
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *what)
{
    double *out; const int threads = 256, blocks = 148 * 2, iters = 8192;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    k<MODE><<<blocks, threads>>>(iters, out, 0.999999, 1e-7, 0); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(iters, out, 0.999999, 1e-7, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %5.2f cycles per DFMA per scheduler\n", what, ms * 1e-3 * 1.965e9 / (4.0 * iters * 8));
    cudaFree(out);
}
int main()
{
    run<0>("DFMA chains only");
    run<1>("hi word of every result through an IADD");
    run<2>("both words through ALU instructions");
    return 0;
}
