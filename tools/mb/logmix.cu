// What bounds the table-driven log on B200?  Four logs per trip (as pass C issues them), 16 warps/SM, with
// pieces knocked out: the int->double conversion of the exponent, the table gather, both.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
__device__ double2 tabg[256];
static __constant__ double LG[5] = {1.0 / 3, -0.5, 1.0 / 5, -0.25, 0.6931471805599453094};
template <int MODE> __device__ __forceinline__ double lg1(double q, const char *base)
{
    const int hi = __double2hiint(q);
    const int hs = hi + (0x3ff00000 - 0x3fe6a09e);
    const int k = (hs >> 20) - 0x3ff;
    const double m = __hiloint2double(hi - (k << 20), __double2loint(q));
    double2 e;
    if (MODE & 2) e = make_double2(1.0 + 1e-9 * (hs & 255), 1e-3);
    else e = *reinterpret_cast<const double2 *>(base + ((hs >> 5) & 0x7f80));
    const double r = fma(m, e.x, -1.0);
    const double r2 = (MODE & 4) ? fma(r, r, 0.0) : r * r;
    const double pa = fma(r, LG[0], LG[1]);
    const double pb = fma(r, LG[2], LG[3]);
    const double p = fma(pb, r2, pa);
    const double lm = fma(r2, p, r);
    double kd;
    if (MODE & 1) kd = __hiloint2double(0x43300000, k ^ 0x80000000) - 4503601774854144.0; // magic-number conversion (1 DADD)
    else kd = (double)k;
    if (MODE & 4) return fma(lm, 1.0, fma(kd, LG[4], e.y));
    return fma(kd, LG[4], e.y) + lm;
}
template <int MODE> __global__ void kern(int trips, double *out)
{
    __shared__ __align__(16) double2 tab[256 * 8];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tab[i] = tabg[i >> 3];
    __syncthreads();
    const char *base = reinterpret_cast<const char *>(tab) + (threadIdx.x & 7) * 16;
    double q0 = 1.0 + threadIdx.x * 1e-3, q1 = q0 * 3.1, q2 = q0 * 1e-5, q3 = q0 * 77.0;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int t = 0; t < trips; t++) {
        const double l0 = lg1<MODE>(q0, base), l1 = lg1<MODE>(q1, base), l2 = lg1<MODE>(q2, base), l3 = lg1<MODE>(q3, base);
        a0 = fma(q0, l0, a0); a1 = fma(q1, l1, a1); a2 = fma(q2, l2, a2); a3 = fma(q3, l3, a3);
        q0 = fma(q0, 1.0000001, 1e-9); q1 = fma(q1, 0.9999999, 1e-9); q2 = fma(q2, 1.0000002, 1e-12); q3 = fma(q3, 0.9999998, 1e-9);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (a0 + a1) + (a2 + a3);
}
// eight logs per trip
template <int MODE> __global__ void kern8(int trips, double *out)
{
    __shared__ __align__(16) double2 tab[256 * 8];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tab[i] = tabg[i >> 3];
    __syncthreads();
    const char *base = reinterpret_cast<const char *>(tab) + (threadIdx.x & 7) * 16;
    double q[8], a[8];
    for (int i = 0; i < 8; i++) { q[i] = (1.0 + threadIdx.x * 1e-3) * (1 + 3.3 * i); a[i] = 0; }
    for (int t = 0; t < trips; t++) {
        double l[8];
#pragma unroll
        for (int i = 0; i < 8; i++) l[i] = lg1<MODE>(q[i], base);
#pragma unroll
        for (int i = 0; i < 8; i++) { a[i] = fma(q[i], l[i], a[i]); q[i] = fma(q[i], 1.0000001, 1e-9); }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run8(const char *what, int ctas)
{
    const int trips = 10000, blocks = 148 * ctas, threads = 256;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    kern8<MODE><<<blocks, threads>>>(trips, out);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern8<MODE><<<blocks, threads>>>(trips, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %6.1f cycles per 4 logs per scheduler (%d warps/SM)\n", what, ms * 1e-3 * 1.965e9 / (trips * 2.0 * ctas) / 2.0 * 1.0, 8 * ctas);
    cudaFree(out);
}
template <int MODE> void run(const char *what)
{
    const int trips = 20000, blocks = 148 * 2, threads = 256;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    kern<MODE><<<blocks, threads>>>(trips, out);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<MODE><<<blocks, threads>>>(trips, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %6.1f cycles per 4-log trip per scheduler\n", what, ms * 1e-3 * 1.965e9 / (trips * 4.0));
    cudaFree(out);
}
int main()
{
    double2 h[256];
    for (int i = 0; i < 256; i++) {
        const unsigned long long a_hi = 0x3fe6a09eull + (unsigned long long)i * 4096ull, ab = a_hi << 32, bb = (a_hi + 4096ull) << 32;
        double a, b;
        memcpy(&a, &ab, 8); memcpy(&b, &bb, 8);
        h[i].x = 1.0 / (0.5 * (a + b));
        h[i].y = (double)(-logl((long double)h[i].x));
    }
    cudaMemcpyToSymbol(tabg, h, sizeof h);
    run<0>("full (I2F.F64 + table gather)");
    run<1>("magic-number int->double instead of I2F");
    run<2>("no table gather");
    run<3>("neither");
    run<4>("full, DMUL and DADD written as DFMA");
    run<6>("no table gather, all DFMA");
    run8<0>("eight logs per trip", 1);
    run8<0>("eight logs per trip", 2);
    run8<0>("eight logs per trip", 4);
    return 0;
}
