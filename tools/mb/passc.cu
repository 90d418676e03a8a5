// Pass C of the entropy epilogue in isolation: cycles per 4-cell trip per scheduler against resident warps
// per SM, for the four-term and the two-term cell.  nvcc -I revealer_b200/csrc ... tools/mb/passc.cu
#include "rvl_common.cuh"
#include <cstdio>
#include <cstring>
#include <cmath>
__device__ double2 tabg[RVL_LOGTAB_ENTRIES];
template <int TERMS>
__global__ void k(int cols, double *out, double epsq)
{
    __shared__ __align__(16) double AT[8][4][32];
    rvl_logtab_fill(tabg);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 8)
        for (int c = 0; c < 4; c++) AT[warp][c][lane] = exp(-0.3 * (lane - 3.0 * c) * (lane - 3.0 * c)) * (1 + c);
    __syncthreads();
    const double(*A)[32] = AT[warp & 7];
    double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    const int G = 25;
    for (int col = 0; col < cols; col++) {
        const double w00 = 1e-3 * (lane + 1 + col), w01 = 2e-5 * (lane + 2), w10 = 3e-7 * (col + 1), w11 = 1e-9 * (lane + 3);
        int i = 0;
#pragma unroll 1
        for (; i + 3 < G; i += 4) {
            double2 lo, hi;
            double q0 = epsq, q1 = epsq, q2 = epsq, q3 = epsq;
            if (TERMS == 4) {
                lo = *reinterpret_cast<const double2 *>(&A[3][i]); hi = *reinterpret_cast<const double2 *>(&A[3][i + 2]);
                q0 = fma(lo.x, w11, q0); q1 = fma(lo.y, w11, q1); q2 = fma(hi.x, w11, q2); q3 = fma(hi.y, w11, q3);
                lo = *reinterpret_cast<const double2 *>(&A[2][i]); hi = *reinterpret_cast<const double2 *>(&A[2][i + 2]);
                q0 = fma(lo.x, w10, q0); q1 = fma(lo.y, w10, q1); q2 = fma(hi.x, w10, q2); q3 = fma(hi.y, w10, q3);
            }
            lo = *reinterpret_cast<const double2 *>(&A[1][i]); hi = *reinterpret_cast<const double2 *>(&A[1][i + 2]);
            q0 = fma(lo.x, w01, q0); q1 = fma(lo.y, w01, q1); q2 = fma(hi.x, w01, q2); q3 = fma(hi.y, w01, q3);
            lo = *reinterpret_cast<const double2 *>(&A[0][i]); hi = *reinterpret_cast<const double2 *>(&A[0][i + 2]);
            q0 = fma(lo.x, w00, q0); q1 = fma(lo.y, w00, q1); q2 = fma(hi.x, w00, q2); q3 = fma(hi.y, w00, q3);
            double l0, l1, l2, l3;
            rvl_log_pos4(q0, q1, q2, q3, l0, l1, l2, l3);
            acc0 = fma(q0, l0, acc0); acc1 = fma(q1, l1, acc1); acc2 = fma(q2, l2, acc2); acc3 = fma(q3, l3, acc3);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (acc0 + acc1) + (acc2 + acc3);
}
template <int TERMS> void run(int warps_per_cta, int ctas_per_sm)
{
    const int cols = 2000, blocks = 148 * ctas_per_sm, threads = 32 * warps_per_cta;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    k<TERMS><<<blocks, threads>>>(cols, out, 1e-12);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<TERMS><<<blocks, threads>>>(cols, out, 1e-12);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double trips = (double)cols * 6 * warps_per_cta * ctas_per_sm / 4.0; // per scheduler
    printf("terms %d  warps/SM %2d : %7.1f cycles per 4-cell trip per scheduler  (%s)\n", TERMS, warps_per_cta * ctas_per_sm,
           ms * 1e-3 * 1.965e9 / trips, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main()
{
    double2 h[RVL_LOGTAB_ENTRIES];
    for (int i = 0; i < RVL_LOGTAB_ENTRIES; i++) {
        const unsigned long long a_hi = 0x3fe6a09eull + (unsigned long long)i * 4096ull, ab = a_hi << 32, bb = (a_hi + 4096ull) << 32;
        double a, b;
        memcpy(&a, &ab, 8); memcpy(&b, &bb, 8);
        h[i].x = 1.0 / (0.5 * (a + b));
        h[i].y = (double)(-logl((long double)h[i].x));
    }
    cudaMemcpyToSymbol(tabg, h, sizeof h);
    for (int w : {1, 2, 4}) { run<4>(4, w); run<2>(4, w); }
    run<4>(8, 2); run<2>(8, 2); run<4>(8, 3); run<2>(8, 3); run<4>(8, 4); run<2>(8, 4);
    return 0;
}
