// Host memory read bandwidth with T threads (AVX2 loads, one stream per thread vs sixteen column streams per thread as
// the cooperative upload's packer reads R's matrix): the ceiling of rvl_load_matrix_f64_colmajor from pageable memory.
//   gcc -O3 -mavx2 -pthread tools/mb/hostbw.c -o /tmp/hostbw && /tmp/hostbw 16
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double *buf;
static size_t N;          // doubles
static int T, mode;
static double sums[256];

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void *work(void *arg)
{
    const int id = (int)(intptr_t)arg;
    const size_t per = N / T, lo = id * per;
    __m256d acc = _mm256_setzero_pd();
    if (mode == 0) {
        for (size_t i = lo; i + 4 <= lo + per; i += 4) acc = _mm256_add_pd(acc, _mm256_loadu_pd(buf + i));
    } else { // sixteen streams, 4096 rows per tile (the packer's access pattern)
        const size_t rows = 4096, ld = per / 16;
        for (size_t r0 = 0; r0 + rows <= ld; r0 += rows)
            for (size_t k = 0; k < rows; k += 4)
                for (int b = 0; b < 16; b++) acc = _mm256_add_pd(acc, _mm256_loadu_pd(buf + lo + ld * b + r0 + k));
    }
    double t[4];
    _mm256_storeu_pd(t, acc);
    sums[id] = t[0] + t[1] + t[2] + t[3];
    return NULL;
}

int main(int argc, char **argv)
{
    T = argc > 1 ? atoi(argv[1]) : 16;
    N = (size_t)100000 * 1000;
    buf = (double *)malloc(N * sizeof(double));
    for (size_t i = 0; i < N; i++) buf[i] = (double)(i & 1);
    for (mode = 0; mode < 2; mode++)
        for (int rep = 0; rep < 4; rep++) {
            pthread_t th[256];
            const double t0 = now();
            for (int i = 0; i < T; i++) pthread_create(&th[i], NULL, work, (void *)(intptr_t)i);
            for (int i = 0; i < T; i++) pthread_join(th[i], NULL);
            const double dt = now() - t0;
            printf("threads %d mode %s: %.2f ms  %.1f GB/s (sum %g)\n", T, mode ? "16-stream tiles" : "linear", dt * 1e3,
                   N * 8 / dt / 1e9, sums[0]);
        }
    return 0;
}
