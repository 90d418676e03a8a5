// hostpack.cpp -- host side of the cooperative upload of R's F x n matrix of doubles (capi.cu,
// rvl_load_matrix_f64_colmajor).  PCIe is the bottleneck of that call (8 bytes cross the link for every bit the
// device keeps), so while the DMA engine ships the leading 64-sample words as doubles, a few host threads turn
// the trailing words into bits (1/64 of the bytes) -- the same packing and the same 0/1 validation
// k_pack_f64_colmajor does on the device, no arithmetic of the path.  Plain C++ (g++), no CUDA.
#include <atomic>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

// One tile: rows [r0, r1) of sample word w.  m2 is column-major (m2[k + ld*s]); out[k - r0] receives bit b for
// sample 64*w + b.  Returns non-zero when an entry is neither 0 nor 1 (NaN included).
static int pack_tile_scalar(const double *m2, int64_t ld, int64_t r0, int64_t r1, int64_t s0, int ncols, uint64_t *out)
{
    const int64_t cnt = r1 - r0;
    for (int64_t k = 0; k < cnt; k++) out[k] = 0;
    uint64_t bad = 0;
    for (int b = 0; b < ncols; b++) {
        const double *col = m2 + ld * (s0 + b) + r0;
        for (int64_t k = 0; k < cnt; k++) {
            const double v = col[k];
            const uint64_t one = v == 1.0, zero = v == 0.0;
            out[k] |= one << b;
            bad |= (one | zero) ^ 1ull;
        }
    }
    return bad != 0;
}

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
// AVX2: four rows per vector, sixteen columns per pass accumulated in a register before the word is written back --
// sixteen concurrent read streams (what the hardware prefetchers follow well) and one read-modify-write of the
// output per sixteen columns instead of one per column: 1.9x the single-column loop on one core.
__attribute__((target("avx2"))) static int pack_tile_avx2(const double *m2, int64_t ld, int64_t r0, int64_t r1, int64_t s0, int ncols,
                                                          uint64_t *out)
{
    const int64_t cnt = r1 - r0;
    const __m256d one = _mm256_set1_pd(1.0), zero = _mm256_set1_pd(0.0);
    const __m256i lsb = _mm256_set1_epi64x(1);
    __m256i badv = _mm256_setzero_si256();
    uint64_t bad = 0;
    for (int64_t k = 0; k < cnt; k++) out[k] = 0;
    for (int b0 = 0; b0 < ncols; b0 += 16) {
        const int nb = ncols - b0 < 16 ? ncols - b0 : 16;
        const double *base = m2 + ld * (s0 + b0) + r0;
        int64_t k = 0;
        for (; k + 4 <= cnt; k += 4) {
            __m256i acc = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(out + k));
            for (int b = 0; b < nb; b++) {
                const __m256d v = _mm256_loadu_pd(base + ld * b + k);
                const __m256i is1 = _mm256_castpd_si256(_mm256_cmp_pd(v, one, _CMP_EQ_OQ));
                const __m256i is0 = _mm256_castpd_si256(_mm256_cmp_pd(v, zero, _CMP_EQ_OQ));
                acc = _mm256_or_si256(acc, _mm256_sll_epi64(_mm256_and_si256(is1, lsb), _mm_cvtsi32_si128(b0 + b)));
                badv = _mm256_or_si256(badv, _mm256_andnot_si256(_mm256_or_si256(is1, is0), lsb));
            }
            _mm256_storeu_si256(reinterpret_cast<__m256i *>(out + k), acc);
        }
        for (; k < cnt; k++)
            for (int b = 0; b < nb; b++) {
                const double v = base[ld * b + k];
                const uint64_t o = v == 1.0, z = v == 0.0;
                out[k] |= o << (b0 + b);
                bad |= (o | z) ^ 1ull;
            }
    }
    return bad != 0 || !_mm256_testz_si256(badv, badv);
}
static int pack_tile(const double *m2, int64_t ld, int64_t r0, int64_t r1, int64_t s0, int ncols, uint64_t *out)
{
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    return have_avx2 ? pack_tile_avx2(m2, ld, r0, r1, s0, ncols, out) : pack_tile_scalar(m2, ld, r0, r1, s0, ncols, out);
}
#else
static int pack_tile(const double *m2, int64_t ld, int64_t r0, int64_t r1, int64_t s0, int ncols, uint64_t *out)
{
    return pack_tile_scalar(m2, ld, r0, r1, s0, ncols, out);
}
#endif

struct RvlHostPackPlan {
    const double *m2;
    int64_t F, n, ld;
    int64_t W;          // sample words = ceil(n / 64)
    uint64_t *packed;   // [W][F] (only words >= back are written)
    int64_t tile_rows;  // rows per job
    std::mutex mu;
    int64_t front = 0;  // next word the DMA path takes
    int64_t back;       // lowest word claimed by the host threads
    int64_t cur_tile;   // next tile of word `back` (tiles_per_word: none left)
    int64_t tiles_per_word;
    std::atomic<int> bad{0};
    std::vector<std::thread> threads;
};

static void worker(RvlHostPackPlan *p)
{
    for (;;) {
        int64_t w, t;
        {
            std::lock_guard<std::mutex> g(p->mu);
            if (p->cur_tile >= p->tiles_per_word) {   // open the next word from the back, if the DMA path left one
                if (p->back - 1 < p->front) return;
                p->back--;
                p->cur_tile = 0;
            }
            w = p->back;
            t = p->cur_tile++;
        }
        const int64_t r0 = t * p->tile_rows, r1 = (r0 + p->tile_rows < p->F) ? r0 + p->tile_rows : p->F;
        const int64_t s0 = w * 64;
        const int ncols = (int)((p->n - s0 < 64) ? (p->n - s0) : 64);
        if (pack_tile(p->m2, p->ld, r0, r1, s0, ncols, p->packed + (size_t)w * p->F + r0)) p->bad.store(1);
    }
}

extern "C" {

// Start `nthreads` packers on the words at the back of the matrix.  `packed` holds W * F uint64.
void *rvl_hostpack_begin(const double *m2, int64_t F, int64_t n, int64_t ld, uint64_t *packed, int nthreads)
{
    RvlHostPackPlan *p = nullptr;
    try { // nothing may be thrown across the C ABI: no plan == the DMA side does the whole upload
        p = new RvlHostPackPlan();
    } catch (...) {
        return nullptr;
    }
    p->m2 = m2; p->F = F; p->n = n; p->ld = ld;
    p->W = (n + 63) / 64;
    p->packed = packed;
    p->tile_rows = 4096;
    p->tiles_per_word = (F + p->tile_rows - 1) / p->tile_rows;
    p->back = p->W;
    p->cur_tile = p->tiles_per_word;
    try { // fewer threads than asked for (or none) is fine: the DMA side takes whatever the packers do not claim
        for (int i = 0; i < nthreads; i++) p->threads.emplace_back(worker, p);
    } catch (...) {
    }
    return p;
}

// The DMA path asks for its next word (from the front); -1 when the packers hold everything that is left.
int64_t rvl_hostpack_claim_front(void *h)
{
    RvlHostPackPlan *p = static_cast<RvlHostPackPlan *>(h);
    std::lock_guard<std::mutex> g(p->mu);
    if (p->front >= p->back) return -1;
    return p->front++;
}

// Wait for the packers; *first_host_word = first word they produced (W when none), returns non-zero when they
// met an entry that is not 0/1.
int rvl_hostpack_end(void *h, int64_t *first_host_word)
{
    RvlHostPackPlan *p = static_cast<RvlHostPackPlan *>(h);
    for (auto &t : p->threads) t.join();
    *first_host_word = p->back;
    const int bad = p->bad.load();
    delete p;
    return bad;
}
}
