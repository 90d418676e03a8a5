// gct.cu -- GCT v1.2 ingest straight to the bit-packed device matrix.
//
// The reference reads its inputs with MSIG.Gct2Frame (LIB:2616-2626):
//   read.delim(filename, header=T, sep="\t", skip=2, row.names=1, blank.lines.skip=T, ...)
// i.e. two header lines ("#1.2", "<rows>\t<cols>"), one line of column names
// ("Name\tDescription\t<sample>..."), then one line per feature: name, description, <cols> values.
// REVEALER.v1 turns that data frame into an F x n matrix of R doubles (data.matrix, LIB:1582):
// 800 MB at config C3, 40 GB at C4 -- only to have every entry compared with 0/1.
//
// Here the file is memory-mapped; the host touches only the header and, on request, the two name
// columns.  The data lines are shipped as TEXT (about 2 bytes per entry instead of 8) in line-aligned
// chunks through pinned staging, and tokenised on the device: a newline scan (cub select) finds the
// line ends, then one warp per line counts tabs (field index by a warp prefix sum), validates every
// value token as 0 or 1 and ORs it into the bit row at the sample position the caller asked for
// (`cols`: output sample s <- file column cols[s], which is how the sample intersection and the
// sort-by-target permutation of LIB:1596-1633 are applied without a second copy).
#include "rvl_common.cuh"

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

struct rvl_gct {
    int fd = -1;
    const char *base = nullptr; // mmap of the whole file
    size_t size = 0;
    size_t data_off = 0;        // first byte after the column-name line
    int64_t nrow = 0, ncol = 0; // from the "<rows>\t<cols>" line
    std::vector<std::string> samples;
    std::vector<size_t> line_start; // data lines (blank lines skipped), filled on first use
    bool indexed = false;
    std::string err;
};

static const char *line_end(const char *p, const char *end)
{
    const char *q = (const char *)memchr(p, '\n', (size_t)(end - p));
    return q ? q : end;
}

// ---- host side: header, names, single rows ------------------------------------------------------

extern "C" int rvl_gct_open_impl(const char *path, rvl_gct **out, char *errbuf, size_t errcap)
{
    auto fail = [&](const std::string &m) {
        if (errbuf && errcap) snprintf(errbuf, errcap, "%s", m.c_str());
        return -2;
    };
    if (!path || !out) return fail("path/out is NULL");
    *out = nullptr;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(std::string("cannot open ") + path + ": " + strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); return fail(std::string(path) + ": empty or unreadable"); }
    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { close(fd); return fail(std::string("mmap failed: ") + strerror(errno)); }
    rvl_gct *g = new rvl_gct();
    g->fd = fd; g->base = (const char *)m; g->size = (size_t)st.st_size;
    const char *p = g->base, *end = g->base + g->size;
    auto bad = [&](const std::string &msg) {
        munmap((void *)g->base, g->size); close(g->fd); delete g;
        return fail(std::string(path) + ": " + msg);
    };
    // line 1: "#1.2"
    const char *e = line_end(p, end);
    if (e - p < 4 || strncmp(p, "#1.2", 4) != 0) return bad("not a GCT v1.2 file (first line is not #1.2)");
    p = e < end ? e + 1 : end;
    // line 2: rows <tab> cols
    e = line_end(p, end);
    {
        std::string l(p, e);
        long long r = -1, c = -1;
        if (sscanf(l.c_str(), "%lld %lld", &r, &c) != 2 || r < 0 || c < 1) return bad("second line must be <rows> <tab> <cols>");
        g->nrow = r; g->ncol = c;
    }
    p = e < end ? e + 1 : end;
    // line 3: Name, Description, sample names
    e = line_end(p, end);
    {
        const char *le = e;
        if (le > p && le[-1] == '\r') le--;
        std::vector<std::string> f;
        const char *q = p;
        while (q <= le) {
            const char *t = (const char *)memchr(q, '\t', (size_t)(le - q));
            if (!t) t = le;
            f.emplace_back(q, t);
            q = t + 1;
        }
        if ((int64_t)f.size() != g->ncol + 2) return bad("column-name line has " + std::to_string(f.size()) + " fields, expected " + std::to_string(g->ncol + 2));
        g->samples.assign(f.begin() + 2, f.end());
    }
    g->data_off = (size_t)((e < end ? e + 1 : end) - g->base);
    *out = g;
    return 0;
}

extern "C" void rvl_gct_close_impl(rvl_gct *g)
{
    if (!g) return;
    if (g->base) munmap((void *)g->base, g->size);
    if (g->fd >= 0) close(g->fd);
    delete g;
}

static bool blank_line(const char *p, const char *e)
{
    for (; p < e; p++)
        if (*p != '\r' && *p != ' ') return false;
    return true;
}

// Start offsets of the data lines (host scan with memchr, a few threads each on its own byte range cut at a
// newline; only needed for names / single rows).
static int gct_index(rvl_gct *g)
{
    if (g->indexed) return 0;
    const char *beg = g->base + g->data_off, *end = g->base + g->size;
    const size_t len = (size_t)(end - beg);
    const int T = len > ((size_t)8 << 20) ? 4 : 1;
    std::vector<const char *> cut(T + 1, end);
    cut[0] = beg;
    for (int i = 1; i < T; i++) { // segment i starts right after the first newline at or after its nominal start
        const char *p = beg + len / T * i;
        const char *e = (const char *)memchr(p, '\n', (size_t)(end - p));
        cut[i] = e ? e + 1 : end;
        if (cut[i] < cut[i - 1]) cut[i] = cut[i - 1];
    }
    std::vector<std::vector<size_t>> part(T);
    auto scan = [&](int i) {
        const char *p = cut[i], *stop = cut[i + 1];
        part[i].reserve((size_t)g->nrow / T + 16);
        while (p < stop) {
            const char *e = line_end(p, end);
            if (!blank_line(p, e)) part[i].push_back((size_t)(p - g->base));
            p = e < end ? e + 1 : end;
        }
    };
    if (T == 1) scan(0);
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < T; i++) th.emplace_back(scan, i);
        for (auto &t : th) t.join();
    }
    g->line_start.clear();
    g->line_start.reserve((size_t)g->nrow);
    for (int i = 0; i < T; i++) g->line_start.insert(g->line_start.end(), part[i].begin(), part[i].end());
    if ((int64_t)g->line_start.size() != g->nrow) {
        g->err = "file has " + std::to_string(g->line_start.size()) + " data lines, header says " + std::to_string(g->nrow);
        return -2;
    }
    g->indexed = true;
    return 0;
}

extern "C" int rvl_gct_dims_impl(const rvl_gct *g, int64_t *nrow, int64_t *ncol)
{
    if (!g) return -2;
    if (nrow) *nrow = g->nrow;
    if (ncol) *ncol = g->ncol;
    return 0;
}

static int copy_out(const char *p, size_t len, char *buf, size_t cap)
{
    if (!buf || cap == 0) return (int)len;
    size_t k = len < cap - 1 ? len : cap - 1;
    memcpy(buf, p, k);
    buf[k] = 0;
    return (int)len;
}

// Returns the length of the name (may exceed cap - 1: truncated copy), or < 0 on error.
extern "C" int rvl_gct_sample_name_impl(const rvl_gct *g, int64_t j, char *buf, size_t cap)
{
    if (!g || j < 0 || j >= g->ncol) return -2;
    return copy_out(g->samples[j].data(), g->samples[j].size(), buf, cap);
}

// field 0 = row name, field 1 = description
extern "C" int rvl_gct_row_field_impl(rvl_gct *g, int64_t i, int field, char *buf, size_t cap)
{
    if (!g || field < 0 || field > 1) return -2;
    if (gct_index(g)) return -2;
    if (i < 0 || i >= g->nrow) return -2;
    const char *p = g->base + g->line_start[i], *end = g->base + g->size;
    const char *e = line_end(p, end);
    for (int f = 0; f < field; f++) {
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        if (!t) return -2;
        p = t + 1;
    }
    const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
    if (!t) t = e;
    return copy_out(p, (size_t)(t - p), buf, cap);
}

// All names (field 0) or descriptions (field 1) joined by '\n'; returns the bytes needed.
extern "C" int64_t rvl_gct_row_names_impl(rvl_gct *g, int field, char *buf, size_t cap)
{
    if (!g || field < 0 || field > 1) return -2;
    if (gct_index(g)) return -2;
    const char *end = g->base + g->size;
    size_t need = 0;
    for (int64_t i = 0; i < g->nrow; i++) {
        const char *p = g->base + g->line_start[i];
        const char *e = line_end(p, end);
        for (int f = 0; f < field; f++) {
            const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
            if (!t) { g->err = "row " + std::to_string(i) + " has no description field"; return -2; }
            p = t + 1;
        }
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        if (!t) t = e;
        const size_t len = (size_t)(t - p);
        if (buf && need + len + 1 <= cap) {
            memcpy(buf + need, p, len);
            buf[need + len] = (i + 1 < g->nrow) ? '\n' : 0;
        }
        need += len + 1;
    }
    return (int64_t)need;
}

// One row as doubles (the continuous target of ds1, LIB:1586): empty field or "NA" -> NaN like
// read.delim(na.strings = "") followed by data.matrix.
extern "C" int rvl_gct_row_values_impl(rvl_gct *g, int64_t i, double *out)
{
    if (!g || !out) return -2;
    if (gct_index(g)) return -2;
    if (i < 0 || i >= g->nrow) return -2;
    const char *p = g->base + g->line_start[i], *end = g->base + g->size;
    const char *e = line_end(p, end);
    if (e > p && e[-1] == '\r') e--;
    int64_t field = 0;
    while (p <= e) {
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        if (!t) t = e;
        if (field >= 2) {
            if (field - 2 >= g->ncol) { g->err = "row " + std::to_string(i) + " has too many fields"; return -2; }
            std::string tok(p, t);
            double v;
            if (tok.empty() || tok == "NA") v = NAN;
            else {
                char *ep = nullptr;
                v = strtod(tok.c_str(), &ep);
                if (ep == tok.c_str() || *ep != 0) v = NAN; // a non-numeric cell: R would make the column a factor; NaN here
            }
            out[field - 2] = v;
        }
        field++;
        p = t + 1;
    }
    if (field != g->ncol + 2) { g->err = "row " + std::to_string(i) + " has " + std::to_string(field) + " fields, expected " + std::to_string(g->ncol + 2); return -2; }
    return 0;
}

extern "C" const char *rvl_gct_error_impl(const rvl_gct *g) { return g ? g->err.c_str() : "null gct handle"; }

// ---- device side ---------------------------------------------------------------------------------

// A newline that ends a non-blank line (blank_lines.skip = T): the byte before it (ignoring one '\r')
// exists and is not a newline.
struct RvlLineEnd {
    const char *text;
    __host__ __device__ bool operator()(int64_t i) const
    {
        if (text[i] != '\n') return false;
        int64_t p = i - 1;
        if (p >= 0 && text[p] == '\r') p--;
        return p >= 0 && text[p] != '\n';
    }
};

// Error bits written to *flag: 1 = a value token is not 0/1, 4 = a line has the wrong number of fields.
// One warp per data line of the chunk.  text is 16-byte aligned and padded; ends[l] = offset of the
// newline ending line l.  row = *row_base + l; rows outside [first_row, first_row + n_rows) are skipped.
// inv[c] = output sample of file column c, or -1 when that column is not wanted.
__global__ void __launch_bounds__(256)
k_gct_parse(const char *__restrict__ text, const int64_t *__restrict__ ends, const int *__restrict__ n_lines_p,
            const long long *__restrict__ row_base_p, int64_t first_row, int64_t n_rows, int64_t ncol,
            const int32_t *__restrict__ inv, int words, uint64_t *__restrict__ bits, int *__restrict__ flag)
{
    extern __shared__ unsigned long long rowbuf_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    unsigned long long *rowbuf = rowbuf_all + (size_t)warp * words;
    const int n_lines = *n_lines_p;
    const long long row_base = *row_base_p;
    for (int64_t l = (int64_t)blockIdx.x * wpb + warp; l < n_lines; l += (int64_t)gridDim.x * wpb) {
        const int64_t row = row_base + l - first_row;
        if (row < 0 || row >= n_rows) continue;
        int64_t start = l == 0 ? 0 : ends[l - 1] + 1;
        int64_t end = ends[l]; // the '\n'
        while (start < end && (text[start] == '\n' || text[start] == '\r')) start++; // skipped blank lines
        if (end > start && text[end - 1] == '\r') end--;
        for (int w = lane; w < words; w += 32) rowbuf[w] = 0ull;
        __syncwarp();
        int bad = 0;
        int64_t tabs_before = 0; // tabs in [start, chunk begin)
        const int64_t abase = start & ~(int64_t)15;
        for (int64_t c0 = abase; c0 < end; c0 += 512) {
            const int64_t a = c0 + lane * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (a < end) v = *reinterpret_cast<const uint4 *>(text + a);
            const unsigned wv[4] = {v.x, v.y, v.z, v.w};
            // tabs of this lane's 16 bytes that lie inside [start, end)
            unsigned tabmask = 0;
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const unsigned ch = (wv[b >> 2] >> ((b & 3) * 8)) & 0xffu;
                const int64_t pos = a + b;
                if (ch == '\t' && pos >= start && pos < end) tabmask |= 1u << b;
            }
            int cnt = __popc(tabmask), incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            int64_t field = tabs_before + (incl - cnt); // field index of the byte at `a`
            tabs_before += __shfl_sync(0xffffffffu, incl, 31);
            // token starts inside this lane's bytes: position `start`, or a byte that follows a tab
            unsigned prev_tab;
            {
                const unsigned pt = __shfl_up_sync(0xffffffffu, tabmask >> 15, 1);
                prev_tab = lane > 0 ? (pt & 1u) : ((c0 > start && text[c0 - 1] == '\t') ? 1u : 0u);
            }
#pragma unroll 1
            for (int b = 0; b < 16; b++) {
                const int64_t pos = a + b;
                const bool is_start = (pos == start) || (prev_tab && pos > start);
                if (is_start && pos < end && field >= 2) {
                    const int64_t col = field - 2;
                    if (col < ncol) {
                        int64_t q = pos;
                        const char c = text[q];
                        int val = c == '1' ? 1 : (c == '0' ? 0 : -1);
                        q++;
                        if (val >= 0 && q < end && text[q] == '.') {
                            q++;
                            while (q < end && text[q] == '0') q++;
                        }
                        if (val < 0 || (q < end && text[q] != '\t')) bad |= 1;
                        else {
                            const int s = inv[col];
                            if (s >= 0 && val) atomicOr(&rowbuf[s >> 6], 1ull << (s & 63));
                        }
                    }
                }
                prev_tab = (tabmask >> b) & 1u;
                field += prev_tab;
            }
        }
        // fields = tabs + 1 must be ncol + 2; a line ending in a tab has an empty last value
        if (tabs_before + 1 != ncol + 2) bad |= 4;
        else if (end > start && text[end - 1] == '\t') bad |= 1;
        __syncwarp();
        for (int w = lane; w < words; w += 32) bits[(size_t)row * words + w] = rowbuf[w];
        if (bad) atomicOr(flag, bad);
        __syncwarp();
    }
}

__global__ void k_gct_advance_rows(long long *row_base, const int *n_lines)
{
    *row_base += *n_lines;
}

// Device scratch of one ingest.
struct RvlGctScratch {
    char *d_text[2] = {nullptr, nullptr};
    char *h_stage[2] = {nullptr, nullptr};
    int64_t *d_ends = nullptr;
    int *d_nlines = nullptr;
    long long *d_rowbase = nullptr;
    int32_t *d_inv = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0, inv_cap = 0;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    cudaStream_t copy = nullptr;
    ~RvlGctScratch()
    {
        for (int i = 0; i < 2; i++) {
            if (d_text[i]) cudaFree(d_text[i]);
            if (h_stage[i]) cudaFreeHost(h_stage[i]);
            if (ev_h2d[i]) cudaEventDestroy(ev_h2d[i]);
            if (ev_done[i]) cudaEventDestroy(ev_done[i]);
        }
        if (d_ends) cudaFree(d_ends);
        if (d_nlines) cudaFree(d_nlines);
        if (d_rowbase) cudaFree(d_rowbase);
        if (d_inv) cudaFree(d_inv);
        if (d_tmp) cudaFree(d_tmp);
        if (copy) cudaStreamDestroy(copy);
    }
};

#define RVL_GCT_CHUNK ((size_t)32 << 20)

extern "C" void rvl_gct_scratch_free(void *p) { delete static_cast<RvlGctScratch *>(p); }

// File bytes [off, off + len) into the pinned staging buffer with pread() from a few threads: the kernel
// copies straight out of the page cache (no page faults on the mapping; one core moves ~5 GB/s, PCIe takes
// ten times that).  Falls back to the mapping for whatever a short read leaves.
static void parallel_read(int fd, const char *base, char *dst, size_t off, size_t len)
{
    auto rd = [=](size_t lo, size_t hi) {
        size_t at = lo;
        while (at < hi) {
            ssize_t k = pread(fd, dst + at, hi - at, (off_t)(off + at));
            if (k <= 0) { memcpy(dst + at, base + off + at, hi - at); break; }
            at += (size_t)k;
        }
    };
    const int T = 4;
    if (len < ((size_t)4 << 20)) { rd(0, len); return; }
    std::thread th[T];
    const size_t per = (len + T - 1) / T;
    for (int i = 0; i < T; i++) {
        const size_t lo = (size_t)i * per < len ? (size_t)i * per : len, hi = lo + per < len ? lo + per : len;
        th[i] = std::thread(rd, lo, hi);
    }
    for (int i = 0; i < T; i++) th[i].join();
}

#define GCK(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            snprintf(errbuf, errcap, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return -1;                                                                                     \
        }                                                                                                  \
    } while (0)

// Parse rows [first_row, first_row + n_rows) of the file into bits[n_rows][words] (device), output sample
// s taken from file column cols[s].  Returns 0, -1 (CUDA), -2 (malformed file), -3 (non-binary value).
// *launches is incremented by the number of kernels launched.
extern "C" int rvl_gct_parse_to_bits(rvl_gct *g, const int32_t *cols, int64_t n, int64_t first_row, int64_t n_rows,
                                     int words, uint64_t *d_bits, int *d_flag, cudaStream_t st, int64_t *launches,
                                     void **scratch_io, char *errbuf, size_t errcap)
{
    if (!g || !cols || n < 1 || first_row < 0 || n_rows < 0 || first_row + n_rows > g->nrow) {
        snprintf(errbuf, errcap, "bad GCT load arguments (rows [%lld, %lld) of %lld)", (long long)first_row,
                 (long long)(first_row + n_rows), g ? (long long)g->nrow : -1ll);
        return -2;
    }
    std::vector<int32_t> inv((size_t)g->ncol, -1);
    for (int64_t s = 0; s < n; s++) {
        if (cols[s] < 0 || cols[s] >= g->ncol || inv[cols[s]] >= 0) {
            snprintf(errbuf, errcap, "cols[%lld] = %d is out of range or repeated", (long long)s, (int)cols[s]);
            return -2;
        }
        inv[cols[s]] = (int32_t)s;
    }
    const size_t CH = RVL_GCT_CHUNK; // text per chunk (whole lines)
    if (!*scratch_io) *scratch_io = new RvlGctScratch();
    RvlGctScratch &S = *static_cast<RvlGctScratch *>(*scratch_io); // kept by the context across loads
    const size_t max_lines = CH / 2 + 2;
    if (!S.copy) {
        GCK(cudaStreamCreateWithFlags(&S.copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            GCK(cudaMalloc((void **)&S.d_text[i], CH + 64));
            GCK(cudaMallocHost((void **)&S.h_stage[i], CH + 64));
            GCK(cudaEventCreateWithFlags(&S.ev_h2d[i], cudaEventDisableTiming));
            GCK(cudaEventCreateWithFlags(&S.ev_done[i], cudaEventDisableTiming));
        }
        GCK(cudaMalloc((void **)&S.d_ends, sizeof(int64_t) * max_lines));
        GCK(cudaMalloc((void **)&S.d_nlines, sizeof(int)));
        GCK(cudaMalloc((void **)&S.d_rowbase, sizeof(long long)));
        RvlLineEnd pred{S.d_text[0]};
        thrust::counting_iterator<int64_t> it(0);
        GCK(cub::DeviceSelect::If(nullptr, S.tmp_bytes, it, S.d_ends, S.d_nlines, (int64_t)(CH + 1), pred, st));
        GCK(cudaMalloc(&S.d_tmp, S.tmp_bytes ? S.tmp_bytes : 1));
    }
    if (S.inv_cap < (size_t)g->ncol) {
        if (S.d_inv) cudaFree(S.d_inv);
        S.d_inv = nullptr;
        S.inv_cap = 0;
        GCK(cudaMalloc((void **)&S.d_inv, sizeof(int32_t) * (size_t)g->ncol));
        S.inv_cap = (size_t)g->ncol;
    }
    // A shard that loads its own row block ships only that block's bytes: the host line index (memchr scan, a few
    // threads) gives the byte range of rows [first_row, first_row + n_rows), and the device row counter starts at
    // first_row.  The whole file: no index needed, every line is counted on the device.
    const bool partial = (first_row > 0 || first_row + n_rows < g->nrow);
    size_t byte_lo = g->data_off, byte_hi = g->size;
    long long expect_rows = (long long)g->nrow;
    if (partial) {
        if (gct_index(g)) { snprintf(errbuf, errcap, "%s", g->err.c_str()); return -2; }
        byte_lo = n_rows > 0 ? g->line_start[(size_t)first_row] : g->size;
        byte_hi = (first_row + n_rows < g->nrow) ? g->line_start[(size_t)(first_row + n_rows)] : g->size;
        if (n_rows == 0) byte_hi = byte_lo;
        expect_rows = (long long)(first_row + n_rows);
    }
    const long long row0 = partial ? (long long)first_row : 0ll;
    GCK(cudaMemcpyAsync(S.d_rowbase, &row0, sizeof(long long), cudaMemcpyHostToDevice, st));
    GCK(cudaStreamSynchronize(st)); // row0 is a stack variable
    GCK(cudaMemcpyAsync(S.d_inv, inv.data(), sizeof(int32_t) * (size_t)g->ncol, cudaMemcpyHostToDevice, st));
    GCK(cudaMemsetAsync(d_bits, 0, sizeof(uint64_t) * (size_t)n_rows * words, st)); // rows the file lacks stay empty
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = sizeof(unsigned long long) * (size_t)words * 8;
    if (smem > 48 * 1024) GCK(cudaFuncSetAttribute(k_gct_parse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GCK(cudaEventRecord(S.ev_done[0], st));
    GCK(cudaEventRecord(S.ev_done[1], st));

    const char *p = g->base + byte_lo, *fend = g->base + byte_hi;
    int half = 0;
    while (p < fend) {
        // a chunk of whole lines: up to CH bytes, cut back to the last newline
        size_t len = (size_t)(fend - p);
        bool add_nl = false;
        if (len > CH) {
            const char *cut = (const char *)memrchr(p, '\n', CH);
            if (!cut) { snprintf(errbuf, errcap, "a line is longer than %zu bytes", CH); return -2; }
            len = (size_t)(cut - p) + 1;
        } else if (fend[-1] != '\n') add_nl = true; // last line without a newline
        GCK(cudaEventSynchronize(S.ev_done[half])); // the staging half is free again
        parallel_read(g->fd, g->base, S.h_stage[half], (size_t)(p - g->base), len);
        if (add_nl) S.h_stage[half][len] = '\n';
        const size_t tot = len + (add_nl ? 1 : 0);
        memset(S.h_stage[half] + tot, '\n', 32); // padding the 16-byte loads may touch
        GCK(cudaMemcpyAsync(S.d_text[half], S.h_stage[half], tot + 32, cudaMemcpyHostToDevice, S.copy));
        GCK(cudaEventRecord(S.ev_h2d[half], S.copy));
        GCK(cudaStreamWaitEvent(st, S.ev_h2d[half], 0));
        RvlLineEnd pred{S.d_text[half]};
        thrust::counting_iterator<int64_t> it(0);
        size_t tb = S.tmp_bytes;
        GCK(cub::DeviceSelect::If(S.d_tmp, tb, it, S.d_ends, S.d_nlines, (int64_t)tot, pred, st));
        k_gct_parse<<<sms * 4, 256, smem, st>>>(S.d_text[half], S.d_ends, S.d_nlines, S.d_rowbase, first_row, n_rows,
                                                 g->ncol, S.d_inv, words, d_bits, d_flag);
        k_gct_advance_rows<<<1, 1, 0, st>>>(S.d_rowbase, S.d_nlines);
        GCK(cudaGetLastError());
        GCK(cudaEventRecord(S.ev_done[half], st));
        GCK(cudaStreamWaitEvent(S.copy, S.ev_done[half], 0));
        if (launches) *launches += 3;
        p += len;
        half ^= 1;
    }
    long long rows_seen = 0;
    int flag = 0;
    GCK(cudaMemcpyAsync(&rows_seen, S.d_rowbase, sizeof rows_seen, cudaMemcpyDeviceToHost, st));
    GCK(cudaMemcpyAsync(&flag, d_flag, sizeof flag, cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    GCK(cudaStreamSynchronize(S.copy));
    if (rows_seen != expect_rows) {
        snprintf(errbuf, errcap, "file has %lld data lines, header says %lld", rows_seen, (long long)g->nrow);
        return -2;
    }
    if (flag & 4) { snprintf(errbuf, errcap, "a data line does not have %lld fields", (long long)(g->ncol + 2)); return -2; }
    if (flag & 1) { snprintf(errbuf, errcap, "feature matrix has an entry that is not 0 or 1 (binary features only; no fallback)"); return -3; }
    return 0;
}
