// pdq_io.cpp -- count-matrix ingestion: a CSV of read counts parsed by host threads straight into the
// (samples, genes) int64 layout the hot path consumes (dds.py:245-249), e.g. into a page-locked buffer that is uploaded at
// full PCIe rate.  The reference loads the same files with pandas (`pd.read_csv(..., index_col=0).T`,
// examples/plot_pandas_io_example.py:57-66): header line = column labels, first field of every line = row label, the rest
// non-negative integers; fields may be double-quoted.  Host code only (no CUDA): it is part of the C ABI library so that the
// product has one native artefact.
#include <fcntl.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pydeseq2_b200.h"

namespace {

struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        n = (size_t)st.st_size;
        if (n == 0) return true;
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        p = (const char*)m;
        madvise(m, n, MADV_SEQUENTIAL);
        return true;
    }
    ~Mapped() {
        if (p) munmap((void*)p, n);
        if (fd >= 0) close(fd);
    }
};

// line starts (offsets) of the file; a trailing newline does not open a new line; '\r' before '\n' is dropped by the field parser
std::vector<size_t> line_starts(const Mapped& f) {
    std::vector<size_t> ls;
    size_t pos = 0;
    while (pos < f.n) {
        ls.push_back(pos);
        const void* nl = memchr(f.p + pos, '\n', f.n - pos);
        if (!nl) break;
        pos = (size_t)((const char*)nl - f.p) + 1;
    }
    ls.push_back(f.n);  // sentinel: end of the last line
    return ls;
}

// next field of [p, end): returns the field's [b, e) with surrounding quotes / blanks / '\r' stripped, advances p past the separator
inline bool next_field(const char*& p, const char* end, char sep, const char*& b, const char*& e) {
    if (p > end) return false;
    const char* q = p;
    if (q < end && *q == '"') {  // quoted label: up to the closing quote ("" inside a label is not expected in count tables)
        ++q;
        b = q;
        while (q < end && *q != '"') ++q;
        e = q;
        while (q < end && *q != sep) ++q;
    } else {
        b = q;
        while (q < end && *q != sep) ++q;
        e = q;
        while (e > b && (e[-1] == '\r' || e[-1] == ' ' || e[-1] == '\n')) --e;
        while (b < e && *b == ' ') ++b;
    }
    p = q + 1;  // past the separator (or past end + 1 after the last field)
    return true;
}

// a read count: non-negative integer, optionally written as a float with zero fraction ("12.0", "1e3"); -1 on anything else
inline int64_t parse_count(const char* b, const char* e) {
    if (b >= e) return -1;
    int64_t v = 0;
    const char* q = b;
    while (q < e && *q >= '0' && *q <= '9') {
        v = v * 10 + (*q - '0');
        if (v < 0) return -1;
        ++q;
    }
    if (q == e) return q == b ? -1 : v;
    char tmp[64];
    const size_t n = (size_t)(e - b);
    if (n >= sizeof tmp) return -1;
    memcpy(tmp, b, n);
    tmp[n] = 0;
    char* stop = nullptr;
    const double d = strtod(tmp, &stop);
    if (stop != tmp + n || !(d >= 0.0) || d > 9.2e18 || d != floor(d)) return -1;
    return (int64_t)d;
}

}  // namespace

extern "C" int pdq_csv_scan(const char* path, char sep, int64_t* n_rows, int64_t* n_cols, size_t* label_bytes) {
    if (!path || !n_rows || !n_cols) return PDQ_ERR_INVALID;
    Mapped f;
    if (!f.open(path)) return PDQ_ERR_INVALID;
    const std::vector<size_t> ls = line_starts(f);
    int64_t rows = 0;
    for (size_t i = 1; i + 1 < ls.size(); ++i) {  // data lines that are not blank
        const char *b = f.p + ls[i], *e = f.p + ls[i + 1];
        while (e > b && (e[-1] == '\n' || e[-1] == '\r' || e[-1] == ' ')) --e;
        if (e > b) ++rows;
    }
    int64_t cols = 0;
    size_t bytes = 0;
    if (ls.size() >= 2) {
        const char *p = f.p + ls[0], *end = f.p + ls[1];
        while (end > p && (end[-1] == '\n' || end[-1] == '\r')) --end;
        const char *b, *e;
        bool first = true;
        while (p <= end && next_field(p, end, sep, b, e)) {
            if (!first) {
                ++cols;
                bytes += (size_t)(e - b) + 1;
            }
            first = false;
        }
    }
    // row labels: bounded by the length of the data lines
    for (size_t i = 1; i + 1 < ls.size(); ++i) {
        const char *p = f.p + ls[i], *end = f.p + ls[i + 1], *b, *e;
        if (next_field(p, end, sep, b, e)) bytes += (size_t)(e - b) + 1;
    }
    *n_rows = rows;
    *n_cols = cols;
    if (label_bytes) *label_bytes = bytes + 2;
    return PDQ_OK;
}

// out[(c * ld_out) + r] when transpose (file rows become matrix columns -- genes in rows -> (samples, genes)), else out[r * ld_out + c].
// labels: '\n'-separated column labels, then a '\0', then '\n'-separated row labels, then a '\0'.
extern "C" int pdq_csv_read_counts(const char* path, char sep, int transpose, int64_t* out, int64_t ld_out, int64_t n_rows,
                                   int64_t n_cols, char* labels, size_t label_cap, int threads, int64_t* bad_row, int64_t* bad_col) {
    if (!path || !out || n_rows < 0 || n_cols < 0) return PDQ_ERR_INVALID;
    Mapped f;
    if (!f.open(path)) return PDQ_ERR_INVALID;
    const std::vector<size_t> ls = line_starts(f);
    std::vector<size_t> data;  // indices of the non-blank data lines
    for (size_t i = 1; i + 1 < ls.size(); ++i) {
        const char *b = f.p + ls[i], *e = f.p + ls[i + 1];
        while (e > b && (e[-1] == '\n' || e[-1] == '\r' || e[-1] == ' ')) --e;
        if (e > b) data.push_back(i);
    }
    if ((int64_t)data.size() != n_rows) return PDQ_ERR_INVALID;
    size_t lp = 0;
    auto put = [&](const char* b, const char* e, char term) {
        if (!labels) return;
        const size_t n = (size_t)(e - b);
        if (lp + n + 1 >= label_cap) return;
        memcpy(labels + lp, b, n);
        lp += n;
        labels[lp++] = term;
    };
    if (ls.size() >= 2) {  // column labels
        const char *p = f.p + ls[0], *end = f.p + ls[1], *b, *e;
        while (end > p && (end[-1] == '\n' || end[-1] == '\r')) --end;
        bool first = true;
        int64_t c = 0;
        while (p <= end && next_field(p, end, sep, b, e)) {
            if (!first) {
                put(b, e, '\n');
                ++c;
            }
            first = false;
        }
        if (c != n_cols) return PDQ_ERR_INVALID;
    }
    if (labels && lp < label_cap) labels[lp++] = 0;
    std::vector<std::pair<const char*, const char*>> row_label((size_t)n_rows);
    std::atomic<int64_t> err_row(-1), err_col(-1);
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt > 32) nt = 32;
    if (nt < 1) nt = 1;
    if ((int64_t)nt > n_rows) nt = n_rows > 0 ? (int)n_rows : 1;
    auto work = [&](int t) {
        const int64_t lo = n_rows * t / nt, hi = n_rows * (t + 1) / nt;
        for (int64_t r = lo; r < hi; ++r) {
            const char *p = f.p + ls[data[(size_t)r]], *end = f.p + ls[data[(size_t)r] + 1], *b, *e;
            while (end > p && (end[-1] == '\n' || end[-1] == '\r')) --end;
            if (!next_field(p, end, sep, b, e)) continue;
            row_label[(size_t)r] = {b, e};
            int64_t c = 0;
            for (; c < n_cols && p <= end; ++c) {
                next_field(p, end, sep, b, e);
                const int64_t v = parse_count(b, e);
                if (v < 0) {
                    int64_t exp = -1;
                    if (err_row.compare_exchange_strong(exp, r)) err_col = c;
                    return;
                }
                if (transpose) out[c * ld_out + r] = v;
                else out[r * ld_out + c] = v;
            }
            if (c != n_cols || p <= end) {  // too few or too many fields on the line
                int64_t exp = -1;
                if (err_row.compare_exchange_strong(exp, r)) err_col = c;
                return;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    if (err_row.load() >= 0) {
        if (bad_row) *bad_row = err_row.load();
        if (bad_col) *bad_col = err_col.load();
        return PDQ_ERR_INVALID;
    }
    for (int64_t r = 0; r < n_rows; ++r) put(row_label[(size_t)r].first, row_label[(size_t)r].second, '\n');
    if (labels && lp < label_cap) labels[lp++] = 0;
    return PDQ_OK;
}
