// Multi-threaded parser for probtrackx `fdt_matrix2.dot` text (host code, part of libnfact_b200.so).
//
// Replaces the `np.loadtxt` call of NFACT/base/matrix_handling.py:93-113 (single-threaded, minutes per
// subject).  Every line is `row col value` separated by blanks; numbers are converted exactly -- plain decimals on
// a one-operation fast path, everything else with std::from_chars(double), both correctly rounded like strtod /
// Python float() --, so the float64 triplets -- and therefore the bit-exact ingest downstream -- are identical to
// what np.loadtxt returns.
#include <algorithm>
#include <charconv>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nfact_b200.h"

namespace nfb {
void set_error(const std::string& msg);
}

namespace {

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// a leading '+' is accepted by loadtxt (float("+5")) but not by from_chars; "+-5" is an error for both
inline bool plus_sign_ok(const char* p, const char* end) {
    if (p + 1 >= end) return false;
    const char c = p[1];
    return (c >= '0' && c <= '9') || c == '.' || c == 'i' || c == 'I' || c == 'n' || c == 'N';
}

inline const char* skip_comment(const char* p, const char* end) {   // from a '#' to the end of its line
    while (p < end && *p != '\n') ++p;
    return p;
}

// number of lines in [begin, end) that hold data: np.loadtxt drops everything from a '#' to the end of the line
// (comments='#') and skips lines that are empty afterwards.  A line holds data when its first non-blank character is
// neither the end of the line nor a '#'; the line ends are found with memchr.
int64_t count_lines(const char* begin, const char* end) {
    int64_t lines = 0;
    const char* p = begin;
    while (p < end) {
        const char* nl = static_cast<const char*>(memchr(p, '\n', static_cast<size_t>(end - p)));
        const char* stop = nl ? nl : end;
        const char* q = p;
        while (q < stop && is_blank(*q)) ++q;
        lines += (q < stop && *q != '#');
        p = stop + 1;
    }
    return lines;
}

// Powers of ten that are exact in float64.
const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// One number at p (no leading blanks).  Plain decimals -- [+-]digits[.digits] with at most 19 significant digits, an
// integer value of the digit string <= 2^53 and at most 22 fraction digits -- convert with ONE correctly rounded IEEE
// operation on two exact operands (Clinger's fast path), which is the correctly rounded value strtod / Python float()
// / np.loadtxt produce; that covers streamline counts and printf's %f / %g output without exponent.  Everything else
// (exponents, long digit strings, inf / nan, errors) goes to std::from_chars, which is correctly rounded as well.
inline bool parse_field(const char*& p, const char* end, double& out) {
    const char* q = p;
    if (q < end && *q == '+') {
        if (!plus_sign_ok(q, end)) return false;
        ++q;
    }
    const char* start = q;
    bool neg = false;
    if (q < end && *q == '-') { neg = true; ++q; }
    uint64_t mant = 0;
    int ndig = 0, frac = 0;
    bool any = false, fast = true;
    while (q < end && static_cast<unsigned>(*q - '0') <= 9u) {
        any = true;
        if (ndig >= 19) { fast = false; break; }
        mant = mant * 10u + static_cast<unsigned>(*q - '0');
        ndig += (mant != 0);
        ++q;
    }
    if (fast && q < end && *q == '.') {
        ++q;
        while (q < end && static_cast<unsigned>(*q - '0') <= 9u) {
            any = true;
            if (ndig >= 19) { fast = false; break; }
            mant = mant * 10u + static_cast<unsigned>(*q - '0');
            ndig += (mant != 0);
            ++frac;
            ++q;
        }
    }
    if (fast && any && !(q < end && (*q == 'e' || *q == 'E')) && mant <= (1ull << 53) && frac <= 22) {
        double v = static_cast<double>(mant);
        if (frac) v /= kPow10[frac];
        out = neg ? -v : v;
        p = q;
        return true;
    }
    auto res = std::from_chars(start, end, out);
    if (res.ec != std::errc()) return false;
    p = res.ptr;
    return true;
}

// parse the lines of [begin, end) into out arrays starting at index `first`; returns lines parsed or -1
int64_t parse_range(const char* begin, const char* end, int64_t first, double* rows, double* cols, double* vals) {
    int64_t idx = first;
    const char* p = begin;
    while (p < end) {
        while (p < end && (is_blank(*p) || *p == '\n' || *p == '#')) p = *p == '#' ? skip_comment(p, end) : p + 1;
        if (p >= end) break;
        double field[3];
        for (int f = 0; f < 3; ++f) {
            while (p < end && is_blank(*p)) ++p;
            if (!parse_field(p, end, field[f])) return -1;
        }
        while (p < end && is_blank(*p)) ++p;
        if (p < end && *p == '#') p = skip_comment(p, end);
        if (p < end && *p != '\n') return -1;  // more than three columns
        rows[idx] = field[0];
        cols[idx] = field[1];
        vals[idx] = field[2];
        ++idx;
    }
    return idx - first;
}

// same lines, but straight into what the ingest kernels take: 0-based int32 row / column (the reference's
// `rows - 1`, truncated like its integer cast) and the float64 value; line `last` (the final line of the file)
// is the matrix shape and goes to shape[3] instead.  Tracks the index range for the bounds check.
struct CooStats {
    int64_t parsed = 0;
    double rmin = 1e300, rmax = -1e300, cmin = 1e300, cmax = -1e300;
    bool bad_index = false;
};

CooStats parse_range_coo(const char* begin, const char* end, int64_t first, int64_t last, int32_t* rows,
                         int32_t* cols, double* vals, double* shape) {
    CooStats st;
    int64_t idx = first;
    const char* p = begin;
    while (p < end) {
        while (p < end && (is_blank(*p) || *p == '\n' || *p == '#')) p = *p == '#' ? skip_comment(p, end) : p + 1;
        if (p >= end) break;
        double field[3];
        for (int f = 0; f < 3; ++f) {
            while (p < end && is_blank(*p)) ++p;
            if (!parse_field(p, end, field[f])) { st.parsed = -1; return st; }
        }
        while (p < end && is_blank(*p)) ++p;
        if (p < end && *p == '#') p = skip_comment(p, end);
        if (p < end && *p != '\n') { st.parsed = -1; return st; }
        if (idx == last) {
            shape[0] = field[0]; shape[1] = field[1]; shape[2] = field[2];
        } else {
            const double r = field[0] - 1.0, c = field[1] - 1.0;
            // NaN or beyond int32: reported as a bad index (the reference fails inside scipy's index cast)
            if (!(r > -2147483648.0 && r < 2147483648.0 && c > -2147483648.0 && c < 2147483648.0)) st.bad_index = true;
            else {
                rows[idx] = static_cast<int32_t>(r);
                cols[idx] = static_cast<int32_t>(c);
            }
            vals[idx] = field[2];
            st.rmin = std::min(st.rmin, r); st.rmax = std::max(st.rmax, r);
            st.cmin = std::min(st.cmin, c); st.cmax = std::max(st.cmax, c);
        }
        ++idx;
    }
    st.parsed = idx - first;
    return st;
}

std::vector<const char*> chunk_starts(const char* buf, int64_t len, int nthreads) {
    std::vector<const char*> starts;
    starts.push_back(buf);
    for (int t = 1; t < nthreads; ++t) {
        const char* p = buf + len * t / nthreads;
        const char* nl = static_cast<const char*>(memchr(p, '\n', static_cast<size_t>(buf + len - p)));
        const char* s = nl ? nl + 1 : buf + len;
        if (s > starts.back()) starts.push_back(s);
    }
    starts.push_back(buf + len);
    return starts;
}

int pick_threads(int64_t len, int nthreads) {
    if (nthreads <= 0) nthreads = static_cast<int>(std::thread::hardware_concurrency());
    nthreads = std::max(1, std::min(nthreads, 64));
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(nthreads, len / (1 << 20) + 1)));
}

}  // namespace

extern "C" {

int nfb_dot_count_lines(const char* buf, int64_t len, int nthreads, int64_t* n_lines_out) {
    if (buf == nullptr || n_lines_out == nullptr || len < 0) {
        nfb::set_error("nfb_dot_count_lines: bad arguments");
        return 2;
    }
    nthreads = pick_threads(len, nthreads);
    auto starts = chunk_starts(buf, len, nthreads);
    std::vector<int64_t> counts(starts.size() - 1, 0);
    std::vector<std::thread> pool;
    for (size_t c = 0; c + 1 < starts.size(); ++c)
        pool.emplace_back([&, c] { counts[c] = count_lines(starts[c], starts[c + 1]); });
    for (auto& th : pool) th.join();
    int64_t total = 0;
    for (int64_t v : counts) total += v;
    *n_lines_out = total;
    return 0;
}

int nfb_dot_parse(const char* buf, int64_t len, int nthreads, int64_t n_lines, double* rows, double* cols,
                  double* vals) {
    if (buf == nullptr || rows == nullptr || cols == nullptr || vals == nullptr || len < 0) {
        nfb::set_error("nfb_dot_parse: bad arguments");
        return 2;
    }
    nthreads = pick_threads(len, nthreads);
    auto starts = chunk_starts(buf, len, nthreads);
    const size_t nchunks = starts.size() - 1;
    std::vector<int64_t> counts(nchunks, 0), parsed(nchunks, 0);
    {
        std::vector<std::thread> pool;
        for (size_t c = 0; c < nchunks; ++c)
            pool.emplace_back([&, c] { counts[c] = count_lines(starts[c], starts[c + 1]); });
        for (auto& th : pool) th.join();
    }
    std::vector<int64_t> first(nchunks + 1, 0);
    for (size_t c = 0; c < nchunks; ++c) first[c + 1] = first[c] + counts[c];
    if (first[nchunks] != n_lines) {
        nfb::set_error("nfb_dot_parse: line count changed between the two passes");
        return 2;
    }
    {
        std::vector<std::thread> pool;
        for (size_t c = 0; c < nchunks; ++c)
            pool.emplace_back([&, c] { parsed[c] = parse_range(starts[c], starts[c + 1], first[c], rows, cols, vals); });
        for (auto& th : pool) th.join();
    }
    for (size_t c = 0; c < nchunks; ++c)
        if (parsed[c] != counts[c]) {
            nfb::set_error("could not convert string to float: malformed line in fdt_matrix2.dot (expected 3 columns)");
            return 2;
        }
    return 0;
}

int nfb_dot_parse_coo(const char* buf, int64_t len, int nthreads, int64_t n_lines, int32_t* rows0, int32_t* cols0,
                      double* vals, int64_t* shape_out) {
    if (buf == nullptr || shape_out == nullptr || len < 0 || n_lines < 1 ||
        (n_lines > 1 && (rows0 == nullptr || cols0 == nullptr || vals == nullptr))) {
        nfb::set_error("nfb_dot_parse_coo: bad arguments");
        return 2;
    }
    nthreads = pick_threads(len, nthreads);
    auto starts = chunk_starts(buf, len, nthreads);
    const size_t nchunks = starts.size() - 1;
    std::vector<int64_t> counts(nchunks, 0);
    {
        std::vector<std::thread> pool;
        for (size_t c = 0; c < nchunks; ++c)
            pool.emplace_back([&, c] { counts[c] = count_lines(starts[c], starts[c + 1]); });
        for (auto& th : pool) th.join();
    }
    std::vector<int64_t> first(nchunks + 1, 0);
    for (size_t c = 0; c < nchunks; ++c) first[c + 1] = first[c] + counts[c];
    if (first[nchunks] != n_lines) {
        nfb::set_error("nfb_dot_parse_coo: line count changed between the two passes");
        return 2;
    }
    double shape[3] = {0.0, 0.0, 0.0};
    std::vector<CooStats> stats(nchunks);
    {
        std::vector<std::thread> pool;
        for (size_t c = 0; c < nchunks; ++c)
            pool.emplace_back([&, c] {
                stats[c] = parse_range_coo(starts[c], starts[c + 1], first[c], n_lines - 1, rows0, cols0, vals, shape);
            });
        for (auto& th : pool) th.join();
    }
    CooStats all;
    for (size_t c = 0; c < nchunks; ++c) {
        if (stats[c].parsed != counts[c]) {
            nfb::set_error("could not convert string to float: malformed line in fdt_matrix2.dot (expected 3 columns)");
            return 2;
        }
        all.bad_index |= stats[c].bad_index;
        all.rmin = std::min(all.rmin, stats[c].rmin); all.rmax = std::max(all.rmax, stats[c].rmax);
        all.cmin = std::min(all.cmin, stats[c].cmin); all.cmax = std::max(all.cmax, stats[c].cmax);
    }
    const int64_t nrows = static_cast<int64_t>(shape[0]), ncols = static_cast<int64_t>(shape[1]);
    shape_out[0] = nrows;
    shape_out[1] = ncols;
    if (n_lines > 1) {   // the messages scipy's coo/csc constructor raises
        if (all.bad_index) { nfb::set_error("row / column index is not a 32-bit integer"); return 2; }
        if (all.rmin <= -1.0) { nfb::set_error("negative row index found"); return 2; }
        if (all.cmin <= -1.0) { nfb::set_error("negative column index found"); return 2; }
        if (static_cast<int64_t>(all.rmax) >= nrows) { nfb::set_error("row index exceeds matrix dimensions"); return 2; }
        if (static_cast<int64_t>(all.cmax) >= ncols) { nfb::set_error("column index exceeds matrix dimensions"); return 2; }
    }
    return 0;
}

}  // extern "C"
