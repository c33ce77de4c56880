// fdt_matrix2.dot text -> COO triplets, on the device.
//
// Replaces np.loadtxt + the index arithmetic of load_single_matrix (NFACT/base/matrix_handling.py:93-137) for the
// resident ingest: the file's bytes are uploaded once (they are about as large as the triplets they encode) and
// parsed by one thread per 128-byte segment, instead of ~9 M lines/s per host thread.
//
// Every line is `row col value` separated by blanks; the last non-empty line holds the matrix shape.  A thread owns
// the lines that START in its segment.  Pass 1 counts them (a line counts when it holds a non-blank character, like
// the host parser and loadtxt), a scan turns the counts into line indices, pass 2 parses each line into
//     rows0[i] = int32(trunc(field0 - 1)), cols0[i] = int32(trunc(field1 - 1)), vals[i] = field2     (float64)
// and the last line into shape[3].  Like loadtxt (comments='#'), everything from a '#' to the end of its line is
// ignored and lines that hold nothing else are skipped.
//
// Exactness.  The ingest downstream is bit-exact against the reference, so a number must convert exactly as
// strtod / Python float() / std::from_chars do.  The parser handles the decimal forms whose conversion is a single
// correctly rounded IEEE operation: at most 19 significant digits collected exactly into a 64-bit integer that is
// <= 2^53, times or divided by a power of ten 10^|e| with |e| <= 22 (both operands exact => the one rounding of
// the multiplication / division is the correct rounding of the decimal value -- Clinger's fast path).  That covers
// integer streamline counts and the %g / %f output of probtrackx.  Anything else (more digits, larger exponents,
// inf / nan) raises the `hard` flag and the caller parses the file on the host instead; a malformed line raises
// `malformed` (loadtxt's ValueError).
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace nfb {
namespace {

constexpr int kSeg = 128;          // bytes per thread
constexpr int kParseThreads = 256;

__device__ __forceinline__ bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }
__device__ __forceinline__ bool is_digit(char c) { return c >= '0' && c <= '9'; }

__constant__ double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

enum : int { kOk = 0, kMalformed = 1, kHard = 2 };

// one decimal number at p (no leading blanks); advances p; returns kOk / kMalformed / kHard
__device__ int parse_number(const char*& p, const char* end, double& out) {
    if (p < end && *p == '+') {                        // loadtxt accepts a leading '+' (not "+-5"), from_chars none
        if (p + 1 >= end) return kMalformed;
        const char c = p[1];
        if (c == 'i' || c == 'I' || c == 'n' || c == 'N') return kHard;
        if (!(is_digit(c) || c == '.')) return kMalformed;
        ++p;
    }
    bool neg = false;
    if (p < end && *p == '-') { neg = true; ++p; }
    if (p < end && (*p == 'i' || *p == 'I' || *p == 'n' || *p == 'N')) return kHard;   // inf / nan: host parser
    unsigned long long mant = 0;
    int ndig = 0, dropped = 0, frac = 0;
    bool any = false, nonzero_dropped = false;
    while (p < end && is_digit(*p)) {
        any = true;
        if (ndig < 19) { mant = mant * 10ull + static_cast<unsigned>(*p - '0'); ndig += (mant != 0); }
        else { ++dropped; nonzero_dropped |= (*p != '0'); }
        ++p;
    }
    if (p < end && *p == '.') {
        ++p;
        while (p < end && is_digit(*p)) {
            any = true;
            if (ndig < 19) { mant = mant * 10ull + static_cast<unsigned>(*p - '0'); ndig += (mant != 0); ++frac; }
            else { nonzero_dropped |= (*p != '0'); }
            ++p;
        }
    }
    if (!any) return kMalformed;
    int e10 = 0;
    if (p < end && (*p == 'e' || *p == 'E')) {
        const char* q = p + 1;
        bool eneg = false;
        if (q < end && (*q == '+' || *q == '-')) { eneg = (*q == '-'); ++q; }
        if (q < end && is_digit(*q)) {                 // otherwise the 'e' is not part of the number
            int ev = 0;
            while (q < end && is_digit(*q)) { ev = min(ev * 10 + (*q - '0'), 100000); ++q; }
            e10 = eneg ? -ev : ev;
            p = q;
        }
    }
    if (mant == 0 && !nonzero_dropped) { out = neg ? -0.0 : 0.0; return kOk; }
    if (nonzero_dropped || mant > (1ull << 53)) return kHard;
    e10 += dropped - frac;
    double v = static_cast<double>(mant);              // exact: mant <= 2^53
    if (e10 < -22 || e10 > 22) return kHard;
    v = e10 >= 0 ? __dmul_rn(v, kPow10[e10]) : __ddiv_rn(v, kPow10[-e10]);
    out = neg ? -v : v;
    return kOk;
}

// does a line starting at p hold data?  The line ends at '\n' or at the end of the text; loadtxt (comments='#')
// drops everything from a '#' on, so a line of blanks and / or a comment holds none
__device__ __forceinline__ bool line_has_content(const char* p, const char* end) {
    for (; p < end && *p != '\n' && *p != '#'; ++p)
        if (!is_blank(*p)) return true;
    return false;
}

__global__ void __launch_bounds__(kParseThreads)
dot_count_kernel(const char* __restrict__ txt, int64_t len, unsigned char* __restrict__ seg_lines,
                 unsigned int* __restrict__ block_lines) {
    __shared__ unsigned int warp_tot[kParseThreads / 32];
    const int64_t seg = static_cast<int64_t>(blockIdx.x) * kParseThreads + threadIdx.x;
    const int64_t p0 = seg * kSeg;
    unsigned int count = 0;
    if (p0 < len) {
        const char* end = txt + len;
        const int64_t p1 = min(p0 + kSeg, len);
        for (int64_t p = p0; p < p1; ++p)
            if ((p == 0 || txt[p - 1] == '\n') && line_has_content(txt + p, end)) ++count;
        seg_lines[seg] = static_cast<unsigned char>(count);
    }
    unsigned int w = count;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = 0;
        for (int i = 0; i < kParseThreads / 32; ++i) t += warp_tot[i];
        block_lines[blockIdx.x] = t;
    }
}

// exclusive scan of the per-block line counts (one CTA; a 6.6 GB file has ~200 k blocks); total -> *total_out
__global__ void __launch_bounds__(1024)
dot_scan_kernel(const unsigned int* __restrict__ block_lines, int64_t nblocks, long long* __restrict__ block_first,
                long long* __restrict__ total_out) {
    __shared__ long long warp_sum_s[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nblocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const long long v = i < nblocks ? block_lines[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sum_s[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            long long ws = warp_sum_s[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, ws, o);
                if (threadIdx.x >= o) ws += t;
            }
            warp_sum_s[threadIdx.x] = ws;              // inclusive over warps
        }
        __syncthreads();
        const long long warp_off = (threadIdx.x >> 5) ? warp_sum_s[(threadIdx.x >> 5) - 1] : 0;
        if (i < nblocks) block_first[i] = carry + warp_off + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_sum_s[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

// flags: [0] malformed line, [1] number outside the exact fast path, [2] index not a 32-bit integer
// range: [0] min row, [1] max row, [2] min col, [3] max col (0-based, after truncation)
__global__ void __launch_bounds__(kParseThreads)
dot_parse_kernel(const char* __restrict__ txt, int64_t len, const unsigned char* __restrict__ seg_lines,
                 const long long* __restrict__ block_first, long long n_lines, int* __restrict__ rows0,
                 int* __restrict__ cols0, double* __restrict__ vals, double* __restrict__ shape,
                 int* __restrict__ flags, int* __restrict__ range) {
    __shared__ unsigned int warp_tot[kParseThreads / 32];
    const int64_t seg = static_cast<int64_t>(blockIdx.x) * kParseThreads + threadIdx.x;
    const int64_t p0 = seg * kSeg;
    const unsigned int mine = p0 < len ? seg_lines[seg] : 0u;
    // exclusive scan of the per-thread line counts inside the block
    unsigned int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned int warp_off = 0;
    for (int i = 0; i < (threadIdx.x >> 5); ++i) warp_off += warp_tot[i];
    if (mine == 0) return;
    long long idx = block_first[blockIdx.x] + warp_off + incl - mine;
    const char* end = txt + len;
    const int64_t p1 = min(p0 + kSeg, len);
    int rmin = 2147483647, rmax = -2147483647 - 1, cmin = 2147483647, cmax = -2147483647 - 1;
    for (int64_t pos = p0; pos < p1; ++pos) {
        if (!(pos == 0 || txt[pos - 1] == '\n')) continue;
        const char* p = txt + pos;
        if (!line_has_content(p, end)) continue;
        double field[3];
        int status = kOk;
        for (int f = 0; f < 3 && status == kOk; ++f) {
            while (p < end && is_blank(*p)) ++p;
            status = parse_number(p, end, field[f]);
        }
        if (status == kOk) {
            while (p < end && is_blank(*p)) ++p;
            if (p < end && *p != '\n' && *p != '#') status = kMalformed;   // more than three columns / garbage
        }
        if (status == kMalformed) flags[0] = 1;
        else if (status == kHard) flags[1] = 1;
        else if (idx == n_lines - 1) {
            shape[0] = field[0]; shape[1] = field[1]; shape[2] = field[2];
        } else {
            const double r = field[0] - 1.0, c = field[1] - 1.0;
            if (!(r > -2147483648.0 && r < 2147483648.0 && c > -2147483648.0 && c < 2147483648.0)) {
                flags[2] = 1;
            } else {
                const int ri = static_cast<int>(r), ci = static_cast<int>(c);
                rows0[idx] = ri;
                cols0[idx] = ci;
                rmin = min(rmin, ri); rmax = max(rmax, ri);
                cmin = min(cmin, ci); cmax = max(cmax, ci);
            }
            vals[idx] = field[2];
        }
        ++idx;
    }
    if (rmin <= rmax) {
        atomicMin(&range[0], rmin); atomicMax(&range[1], rmax);
        atomicMin(&range[2], cmin); atomicMax(&range[3], cmax);
    }
}

__global__ void dot_init_range_kernel(int* range) {
    range[0] = 2147483647; range[1] = -2147483647 - 1; range[2] = 2147483647; range[3] = -2147483647 - 1;
}

}  // namespace

int64_t dot_gpu_segments(int64_t len) { return ceil_div64(len, kSeg); }
int64_t dot_gpu_blocks(int64_t len) { return ceil_div64(dot_gpu_segments(len), kParseThreads); }

int dot_gpu_count(const char* txt, int64_t len, unsigned char* seg_lines, unsigned int* block_lines,
                  long long* block_first, long long* total_out, cudaStream_t stream) {
    const int64_t nblocks = dot_gpu_blocks(len);
    if (nblocks == 0) {
        NFB_CUDA(cudaMemsetAsync(total_out, 0, sizeof(long long), stream));
        return 0;
    }
    dot_count_kernel<<<static_cast<unsigned>(nblocks), kParseThreads, 0, stream>>>(txt, len, seg_lines, block_lines);
    NFB_CUDA(cudaGetLastError());
    dot_scan_kernel<<<1, 1024, 0, stream>>>(block_lines, nblocks, block_first, total_out);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

int dot_gpu_parse(const char* txt, int64_t len, const unsigned char* seg_lines, const long long* block_first,
                  long long n_lines, int* rows0, int* cols0, double* vals, double* shape, int* flags, int* range,
                  cudaStream_t stream) {
    const int64_t nblocks = dot_gpu_blocks(len);
    NFB_CUDA(cudaMemsetAsync(flags, 0, 3 * sizeof(int), stream));
    dot_init_range_kernel<<<1, 1, 0, stream>>>(range);
    if (nblocks == 0) return 0;
    dot_parse_kernel<<<static_cast<unsigned>(nblocks), kParseThreads, 0, stream>>>(
        txt, len, seg_lines, block_first, n_lines, rows0, cols0, vals, shape, flags, range);
    NFB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace nfb
