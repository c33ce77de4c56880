// LZ4 frame decoder for `fdt_matrix2.dot.lz4` (host code, part of libnfact_b200.so).
//
// The reference reads compressed matrices with `lz4.frame.open(matfile, "rb")`
// (NFACT/base/matrix_handling.py:108-110; the files are written by nfact_config,
// NFACT/config/nfact_config_functions.py:601-604, with the python-lz4 defaults).  python-lz4 is an
// un-vendored, unpinned dependency (pyproject.toml:27) that is absent from this image, so this file restates the
// published format it implements -- LZ4 Frame Format 1.6.x and LZ4 Block Format -- from the specification:
//   frame  = magic 0x184D2204 | FLG | BD | [content size u64] | [dict id u32] | HC | blocks... | EndMark 0 |
//            [content checksum u32]            (skippable frames 0x184D2A5x and concatenated frames allowed)
//   block  = u32 size (bit 31: stored uncompressed) | data | [block checksum u32]
//   data   = sequences: token (hi nibble literal length, lo nibble match length - 4, 15 = more length bytes of
//            255...) | literals | u16 offset | ...; the last sequence ends after its literals
// Checksums are XXH32 (seed 0); HC = (XXH32(descriptor) >> 8) & 0xff.  All of them are verified, like LZ4F does.
//
// Frames with independent blocks (the `lz4` command line default) are decoded by all host threads: one pass
// measures every block's decoded size (token walk, no copies), a prefix sum places the blocks, a second pass
// decodes them in place.  Linked blocks (python-lz4's default) may reference the previous 64 KB of output and
// are decoded in order into the same contiguous buffer (optionally in side-by-side chunks, NFB_LZ4_CHUNK_BYTES).
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nfact_b200.h"

namespace nfb {
void set_error(const std::string& msg);
}

namespace {

constexpr uint32_t kMagic = 0x184D2204u;
constexpr uint32_t kSkippableLo = 0x184D2A50u, kSkippableHi = 0x184D2A5Fu;

inline uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<uint32_t>(p[3]) << 24); }
inline uint64_t rd64(const uint8_t* p) { return rd32(p) | (static_cast<uint64_t>(rd32(p + 4)) << 32); }
inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// XXH32, seed 0 (Yann Collet's xxHash specification)
uint32_t xxh32(const uint8_t* p, size_t len) {
    constexpr uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    const uint8_t* const end = p + len;
    uint32_t h;
    if (len >= 16) {
        uint32_t v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0u - P1;
        const uint8_t* const limit = end - 16;
        do {
            v1 = rotl(v1 + rd32(p) * P2, 13) * P1;
            v2 = rotl(v2 + rd32(p + 4) * P2, 13) * P1;
            v3 = rotl(v3 + rd32(p + 8) * P2, 13) * P1;
            v4 = rotl(v4 + rd32(p + 12) * P2, 13) * P1;
            p += 16;
        } while (p <= limit);
        h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
    } else {
        h = P5;
    }
    h += static_cast<uint32_t>(len);
    while (p + 4 <= end) {
        h = rotl(h + rd32(p) * P3, 17) * P4;
        p += 4;
    }
    while (p < end) {
        h = rotl(h + (*p) * P5, 11) * P1;
        ++p;
    }
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}

struct Block {
    const uint8_t* src;
    uint32_t csize;
    bool stored;       // not compressed
    bool has_sum;
    uint32_t sum;
    int frame;
    int64_t dsize = -1;   // decoded size
    int64_t dst_off = 0;
};

struct Frame {
    bool independent, content_sum, has_size;
    uint64_t content_size;
    uint32_t block_max;
    uint32_t content_checksum = 0;
    size_t first_block = 0, n_blocks = 0;
};

// decoded size of one compressed block, or -1 when malformed (walks the sequences, copies nothing)
int64_t block_decoded_size(const uint8_t* p, size_t n, int64_t limit) {
    const uint8_t* const end = p + n;
    int64_t out = 0;
    if (n == 0) return -1;
    for (;;) {
        if (p >= end) return -1;
        const unsigned token = *p++;
        size_t lit = token >> 4;
        if (lit == 15) {
            unsigned b;
            do {
                if (p >= end) return -1;
                b = *p++;
                lit += b;
            } while (b == 255);
        }
        if (lit > static_cast<size_t>(end - p)) return -1;
        p += lit;
        out += static_cast<int64_t>(lit);
        if (p == end) break;   // last sequence: literals only
        if (end - p < 2) return -1;
        const unsigned off = p[0] | (p[1] << 8);
        p += 2;
        if (off == 0) return -1;
        size_t ml = token & 15;
        if (ml == 15) {
            unsigned b;
            do {
                if (p >= end) return -1;
                b = *p++;
                ml += b;
            } while (b == 255);
        }
        out += static_cast<int64_t>(ml + 4);
        if (out > limit) return -1;
    }
    return out <= limit ? out : -1;
}

// decode one compressed block to dst[0, cap); `hist` = bytes of valid output before dst a match may reach into
// (0 for an independent block).  Returns the decoded size or -1.
int64_t block_decode(const uint8_t* p, size_t n, uint8_t* dst, int64_t cap, int64_t hist) {
    const uint8_t* const end = p + n;
    uint8_t* op = dst;
    uint8_t* const oend = dst + cap;
    if (n == 0) return -1;
    for (;;) {
        if (p >= end) return -1;
        const unsigned token = *p++;
        size_t lit = token >> 4;
        if (lit == 15) {
            unsigned b;
            do {
                if (p >= end) return -1;
                b = *p++;
                lit += b;
            } while (b == 255);
        }
        if (lit > static_cast<size_t>(end - p) || lit > static_cast<size_t>(oend - op)) return -1;
        // short literal runs dominate text: one fixed 16 / 32-byte copy (a single vector move) instead of a memcpy
        // call of variable length, when both buffers have the slack; the surplus bytes are overwritten by what follows
        if (lit <= 16 && end - p >= 16 && oend - op >= 16) memcpy(op, p, 16);
        else if (lit <= 32 && end - p >= 32 && oend - op >= 32) memcpy(op, p, 32);
        else memcpy(op, p, lit);
        p += lit;
        op += lit;
        if (p == end) break;
        if (end - p < 2) return -1;
        const size_t off = p[0] | (p[1] << 8);
        p += 2;
        size_t ml = token & 15;
        if (ml == 15) {
            unsigned b;
            do {
                if (p >= end) return -1;
                b = *p++;
                ml += b;
            } while (b == 255);
        }
        ml += 4;
        if (off == 0 || static_cast<int64_t>(off) > (op - dst) + hist) return -1;
        if (ml > static_cast<size_t>(oend - op)) return -1;
        const uint8_t* m = op - off;
        uint8_t* const match_end = op + ml;
        if (off >= 16 && oend - match_end >= 16) {
            // fixed 16-byte steps that may run up to 15 bytes past the match (inside this block's own output, which
            // the next sequences overwrite); a step never reads what it writes because the distance is >= 16
            do { memcpy(op, m, 16); op += 16; m += 16; } while (op < match_end);
            op = match_end;
        } else if (off >= ml) {
            memcpy(op, m, ml);
            op += ml;
        } else if (off >= 8) {          // overlapping, period >= 8: forward 8-byte copies are safe
            uint8_t* const mend = op + ml;
            while (op + 8 <= mend) { memcpy(op, m, 8); op += 8; m += 8; }
            while (op < mend) *op++ = *m++;
        } else {                        // short period (run-length style): byte by byte
            for (size_t i = 0; i < ml; ++i) op[i] = m[i];
            op += ml;
        }
    }
    return op - dst;
}

// The same block format decoded into a scratch stream whose beginning is NOT the beginning of the frame: `pos` bytes
// of the stream exist, `known` (<= 65536) bytes of real history precede this block in the frame.  A match that reaches
// before the stream copies unknown bytes; taint[] marks every byte whose value depends on them (copied from a marked
// byte, directly or through a chain of matches).  Bytes with taint 0 are exactly what a decode from the start of the
// frame produces.  Plain byte loops: this only runs over the few MB in front of a chunk.
int64_t block_decode_taint(const uint8_t* p, size_t n, uint8_t* buf, uint8_t* taint, int64_t pos, int64_t cap,
                           int64_t known) {
    const uint8_t* const end = p + n;
    int64_t op = pos;
    const int64_t oend = pos + cap;
    if (n == 0) return -1;
    for (;;) {
        if (p >= end) return -1;
        const unsigned token = *p++;
        size_t lit = token >> 4;
        if (lit == 15) {
            unsigned b;
            do {
                if (p >= end) return -1;
                b = *p++;
                lit += b;
            } while (b == 255);
        }
        if (lit > static_cast<size_t>(end - p) || static_cast<int64_t>(lit) > oend - op) return -1;
        memcpy(buf + op, p, lit);
        memset(taint + op, 0, lit);
        p += lit;
        op += static_cast<int64_t>(lit);
        if (p == end) break;
        if (end - p < 2) return -1;
        const int64_t off = p[0] | (p[1] << 8);
        p += 2;
        size_t ml = token & 15;
        if (ml == 15) {
            unsigned b;
            do {
                if (p >= end) return -1;
                b = *p++;
                ml += b;
            } while (b == 255);
        }
        ml += 4;
        if (off == 0 || off > (op - pos) + known) return -1;
        if (static_cast<int64_t>(ml) > oend - op) return -1;
        for (size_t i = 0; i < ml; ++i, ++op) {
            const int64_t from = op - off;
            if (from < 0) { buf[op] = 0; taint[op] = 1; }
            else { buf[op] = buf[from]; taint[op] = taint[from]; }
        }
    }
    return op - pos;
}

// diagnostics of the last nfb_lz4_frame_decompress call on this thread (nfb_lz4_last_chunks)
thread_local int64_t g_last_chunks = 0, g_last_deferred = 0;

struct Parsed {
    std::vector<Frame> frames;
    std::vector<Block> blocks;
};

bool fail(const std::string& why) {
    nfb::set_error("LZ4F: " + why);
    return false;
}

// walk the container: frame headers, block table, checksums' positions
bool parse_container(const uint8_t* src, int64_t len, Parsed& out) {
    const uint8_t* p = src;
    const uint8_t* const end = src + len;
    if (len == 0) return true;     // lz4.frame.open(...).read() of an empty file is b""
    while (p < end) {
        if (end - p < 4) return fail("truncated frame magic");
        const uint32_t magic = rd32(p);
        if (magic >= kSkippableLo && magic <= kSkippableHi) {
            if (end - p < 8) return fail("truncated skippable frame");
            const uint32_t sz = rd32(p + 4);
            if (static_cast<uint64_t>(end - p - 8) < sz) return fail("truncated skippable frame");
            p += 8 + sz;
            continue;
        }
        if (magic != kMagic) return fail("ERROR_frameType_unknown (not an LZ4 frame)");
        p += 4;
        const uint8_t* const desc = p;
        if (end - p < 3) return fail("truncated frame header");
        const unsigned flg = p[0], bd = p[1];
        p += 2;
        if ((flg >> 6) != 1) return fail("ERROR_headerVersion_wrong");
        if (flg & 0x02) return fail("ERROR_reservedFlag_set");
        if (bd & 0x8F) return fail("ERROR_reservedFlag_set");
        const unsigned bs_id = (bd >> 4) & 7;
        if (bs_id < 4) return fail("ERROR_maxBlockSize_invalid");
        Frame fr;
        fr.independent = flg & 0x20;
        const bool block_sum = flg & 0x10;
        fr.has_size = flg & 0x08;
        fr.content_sum = flg & 0x04;
        const bool dict_id = flg & 0x01;
        fr.block_max = 1u << (8 + 2 * bs_id);   // 4 -> 64 KB, 5 -> 256 KB, 6 -> 1 MB, 7 -> 4 MB
        fr.content_size = 0;
        if (fr.has_size) {
            if (end - p < 8) return fail("truncated frame header");
            fr.content_size = rd64(p);
            p += 8;
        }
        if (dict_id) {
            if (end - p < 4) return fail("truncated frame header");
            p += 4;   // a dictionary id without the dictionary: matches into it fail the offset check
        }
        if (end - p < 1) return fail("truncated frame header");
        if (*p != ((xxh32(desc, static_cast<size_t>(p - desc)) >> 8) & 0xFF)) return fail("ERROR_headerChecksum_invalid");
        ++p;
        fr.first_block = out.blocks.size();
        for (;;) {
            if (end - p < 4) return fail("truncated frame (no EndMark)");
            const uint32_t word = rd32(p);
            p += 4;
            if (word == 0) break;
            Block b;
            b.stored = word >> 31;
            b.csize = word & 0x7FFFFFFFu;
            if (b.csize > fr.block_max) return fail("ERROR_maxBlockSize_invalid (block larger than the frame's maximum)");
            if (static_cast<uint64_t>(end - p) < static_cast<uint64_t>(b.csize) + (block_sum ? 4u : 0u))
                return fail("truncated block");
            b.src = p;
            p += b.csize;
            b.has_sum = block_sum;
            b.sum = 0;
            if (block_sum) { b.sum = rd32(p); p += 4; }
            b.frame = static_cast<int>(out.frames.size());
            if (b.stored) b.dsize = b.csize;
            out.blocks.push_back(b);
        }
        fr.n_blocks = out.blocks.size() - fr.first_block;
        if (fr.content_sum) {
            if (end - p < 4) return fail("truncated frame (content checksum missing)");
            fr.content_checksum = rd32(p);
            p += 4;
        }
        out.frames.push_back(fr);
    }
    return true;
}

int pick_threads(int nthreads, size_t work_items) {
    if (nthreads <= 0) nthreads = static_cast<int>(std::thread::hardware_concurrency());
    nthreads = std::max(1, std::min(nthreads, 64));
    return static_cast<int>(std::max<size_t>(1, std::min<size_t>(nthreads, work_items)));
}

template <typename F>
void parallel_for(size_t n, int nthreads, F&& body) {
    nthreads = pick_threads(nthreads, n);
    if (nthreads == 1) {
        for (size_t i = 0; i < n; ++i) body(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
        pool.emplace_back([&] {
            for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) body(i);
        });
    for (auto& th : pool) th.join();
}

// fills dsize of every block that can be sized without decoding; returns false on a malformed block.
// Blocks of frames with linked blocks are sized too (the walk needs no history).
bool size_blocks(Parsed& c, int nthreads) {
    std::atomic<bool> ok{true};
    parallel_for(c.blocks.size(), nthreads, [&](size_t i) {
        Block& b = c.blocks[i];
        if (b.has_sum && xxh32(b.src, b.csize) != b.sum) { ok = false; b.dsize = -2; return; }
        if (b.stored) return;
        b.dsize = block_decoded_size(b.src, b.csize, c.frames[b.frame].block_max);
        if (b.dsize < 0) ok = false;
    });
    if (!ok) {
        for (const Block& b : c.blocks)
            if (b.dsize == -2) return fail("ERROR_blockChecksum_invalid");
        return fail("ERROR_decompressionFailed (malformed block)");
    }
    int64_t off = 0;
    for (Block& b : c.blocks) {
        b.dst_off = off;
        off += b.dsize;
    }
    for (const Frame& fr : c.frames) {
        if (!fr.has_size) continue;
        int64_t total = 0;
        for (size_t i = 0; i < fr.n_blocks; ++i) total += c.blocks[fr.first_block + i].dsize;
        if (static_cast<uint64_t>(total) != fr.content_size) return fail("ERROR_frameSize_wrong");
    }
    return true;
}

}  // namespace

extern "C" {

int nfb_lz4_frame_size(const void* src, int64_t src_len, int nthreads, int64_t* size_out) {
    if ((src == nullptr && src_len != 0) || src_len < 0 || size_out == nullptr) {
        nfb::set_error("nfb_lz4_frame_size: bad arguments");
        return 2;
    }
    Parsed c;
    if (!parse_container(static_cast<const uint8_t*>(src), src_len, c)) return 2;
    if (!size_blocks(c, nthreads)) return 2;
    *size_out = c.blocks.empty() ? 0 : c.blocks.back().dst_off + c.blocks.back().dsize;
    return 0;
}

int nfb_lz4_frame_decompress(const void* src, int64_t src_len, void* dst, int64_t dst_cap, int nthreads,
                             int64_t* written_out) {
    if ((src == nullptr && src_len != 0) || src_len < 0 || dst_cap < 0 || (dst == nullptr && dst_cap != 0) ||
        written_out == nullptr) {
        nfb::set_error("nfb_lz4_frame_decompress: bad arguments");
        return 2;
    }
    Parsed c;
    if (!parse_container(static_cast<const uint8_t*>(src), src_len, c)) return 2;
    if (!size_blocks(c, nthreads)) return 2;
    const int64_t total = c.blocks.empty() ? 0 : c.blocks.back().dst_off + c.blocks.back().dsize;
    if (total > dst_cap) {
        nfb::set_error("nfb_lz4_frame_decompress: destination too small");
        return 2;
    }
    uint8_t* const out = static_cast<uint8_t*>(dst);
    // Work items: every block of an independent-block frame; a linked-block frame (python-lz4's default: every block
    // may reference the previous 64 KB of output) as one sequential item -- or, as an option, as several chunks
    // of consecutive blocks decoded side by side.  A chunk that does not start the frame first decodes the few MB in
    // front of it into a scratch stream with unknown history and exact taint tracking (block_decode_taint): once its
    // own first 64 KB come out untainted they are the true bytes, are copied into place, and the rest of the chunk
    // is decoded in place at full speed.  A chunk whose start stays tainted (long copy chains) is decoded after its
    // predecessor, from the real history.
    struct Item { size_t first, count; bool linked; int64_t frame_start; bool chunk_head; };
    std::vector<Item> items;
    const int threads_avail = pick_threads(nthreads, static_cast<size_t>(1) << 30);
    // Off unless NFB_LZ4_CHUNK_BYTES is set: on probtrackx text (rows sorted, the same prefixes copied forward from
    // match to match) the taint of a chunk's start never clears -- measured: every chunk of a 40 MB .dot frame had to
    // wait for its predecessor, and the warm-up is then pure overhead.  It pays on data whose copy chains are short.
    const int64_t chunk_min = getenv("NFB_LZ4_CHUNK_BYTES") ? std::max<int64_t>(1 << 17, atoll(getenv("NFB_LZ4_CHUNK_BYTES")))
                                                            : INT64_MAX;
    const int64_t warm_bytes = getenv("NFB_LZ4_WARM_BYTES") ? std::max<int64_t>(1 << 16, atoll(getenv("NFB_LZ4_WARM_BYTES")))
                                                                  : (static_cast<int64_t>(4) << 20);
    for (const Frame& fr : c.frames) {
        if (fr.independent) {
            for (size_t i = 0; i < fr.n_blocks; ++i)
                items.push_back({fr.first_block + i, 1, false, c.blocks[fr.first_block + i].dst_off, true});
        } else if (fr.n_blocks) {
            const Block& first = c.blocks[fr.first_block];
            const Block& last = c.blocks[fr.first_block + fr.n_blocks - 1];
            const int64_t bytes = last.dst_off + last.dsize - first.dst_off;
            const int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(threads_avail, bytes / chunk_min));
            const int64_t per_chunk = (bytes + n_chunks - 1) / n_chunks;
            size_t begin = fr.first_block;
            for (int64_t ch = 0; ch < n_chunks && begin < fr.first_block + fr.n_blocks; ++ch) {
                size_t stop = begin;
                const int64_t limit = first.dst_off + (ch + 1) * per_chunk;
                while (stop < fr.first_block + fr.n_blocks && (c.blocks[stop].dst_off < limit || ch == n_chunks - 1)) ++stop;
                if (stop == begin) ++stop;
                items.push_back({begin, stop - begin, true, first.dst_off, begin == fr.first_block});
                begin = stop;
            }
        }
    }
    std::atomic<bool> ok{true};
    std::vector<char> deferred(items.size(), 0);
    g_last_chunks = 0;
    for (const Item& it : items) g_last_chunks += (it.linked && !it.chunk_head);
    g_last_deferred = 0;
    // in-place decode of blocks [from, to) of a linked frame whose history is already in the output buffer
    auto decode_in_place = [&](size_t from, size_t to, int64_t frame_start, bool linked) {
        for (size_t i = from; i < to; ++i) {
            const Block& b = c.blocks[i];
            uint8_t* d = out + b.dst_off;
            if (b.stored) {
                memcpy(d, b.src, b.csize);
                continue;
            }
            const int64_t hist = linked ? std::min<int64_t>(b.dst_off - frame_start, 65536) : 0;
            if (block_decode(b.src, b.csize, d, b.dsize, hist) != b.dsize) ok = false;
        }
    };
    parallel_for(items.size(), nthreads, [&](size_t w) {
        const Item& it = items[w];
        if (!it.linked || it.chunk_head) {
            decode_in_place(it.first, it.first + it.count, it.frame_start, it.linked);
            return;
        }
        // warm-up: walk back until ~warm_bytes of output (or the start of the frame) lie in front of the chunk
        const size_t frame_first = [&] { size_t i = it.first; while (c.blocks[i].dst_off > it.frame_start) --i; return i; }();
        size_t warm = it.first;
        while (warm > frame_first && c.blocks[it.first].dst_off - c.blocks[warm].dst_off < warm_bytes) --warm;
        // own blocks decoded into the scratch stream as well, until 64 KB of own output exist
        size_t own_end = it.first;
        int64_t own_bytes = 0;
        while (own_end < it.first + it.count && own_bytes < 65536) own_bytes += c.blocks[own_end++].dsize;
        const int64_t stream_start = c.blocks[warm].dst_off;
        const int64_t stream_len = c.blocks[own_end - 1].dst_off + c.blocks[own_end - 1].dsize - stream_start;
        std::vector<uint8_t> buf(static_cast<size_t>(stream_len)), taint(static_cast<size_t>(stream_len));
        bool bad = false;
        for (size_t i = warm; i < own_end && !bad; ++i) {
            const Block& b = c.blocks[i];
            const int64_t pos = b.dst_off - stream_start;
            if (b.stored) {
                memcpy(buf.data() + pos, b.src, b.csize);
                memset(taint.data() + pos, 0, b.csize);
                continue;
            }
            // real history in front of this block; what lies before the stream is unknown unless the stream starts the frame
            const int64_t known = std::min<int64_t>(b.dst_off - it.frame_start, 65536);
            // (a stream that begins at the frame start has known <= pos: nothing can reach in front of it)
            if (block_decode_taint(b.src, b.csize, buf.data(), taint.data(), pos, b.dsize, known) != b.dsize) bad = true;
        }
        if (bad) { ok = false; return; }
        const int64_t own_pos = c.blocks[it.first].dst_off - stream_start;
        const bool clean = memchr(taint.data() + own_pos, 1, static_cast<size_t>(stream_len - own_pos)) == nullptr;
        if (!clean) { deferred[w] = 1; return; }
        memcpy(out + c.blocks[it.first].dst_off, buf.data() + own_pos, static_cast<size_t>(stream_len - own_pos));
        decode_in_place(own_end, it.first + it.count, it.frame_start, true);
    });
    // chunks whose first bytes stayed tainted: in order, each after everything in front of it is final
    for (size_t w = 0; w < items.size(); ++w)
        if (deferred[w]) {
            ++g_last_deferred;
            decode_in_place(items[w].first, items[w].first + items[w].count, items[w].frame_start, true);
        }
    if (!ok) { fail("ERROR_decompressionFailed (match offset outside the window)"); return 2; }
    // content checksums: one frame per thread (XXH32 is sequential within a frame)
    std::atomic<bool> sums_ok{true};
    parallel_for(c.frames.size(), nthreads, [&](size_t f) {
        const Frame& fr = c.frames[f];
        if (!fr.content_sum) return;
        int64_t begin = 0, bytes = 0;
        if (fr.n_blocks) {
            begin = c.blocks[fr.first_block].dst_off;
            const Block& last = c.blocks[fr.first_block + fr.n_blocks - 1];
            bytes = last.dst_off + last.dsize - begin;
        }
        if (xxh32(out + begin, static_cast<size_t>(bytes)) != fr.content_checksum) sums_ok = false;
    });
    if (!sums_ok) { fail("ERROR_contentChecksum_invalid"); return 2; }
    *written_out = total;
    return 0;
}

int nfb_lz4_last_chunks(int64_t* parallel_chunks_out, int64_t* deferred_out) {
    if (parallel_chunks_out) *parallel_chunks_out = g_last_chunks;
    if (deferred_out) *deferred_out = g_last_deferred;
    return 0;
}

}  // extern "C"
