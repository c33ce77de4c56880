// Split-precision tcgen05 GEMM: the hot contraction of the NMF / dual-regression path.
//
//   D[r, t] = sum_K  A[r, K] * B[t, K]            r < M_total,  t < N (<= 256),  K streamed
//
// A is the connectivity matrix X (or X^T), B is a factor matrix stored component-major
// ([k_pad, dim], K contiguous).  Every operand is held as two fp16 planes (hi + lo of a
// power-of-two pre-scaled fp32 value, 22 significant bits) and each K step issues three
// tcgen05.mma.kind::f16 (hi*hi + lo*hi + hi*lo) into one fp32 TMEM accumulator, so the result
// matches an fp32 GEMM to ~2^-22 per product while X still costs 4 B/element of HBM traffic.
//
// This replaces the two sgemm calls sklearn makes per NMF iteration
// (sklearn/decomposition/_nmf.py:376 `safe_sparse_dot(X, Ht)` and :631 `safe_sparse_dot(W.T, X)`,
// reached from NFACT/decomp/decomposition/sso/nmfsso.py:54-56) and the k x k Gram / denominator
// products around them (:375, :550, :632).
//
// Kernel shape: persistent, warp specialised (1 TMA producer thread, 1 MMA issuer thread, 8 epilogue
// warps), multi-stage smem ring fed by TMA with 128B swizzle, two TMEM accumulator stages,
// split-K over `splits` deterministic partial outputs (no atomics), optional 2-CTA pairs
// (cta_group::2: M = 256 per pair, each CTA stages half of B).
//
// Accumulation promotion: the tensor core truncates when it adds into the fp32 TMEM accumulator
// (measured on B200: about -7e-8 relative per 64-wide K block of chain length for non-negative data,
// -1.3e-4 over K = 110k), so a TMEM chain is at most kChunkKB K blocks long; the epilogue warps drain
// every finished chain into fp32 registers with round-to-nearest adds while the issuer fills the other
// TMEM stage (the FP8-style "promotion" scheme, applied here to keep fp32-faithful sums).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace nfb {

namespace {

constexpr int kBlockM = 128;       // accumulator rows per CTA == TMEM lanes
constexpr int kBlockK = 64;        // fp16 elements per K block: 128 B rows, SWIZZLE_128B
constexpr int kUmmaK = 16;         // K per tcgen05.mma for 16-bit inputs
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB per plane
constexpr int kTmemCols = 512;
constexpr int kAccStageCols = 256;
constexpr int kNumEpilogueWarps = 8;
constexpr int kNumThreads = 128 + kNumEpilogueWarps * 32;
constexpr int kMaxStages = 8;
constexpr int kChunkKB = 8;        // K blocks per TMEM accumulation chain (512 K elements)

struct GemmArgs {
    float* out;              // [splits][n_store rows][ld_out]   (component-major: D^T)
    const float* col_scale;  // [N] per-column multiplier folded into the store (nullptr -> 1)
    const float* row_scale;  // [m_total] per-row multiplier (nullptr -> 1)
    float scale;             // global multiplier
    int m_total;             // valid rows of D
    int n_umma;              // UMMA N (multiple of 16, 16..256)
    int n_store;             // columns of D that are stored (<= n_umma)
    int num_kb;              // K blocks in total
    int splits;              // split-K factor (<= num_kb)
    int m_blocks;            // ceil(m_total / (128 * cta_group))
    int ld_out;
    long long split_stride;  // elements between consecutive split slabs
    int num_stages;
    int b_tile_bytes;        // per plane, per CTA
    // Fused reduce-scatter (row-sharded W'X, p2p.cu): row r of D (an H column) is stored straight into the inbox of
    // the rank that owns it -- peer memory over NVLink -- instead of into `out`:
    //   scatter_base[r / scatter_slice][(scatter_src * scatter_rows + col0 + col) * scatter_slice + r % scatter_slice]
    int scatter_slice;       // H columns per rank (0: plain store into `out`)
    int scatter_rows;        // k: component rows per source slab
    int scatter_src;         // this rank
    int scatter_col0;        // first component of this N slab
    float* scatter_base[kGemmMaxPeers];
    // Tile order of the scatter epilogue.  In natural order the 74 tiles that run side by side cover ~1.4 owners,
    // and they do so on every rank at once: all ranks store into the same one or two inboxes (incast on one
    // NVLink ingress port while the other seven idle).  With perm_groups = P the M blocks are dealt out round-robin
    // over P contiguous groups of perm_bpo blocks (about one owner each), starting at group perm_rot, so that at
    // any time every rank's stores are spread evenly over all owners.  0: natural order.
    int perm_groups, perm_bpo, perm_rot;
    int num_tiles;           // tiles to walk (m_blocks * splits, or perm_groups * perm_bpo: some are empty)
    // Stream-K tail (unsplit launches): every worker first takes sk_full whole tiles (tile = worker + i * workers); the
    // remaining sk_tail tiles are cut into equal shares of sk_share K blocks per worker, so the last, partial wave of
    // tiles costs its share of the K loop instead of a whole tile time.  A share covers the end of one tile and / or
    // the beginning of the next: the worker that holds a tile's LAST K blocks owns the tile -- it waits for the other
    // workers' raw fp32 partials (sk_ws, one slot per worker, arrivals counted in sk_cnt), adds them in worker order
    // (deterministic) and stores the tile once.  sk_share == 0: off.
    int sk_full, sk_tail, sk_share;
    float* sk_ws;            // [workers][n_store][256] raw partial accumulators of a worker's contributed segment
    unsigned int* sk_cnt;    // [workers][2]: epilogue warps that have stored / read slot w (reset by the last reader)
};

struct WorkItem {
    int mb, ks, kb0, kb1;
    int role;                // 0: whole tile (or split slab), 1: contributes a partial, 2: owns a tile others contribute to
    int first;               // owner: contributors are the workers first .. worker - 1
};
constexpr int kItemEnd = -1, kItemSkip = 0, kItemWork = 1;

__device__ __forceinline__ bool map_tile(const GemmArgs& args, int tile, int& ks, int& mb);

// idx-th work item of a worker; the three roles of a CTA pair walk the same list
__device__ __forceinline__ int next_item(const GemmArgs& args, int worker, int num_workers, int idx, WorkItem& it) {
    it.role = 0;
    it.first = 0;
    if (args.sk_share == 0) {
        const int tile = worker + idx * num_workers;
        if (tile >= args.num_tiles) return kItemEnd;
        if (!map_tile(args, tile, it.ks, it.mb)) return kItemSkip;
        it.kb0 = static_cast<int>(static_cast<long long>(it.ks) * args.num_kb / args.splits);
        it.kb1 = static_cast<int>(static_cast<long long>(it.ks + 1) * args.num_kb / args.splits);
        return kItemWork;
    }
    it.ks = 0;
    if (idx < args.sk_full) {
        it.mb = worker + idx * num_workers;
        it.kb0 = 0;
        it.kb1 = args.num_kb;
        return kItemWork;
    }
    const int kb = args.num_kb;
    const long long u0 = static_cast<long long>(worker) * args.sk_share;
    const long long u1 = min(u0 + args.sk_share, static_cast<long long>(args.sk_tail) * kb);
    if (u0 >= u1) return kItemEnd;
    const int ta = static_cast<int>(u0 / kb);
    const long long end_a = min(u1, static_cast<long long>(ta + 1) * kb);
    const bool has_b = u1 > end_a;
    const int j = idx - args.sk_full;
    // contribution first, owned tile last: an owner never waits for a worker that is itself still waiting
    if (has_b && j == 0) {                       // the beginning of tile ta + 1
        it.mb = args.sk_full * num_workers + ta + 1;
        it.kb0 = 0;
        it.kb1 = static_cast<int>(u1 - end_a);
        it.role = 1;
        return kItemWork;
    }
    if (j > (has_b ? 1 : 0)) return kItemEnd;
    it.mb = args.sk_full * num_workers + ta;
    it.kb0 = static_cast<int>(u0 - static_cast<long long>(ta) * kb);
    it.kb1 = static_cast<int>(end_a - static_cast<long long>(ta) * kb);
    if (it.kb1 < kb) it.role = 1;
    else if (it.kb0 > 0) {
        it.role = 2;
        it.first = static_cast<int>((static_cast<long long>(ta) * kb) / args.sk_share);
    }
    return kItemWork;
}

// tile index -> (K split, M block); false for the empty tiles of the permuted order
__device__ __forceinline__ bool map_tile(const GemmArgs& args, int tile, int& ks, int& mb) {
    if (args.perm_groups > 0) {
        ks = 0;
        mb = ((tile % args.perm_groups + args.perm_rot) % args.perm_groups) * args.perm_bpo + tile / args.perm_groups;
        return mb < args.m_blocks;
    }
    ks = tile / args.m_blocks;
    mb = tile % args.m_blocks;
    return true;
}

// 64-bit shared-memory matrix descriptor (sm_100 layout: version=1 at bit 46, SWIZZLE_128B = 2 at bit 61)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t desc = 0;
    desc |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    desc |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    desc |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    desc |= static_cast<uint64_t>(1) << 46;
    desc |= static_cast<uint64_t>(2) << 61;
    return desc;
}

// instruction descriptor for kind::f16: fp16 x fp16 -> fp32, B always K-major
__device__ __forceinline__ uint32_t make_idesc(int m, int n, int a_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                                 // D format: f32
    d |= 0u << 7;                                 // A format: f16
    d |= 0u << 10;                                // B format: f16
    d |= static_cast<uint32_t>(a_mn_major) << 15; // A major
    d |= 0u << 16;                                // B major: K
    d |= static_cast<uint32_t>(n >> 3) << 17;
    d |= static_cast<uint32_t>(m >> 4) << 24;
    return d;
}

// kNch = 16-column accumulator chunks owned by one epilogue thread (two warps share a TMEM lane quarter
// and split the N columns between them): N <= 32 * kNch.
template <int kAMajorMN, int kCtaGroup, int kNch>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_split_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                  const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                  const GemmArgs args) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t warp_idx = threadIdx.x >> 5;
    const uint32_t lane_idx = threadIdx.x & 31;
    const uint32_t cta_rank = kCtaGroup == 2 ? cluster_ctarank() : 0;
    const bool is_leader = cta_rank == 0;

    // ---- smem carve-up (1024 B aligned for SWIZZLE_128B) ---------------------------------------
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int stage_bytes = 2 * kATileBytes + 2 * args.b_tile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(args.num_stages) * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kMaxStages;
    uint64_t* tmem_full_bar = bars + 2 * kMaxStages;
    uint64_t* tmem_empty_bar = bars + 2 * kMaxStages + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

    if (warp_idx == 0 && lane_idx == 0) {
        tma_prefetch_desc(&tm_a_hi);
        tma_prefetch_desc(&tm_a_lo);
        tma_prefetch_desc(&tm_b_hi);
        tma_prefetch_desc(&tm_b_lo);
    }
    if (warp_idx == 1 && lane_idx == 0) {
        for (int s = 0; s < args.num_stages; ++s) {
            mbar_init(&full_bar[s], kCtaGroup);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], kCtaGroup * kNumEpilogueWarps * 32);
        }
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<kCtaGroup>(tmem_ptr_smem, kTmemCols);
    tc_fence_before();
    if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_workers = gridDim.x / kCtaGroup;
    const int worker = blockIdx.x / kCtaGroup;
    const int n_load = args.n_umma / kCtaGroup;  // rows of B staged by this CTA

    uint32_t stage = 0, phase = 0;
    auto advance = [&]() {
        if (++stage == static_cast<uint32_t>(args.num_stages)) { stage = 0; phase ^= 1; }
    };

    if (warp_idx == 0) {
      if (lane_idx == 0) {
        // =========================== TMA producer ===============================================
        for (int idx = 0;; ++idx) {
            WorkItem item;
            const int what = next_item(args, worker, num_workers, idx, item);
            if (what == kItemEnd) break;
            if (what == kItemSkip) continue;
            const int mb = item.mb, kb0 = item.kb0, kb1 = item.kb1;
            const int m_idx = (mb * kCtaGroup + static_cast<int>(cta_rank)) * kBlockM;
            const int n_idx = static_cast<int>(cta_rank) * n_load;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa_hi = smem + static_cast<size_t>(stage) * stage_bytes;
                uint8_t* sa_lo = sa_hi + kATileBytes;
                uint8_t* sb_hi = sa_lo + kATileBytes;
                uint8_t* sb_lo = sb_hi + args.b_tile_bytes;
                const int k_idx = kb * kBlockK;
                auto load = [&](const CUtensorMap* tm, void* dst, int c0, int c1) {
                    if constexpr (kCtaGroup == 2) tma_load_2d_2sm(tm, &full_bar[stage], dst, c0, c1);
                    else tma_load_2d(tm, &full_bar[stage], dst, c0, c1);
                };
                if constexpr (kAMajorMN) {
                    // A = X^T: global rows are K, global columns are M.  Two 64-column atoms.
                    for (int atom = 0; atom < 2; ++atom) {
                        load(&tm_a_hi, sa_hi + atom * (kBlockK * 128), m_idx + atom * 64, k_idx);
                        load(&tm_a_lo, sa_lo + atom * (kBlockK * 128), m_idx + atom * 64, k_idx);
                    }
                } else {
                    load(&tm_a_hi, sa_hi, k_idx, m_idx);
                    load(&tm_a_lo, sa_lo, k_idx, m_idx);
                }
                load(&tm_b_hi, sb_hi, k_idx, n_idx);
                load(&tm_b_lo, sb_lo, k_idx, n_idx);
                if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(stage_bytes) * kCtaGroup);
                else mbar_arrive_cluster(&full_bar[stage], 0);
                advance();
            }
        }
      }
      __syncwarp();
    } else if (warp_idx == 1) {
      if (lane_idx == 0 && is_leader) {
        // =========================== MMA issuer =================================================
        const uint32_t idesc = make_idesc(kBlockM * kCtaGroup, args.n_umma, kAMajorMN);
        // A: K-major  -> SBO = 8 rows * 128 B;              K step = 32 B inside the swizzle atom
        //    MN-major -> LBO = next 64-wide M atom (8 KB),  SBO = 8 K-rows * 128 B, K step = 16 rows * 128 B
        const uint32_t a_lbo = kAMajorMN ? kBlockK * 128 : 0;
        const uint32_t a_kstep = kAMajorMN ? kUmmaK * 128 : kUmmaK * 2;
        uint32_t chain = 0;  // running index of TMEM accumulation chains (stage = chain & 1)
        for (int idx = 0;; ++idx) {
            WorkItem item;
            const int what = next_item(args, worker, num_workers, idx, item);
            if (what == kItemEnd) break;
            if (what == kItemSkip) continue;
            const int kb0 = item.kb0, kb1 = item.kb1;
            for (int kc0 = kb0; kc0 < kb1; kc0 += kChunkKB, ++chain) {
                const int kc1 = min(kc0 + kChunkKB, kb1);
                const uint32_t as = chain & 1, aphase = (chain >> 1) & 1;
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * kAccStageCols;
                for (int kb = kc0; kb < kc1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa_hi = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
                    const uint32_t sa_lo = sa_hi + kATileBytes;
                    const uint32_t sb_hi = sa_lo + kATileBytes;
                    const uint32_t sb_lo = sb_hi + args.b_tile_bytes;
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        const uint64_t da_hi = make_smem_desc(sa_hi + k * a_kstep, a_lbo, 1024);
                        const uint64_t da_lo = make_smem_desc(sa_lo + k * a_kstep, a_lbo, 1024);
                        const uint64_t db_hi = make_smem_desc(sb_hi + k * kUmmaK * 2, 0, 1024);
                        const uint64_t db_lo = make_smem_desc(sb_lo + k * kUmmaK * 2, 0, 1024);
                        umma_f16<kCtaGroup>(d_tmem, da_hi, db_hi, idesc, (kb > kc0 || k > 0) ? 1u : 0u);
                        umma_f16<kCtaGroup>(d_tmem, da_lo, db_hi, idesc, 1u);
                        umma_f16<kCtaGroup>(d_tmem, da_hi, db_lo, idesc, 1u);
                    }
                    umma_commit<kCtaGroup>(&empty_bar[stage]);
                    if (kb == kc1 - 1) umma_commit<kCtaGroup>(&tmem_full_bar[as]);
                    advance();
                }
            }
        }
      }
      __syncwarp();
    } else if (warp_idx >= 4) {
        // ====== epilogue: drain every chain TMEM -> fp32 registers (RN adds), store once per tile ======
        const uint32_t ewarp = warp_idx - 4;
        const uint32_t quarter = warp_idx & 3;          // TMEM lanes this warp may touch
        const uint32_t col_half = ewarp >> 2;           // which half of the N columns
        const uint32_t row = quarter * 32 + lane_idx;
        const int nch_total = args.n_umma >> 4;
        const int nch_first = (nch_total + 1) >> 1;
        const int ch0 = col_half == 0 ? 0 : nch_first;                       // first 16-col chunk
        const int nch = col_half == 0 ? nch_first : nch_total - nch_first;   // chunks owned (<= kNch)
        uint32_t chain = 0;
        for (int idx = 0;; ++idx) {
            WorkItem item;
            const int what = next_item(args, worker, num_workers, idx, item);
            if (what == kItemEnd) break;
            if (what == kItemSkip) continue;
            const int ks = item.ks, mb = item.mb, kb0 = item.kb0, kb1 = item.kb1;
            float acc[kNch * 16];
#pragma unroll
            for (int i = 0; i < kNch * 16; ++i) acc[i] = 0.0f;
            for (int kc0 = kb0; kc0 < kb1; kc0 += kChunkKB, ++chain) {
                const uint32_t as = chain & 1, aphase = (chain >> 1) & 1;
                mbar_wait(&tmem_full_bar[as], aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + as * kAccStageCols + ch0 * 16;
#pragma unroll
                for (int c = 0; c < kNch; ++c) {
                    if (c < nch) {  // warp-uniform
                        uint32_t v[32];
                        tmem_ld_32x32b_x16(taddr + c * 16, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[c * 16 + j] = __fadd_rn(acc[c * 16 + j], __uint_as_float(v[j]));
                    }
                }
                tc_fence_before();
                if constexpr (kCtaGroup == 2) mbar_arrive_cluster(&tmem_empty_bar[as], 0);
                else mbar_arrive(&tmem_empty_bar[as]);
            }
            if (item.role != 0) {
                // stream-K: slot layout [n_store][kCtaGroup * 128 rows], rows of a warp contiguous (coalesced)
                const int pair_rows = kCtaGroup * kBlockM;
                const int prow = static_cast<int>(cta_rank) * kBlockM + static_cast<int>(row);
                constexpr unsigned int kWarpsPerSlot = kCtaGroup * kNumEpilogueWarps;
                if (item.role == 1) {
                    float* ws = args.sk_ws + static_cast<size_t>(worker) * args.n_store * pair_rows + prow;
#pragma unroll
                    for (int c = 0; c < kNch; ++c) {
                        if (c < nch) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int col = (ch0 + c) * 16 + j;
                                if (col < args.n_store) ws[static_cast<size_t>(col) * pair_rows] = acc[c * 16 + j];
                            }
                        }
                    }
                    __threadfence();
                    __syncwarp();
                    if (lane_idx == 0) atomicAdd(args.sk_cnt + 2 * worker, 1u);
                    continue;
                }
                for (int src = item.first; src < worker; ++src) {
                    unsigned int* cnt = args.sk_cnt + 2 * src;
                    if (lane_idx == 0) {
                        unsigned int seen;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
                            if (seen < kWarpsPerSlot) __nanosleep(40);
                        } while (seen < kWarpsPerSlot);
                    }
                    __syncwarp();
                    __threadfence();
                    const float* ws = args.sk_ws + static_cast<size_t>(src) * args.n_store * pair_rows + prow;
#pragma unroll
                    for (int c = 0; c < kNch; ++c) {
                        if (c < nch) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int col = (ch0 + c) * 16 + j;
                                if (col < args.n_store)
                                    acc[c * 16 + j] = __fadd_rn(acc[c * 16 + j], __ldcg(ws + static_cast<size_t>(col) * pair_rows));
                            }
                        }
                    }
                    __syncwarp();
                    if (lane_idx == 0) {
                        // the last of the owner's warps to have read the slot re-arms it for the next launch
                        if (atomicAdd(cnt + 1, 1u) == kWarpsPerSlot - 1) {
                            cnt[1] = 0u;
                            __threadfence();
                            cnt[0] = 0u;
                        }
                    }
                }
            }
            const long long m_row = static_cast<long long>(mb * kCtaGroup + static_cast<int>(cta_rank)) * kBlockM + row;
            if (m_row < args.m_total) {
                float* outp = args.out + static_cast<long long>(ks) * args.split_stride + m_row;
                if (args.scatter_slice > 0) {
                    // the tile's rows may straddle two owners: every thread (= row) picks its own destination
                    const int owner = static_cast<int>(m_row / args.scatter_slice);
                    outp = args.scatter_base[owner] +
                           static_cast<long long>(args.scatter_src * args.scatter_rows + args.scatter_col0) * args.scatter_slice +
                           (m_row - static_cast<long long>(owner) * args.scatter_slice);
                }
                const float rs = args.row_scale ? __ldg(args.row_scale + m_row) * args.scale : args.scale;
#pragma unroll
                for (int c = 0; c < kNch; ++c) {
                    if (c < nch) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int col = (ch0 + c) * 16 + j;
                            if (col < args.n_store) {
                                const float cs = args.col_scale ? __ldg(args.col_scale + col) : 1.0f;
                                outp[static_cast<long long>(col) * args.ld_out] = (acc[c * 16 + j] * cs) * rs;
                            }
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
    if (warp_idx == 2) tmem_dealloc<kCtaGroup>(tmem_base, kTmemCols);
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// fp16 plane [rows, ld] row-major; box = box_inner (<= 64 elements, 128 B) x box_outer rows
int make_plane_map(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_inner,
                   int box_outer) {
    EncodeTiledFn fn = encode_fn();
    NFB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    NFB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0, "plane must be 16 B aligned");
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NFB_CHECK(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(rc)));
    return 0;
}

struct Launch {
    const CUtensorMap *a_hi, *a_lo, *b_hi, *b_lo;
    const GemmArgs* args;
    int grid;
    size_t smem_bytes;
    cudaStream_t stream;
};

template <int kAMajorMN, int kCtaGroup, int kNch>
int launch(const Launch& l) {
    auto kernel = gemm_split_kernel<kAMajorMN, kCtaGroup, kNch>;
    NFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(l.smem_bytes)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(l.grid);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = l.smem_bytes;
    cfg.stream = l.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCtaGroup;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NFB_CUDA(cudaLaunchKernelEx(&cfg, kernel, *l.a_hi, *l.a_lo, *l.b_hi, *l.b_lo, *l.args));
    return 0;
}

template <int kAMajorMN, int kCtaGroup>
int launch_nch(const Launch& l, int nch_per_thread) {
    if (nch_per_thread <= 2) return launch<kAMajorMN, kCtaGroup, 2>(l);
    if (nch_per_thread <= 4) return launch<kAMajorMN, kCtaGroup, 4>(l);
    if (nch_per_thread <= 7) return launch<kAMajorMN, kCtaGroup, 7>(l);
    return launch<kAMajorMN, kCtaGroup, 8>(l);
}

}  // namespace

int gemm_default_cta_group() {
    static int value = -1;
    if (value < 0) {
        const char* env = getenv("NFB_GEMM_CTA_GROUP");
        value = env ? atoi(env) : 2;
        if (value != 1 && value != 2) value = 2;
    }
    return value;
}


int gemm_pick_splits(int64_t m_total, int64_t k_total, int cta_group, int64_t n_out) {
    const int64_t m_blocks = ceil_div64(m_total, static_cast<int64_t>(kBlockM) * cta_group);
    const int64_t num_kb = ceil_div64(k_total, kBlockK);
    const int64_t workers = sm_count() / cta_group;
    // Cost in bytes of HBM time: the operand stream (4 B per element of A, at ~80 % of peak when the waves are
    // full) divided by the wave efficiency, plus every split slab written once and read once by the consumer.
    // Every split stays at least 4 K blocks long.
    int64_t best = 1;
    double best_cost = 0.0;
    const int64_t max_splits = std::max<int64_t>(1, std::min<int64_t>(num_kb / 4, 64));
    for (int64_t s = 1; s <= max_splits; ++s) {
        const int64_t tiles = m_blocks * s;
        const int64_t waves = ceil_div64(tiles, workers);
        const double eff = static_cast<double>(tiles) / static_cast<double>(waves * workers);
        const double stream = 4.0 * static_cast<double>(m_total) * static_cast<double>(k_total) / (0.8 * eff);
        const double slabs = s > 1 ? 8.0 * static_cast<double>(s) * static_cast<double>(n_out) * static_cast<double>(m_total) : 0.0;
        const double cost = stream + slabs;
        if (s == 1 || cost < best_cost) { best_cost = cost; best = s; }
    }
    return static_cast<int>(best);
}

bool gemm_streamk_enabled() {
    // Off unless NFB_STREAMK=1.  Measured on B200: the tail engages as planned and buys nothing --
    // C3's W'X runs at the power cap (the SMs a partial wave leaves idle are power the busy ones use), C2's is HBM-bound
    // (the tiles of a partial wave get the whole memory system and finish early); wave quantisation is not what these
    // launches pay for.
    static const bool on = getenv("NFB_STREAMK") && atoi(getenv("NFB_STREAMK")) != 0;
    return on;
}

// the stream-K decision for an unsplit launch over m_total rows and k_total contraction elements: tiles per worker
// before the shared tail, tiles in the tail, K blocks of the tail per worker; false when whole tiles do as well
static bool streamk_plan(int64_t m_total, int64_t k_total, int cta_group, int* full_out, int* tail_out, int* share_out) {
    const int64_t m_blocks = ceil_div64(m_total, static_cast<int64_t>(kBlockM) * cta_group);
    const int64_t num_kb = ceil_div64(k_total, kBlockK);
    const int64_t w_all = sm_count() / cta_group;
    const int64_t full = m_blocks / w_all, tail = m_blocks - full * w_all;
    const int64_t share = tail > 0 ? (tail * num_kb + w_all - 1) / w_all : 0;
    // Worth it when the shared tail beats a whole extra wave by more than the fix-up costs: the partials of a shared
    // tile are written, awaited and re-read at the very end of the launch, ~30-50 us that nothing hides (the 8-GPU
    // shard's W'X, where the tail saves 22 of 700 K blocks, ran 0.03 ms slower), so the saving must be worth ~100 K
    // blocks.  At C3 the tail saves 175 of 5574 K-block times of W'X -- and buys nothing measurable there: the launch
    // runs at the power cap, and the SMs a partial wave leaves idle are power the busy ones use (5.60-5.64 ms either
    // way); at C2 (HBM-bound) the unsplit launch + tail is 1 % of the iteration faster than the split-K slabs.
    const double t_plain = static_cast<double>(full + 1) * num_kb, t_sk = static_cast<double>(full) * num_kb + share;
    if (!(tail > 0 && share >= 2 * kChunkKB && share < num_kb && t_plain - t_sk >= std::max(96.0, 0.02 * t_plain)))
        return false;
    if (full_out) *full_out = static_cast<int>(full);
    if (tail_out) *tail_out = static_cast<int>(tail);
    if (share_out) *share_out = static_cast<int>(share);
    return true;
}
bool gemm_streamk_pays(int64_t m_total, int64_t k_total, int cta_group) {
    return gemm_streamk_enabled() && streamk_plan(m_total, k_total, cta_group, nullptr, nullptr, nullptr);
}

size_t gemm_streamk_ws_floats(int n_store) {
    return static_cast<size_t>(sm_count()) * std::min(n_store, 256) * kBlockM;   // workers * cta_group * 128 rows
}
int gemm_streamk_max_workers() { return sm_count(); }

static int gemm_split_tile(const GemmOperand& a, bool a_is_mn_major, const GemmOperand& b, int64_t m_total,
                           int64_t k_total, int n_umma, int n_store, float* out, int64_t ld_out,
                           long long split_stride, int splits, const float* col_scale, const float* row_scale,
                           float scale, int cta_group, cudaStream_t stream, const GemmScatter* scatter, int col0,
                           const GemmStreamK* sk) {
    NFB_CHECK(n_umma % 16 == 0 && n_umma >= 16 && n_umma <= 256, "n_umma must be a multiple of 16 in [16,256]");
    NFB_CHECK(cta_group == 1 || cta_group == 2, "cta_group must be 1 or 2");
    NFB_CHECK(n_store <= n_umma, "n_store exceeds n_umma");
    NFB_CHECK(m_total < (1LL << 31) && k_total < (1LL << 31), "dimension exceeds int32");
    const int64_t num_kb = ceil_div64(k_total, kBlockK);
    NFB_CHECK(splits >= 1 && splits <= num_kb, "splits must be in [1, num_k_blocks]");

    CUtensorMap tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo;
    if (a_is_mn_major) {
        // plane rows = K, plane cols = M
        NFB_TRY(make_plane_map(&tm_a_hi, a.hi, k_total, m_total, a.ld, 64, kBlockK));
        NFB_TRY(make_plane_map(&tm_a_lo, a.lo, k_total, m_total, a.ld, 64, kBlockK));
    } else {
        NFB_TRY(make_plane_map(&tm_a_hi, a.hi, m_total, k_total, a.ld, kBlockK, kBlockM));
        NFB_TRY(make_plane_map(&tm_a_lo, a.lo, m_total, k_total, a.ld, kBlockK, kBlockM));
    }
    const int n_load = n_umma / cta_group;
    NFB_CHECK(n_load % 8 == 0, "per-CTA B rows must be a multiple of 8");
    NFB_CHECK(b.rows >= n_umma, "B plane has fewer rows than n_umma");
    NFB_TRY(make_plane_map(&tm_b_hi, b.hi, b.rows, k_total, b.ld, kBlockK, n_load));
    NFB_TRY(make_plane_map(&tm_b_lo, b.lo, b.rows, k_total, b.ld, kBlockK, n_load));

    GemmArgs args;
    args.out = out;
    args.col_scale = col_scale;
    args.row_scale = row_scale;
    args.scale = scale;
    args.m_total = static_cast<int>(m_total);
    args.n_umma = n_umma;
    args.n_store = n_store;
    args.num_kb = static_cast<int>(num_kb);
    args.splits = splits;
    args.m_blocks = static_cast<int>(ceil_div64(m_total, static_cast<int64_t>(kBlockM) * cta_group));
    args.ld_out = static_cast<int>(ld_out);
    args.split_stride = split_stride;
    args.scatter_slice = 0;
    args.scatter_rows = args.scatter_src = args.scatter_col0 = 0;
    for (int p = 0; p < kGemmMaxPeers; ++p) args.scatter_base[p] = nullptr;
    args.perm_groups = args.perm_bpo = args.perm_rot = 0;
    args.num_tiles = args.m_blocks * splits;
    if (scatter != nullptr && scatter->slice > 0) {
        NFB_CHECK(splits == 1, "the fused reduce-scatter epilogue needs an unsplit GEMM");
        NFB_CHECK(ceil_div64(m_total, scatter->slice) <= kGemmMaxPeers, "more owners than peers in the scatter epilogue");
        args.scatter_slice = static_cast<int>(scatter->slice);
        args.scatter_rows = scatter->rows;
        args.scatter_src = scatter->src;
        args.scatter_col0 = col0;
        args.ld_out = static_cast<int>(scatter->slice);
        for (int p = 0; p < kGemmMaxPeers; ++p) args.scatter_base[p] = scatter->base[p];
        const int owners = static_cast<int>(ceil_div64(m_total, scatter->slice));
        static const bool permute = !(getenv("NFB_SCATTER_PERM") && atoi(getenv("NFB_SCATTER_PERM")) == 0);
        if (permute && owners > 1 && args.m_blocks >= 2 * owners) {
            args.perm_groups = owners;
            args.perm_bpo = (args.m_blocks + owners - 1) / owners;
            args.perm_rot = scatter->src % owners;
            args.num_tiles = args.perm_groups * args.perm_bpo;
        }
    }
    args.b_tile_bytes = n_load * 128;
    const int stage_bytes = 2 * kATileBytes + 2 * args.b_tile_bytes;
    const int smem_budget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/;
    args.num_stages = std::min(kMaxStages, smem_budget / stage_bytes);
    NFB_CHECK(args.num_stages >= 2, "not enough shared memory for a 2-stage pipeline");
    const size_t smem_bytes = static_cast<size_t>(args.num_stages) * stage_bytes + 1024 + 256;

    const int64_t tiles = args.num_tiles;
    int workers = static_cast<int>(std::min<int64_t>(tiles, sm_count() / cta_group));
    args.sk_full = args.sk_tail = args.sk_share = 0;
    args.sk_ws = nullptr;
    args.sk_cnt = nullptr;
    if (sk != nullptr && (sk->force || gemm_streamk_enabled()) && splits == 1 && args.scatter_slice == 0 &&
        args.perm_groups == 0) {
        const int w_all = sm_count() / cta_group;
        const size_t need = static_cast<size_t>(w_all) * n_store * cta_group * kBlockM;
        int full = 0, tail = 0, share = 0;
        if (w_all <= sk->workers && need <= sk->ws_floats && streamk_plan(m_total, k_total, cta_group, &full, &tail, &share)) {
            args.sk_full = full;
            args.sk_tail = tail;
            args.sk_share = share;
            args.sk_ws = sk->ws;
            args.sk_cnt = sk->cnt;
            workers = w_all;
        }
    }

    Launch l{&tm_a_hi, &tm_a_lo, &tm_b_hi, &tm_b_lo, &args, workers * cta_group, smem_bytes, stream};
    const int nch_per_thread = ((n_umma >> 4) + 1) >> 1;
    if (a_is_mn_major) {
        if (cta_group == 2) return launch_nch<1, 2>(l, nch_per_thread);
        return launch_nch<1, 1>(l, nch_per_thread);
    }
    if (cta_group == 2) return launch_nch<0, 2>(l, nch_per_thread);
    return launch_nch<0, 1>(l, nch_per_thread);
}

// N (the component dimension) is tiled in slabs of <= 256 columns: one launch per slab, each streaming A once.
int gemm_split(const GemmOperand& a, bool a_is_mn_major, const GemmOperand& b, int64_t m_total, int64_t k_total,
               int n_umma, int n_store, float* out, int64_t ld_out, int splits, const float* col_scale,
               const float* row_scale, float scale, int cta_group, cudaStream_t stream, const GemmScatter* scatter,
               const GemmStreamK* sk) {
    NFB_CHECK(n_umma % 16 == 0 && n_umma >= 16, "n_umma must be a positive multiple of 16");
    const long long split_stride = static_cast<long long>(n_store) * ld_out;
    // NFB_GEMM_TIMING=1: device time of every call (debug aid; synchronises the stream)
    static const bool timing = getenv("NFB_GEMM_TIMING") != nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timing && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) cudaEventRecord(e0, stream);
    struct Report {
        cudaEvent_t e0, e1; cudaStream_t st; int64_t m, k; int n, splits, cg; bool mn;
        ~Report() {
            if (!e0 || !e1) return;
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, e0, e1);
            fprintf(stderr, "nfact_b200 gemm: M=%lld K=%lld N=%d splits=%d cta_group=%d a_mn_major=%d  %.3f ms\n",
                    static_cast<long long>(m), static_cast<long long>(k), n, splits, cg, mn ? 1 : 0, ms);
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
    } report{e0, e1, stream, m_total, k_total, n_umma, splits, cta_group, a_is_mn_major};
    for (int t0 = 0; t0 < n_umma && t0 < n_store; t0 += 256) {
        const int n_tile = std::min(256, n_umma - t0);
        GemmOperand bt{b.hi + static_cast<int64_t>(t0) * b.ld, b.lo + static_cast<int64_t>(t0) * b.ld, b.rows - t0,
                       b.ld};
        NFB_TRY(gemm_split_tile(a, a_is_mn_major, bt, m_total, k_total, n_tile, std::min(n_tile, n_store - t0),
                                out + static_cast<int64_t>(t0) * ld_out, ld_out, split_stride, splits,
                                col_scale ? col_scale + t0 : nullptr, row_scale, scale, cta_group, stream, scatter, t0,
                                sk));
    }
    return 0;
}

}  // namespace nfb
