// pgx_gm.cu -- kernels over the group-major ("path-major") copy of the abacus bitmap:
//   * k_transpose      node-major rows -> one bit-row per group (64 items per u64 word)
//   * k_gm_growth      ordered growth under an arbitrary group order (permuted growth, any quorum):
//                      OR-accumulate rows in the given order; bit-sliced per-item rank counters
//                      compared against the per-position quorum threshold
//                      (same arithmetic as AbacusByGroup::calc_growth, abacus.rs:989-1032, applied to
//                      the abacus the reference would rebuild under `--order`, abacus.rs:324-326)
//   * k_gm_similarity  all-pairs group intersections, AND + POPC over item words
//                      (integer part of Similarity::set_table, src/analyses/similarity.rs:125-150)
//   * k_weight_planes  bit-planes of the u32 item weights (bp-weighted intersections)
//   * k_scatter        ItemTable slice -> bitmap bits (abacus.rs:719-744 de-duplication = idempotent OR)
#include <cstdlib>
#include <cstring>

#include <cub/device/device_radix_sort.cuh>

#include "pgx_common.cuh"
#include "pgx_internal.h"

namespace pgx {

uint64_t gm_stride_words(uint64_t n_rows) {
    const uint64_t w = (n_rows + 63u) / 64u;
    return (w + 15u) / 16u * 16u;
}

namespace {

// ---- transpose -----------------------------------------------------------------------------------
// A CTA stages 256 items x COLS word-columns of the node-major bitmap in shared memory with coalesced
// 128-bit loads (row pitch padded by one word: the column reads below are then conflict free for 64-bit
// accesses); warp w then turns word-column w into 64 group rows x 8 u32 (32 contiguous bytes per row =
// one full sector per store).  The 32x32 bit blocks are transposed across the lanes of a warp with the
// 5-stage butterfly: per stage one SHFL and one PRMT (the byte-granular stages) or one SHF + one LOP3 (the
// bit stages), all per-lane constants hoisted -- the first version spent ~6 ALU operations per stage and
// was ALU-pipe bound (1.04 ms for 10M x 1024; a variant with 1024-item tiles and full 128-byte output
// lines but 8 warps per CTA was slower still, 1.19 ms: occupancy, not the write granularity, is what
// counts here -- profiles/r2_transpose_*.jsonl).
constexpr int kTrItems = 256;

struct TrLane {  // per-lane constants of the butterfly
    uint32_t sel16, sel8;          // PRMT selectors of the 16- and 8-bit stages
    uint32_t keep4, keep2, keep1;  // bits that stay in place in the 4-, 2-, 1-bit stages
    uint32_t rot4, rot2, rot1;     // left-rotation of the partner's word
};

__device__ __forceinline__ TrLane tr_lane(uint32_t lane) {
    TrLane t;
    t.sel16 = (lane & 16u) ? 0x3276u : 0x5410u;
    t.sel8 = (lane & 8u) ? 0x3715u : 0x6240u;
    t.keep4 = (lane & 4u) ? 0xF0F0F0F0u : 0x0F0F0F0Fu;
    t.keep2 = (lane & 2u) ? 0xCCCCCCCCu : 0x33333333u;
    t.keep1 = (lane & 1u) ? 0xAAAAAAAAu : 0x55555555u;
    t.rot4 = (lane & 4u) ? 28u : 4u;
    t.rot2 = (lane & 2u) ? 30u : 2u;
    t.rot1 = (lane & 1u) ? 31u : 1u;
    return t;
}

// lane r holds row r of a 32x32 bit matrix in x; on return lane c holds column c (bit r = old bit c of lane r)
__device__ __forceinline__ uint32_t transpose32(uint32_t x, const TrLane &t) {
    uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, x, 16);
    x = __byte_perm(x, y, t.sel16);
    y = __shfl_xor_sync(0xFFFFFFFFu, x, 8);
    x = __byte_perm(x, y, t.sel8);
    y = __shfl_xor_sync(0xFFFFFFFFu, x, 4);
    y = __funnelshift_l(y, y, t.rot4);
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(x) : "r"(t.keep4), "r"(x), "r"(y));  // keep ? x : y
    y = __shfl_xor_sync(0xFFFFFFFFu, x, 2);
    y = __funnelshift_l(y, y, t.rot2);
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(x) : "r"(t.keep2), "r"(x), "r"(y));
    y = __shfl_xor_sync(0xFFFFFFFFu, x, 1);
    y = __funnelshift_l(y, y, t.rot1);
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(x) : "r"(t.keep1), "r"(x), "r"(y));
    return x;
}

template <int COLS>
__global__ void __launch_bounds__(COLS * 32) k_transpose(const uint64_t *__restrict__ bitmap, uint64_t n_rows, uint32_t G,
                                                         uint32_t W, uint32_t Wp, uint32_t *__restrict__ gm32,
                                                         uint64_t gm_stride32, const uint32_t *__restrict__ perm) {
    // perm != nullptr: bit position i of the output rows holds item perm[i] (weight-sorted copy for similarity)
    constexpr int kPitch = COLS + 1, kThreads = COLS * 32;
    __shared__ uint64_t tile[kTrItems * kPitch];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t item0 = (uint64_t)blockIdx.x * kTrItems;
    const uint32_t wc0 = blockIdx.y * COLS;
    const uint32_t ncols = min((uint32_t)COLS, Wp - wc0);  // word-columns staged by this CTA (Wp is even or 1)
    if (ncols >= 2u && !(ncols & 1u)) {
        const uint32_t chunks = ncols >> 1;  // 16-byte chunks per row
        for (uint32_t e = tid; e < kTrItems * chunks; e += kThreads) {
            const uint32_t r = e / chunks, c = e - r * chunks;
            uint64_t item = item0 + r;
            if (perm && item < n_rows) item = __ldg(perm + item);
            uint64_t x0 = 0, x1 = 0;
            if (item != 0 && item < n_rows) {
                const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(bitmap + item * Wp + wc0) + c);
                x0 = v.x;
                x1 = v.y;
            }
            tile[r * kPitch + 2u * c] = x0;
            tile[r * kPitch + 2u * c + 1u] = x1;
        }
    } else {
        for (uint32_t e = tid; e < kTrItems * ncols; e += kThreads) {
            const uint32_t r = e / ncols, c = e - r * ncols;
            uint64_t item = item0 + r;
            if (perm && item < n_rows) item = __ldg(perm + item);
            tile[r * kPitch + c] = (item != 0 && item < n_rows) ? __ldg(bitmap + item * Wp + wc0 + c) : 0ull;
        }
    }
    __syncthreads();
    const uint32_t wc = wc0 + warp;  // word column of the node-major row handled by this warp
    if (warp >= ncols || wc >= W) return;
    const TrLane tl = tr_lane(lane);
    uint32_t keep0[8], keep1[8];
#pragma unroll
    for (int ib = 0; ib < 8; ++ib) {
        const uint64_t x = tile[(ib * 32 + lane) * kPitch + warp];
        keep0[ib] = transpose32((uint32_t)x, tl);
        keep1[ib] = transpose32((uint32_t)(x >> 32), tl);
    }
    // lane b owns groups wc*64 + b and wc*64 + 32 + b; 8 u32 = items item0 .. item0+255
    const uint64_t col32 = item0 / 32u;
    const uint32_t g0 = wc * 64u + lane, g1 = g0 + 32u;
    if (g0 < G) {
        uint4 *dst = reinterpret_cast<uint4 *>(gm32 + (uint64_t)g0 * gm_stride32 + col32);
        dst[0] = make_uint4(keep0[0], keep0[1], keep0[2], keep0[3]);
        dst[1] = make_uint4(keep0[4], keep0[5], keep0[6], keep0[7]);
    }
    if (g1 < G) {
        uint4 *dst = reinterpret_cast<uint4 *>(gm32 + (uint64_t)g1 * gm_stride32 + col32);
        dst[0] = make_uint4(keep1[0], keep1[1], keep1[2], keep1[3]);
        dst[1] = make_uint4(keep1[4], keep1[5], keep1[6], keep1[7]);
    }
}

// ---- transpose, second version: 32x32 bit blocks transposed inside one thread ---------------------------------------
// The shuffle butterfly above is bound by the MIO queue (10 SHFL per 64-bit word next to the tile's STS / LDS).  Here a
// thread owns one 32x32 block -- 32 items x one 32-bit column of the node-major rows -- reads it from the staged tile
// with 32 conflict-free LDS.32, transposes it in registers (5 stages x 16 word pairs: 2 PRMT for the 16- and 8-bit
// stages, 2 SHF + 2 LOP3 for the bit stages = 256 ALU operations, no shuffles) and stores word b straight to group row
// 32 c + b.  The lanes of a warp own CONSECUTIVE 32-item blocks of the same column, so every warp store writes 128
// contiguous bytes of one group row (64 for 128-byte-wide tiles, where a warp covers two columns).
// Tile: 16384 32-bit words = 32 IB items x WC columns, WC = 2^WC_LOG 32-bit columns (8 .. 128 bytes of a row), IB item
// blocks; 512 threads, one block each.  Lanes 32 rows apart would all hit one bank, so word (R, c) of the tile lives at
// L ^ s(R / 32) with L = R WC + c: the XOR only permutes the 32 words of a 128-byte line (all of one item block), and
// the lanes' banks become (const ^ s(ib)) -- all different.
template <int WC_LOG>
__global__ void __launch_bounds__(512, 2) k_transpose_reg(const uint64_t *__restrict__ bitmap, uint64_t n_rows, uint32_t G,
                                                          uint32_t Wp, uint32_t *__restrict__ gm32, uint64_t gm_stride32,
                                                          const uint32_t *__restrict__ perm) {
    constexpr uint32_t WC = 1u << WC_LOG;          // 32-bit columns per tile row
    constexpr uint32_t IB = 512u >> WC_LOG;        // 32-item blocks per tile
    constexpr uint32_t kRows = IB * 32u;
    extern __shared__ __align__(16) uint32_t tile[];  // 16384 words
    const uint32_t tid = threadIdx.x;
    const uint64_t item0 = (uint64_t)blockIdx.x * kRows;
    const uint32_t col0 = blockIdx.y * WC;          // first 32-bit column of this tile
    const uint32_t row_cols = Wp * 2u;              // 32-bit columns per bitmap row
    auto swz = [](uint32_t ib) -> uint32_t { return IB >= 32u ? (ib & 31u) : ((ib << 1) & 31u); };

    // ---- stage: coalesced 16-byte (8-byte for one-word rows) loads -> swizzled shared-memory tile ----
    if (WC_LOG >= 2) {
        constexpr uint32_t CH = WC / 4u;            // 16-byte chunks per tile row
#pragma unroll
        for (uint32_t k = 0; k < 8u; ++k) {
            const uint32_t e = tid + k * 512u;
            const uint32_t R = e / CH, q = e % CH;
            uint64_t item = item0 + R;
            if (perm && item < n_rows) item = __ldg(perm + item);
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (item != 0 && item < n_rows && col0 + 4u * q < row_cols)
                v = __ldg(reinterpret_cast<const uint4 *>(bitmap + item * Wp) + (col0 / 4u + q));
            const uint32_t s = swz(R >> 5);
            if (s & 1u) {
                uint32_t t = v.x; v.x = v.y; v.y = t;
                t = v.z; v.z = v.w; v.w = t;
            }
            if (s & 2u) {
                uint32_t t = v.x; v.x = v.z; v.z = t;
                t = v.y; v.y = v.w; v.w = t;
            }
            *reinterpret_cast<uint4 *>(tile + ((R * WC + 4u * q) ^ (s & 28u))) = v;
        }
    } else {
#pragma unroll
        for (uint32_t k = 0; k < 16u; ++k) {
            const uint32_t R = tid + k * 512u;     // one 8-byte row per element
            uint64_t item = item0 + R;
            if (perm && item < n_rows) item = __ldg(perm + item);
            uint2 v = make_uint2(0u, 0u);
            if (item != 0 && item < n_rows) v = __ldg(reinterpret_cast<const uint2 *>(bitmap + item * Wp));
            const uint32_t s = swz(R >> 5);
            if (s & 1u) {
                const uint32_t t = v.x; v.x = v.y; v.y = t;
            }
            *reinterpret_cast<uint2 *>(tile + ((R * 2u) ^ (s & 30u))) = v;
        }
    }
    __syncthreads();

    // ---- one 32x32 block per thread ----
    const uint32_t ib = tid % IB, c = tid / IB;
    const uint32_t s = swz(ib);
    uint32_t a[32];
#pragma unroll
    for (uint32_t r = 0; r < 32u; ++r) {
        const uint32_t L = (32u * ib + r) * WC + c;
        a[r] = tile[L ^ s];
    }
    // rows -> columns: swap the off-diagonal blocks of size 16, 8 (byte permutes), 4, 2, 1 (funnel shift + select)
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint32_t x = a[k], y = a[k + 16];
        a[k] = __byte_perm(x, y, 0x5410);
        a[k + 16] = __byte_perm(x, y, 0x7632);
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        if (k & 8) continue;
        const uint32_t x = a[k], y = a[k + 8];
        a[k] = __byte_perm(x, y, 0x6240);
        a[k + 8] = __byte_perm(x, y, 0x7351);
    }
#pragma unroll
    for (int sh = 4; sh >= 1; sh >>= 1) {
        const uint32_t m = sh == 4 ? 0x0F0F0F0Fu : (sh == 2 ? 0x33333333u : 0x55555555u);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if (k & sh) continue;
            const uint32_t x = a[k], y = a[k + sh];
            asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(a[k]) : "r"(m), "r"(x), "r"(y << sh));       // m ? x : y << sh
            asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(a[k + sh]) : "r"(m), "r"(x >> sh), "r"(y));  // m ? x >> sh : y
        }
    }
    // a[b]: bit r = item 32 (item block) + r of group 32 (col0 + c) + b
    const uint64_t ibg = item0 / 32u + ib;
    const uint32_t g0 = (col0 + c) * 32u;
    if (ibg < gm_stride32 && g0 < G) {
        uint32_t *dst = gm32 + (uint64_t)g0 * gm_stride32 + ibg;
        if (g0 + 32u <= G) {
#pragma unroll
            for (uint32_t b = 0; b < 32u; ++b, dst += gm_stride32) *dst = a[b];
        } else {
#pragma unroll
            for (uint32_t b = 0; b < 32u; ++b, dst += gm_stride32)
                if (g0 + b < G) *dst = a[b];
        }
    }
}

// ---- group-major growth -----------------------------------------------------------------------
constexpr int kGmThreads = 256;
constexpr int kRankPlanes = 21;  // supports G up to 2^20

__device__ __forceinline__ uint64_t weighted_bits(uint64_t m, const uint32_t *__restrict__ wrow) {
    uint64_t s = 0;
    while (m) {
        const uint32_t b = (uint32_t)__ffsll((long long)m) - 1u;
        m &= m - 1;
        s += __ldg(wrow + b);
    }
    return s;
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
    // values are < 2^44 here (<= 64 weights of < 2^32 per lane ... summed over 32 lanes)
    const uint32_t lo = (uint32_t)(v & 0xFFFFFFu), mid = (uint32_t)((v >> 24) & 0xFFFFFFu), hi = (uint32_t)(v >> 48);
    const uint64_t slo = __reduce_add_sync(0xFFFFFFFFu, lo);
    const uint64_t smid = __reduce_add_sync(0xFFFFFFFFu, mid);
    const uint64_t shi = __reduce_add_sync(0xFFFFFFFFu, hi);
    return slo + (smid << 24) + (shi << 48);
}

// One thread owns 64 items (one u64 column of the group-major bitmap) and walks the groups in the given
// order.  q = 0 thresholds: an item starts counting at its first group -> popcount of the newly seen
// bits (HBM-bound; rows are prefetched kGmPrefetch steps ahead).  General thresholds: bit-sliced rank
// counters R (P planes, one bit per item) are incremented by the row and compared against the uniform
// per-position cutoff K = thr[t][j] with one 3-input LOP per plane; the verdict of an item only changes
// at its own set bits (abacus.rs:1007-1010), and the curve's first difference is the net number (or
// weight) of verdict flips.
template <int P, bool GENERAL, int TMAX>  // P: rank bit-planes (2^P > G); TMAX: compile-time bound on p.T
__global__ void __launch_bounds__(kGmThreads, GENERAL ? 3 : 4) k_gm_growth(const __grid_constant__ GmGrowthParams p) {
    constexpr int kGmPrefetch = GENERAL ? 2 : 8;  // rows in flight per thread (the q = 0 kernel is HBM-bound)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // smem: order[G] u32 | thr[T*G] u32 (if any general) | delta[T*G] u64
    uint32_t *s_order = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *s_thr = s_order + p.G;
    const uint32_t thr_words = GENERAL ? p.T * p.G : 0u;
    unsigned long long *s_delta =
        reinterpret_cast<unsigned long long *>(smem_raw + (((size_t)(p.G + thr_words) * 4u + 15u) & ~(size_t)15u));
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    // 1-D grid, order fastest: the CTAs resident at any time cover a few 64-item columns under many orders, so all but
    // the first read of a column block hit L2 (the group-major copy of a large table does not fit it as a whole)
    const uint32_t n_col_blocks = (uint32_t)((p.n_words + kGmThreads - 1) / kGmThreads);
    const uint32_t order_id = p.col_fastest ? blockIdx.x / n_col_blocks : blockIdx.x % p.n_orders;
    const uint64_t col_block = p.col_fastest ? blockIdx.x % n_col_blocks : blockIdx.x / p.n_orders;
    const uint32_t *order = p.order + (size_t)order_id * p.G;
    for (uint32_t i = tid; i < p.G; i += kGmThreads) s_order[i] = order[i];
    if (GENERAL)
        for (uint32_t i = tid; i < p.T * p.G; i += kGmThreads) s_thr[i] = p.thr[i];
    if (!p.direct_out)
        for (uint32_t i = tid; i < p.T * p.G; i += kGmThreads) s_delta[i] = 0ull;
    __syncthreads();

    const uint64_t wi = col_block * kGmThreads + tid;
    const bool active = wi < p.n_words;
    const uint64_t wsafe = active ? wi : 0;
    const uint32_t *wrow = (p.weighted && p.weight) ? p.weight + wsafe * 64u : nullptr;

    // eligibility masks: item counted for threshold t only if its total coverage >= cov[t]
    uint64_t elig[TMAX];
#pragma unroll
    for (int t = 0; t < TMAX; ++t) elig[t] = ~0ull;
    bool need_cov = false;
    for (uint32_t t = 0; t < p.T; ++t) need_cov |= p.cov[t] > 1u;
    if (need_cov && active) {
#pragma unroll
        for (int t = 0; t < TMAX; ++t) elig[t] = 0ull;
        for (uint32_t b = 0; b < 64u; ++b) {
            const uint64_t item = wi * 64u + b;
            const uint32_t c = (item < p.n_rows && item != 0) ? __ldg(p.countable + item) : 0u;
#pragma unroll
            for (int t = 0; t < TMAX; ++t)
                if ((uint32_t)t < p.T && c >= p.cov[t]) elig[t] |= 1ull << b;
        }
    }

    uint64_t seen = 0;
    uint64_t R[P];
#pragma unroll
    for (int i = 0; i < P; ++i) R[i] = 0ull;
    uint64_t verdict[TMAX];
#pragma unroll
    for (int t = 0; t < TMAX; ++t) verdict[t] = 0ull;

    const uint64_t *col = p.gm + wsafe;
    auto load_row = [&](uint32_t j) -> uint64_t {
        return (active && j < p.G) ? __ldg(col + (uint64_t)s_order[j] * p.gm_stride) : 0ull;
    };
    uint64_t nxt[kGmPrefetch];
#pragma unroll
    for (int u = 0; u < kGmPrefetch; ++u) nxt[u] = load_row((uint32_t)u);

    for (uint32_t j0 = 0; j0 < p.G; j0 += kGmPrefetch) {
        uint64_t cur[kGmPrefetch];
#pragma unroll
        for (int u = 0; u < kGmPrefetch; ++u) cur[u] = nxt[u];
#pragma unroll
        for (int u = 0; u < kGmPrefetch; ++u) nxt[u] = load_row(j0 + kGmPrefetch + (uint32_t)u);  // in flight during the math
#pragma unroll
        for (int u = 0; u < kGmPrefetch; ++u) {
            const uint32_t j = j0 + (uint32_t)u;
            if (j >= p.G) break;
            const uint64_t b = cur[u];
            const uint64_t fresh = b & ~seen;
            seen |= b;
            if (GENERAL) {  // R += b (bit-sliced ripple increment)
                uint64_t carry = b;
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const uint64_t t2 = R[i] & carry;
                    R[i] ^= carry;
                    carry = t2;
                }
            }
#pragma unroll
            for (int t = 0; t < TMAX; ++t) {
                if ((uint32_t)t >= p.T) break;
                uint64_t up, down = 0ull;
                if (GENERAL && ((p.general_mask >> t) & 1u)) {
                    // ge = (R >= K) per item, K uniform: scan the planes from the LSB,
                    // ge <- K_i ? (ge & R_i) : (ge | R_i)  == one 3-input LOP with m = -K_i
                    const uint32_t K = s_thr[t * p.G + j];
                    uint64_t ge = ~0ull;
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        const uint64_t m = 0ull - (uint64_t)((K >> i) & 1u);
                        ge = (m & (ge & R[i])) | (~m & (ge | R[i]));
                    }
                    if (K >> P) ge = 0ull;
                    const uint64_t vnew = (b & ge) | (~b & verdict[t]);
                    up = vnew & ~verdict[t];
                    down = verdict[t] & ~vnew;
                    verdict[t] = vnew;
                } else {
                    up = fresh;
                }
                up &= elig[t];
                down &= elig[t];
                long long net;
                if (p.weighted) {
                    const uint64_t su = warp_sum_u64(wrow ? weighted_bits(up, wrow) : (uint64_t)__popcll(up));
                    uint64_t sd = 0;
                    if (GENERAL) sd = warp_sum_u64(wrow ? weighted_bits(down, wrow) : (uint64_t)__popcll(down));
                    net = (long long)(su - sd);
                } else {
                    const int diff = __popcll(up) - (GENERAL ? __popcll(down) : 0);
                    net = (long long)__reduce_add_sync(0xFFFFFFFFu, diff);
                }
                if (lane == 0 && net != 0) {
                    if (p.direct_out)
                        atomicAdd(reinterpret_cast<unsigned long long *>(p.out + (size_t)order_id * p.out_order_stride +
                                                                         (size_t)p.slot[t] * p.G + j),
                                  (unsigned long long)net);
                    else
                        atomicAdd(&s_delta[t * p.G + j], (unsigned long long)net);
                }
            }
        }
    }
    __syncthreads();
    if (p.direct_out) return;
    uint64_t *out = p.out + (size_t)order_id * p.out_order_stride;
    for (uint32_t i = tid; i < p.T * p.G; i += kGmThreads) {
        const unsigned long long v = s_delta[i];
        if (v) {
            const uint32_t t = i / p.G, j = i - t * p.G;
            atomicAdd(reinterpret_cast<unsigned long long *>(out + (size_t)p.slot[t] * p.G + j), v);
        }
    }
}

// ---- similarity -----------------------------------------------------------------------------------
constexpr int kSimTile = 64;    // groups per tile edge
constexpr int kSimKW = 32;      // item words per smem stage
constexpr int kSimPad = kSimTile + 2;

// one LOP3 each (nvcc splits the three-term majority into two ops)
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// CSA (unweighted only): the AND words of two consecutive item words go through a carry-save adder per group pair
// (ones' = ones ^ a1 ^ a2, carry = maj(ones, a1, a2)); only the carry word is popcounted (weight 2) and the `ones`
// word once at the very end.  Halves the POPC count -- the kernel's bound, 16 lanes/clk/SM against 64 for LOP3 -- for
// 4 extra logic ops per pair and two words, which roughly balances the two pipes.
template <bool WEIGHTED, bool CSA>
__global__ void __launch_bounds__(256) k_gm_similarity(const __grid_constant__ GmSimParams p, uint32_t words_per_split) {
    static_assert(!(WEIGHTED && CSA), "the carry-save variant counts unweighted intersections");
    __shared__ uint64_t Xs[kSimKW][kSimPad];
    __shared__ uint64_t Ys[kSimKW][kSimPad];
    __shared__ uint64_t Ps[32][kSimKW];
    __shared__ uint64_t Us[kSimKW];   // (1 << 32) | w if all 64 items of the word share weight w, else 0
    __shared__ uint32_t Ms[kSimKW];   // non-empty weight planes of the word
    const uint32_t tid = threadIdx.x;
    const uint32_t tx = tid & 15u, ty = tid >> 4;
    uint32_t bx = blockIdx.x, by = blockIdx.y;
    if (p.triangular) {  // full square requested: only tiles on or above the diagonal (bx >= by), mirrored afterwards
        const uint32_t nt = (p.G + kSimTile - 1u) / kSimTile;  // tiles per edge; blockIdx.x enumerates the upper tiles
        uint32_t lin = blockIdx.x;
        by = 0;
        while (lin >= nt - by) {
            lin -= nt - by;
            ++by;
        }
        bx = by + lin;
    }
    const uint32_t x0 = p.row_begin + by * kSimTile;  // output rows
    const uint32_t y0 = p.col_begin + bx * kSimTile;  // output columns (col_begin > 0: upper-triangle sharding)
    if (p.upper_only && y0 + kSimTile <= x0) return;  // tile strictly below the diagonal: its mirror image is computed instead
    const uint64_t k_begin = (uint64_t)blockIdx.z * words_per_split;
    uint64_t k_end = k_begin + words_per_split;
    if (k_end > p.n_words) k_end = p.n_words;

    uint64_t acc[4][4];
    const uint32_t one = p.csa;  // == 1 whenever CSA is instantiated
    uint32_t ones_lo[CSA ? 4 : 1][CSA ? 4 : 1], ones_hi[CSA ? 4 : 1][CSA ? 4 : 1];  // CSA: pending weight-1 bits per pair
    uint32_t twos[CSA ? 4 : 1][CSA ? 4 : 1];  // CSA: number of carries (weight 2); < 2^31 for any n_items < 2^32
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc[i][j] = 0ull;
            if (CSA) {
                ones_lo[i][j] = ones_hi[i][j] = 0u;
                twos[i][j] = 0u;
            }
        }

    for (uint64_t k0 = k_begin; k0 < k_end; k0 += kSimKW) {
        // stage 64 x-rows and 64 y-rows x 32 words (coalesced along words), stored word-major
        for (uint32_t e = tid; e < kSimTile * kSimKW; e += 256u) {
            const uint32_t r = e / kSimKW, kw = e % kSimKW;
            const uint64_t k = k0 + kw;
            const uint32_t gx = x0 + r, gy = y0 + r;
            Xs[kw][r] = (gx < p.row_end && k < k_end) ? __ldg(p.gm + (uint64_t)gx * p.gm_stride + k) : 0ull;
            Ys[kw][r] = (gy < p.G && k < k_end) ? __ldg(p.gm + (uint64_t)gy * p.gm_stride + k) : 0ull;
        }
        if (WEIGHTED) {
            for (uint32_t e = tid; e < p.n_planes * kSimKW; e += 256u) {
                const uint32_t pl = e / kSimKW, kw = e % kSimKW;
                const uint64_t k = k0 + kw;
                Ps[pl][kw] = (k < k_end) ? __ldg(p.planes + (uint64_t)pl * p.gm_stride + k) : 0ull;
            }
            if (tid < kSimKW) {
                const uint64_t k = k0 + tid;
                Us[tid] = (k < k_end) ? __ldg(p.uniform_w + k) : (1ull << 32);
                Ms[tid] = (k < k_end) ? __ldg(p.plane_mask + k) : 0u;
            }
        }
        __syncthreads();
        if (CSA) {
#pragma unroll 2
            for (uint32_t kw = 0; kw < kSimKW; kw += 2) {
                uint64_t xa[4], ya[4], xb[4], yb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    xa[i] = Xs[kw][ty * 4 + i];
                    ya[i] = Ys[kw][tx * 4 + i];
                    xb[i] = Xs[kw + 1][ty * 4 + i];
                    yb[i] = Ys[kw + 1][tx * 4 + i];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint64_t a1 = xa[i] & ya[j], a2 = xb[i] & yb[j];
                        const uint32_t a1l = (uint32_t)a1, a1h = (uint32_t)(a1 >> 32), a2l = (uint32_t)a2, a2h = (uint32_t)(a2 >> 32);
                        // accumulate with IMAD (x * 1 + acc, multiplier from a kernel parameter so it stays a multiply): the
                        // FMA pipe is idle here while LOP3 saturates the ALU pipe an IADD3 would share
                        twos[i][j] = (uint32_t)__popc(lop3_maj(ones_lo[i][j], a1l, a2l)) * one + twos[i][j];
                        twos[i][j] = (uint32_t)__popc(lop3_maj(ones_hi[i][j], a1h, a2h)) * one + twos[i][j];
                        ones_lo[i][j] = lop3_xor3(ones_lo[i][j], a1l, a2l);
                        ones_hi[i][j] = lop3_xor3(ones_hi[i][j], a1h, a2h);
                    }
            }
            __syncthreads();
            continue;
        }
#pragma unroll 4
        for (uint32_t kw = 0; kw < kSimKW; ++kw) {
            uint64_t xv[4], yv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                xv[i] = Xs[kw][ty * 4 + i];
                yv[i] = Ys[kw][tx * 4 + i];
            }
            if (!WEIGHTED) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] += (uint64_t)__popcll(xv[i] & yv[j]);
            } else {
                const uint64_t u = Us[kw];
                if (u) {  // items are sorted by weight: most words carry one weight -> a single popcount pass
                    const uint64_t w = u & 0xFFFFFFFFull;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] += (uint64_t)__popcll(xv[i] & yv[j]) * w;
                } else {
                    for (uint32_t m = Ms[kw]; m; m &= m - 1u) {
                        const uint32_t pl = (uint32_t)__ffs((int)m) - 1u;
                        const uint64_t pw = Ps[pl][kw];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] += (uint64_t)__popcll(xv[i] & yv[j] & pw) << pl;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (CSA) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 2ull * twos[i][j] + (uint64_t)(__popc(ones_lo[i][j]) + __popc(ones_hi[i][j]));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t gx = x0 + ty * 4 + i;
        if (gx >= p.row_end) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t gy = y0 + tx * 4 + j;
            if (gy < p.G && acc[i][j])
                atomicAdd(reinterpret_cast<unsigned long long *>(p.inter + (uint64_t)(gx - p.row_begin) * p.G + gy),
                          (unsigned long long)acc[i][j]);
        }
    }
}

// lower triangle <- upper triangle (the intersection matrix is symmetric)
__global__ void __launch_bounds__(256) k_sim_mirror(uint64_t *inter, uint32_t G) {
    const uint32_t x = blockIdx.y * 16u + (threadIdx.x >> 4), y = blockIdx.x * 16u + (threadIdx.x & 15u);
    if (x < G && y < G && y / kSimTile < x / kSimTile) inter[(uint64_t)x * G + y] = inter[(uint64_t)y * G + x];
}

// first differences -> curves in place (wrapping u64 prefix sums); one warp per curve of G entries
__global__ void __launch_bounds__(256) k_prefix_curves(uint64_t *d, uint64_t n_curves, uint32_t G) {
    const uint64_t curve = ((uint64_t)blockIdx.x * 256u + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (curve >= n_curves) return;
    uint64_t *row = d + curve * G;
    unsigned long long carry = 0;
    for (uint32_t j0 = 0; j0 < G; j0 += 32u) {
        const uint32_t j = j0 + lane;
        unsigned long long v = j < G ? row[j] : 0ull;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, v, o);
            if ((int)lane >= o) v += u;
        }
        v += carry;
        if (j < G) row[j] = v;
        carry = __shfl_sync(0xFFFFFFFFu, v, 31);
    }
}

// ---- per-group totals: len[g] = sum_i w_i [g in i] (similarity.rs:133-137) -------------------------
__global__ void __launch_bounds__(256) k_gm_rowsum(const uint64_t *__restrict__ gm, uint64_t gm_stride,
                                                   uint64_t n_words, const uint64_t *__restrict__ planes,
                                                   uint32_t n_planes, const uint64_t *__restrict__ uniform_w,
                                                   uint64_t *__restrict__ len) {
    const uint32_t g = blockIdx.x;
    const uint64_t *row = gm + (uint64_t)g * gm_stride;
    unsigned long long s = 0;
    for (uint64_t k = threadIdx.x; k < n_words; k += 256u) {
        const uint64_t x = __ldg(row + k);
        if (!planes) {
            s += (unsigned long long)__popcll(x);
        } else if (x) {
            const uint64_t u = uniform_w ? __ldg(uniform_w + k) : 0ull;
            if (u) {
                s += (unsigned long long)__popcll(x) * (u & 0xFFFFFFFFull);
            } else {
                for (uint32_t pl = 0; pl < n_planes; ++pl)
                    s += (unsigned long long)__popcll(x & __ldg(planes + (uint64_t)pl * gm_stride + k)) << pl;
            }
        }
    }
    __shared__ unsigned long long part[8];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31u) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; ++i) t += part[i];
        len[g] = t;
    }
}

// ---- weight bit-planes -----------------------------------------------------------------------------
// One warp per 64-item word of the (position-indexed) weight vector: plane k = bit k of every weight;
// uniform_w[word] = (1 << 32) | w when all valid items of the word share the weight w; plane_mask[word] =
// which planes are non-empty.  skip0: position 0 is the dummy item (natural order).
__global__ void __launch_bounds__(256) k_weight_planes(const uint32_t *__restrict__ weight, uint64_t n_rows,
                                                       uint64_t *__restrict__ planes, uint64_t stride,
                                                       uint32_t n_planes, uint64_t n_words_padded,
                                                       uint64_t *__restrict__ uniform_w, uint32_t *__restrict__ plane_mask,
                                                       int skip0) {
    const uint64_t word = ((uint64_t)blockIdx.x * 256u + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (word >= n_words_padded) return;
    const uint64_t i0 = word * 64u + lane, i1 = i0 + 32u;
    const bool v0 = i0 < n_rows && !(skip0 && i0 == 0), v1 = i1 < n_rows;
    const uint32_t w0 = v0 ? __ldg(weight + i0) : 0u, w1 = v1 ? __ldg(weight + i1) : 0u;
    uint32_t mask = 0;
    for (uint32_t k = 0; k < n_planes; ++k) {
        const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, (w0 >> k) & 1u), m1 = __ballot_sync(0xFFFFFFFFu, (w1 >> k) & 1u);
        if (lane == 0) planes[(uint64_t)k * stride + word] = (uint64_t)m0 | ((uint64_t)m1 << 32);
        if (m0 | m1) mask |= 1u << k;
    }
    // reference weight: the first valid item of the word
    const uint32_t valid0 = __ballot_sync(0xFFFFFFFFu, v0), valid1 = __ballot_sync(0xFFFFFFFFu, v1);
    uint32_t ref = 0;
    if (valid0) ref = __shfl_sync(0xFFFFFFFFu, w0, __ffs((int)valid0) - 1);
    else if (valid1) ref = __shfl_sync(0xFFFFFFFFu, w1, __ffs((int)valid1) - 1);
    const bool same = __all_sync(0xFFFFFFFFu, (!v0 || w0 == ref) && (!v1 || w1 == ref));
    if (lane == 0) {
        if (uniform_w) uniform_w[word] = same ? ((1ull << 32) | ref) : 0ull;
        if (plane_mask) plane_mask[word] = mask;
    }
}

__global__ void __launch_bounds__(256) k_sort_keys(const uint32_t *__restrict__ weight, uint64_t n_rows,
                                                   uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    if (i < n_rows) {
        keys[i] = i ? __ldg(weight + i) : 0u;  // the dummy item sorts with the zero weights
        vals[i] = (uint32_t)i;
    }
}

// ---- scatter build ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter(uint64_t *bitmap, uint32_t Wp, uint64_t n_rows,
                                                 const uint64_t *__restrict__ items, uint64_t n_steps,
                                                 uint32_t group_id, const uint8_t *__restrict__ exclude,
                                                 unsigned int *err) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t w = group_id >> 6;
    const unsigned long long bit = 1ull << (group_id & 63u);
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_steps; s += stride) {
        const uint64_t id = __ldg(items + s);
        if (id == 0 || id >= n_rows) {
            atomicOr(err, 1u);
            continue;
        }
        if (exclude && __ldg(exclude + id)) continue;
        unsigned long long *word = reinterpret_cast<unsigned long long *>(bitmap + id * Wp + w);
        if (!(*word & bit)) atomicOr(word, bit);  // repeated visits of a group count once (abacus.rs:736-741)
    }
}

// whole-ItemTable build: items[] holds steps [step0, step0 + n_steps) of the table (ids as u32 or as the reference's
// u64 ItemIdSize).  A thread's steps ascend, so the path that owns a step is found by one binary search for the
// thread's first step and a forward walk of the cursor afterwards (paths are long runs of steps); the ids of four
// iterations are loaded before the first is processed so that the scattered atomics overlap the streaming loads.
__device__ __forceinline__ uint64_t owner_path(const uint64_t *__restrict__ prefsum, uint64_t n_paths, uint64_t s) {
    uint64_t lo = 0, hi = n_paths;  // largest p with prefsum[p] <= s (empty paths share a boundary; the search lands on the owner)
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (__ldg(prefsum + mid) <= s) lo = mid; else hi = mid;
    }
    return lo;
}

template <typename IdT>
__global__ void __launch_bounds__(256) k_build(uint64_t *bitmap, uint32_t Wp, uint64_t n_rows, uint32_t G,
                                               const IdT *__restrict__ items, uint64_t step0, uint64_t n_steps,
                                               const uint64_t *__restrict__ prefsum, uint64_t n_paths,
                                               const int64_t *__restrict__ path_group,
                                               const uint8_t *__restrict__ exclude, unsigned int *err) {
    constexpr int kUnroll = 4;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_steps) return;
    uint64_t p = owner_path(prefsum, n_paths, step0 + k);
    uint64_t p_end = __ldg(prefsum + p + 1);
    long long grp = __ldg(path_group + p);
    for (; k < n_steps; k += kUnroll * stride) {
        uint64_t id[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint64_t ku = k + (uint64_t)u * stride;
            id[u] = ku < n_steps ? (uint64_t)__ldg(items + ku) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint64_t ku = k + (uint64_t)u * stride;
            if (ku >= n_steps) break;
            const uint64_t s = step0 + ku;
            if (s >= p_end) {  // next path(s): walk a few boundaries, search when the stride skipped many short paths
                int tries = 0;
                do {
                    ++p;
                    p_end = __ldg(prefsum + p + 1);
                } while (s >= p_end && ++tries < 8);
                if (s >= p_end) {
                    p = owner_path(prefsum, n_paths, s);
                    p_end = __ldg(prefsum + p + 1);
                }
                grp = __ldg(path_group + p);
            }
            if (grp < 0) continue;
            if ((unsigned long long)grp >= G) {
                atomicOr(err, 4u);
                continue;
            }
            if (id[u] == 0 || id[u] >= n_rows) {
                atomicOr(err, 1u);
                continue;
            }
            if (exclude && __ldg(exclude + id[u])) continue;
            unsigned long long *word = reinterpret_cast<unsigned long long *>(bitmap + id[u] * Wp + ((uint32_t)grp >> 6));
            const unsigned long long bit = 1ull << ((uint32_t)grp & 63u);
            if (!(*word & bit)) atomicOr(word, bit);  // repeated visits of a group count once (abacus.rs:736-741)
        }
    }
}

size_t gm_growth_smem(const GmGrowthParams &p) { return gm_growth_smem_bytes(p.G, p.T, p.general_mask != 0, p.direct_out != 0); }

template <int P, bool GENERAL, int TMAX>
int launch_gm_growth_t(const GmGrowthParams &p, cudaStream_t stream) {
    const size_t smem = gm_growth_smem(p);
    if (smem > 232448u) return fail(PGX_ERR_UNSUPPORTED, "group-major growth: G*T too large for shared memory");
    auto kern = k_gm_growth<P, GENERAL, TMAX>;
    PGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t blocks = (p.n_words + kGmThreads - 1) / kGmThreads * p.n_orders;
    if (blocks > 0x7FFFFFFFull) return fail(PGX_ERR_UNSUPPORTED, "group-major growth: too many column blocks x orders in one launch");
    kern<<<(unsigned)blocks, kGmThreads, smem, stream>>>(p);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

}  // namespace

size_t gm_growth_smem_bytes(uint32_t G, uint32_t T, bool any_general, bool direct_out) {
    const size_t thr_words = any_general ? (size_t)T * G : 0u;
    return (((size_t)(G + thr_words) * 4u + 15u) & ~(size_t)15u) + (direct_out ? 0u : (size_t)T * G * 8u);
}

int launch_transpose(const uint64_t *bitmap, uint64_t n_rows, uint32_t G, uint32_t Wp, uint64_t *gm,
                     uint64_t gm_stride, const uint32_t *perm, cudaStream_t stream) {
    const uint32_t W = (G + 63u) / 64u;
    const char *env = getenv("PGX_TRANSPOSE");
    if (!(env && !strcmp(env, "shfl"))) {  // default: in-register 32x32 transposes (PGX_TRANSPOSE=shfl: the shuffle butterfly)
        const uint32_t row_cols = Wp * 2u;
        uint32_t wl = 1;
        while (wl < 5u && (1u << wl) < row_cols) ++wl;  // widest tile row that the bitmap row fills: 2, 4, 8, 16 or 32 columns
        const uint64_t rows_per_tile = (512u >> wl) * 32u;
        const dim3 grid((unsigned)((gm_stride * 64u + rows_per_tile - 1u) / rows_per_tile), (row_cols + (1u << wl) - 1u) >> wl);
        uint32_t *gm32 = reinterpret_cast<uint32_t *>(gm);
        auto go = [&](auto kern) -> cudaError_t {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            if (e != cudaSuccess) return e;
            kern<<<grid, 512, 65536, stream>>>(bitmap, n_rows, G, Wp, gm32, gm_stride * 2u, perm);
            return cudaSuccess;
        };
        switch (wl) {
            case 1: PGX_CUDA(go(k_transpose_reg<1>)); break;
            case 2: PGX_CUDA(go(k_transpose_reg<2>)); break;
            case 3: PGX_CUDA(go(k_transpose_reg<3>)); break;
            case 4: PGX_CUDA(go(k_transpose_reg<4>)); break;
            default: PGX_CUDA(go(k_transpose_reg<5>)); break;
        }
        PGX_CUDA(cudaGetLastError());
        return PGX_OK;
    }
    const unsigned gx = (unsigned)(gm_stride * 64u / kTrItems);
    if (Wp >= 16u)  // 128-byte (or wider) rows: one CTA reads whole lines
        k_transpose<16><<<dim3(gx, (W + 15u) / 16u), 512, 0, stream>>>(bitmap, n_rows, G, W, Wp,
                                                                      reinterpret_cast<uint32_t *>(gm), gm_stride * 2u, perm);
    else
        k_transpose<8><<<dim3(gx, (W + 7u) / 8u), 256, 0, stream>>>(bitmap, n_rows, G, W, Wp,
                                                                    reinterpret_cast<uint32_t *>(gm), gm_stride * 2u, perm);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

template <int P, bool GENERAL>
int launch_gm_growth_p(const GmGrowthParams &p, cudaStream_t stream) {
    if (p.T <= 1u) return launch_gm_growth_t<P, GENERAL, 1>(p, stream);
    if (p.T <= 2u) return launch_gm_growth_t<P, GENERAL, 2>(p, stream);
    if (p.T <= 4u) return launch_gm_growth_t<P, GENERAL, 4>(p, stream);
    return launch_gm_growth_t<P, GENERAL, kMaxThresholds>(p, stream);
}

int launch_gm_growth(const GmGrowthParams &p, int /*sm_count*/, cudaStream_t stream) {
    if (p.n_orders == 0 || p.n_orders > 65535u) return fail(PGX_ERR_INVALID, "n_orders must be in 1..65535 per launch");
    if (!p.general_mask) return launch_gm_growth_p<1, false>(p, stream);
    if (p.G <= 255u) return launch_gm_growth_p<8, true>(p, stream);
    if (p.G <= 1023u) return launch_gm_growth_p<10, true>(p, stream);
    if (p.G <= 4095u) return launch_gm_growth_p<12, true>(p, stream);
    if (p.G <= 65535u) return launch_gm_growth_p<16, true>(p, stream);
    return launch_gm_growth_p<kRankPlanes, true>(p, stream);
}

int launch_gm_similarity(const GmSimParams &p, int sm_count, cudaStream_t stream) {
    const uint32_t rows = p.row_end - p.row_begin;
    if (rows == 0) return PGX_OK;
    if (p.col_begin >= p.G) return PGX_OK;
    const uint32_t ty = (rows + kSimTile - 1) / kSimTile, tx = (p.G - p.col_begin + kSimTile - 1) / kSimTile;
    // split the item-word range so that the grid fills the GPU a few times over
    uint64_t splits = ((uint64_t)sm_count * 8u + (uint64_t)tx * ty - 1) / ((uint64_t)tx * ty);
    const uint64_t max_splits = (p.n_words + kSimKW - 1) / kSimKW;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535u) splits = 65535u;
    uint64_t wps = (p.n_words + splits - 1) / splits;
    wps = (wps + kSimKW - 1) / kSimKW * kSimKW;
    splits = (p.n_words + wps - 1) / wps;
    dim3 grid(tx, ty, (unsigned)splits);
    GmSimParams q = p;
    q.triangular = (p.row_begin == 0 && p.row_end == p.G && p.col_begin == 0 && tx == ty && tx > 1u) ? 1u : 0u;
    if (q.triangular) grid = dim3(tx * (tx + 1u) / 2u, 1u, (unsigned)splits);
    if (p.planes) {
        if (p.n_planes > 32u) return fail(PGX_ERR_INVALID, "n_planes > 32");
        k_gm_similarity<true, false><<<grid, 256, 0, stream>>>(q, (uint32_t)wps);
    } else if (p.csa) {
        k_gm_similarity<false, true><<<grid, 256, 0, stream>>>(q, (uint32_t)wps);
    } else {
        k_gm_similarity<false, false><<<grid, 256, 0, stream>>>(q, (uint32_t)wps);
    }
    PGX_CUDA(cudaGetLastError());
    if (q.triangular) {
        k_sim_mirror<<<dim3((p.G + 15u) / 16u, (p.G + 15u) / 16u), 256, 0, stream>>>(p.inter, p.G);
        PGX_CUDA(cudaGetLastError());
    }
    return PGX_OK;
}

int launch_sim_mirror(uint64_t *inter, uint32_t G, cudaStream_t stream) {
    k_sim_mirror<<<dim3((G + 15u) / 16u, (G + 15u) / 16u), 256, 0, stream>>>(inter, G);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

int launch_gm_rowsum(const uint64_t *gm, uint64_t gm_stride, uint64_t n_words, const uint64_t *planes,
                     uint32_t n_planes, const uint64_t *uniform_w, uint32_t G, uint64_t *len, cudaStream_t stream) {
    k_gm_rowsum<<<G, 256, 0, stream>>>(gm, gm_stride, n_words, planes, n_planes, uniform_w, len);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

int launch_prefix_curves(uint64_t *d, uint64_t n_curves, uint32_t G, cudaStream_t stream) {
    if (n_curves == 0 || G == 0) return PGX_OK;
    k_prefix_curves<<<(unsigned)((n_curves * 32u + 255u) / 256u), 256, 0, stream>>>(d, n_curves, G);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

int launch_weight_planes(const uint32_t *weight, uint64_t n_rows, uint64_t *planes, uint64_t gm_stride,
                         uint32_t n_planes, uint64_t *uniform_w, uint32_t *plane_mask, int skip0, cudaStream_t stream) {
    const uint64_t threads = gm_stride * 32u;  // one warp per (padded) 64-item word
    k_weight_planes<<<(unsigned)((threads + 255u) / 256u), 256, 0, stream>>>(weight, n_rows, planes, gm_stride, n_planes,
                                                                           gm_stride, uniform_w, plane_mask, skip0);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

__global__ void __launch_bounds__(256) k_gather_keys(const uint32_t *__restrict__ weight, const uint32_t *__restrict__ order,
                                                     uint64_t n_rows, uint32_t *__restrict__ keys) {
    const uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    if (i < n_rows) {
        const uint32_t item = order[i];
        keys[i] = item ? __ldg(weight + item) : 0u;  // the dummy item sorts with the zero weights
    }
}

// perm[i] = item at sorted position i (weights descending), sorted_w[i] = its weight.  secondary != nullptr: items of
// equal weight are ordered by that key (descending) -- two passes of the stable radix sort, the secondary key first.
int sort_items_by_weight(const uint32_t *weight, uint64_t n_rows, uint32_t *perm, uint32_t *sorted_w, cudaStream_t stream,
                         const uint32_t *secondary) {
    uint32_t *keys = nullptr, *vals = nullptr, *vals2 = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    const int64_t n = (int64_t)n_rows;  // CUB takes a 64-bit count: n_rows may exceed 2^31
    auto cleanup = [&]() {
        cudaFree(keys);
        cudaFree(vals);
        cudaFree(vals2);
        cudaFree(tmp);
    };
    cudaError_t e;
    if ((e = cudaMalloc(reinterpret_cast<void **>(&keys), n_rows * 4u)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&vals), n_rows * 4u)) != cudaSuccess ||
        (secondary && (e = cudaMalloc(reinterpret_cast<void **>(&vals2), n_rows * 4u)) != cudaSuccess)) {
        cleanup();
        return fail(PGX_ERR_NOMEM, cudaGetErrorString(e));
    }
    const unsigned blocks = (unsigned)((n_rows + 255u) / 256u);
    e = cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, keys, sorted_w, vals, perm, n, 0, 32, stream);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess && secondary) {
        // pass 1: by the secondary key (sorted_w doubles as the scratch output for the sorted keys)
        k_sort_keys<<<blocks, 256, 0, stream>>>(secondary, n_rows, keys, vals);
        e = cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, keys, sorted_w, vals, vals2, n, 0, 32, stream);
        // pass 2: by weight, stable: the order of pass 1 survives among equal weights
        if (e == cudaSuccess) {
            k_gather_keys<<<blocks, 256, 0, stream>>>(weight, vals2, n_rows, keys);
            e = cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, keys, sorted_w, vals2, perm, n, 0, 32, stream);
        }
    } else if (e == cudaSuccess) {
        k_sort_keys<<<blocks, 256, 0, stream>>>(weight, n_rows, keys, vals);
        e = cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, keys, sorted_w, vals, perm, n, 0, 32, stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cleanup();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(PGX_ERR_CUDA, std::string("sort_items_by_weight: ") + cudaGetErrorString(e));
    }
    return PGX_OK;
}

int launch_build(uint64_t *bitmap, uint32_t Wp, uint64_t n_rows, uint32_t G, const void *d_items, int id_bytes, uint64_t step0,
                 uint64_t n_steps, const uint64_t *d_prefsum, uint64_t n_paths, const int64_t *d_path_group,
                 const uint8_t *d_exclude, unsigned int *d_err, cudaStream_t stream) {
    if (n_steps == 0) return PGX_OK;
    uint64_t blocks = (n_steps + 255u) / 256u;
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    if (id_bytes == 4)
        k_build<uint32_t><<<(unsigned)blocks, 256, 0, stream>>>(bitmap, Wp, n_rows, G, static_cast<const uint32_t *>(d_items), step0,
                                                                n_steps, d_prefsum, n_paths, d_path_group, d_exclude, d_err);
    else
        k_build<uint64_t><<<(unsigned)blocks, 256, 0, stream>>>(bitmap, Wp, n_rows, G, static_cast<const uint64_t *>(d_items), step0,
                                                                n_steps, d_prefsum, n_paths, d_path_group, d_exclude, d_err);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

int launch_scatter(uint64_t *bitmap, uint32_t Wp, uint64_t n_rows, const uint64_t *d_items, uint64_t n_steps,
                   uint32_t group_id, const uint8_t *d_exclude, unsigned int *d_err, cudaStream_t stream) {
    if (n_steps == 0) return PGX_OK;
    uint64_t blocks = (n_steps + 255u) / 256u;
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    k_scatter<<<(unsigned)blocks, 256, 0, stream>>>(bitmap, Wp, n_rows, d_items, n_steps, group_id, d_exclude, d_err);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

}  // namespace pgx
