// pgx_quorum.cu -- ordered growth under arbitrary group orders on the group-major copy of the bitmap:
//   k_gm_quorum  general quorum thresholds (q > 0), optionally with one q = 0 threshold riding along
//   k_gm_union   q = 0 thresholds only (the permutation-sampled union / coverage >= c growth)
// Same arithmetic as AbacusByGroup::calc_growth (src/graph_broker/abacus.rs:989-1032) applied to the abacus the
// reference would rebuild under `--order` (abacus.rs:324-326); the thresholds are the host's f64 `ceil((j + 1) * q)`
// per position (abacus.rs:1010).
//
// k_gm_quorum: one thread owns 64 items and walks the groups in order, keeping their ranks in P bit-planes
// (pgx_rank.cuh).  Integer-ALU bound by construction (~4P + 2PT logic ops per 64 items and position for T thresholds),
// so the inner loop is kept to the plane operations themselves:
//   * the per-plane cutoff masks come from a shared-memory table built once per CTA ([position][threshold][plane],
//     16-byte rows read with LDS.128 on the otherwise idle LSU pipe) instead of being rebuilt from K in every lane;
//   * T general thresholds and the weighted mode are template parameters: no per-threshold branches.  One q = 0
//     threshold may ride along (p.n_fast: "item counts from its first group on", a popcount of the seen mask, a few
//     ops per position); further q = 0 thresholds run on k_gm_union;
//   * counting nodes: a thread tracks popc(verdict & eligible) and adds the change; one REDUX per warp and position,
//     whose result is parked in lane (position % 32) and added to the CTA's shared first differences once per 32
//     positions -- one conflict-free native 32-bit shared reduction instead of an elected-lane atomic per position
//     (and no 64-bit shared atomics, which are CAS loops on sm_100);
//   * summing bp: runs on the weight-sorted group-major copy (the one similarity uses): most 64-item columns then carry
//     a single weight w and their contribution is (change of the popcount) x w; only mixed columns walk the flipped
//     bits.  The per-lane net is reduced as three 24/24/16-bit pieces of its two's complement (3 REDUX);
//   * the byte offsets of the rows (order[j] x row pitch) are pre-multiplied in shared memory and padded by repeats of
//     the last row: a row load is LDS.64 + one 64-bit add + LDG, without bounds predicates.
#include <type_traits>

#include "pgx_common.cuh"
#include "pgx_internal.h"
#include "pgx_rank.cuh"

// rows in flight per thread for counting with T <= 2 (tools/gpu_run38.sh compiles the variant on the GPU box): 4 -> 18.39 ms,
// 8 -> 19.78 ms at c3 (80 registers, 32 bytes of spills)
#ifndef PGX_QUORUM_PREFETCH
#define PGX_QUORUM_PREFETCH 4
#endif

namespace pgx {

namespace {

constexpr int kQThreads = 256;

__device__ __forceinline__ long long weight_of_bits(uint32_t lo, uint32_t hi, const uint32_t *__restrict__ wrow) {
    long long s = 0;
    if (!wrow) return (long long)(__popc(lo) + __popc(hi));
    while (lo) {
        const uint32_t b = (uint32_t)__ffs((int)lo) - 1u;
        lo &= lo - 1u;
        s += (long long)__ldg(wrow + b);
    }
    while (hi) {
        const uint32_t b = (uint32_t)__ffs((int)hi) - 1u;
        hi &= hi - 1u;
        s += (long long)__ldg(wrow + 32u + b);
    }
    return s;
}

// native 32-bit shared-memory reduction by the calling lane (wrapping add: signed values as two's complement); inline PTX
// keeps ptxas from wrapping it into its vote / elect / popc aggregation sequence, the caller already is one elected lane
__device__ __forceinline__ void red_shared_add_u32(uint32_t *addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_u32(addr)), "r"(v) : "memory");
}

// exact warp sum of signed 64-bit values with |v| < 2^47: v = hi * 2^48 + mid * 2^24 + lo (hi signed)
__device__ __forceinline__ long long warp_sum_i64(long long v) {
    const unsigned long long u = (unsigned long long)v;
    const uint32_t lo = (uint32_t)(u & 0xFFFFFFull), mid = (uint32_t)((u >> 24) & 0xFFFFFFull);
    const int hi = (int)(v >> 48);
    const unsigned long long slo = __reduce_add_sync(0xFFFFFFFFu, lo);
    const unsigned long long smid = __reduce_add_sync(0xFFFFFFFFu, mid);
    const long long shi = (long long)__reduce_add_sync(0xFFFFFFFFu, hi);
    return (long long)(slo + (smid << 24) + ((unsigned long long)shi << 48));
}

// T general thresholds (cov / slot index 0 .. T-1) and p.n_fast <= NF thresholds with q = 0 (index T .. T+n_fast-1).
template <int P, int T, int NF, bool WEIGHTED>
__global__ void __launch_bounds__(kQThreads, (P >= 16) ? 1 : ((T <= 2 && !WEIGHTED && P <= 10) ? 3 : 2)) k_gm_quorum(const __grid_constant__ GmGrowthParams p) {
    static_assert(T >= 1 && T <= (int)kGmQuorumMaxT && NF == 1, "1..4 general thresholds and at most one q = 0 rider");
    constexpr int PP = RankMaskWords<P>::value;
    // rows in flight per thread (divides 32, at most 8: the padding of s_off); a single order streams from DRAM, many orders from L2
    constexpr int kPrefetch = (!WEIGHTED && T <= 2) ? PGX_QUORUM_PREFETCH : 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem_raw);  // [G][T][PP]
    // byte offset of the row added at position j (+ 8 entries repeating the last row: the prefetch needs no bounds check)
    unsigned long long *s_off = reinterpret_cast<unsigned long long *>(s_mask + (size_t)p.G * T * PP);  // [G + 8]
    uint32_t *s_dlo = reinterpret_cast<uint32_t *>(s_off + p.G + 8u);  // [T + NF][G] first differences (low or only word)
    uint32_t *s_dhi = s_dlo + (size_t)(T + NF) * p.G;                  // [T + NF][G] high words (WEIGHTED only)
    const uint32_t n_fast = p.n_fast < (uint32_t)NF ? p.n_fast : (uint32_t)NF;  // q = 0 thresholds (warp-uniform)
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    // 1-D grid, order fastest (see k_gm_growth): co-resident CTAs share their column blocks through L2
    const uint32_t n_col_blocks = (uint32_t)((p.n_words + kQThreads - 1) / kQThreads);
    const uint32_t order_id = p.col_fastest ? blockIdx.x / n_col_blocks : blockIdx.x % p.n_orders;
    const uint64_t col_block = p.col_fastest ? blockIdx.x % n_col_blocks : blockIdx.x / p.n_orders;
    const uint32_t *order = p.order + (size_t)order_id * p.G;
    for (uint32_t i = tid; i < p.G + 8u; i += kQThreads)
        s_off[i] = (unsigned long long)order[i < p.G ? i : p.G - 1u] * p.gm_stride * 8ull;
    for (uint32_t i = tid; i < p.G * T; i += kQThreads) {
        const uint32_t j = i / T, t = i - j * T;
        // ranks are <= j + 1 at position j: any cutoff above that never passes, j + 2 stands for all of them -- which keeps
        // every cutoff below 2^PE for the plane-skipping blocks further down (jend + 1 < 2^PE)
        const uint32_t K = __ldg(p.thr + (size_t)t * p.G + j);
        rank_mask_row<P>(K < j + 2u ? K : j + 2u, p.G, s_mask + (size_t)i * PP);
    }
    for (uint32_t i = tid; i < p.G * (T + NF) * (WEIGHTED ? 2u : 1u); i += kQThreads) s_dlo[i] = 0u;
    __syncthreads();

    const uint64_t wi = col_block * kQThreads + tid;
    const bool active = wi < p.n_words;
    const uint64_t wsafe = active ? wi : 0;
    // WEIGHTED with weights: p.gm is the weight-sorted copy -- bit b of column wi is item p.perm[wi * 64 + b], p.weight
    // holds the weights in that order and p.uniform_w[wi] = (1 << 32) | w when all items of the column weigh w
    const uint32_t *wrow = (WEIGHTED && p.weight) ? p.weight + wsafe * 64u : nullptr;
    bool uniform = WEIGHTED && !p.weight;  // unit weights
    uint32_t uw = 1u;
    if (WEIGHTED && p.weight && p.uniform_w) {
        const unsigned long long u = active ? __ldg(reinterpret_cast<const unsigned long long *>(p.uniform_w) + wi) : (1ull << 32);
        uniform = (u >> 32) != 0ull;
        uw = (uint32_t)u;
    }

    // eligibility: an item counts for threshold t only if its total coverage >= cov[t] (abacus.rs:1003)
    uint32_t elo[T + NF], ehi[T + NF];
#pragma unroll
    for (int t = 0; t < T + NF; ++t) elo[t] = ehi[t] = ~0u;
    bool need_cov = false;
#pragma unroll
    for (int t = 0; t < T + NF; ++t) need_cov |= (uint32_t)t < (uint32_t)T + n_fast && p.cov[t] > 1u;
    if (need_cov) {
#pragma unroll
        for (int t = 0; t < T + NF; ++t) elo[t] = ehi[t] = 0u;
        if (active) {
            for (uint32_t b = 0; b < 64u; ++b) {
                const uint64_t pos = wi * 64u + b;
                uint64_t item = pos;
                if (p.perm && pos < p.n_rows) item = __ldg(p.perm + pos);
                const uint32_t c = (pos < p.n_rows && item != 0) ? __ldg(p.countable + item) : 0u;
#pragma unroll
                for (int t = 0; t < T + NF; ++t)
                    if (c >= p.cov[t]) {
                        if (b < 32u) elo[t] |= 1u << b; else ehi[t] |= 1u << (b - 32u);
                    }
            }
        }
    }

    if (!active) {
#pragma unroll
        for (int t = 0; t < T + NF; ++t) elo[t] = ehi[t] = 0u;
    }
    // warp-uniform: does any of the warp's 2048 items reach threshold t's coverage cutoff?  On the coverage-sorted copy
    // (p.perm, counting) most warps are all-or-nothing, and a warp without eligible items skips the threshold's rank
    // comparison -- without any general threshold left, the rank counters themselves.
    bool act[T], any_act = false;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        act[t] = __any_sync(0xFFFFFFFFu, (elo[t] | ehi[t]) != 0u) != 0;
        any_act |= act[t];
    }
    RankColumn<P> R;
    R.clear();
    uint32_t vlo[T], vhi[T];
    int cnt[T + NF];
#pragma unroll
    for (int t = 0; t < T; ++t) vlo[t] = vhi[t] = 0u;
#pragma unroll
    for (int t = 0; t < T + NF; ++t) cnt[t] = 0;
    uint32_t slo = 0u, shi = 0u;  // items seen so far (q = 0 threshold)

    // threads past the last column read column 0 like everybody else; their eligibility masks are cleared below, so
    // they never count
    const unsigned char *col = reinterpret_cast<const unsigned char *>(p.gm + wsafe);
    auto load_row = [&](uint32_t j) -> uint64_t { return __ldg(reinterpret_cast<const unsigned long long *>(col + s_off[j])); };
    uint64_t ring[kPrefetch];  // slot u holds row j0 + u; refilled with row j0 + kPrefetch + u as soon as it is consumed
#pragma unroll
    for (int u = 0; u < kPrefetch; ++u) ring[u] = load_row((uint32_t)u);

    // The warp sum of position j is parked in lane j % 32; every 32 positions each lane adds its value to the CTA's
    // shared first differences -- one conflict-free reduction per 32 positions instead of an elected-lane one per position.
    typename std::conditional<WEIGHTED, long long, int>::type park[T + NF];
#pragma unroll
    for (int t = 0; t < T + NF; ++t) park[t] = 0;
    for (uint32_t jb = 0; jb < p.G; jb += 32u) {
      const uint32_t jend = jb + 32u < p.G ? jb + 32u : p.G;
      // 32 positions with the rank planes that can be non-zero by their end (ranks <= jend, cutoffs <= jend): PE planes
      auto block = [&](auto pe_tag) {
      constexpr int PE = decltype(pe_tag)::value;
      for (uint32_t j0 = jb; j0 < jend; j0 += kPrefetch) {
#pragma unroll
        for (int u = 0; u < kPrefetch; ++u) {
            const uint32_t j = j0 + (uint32_t)u;
            if (j >= jend) break;
            const bool my_slot = lane == j - jb;
            const uint64_t rowbits = ring[u];
            ring[u] = load_row(j + kPrefetch);
            const uint32_t blo = (uint32_t)rowbits, bhi = (uint32_t)(rowbits >> 32);
            if (any_act) R.template add_n<PE>(blo, bhi);
            if (n_fast != 0u) {  // warp-uniform: an item starts counting at its first group and never stops (q = 0)
                const uint32_t olo = slo, ohi = shi;
                slo |= blo;
                shi |= bhi;
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    if ((uint32_t)f >= n_fast) break;
                    int d = 0;
                    if (!WEIGHTED || uniform) {
                        const int c = __popc(slo & elo[T + f]) + __popc(shi & ehi[T + f]);
                        d = c - cnt[T + f];
                        cnt[T + f] = c;
                    }
                    if (!WEIGHTED) {
                        const int net = __reduce_add_sync(0xFFFFFFFFu, d);
                        if (my_slot) park[T + f] = net;
                    } else {
                        long long mine;
                        if (uniform) {  // every item of the column weighs uw
                            mine = (long long)d * (long long)uw;
                        } else {
                            const uint32_t flo = blo & ~olo & elo[T + f], fhi = bhi & ~ohi & ehi[T + f];
                            mine = (flo | fhi) ? weight_of_bits(flo, fhi, wrow) : 0ll;
                        }
                        const long long net = warp_sum_i64(mine);
                        if (my_slot) park[T + f] = net;
                    }
                }
            }
            const uint4 *row = reinterpret_cast<const uint4 *>(s_mask + (size_t)j * (T * PP));
#pragma unroll
            for (int t = 0; t < T; ++t) {
                if (!act[t]) continue;  // (park[t] stays 0: the warp adds nothing to this curve)
                uint32_t m[PP];
#pragma unroll
                for (int q = 0; q < PP / 4; ++q) {
                    const uint4 x = row[t * (PP / 4) + q];
                    m[4 * q + 0] = x.x;
                    m[4 * q + 1] = x.y;
                    m[4 * q + 2] = x.z;
                    m[4 * q + 3] = x.w;
                }
                uint32_t glo, ghi;
                R.template ge_n<PE>(m, glo, ghi);
                const uint32_t nlo = verdict_update(blo, glo, vlo[t]), nhi = verdict_update(bhi, ghi, vhi[t]);
                if (!WEIGHTED) {
                    const int c = __popc(nlo & elo[t]) + __popc(nhi & ehi[t]);
                    const int net = __reduce_add_sync(0xFFFFFFFFu, c - cnt[t]);
                    cnt[t] = c;
                    if (my_slot) park[t] = net;
                } else {
                    long long mine = 0;
                    if (uniform) {  // every item of the column weighs uw: (change of the count) x uw
                        const int c = __popc(nlo & elo[t]) + __popc(nhi & ehi[t]);
                        mine = (long long)(c - cnt[t]) * (long long)uw;
                        cnt[t] = c;
                    } else {
                        const uint32_t ulo = nlo & ~vlo[t] & elo[t], uhi = nhi & ~vhi[t] & ehi[t];
                        const uint32_t dlo = vlo[t] & ~nlo & elo[t], dhi = vhi[t] & ~nhi & ehi[t];
                        if (ulo | uhi) mine += weight_of_bits(ulo, uhi, wrow);
                        if (dlo | dhi) mine -= weight_of_bits(dlo, dhi, wrow);
                    }
                    const long long net = warp_sum_i64(mine);
                    if (my_slot) park[t] = net;
                }
                vlo[t] = nlo;
                vhi[t] = nhi;
            }
        }
      }
      };
      // Three instantiations of the block triple the loop's code.  Counting with T <= 2 gains from it (19.4 -> 18.6 ms at
      // c3); the larger weighted body then no longer fits the instruction cache -- ncu: "no_instruction" became the top
      // stall and the kernel went from 8.0 to 10.1 ms per 20 orders (profiles/r2b_quorum_ncu_summary.txt) -- so only the
      // small bodies skip planes.
      constexpr bool kSkipPlanes = !WEIGHTED && T <= 2 && P >= 4;
      if (kSkipPlanes && jend + 1u < (1u << (P - 2)))
          block(std::integral_constant<int, (kSkipPlanes ? P - 2 : P)>{});
      else if (kSkipPlanes && jend + 1u < (1u << (P - 1)))
          block(std::integral_constant<int, (kSkipPlanes ? P - 1 : P)>{});
      else
          block(std::integral_constant<int, P>{});
      if (jb + lane < jend) {  // lane l holds the warp sums of position jb + l
#pragma unroll
          for (int t = 0; t < T + NF; ++t) {
              if ((uint32_t)t >= (uint32_t)T + n_fast) break;
              if (WEIGHTED) {
                  const unsigned long long v = (unsigned long long)park[t];
                  if (v) smem_add64(s_dlo, s_dhi, t * p.G + jb + lane, (uint32_t)v, (uint32_t)(v >> 32));
              } else {
                  red_shared_add_u32(s_dlo + t * p.G + jb + lane, (uint32_t)park[t]);
              }
          }
      }
    }
    __syncthreads();
    uint64_t *out = p.out + (size_t)order_id * p.out_order_stride;
    for (uint32_t i = tid; i < p.G * ((uint32_t)T + n_fast); i += kQThreads) {
        unsigned long long v;
        if (WEIGHTED) v = (unsigned long long)s_dlo[i] | ((unsigned long long)s_dhi[i] << 32);
        else v = (unsigned long long)(long long)(int)s_dlo[i];  // signed count -> two's-complement u64
        if (v) {
            const uint32_t t = i / p.G, j = i - t * p.G;
            atomicAdd(reinterpret_cast<unsigned long long *>(out + (size_t)p.slot[t] * p.G + j), v);
        }
    }
}

// ---- q = 0 thresholds only: "an item counts from its first group on" (abacus.rs:1007-1010 with ceil((c+1)*0) = 0) ----------
// The permutation-sampled union / coverage >= c growth.  No ranks: a thread keeps the `seen` mask of 128 items (two
// adjacent u64 columns, one 16-byte load per row) and adds the change of popc(seen & eligible) at every position.  With
// several orders in flight the column blocks are served from L2 (one DRAM pass for all orders), so the loop is
// issue-bound: two columns per thread halve the per-item share of the row lookup, the REDUX and the shared atomic.
// NF thresholds (p.n_fast <= NF of them in use) differ only in their coverage cutoff; COV = any cutoff > 1.
template <int NF, bool WEIGHTED, bool COV>
__global__ void __launch_bounds__(kQThreads, (NF >= 4 || (WEIGHTED && NF >= 2)) ? 2 : 3) k_gm_union(const __grid_constant__ GmGrowthParams p) {
    constexpr int kPrefetch = 8;  // divides 32 (the parking scheme below)
    constexpr int NE = COV ? NF : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // byte offset of the row added at position j; kPrefetch extra entries repeat the last row so that the prefetch
    // needs no bounds predicate
    unsigned long long *s_off = reinterpret_cast<unsigned long long *>(smem_raw);  // [G + kPrefetch]
    uint32_t *s_dlo = reinterpret_cast<uint32_t *>(s_off + p.G + kPrefetch);       // [NF][G]
    uint32_t *s_dhi = s_dlo + (size_t)NF * p.G;                                    // [NF][G] (WEIGHTED only)
    const uint32_t n_fast = p.n_fast < (uint32_t)NF ? p.n_fast : (uint32_t)NF;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint64_t n_pairs = (p.n_words + 1u) / 2u;  // pairs of columns
    const uint32_t n_col_blocks = (uint32_t)((n_pairs + kQThreads - 1) / kQThreads);
    const uint32_t order_id = p.col_fastest ? blockIdx.x / n_col_blocks : blockIdx.x % p.n_orders;
    const uint64_t col_block = p.col_fastest ? blockIdx.x % n_col_blocks : blockIdx.x / p.n_orders;
    const uint32_t *order = p.order + (size_t)order_id * p.G;
    for (uint32_t i = tid; i < p.G + kPrefetch; i += kQThreads)
        s_off[i] = (unsigned long long)order[i < p.G ? i : p.G - 1u] * p.gm_stride * 8ull;
    for (uint32_t i = tid; i < p.G * NF * (WEIGHTED ? 2u : 1u); i += kQThreads) s_dlo[i] = 0u;
    __syncthreads();

    const uint64_t pi = col_block * kQThreads + tid;
    const bool active = pi < n_pairs;
    const uint64_t w0 = active ? pi * 2u : 0;  // first column of the pair (the rows are padded with zero words to 128 bytes)
    // weight-sorted copy (see k_gm_quorum): per column, (1 << 32) | w when all its items weigh w
    const uint32_t *wrow = (WEIGHTED && p.weight) ? p.weight + w0 * 64u : nullptr;
    bool uni[2] = {WEIGHTED && !p.weight, WEIGHTED && !p.weight};
    uint32_t uw[2] = {1u, 1u};
    if (WEIGHTED && p.weight && p.uniform_w) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const bool in = active && w0 + h < p.n_words;
            const unsigned long long u = in ? __ldg(reinterpret_cast<const unsigned long long *>(p.uniform_w) + w0 + h) : (1ull << 32);
            uni[h] = (u >> 32) != 0ull;
            uw[h] = (uint32_t)u;
        }
    }
    uint32_t elig[NE][4];  // [threshold][32-item quarter]
    if (COV) {
#pragma unroll
        for (int f = 0; f < NE; ++f)
#pragma unroll
            for (int q = 0; q < 4; ++q) elig[f][q] = 0u;
        if (active) {
            for (uint32_t b = 0; b < 128u; ++b) {
                const uint64_t pos = w0 * 64u + b;
                uint64_t item = pos;
                if (p.perm && pos < p.n_rows) item = __ldg(p.perm + pos);
                const uint32_t c = (pos < p.n_rows && item != 0) ? __ldg(p.countable + item) : 0u;
#pragma unroll
                for (int f = 0; f < NE; ++f)
                    if (c >= p.cov[f]) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if ((b >> 5) == (uint32_t)q) elig[f][q] |= 1u << (b & 31u);
                    }
            }
        }
    }
    uint32_t seen[4] = {0u, 0u, 0u, 0u};
    int cnt[NF][2];
#pragma unroll
    for (int f = 0; f < NF; ++f) cnt[f][0] = cnt[f][1] = 0;

    // threads past the last pair read pair 0 (w0 = 0) like everybody else and contribute nothing (see `live` below)
    const unsigned char *col = reinterpret_cast<const unsigned char *>(p.gm + w0);
    auto load_row = [&](uint32_t j) -> uint4 { return __ldg(reinterpret_cast<const uint4 *>(col + s_off[j])); };
    const int live = active ? 1 : 0;
    uint4 ring[kPrefetch];  // slot u holds row j0 + u; it is refilled with row j0 + kPrefetch + u as soon as it is consumed
#pragma unroll
    for (int u = 0; u < kPrefetch; ++u) ring[u] = load_row((uint32_t)u);
    // warp sums are parked in lane j % 32 and added to shared memory once per 32 positions (see k_gm_quorum)
    typename std::conditional<WEIGHTED, long long, int>::type park[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) park[f] = 0;
    for (uint32_t jb = 0; jb < p.G; jb += 32u) {
      const uint32_t jend = jb + 32u < p.G ? jb + 32u : p.G;
      for (uint32_t j0 = jb; j0 < jend; j0 += kPrefetch) {
#pragma unroll
        for (int u = 0; u < kPrefetch; ++u) {
            const uint32_t j = j0 + (uint32_t)u;
            if (j >= jend) break;
            const bool my_slot = lane == j - jb;
            const uint32_t b[4] = {ring[u].x, ring[u].y, ring[u].z, ring[u].w};
            uint32_t old[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                old[q] = seen[q];
                seen[q] |= b[q];
            }
            if (!WEIGHTED) ring[u] = load_row(j + kPrefetch);  // the slot is consumed: refill it (no register copies)
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                if (f > 0 && (uint32_t)f >= n_fast) break;  // n_fast >= 1
                const int fe = COV ? f : 0;
                int d[2] = {0, 0};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (!WEIGHTED || uni[h]) {
                        const int c = COV ? __popc(seen[2 * h] & elig[fe][2 * h]) + __popc(seen[2 * h + 1] & elig[fe][2 * h + 1])
                                          : __popc(seen[2 * h]) + __popc(seen[2 * h + 1]);
                        d[h] = c - cnt[f][h];
                        cnt[f][h] = c;
                    }
                }
                if (!WEIGHTED) {
                    const int net = __reduce_add_sync(0xFFFFFFFFu, (d[0] + d[1]) * live);
                    if (my_slot) park[f] = net;
                } else {
                    long long mine = 0;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (!live) break;
                        if (uni[h]) {
                            mine += (long long)d[h] * (long long)uw[h];
                        } else {
                            uint32_t flo = b[2 * h] & ~old[2 * h], fhi = b[2 * h + 1] & ~old[2 * h + 1];
                            if (COV) {
                                flo &= elig[fe][2 * h];
                                fhi &= elig[fe][2 * h + 1];
                            }
                            if (flo | fhi) mine += weight_of_bits(flo, fhi, wrow + 64 * h);
                        }
                    }
                    const long long net = warp_sum_i64(mine);
                    if (my_slot) park[f] = net;
                }
            }
            if (WEIGHTED) ring[u] = load_row(j + kPrefetch);  // mixed-weight columns still needed the row above
        }
      }
      if (jb + lane < jend) {
#pragma unroll
          for (int f = 0; f < NF; ++f) {
              if (f > 0 && (uint32_t)f >= n_fast) break;
              if (WEIGHTED) {
                  const unsigned long long v = (unsigned long long)park[f];
                  if (v) smem_add64(s_dlo, s_dhi, f * p.G + jb + lane, (uint32_t)v, (uint32_t)(v >> 32));
              } else {
                  red_shared_add_u32(s_dlo + f * p.G + jb + lane, (uint32_t)park[f]);
              }
          }
      }
    }
    __syncthreads();
    uint64_t *out = p.out + (size_t)order_id * p.out_order_stride;
    for (uint32_t i = tid; i < p.G * n_fast; i += kQThreads) {
        unsigned long long v;
        if (WEIGHTED) v = (unsigned long long)s_dlo[i] | ((unsigned long long)s_dhi[i] << 32);
        else v = (unsigned long long)s_dlo[i];  // counts only grow for q = 0
        if (v) {
            const uint32_t f = i / p.G, j = i - f * p.G;
            atomicAdd(reinterpret_cast<unsigned long long *>(out + (size_t)p.slot[f] * p.G + j), v);
        }
    }
}

template <int NF, bool WEIGHTED, bool COV>
int launch_union(const GmGrowthParams &p, cudaStream_t stream) {
    const size_t smem = ((size_t)p.G + 8u) * 8u + (size_t)NF * p.G * (WEIGHTED ? 2u : 1u) * 4u;  // 8 = kPrefetch of k_gm_union
    auto kern = k_gm_union<NF, WEIGHTED, COV>;
    PGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t n_pairs = (p.n_words + 1u) / 2u;
    const uint64_t blocks = (n_pairs + kQThreads - 1) / kQThreads * p.n_orders;
    if (blocks > 0x7FFFFFFFull) return fail(PGX_ERR_UNSUPPORTED, "k_gm_union: too many column blocks x orders in one launch");
    kern<<<(unsigned)blocks, kQThreads, smem, stream>>>(p);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

template <bool WEIGHTED, bool COV>
int launch_union_n(const GmGrowthParams &p, cudaStream_t stream) {
    if (p.n_fast <= 1u) return launch_union<1, WEIGHTED, COV>(p, stream);
    if (p.n_fast <= 2u) return launch_union<2, WEIGHTED, COV>(p, stream);
    return launch_union<4, WEIGHTED, COV>(p, stream);
}

template <int P, int T, int NF, bool WEIGHTED>
int launch_q(const GmGrowthParams &p, cudaStream_t stream) {
    const size_t smem = gm_quorum_smem_bytes(p.G, T, NF, WEIGHTED);
    auto kern = k_gm_quorum<P, T, NF, WEIGHTED>;
    PGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t blocks = (p.n_words + kQThreads - 1) / kQThreads * p.n_orders;
    if (blocks > 0x7FFFFFFFull) return fail(PGX_ERR_UNSUPPORTED, "k_gm_quorum: too many column blocks x orders in one launch");
    kern<<<(unsigned)blocks, kQThreads, smem, stream>>>(p);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

template <int P, bool WEIGHTED>
int launch_q_t(const GmGrowthParams &p, cudaStream_t stream) {
    switch (p.T) {
        case 1: return launch_q<P, 1, 1, WEIGHTED>(p, stream);
        case 2: return launch_q<P, 2, 1, WEIGHTED>(p, stream);
        case 3: return launch_q<P, 3, 1, WEIGHTED>(p, stream);
        case 4: return launch_q<P, 4, 1, WEIGHTED>(p, stream);
        default: return fail(PGX_ERR_INVALID, "k_gm_quorum takes 1..4 general thresholds per launch");
    }
}

template <int P>
int launch_q_p(const GmGrowthParams &p, cudaStream_t stream) {
    return p.weighted ? launch_q_t<P, true>(p, stream) : launch_q_t<P, false>(p, stream);
}

}  // namespace

int gm_quorum_planes(uint32_t G) {
    const int need = rank_planes_needed(G);
    for (int P : {7, 8, 9, 10, 11, 12, 14, 16, 21})
        if (need <= P) return P;
    return 0;
}

// NF: q = 0 slots of the instantiation (1 next to general thresholds; 1, 2 or 4 when T = 0)
size_t gm_quorum_smem_bytes(uint32_t G, uint32_t T, uint32_t NF, bool weighted) {
    const int P = T ? gm_quorum_planes(G) : 0;
    return ((size_t)G * T * (size_t)((P + 3) & ~3) + 2u * ((size_t)G + 8u) + (size_t)(T + NF) * G * (weighted ? 2u : 1u)) * 4u;
}

uint32_t gm_quorum_fast_slots(uint32_t T, uint32_t n_fast) { return T ? 1u : (n_fast <= 1u ? 1u : n_fast <= 2u ? 2u : 4u); }

// Thresholds 0 .. p.T-1 (p.T <= 4) must be general ones (p.thr holds T x G cutoffs by position); cov / slot index
// p.T .. p.T+p.n_fast-1 describe q = 0 thresholds computed in the same pass (at most 1 when p.T > 0, at most 4 when
// p.T = 0).  p.direct_out is not supported.
int launch_gm_quorum(const GmGrowthParams &p, cudaStream_t stream) {
    if (p.n_orders == 0 || p.n_orders > 65535u) return fail(PGX_ERR_INVALID, "n_orders must be in 1..65535 per launch");
    if (p.direct_out || (p.T == 0 && p.n_fast == 0) || (p.T && !p.thr) || p.n_fast > (p.T ? 1u : 4u) || p.T > kGmQuorumMaxT)
        return fail(PGX_ERR_INVALID, "k_gm_quorum: bad parameters");
    if (gm_quorum_smem_bytes(p.G, p.T, gm_quorum_fast_slots(p.T, p.n_fast), p.weighted != 0) > kGmQuorumSmemMax)
        return fail(PGX_ERR_UNSUPPORTED, "k_gm_quorum: tables do not fit in shared memory");
    if (p.T == 0) {  // q = 0 only: k_gm_union
        bool cov = false;
        for (uint32_t f = 0; f < p.n_fast; ++f) cov |= p.cov[f] > 1u;
        if (p.weighted) return cov ? launch_union_n<true, true>(p, stream) : launch_union_n<true, false>(p, stream);
        return cov ? launch_union_n<false, true>(p, stream) : launch_union_n<false, false>(p, stream);
    }
    switch (gm_quorum_planes(p.G)) {
        case 7: return launch_q_p<7>(p, stream);
        case 8: return launch_q_p<8>(p, stream);
        case 9: return launch_q_p<9>(p, stream);
        case 10: return launch_q_p<10>(p, stream);
        case 11: return launch_q_p<11>(p, stream);
        case 12: return launch_q_p<12>(p, stream);
        case 14: return launch_q_p<14>(p, stream);
        case 16: return launch_q_p<16>(p, stream);
        case 21: return launch_q_p<21>(p, stream);
        default: return fail(PGX_ERR_UNSUPPORTED, "k_gm_quorum: too many groups");
    }
}

}  // namespace pgx
