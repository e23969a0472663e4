// pgx_quorum.cu -- k_gm_quorum: ordered growth for general quorum thresholds (q > 0) under an arbitrary group
// order, on the group-major copy of the bitmap.  Same arithmetic as AbacusByGroup::calc_growth
// (src/graph_broker/abacus.rs:989-1032) applied to the abacus the reference would rebuild under `--order`
// (abacus.rs:324-326); the thresholds are the host's f64 `ceil((j + 1) * q)` per position (abacus.rs:1010).
//
// One thread owns 64 items and walks the groups in order, keeping their ranks in P bit-planes (pgx_rank.cuh).
// Integer-ALU bound by construction (~4P + 2PT logic ops per 64 items and position for T thresholds), so the
// inner loop is kept to the plane operations themselves:
//   * the per-plane cutoff masks come from a shared-memory table built once per CTA ([position][threshold][plane],
//     16-byte rows read with LDS.128 on the otherwise idle LSU pipe) instead of being rebuilt from K in every lane;
//   * T general thresholds and the weighted mode are template parameters: no per-threshold branches.  One q = 0
//     threshold may ride along (p.n_fast: "item counts from its first group on", a popcount of the seen mask, a few
//     ops per position); further q = 0 thresholds run in a T = 0 launch of their own;
//   * counting nodes: a thread tracks popc(verdict) and adds the change, one REDUX per warp and position and a
//     native 32-bit shared atomic (64-bit shared atomics are CAS loops on sm_100);
//   * summing bp: runs on the weight-sorted group-major copy (the one similarity uses): most 64-item words then carry
//     a single weight w and their contribution is (change of the popcount) x w; only mixed words walk the flipped
//     bits.  The per-lane net is reduced as three 24/24/16-bit pieces of its two's complement (3 REDUX).
//   * T = 0 instantiations (q = 0 thresholds only, up to NF of them differing in their coverage cutoff) replace the
//     first-generation k_gm_growth<.,false> for permuted union growth: same loop without ranks, deeper prefetch.
#include "pgx_common.cuh"
#include "pgx_internal.h"
#include "pgx_rank.cuh"

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
__global__ void __launch_bounds__(kQThreads, (P >= 16) ? 1 : (T == 0 ? 3 : 2)) k_gm_quorum(const __grid_constant__ GmGrowthParams p) {
    constexpr int PP = RankMaskWords<P>::value;
    constexpr int TT = T > 0 ? T : 1;  // array extents (no zero-length arrays)
    constexpr int kPrefetch = T == 0 ? 8 : 2;  // rows in flight per thread: q = 0 only is memory-bound, ranks are ALU-bound
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem_raw);  // [G][T][PP]
    uint32_t *s_order = s_mask + (size_t)p.G * T * PP;          // [G]
    uint32_t *s_dlo = s_order + p.G;                            // [T + NF][G] first differences (low or only word)
    uint32_t *s_dhi = s_dlo + (size_t)(T + NF) * p.G;           // [T + NF][G] high words (WEIGHTED only)
    const uint32_t n_fast = p.n_fast < (uint32_t)NF ? p.n_fast : (uint32_t)NF;  // q = 0 thresholds (warp-uniform)
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    // 1-D grid, order fastest (see k_gm_growth): co-resident CTAs share their column blocks through L2
    const uint32_t n_col_blocks = (uint32_t)((p.n_words + kQThreads - 1) / kQThreads);
    const uint32_t order_id = p.col_fastest ? blockIdx.x / n_col_blocks : blockIdx.x % p.n_orders;
    const uint64_t col_block = p.col_fastest ? blockIdx.x % n_col_blocks : blockIdx.x / p.n_orders;
    const uint32_t *order = p.order + (size_t)order_id * p.G;
    for (uint32_t i = tid; i < p.G; i += kQThreads) s_order[i] = order[i];
    for (uint32_t i = tid; i < p.G * T; i += kQThreads) {
        const uint32_t j = i / T, t = i - j * T;
        rank_mask_row<P>(__ldg(p.thr + (size_t)t * p.G + j), p.G, s_mask + (size_t)i * PP);
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

    RankColumn<P> R;
    R.clear();
    uint32_t vlo[TT], vhi[TT];
    int cnt[T + NF];
#pragma unroll
    for (int t = 0; t < TT; ++t) vlo[t] = vhi[t] = 0u;
#pragma unroll
    for (int t = 0; t < T + NF; ++t) cnt[t] = 0;
    uint32_t slo = 0u, shi = 0u;  // items seen so far (q = 0 threshold)

    const uint64_t *col = p.gm + wsafe;
    auto load_row = [&](uint32_t j) -> uint64_t {
        return (active && j < p.G) ? __ldg(col + (uint64_t)s_order[j] * p.gm_stride) : 0ull;
    };
    uint64_t nxt[kPrefetch];
#pragma unroll
    for (int u = 0; u < kPrefetch; ++u) nxt[u] = load_row((uint32_t)u);

    for (uint32_t j0 = 0; j0 < p.G; j0 += kPrefetch) {
        uint64_t cur[kPrefetch];
#pragma unroll
        for (int u = 0; u < kPrefetch; ++u) cur[u] = nxt[u];
#pragma unroll
        for (int u = 0; u < kPrefetch; ++u) nxt[u] = load_row(j0 + kPrefetch + (uint32_t)u);
#pragma unroll
        for (int u = 0; u < kPrefetch; ++u) {
            const uint32_t j = j0 + (uint32_t)u;
            if (j >= p.G) break;
            const uint32_t blo = (uint32_t)cur[u], bhi = (uint32_t)(cur[u] >> 32);
            if (T > 0) R.add(blo, bhi);
            if (T == 0 || n_fast != 0u) {  // warp-uniform: an item starts counting at its first group and never stops (q = 0)
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
                        if (lane == 0 && net != 0) atomicAdd(reinterpret_cast<int *>(s_dlo) + (T + f) * p.G + j, net);
                    } else {
                        long long mine;
                        if (uniform) {  // every item of the column weighs uw
                            mine = (long long)d * (long long)uw;
                        } else {
                            const uint32_t flo = blo & ~olo & elo[T + f], fhi = bhi & ~ohi & ehi[T + f];
                            mine = (flo | fhi) ? weight_of_bits(flo, fhi, wrow) : 0ll;
                        }
                        const long long net = warp_sum_i64(mine);
                        if (lane == 0 && net != 0)
                            smem_add64(s_dlo, s_dhi, (T + f) * p.G + j, (uint32_t)(unsigned long long)net,
                                       (uint32_t)((unsigned long long)net >> 32));
                    }
                }
            }
            if (T == 0) continue;
            const uint4 *row = reinterpret_cast<const uint4 *>(s_mask + (size_t)j * (T * PP));
#pragma unroll
            for (int t = 0; t < T; ++t) {
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
                R.ge(m, glo, ghi);
                const uint32_t nlo = verdict_update(blo, glo, vlo[t]), nhi = verdict_update(bhi, ghi, vhi[t]);
                if (!WEIGHTED) {
                    const int c = __popc(nlo & elo[t]) + __popc(nhi & ehi[t]);
                    const int net = __reduce_add_sync(0xFFFFFFFFu, c - cnt[t]);
                    cnt[t] = c;
                    if (lane == 0 && net != 0) atomicAdd(reinterpret_cast<int *>(s_dlo) + t * p.G + j, net);
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
                    if (lane == 0 && net != 0)
                        smem_add64(s_dlo, s_dhi, t * p.G + j, (uint32_t)(unsigned long long)net,
                                   (uint32_t)((unsigned long long)net >> 32));
                }
                vlo[t] = nlo;
                vhi[t] = nhi;
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

template <bool WEIGHTED>
int launch_q0(const GmGrowthParams &p, cudaStream_t stream) {  // q = 0 thresholds only
    if (p.n_fast <= 1u) return launch_q<1, 0, 1, WEIGHTED>(p, stream);
    if (p.n_fast <= 2u) return launch_q<1, 0, 2, WEIGHTED>(p, stream);
    return launch_q<1, 0, 4, WEIGHTED>(p, stream);
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
    return ((size_t)G * T * (size_t)((P + 3) & ~3) + G + (size_t)(T + NF) * G * (weighted ? 2u : 1u)) * 4u;
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
    if (p.T == 0) return p.weighted ? launch_q0<true>(p, stream) : launch_q0<false>(p, stream);
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
