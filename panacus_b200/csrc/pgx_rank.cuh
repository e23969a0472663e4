// pgx_rank.cuh -- bit-sliced per-item rank counters and the quorum verdict of AbacusByGroup::calc_growth
// (src/graph_broker/abacus.rs:1003-1014) for 64 items at a time.
//
// One "column" = 64 items (one u64 word of the group-major bitmap).  While the groups are walked in counting
// order, rank[i] = number of groups seen so far that contain item i lives in P bit-planes (plane k holds bit k
// of the 64 ranks), kept as 32-bit halves so that every plane operation is exactly one LOP3 per half.  At
// position j the reference tests `rank >= ceil((j + 1) * q)` for the items of group j only (the verdict of an
// item stays what it was at its last own group); the cutoff K is the same for all items, so `rank >= K` is one
// 3-input logic op per plane:  ge <- K_k ? (ge & R_k) : (ge | R_k), scanned from the least significant plane.
// The per-plane masks -K_k (0 or ~0) come from a table (shared memory in the kernel), one row per
// (position, threshold), so the inner loop spends no instruction on rebuilding them.
//
// Everything here is plain integer code shared by the CUDA kernel (pgx_gm.cu: k_gm_quorum) and by the host
// check library tests/native/rank_sim.cpp, which replays the same step functions on the CPU against the oracle.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define PGX_HD __host__ __device__ __forceinline__
#else
#define PGX_HD inline
#endif

namespace pgx {

// masks per (position, threshold) row, padded to a multiple of four words (one or more 16-byte loads)
template <int P>
struct RankMaskWords {
    static constexpr int value = (P + 3) & ~3;
};

// smallest supported plane count for G groups: ranks reach G, and the clamp value G + 1 ("never") must fit
PGX_HD int rank_planes_needed(uint32_t G) {
    int p = 1;
    while (p < 32 && ((1ull << p) - 1ull) < (uint64_t)G + 1ull) ++p;
    return p;
}

// row of the mask table for cutoff K (clamped to G + 1: no rank reaches it): row[k] = 0 - bit k of K
template <int P>
PGX_HD void rank_mask_row(uint32_t K, uint32_t G, uint32_t *row) {
    if (K > G + 1u) K = G + 1u;
#pragma unroll
    for (int k = 0; k < RankMaskWords<P>::value; ++k) row[k] = (k < P) ? (0u - ((K >> k) & 1u)) : 0u;
}

// m ? (g & r) : (g | r) for every bit -- one LOP3 (LUT 0xD4) on the device; nvcc does not merge the three-op form
PGX_HD uint32_t ge_plane(uint32_t g, uint32_t r, uint32_t m) {
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xD4;" : "=r"(d) : "r"(g), "r"(r), "r"(m));
    return d;
#else
    return (m & (g & r)) | (~m & (g | r));
#endif
}

template <int P>
struct RankColumn {
    uint32_t lo[P], hi[P];  // plane k: bit k of the ranks of items 0..31 / 32..63

    PGX_HD void clear() {
#pragma unroll
        for (int k = 0; k < P; ++k) lo[k] = hi[k] = 0u;
    }

    // rank += b (one bit per item): ripple-carry increment.  PE <= P: only the low PE planes take part -- exact while
    // every rank stays below 2^PE (at position j ranks are <= j + 1: the planes above bitlen(j + 1) are still zero)
    template <int PE>
    PGX_HD void add_n(uint32_t blo, uint32_t bhi) {
        uint32_t clo = blo, chi = bhi;
#pragma unroll
        for (int k = 0; k < PE; ++k) {
            const uint32_t tlo = lo[k] & clo, thi = hi[k] & chi;
            lo[k] ^= clo;
            hi[k] ^= chi;
            clo = tlo;
            chi = thi;
        }
    }
    PGX_HD void add(uint32_t blo, uint32_t bhi) { add_n<P>(blo, bhi); }

    // per item: rank >= K, with K given as its mask row (PE as above; K < 2^PE as well: K <= j + 1)
    template <int PE>
    PGX_HD void ge_n(const uint32_t *mask_row, uint32_t &glo, uint32_t &ghi) const {
        glo = ghi = ~0u;
#pragma unroll
        for (int k = 0; k < PE; ++k) {
            const uint32_t m = mask_row[k];
            glo = ge_plane(glo, lo[k], m);
            ghi = ge_plane(ghi, hi[k], m);
        }
    }
    PGX_HD void ge(const uint32_t *mask_row, uint32_t &glo, uint32_t &ghi) const { ge_n<P>(mask_row, glo, ghi); }
};

// Verdict update of one threshold at one position: items of the group (b) take the fresh test result, all
// others keep theirs (abacus.rs:1007-1010: k only advances at the item's own groups).
PGX_HD uint32_t verdict_update(uint32_t b, uint32_t ge, uint32_t verdict) { return (b & ge) | (~b & verdict); }

}  // namespace pgx
