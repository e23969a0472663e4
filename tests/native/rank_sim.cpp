// rank_sim.cpp -- host replay of the step functions of k_gm_quorum (panacus_b200/csrc/pgx_rank.cuh) for the CPU
// test-suite: the same RankColumn / mask-table / verdict code the CUDA kernel is built from, run column by column
// on the CPU, so tests/test_rank_sim.py can compare it with the oracle without a GPU.  Test infrastructure only.
#include <cstdint>
#include <vector>

#include "../../panacus_b200/csrc/pgx_rank.cuh"

namespace {

template <int P>
void run(const uint64_t *gm, uint64_t stride, uint64_t n_words, uint64_t n_rows, uint32_t G, const uint32_t *order,
         const uint32_t *thr, uint32_t cov, const uint32_t *countable, const uint32_t *weight, int64_t *delta) {
    constexpr int PP = pgx::RankMaskWords<P>::value;
    std::vector<uint32_t> table((size_t)G * PP);
    for (uint32_t j = 0; j < G; ++j) pgx::rank_mask_row<P>(thr[j], G, table.data() + (size_t)j * PP);
    for (uint32_t j = 0; j < G; ++j) delta[j] = 0;
    for (uint64_t w = 0; w < n_words; ++w) {
        uint32_t elo = ~0u, ehi = ~0u;
        if (cov > 1) {
            elo = ehi = 0;
            for (uint32_t b = 0; b < 64; ++b) {
                const uint64_t item = w * 64 + b;
                const uint32_t c = (item < n_rows && item != 0) ? countable[item] : 0u;
                if (c >= cov) (b < 32 ? elo : ehi) |= 1u << (b & 31u);
            }
        }
        pgx::RankColumn<P> R;
        R.clear();
        uint32_t vlo = 0, vhi = 0;
        int cnt = 0;
        for (uint32_t j = 0; j < G; ++j) {
            const uint64_t b = gm[(uint64_t)order[j] * stride + w];
            const uint32_t blo = (uint32_t)b, bhi = (uint32_t)(b >> 32);
            R.add(blo, bhi);
            uint32_t glo, ghi;
            R.ge(table.data() + (size_t)j * PP, glo, ghi);
            const uint32_t nlo = pgx::verdict_update(blo, glo, vlo), nhi = pgx::verdict_update(bhi, ghi, vhi);
            if (!weight) {
                const int c = __builtin_popcount(nlo & elo) + __builtin_popcount(nhi & ehi);
                delta[j] += c - cnt;
                cnt = c;
            } else {
                const uint64_t up = ((uint64_t)(nhi & ~vhi & ehi) << 32) | (nlo & ~vlo & elo);
                const uint64_t dn = ((uint64_t)(vhi & ~nhi & ehi) << 32) | (vlo & ~nlo & elo);
                for (uint32_t bit = 0; bit < 64; ++bit) {
                    const uint64_t item = w * 64 + bit;
                    if ((up >> bit) & 1u) delta[j] += (int64_t)(item < n_rows ? weight[item] : 0u);
                    if ((dn >> bit) & 1u) delta[j] -= (int64_t)(item < n_rows ? weight[item] : 0u);
                }
            }
            vlo = nlo;
            vhi = nhi;
        }
    }
}

}  // namespace

// planes < 0: the smallest count that fits (rank_planes_needed); otherwise the given instantiation
extern "C" int rank_sim(const uint64_t *gm, uint64_t stride, uint64_t n_words, uint64_t n_rows, uint32_t G,
                        const uint32_t *order, const uint32_t *thr, uint32_t cov, const uint32_t *countable,
                        const uint32_t *weight, int planes, int64_t *delta) {
    const int need = pgx::rank_planes_needed(G);
    const int P = planes < 0 ? need : planes;
    if (P < need) return -1;
#define CASE(N) \
    case N: run<N>(gm, stride, n_words, n_rows, G, order, thr, cov, countable, weight, delta); return N;
    switch (P) {
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(14) CASE(16) CASE(21)
        default: return -2;
    }
#undef CASE
}

extern "C" int rank_planes_needed_c(uint32_t G) { return pgx::rank_planes_needed(G); }
