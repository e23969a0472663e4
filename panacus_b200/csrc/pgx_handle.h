// pgx_handle.h -- the abacus handle behind the C ABI and the helpers its translation units share (pgx_api.cu: handle,
// staging, launch sequences; pgx_comm.cu: NCCL communicator + the sharded entry points).  Not installed.
#pragma once

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "pgx_common.cuh"
#include "pgx_internal.h"

struct pgx_abacus {
    int device = 0;
    int sm_count = 148;
    uint64_t n_items = 0, n_rows = 0;
    uint32_t G = 0, W = 0, Wp = 0;

    uint64_t *d_bitmap = nullptr;
    bool own_bitmap = false;
    uint32_t *d_weight = nullptr;  // nullptr = unit weights
    bool own_weight = false;
    uint32_t max_weight = 1;
    bool max_weight_known = true;

    uint32_t *d_countable = nullptr;  // N+1, lazily allocated
    uint64_t *d_hist_tmp = nullptr;   // histogram by-product of the countable pass
    bool countable_valid = false;

    uint64_t *d_gm = nullptr;  // group-major copy, lazily built
    uint64_t gm_stride = 0;
    bool gm_valid = false;
    // weighted similarity: a second group-major copy with the items sorted by weight (descending), so that
    // most 64-item words carry a single weight (one popcount pass) and high weight planes are empty
    uint64_t *d_gm_w = nullptr;
    uint32_t *d_perm = nullptr, *d_sorted_w = nullptr;
    uint64_t *d_planes = nullptr, *d_uniform_w = nullptr;
    uint32_t *d_plane_mask = nullptr;
    uint32_t n_planes = 0;
    bool planes_valid = false;
    bool planes_cov_order = false;  // items of equal weight are ordered by coverage (k_gm_quorum's warp-uniform skips)

    // general-quorum growth under many orders (counting): a third group-major copy with the items sorted by coverage,
    // so that whole warps of k_gm_quorum hold only items below a threshold's coverage cutoff and skip its rank logic
    uint64_t *d_gm_c = nullptr;
    uint32_t *d_perm_c = nullptr;
    bool gm_c_valid = false;

    uint64_t *d_csr_r = nullptr;  // AbacusByGroup::r (N + 2 row offsets), lazily derived from the bitmap
    bool csr_valid = false;

    uint64_t *d_acc = nullptr;  // self-cleaning global accumulators of k_scan
    size_t acc_words = 0;
    unsigned int *d_ticket = nullptr;  // [0] completion ticket of the exchange epilogue, [1] epoch flag of the direct epilogue
    uint32_t scan_epoch = 0;           // launches of the direct epilogue (value published in d_ticket[1])
    uint32_t scan_launches = 0;        // parity selects the tile counter d_ticket[2 + parity] of a launch
    unsigned int *d_err = nullptr;   // [0]: build / scatter / csr input errors, [1]: fused-exchange watchdog (own word: never mixed)

    uint32_t *d_thr = nullptr;  // quorum thresholds of the current call
    size_t thr_cap = 0;
    std::vector<uint32_t> thr_cache;
    uint32_t *d_order = nullptr;
    size_t order_cap = 0;
    uint32_t *d_identity = nullptr;  // 0..G-1, for general-quorum growth in group order on the group-major copy
    uint64_t *d_scratch = nullptr;  // generic device result buffer
    size_t scratch_cap = 0;

    uint64_t *h_pinned = nullptr;
    size_t pinned_words = 0;

    // ItemTable -> bitmap build pipeline: two device staging buffers filled by a copy stream while k_build consumes
    // the other one; two pinned host buffers when the caller's table is pageable (narrowed to u32 on the way)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_ready[2] = {}, ev_free[2] = {};
    void *d_stage[2] = {};
    size_t d_stage_bytes = 0;
    void *h_stage[2] = {};
    size_t h_stage_bytes = 0;
    uint64_t *d_prefsum = nullptr;
    size_t prefsum_cap = 0;
    int64_t *d_path_group = nullptr;
    size_t path_group_cap = 0;
    uint8_t *d_exclude = nullptr;
    size_t exclude_cap = 0;

    // fused NVLink exchange (item-range sharding)
    unsigned char *d_xchg = nullptr;  // [2 parities][kMaxRanks][acc_words][2] u64 packets {epoch:32 | half:32}
    void *peer_base[pgx::kMaxRanks] = {};
    pgx::Exchange x = {};
    uint32_t epoch = 0;

    // optional kernel timing (pgx_abacus_set_timing): CUDA events around the hot-path kernels of a call, on the handle's stream
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;  // created on demand, reused
    size_t ev_used = 0;

    cudaStream_t own_stream = nullptr, stream = nullptr;
    uint64_t launches = 0;
    std::string last_launch;
};

namespace pgx {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

template <typename T>
int ensure_dev(T **ptr, size_t *cap, size_t count) {
    if (*cap >= count && *ptr) return PGX_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(ptr), std::max<size_t>(count, 1) * sizeof(T)));
    *cap = count;
    return PGX_OK;
}

// RAII bracket of CUDA events around the hot kernels of a call (no-op unless timing is enabled on the handle)
struct KernelTimer {
    pgx_abacus *a;
    cudaEvent_t stop = nullptr;
    explicit KernelTimer(pgx_abacus *h);
    ~KernelTimer();
};
int ensure_pinned(pgx_abacus *a, size_t words);
void invalidate_derived(pgx_abacus *a);
int check_handle(const pgx_abacus *a);
int check_exchange(pgx_abacus *a);
bool all_zero(const uint32_t *thr, uint32_t G);
int validate_thresholds(const pgx_abacus *a, uint32_t T, const uint32_t *cov);
int is_permutation(const uint32_t *order, uint32_t G);
int ensure_countable(pgx_abacus *a);
int ensure_gm(pgx_abacus *a);
int ensure_planes(pgx_abacus *a);
int ensure_max_weight(pgx_abacus *a);
// fused node-major pass -> device buffer in the fused layout (see pgx_fused_pass_async)
int fused_pass(pgx_abacus *a, bool want_cnt, bool want_w, uint32_t T, const uint32_t *cov, const uint32_t *thr, int weighted,
               uint32_t *d_countable, uint64_t *d_out);
// growth on the group-major copy under n_orders device-resident orders -> first differences in device memory
int gm_growth_launch(pgx_abacus *a, uint32_t n_orders, const uint32_t *d_orders, const std::vector<uint32_t> &ts,
                     const uint32_t *cov, const uint32_t *thr, int weighted, uint64_t *d_out_base, uint64_t out_order_stride);
// growth under host-supplied orders -> prefix-summed curves in device memory (d_out: n_orders x T x G, no sync)
int gm_growth_device(pgx_abacus *a, uint32_t n_orders, const uint32_t *orders, uint32_t T, const uint32_t *cov,
                     const uint32_t *thr, int weighted, uint64_t *d_out);
// similarity building blocks on device buffers (no sync): rows [row_begin, row_end) x columns >= col_begin; group totals
// (word_begin / word_end: restrict the sum to the items of these 64-item words; 0 / ~0 = all)
int sim_rows_device(pgx_abacus *a, int weighted, uint32_t row_begin, uint32_t row_end, uint32_t col_begin, uint64_t *d_inter,
                    bool upper_only, uint64_t word_begin, uint64_t word_end);
void sim_block_bounds(uint32_t G, uint32_t world, uint32_t *bounds /* 2 * world + 1 */);
int sim_len_device(pgx_abacus *a, int weighted, uint32_t g_begin, uint32_t g_end, uint64_t *d_len);
// device -> host copy of `words` u64 on the handle's stream + synchronise: straight into `dst` when it is pinned /
// registered host memory, otherwise through the handle's pinned staging buffer
int copy_to_host(pgx_abacus *a, uint64_t *dst, const uint64_t *d_src, size_t words);

}  // namespace pgx
