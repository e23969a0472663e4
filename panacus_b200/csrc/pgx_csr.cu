// pgx_csr.cu -- AbacusByGroup's CSR {r, c, v} (src/graph_broker/abacus.rs:790-799) derived on the device.
//
// The reference builds it with two serial scatter passes over the ItemTable (compute_row_storage_space,
// abacus.rs:859-899; compute_column_values, abacus.rs:901-986).  Here the bitmap already holds the
// de-duplicated incidence, so
//   r  = exclusive prefix sum of the row popcounts (k_csr_row_len + cub::DeviceScan), N + 2 entries, r[0] = r[1] = 0
//   c  = the positions of the set bits of each row, ascending (k_csr_cols)
//   v  = occurrence counts: every ItemTable step adds 1 at r[item] + rank(group bit in the item's row) (k_csr_vals);
//        integer atomics, so the result does not depend on the step order
// which is exactly the sorted CSR the reference's cursor scheme ends up with.
// Consumers: AbacusByGroup::to_tsv (abacus.rs:1056-1178), the `table` analysis.
#include <cub/device/device_scan.cuh>

#include "pgx_common.cuh"
#include "pgx_internal.h"

namespace pgx {

namespace {

__device__ __forceinline__ uint64_t row_word(const uint64_t *__restrict__ bitmap, uint64_t item, uint32_t Wp, uint32_t w,
                                             uint32_t G) {
    uint64_t x = __ldg(bitmap + item * Wp + w);
    const uint32_t base = w << 6;
    if (base + 64u > G) x &= (G > base) ? (~0ull >> (64u - (G - base))) : 0ull;  // bits >= G are padding
    return x;
}

// len[i] = number of groups containing item i (i = 1..N); len[0] = len[N+1] = 0
__global__ void __launch_bounds__(256) k_csr_row_len(const uint64_t *__restrict__ bitmap, uint64_t n_rows, uint32_t G, uint32_t W,
                                                     uint32_t Wp, uint64_t *__restrict__ len) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_rows; i += stride) {
        uint32_t n = 0;
        if (i != 0 && i < n_rows)
            for (uint32_t w = 0; w < W; ++w) n += (uint32_t)__popcll(row_word(bitmap, i, Wp, w, G));
        len[i] = n;
    }
}

__global__ void __launch_bounds__(256) k_csr_cols(const uint64_t *__restrict__ bitmap, uint64_t n_rows, uint32_t G, uint32_t W,
                                                  uint32_t Wp, const uint64_t *__restrict__ r, uint64_t *__restrict__ c) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1u; i < n_rows; i += stride) {
        uint64_t k = r[i];
        for (uint32_t w = 0; w < W; ++w) {
            uint64_t x = row_word(bitmap, i, Wp, w, G);
            while (x) {
                const uint32_t b = (uint32_t)__ffsll((long long)x) - 1u;
                c[k++] = (uint64_t)(w << 6) + b;
                x &= x - 1u;
            }
        }
    }
}

// items[] holds steps [step0, step0 + n_steps) of the ItemTable; same path lookup and filters as k_build
__global__ void __launch_bounds__(256) k_csr_vals(const uint64_t *__restrict__ bitmap, uint64_t n_rows, uint32_t G, uint32_t Wp,
                                                  const uint64_t *__restrict__ items, uint64_t step0, uint64_t n_steps,
                                                  const uint64_t *__restrict__ prefsum, uint64_t n_paths,
                                                  const int64_t *__restrict__ path_group, const uint8_t *__restrict__ exclude,
                                                  const uint64_t *__restrict__ r, uint32_t *__restrict__ v, unsigned int *err) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_steps; k += stride) {
        const uint64_t s = step0 + k;
        uint64_t lo = 0, hi = n_paths;  // largest p with prefsum[p] <= s
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (__ldg(prefsum + mid) <= s) lo = mid; else hi = mid;
        }
        const long long grp = __ldg(path_group + lo);
        if (grp < 0) continue;
        if ((unsigned long long)grp >= G) {
            atomicOr(err, 4u);
            continue;
        }
        const uint64_t id = __ldg(items + k);
        if (id == 0 || id >= n_rows) {
            atomicOr(err, 1u);
            continue;
        }
        if (exclude && __ldg(exclude + id)) continue;
        const uint32_t g = (uint32_t)grp, gw = g >> 6;
        uint32_t rank = 0;
        for (uint32_t w = 0; w < gw; ++w) rank += (uint32_t)__popcll(row_word(bitmap, id, Wp, w, G));
        const uint64_t x = row_word(bitmap, id, Wp, gw, G);
        const uint64_t bit = 1ull << (g & 63u);
        if (!(x & bit)) {  // the bitmap was not built from this table
            atomicOr(err, 8u);
            continue;
        }
        rank += (uint32_t)__popcll(x & (bit - 1u));
        atomicAdd(v + r[id] + rank, 1u);
    }
}

unsigned grid_for(uint64_t n) {
    uint64_t blocks = (n + 255u) / 256u;
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    return (unsigned)(blocks ? blocks : 1u);
}

}  // namespace

int launch_csr_rows(const uint64_t *bitmap, uint64_t n_rows, uint32_t G, uint32_t W, uint32_t Wp, uint64_t *d_r,
                    cudaStream_t stream) {
    const uint64_t n = n_rows + 1u;  // N + 2 entries
    k_csr_row_len<<<grid_for(n), 256, 0, stream>>>(bitmap, n_rows, G, W, Wp, d_r);
    PGX_CUDA(cudaGetLastError());
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_r, d_r, n, stream);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_r, d_r, n, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(tmp);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(PGX_ERR_CUDA, std::string("csr row scan: ") + cudaGetErrorString(e));
    }
    return PGX_OK;
}

int launch_csr_cols(const uint64_t *bitmap, uint64_t n_rows, uint32_t G, uint32_t W, uint32_t Wp, const uint64_t *d_r,
                    uint64_t *d_c, cudaStream_t stream) {
    if (n_rows <= 1) return PGX_OK;
    k_csr_cols<<<grid_for(n_rows), 256, 0, stream>>>(bitmap, n_rows, G, W, Wp, d_r, d_c);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

int launch_csr_vals(const uint64_t *bitmap, uint64_t n_rows, uint32_t G, uint32_t Wp, const uint64_t *d_items, uint64_t step0,
                    uint64_t n_steps, const uint64_t *d_prefsum, uint64_t n_paths, const int64_t *d_path_group,
                    const uint8_t *d_exclude, const uint64_t *d_r, uint32_t *d_v, unsigned int *d_err, cudaStream_t stream) {
    if (n_steps == 0) return PGX_OK;
    k_csr_vals<<<grid_for(n_steps), 256, 0, stream>>>(bitmap, n_rows, G, Wp, d_items, step0, n_steps, d_prefsum, n_paths,
                                                      d_path_group, d_exclude, d_r, d_v, d_err);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

}  // namespace pgx
