// pgx_common.cuh -- shared device helpers (mbarrier / TMA bulk-copy PTX wrappers, smem 64-bit
// accumulation) and host-side error plumbing for libpanacus_b200.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace pgx {

// ---- host-side error handling ------------------------------------------------------------------
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define PGX_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            cudaGetLastError();                                                                    \
            return ::pgx::fail(PGX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
        }                                                                                          \
    } while (0)

// ---- device helpers ----------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
    // make the initialised barriers visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t tx_bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx_bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {  // try_wait suspends the thread in hardware for a bounded time; loop until the phase flips
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes,
                                             uint32_t bar, uint64_t l2_policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
        "l"(src_gmem), "r"(bytes), "r"(bar), "l"(l2_policy)
        : "memory");
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ void lds_v2_u64(uint32_t addr, uint64_t &a, uint64_t &b) {
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}

__device__ __forceinline__ uint64_t lds_u64(uint32_t addr) {
    uint64_t a;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(a) : "r"(addr));
    return a;
}

// 64-bit accumulation in shared memory from 32-bit native atomics (64-bit shared atomics compile
// to CAS spin loops on sm_100): lo += v with the carry forwarded to hi.  Wrapping arithmetic, so
// negative contributions are passed as two's complement (v_hi = 0xFFFFFFFF).
__device__ __forceinline__ void smem_add64(uint32_t *lo, uint32_t *hi, uint32_t idx, uint32_t v_lo,
                                           uint32_t v_hi) {
    const uint32_t old = atomicAdd(&lo[idx], v_lo);
    const uint32_t carry = (old + v_lo < old) ? 1u : 0u;
    if (v_hi + carry != 0u) atomicAdd(&hi[idx], v_hi + carry);
}

#endif  // __CUDACC__

}  // namespace pgx
