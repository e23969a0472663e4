// pgx_simmma.cu -- all-pairs group intersections on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// inter[x][y] = #items in both groups = (B^T B)[x][y] for the 0/1 incidence matrix B -- the integer part of
// Similarity::set_table (reference src/analyses/similarity.rs:125-150).  On the CUDA cores this is AND + POPC over
// 64-item words (k_gm_similarity: 30.7 ms at 10M x 1024, both integer pipes ~75 % busy).  Here the bits of the
// group-major bitmap are expanded to u8 0/1 operands in shared memory and multiplied with tcgen05.mma.kind::i8
// (u8 x u8 -> s32, exact): one CTA owns a 128 x 256 tile of the matrix over a range of items,
//   * 12 loader warps: one thread per operand row (128 A rows + 256 B rows) streams the row's bits, 16 bytes = 128 items
//     per stage, global -> registers (4 loads in flight) -> a ring of raw slots in shared memory.  They execute no
//     fence, so their loads overlap freely;
//   * 12 expander warps: one thread per operand row turns the 128 bits of its row into 128 bytes of the K-major
//     SWIZZLE_128B UMMA layout (8-row groups of 1 KB, 16-byte chunks XORed with the row index), fences the
//     generic-proxy stores for the async proxy and arrives on full[stage].  (In the first version the expanders loaded
//     their bits themselves: `fence.proxy.async` is a MEMBAR.ALL.CTA in SASS and waited for the prefetches it followed --
//     one DRAM latency per stage.)
//   * 1 MMA warp: one elected lane issues four m128 n256 k32 MMAs per stage into a 128-lane x 256-column s32
//     accumulator in TMEM and commits them to empty[stage];
//   * epilogue (8 warps): tcgen05.ld 32 lanes x 32 columns at a time -> u64 atomicAdd into the caller's matrix (the item
//     range is split over gridDim.z CTAs per tile).
// Counts only (unit weights); bp-weighted similarity stays on k_gm_similarity.
#include <cstdlib>
#include <cstring>

#include "pgx_common.cuh"
#include "pgx_internal.h"

namespace pgx {

namespace {

constexpr int kMmaM = 128, kMmaN = 256, kMmaK = 32;  // one tcgen05.mma: u8, K = 32 bytes
constexpr int kStageItems = 128;                      // items per stage = one 16-byte load per operand row
constexpr int kKBlocks = kStageItems / kMmaK;         // MMAs per stage
constexpr int kStages = 3;
constexpr int kRawSlots = 8;                          // ring of raw (bit) slots between loaders and expanders
constexpr int kExpanders = kMmaM + kMmaN;             // one thread per operand row
constexpr int kMmaWarp = kExpanders / 32;             // warp 12
constexpr int kLoaderWarp0 = kMmaWarp + 1;            // warps 13 .. 24: loaders, one thread per operand row
constexpr int kSimMmaThreads = 2 * kExpanders + 32;
constexpr uint32_t kRawBytes = kExpanders * 16u;      // 6 KB per raw slot
constexpr uint32_t kABytes = kMmaM * kStageItems;     // 16 KB per stage
constexpr uint32_t kBBytes = kMmaN * kStageItems;     // 32 KB per stage
constexpr uint32_t kStageBytes = kABytes + kBBytes;
constexpr uint32_t kTmemCols = 256;
constexpr int kPrefetch = 3;                          // global loads in flight per expander thread

// K-major SWIZZLE_128B operand (cute/arch/mma_sm100_desc.hpp): rows of 128 bytes, 8-row groups of 1 KB (SBO), the
// 16-byte chunks of a row XORed with (row % 8); LBO is unused for swizzled K-major layouts (1 by convention); version 1;
// layout type 2 in bits 61..63.  K advances inside the 128-byte atom by moving the start address (32 bytes per MMA).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// instruction descriptor of kind::i8: D = s32 (bits 4-5 = 2), A and B unsigned 8 bit (0), both K-major (bits 15, 16 = 0),
// N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (2u << 4) | ((uint32_t)(kMmaN >> 3) << 17) | ((uint32_t)(kMmaM >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(kIdesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 4 bits -> 4 bytes of 0 / 1: the four partial products of the multiply land on disjoint bit ranges (no carries)
__device__ __forceinline__ uint32_t spread4(uint32_t nibble) { return (nibble * 0x00204081u) & 0x01010101u; }

__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(kSimMmaThreads, 1) k_sim_mma(const __grid_constant__ GmSimParams p, uint32_t words_per_split) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long s_bar[2 * kStages + 2 * kRawSlots + 1];
    __shared__ uint32_t s_tmem;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t x0 = p.row_begin + blockIdx.y * kMmaM;  // output rows of the tile
    const uint32_t y0 = p.col_begin + blockIdx.x * kMmaN;  // output columns
    if (x0 >= p.row_end || y0 >= p.G) return;
    if ((p.upper_only || p.triangular) && y0 + kMmaN <= x0) return;  // strictly below the diagonal: the mirror image is computed
    const uint64_t k_begin = (uint64_t)blockIdx.z * words_per_split;  // 64-item words, even
    uint64_t k_end = k_begin + words_per_split;
    if (k_end > p.n_words) k_end = p.n_words;
    if (k_begin >= k_end) return;
    const uint32_t n_it = (uint32_t)((k_end - k_begin + 1u) / 2u);  // stages of two words

    const uint32_t stage0 = (smem_u32(smem) + 1023u) & ~1023u;  // SWIZZLE_128B atoms are 1 KB aligned (2 KB of slack is allocated)
    const uint32_t raw0 = stage0 + (uint32_t)kStages * kStageBytes;
    // barriers: full[kStages] | empty[kStages] | raw_full[kRawSlots] | raw_empty[kRawSlots] | accum
    const uint32_t full0 = smem_u32(s_bar), empty0 = full0 + 8u * kStages, rawfull0 = empty0 + 8u * kStages,
                   rawempty0 = rawfull0 + 8u * kRawSlots, accum_bar = rawempty0 + 8u * kRawSlots;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8u * s, (uint32_t)kMmaWarp);  // one arrival per expander warp
            mbar_init(empty0 + 8u * s, 1u);                 // tcgen05.commit
        }
        for (int s = 0; s < kRawSlots; ++s) {
            mbar_init(rawfull0 + 8u * s, (uint32_t)kMmaWarp);   // one arrival per loader warp
            mbar_init(rawempty0 + 8u * s, (uint32_t)kMmaWarp);  // one arrival per expander warp
        }
        mbar_init(accum_bar, 1u);
        mbar_fence_init();
    }
    if (warp == 0) {  // TMEM: 256 columns x 128 lanes of s32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp >= (uint32_t)kLoaderWarp0) {
        // ===== loaders: one operand row per thread, global -> registers -> raw slot =====
        const uint32_t t = tid - (uint32_t)kLoaderWarp0 * 32u;  // operand row 0 .. 383 (A rows first)
        const bool is_a = t < (uint32_t)kMmaM;
        const uint32_t g = is_a ? x0 + t : y0 + (t - kMmaM);
        const bool live = is_a ? g < p.row_end : g < p.G;
        const uint64_t *row = p.gm + (uint64_t)(live ? g : 0u) * p.gm_stride;
        auto fetch = [&](uint32_t it) -> ulonglong2 {
            const uint64_t w = k_begin + 2ull * it;
            ulonglong2 v = make_ulonglong2(0ull, 0ull);
            if (live && it < n_it) {
                if (w + 1u < k_end) {
                    v = __ldg(reinterpret_cast<const ulonglong2 *>(row + w));
                } else if (w < k_end) {
                    v.x = __ldg(row + w);
                }
            }
            return v;
        };
        constexpr int kRing = 4;  // loads in flight per thread (kRawSlots is a multiple of it)
        ulonglong2 ring[kRing];
#pragma unroll
        for (int u = 0; u < kRing; ++u) ring[u] = fetch((uint32_t)u);
        uint32_t slot = 0, ph = 0;
        for (uint32_t it0 = 0; it0 < n_it; it0 += kRing) {
#pragma unroll
            for (int u = 0; u < kRing; ++u) {
                const uint32_t it = it0 + (uint32_t)u;
                if (it >= n_it) break;
                const ulonglong2 v = ring[u];
                ring[u] = fetch(it + kRing);
                if (it >= (uint32_t)kRawSlots) mbar_wait(rawempty0 + 8u * slot, ph ^ 1u);
                sts_v4(raw0 + slot * kRawBytes + t * 16u, (uint32_t)v.x, (uint32_t)(v.x >> 32), (uint32_t)v.y, (uint32_t)(v.y >> 32));
                __syncwarp();
                if (lane == 0) mbar_arrive(rawfull0 + 8u * slot);
                if (++slot == (uint32_t)kRawSlots) {
                    slot = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp < (uint32_t)kMmaWarp) {
        // ===== expanders: one operand row per thread, raw bits -> u8 operand rows =====
        const bool is_a = tid < (uint32_t)kMmaM;
        const uint32_t r = is_a ? tid : tid - kMmaM;  // row inside the A / B block
        // the row's 128 bytes inside a stage: 8-row groups of 1 KB; chunk c of the row lives at chunk (c ^ (r % 8))
        const uint32_t row_off = (is_a ? 0u : kABytes) + (r >> 3) * 1024u + (r & 7u) * 128u;
        const uint32_t sw = r & 7u;
        uint32_t st = 0, ph = 0, slot = 0, rph = 0;
        for (uint32_t it = 0; it < n_it; ++it) {
            mbar_wait(rawfull0 + 8u * slot, rph);
            uint32_t w32[4];
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w32[0]), "=r"(w32[1]), "=r"(w32[2]), "=r"(w32[3]) : "r"(raw0 + slot * kRawBytes + tid * 16u));
            if (it >= (uint32_t)kStages) mbar_wait(empty0 + 8u * st, ph ^ 1u);  // the MMAs that read this stage are done
            const uint32_t base = stage0 + st * kStageBytes + row_off;
            if (p.csa == 5u) {  // (timing experiment PGX_SIM_DEBUG=5: the ALU work of the expansion, one store instead of eight)
                uint32_t acc[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int kb = 0; kb < kKBlocks; ++kb) {
                    const uint32_t bits = w32[kb];
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j & 3] ^= spread4((bits >> (4 * j)) & 0xFu) << (j >> 2);
                }
                sts_v4(base + (sw << 4), acc[0], acc[1], acc[2], acc[3]);
            } else if (p.csa != 2u)  // (timing experiment PGX_SIM_DEBUG=2: no expansion, results are wrong)
#pragma unroll
                for (int kb = 0; kb < kKBlocks; ++kb) {  // 32 items = one MMA's K = chunks 2 kb, 2 kb + 1
                    const uint32_t bits = w32[kb];
                    sts_v4(base + (((uint32_t)(2 * kb) ^ sw) << 4), spread4(bits & 0xFu), spread4((bits >> 4) & 0xFu),
                           spread4((bits >> 8) & 0xFu), spread4((bits >> 12) & 0xFu));
                    sts_v4(base + (((uint32_t)(2 * kb + 1) ^ sw) << 4), spread4((bits >> 16) & 0xFu), spread4((bits >> 20) & 0xFu),
                           spread4((bits >> 24) & 0xFu), spread4(bits >> 28));
                }
            if (p.csa != 4u)  // (timing experiment PGX_SIM_DEBUG=4: no proxy fence, results may be wrong)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(rawempty0 + 8u * slot);  // (the raw bits are in registers and used)
                mbar_arrive(full0 + 8u * st);
            }
            if (++st == (uint32_t)kStages) {
                st = 0;
                ph ^= 1u;
            }
            if (++slot == (uint32_t)kRawSlots) {
                slot = 0;
                rph ^= 1u;
            }
        }
    } else {
        // ===== MMA issuer: one elected lane =====
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            for (uint32_t it = 0; it < n_it; ++it) {
                mbar_wait(full0 + 8u * st, ph);
                tc_fence_after();
                const uint32_t a0 = stage0 + st * kStageBytes, b0 = a0 + kABytes;
                if (p.csa != 3u || it == 0u)  // (timing experiment PGX_SIM_DEBUG=3: only the first stage's MMAs)
#pragma unroll
                    for (int kb = 0; kb < kKBlocks; ++kb)
                        umma_i8(tmem, umma_desc_sw128(a0 + (uint32_t)kb * kMmaK), umma_desc_sw128(b0 + (uint32_t)kb * kMmaK),
                                (it | (uint32_t)kb) ? 1u : 0u);
                umma_commit(empty0 + 8u * st);  // frees the stage when these MMAs have read it
                if (++st == (uint32_t)kStages) {
                    st = 0;
                    ph ^= 1u;
                }
            }
            umma_commit(accum_bar);  // all MMAs of the tile done: the accumulator is final
        }
    }

    // ===== epilogue: TMEM -> u64 atomics (8 warps: lane quadrant = warp % 4, column half = warp / 4) =====
    if (warp < 8u) {
        mbar_wait(accum_bar, 0u);
        tc_fence_after();
        const uint32_t q = warp & 3u, half = warp >> 2;
        const uint32_t x = x0 + q * 32u + lane;
#pragma unroll 1
        for (uint32_t cb = 0; cb < 4u; ++cb) {
            const uint32_t col0 = half * 128u + cb * 32u;
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
                "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                  "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                  "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(tmem + ((q * 32u) << 16) + col0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (x < p.row_end) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const uint32_t y = y0 + col0 + (uint32_t)j;
                    if (y < p.G && v[j])
                        atomicAdd(reinterpret_cast<unsigned long long *>(p.inter + (uint64_t)(x - p.row_begin) * p.G + y),
                                  (unsigned long long)v[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace

// Unweighted intersections of the rows [row_begin, row_end) x columns >= col_begin on the tensor cores; same contract as
// launch_gm_similarity (p.inter zeroed by the caller; triangular / upper_only skip tiles strictly below the diagonal).
int launch_sim_mma(const GmSimParams &p, int sm_count, cudaStream_t stream) {
    const uint32_t rows = p.row_end - p.row_begin;
    if (rows == 0 || p.col_begin >= p.G) return PGX_OK;
    const uint32_t ty = (rows + kMmaM - 1) / kMmaM, tx = (p.G - p.col_begin + kMmaN - 1) / kMmaN;
    // tiles that are actually computed (upper part when the launch is symmetric), to size the item split
    uint64_t live = 0;
    for (uint32_t by = 0; by < ty; ++by)
        for (uint32_t bx = 0; bx < tx; ++bx) {
            const uint32_t x0 = p.row_begin + by * kMmaM, y0 = p.col_begin + bx * kMmaN;
            if ((p.upper_only || p.triangular) && y0 + kMmaN <= x0) continue;
            ++live;
        }
    if (live == 0) return PGX_OK;
    uint64_t splits = ((uint64_t)sm_count * 2u + live - 1u) / live;
    const uint64_t max_splits = (p.n_words + 1u) / 2u;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1u) splits = 1u;
    if (splits > 65535u) splits = 65535u;
    uint64_t wps = (p.n_words + splits - 1u) / splits;
    wps = (wps + 1u) & ~1ull;  // whole stages of two words; keeps the 16-byte loads aligned
    splits = (p.n_words + wps - 1u) / wps;
    const size_t smem = (size_t)kStages * kStageBytes + (size_t)kRawSlots * kRawBytes + 2048u;
    PGX_CUDA(cudaFuncSetAttribute(k_sim_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sim_mma<<<dim3(tx, ty, (unsigned)splits), kSimMmaThreads, smem, stream>>>(p, (uint32_t)wps);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

}  // namespace pgx
