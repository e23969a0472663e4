// pgx_scan.cu -- the fused node-major pass: coverage histogram + ordered growth in one sweep over
// the abacus bitmap.  Replaces AbacusByTotal::coverage + construct_hist[_bps]
// (reference src/graph_broker/abacus.rs:719-787) and AbacusByGroup::calc_growth (abacus.rs:989-1032).
//
// Structure (one persistent CTA per SM slot, 8 scanning warps + 1 TMA producer warp):
//   producer lane : tiles from a global counter (a CTA's first one is blockIdx.x); 1-D TMA bulk copies
//                   (cp.async.bulk, SASS UBLKCP) of `tile_items` bitmap rows (+ their u32 weights)
//                   into a ring of shared-memory stages, completion on mbarriers (scan_producer)
//   consumers     : k_scan       one thread per item; 128-bit shared loads in a per-lane rotated chunk
//                                order (bank-conflict free for any row width), popcount -> coverage,
//                                first set bit -> growth column, 32-bit shared atomics
//                   k_scan_priv  G <~ 300: lane-private narrow counters instead of the atomics
//                   k_scan_vert  G <= 128, counting: bit-sliced vertical counters (carry-save adders
//                                over one-hot words), no atomics in the loop
//   epilogue      : single GPU: CTA 0 zeroes the result vector and publishes an epoch, every CTA adds
//                   its sums straight into it (RED.64) -- the end of the kernel is the end of the pass;
//                   multi-GPU exchange: global accumulators, completion ticket, the last CTA pushes the
//                   vector to its peers over NVLink and sums their slots (scan_epilogue)
//
// Integer-only, HBM-bandwidth-bound: algorithmic bytes per item = W*8 (+4 weighted).
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "pgx_common.cuh"
#include "pgx_internal.h"

namespace pgx {

namespace {

__device__ __forceinline__ uint32_t first_bit(uint64_t x) { return (uint32_t)__ffsll((long long)x) - 1u; }
__device__ __forceinline__ uint32_t last_bit(uint64_t x) { return 63u - (uint32_t)__clzll((long long)x); }

struct SmemAcc {
    uint32_t *hist_cnt, *hist_wlo, *hist_whi, *delta_lo, *delta_hi;
    uint32_t *joint_cnt, *joint_wlo, *joint_whi;
    const uint32_t *thr;
};

// ---- q = 0 path: hist bin + first-set-bit delta -------------------------------------------------
// (Keeping the dominant bins -- coverage 1, first group 0 -- in registers and only the rest on shared atomics was
// measured SLOWER, 10M x 1024: 0.195 -> 0.225 ms, 10M x 256: 80 -> 96 us, profiles/r2_scan_shapes_v3_regacc.jsonl: the
// divergent branch costs more than the same-address serialisation it removes.)
__device__ __forceinline__ void account_fast(const ScanParams &p, const SmemAcc &s, uint64_t item, uint32_t cov,
                                             uint32_t first, uint32_t wgt) {
    if (p.countable) p.countable[item] = cov;
    if (p.flags & kJoint) {
        // small G: one bin per (coverage, first group); hist and every q = 0 curve are marginals of it, so an
        // item costs one shared atomic (count) / one lo+carry pair (weight) no matter how many thresholds.
        // (Warp-aggregating equal bins with match.any first was measured slower: 79 vs 51 us at 10M x 44.)
        const uint32_t bin = cov * p.G + (cov ? first : 0u);
        if ((p.flags & kHistCount) || !(p.flags & kWeighted)) atomicAdd(&s.joint_cnt[bin], 1u);
        if (p.flags & (kHistWeight | kWeighted)) smem_add64(s.joint_wlo, s.joint_whi, bin, wgt, 0u);
        return;
    }
    if (p.flags & kHistCount) atomicAdd(&s.hist_cnt[cov], 1u);
    if (p.flags & kHistWeight) smem_add64(s.hist_wlo, s.hist_whi, cov, wgt, 0u);
    if (cov == 0) return;
    for (uint32_t t = 0; t < p.T; ++t) {
        if (cov >= p.cov[t]) {
            if (p.flags & kWeighted)
                smem_add64(s.delta_lo + t * p.G, s.delta_hi + t * p.G, first, wgt, 0u);
            else
                atomicAdd(&s.delta_lo[t * p.G + first], 1u);
        }
    }
}

// word mask for natural-order readers: bits >= G and the padding word are ignored
__device__ __forceinline__ uint64_t word_mask(const ScanParams &p, uint32_t w) {
    if (w + 1 < p.W) return ~0ull;
    if (w + 1 == p.W) return (p.G & 63u) ? ((1ull << (p.G & 63u)) - 1ull) : ~0ull;
    return 0ull;
}

template <typename Reader>
__device__ __forceinline__ void item_fast_natural(const ScanParams &p, const SmemAcc &s, uint64_t item,
                                                  uint32_t wgt, Reader rd) {
    uint32_t cov = 0, first = 0xFFFFFFFFu;
    for (uint32_t w = 0; w < p.W; ++w) {
        const uint64_t x = rd(w) & word_mask(p, w);
        cov += __popcll(x);
        if (x && first == 0xFFFFFFFFu) first = w * 64u + first_bit(x);
    }
    account_fast(p, s, item, cov, first, wgt);
}

// ---- general quorum path (abacus.rs:1004-1010 on the bitmap) --------------------------------------
// For an item with bits b: at its k-th set bit (1-based) in column g the verdict becomes
// (k >= thr[g]) and holds until the next set bit; the curve's first difference at g changes by
// +-wgt whenever the verdict flips (it starts at "not counted").
__device__ __forceinline__ void delta_add(const ScanParams &p, const SmemAcc &s, uint32_t t, uint32_t col,
                                          uint32_t wgt, bool up) {
    if (p.flags & kWeighted) {
        if (up)
            smem_add64(s.delta_lo + t * p.G, s.delta_hi + t * p.G, col, wgt, 0u);
        else if (wgt)
            smem_add64(s.delta_lo + t * p.G, s.delta_hi + t * p.G, col, 0u - wgt, 0xFFFFFFFFu);
    } else {
        atomicAdd(&s.delta_lo[t * p.G + col], up ? 1u : 0xFFFFFFFFu);
    }
}

template <typename Reader>
__device__ __forceinline__ void item_quorum_natural(const ScanParams &p, const SmemAcc &s, uint64_t item,
                                                    uint32_t wgt, Reader rd) {
    uint32_t cov = 0;
    for (uint32_t w = 0; w < p.W; ++w) cov += __popcll(rd(w) & word_mask(p, w));
    if (p.countable) p.countable[item] = cov;
    uint32_t elig = 0;
    for (uint32_t t = 0; t < p.T; ++t) elig |= (cov >= p.cov[t] ? 1u : 0u) << t;
    if (cov == 0 || elig == 0) return;
    uint32_t rank = 0, verdict = 0;  // verdict: bit t = item currently counted for threshold t
    for (uint32_t w = 0; w < p.W; ++w) {
        const uint64_t x = rd(w) & word_mask(p, w);
        if (!x) continue;
        const uint32_t c = __popcll(x), fb = first_bit(x), lb = last_bit(x), col0 = w * 64u;
        for (uint32_t t = 0; t < p.T; ++t) {
            if (!((elig >> t) & 1u)) continue;
            const uint32_t *th = s.thr + t * p.G + col0;
            bool prev = (verdict >> t) & 1u;
            if (rank + 1 >= th[lb]) {  // every set bit of this word passes
                if (!prev) delta_add(p, s, t, col0 + fb, wgt, true);
                prev = true;
            } else if (rank + c < th[fb]) {  // none passes
                if (prev) delta_add(p, s, t, col0 + fb, wgt, false);
                prev = false;
            } else {
                uint64_t y = x;
                uint32_t k = rank;
                while (y) {
                    const uint32_t b = first_bit(y);
                    y &= y - 1;
                    ++k;
                    const bool v = k >= th[b];
                    if (v != prev) {
                        delta_add(p, s, t, col0 + b, wgt, v);
                        prev = v;
                    }
                }
            }
            verdict = (verdict & ~(1u << t)) | ((prev ? 1u : 0u) << t);
        }
        rank += c;
    }
}

struct SmemRowReader {
    uint32_t addr;
    __device__ __forceinline__ uint64_t operator()(uint32_t w) const { return lds_u64(addr + w * 8u); }
};
struct GlobalRowReader {
    const uint64_t *row;
    __device__ __forceinline__ uint64_t operator()(uint32_t w) const { return __ldg(row + w); }
};

// ---- rotated-chunk fast path over a shared-memory row --------------------------------------------
// A row is C = Wp/2 chunks of 16 bytes.  Lane i visits the chunks in a rotated order so that the 8
// lanes of a quarter-warp hit 8 different 16-byte bank groups for every row width:
//   bank group of (i, c) = (i*C + c) mod 8;  with 2^a | C,  r(i) = (i >> (3-a')) & (2^a' - 1), a' = min(a,3);
//   chunk visited at step k: k ^ r(i) when C is a power of two, (k + r(i)) mod C otherwise.
// popcount and "first non-empty chunk" are order independent, so the rotation costs nothing; the
// first set bit is extracted once, after the loop, from the first non-empty chunk.
template <int C_T, bool MASK>
__device__ __forceinline__ void row_cov_first(const ScanParams &p, uint32_t row_addr, uint32_t li, uint32_t C_rt,
                                              uint32_t rot_shift, uint32_t rot_mask, uint32_t &cov_out, uint32_t &first_out) {
    constexpr bool kXor = C_T > 0 && (C_T & (C_T - 1)) == 0;
    const uint32_t C = C_T > 0 ? (uint32_t)C_T : C_rt;
    const uint32_t r = (li >> rot_shift) & rot_mask;
    uint32_t cov = 0, firstc = 0xFFFFFFFFu;
#pragma unroll(C_T > 0 ? C_T : 4)
    for (uint32_t k = 0; k < C; ++k) {
        uint32_t c;
        if (kXor) {
            c = k ^ r;
        } else {
            c = k + r;
            if (c >= C) c -= C;
        }
        uint64_t x, y;
        lds_v2_u64(row_addr + c * 16u, x, y);
        if (MASK && c == C - 1) {
            x &= p.last_mask0;
            y &= p.last_mask1;
        }
        cov += __popcll(x) + __popcll(y);
        if ((x | y) != 0ull) firstc = min(firstc, c);
    }
    uint32_t first = 0xFFFFFFFFu;
    if (cov) {
        uint64_t x, y;
        lds_v2_u64(row_addr + firstc * 16u, x, y);
        if (MASK && firstc == C - 1) {
            x &= p.last_mask0;
            y &= p.last_mask1;
        }
        first = firstc * 128u + (x ? first_bit(x) : 64u + first_bit(y));
    }
    cov_out = cov;
    first_out = first;
}

template <int C_T, bool MASK>
__device__ __forceinline__ void item_fast_smem(const ScanParams &p, const SmemAcc &s, uint64_t item, uint32_t wgt,
                                               uint32_t row_addr, uint32_t li, uint32_t C_rt, uint32_t rot_shift,
                                               uint32_t rot_mask) {
    uint32_t cov, first;
    row_cov_first<C_T, MASK>(p, row_addr, li, C_rt, rot_shift, rot_mask, cov, first);
    account_fast(p, s, item, cov, first, wgt);
}

// tuning overrides for experiments (PGX_SCAN_TILE / PGX_SCAN_STAGES / PGX_SCAN_CTAS), 0 = automatic
uint32_t env_u32(const char *name) {
    const char *v = getenv(name);
    return v ? (uint32_t)strtoul(v, nullptr, 10) : 0u;
}

// ---- tile scheduling + TMA producer (one elected lane per CTA; shared by k_scan, k_scan_priv, k_scan_vert) --------------
// A CTA's first tile is blockIdx.x; every further one comes from a global counter (p.tile_ctr[parity], requested one tile
// ahead so that the atomic's round trip hides behind the ring): CTAs whose first bytes arrive late -- the first tiles of
// a launch take 2 .. 8 us on this part, profiles/r2_scan_timeline_v1.txt -- simply take fewer tiles instead of making
// the whole grid wait for them.  The tile a stage holds is published in s_tile[stage] before the stage's "full" barrier
// is armed; index >= n_tiles is the end-of-work marker.  CTA 0 zeroes the OTHER parity's counter for the next launch.
// max_tiles: most tiles this CTA may take (k_scan_vert's counters have a capacity; the others pass ~0).
__device__ __forceinline__ void scan_producer(const ScanParams &p, uint32_t full0, uint32_t empty0, uint32_t stage0,
                                              volatile uint32_t *s_tile, uint32_t max_tiles) {
    const uint64_t pol = l2_policy_evict_first();
    const uint32_t S = p.stages, rowbytes = p.Wp * 8u;
    const uint32_t last_tile = p.n_tiles - 1u;
    const uint32_t last_rows = (uint32_t)(p.n_rows - (uint64_t)last_tile * p.tile_items);
    unsigned int *ctr = p.ticket + 2u + (p.sched_parity & 1u);
    if (blockIdx.x == 0u) p.ticket[2u + ((p.sched_parity & 1u) ^ 1u)] = 0u;
    const bool dynamic = p.sched_dynamic != 0u;
    uint32_t st = 0, ph = 0, taken = 0;
    bool ring_full = false;  // true once every stage has been used at least once
    uint32_t tile = blockIdx.x;
    for (;;) {
        // the tile after this one: requested now, needed one iteration later
        uint32_t nxt = 0xFFFFFFFFu;
        ++taken;
        if (tile < p.n_tiles && taken < max_tiles) nxt = dynamic ? gridDim.x + atomicAdd(ctr, 1u) : tile + gridDim.x;
        if (ring_full) mbar_wait(empty0 + 8u * st, ph ^ 1u);
        s_tile[st] = tile;
        const uint32_t full = full0 + 8u * st;
        if (tile >= p.n_tiles) {  // end marker
            mbar_arrive(full);
            break;
        }
        const uint64_t row0 = (uint64_t)tile * p.tile_items;
        const uint32_t rows = tile == last_tile ? last_rows : p.tile_items;
        const uint32_t trows = rows & ~3u;  // TMA needs 16-byte multiples; <= 3 tail rows are read directly
        if (trows) {
            const uint32_t bytes_b = trows * rowbytes;
            const uint32_t bytes_w = p.weight ? trows * 4u : 0u;
            const uint32_t dst = stage0 + st * p.L.stage_stride;
            mbar_arrive_expect_tx(full, bytes_b + bytes_w);
            tma_bulk_g2s(dst, p.bitmap + row0 * p.Wp, bytes_b, full, pol);
            if (bytes_w) tma_bulk_g2s(dst + p.L.off_stage_w, p.weight + row0, bytes_w, full, pol);
        } else {
            mbar_arrive(full);
        }
        if (++st == S) {
            st = 0;
            ph ^= 1u;
            ring_full = true;
        }
        tile = nxt;
    }
}

// PGX_SCAN_TS=1: thread-0 time stamps of a CTA's phases (measurement aid, tools/scan_timeline.py)
__device__ __forceinline__ void ts_mark(const ScanParams &p, uint32_t slot) {
    if (p.dbg_ts) {
        uint64_t t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.dbg_ts[(size_t)blockIdx.x * 8u + slot] = t;
    }
}

// ---- direct epilogue (single GPU): the CTAs add their sums straight into the caller's result vector ------------------
// CTA 0 zeroes the requested words of `out` while the first tiles are in flight and publishes the launch's epoch; every
// CTA's TMA lane, idle once its last copy is issued, waits for that epoch before the CTA-wide barrier that precedes
// the RED flush.  The kernel's end is the end of the pass: no completion ticket, no fence, no snapshot by a last CTA
// (three dependent L2 round trips of the round-1 epilogue, ~2-3 us of every launch).
__device__ __forceinline__ void direct_zero_out(const ScanParams &p, uint32_t tid) {  // CTA 0, all threads, before a __syncthreads()
    const uint32_t G1 = p.G + 1u;
    if (p.flags & kHistCount)
        for (uint32_t i = tid; i < G1; i += kScanThreads) p.out[i] = 0ull;
    if (p.flags & kHistWeight)
        for (uint32_t i = tid; i < G1; i += kScanThreads) p.out[G1 + i] = 0ull;
    for (uint32_t t = 0; t < p.T; ++t)
        for (uint32_t j = tid; j < p.G; j += kScanThreads) p.out[2u * G1 + p.slot[t] * p.G + j] = 0ull;
}
__device__ __forceinline__ void direct_publish(const ScanParams &p) {  // CTA 0, one thread, after that __syncthreads()
    __threadfence();
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.ticket + 1), "r"(p.zero_epoch) : "memory");
}
__device__ __forceinline__ void direct_wait_zeroed(const ScanParams &p) {  // one thread per CTA, before the barrier ahead of the epilogue
    uint32_t v;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ticket + 1) : "memory");
    } while (v != p.zero_epoch);
}

// Per-CTA accumulators -> global u64 accumulators -> (last CTA) the caller's result vector, or the multi-GPU exchange.
// Shared by k_scan and k_scan_priv; expects every thread of the CTA, after a __syncthreads().
__device__ __forceinline__ void scan_epilogue(const ScanParams &p, const SmemAcc &s, const uint32_t tid) {
    // ===== epilogue: per-CTA accumulators -> global u64 accumulators =====
    const uint32_t G1 = p.G + 1u;
    const bool direct = p.zero_epoch != 0u;
    uint64_t *const dst = direct ? p.out : p.acc;
    if (tid == 0) ts_mark(p, 4);
    if (p.flags & kJoint) {  // marginalise the joint histogram into hist[] and the curves' first differences
        const bool use_cnt = (p.flags & kHistCount) || !(p.flags & kWeighted);
        const bool use_w = (p.flags & (kHistWeight | kWeighted)) != 0;
        for (uint32_t c = tid; c < G1; c += kScanThreads) {
            uint32_t n = 0;
            uint64_t w = 0;
            for (uint32_t f = 0; f < p.G; ++f) {
                if (use_cnt) n += s.joint_cnt[c * p.G + f];
                if (use_w) w += ((uint64_t)s.joint_whi[c * p.G + f] << 32) | s.joint_wlo[c * p.G + f];
            }
            if (p.flags & kHistCount) s.hist_cnt[c] = n;
            if (p.flags & kHistWeight) {
                s.hist_wlo[c] = (uint32_t)w;
                s.hist_whi[c] = (uint32_t)(w >> 32);
            }
        }
        for (uint32_t i = tid; i < p.T * p.G; i += kScanThreads) {
            const uint32_t t = i / p.G, f = i - t * p.G;
            uint32_t n = 0;
            uint64_t w = 0;
            for (uint32_t c = p.cov[t]; c < G1; ++c) {  // cov[t] >= 1: the coverage-0 row never counts
                if (p.flags & kWeighted)
                    w += ((uint64_t)s.joint_whi[c * p.G + f] << 32) | s.joint_wlo[c * p.G + f];
                else
                    n += s.joint_cnt[c * p.G + f];
            }
            if (p.flags & kWeighted) {
                s.delta_lo[i] = (uint32_t)w;
                s.delta_hi[i] = (uint32_t)(w >> 32);
            } else {
                s.delta_lo[i] = n;
            }
        }
        __syncthreads();
    }
    for (uint32_t i = tid; i < G1; i += kScanThreads) {
        if (p.flags & kHistCount) {
            const uint32_t c = s.hist_cnt[i];
            if (c) atomicAdd(reinterpret_cast<unsigned long long *>(dst + i), (unsigned long long)c);
        }
        if (p.flags & kHistWeight) {
            const uint64_t v = ((uint64_t)s.hist_whi[i] << 32) | s.hist_wlo[i];
            if (v) atomicAdd(reinterpret_cast<unsigned long long *>(dst + G1 + i), (unsigned long long)v);
        }
    }
    const uint32_t nd = p.T * p.G;
    for (uint32_t t = 0; t < p.T; ++t) {
        // accumulators: launch-local threshold index; the caller's vector: its fused-layout slot
        uint64_t *const row = dst + 2u * G1 + (direct ? p.slot[t] : t) * p.G;
        for (uint32_t j = tid; j < p.G; j += kScanThreads) {
            const uint32_t i = t * p.G + j;
            uint64_t v;
            if (p.flags & kWeighted)
                v = ((uint64_t)s.delta_hi[i] << 32) | s.delta_lo[i];
            else
                v = (uint64_t)(int64_t)(int32_t)s.delta_lo[i];  // signed per-CTA net flip count
            if (v) atomicAdd(reinterpret_cast<unsigned long long *>(row + j), (unsigned long long)v);
        }
    }
    if (direct) {
        if (tid == 0) ts_mark(p, 5);
        return;
    }

    // ===== last CTA: snapshot + re-zero (threadfence reduction pattern) =====
    __shared__ uint32_t s_is_last;
    __syncthreads();
    if (tid == 0) {
        __threadfence();  // (after the barrier: cumulative over the CTA's RED.ADDs, the grid-sync pattern of cooperative groups)
        const unsigned int t = atomicAdd(p.ticket, 1u);
        s_is_last = (t == gridDim.x - 1u) ? 1u : 0u;
    }
    __syncthreads();
    if (s_is_last) {
        __threadfence();
        const uint32_t total = 2u * G1 + nd;
        // accumulator index -> index in the caller's fused layout (0xFFFFFFFF: not requested)
        auto out_index = [&](uint32_t i) -> uint32_t {
            if (i < 2u * G1) {
                const bool wanted = (i < G1) ? (p.flags & kHistCount) : (p.flags & kHistWeight);
                return wanted ? i : 0xFFFFFFFFu;
            }
            const uint32_t d = i - 2u * G1, t = d / p.G, j = d - t * p.G;
            return 2u * G1 + p.slot[t] * p.G + j;
        };
        constexpr uint32_t kBatch = 6;  // independent L2 loads in flight per thread
        const bool xchg = p.x.world > 1u;
        const uint32_t par = p.x.epoch & 1u;
        for (uint32_t base = 0; base < total; base += kBatch * kScanThreads) {
            uint64_t v[kBatch];
#pragma unroll
            for (uint32_t u = 0; u < kBatch; ++u) {
                const uint32_t i = base + u * kScanThreads + tid;
                v[u] = i < total ? __ldcg(p.acc + i) : 0ull;
            }
#pragma unroll
            for (uint32_t u = 0; u < kBatch; ++u) {
                const uint32_t i = base + u * kScanThreads + tid;
                if (i >= total) continue;
                const uint32_t oi = out_index(i);
                if (oi == 0xFFFFFFFFu) continue;
                if (!xchg) {
                    p.out[oi] = v[u];
                } else {
                    // push my partial result into slot [par][rank] of every rank (NVLink peer stores).  Each u64 travels
                    // as two 8-byte packets {epoch : 32 | half : 32}: an aligned 8-byte store is atomic, so the
                    // receiver needs no separate flag, fence or barrier (the idea of NCCL's LL protocol)
                    // slots are indexed by the launch-local accumulator index i (< acc_words = the slot stride for any
                    // number of thresholds in the call), not by the caller's fused-layout index oi
                    const size_t off = ((size_t)(par * kMaxRanks + p.x.rank) * p.x.stride + i) * 2u;
                    const uint64_t tag = (uint64_t)p.x.epoch << 32;
                    const uint64_t lo = tag | (v[u] & 0xFFFFFFFFull), hi = tag | (v[u] >> 32);
                    for (uint32_t r = 0; r < p.x.world; ++r) {
                        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p.x.data[r] + off), "l"(lo) : "memory");
                        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p.x.data[r] + off + 1), "l"(hi) : "memory");
                    }
                }
                if (v[u]) p.acc[i] = 0ull;
            }
        }
        if (xchg) {
            // every thread collects the words it pushed itself: spin until all ranks' packets carry this epoch
            const uint64_t *mine = p.x.data[p.x.rank] + (size_t)par * kMaxRanks * p.x.stride * 2u;
            uint64_t t0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            bool dead = false;
            for (uint32_t i = tid; i < total; i += kScanThreads) {
                const uint32_t oi = out_index(i);
                if (oi == 0xFFFFFFFFu) continue;
                // all ranks' packets of this word are loaded together (one L2 round trip per attempt), then checked
                uint64_t a[kMaxRanks], b[kMaxRanks];
                for (;;) {
#pragma unroll
                    for (uint32_t r = 0; r < (uint32_t)kMaxRanks; ++r) {
                        a[r] = b[r] = (uint64_t)p.x.epoch << 32;
                        if (r < p.x.world) {
                            const uint64_t *q = mine + ((size_t)r * p.x.stride + i) * 2u;
                            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a[r]), "=l"(b[r]) : "l"(q) : "memory");
                        }
                    }
                    bool ok = true;
#pragma unroll
                    for (uint32_t r = 0; r < (uint32_t)kMaxRanks; ++r)
                        ok = ok && (uint32_t)(a[r] >> 32) == p.x.epoch && (uint32_t)(b[r] >> 32) == p.x.epoch;
                    if (ok || dead) break;
                    uint64_t t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > 10000000000ull) {  // 10 s: a peer never arrived; flag the error, do not hang
                        atomicExch(p.x.err, 2u);
                        dead = true;
                    }
                }
                uint64_t sum = 0;
#pragma unroll
                for (uint32_t r = 0; r < (uint32_t)kMaxRanks; ++r)
                    if (r < p.x.world) sum += (a[r] & 0xFFFFFFFFull) | (b[r] << 32);
                p.out[oi] = sum;
            }
        }
        if (tid == 0) *p.ticket = 0u;
    }
}

template <bool QUORUM, int C_T, bool MASK>
__global__ void __launch_bounds__(kScanThreads, 2) k_scan(const __grid_constant__ ScanParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t tid = threadIdx.x;
    if (tid == 0) ts_mark(p, 0);
    const uint32_t warp = tid >> 5, lane = tid & 31u;
    const uint32_t S = p.stages;
    const uint32_t full0 = smem_u32(smem), empty0 = full0 + 8u * kMaxStages;
    const uint32_t stage0 = smem_u32(smem + p.L.off_stage0);
    __shared__ uint32_t s_tile_words[kMaxStages];  // tile held by each stage (written by the producer lane)
    volatile uint32_t *s_tile = s_tile_words;
    const uint32_t rowbytes = p.Wp * 8u;

    SmemAcc s;
    s.hist_cnt = reinterpret_cast<uint32_t *>(smem + p.L.off_hist_cnt);
    s.hist_wlo = reinterpret_cast<uint32_t *>(smem + p.L.off_hist_wlo);
    s.hist_whi = reinterpret_cast<uint32_t *>(smem + p.L.off_hist_whi);
    s.delta_lo = reinterpret_cast<uint32_t *>(smem + p.L.off_delta_lo);
    s.delta_hi = reinterpret_cast<uint32_t *>(smem + p.L.off_delta_hi);
    uint32_t *s_thr = reinterpret_cast<uint32_t *>(smem + p.L.off_thr);
    s.thr = s_thr;
    s.joint_cnt = reinterpret_cast<uint32_t *>(smem + p.L.off_joint_cnt);
    s.joint_wlo = reinterpret_cast<uint32_t *>(smem + p.L.off_joint_wlo);
    s.joint_whi = reinterpret_cast<uint32_t *>(smem + p.L.off_joint_whi);

    {
        uint32_t *acc32 = reinterpret_cast<uint32_t *>(smem + p.L.off_acc);
        for (uint32_t i = tid; i < p.L.acc_words; i += kScanThreads) acc32[i] = 0u;
        if (QUORUM)
            for (uint32_t i = tid; i < p.T * p.G; i += kScanThreads) s_thr[i] = p.thr[i];
    }
    if (tid == 0) {
        for (uint32_t i = 0; i < S; ++i) {
            mbar_init(full0 + 8u * i, 1u);
            mbar_init(empty0 + 8u * i, (uint32_t)kConsumerWarps);
        }
        mbar_fence_init();
    }
    const bool zeroes_out = p.zero_epoch != 0u && blockIdx.x == 0u;
    if (zeroes_out) direct_zero_out(p, tid);
    __syncthreads();
    if (zeroes_out && tid == kScanThreads - 1) direct_publish(p);
    if (tid == 0) ts_mark(p, 1);

    const uint32_t last_tile = p.n_tiles - 1u;
    // rows of the last tile (may be partial); every other tile is full
    const uint32_t last_rows = (uint32_t)(p.n_rows - (uint64_t)last_tile * p.tile_items);

    if (warp == (uint32_t)kConsumerWarps) {
        // ===== TMA producer =====
        if (lane == 0) {
            scan_producer(p, full0, empty0, stage0, s_tile, p.L.vert_planes ? ((1u << p.L.vert_planes) - 1u) : 0xFFFFFFFFu);
            if (p.zero_epoch) direct_wait_zeroed(p);
        }
    } else {
        // ===== consumers: one thread per item =====
        const uint32_t C_rt = p.Wp >> 1;
        uint32_t a = 0;
        while (a < 3 && C_rt && !((C_rt >> a) & 1u)) ++a;  // a' = min(ctz(C), 3)
        const uint32_t rot_shift = 3u - a, rot_mask = (1u << a) - 1u;
        uint32_t st = 0, ph = 0;
        for (bool first_wait = true;; first_wait = false) {
            mbar_wait(full0 + 8u * st, ph);
            const uint32_t tile = s_tile[st];
            if (tile >= p.n_tiles) break;
            if (tid == 0 && first_wait) ts_mark(p, 2);
            const uint64_t row0 = (uint64_t)tile * p.tile_items;
            const uint32_t rows = tile == last_tile ? last_rows : p.tile_items;
            const uint32_t trows = rows & ~3u;
            const uint32_t base = stage0 + st * p.L.stage_stride;
            const uint32_t wbase = base + p.L.off_stage_w;
            for (uint32_t li = tid; li < rows; li += kConsumerThreads) {
                const uint64_t item = row0 + li;
                if (item == 0) {  // the reference's dummy item (abacus.rs:551, 1000-1002)
                    if (p.countable) p.countable[0] = 0xFFFFFFFFu;
                    continue;
                }
                if (li < trows) {
                    uint32_t wgt = 1u;
                    if (p.weight) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wgt) : "r"(wbase + li * 4u));
                    const uint32_t row_addr = base + li * rowbytes;
                    if (QUORUM) {
                        item_quorum_natural(p, s, item, wgt, SmemRowReader{row_addr});
                    } else if (C_T < 0) {  // 8-byte rows (G <= 64)
                        const uint64_t x = lds_u64(row_addr) & p.last_mask0;
                        account_fast(p, s, item, __popcll(x), x ? first_bit(x) : 0xFFFFFFFFu, wgt);
                    } else {
                        item_fast_smem<C_T, MASK>(p, s, item, wgt, row_addr, li, C_rt, rot_shift, rot_mask);
                    }
                } else {  // <= 3 tail rows of the last tile, straight from global memory
                    const uint32_t wgt = p.weight ? __ldg(p.weight + item) : 1u;
                    GlobalRowReader rd{p.bitmap + item * p.Wp};
                    if (QUORUM)
                        item_quorum_natural(p, s, item, wgt, rd);
                    else
                        item_fast_natural(p, s, item, wgt, rd);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8u * st);
            if (++st == S) {
                st = 0;
                ph ^= 1u;
            }
        }
    }
    if (tid == 0) ts_mark(p, 3);
    __syncthreads();

    scan_epilogue(p, s, tid);
}

// ---- lane-private counters (kPrivate) ---------------------------------------------------------------
// Shared atomics cost ~2 cycles per lane on this part (and serialise on equal addresses, which a U-shaped coverage
// histogram produces all the time), plain shared loads / stores 1/32: below ~300 groups the atomics, not HBM, bound
// k_scan.  Here every consumer thread owns one narrow counter per bin: row `bin` of the private region holds the 256
// threads' counters (u8 for counts, u16 for bp sums), laid out so that the 32 lanes of a warp always hit 32 different
// banks whatever bins they address.  An item costs two plain read-modify-writes (its coverage bin; its (coverage class,
// first group) bin -- one bin for ANY number of q = 0 thresholds: a class is "how many of the distinct coverage cutoffs
// the item reaches").  A counter that wraps (once per 256 items of a thread and bin; for bp sums whenever 64 Ki bp have
// piled up, plus the bits of a weight above 2^16) carries into a per-CTA u32 word with a shared atomic -- rare.  After
// the last tile each bin is folded by one thread (128-bit loads, DP4A byte sums): low parts + (carries << 8 or 16).
template <int CW>
__device__ __forceinline__ uint32_t priv_thread_off(uint32_t tid) {
    const uint32_t w = tid >> 5, l = tid & 31u;
    constexpr uint32_t kPerWord = 4u / CW;  // threads sharing one 32-bit word: warps w, w + 1, ... of the same 128-byte segment
    return (w / kPerWord) * 128u + l * 4u + (w % kPerWord) * CW;
}

template <int CW>
__device__ __forceinline__ uint32_t priv_ld(uint32_t addr) {
    uint32_t v;
    if (CW == 1)
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    else
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <int CW>
__device__ __forceinline__ void priv_st(uint32_t addr, uint32_t v) {  // stores the low CW bytes
    if (CW == 1)
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
    else
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// fold the lane-private rows (+ carries) into the CTA accumulators: every thread of the CTA, after a __syncthreads()
template <int CW, bool WEIGHTED>
__device__ __forceinline__ void priv_fold(const ScanParams &p, const SmemAcc &s, uint32_t *cls_lo, uint32_t *cls_hi,
                                          const uint32_t *carry, uint32_t priv_base, uint32_t tid) {
    constexpr uint32_t kRow = 256u * CW, kChunks = kRow / 16u, kBits = 8u * CW;
    for (uint32_t bin = tid; bin < p.L.priv_bins; bin += kScanThreads) {
        const uint32_t row = priv_base + bin * kRow;
        uint32_t sum32 = 0;  // 256 x (2^16 - 1) < 2^24
#pragma unroll 4
        for (uint32_t k = 0; k < kChunks; ++k) {
            const uint32_t a = row + ((k + bin) & (kChunks - 1u)) * 16u;  // rotated: the lanes of a quarter-warp hit different banks
            uint32_t x0, x1, x2, x3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(a));
            if (CW == 1) {
                sum32 = __dp4a(x0, 0x01010101u, sum32);
                sum32 = __dp4a(x1, 0x01010101u, sum32);
                sum32 = __dp4a(x2, 0x01010101u, sum32);
                sum32 = __dp4a(x3, 0x01010101u, sum32);
            } else {
                sum32 += (x0 & 0xFFFFu) + (x0 >> 16) + (x1 & 0xFFFFu) + (x1 >> 16) + (x2 & 0xFFFFu) + (x2 >> 16) + (x3 & 0xFFFFu) + (x3 >> 16);
            }
        }
        const uint64_t sum = (uint64_t)sum32 + ((uint64_t)carry[bin] << kBits);
        uint32_t *lo, *hi;
        uint32_t idx;
        if (bin < p.L.priv_hist_bins) {
            lo = WEIGHTED ? s.hist_wlo : s.hist_cnt;
            hi = s.hist_whi;
            idx = bin;
        } else {
            lo = cls_lo;
            hi = cls_hi;
            idx = bin - p.L.priv_hist_bins;
        }
        lo[idx] = (uint32_t)sum;
        if (WEIGHTED) hi[idx] = (uint32_t)(sum >> 32);
    }
}

template <int CW, bool WEIGHTED, int C_T>
__global__ void __launch_bounds__(kScanThreads, 2) k_scan_priv(const __grid_constant__ ScanParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t tid = threadIdx.x;
    if (tid == 0) ts_mark(p, 0);
    const uint32_t warp = tid >> 5, lane = tid & 31u;
    const uint32_t S = p.stages;
    const uint32_t full0 = smem_u32(smem), empty0 = full0 + 8u * kMaxStages;
    const uint32_t stage0 = smem_u32(smem + p.L.off_stage0);
    __shared__ uint32_t s_tile_words[kMaxStages];  // tile held by each stage (written by the producer lane)
    volatile uint32_t *s_tile = s_tile_words;
    const uint32_t rowbytes = p.Wp * 8u;

    SmemAcc s;
    s.hist_cnt = reinterpret_cast<uint32_t *>(smem + p.L.off_hist_cnt);
    s.hist_wlo = reinterpret_cast<uint32_t *>(smem + p.L.off_hist_wlo);
    s.hist_whi = reinterpret_cast<uint32_t *>(smem + p.L.off_hist_whi);
    s.delta_lo = reinterpret_cast<uint32_t *>(smem + p.L.off_delta_lo);
    s.delta_hi = reinterpret_cast<uint32_t *>(smem + p.L.off_delta_hi);
    s.thr = nullptr;
    s.joint_cnt = s.joint_wlo = s.joint_whi = nullptr;
    uint32_t *cls_lo = reinterpret_cast<uint32_t *>(smem + p.L.off_cls_lo);
    uint32_t *cls_hi = reinterpret_cast<uint32_t *>(smem + p.L.off_cls_hi);
    uint32_t *carry = reinterpret_cast<uint32_t *>(smem + p.L.off_carry);
    uint32_t *s_cbase = reinterpret_cast<uint32_t *>(smem + p.L.off_cbase);  // coverage -> first bin of its class (or ~0)
    const uint32_t priv_base = smem_u32(smem + p.L.off_priv);

    {  // accumulators and the private region (contiguous) start at zero
        uint4 *z = reinterpret_cast<uint4 *>(smem + p.L.off_acc);
        const uint32_t n16 = (p.L.off_cbase - p.L.off_acc + 15u) / 16u;
        for (uint32_t i = tid; i < n16; i += kScanThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        for (uint32_t c = tid; c <= p.G; c += kScanThreads) {
            uint32_t cls = 0;
            for (uint32_t d = 0; d < p.n_classes; ++d) cls += c >= p.cls_thr[d] ? 1u : 0u;
            s_cbase[c] = (cls && c) ? p.L.priv_hist_bins + (cls - 1u) * p.G : 0xFFFFFFFFu;
        }
    }
    if (tid == 0) {
        for (uint32_t i = 0; i < S; ++i) {
            mbar_init(full0 + 8u * i, 1u);
            mbar_init(empty0 + 8u * i, (uint32_t)kConsumerWarps);
        }
        mbar_fence_init();
    }
    const bool zeroes_out = p.zero_epoch != 0u && blockIdx.x == 0u;
    if (zeroes_out) direct_zero_out(p, tid);
    __syncthreads();
    if (zeroes_out && tid == kScanThreads - 1) direct_publish(p);
    if (tid == 0) ts_mark(p, 1);

    const uint32_t last_tile = p.n_tiles - 1u;
    const uint32_t last_rows = (uint32_t)(p.n_rows - (uint64_t)last_tile * p.tile_items);

    if (warp == (uint32_t)kConsumerWarps) {
        // ===== TMA producer (same ring as k_scan) =====
        if (lane == 0) {
            scan_producer(p, full0, empty0, stage0, s_tile, p.L.vert_planes ? ((1u << p.L.vert_planes) - 1u) : 0xFFFFFFFFu);
            if (p.zero_epoch) direct_wait_zeroed(p);
        }
    } else {
        // ===== consumers: two items per thread and step, lane-private counters =====
        const uint32_t C_rt = p.Wp >> 1;
        uint32_t a = 0;
        while (a < 3 && C_rt && !((C_rt >> a) & 1u)) ++a;
        const uint32_t rot_shift = 3u - a, rot_mask = (1u << a) - 1u;
        const uint32_t my = priv_base + priv_thread_off<CW>(tid);
        uint32_t st = 0, ph = 0;
        constexpr uint32_t kRow = 256u * CW;
        constexpr uint32_t kBits = 8u * CW, kMask = (1u << kBits) - 1u;
        const bool has_hist = p.L.priv_hist_bins != 0u, has_cls = p.n_classes != 0u;
        // One item's two read-modify-writes on the thread's own counters.  The two loads are in flight together (disjoint
        // regions); the narrow counters keep the low bits, and whatever wraps out of them (rarely) is added to the bin's
        // u32 carry word of the CTA -- one branch for both.  cb = s_cbase[cov]: first bin of the item's coverage class.
        auto account = [&](bool valid, uint32_t cov, uint32_t first, uint32_t wgt, uint32_t cb) {
            const uint32_t inc = valid ? (WEIGHTED ? wgt : 1u) : 0u;
            const bool counted = cb != 0xFFFFFFFFu;
            const uint32_t cbin = counted ? cb + first : 0u;
            const uint32_t ha = my + cov * kRow, ca = my + cbin * kRow;
            const uint32_t cinc = counted ? inc : 0u;
            uint32_t hv = 0, cv = 0;
            if (has_hist) hv = priv_ld<CW>(ha);
            if (has_cls) cv = priv_ld<CW>(ca);
            const uint32_t nh = hv + (WEIGHTED ? (inc & kMask) : inc), nc = cv + (WEIGHTED ? (cinc & kMask) : cinc);
            if (has_hist) priv_st<CW>(ha, nh);
            if (has_cls && counted) priv_st<CW>(ca, nc);  // (an uncounted item's class address is a dummy: row 0)
            if (((nh | nc) >> kBits) | (WEIGHTED ? (inc >> kBits) : 0u)) {
                const uint32_t ch = (nh >> kBits) + (WEIGHTED ? (inc >> kBits) : 0u), cc = (nc >> kBits) + (WEIGHTED ? (cinc >> kBits) : 0u);
                if (has_hist && ch) atomicAdd(carry + cov, ch);
                if (has_cls && cc) atomicAdd(carry + cbin, cc);
            }
        };
        // kPrivGrowthAtomics: the curves' first differences keep their shared atomics (as in k_scan)
        const bool growth_atomics = (p.flags & kPrivGrowthAtomics) != 0u;
        auto growth_atomic = [&](uint32_t cov, uint32_t first, uint32_t wgt) {
            for (uint32_t t = 0; t < p.T; ++t)
                if (cov >= p.cov[t]) {
                    if (WEIGHTED)
                        smem_add64(s.delta_lo + t * p.G, s.delta_hi + t * p.G, first, wgt, 0u);
                    else
                        atomicAdd(&s.delta_lo[t * p.G + first], 1u);
                }
        };
        // items per thread and step: their row loads, popcounts and class look-ups are independent and overlap; only the
        // counter updates run one item after the other (two items of a thread may share a bin)
        constexpr int K = (C_T < 0 || C_T == 1) ? 4 : 2;
        for (bool first_wait = true;; first_wait = false) {
            mbar_wait(full0 + 8u * st, ph);
            const uint32_t tile = s_tile[st];
            if (tile >= p.n_tiles) break;
            if (tid == 0 && first_wait) ts_mark(p, 2);
            const uint64_t row0 = (uint64_t)tile * p.tile_items;
            const uint32_t rows = tile == last_tile ? last_rows : p.tile_items;
            const uint32_t trows = rows & ~3u;
            const uint32_t base = stage0 + st * p.L.stage_stride;
            const uint32_t wbase = base + p.L.off_stage_w;
            for (uint32_t li0 = tid; li0 < trows; li0 += K * kConsumerThreads) {
                uint32_t cov[K], first[K], wgt[K], cb[K];
                bool valid[K];
                uint64_t x[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const uint32_t lk = li0 + (uint32_t)k * kConsumerThreads;
                    valid[k] = lk < trows && (tile | lk) != 0u;  // item 0: the reference's dummy item, never counted
                    const uint32_t li = lk < trows ? lk : li0;     // out-of-tile slots re-read a valid row and add nothing
                    wgt[k] = 1u;
                    if (WEIGHTED) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wgt[k]) : "r"(wbase + li * 4u));
                    if (C_T < 0) x[k] = lds_u64(base + li * rowbytes);
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const uint32_t lk = li0 + (uint32_t)k * kConsumerThreads;
                    const uint32_t li = lk < trows ? lk : li0;
                    if (C_T < 0) {  // 8-byte rows (G <= 64): 2 POPC for the coverage, 1 for the first group (trailing zeros)
                        const uint64_t xm = x[k] & p.last_mask0;
                        const uint32_t lo = (uint32_t)xm, hi = (uint32_t)(xm >> 32);
                        cov[k] = __popc(lo) + __popc(hi);
                        const uint32_t sel = lo ? lo : hi;
                        first[k] = (lo ? 0u : 32u) + __popc((sel - 1u) & ~sel);
                    } else {
                        row_cov_first<C_T, true>(p, base + li * rowbytes, li, C_rt, rot_shift, rot_mask, cov[k], first[k]);
                    }
                    cb[k] = has_cls ? s_cbase[cov[k]] : 0xFFFFFFFFu;
                    if (p.countable && lk < trows) p.countable[row0 + lk] = (tile | lk) != 0u ? cov[k] : 0xFFFFFFFFu;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) account(valid[k], cov[k], first[k], wgt[k], cb[k]);
                if (growth_atomics) {
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        if (valid[k] && cov[k]) growth_atomic(cov[k], first[k], wgt[k]);
                }
            }
            if (tid < rows - trows) {  // <= 3 tail rows of the last tile, straight from global memory
                const uint64_t item = row0 + trows + tid;
                const uint32_t wgt = WEIGHTED ? __ldg(p.weight + item) : 1u;
                uint32_t cov = 0, first = 0xFFFFFFFFu;
                for (uint32_t w = 0; w < p.W; ++w) {
                    const uint64_t x = __ldg(p.bitmap + item * p.Wp + w) & word_mask(p, w);
                    cov += __popcll(x);
                    if (x && first == 0xFFFFFFFFu) first = w * 64u + first_bit(x);
                }
                if (p.countable) p.countable[item] = item ? cov : 0xFFFFFFFFu;
                account(item != 0, cov, first, wgt, has_cls ? s_cbase[cov] : 0xFFFFFFFFu);
                if (growth_atomics && item != 0 && cov) growth_atomic(cov, first, wgt);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8u * st);
            if (++st == S) {
                st = 0;
                ph ^= 1u;
            }
        }
    }
    if (tid == 0) ts_mark(p, 3);
    __syncthreads();
    priv_fold<CW, WEIGHTED>(p, s, cls_lo, cls_hi, carry, priv_base, tid);
    __syncthreads();
    // classes -> first differences of every threshold's curve: threshold t counts the classes >= cls_rank[t]
    // (kPrivGrowthAtomics: the first differences are in place already)
    for (uint32_t i = tid; i < ((p.flags & kPrivGrowthAtomics) ? 0u : p.T * p.G); i += kScanThreads) {
        const uint32_t t = i / p.G, f = i - t * p.G;
        uint64_t v = 0;
        for (uint32_t k = p.cls_rank[t] - 1u; k < p.n_classes; ++k)
            v += WEIGHTED ? (((uint64_t)cls_hi[k * p.G + f] << 32) | cls_lo[k * p.G + f]) : (uint64_t)cls_lo[k * p.G + f];
        s.delta_lo[i] = (uint32_t)v;
        if (WEIGHTED) s.delta_hi[i] = (uint32_t)(v >> 32);
    }
    __syncthreads();
    scan_epilogue(p, s, tid);
}

// ---- bit-sliced vertical counters (kVertical): G <= 128, counting nodes / edges --------------------------------------
// An item's row is one 64-bit word x.  Its two facts are turned into ONE-HOT words -- 1 << popc(x) (coverage) and x & -x
// (first group) -- so a histogram is the column sum of a 64-column bit matrix.  Column sums of words are what carry-save
// adders do: a thread adds 16 one-hot words with 15 full adders (xor3 / majority = 2 LOP3 per 32-bit half) into the four
// bit-planes ones / twos / fours / eights it keeps in registers; the "sixteens" word that falls out ripples into P more
// planes in shared memory (the thread's own words: no conflicts, no atomics), P chosen by the host from the tiles a CTA
// walks.  ~22 instructions per item for a histogram plus one curve (the lane-private kernel: ~45, the atomics kernel is
// bound by ATOMS throughput).  After the last tile the threads' counters are added with the same full adders across the
// warp (shuffles), then across the warps through shared memory, and only then read out bin by bin.
// One counter for the histogram (HIST), one per distinct coverage cutoff (D <= 3: items with coverage >= cutoff, by
// first group); C1: the lowest cutoff is 1 (any covered item counts: no mask needed).  Coverage 64 (G = 64) does not fit
// the one-hot word: those items are counted in a register.
struct V64 {
    uint32_t lo, hi;
};
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// full adder on 64 one-bit columns: (h, l) = a + b + c; l may alias a
__device__ __forceinline__ void csa(V64 &h, V64 &l, const V64 a, const V64 b, const V64 c) {
    const V64 hh = {lop3_maj(a.lo, b.lo, c.lo), lop3_maj(a.hi, b.hi, c.hi)};
    l = {lop3_xor3(a.lo, b.lo, c.lo), lop3_xor3(a.hi, b.hi, c.hi)};
    h = hh;
}
struct VertTree {  // ones, twos, fours, eights of one counter + the pending fours / eights words of the running 16-block
    V64 p1 = {0, 0}, p2 = {0, 0}, p4 = {0, 0}, p8 = {0, 0}, q4 = {0, 0}, q8 = {0, 0};
};
constexpr int kVertK = 16;          // items per thread and step = inputs of one adder tree
constexpr int kVertMaxPlanes = 12;  // shared-memory planes per counter (sixteens .. ): 16 * 2^12 items per thread
constexpr int kVertFoldPlanes = 4 + kVertMaxPlanes + 8;  // a CTA's total: 256 threads = 8 more bits

// four more words of the running block (quarter Q of NQ): a block is 16 words (NQ = 4; its last quarter returns the
// "sixteens" word) or 8 words (NQ = 2, two-word rows: the last quarter returns the "eights" word and p8 stays unused)
template <int Q, int NQ>
__device__ __forceinline__ void vert_feed4(VertTree &t, const V64 v0, const V64 v1, const V64 v2, const V64 v3, V64 &out) {
    V64 a2, b2, f4, e8;
    csa(a2, t.p1, t.p1, v0, v1);
    csa(b2, t.p1, t.p1, v2, v3);
    csa(f4, t.p2, t.p2, a2, b2);
    if (Q % 2 == 0) {
        t.q4 = f4;
    } else {
        csa(e8, t.p4, t.p4, t.q4, f4);
        if (NQ == 2)
            out = e8;
        else if (Q == 1)
            t.q8 = e8;
        else
            csa(out, t.p8, t.p8, t.q8, e8);
    }
}
// ripple the sixteens word into the thread's shared-memory planes (plane stride: 256 threads x 8 bytes)
__device__ __forceinline__ void vert_ripple(uint32_t addr, uint32_t n_planes, V64 e) {
    for (uint32_t k = 0; k < n_planes && (e.lo | e.hi); ++k, addr += 256u * 8u) {
        uint32_t vlo, vhi;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(vlo), "=r"(vhi) : "r"(addr));
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(vlo ^ e.lo), "r"(vhi ^ e.hi) : "memory");
        e.lo &= vlo;
        e.hi &= vhi;
    }
}

// One step of a consumer thread: K = 16 / NW rows of NW 64-bit words of the stage at `base` (rows li0 + k * 256) into the
// trees; out[c * NW + w] = the word that falls out of the tree of counter c, row word w.  SLOW: per-slot validity (item 0,
// rows past the end of a short tile) and the per-item coverage output; every other tile takes the branch-free path.
template <int NW, int HIST, int D, bool C1, bool SLOW>
__device__ __forceinline__ void vert_step(const ScanParams &p, VertTree (&tree)[(HIST + D) * NW], V64 (&out)[(HIST + D) * NW],
                                          uint32_t base, uint32_t li0, uint32_t trows, uint32_t tile, uint64_t row0,
                                          const uint32_t *cthr) {
    constexpr int K = kVertK / NW, NQ = K / 4;
    uint64_t x[K][NW];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t lk = li0 + (uint32_t)k * kConsumerThreads;
        const uint32_t addr = base + ((SLOW && lk >= trows) ? li0 : lk) * (8u * NW);
        if (NW == 1)
            x[k][0] = lds_u64(addr);
        else
            lds_v2_u64(addr, x[k][0], x[k][NW - 1]);
    }
    auto quarter = [&](auto qtag) {
        constexpr int Q = decltype(qtag)::value;
        V64 oh[NW][4], fw[D > 0 ? D : 1][NW][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = Q * 4 + j;
            const uint32_t lk = li0 + (uint32_t)k * kConsumerThreads;
            uint64_t xm[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) xm[w] = x[k][w];
            xm[NW - 1] &= (NW == 1) ? p.last_mask0 : p.last_mask1;  // bits >= G of the last word are ignored
            bool valid = true;
            if (SLOW) {
                valid = lk < trows && (tile | lk) != 0u;  // item 0: the reference's dummy item, never counted
#pragma unroll
                for (int w = 0; w < NW; ++w) xm[w] = valid ? xm[w] : 0ull;
            }
            uint32_t cov = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) cov += __popc((uint32_t)xm[w]) + __popc((uint32_t)(xm[w] >> 32));
            if (SLOW && p.countable && lk < trows) p.countable[row0 + lk] = valid ? cov : 0xFFFFFFFFu;
            if (HIST) {
                // 1 << cov in 32-bit pieces (shl.b32 clamps: shifts >= 32, also "negative" ones, give 0); the top coverage
                // 64 NW has no bit and is recovered in the read-out
                const uint32_t one = SLOW ? (valid ? 1u : 0u) : 1u;
                uint32_t h[2 * NW];
#pragma unroll
                for (int i = 0; i < 2 * NW; ++i) asm("shl.b32 %0, %1, %2;" : "=r"(h[i]) : "r"(one), "r"(cov - 32u * (uint32_t)i));
#pragma unroll
                for (int w = 0; w < NW; ++w) oh[w][j] = {h[2 * w], h[2 * w + 1]};
            }
            if (D > 0) {
                uint64_t f[NW];
                f[0] = xm[0] & (0ull - xm[0]);
                if (NW == 2) f[NW - 1] = xm[0] ? 0ull : (xm[NW - 1] & (0ull - xm[NW - 1]));
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const uint32_t m = (C1 && d == 0) ? 0xFFFFFFFFu : (cov >= cthr[d] ? 0xFFFFFFFFu : 0u);
#pragma unroll
                    for (int w = 0; w < NW; ++w) fw[d][w][j] = {(uint32_t)f[w] & m, (uint32_t)(f[w] >> 32) & m};
                }
            }
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            if (HIST) vert_feed4<Q, NQ>(tree[w], oh[w][0], oh[w][1], oh[w][2], oh[w][3], out[w]);
#pragma unroll
            for (int d = 0; d < D; ++d)
                vert_feed4<Q, NQ>(tree[(HIST + d) * NW + w], fw[d][w][0], fw[d][w][1], fw[d][w][2], fw[d][w][3], out[(HIST + d) * NW + w]);
        }
    };
    quarter(std::integral_constant<int, 0>{});
    quarter(std::integral_constant<int, 1>{});
    if (NQ == 4) {
        quarter(std::integral_constant<int, NQ == 4 ? 2 : 0>{});
        quarter(std::integral_constant<int, NQ == 4 ? 3 : 1>{});
    }
}

template <int NW, int HIST, int D, bool C1>
__global__ void __launch_bounds__(kScanThreads, 2) k_scan_vert(const __grid_constant__ ScanParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int C = HIST + D, CW = C * NW;      // counters, counter words (one 64-bin word per row word)
    constexpr int K = kVertK / NW;                // rows per thread and step
    constexpr uint32_t TREE = NW == 1 ? 4u : 3u;  // register planes of a tree: 1, 2, 4 (, 8); the shared-memory planes follow
    const uint32_t tid = threadIdx.x;
    if (tid == 0) ts_mark(p, 0);
    const uint32_t warp = tid >> 5, lane = tid & 31u;
    const uint32_t S = p.stages;
    const uint32_t full0 = smem_u32(smem), empty0 = full0 + 8u * kMaxStages;
    const uint32_t stage0 = smem_u32(smem + p.L.off_stage0);
    __shared__ uint32_t s_tile_words[kMaxStages];  // tile held by each stage (written by the producer lane)
    volatile uint32_t *s_tile = s_tile_words;
    const uint32_t P = p.L.vert_planes;

    SmemAcc s;
    s.hist_cnt = reinterpret_cast<uint32_t *>(smem + p.L.off_hist_cnt);
    s.hist_wlo = s.hist_whi = nullptr;
    s.delta_lo = reinterpret_cast<uint32_t *>(smem + p.L.off_delta_lo);
    s.delta_hi = nullptr;
    s.thr = nullptr;
    s.joint_cnt = s.joint_wlo = s.joint_whi = nullptr;
    uint32_t *cls_tot = reinterpret_cast<uint32_t *>(smem + p.L.off_cls_lo);  // [D][64 NW]
    uint64_t *fold = reinterpret_cast<uint64_t *>(smem + p.L.off_carry);      // [CW][kVertFoldPlanes]: the CTA's totals
    const uint32_t planes0 = smem_u32(smem + p.L.off_priv);                   // [CW][TREE + P][256] u64
    __shared__ uint32_t s_missing_word;
    uint32_t *s_missing = &s_missing_word;

    {
        uint4 *z = reinterpret_cast<uint4 *>(smem + p.L.off_acc);
        const uint32_t n16 = (p.L.vert_end - p.L.off_acc) / 16u;
        for (uint32_t i = tid; i < n16; i += kScanThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        for (uint32_t i = 0; i < S; ++i) {
            mbar_init(full0 + 8u * i, 1u);
            mbar_init(empty0 + 8u * i, (uint32_t)kConsumerWarps);
        }
        mbar_fence_init();
    }
    const bool zeroes_out = p.zero_epoch != 0u && blockIdx.x == 0u;
    if (zeroes_out) direct_zero_out(p, tid);
    __syncthreads();
    if (zeroes_out && tid == kScanThreads - 1) direct_publish(p);
    if (tid == 0) ts_mark(p, 1);

    const uint32_t last_tile = p.n_tiles - 1u;
    const uint32_t last_rows = (uint32_t)(p.n_rows - (uint64_t)last_tile * p.tile_items);

    if (warp == (uint32_t)kConsumerWarps) {
        // ===== TMA producer (same ring as k_scan; rows only: counting never stages the weights) =====
        if (lane == 0) {
            scan_producer(p, full0, empty0, stage0, s_tile, p.L.vert_planes ? ((1u << p.L.vert_planes) - 1u) : 0xFFFFFFFFu);
            if (p.zero_epoch) direct_wait_zeroed(p);
        }
    } else {
        // ===== consumers: K rows per thread and step =====
        VertTree tree[CW];
        uint32_t seen = 0;  // (thread 0) items fed to the trees: the histogram's bins must add up to it (top coverage, see below)
        const uint32_t my_planes = planes0 + tid * 8u;
        uint32_t cthr[D > 0 ? D : 1];
#pragma unroll
        for (int d = 0; d < D; ++d) cthr[d] = p.cls_thr[d];
        uint32_t st = 0, ph = 0;
        for (bool first_wait = true;; first_wait = false) {
            mbar_wait(full0 + 8u * st, ph);
            const uint32_t tile = s_tile[st];
            if (tile >= p.n_tiles) break;
            if (tid == 0 && first_wait) ts_mark(p, 2);
            const uint64_t row0 = (uint64_t)tile * p.tile_items;
            const uint32_t rows = tile == last_tile ? last_rows : p.tile_items;
            const uint32_t trows = rows & ~3u;
            const uint32_t base = stage0 + st * p.L.stage_stride;
            const bool slow = tile == 0u || trows != p.tile_items || p.countable != nullptr;  // item 0 / a short tile / coverage output
            for (uint32_t li0 = tid; li0 < trows; li0 += (uint32_t)K * kConsumerThreads) {
                V64 out[CW];
                if (slow)
                    vert_step<NW, HIST, D, C1, true>(p, tree, out, base, li0, trows, tile, row0, cthr);
                else
                    vert_step<NW, HIST, D, C1, false>(p, tree, out, base, li0, trows, tile, row0, cthr);
#pragma unroll
                for (int cw = 0; cw < CW; ++cw) vert_ripple(my_planes + ((uint32_t)cw * (P + TREE) + TREE) * 2048u, P, out[cw]);
            }
            if (tid == 0) seen += trows - (tile == 0u && trows ? 1u : 0u);  // items this CTA fed to the trees
            if (tid < rows - trows) {  // <= 3 tail rows of the last tile, straight from global memory, plain shared atomics
                const uint64_t item = row0 + trows + tid;
                uint32_t cov = 0, first = 0xFFFFFFFFu;
                for (uint32_t w = 0; w < p.W; ++w) {
                    const uint64_t xw = __ldg(p.bitmap + item * p.Wp + w) & word_mask(p, w);
                    cov += __popcll(xw);
                    if (xw && first == 0xFFFFFFFFu) first = w * 64u + first_bit(xw);
                }
                if (p.countable) p.countable[item] = item ? cov : 0xFFFFFFFFu;
                if (item) {
                    if (HIST) atomicAdd(&s.hist_cnt[cov], 1u);
                    if (cov)
                        for (uint32_t t = 0; t < p.T; ++t)
                            if (cov >= p.cov[t]) atomicAdd(&s.delta_lo[t * p.G + first], 1u);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8u * st);
            if (++st == S) {
                st = 0;
                ph ^= 1u;
            }
        }
        if (tid == 0) ts_mark(p, 3);
        if (HIST && tid == 0) *s_missing = seen;
        // the trees' register planes join the thread's shared-memory planes (slots 0 .. TREE - 1 of each counter word) ...
#pragma unroll
        for (int cw = 0; cw < CW; ++cw) {
            const uint32_t a = my_planes + (uint32_t)cw * (P + TREE) * 2048u;
            const VertTree &t = tree[cw];
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(t.p1.lo), "r"(t.p1.hi) : "memory");
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a + 2048u), "r"(t.p2.lo), "r"(t.p2.hi) : "memory");
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a + 4096u), "r"(t.p4.lo), "r"(t.p4.hi) : "memory");
            if (TREE == 4u) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a + 6144u), "r"(t.p8.lo), "r"(t.p8.hi) : "memory");
        }
    }
    __syncthreads();
    // ===== fold: warp cw adds up counter word cw =====
    // Shuffles are the scarce resource here (~0.25 warp-SHFL per clock and SM: a butterfly over all 8 warps of both
    // resident CTAs cost 5-9 us, profiles/r2_scan_timeline_v2.txt), shared-memory loads are not: lane l first adds the
    // planes of threads l, l + 32, .. l + 224 (conflict-free LDS.64, ripple-carry full adders), then one butterfly over
    // the warp's 32 lanes finishes the sum; lane 0 keeps it for the read-out.
    if (tid == 0) ts_mark(p, 6);
    if (warp < (uint32_t)CW) {
        const uint32_t c = warp;  // counter word
        V64 pl[kVertFoldPlanes];
#pragma unroll
        for (int k = 0; k < kVertFoldPlanes; ++k) pl[k] = {0u, 0u};
        // plane-major: the eight threads' words of a plane are loaded together (independent LDS), then added one after
        // the other, each with its own carry word -- the eight ripple chains run pipelined across the planes instead of
        // one after the other.  (This phase is pure latency, one warp per counter: 3.2 us as written here; rolled loops over
        // the thread groups / butterfly levels were measured at 8.6 us, a butterfly over all eight warps at 5-9 us:
        // profiles/r2_scan_timeline_v*.txt.)
        const uint32_t src = planes0 + ((c * (P + TREE)) * 256u + lane) * 8u;
        V64 carry[kConsumerWarps];
#pragma unroll
        for (int j = 0; j < kConsumerWarps; ++j) carry[j] = {0u, 0u};
#pragma unroll
        for (int k = 0; k < kVertFoldPlanes; ++k) {
            if ((uint32_t)k >= P + TREE + 3u) break;  // 8 numbers of P + TREE planes: 3 planes more
            V64 o[kConsumerWarps];
#pragma unroll
            for (int j = 0; j < kConsumerWarps; ++j) {
                o[j] = {0u, 0u};
                if ((uint32_t)k < P + TREE)
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(o[j].lo), "=r"(o[j].hi) : "r"(src + (uint32_t)k * 2048u + (uint32_t)j * 256u));
            }
#pragma unroll
            for (int j = 0; j < kConsumerWarps; ++j) csa(carry[j], pl[k], pl[k], o[j], carry[j]);
        }
        uint32_t n = TREE + P + 3u;  // planes that can be non-zero
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            V64 carry2 = {0u, 0u};
#pragma unroll
            for (int k0 = 0; k0 < kVertFoldPlanes; k0 += 8) {  // eight planes' shuffles in flight, then their carry chain
                if ((uint32_t)k0 > n) break;
                V64 o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    o[k] = {0u, 0u};
                    if ((uint32_t)(k0 + k) < n) o[k] = {__shfl_xor_sync(0xFFFFFFFFu, pl[k0 + k].lo, off), __shfl_xor_sync(0xFFFFFFFFu, pl[k0 + k].hi, off)};
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if ((uint32_t)(k0 + k) <= n) csa(carry2, pl[k0 + k], pl[k0 + k], o[k], carry2);
            }
            ++n;
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < kVertFoldPlanes; ++k) fold[(size_t)c * kVertFoldPlanes + k] = ((uint64_t)pl[k].hi << 32) | pl[k].lo;
        }
    }
    if (tid == 0) ts_mark(p, 7);
    __syncthreads();
    // ===== read the bins out: bit b of plane k of a counter word's total weighs 2^k =====
    for (uint32_t i = tid; i < (uint32_t)(CW * 64); i += kScanThreads) {
        const uint32_t cw = i >> 6, c = cw / (uint32_t)NW, bin = (cw % (uint32_t)NW) * 64u + (i & 63u);
        const uint64_t *f = fold + (size_t)cw * kVertFoldPlanes;
        uint32_t total = 0;
        for (uint32_t k = 0; k < TREE + P + 8u; ++k) total += (uint32_t)((f[k] >> (i & 63u)) & 1ull) << k;
        if (!total) continue;
        if (HIST && c == 0) {
            atomicAdd(&s.hist_cnt[bin], total);  // (the tail rows' atomics are in there already; bits above G are never set)
            atomicSub(s_missing, total);
        } else {
            cls_tot[(c - HIST) * (64u * NW) + bin] = total;
        }
    }
    __syncthreads();
    // the top coverage 64 NW (G = 64 NW) has no bit in the one-hot words: those items are the ones the bins did not account for
    if (HIST && tid == 0 && p.G == 64u * NW && *s_missing) s.hist_cnt[p.G] += *s_missing;
    for (uint32_t i = tid; i < p.T * p.G; i += kScanThreads) {
        const uint32_t t = i / p.G, f = i - t * p.G;
        s.delta_lo[i] += cls_tot[(p.cls_rank[t] - 1u) * (64u * NW) + f];
    }
    __syncthreads();
    scan_epilogue(p, s, tid);
}

inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1u) / a * a; }

template <bool QUORUM, int C_T, bool MASK>
int launch_one(const ScanParams &p, int grid, cudaStream_t stream) {
    auto kern = k_scan<QUORUM, C_T, MASK>;
    PGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.L.total));
    kern<<<grid, kScanThreads, p.L.total, stream>>>(p);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

}  // namespace

int plan_scan(ScanParams &p, bool quorum, int sm_count, int *grid_out) {
    if (p.G == 0 || p.G > (1u << 20)) return fail(PGX_ERR_UNSUPPORTED, "n_groups must be in 1..2^20");
    if (p.T > (uint32_t)kMaxThresholds) return fail(PGX_ERR_INVALID, "too many thresholds per launch");
    if (p.n_rows >= (1ull << 32)) return fail(PGX_ERR_UNSUPPORTED, "n_items must be < 2^32 - 1");
    const uint32_t rowbytes = p.Wp * 8u;
    const uint32_t G1 = p.G + 1u, TG = p.T * p.G;

    // accumulators
    ScanLayout &L = p.L;
    uint32_t off = 2u * 8u * kMaxStages;  // full[] + empty[] mbarriers
    L.off_acc = off;
    L.off_hist_cnt = off;
    off += (p.flags & kHistCount) ? G1 * 4u : 0u;
    L.off_hist_wlo = off;
    off += (p.flags & kHistWeight) ? G1 * 4u : 0u;
    L.off_hist_whi = off;
    off += (p.flags & kHistWeight) ? G1 * 4u : 0u;
    L.off_delta_lo = off;
    off += TG * 4u;
    L.off_delta_hi = off;
    off += (p.flags & kWeighted) ? TG * 4u : 0u;
    // lane-private counters (no shared atomics) when the private region fits next to a pipeline: G <~ 300 for counts
    L.off_cls_lo = L.off_cls_hi = L.off_carry = L.off_priv = L.off_cbase = off;
    L.priv_bins = L.priv_cw = L.priv_hist_bins = 0;
    p.n_classes = 0;
    bool priv = false, vert = false;
    uint32_t priv_tile = 0, priv_min_stages = 3;
    L.vert_planes = L.vert_counters = L.vert_end = 0;
    // distinct coverage cutoffs, ascending (k_scan_priv: threshold t sums the classes >= its rank; k_scan_vert: counter
    // cls_rank[t] - 1 holds the items of coverage >= cutoff)
    uint32_t D = 0;
    if (!quorum) {
        for (uint32_t t = 0; t < p.T; ++t) {
            bool seen = false;
            for (uint32_t d = 0; d < D; ++d) seen |= p.cls_thr[d] == p.cov[t];
            if (!seen) p.cls_thr[D++] = p.cov[t];
        }
        for (uint32_t i = 1; i < D; ++i)
            for (uint32_t j = i; j > 0 && p.cls_thr[j - 1] > p.cls_thr[j]; --j) {
                const uint32_t tmp = p.cls_thr[j];
                p.cls_thr[j] = p.cls_thr[j - 1];
                p.cls_thr[j - 1] = tmp;
            }
        for (uint32_t t = 0; t < p.T; ++t)
            for (uint32_t d = 0; d < D; ++d)
                if (p.cls_thr[d] == p.cov[t]) p.cls_rank[t] = d + 1u;
    }
    // bit-sliced vertical counters: rows of one or two words (G <= 128), counting (no weights), few distinct cutoffs
    // (PGX_SCAN_VERT=2 switches it off; PGX_SCAN_PRIV=1 / 2 ask for the lane-private / the shared-atomics kernel explicitly)
    if (!quorum && p.Wp <= 2u && !(p.flags & (kWeighted | kHistWeight)) && D <= 3u && env_u32("PGX_SCAN_VERT") != 2u &&
        env_u32("PGX_SCAN_PRIV") == 0u) {
        const uint32_t NW = p.Wp, tree_slots = NW == 1u ? 4u : 3u;
        const uint32_t C = ((p.flags & kHistCount) ? 1u : 0u) + D;
        const uint32_t tile_rows = (uint32_t)kVertK / NW * (uint32_t)kConsumerThreads;  // one block per thread and tile
        const uint64_t n_tiles = (p.n_rows + tile_rows - 1u) / tile_rows;
        uint64_t worst_grid = env_u32("PGX_SCAN_GRID") ? env_u32("PGX_SCAN_GRID") : (uint64_t)sm_count;  // fewest CTAs the launch may use
        if (worst_grid > n_tiles) worst_grid = n_tiles;
        const uint64_t blocks = worst_grid ? (n_tiles + worst_grid - 1u) / worst_grid : 0u;  // blocks a thread adds up
        // 2^P - 1 >= 2 x blocks: with the dynamic tile scheduler a CTA may take up to twice its share before it stops
        // asking for tiles (scan_producer's max_tiles); the grid's joint capacity then still covers every tile
        uint32_t P = 1;
        while (P <= 32u && ((2u * blocks + 1u) >> P)) ++P;
        // (measured: every counter costs ~9 LOP3 per item and row word; from three counters on the loop is ALU-bound and
        // the lane-private kernel is faster -- profiles/r2_scan_shapes_v5.jsonl: 10M x 44, T = 3: 45 vs 43 us)
        const uint32_t max_c = (NW == 1u && env_u32("PGX_SCAN_VERT") == 1u) ? 4u : 2u;  // (PGX_SCAN_VERT=1: up to 4 for one-word rows)
        if (C >= 1u && C <= max_c && P <= (uint32_t)kVertMaxPlanes) {
            vert = true;
            p.flags |= kVertical;
            p.n_classes = D;
            off = align_up(off, 16u);
            L.off_cls_lo = off;
            off += D * 64u * NW * 4u;
            L.off_carry = off = align_up(off, 16u);
            off += C * NW * (uint32_t)kVertFoldPlanes * 8u;
            L.off_priv = off = align_up(off, 16u);
            off += C * NW * (P + tree_slots) * 2048u;  // per counter word: the trees' register planes (fold) + P ripple planes
            L.vert_planes = P;
            L.vert_counters = C;
            L.vert_end = off;
            L.off_cls_hi = L.off_cbase = off;
        }
    }
    if (!quorum && !vert && env_u32("PGX_SCAN_PRIV") != 2u) {
        const bool count_mode = !(p.flags & (kWeighted | kHistWeight));
        const bool weight_mode = !(p.flags & kHistCount) && ((p.flags & kWeighted) || p.T == 0) &&
                                 (p.flags & (kHistWeight | kWeighted)) && p.weight != nullptr;
        if (count_mode || weight_mode) {
            const uint32_t cw = count_mode ? 1u : 2u;
            const uint32_t hist_bins = (count_mode ? (p.flags & kHistCount) : (p.flags & kHistWeight)) ? G1 : 0u;
            // Dp: coverage classes that get lane-private bins.  When histogram + classes do not leave room for two CTAs per
            // SM but the histogram alone does (G ~ 130 .. 300), only the histogram goes private and the curves keep their
            // shared atomics (kPrivGrowthAtomics): half the ATOMS of the atomics kernel, which are what bounds it there.
            uint32_t Dp = D;
            uint32_t min_stages = 3u;
            auto fixed_bytes = [&](uint32_t dp) {
                const uint64_t b = std::max<uint64_t>(1u, (uint64_t)hist_bins + (uint64_t)dp * p.G);
                return (uint64_t)off + (uint64_t)dp * p.G * 4u * (count_mode ? 1u : 2u) + 16u + b * (256u * cw + 4u) + G1 * 4u + 128u;
            };
            auto fits_two = [&](uint32_t dp, uint32_t stages) {
                const uint32_t mt = rowbytes <= 16u ? 1024u : 512u;
                uint32_t t0;
                if (rowbytes <= 8u) t0 = 3072u;
                else if (rowbytes <= 16u) t0 = 2048u;
                else if (rowbytes <= 32u) t0 = 1024u;
                else t0 = std::max(512u, 32768u / rowbytes / 512u * 512u);
                for (uint32_t t = t0; t >= mt; t >>= 1) {
                    const uint64_t st_bytes = align_up(t * rowbytes, 128u) + (p.weight ? align_up(t * 4u, 128u) : 0u);
                    if (fixed_bytes(dp) + stages * st_bytes <= 232448u / 2u - 1024u) return true;
                }
                return false;
            };
            // (not where the joint histogram of the atomics kernel applies, G <~ 100: that is one atomic per item already)
            const bool joint_fits = (uint64_t)G1 * p.G * 4u * (count_mode ? 1u : 2u) <= 44u * 1024u;
            // Measured (profiles/r2_scan_shapes_v6_hybrid.jsonl): pays for bp sums only (10M x 100 / 128, hist + one curve:
            // 86.1 -> 77.9 / 82.4 -> 78.4 us -- a weighted atomic is a lo / carry pair); when counting, the private update is
            // slower than the ATOMS it replaces (10M x 256: 86.8 -> 105 us, c2: 18.8 -> 21.3 us), so counting needs
            // PGX_SCAN_HYBRID=1 to get it.
            const uint32_t hyb = env_u32("PGX_SCAN_HYBRID");
            if (D > 0u && hist_bins > 0u && !joint_fits && !fits_two(D, 3u) && fits_two(0u, 2u) && env_u32("PGX_SCAN_PRIV") != 1u &&
                hyb != 2u && (!count_mode || hyb == 1u)) {
                Dp = 0u;
                min_stages = fits_two(0u, 3u) ? 3u : 2u;
            }
            const uint64_t bins = std::max<uint64_t>(1u, (uint64_t)hist_bins + (uint64_t)Dp * p.G);
            uint32_t tile;
            if (rowbytes <= 8u) tile = 3072u;
            else if (rowbytes <= 16u) tile = 2048u;
            else if (rowbytes <= 32u) tile = 1024u;
            else tile = std::max(512u, 32768u / rowbytes / 512u * 512u);
            if (const uint32_t t_env = env_u32("PGX_SCAN_TILE")) tile = std::max(512u, t_env / 512u * 512u);
            const uint32_t stage = align_up(tile * rowbytes, 128u) + (p.weight ? align_up(tile * 4u, 128u) : 0u);
            // measured (profiles/r2_scan_shapes_*.txt): the private counters only pay with two CTAs per SM -- with one
            // (8 consumer warps) the dependent shared-memory round trips of the updates are not hidden and the atomics
            // kernel (16 warps, fire-and-forget ATOMS) is faster; so: eligible iff 3 stages of some tile fit in half an SM
            const uint64_t fixed = fixed_bytes(Dp);
            const bool two_ctas = fits_two(Dp, min_stages);
            (void)stage;
            if ((two_ctas || env_u32("PGX_SCAN_PRIV") == 1u) && fixed + 2ull * stage <= 232448u && rowbytes <= 64u) {
                priv = true;
                priv_tile = tile;
                priv_min_stages = min_stages;
                p.flags |= kPrivate;
                if (Dp != D) p.flags |= kPrivGrowthAtomics;
                p.n_classes = Dp;
                L.off_cls_lo = off;
                off += Dp * p.G * 4u;
                L.off_cls_hi = off;
                off += count_mode ? 0u : Dp * p.G * 4u;
                L.off_carry = off;
                off += (uint32_t)bins * 4u;
                off = align_up(off, 16u);
                L.off_priv = off;
                L.priv_bins = (uint32_t)bins;
                L.priv_cw = cw;
                L.priv_hist_bins = hist_bins;
                off += (uint32_t)bins * 256u * cw;
                L.off_cbase = off;  // (thr region below keeps it out of the zeroed accumulator words: see acc_words / off_thr)
            }
        }
    }
    // small G: joint (coverage, first group) histogram -- one atomic per item instead of 1 + T
    L.off_joint_cnt = L.off_joint_wlo = L.off_joint_whi = off;
    if (!priv && !vert && !quorum && env_u32("PGX_SCAN_JOINT") != 2u) {
        const bool use_cnt = (p.flags & kHistCount) || !(p.flags & kWeighted);
        const bool use_w = (p.flags & (kHistWeight | kWeighted)) != 0;
        const uint32_t one = G1 * p.G * 4u, bytes = one * ((use_cnt ? 1u : 0u) + (use_w ? 2u : 0u));
        if (bytes <= 44u * 1024u || env_u32("PGX_SCAN_JOINT") == 1u) {
            p.flags |= kJoint;
            L.off_joint_cnt = off;
            off += use_cnt ? one : 0u;
            L.off_joint_wlo = off;
            off += use_w ? one : 0u;
            L.off_joint_whi = off;
            off += use_w ? one : 0u;
        }
    }
    L.acc_words = (off - L.off_acc) / 4u;
    L.off_thr = off;
    off += quorum ? TG * 4u : 0u;
    off += priv ? G1 * 4u : 0u;  // the coverage -> class-bin table of k_scan_priv (at off_cbase == off_thr)
    off = align_up(off, 128u);
    L.off_stage0 = off;

    // tile geometry (rows per stage, CTAs per SM, ring depth): measured optima on B200
    // (tools/sweep_scan.py, profiles/r1_sweep_*.txt).  The quorum kernel keeps one item per thread.
    uint32_t tile, want_ctas, want_stages;
    const bool heavy = quorum || (p.flags & (kHistWeight | kWeighted)) != 0;  // more atomics per item
    if (vert) {
        tile = (uint32_t)kVertK / p.Wp * (uint32_t)kConsumerThreads;  // one block (16 / NW rows) per thread and tile
        want_ctas = 2, want_stages = 3;
    } else if (priv) {
        // two CTAs per SM (16 consumer warps hide the shared-memory latencies of the counter updates) whenever three
        // stages of some tile size fit next to the private region in half an SM's shared memory; else one CTA, deep ring
        tile = priv_tile;
        want_stages = 4;
        want_ctas = 1;
        const uint32_t min_tile = rowbytes <= 16u ? 1024u : 512u;  // a thread takes 4 (narrow rows) or 2 items per step
        for (uint32_t t = priv_tile; t >= min_tile; t >>= 1) {
            const uint32_t stage = align_up(t * rowbytes, 128u) + (p.weight ? align_up(t * 4u, 128u) : 0u);
            if (off + priv_min_stages * stage <= 232448u / 2u - 1024u) {
                tile = t;
                want_ctas = 2;
                want_stages = priv_min_stages;
                break;
            }
        }
    } else if (rowbytes >= 256u) {
        tile = 256u, want_ctas = 1, want_stages = 3;
        // wide rows: fewer rows per stage; shrink further while accumulators + two stages do not fit
        const uint32_t wbytes = p.weight ? 4u : 0u;
        while (tile > 4u && (tile * rowbytes > 65536u || off + 2u * (tile * (rowbytes + wbytes) + 256u) > 232448u)) tile >>= 1;
        if (tile * rowbytes > 98304u) return fail(PGX_ERR_UNSUPPORTED, "n_groups too large for the shared-memory pipeline");
    } else if (rowbytes == 128u) {
        if (heavy) tile = 256u, want_ctas = 2, want_stages = 2;
        else tile = 512u, want_ctas = 1, want_stages = 3;
    } else if (rowbytes >= 64u) {
        tile = 256u, want_ctas = 2, want_stages = 4;
    } else {
        tile = 24576u / rowbytes / 256u * 256u;  // ~24 KB stages for narrow rows
        if (tile > 3072u) tile = 3072u;
        want_ctas = 2, want_stages = 3;
    }
    if (const uint32_t t_env = env_u32("PGX_SCAN_TILE")) tile = (priv || vert) ? tile : ((t_env + 3u) & ~3u);
    p.tile_items = tile;
    L.off_stage_w = align_up(tile * rowbytes, 128u);
    L.stage_stride = L.off_stage_w + (p.weight ? align_up(tile * 4u, 128u) : 0u);

    const uint32_t max_smem = 232448u;  // 227 KB opt-in limit per CTA
    if (const uint32_t c_env = env_u32("PGX_SCAN_CTAS")) want_ctas = c_env;
    if (const uint32_t s_env = env_u32("PGX_SCAN_STAGES")) want_stages = s_env;
    int ctas = (int)want_ctas;
    uint32_t stages = 0;
    for (; ctas >= 1; --ctas) {  // fall back to fewer CTAs per SM when the accumulators are large
        const uint32_t budget = ctas == 1 ? max_smem : max_smem / (uint32_t)ctas - 1024u;
        if (off + 2u * L.stage_stride > budget) continue;
        stages = (budget - off) / L.stage_stride;
        break;
    }
    if (ctas < 1) return fail(PGX_ERR_UNSUPPORTED, "accumulators + pipeline exceed shared memory for this G / T");
    if (stages > want_stages) stages = want_stages;
    if (stages > (uint32_t)kMaxStages) stages = kMaxStages;
    p.stages = stages;
    L.total = off + stages * L.stage_stride;

    const uint64_t n_tiles = (p.n_rows + tile - 1u) / tile;
    p.n_tiles = (uint32_t)n_tiles;
    const uint64_t max_grid = (uint64_t)sm_count * ctas;
    *grid_out = (int)(n_tiles < max_grid ? n_tiles : max_grid);
    if (const uint32_t g_env = env_u32("PGX_SCAN_GRID"))  // test hook: few CTAs walk many tiles each (ring wrap-around, counter folds)
        *grid_out = (int)std::min<uint64_t>(g_env, n_tiles);

    // masks for the two words of a row's last 16-byte chunk
    const uint64_t lastmask = (p.G & 63u) ? ((1ull << (p.G & 63u)) - 1ull) : ~0ull;
    if (p.Wp == 1u) {
        p.last_mask0 = lastmask;
        p.last_mask1 = 0ull;
    } else if (p.W == p.Wp) {
        p.last_mask0 = ~0ull;
        p.last_mask1 = lastmask;
    } else {  // odd W: last real word, then the padding word
        p.last_mask0 = lastmask;
        p.last_mask1 = 0ull;
    }
    return PGX_OK;
}

template <int C_T>
int launch_fast(const ScanParams &p, int grid, cudaStream_t stream) {
    const bool mask = p.last_mask0 != ~0ull || p.last_mask1 != ~0ull;
    return mask ? launch_one<false, C_T, true>(p, grid, stream) : launch_one<false, C_T, false>(p, grid, stream);
}

template <int CW, bool WEIGHTED, int C_T>
int launch_priv_one(const ScanParams &p, int grid, cudaStream_t stream) {
    auto kern = k_scan_priv<CW, WEIGHTED, C_T>;
    PGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.L.total));
    kern<<<grid, kScanThreads, p.L.total, stream>>>(p);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

template <int CW, bool WEIGHTED>
int launch_priv(const ScanParams &p, int grid, cudaStream_t stream) {
    if (p.Wp == 1u) return launch_priv_one<CW, WEIGHTED, -1>(p, grid, stream);
    switch (p.Wp >> 1) {
        case 1: return launch_priv_one<CW, WEIGHTED, 1>(p, grid, stream);
        case 2: return launch_priv_one<CW, WEIGHTED, 2>(p, grid, stream);
        default: return launch_priv_one<CW, WEIGHTED, 0>(p, grid, stream);
    }
}

template <int NW, int HIST, int D, bool C1>
int launch_vert_one(const ScanParams &p, int grid, cudaStream_t stream) {
    auto kern = k_scan_vert<NW, HIST, D, C1>;
    PGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.L.total));
    kern<<<grid, kScanThreads, p.L.total, stream>>>(p);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

template <int NW, int HIST>
int launch_vert(const ScanParams &p, int grid, cudaStream_t stream) {
    const bool c1 = p.n_classes && p.cls_thr[0] <= 1u;
    switch (p.n_classes) {
        case 0:
            if constexpr (HIST != 0) return launch_vert_one<NW, 1, 0, false>(p, grid, stream);
            return fail(PGX_ERR_INVALID, "k_scan_vert without a counter");
        case 1: return c1 ? launch_vert_one<NW, HIST, 1, true>(p, grid, stream) : launch_vert_one<NW, HIST, 1, false>(p, grid, stream);
        case 2:
            if constexpr (NW == 1 || HIST == 0)
                return c1 ? launch_vert_one<NW, HIST, 2, true>(p, grid, stream) : launch_vert_one<NW, HIST, 2, false>(p, grid, stream);
            break;
        default:
            if constexpr (NW == 1)
                return c1 ? launch_vert_one<NW, HIST, 3, true>(p, grid, stream) : launch_vert_one<NW, HIST, 3, false>(p, grid, stream);
            break;
    }
    return fail(PGX_ERR_INVALID, "k_scan_vert: too many counters for two-word rows");
}

int launch_scan(const ScanParams &p, bool quorum, int grid, cudaStream_t stream) {
    if (p.flags & kVertical) {
        if (p.Wp == 1u) return (p.flags & kHistCount) ? launch_vert<1, 1>(p, grid, stream) : launch_vert<1, 0>(p, grid, stream);
        return (p.flags & kHistCount) ? launch_vert<2, 1>(p, grid, stream) : launch_vert<2, 0>(p, grid, stream);
    }
    if (p.flags & kPrivate) return p.L.priv_cw == 1u ? launch_priv<1, false>(p, grid, stream) : launch_priv<2, true>(p, grid, stream);
    if (quorum) return launch_one<true, 0, true>(p, grid, stream);
    if (p.Wp == 1u) return launch_one<false, -1, true>(p, grid, stream);
    switch (p.Wp >> 1) {
        case 1: return launch_fast<1>(p, grid, stream);
        case 2: return launch_fast<2>(p, grid, stream);
        case 4: return launch_fast<4>(p, grid, stream);
        case 8: return launch_fast<8>(p, grid, stream);
        case 16: return launch_fast<16>(p, grid, stream);
        default: return launch_fast<0>(p, grid, stream);
    }
}

}  // namespace pgx
