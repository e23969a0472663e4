// pgx_internal.h -- interfaces between the translation units of libpanacus_b200 (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>

#include "../../include/panacus_b200.h"

namespace pgx {

constexpr int kMaxThresholds = 8;       // (coverage, quorum) pairs per launch
constexpr int kConsumerWarps = 8;       // warps that scan items
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kScanThreads = kConsumerThreads + 32;  // + one TMA producer warp
constexpr int kMaxStages = 8;

enum ScanFlags : uint32_t {
    kHistCount = 1u,   // accumulate hist_count
    kHistWeight = 2u,  // accumulate hist_weight
    kWeighted = 4u,    // growth deltas sum weight[i] instead of 1
    kJoint = 8u,       // small G: one joint (coverage, first group) histogram, marginalised in the epilogue
    kPrivate = 16u,    // lane-private narrow counters (plain LDS / STS; a shared atomic only when one wraps), folded after the last tile
    kPrivGrowthAtomics = 64u,  // kPrivate, but only the histogram is lane-private: the curves' bins keep their shared atomics
    kVertical = 32u,   // G <= 64, counts: bit-sliced vertical counters (carry-save adders over one-hot words), no atomics in the loop
};

// Byte offsets into the dynamic shared memory of k_scan (identical on host and device).
struct ScanLayout {
    uint32_t off_acc;        // start of the u32 accumulator region (zeroed at kernel start)
    uint32_t acc_words;      // its length in u32
    uint32_t off_hist_cnt;   // u32[G+1]
    uint32_t off_hist_wlo;   // u32[G+1]
    uint32_t off_hist_whi;   // u32[G+1]
    uint32_t off_delta_lo;   // u32[T*G]  (counts when not weighted)
    uint32_t off_delta_hi;   // u32[T*G]  (weighted only)
    uint32_t off_thr;        // u32[T*G]  (quorum kernel only)
    uint32_t off_joint_cnt;  // u32[(G+1)*G] joint histogram of counts      (kJoint)
    uint32_t off_joint_wlo;  // u32[(G+1)*G] joint histogram of weights, low  (kJoint, weights in use)
    uint32_t off_joint_whi;  //                                       high
    uint32_t off_cls_lo;     // u32[D*G] per coverage class (kPrivate): items whose coverage reaches exactly d thresholds, by first group
    uint32_t off_cls_hi;     //          high words (weights)
    uint32_t off_cbase;      // u32[G+1]: coverage -> first bin of the item's coverage class, ~0 = not counted (kPrivate)
    uint32_t off_carry;      // u32[priv_bins]: what wrapped out of the narrow lane-private counters
    uint32_t off_priv;       // lane-private counters: priv_bins rows of 256 counters of priv_cw bytes (kPrivate)
    uint32_t priv_bins, priv_cw, priv_hist_bins;
    // kVertical (k_scan_vert): per-thread high bit-planes [counter][plane][thread] u64 at off_priv, the warps' folded
    // planes at off_carry, per-class totals u32[D][64] at off_cls_lo; everything in [off_acc, vert_end) starts at zero
    uint32_t vert_planes, vert_counters, vert_end;
    uint32_t off_stage0;     // first pipeline stage (128-byte aligned)
    uint32_t stage_stride;   // bytes per stage
    uint32_t off_stage_w;    // offset of the weight tile inside a stage
    uint32_t total;          // dynamic shared memory bytes
};

constexpr int kMaxRanks = 8;  // GPUs of one NVSwitch domain taking part in the fused exchange

// In-kernel all-reduce of the KB-sized result vector over NVLink peer memory (world > 1): the last CTA
// of every rank stores its vector into slot [epoch parity][rank] of every rank's buffer as 8-byte
// {epoch, half-word} packets (atomic stores: no flag / fence round trip), then spins on its own buffer
// until every rank's packets carry the current epoch and sums them.
struct Exchange {
    uint64_t *data[kMaxRanks];   // data[r]: rank r's buffer as mapped into this process (r == rank: local)
    unsigned int *err;           // set to 2 if a peer did not arrive within the timeout
    uint32_t world, rank, epoch, stride;  // stride: result words per slot (each word = 2 packets)
};

struct ScanParams {
    const uint64_t *bitmap;  // n_rows x Wp, node-major
    const uint32_t *weight;  // n_rows, or nullptr (unit weights)
    uint32_t *countable;     // n_rows, or nullptr
    const uint32_t *thr;     // device T*G quorum thresholds (quorum kernel), else nullptr
    uint64_t *acc;           // global accumulators (zero on entry, zero again on exit)
    uint64_t *out;           // caller's fused-layout result buffer, written by the last CTA
    unsigned int *ticket;    // CTA completion counter (zero on entry and exit); ticket[1]: "out is zeroed" epoch flag;
                             // ticket[2], ticket[3]: tile counters of even / odd launches
    uint32_t zero_epoch;     // != 0 (single GPU): CTA 0 zeroes the requested words of `out` and publishes this epoch in
                             // ticket[1]; every CTA then adds its sums straight into `out` -- no ticket, no snapshot
    uint64_t n_rows;         // N + 1
    uint32_t G, W, Wp;
    uint32_t T;
    uint32_t cov[kMaxThresholds];
    uint32_t slot[kMaxThresholds];  // threshold k's deltas go to out[2(G+1) + slot[k]*G ...]
    uint32_t flags;
    // kPrivate: the D distinct coverage cutoffs (ascending); an item of coverage c is in class #{d : c >= cls_thr[d]};
    // threshold t's curve sums the classes >= cls_rank[t] (1-based)
    uint32_t n_classes, cls_thr[kMaxThresholds], cls_rank[kMaxThresholds];
    uint32_t tile_items;     // rows per pipeline stage (multiple of 4)
    uint32_t stages;
    uint32_t n_tiles;
    uint64_t last_mask0, last_mask1;  // masks of the two words of the last 16-byte chunk of a row
    ScanLayout L;
    Exchange x;              // x.world <= 1: no exchange
    uint32_t sched_dynamic;  // 1: tiles after a CTA's first one come from the global counter ticket[2 + sched_parity]
    uint32_t sched_parity;   // launch parity: this launch's counter; CTA 0 zeroes the other one for the next launch
    uint64_t *dbg_ts;        // PGX_SCAN_TS=1: per CTA 8 globaltimer stamps (phase timeline of one launch), else nullptr
};

struct ScanPlan {
    ScanParams p;
    int grid;
    int ctas_per_sm;
};

// Fills tile geometry, shared-memory layout and grid for one fused pass.  Returns PGX_OK or an error.
int plan_scan(ScanParams &p, bool quorum, int sm_count, int *grid_out);
// Enqueues the pass.  quorum = false: hist + first-set-bit growth (all thresholds have q = 0);
// quorum = true: general thresholds (p.thr must be set), no hist.
int launch_scan(const ScanParams &p, bool quorum, int grid, cudaStream_t stream);

// ---- group-major ("path-major") side ----------------------------------------------------------
// gm layout: G rows x gm_stride u64, bit (i % 64) of word (i / 64) of row g set iff item i is in
// group g; gm_stride = ceil((N+1)/64) rounded up to a multiple of 16 words (128-byte rows).
uint64_t gm_stride_words(uint64_t n_rows);
int launch_transpose(const uint64_t *bitmap, uint64_t n_rows, uint32_t G, uint32_t Wp, uint64_t *gm,
                     uint64_t gm_stride, const uint32_t *perm /*nullptr = natural item order*/, cudaStream_t stream);
int sort_items_by_weight(const uint32_t *weight, uint64_t n_rows, uint32_t *perm, uint32_t *sorted_w, cudaStream_t stream,
                         const uint32_t *secondary = nullptr);  // secondary: tie order among equal weights (descending)

struct GmGrowthParams {
    const uint64_t *gm;      // G x gm_stride
    uint64_t gm_stride;
    uint64_t n_words;        // ceil(n_rows / 64)
    uint64_t n_rows;
    const uint32_t *weight;  // n_rows or nullptr
    const uint32_t *countable;  // n_rows (needed when any cov > 1)
    const uint32_t *order;   // device: n_orders x G entries, group at position j
    uint32_t n_orders;
    const uint32_t *thr;     // device: T*G thresholds indexed by position, or nullptr (all q = 0)
    uint64_t *out;           // device: first differences, order o at out + o*out_order_stride (zeroed by the launcher)
    uint64_t out_order_stride;  // words between consecutive orders in out (>= T*G)
    uint32_t G, T;
    uint32_t cov[kMaxThresholds];
    uint32_t slot[kMaxThresholds];  // threshold k's first differences go to out + slot[k]*G
    uint32_t general_mask;   // bit t set: threshold t needs the rank comparison (q > 0)
    uint32_t direct_out;     // very large G: no shared-memory staging of the deltas, atomics go straight to `out`
    uint32_t col_fastest;    // grid mapping: 0 = consecutive CTAs take the same column block under different orders (L2 reuse),
                             // 1 = consecutive CTAs walk the column blocks of one order (PGX_GM_GRID=col; for measurements)
    uint32_t n_fast;         // k_gm_quorum / k_gm_union: cov / slot index T .. T+n_fast-1 are q = 0 thresholds of the pass
    const uint32_t *perm;       // k_gm_quorum, weighted: gm / weight are in weight-sorted item order, perm[pos] = item
    const uint64_t *uniform_w;  // k_gm_quorum, weighted: (1 << 32) | w per column whose items all weigh w, else 0
    int weighted;
};
int launch_gm_growth(const GmGrowthParams &p, int sm_count, cudaStream_t stream);
size_t gm_growth_smem_bytes(uint32_t G, uint32_t T, bool any_general, bool direct_out = false);
// table-driven general-quorum kernel (pgx_quorum.cu): T <= kGmQuorumMaxT general thresholds (+ one q = 0) per launch
constexpr uint32_t kGmQuorumMaxT = 4;
constexpr size_t kGmQuorumSmemMax = 227u * 1024u;
int gm_quorum_planes(uint32_t G);  // rank bit-planes the kernel is instantiated with for G groups (0: unsupported)
size_t gm_quorum_smem_bytes(uint32_t G, uint32_t T, uint32_t NF, bool weighted);
uint32_t gm_quorum_fast_slots(uint32_t T, uint32_t n_fast);  // NF of the instantiation that serves (T, n_fast)
int launch_gm_quorum(const GmGrowthParams &p, cudaStream_t stream);

struct GmSimParams {
    const uint64_t *gm;
    uint64_t gm_stride;
    uint64_t n_words;
    const uint64_t *planes;  // weighted: n_planes x gm_stride weight bit-planes, else nullptr
    const uint64_t *uniform_w;   // weighted: per word (1 << 32) | w if the word's items share one weight, else 0
    const uint32_t *plane_mask;  // weighted: non-empty planes per word
    uint32_t n_planes;
    uint32_t G;
    uint32_t row_begin, row_end;
    uint32_t col_begin;      // only columns >= col_begin are computed (the rest of `inter` stays zero)
    uint32_t triangular;     // set by the launcher: full square, compute upper tiles only and mirror
    uint32_t csa;            // unweighted: carry-save variant (one POPC per two item words and pair)
    uint32_t upper_only;     // skip tiles strictly below the diagonal (sharded runs read only (x, y) with y >= x)
    uint64_t *inter;         // device (row_end-row_begin) x G, zeroed by the launcher
};
int launch_gm_similarity(const GmSimParams &p, int sm_count, cudaStream_t stream);
// the same contract on the tensor cores (pgx_simmma.cu: tcgen05.mma.kind::i8 on the bits expanded to u8), unweighted only;
// p.triangular must be set by the caller for a full square (then follow with launch_sim_mirror)
int launch_sim_mma(const GmSimParams &p, int sm_count, cudaStream_t stream);
int launch_sim_mirror(uint64_t *inter, uint32_t G, cudaStream_t stream);
// first differences -> curves, in place: n_curves rows of G u64
int launch_prefix_curves(uint64_t *d, uint64_t n_curves, uint32_t G, cudaStream_t stream);
int launch_gm_rowsum(const uint64_t *gm, uint64_t gm_stride, uint64_t n_words, const uint64_t *planes,
                     uint32_t n_planes, const uint64_t *uniform_w, uint32_t G, uint64_t *len, cudaStream_t stream);
int launch_weight_planes(const uint32_t *weight, uint64_t n_rows, uint64_t *planes, uint64_t gm_stride,
                         uint32_t n_planes, uint64_t *uniform_w, uint32_t *plane_mask, int skip0, cudaStream_t stream);

// scatter build (ItemTable slice -> bitmap bits)
int launch_scatter(uint64_t *bitmap, uint32_t Wp, uint64_t n_rows, const uint64_t *d_items, uint64_t n_steps,
                   uint32_t group_id, const uint8_t *d_exclude, unsigned int *d_err, cudaStream_t stream);

// whole-ItemTable build: step s belongs to path p with prefsum[p] <= s < prefsum[p+1]
// (d_items: ids as u32 (id_bytes = 4) or u64 (8))
int launch_build(uint64_t *bitmap, uint32_t Wp, uint64_t n_rows, uint32_t G, const void *d_items, int id_bytes, uint64_t step0,
                 uint64_t n_steps, const uint64_t *d_prefsum, uint64_t n_paths, const int64_t *d_path_group,
                 const uint8_t *d_exclude, unsigned int *d_err, cudaStream_t stream);


// AbacusByGroup CSR {r, c, v} from the bitmap (+ the ItemTable for the occurrence counts v); pgx_csr.cu
int launch_csr_rows(const uint64_t *bitmap, uint64_t n_rows, uint32_t G, uint32_t W, uint32_t Wp, uint64_t *d_r /*n_rows + 1*/,
                    cudaStream_t stream);  // synchronises the stream
int launch_csr_cols(const uint64_t *bitmap, uint64_t n_rows, uint32_t G, uint32_t W, uint32_t Wp, const uint64_t *d_r,
                    uint64_t *d_c, cudaStream_t stream);
int launch_csr_vals(const uint64_t *bitmap, uint64_t n_rows, uint32_t G, uint32_t Wp, const uint64_t *d_items, uint64_t step0,
                    uint64_t n_steps, const uint64_t *d_prefsum, uint64_t n_paths, const int64_t *d_path_group,
                    const uint8_t *d_exclude, const uint64_t *d_r, uint32_t *d_v /*zeroed by the caller*/, unsigned int *d_err,
                    cudaStream_t stream);

}  // namespace pgx
