/*
 * panacus_b200.h -- C ABI of libpanacus_b200.so: the B200 (sm_100a) implementation of panacus's
 * counting hot path (coverage histogram, ordered / permuted growth, all-pairs group intersections).
 *
 * The reference (marschall-lab/panacus @ 395ba41, v0.4.1) has no FFI; the seam this library serves
 * is the set of in-process Rust call sites listed next to each entry point below.  A Rust
 * `extern "C"` block binding exactly these symbols is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns PGX_OK (0) or a negative pgx_status; nothing throws or unwinds;
 *     pgx_last_error() returns a thread-local description of the last failure.
 *   - all pointers are plain host pointers unless the parameter name starts with `d_`
 *     (device pointer on the handle's device).  The caller owns every buffer it passes.
 *   - one handle is driven by one host thread at a time; handles on different devices are
 *     independent.
 *   - shapes: n_items < 2^32 - 2; the histogram / growth kernels keep their per-CTA accumulators in shared
 *     memory, which bounds n_groups to ~17,000 (both histograms in one call), ~26,000 (one of them), and the
 *     number of thresholds per launch (more are split over several passes); larger shapes return
 *     PGX_ERR_UNSUPPORTED.
 *   - item ids are 1..=n_items; row 0 of every per-item array is the reference's dummy item
 *     (src/graph_broker/graph.rs:323-324, abacus.rs:551,1000-1002) and is never counted.
 *
 * Data layout ("abacus bitmap", node-major): (n_items + 1) rows of `row_words` u64, row i = item i,
 * bit (g % 64) of word (g / 64) set iff group g contains item i (after the reference's de-duplication
 * and exclude rules, abacus.rs:719-744 / 859-899).  row_words = pgx_row_words(n_groups): ceil(G/64)
 * rounded up to an even number when > 1, so rows are 16-byte aligned for 128-bit loads / TMA bulk copies.
 * Bits >= n_groups and the padding word are ignored.  weight[i] = node length in bp (u32,
 * graph.rs:341-350) minus the item's uncovered bps (abacus.rs:1016-1023) when counting bp.
 */
#ifndef PANACUS_B200_H
#define PANACUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PGX_OK = 0,
    PGX_ERR_INVALID = -1,     /* bad argument */
    PGX_ERR_CUDA = -2,        /* CUDA runtime error (message in pgx_last_error) */
    PGX_ERR_NOMEM = -3,       /* host or device allocation failed */
    PGX_ERR_UNSUPPORTED = -4, /* shape outside what the kernels support */
    PGX_ERR_STATE = -5,       /* call order violated (e.g. no bitmap uploaded yet) */
    PGX_ERR_EXCHANGE = -6,    /* multi-GPU exchange failed: a peer never arrived (watchdog) or a collective call failed */
    PGX_ERR_NCCL = -7         /* NCCL could not be loaded or returned an error (message in pgx_last_error) */
} pgx_status;

typedef struct pgx_abacus pgx_abacus;

/* library / device ------------------------------------------------------------------------------ */
const char *pgx_version(void);
const char *pgx_last_error(void);
int pgx_device_count(int *n);
/* Creates the CUDA context of `device` (what the first pgx_abacus_create would otherwise pay, ~1-2 s in a cold
 * process).  Thread safe and idempotent: a host program calls it from a side thread at start-up so that context
 * creation overlaps its own input parsing. */
int pgx_device_warmup(int device);
/* u64 words per bitmap row for n_groups groups (see layout above). */
uint32_t pgx_row_words(uint32_t n_groups);

/* handle ---------------------------------------------------------------------------------------- */
/* Allocates a zeroed device bitmap of (n_items+1) x pgx_row_words(n_groups) u64 and a weight vector
 * (all ones).  Replaces the host-side AbacusByGroup{r,c,v} / AbacusByTotal{countable}
 * (src/graph_broker/abacus.rs:486-503, 790-800). */
int pgx_abacus_create(pgx_abacus **out, int device, uint64_t n_items, uint32_t n_groups);
void pgx_abacus_destroy(pgx_abacus *a);
/* Work is enqueued on `cuda_stream` (a cudaStream_t; NULL = the handle's own stream). */
int pgx_abacus_set_stream(pgx_abacus *a, void *cuda_stream);
int pgx_abacus_shape(const pgx_abacus *a, uint64_t *n_items, uint32_t *n_groups, uint32_t *row_words);

/* Host -> device copy of a packed bitmap ((n_items+1) x host_row_words u64, host_row_words >=
 * ceil(G/64)) and of the weights (n_items+1 u32, or NULL = keep).  bitmap may be NULL (= keep). */
int pgx_abacus_upload(pgx_abacus *a, const uint64_t *bitmap, uint32_t host_row_words, const uint32_t *weight);
/* Use caller-owned device buffers in place (no copy); d_weight may be NULL (= unit weights).
 * The bitmap must use pgx_row_words(n_groups) words per row. */
int pgx_abacus_adopt_device(pgx_abacus *a, uint64_t *d_bitmap, uint32_t *d_weight);
/* Device-side build from ItemTable slices (src/util.rs:80-93): ORs bit `group_id` into the row of
 * every item in items[0..n_steps) that is not flagged in `exclude` (n_items+1 bytes, or NULL).
 * Same result as AbacusByTotal::coverage / compute_row_storage_space's de-duplicated incidence
 * (abacus.rs:719-744, 859-899).  Ids are the reference's u64 ItemIdSize. */
int pgx_abacus_scatter(pgx_abacus *a, const uint64_t *items, uint64_t n_steps, uint32_t group_id,
                       const uint8_t *exclude);
/* Whole-graph device-side build: the complete ItemTable (items[0..n_steps), id_prefsum[0..n_paths]) plus the
 * counting order as a per-path group id (path_group[p] = group of path p, or -1 if the path is not counted:
 * excluded / outside the subset / absent from the order, abacus.rs:310-347, 555-569).  One upload, one kernel:
 * every step finds its path by binary search in id_prefsum and ORs its group's bit into the item's row.
 * Replaces the per-path loops of item_table_to_abacus (abacus.rs:539-586) and both CSR passes
 * (abacus.rs:859-986). */
int pgx_abacus_build(pgx_abacus *a, const uint64_t *items, uint64_t n_steps, const uint64_t *id_prefsum,
                     uint64_t n_paths, const int64_t *path_group, const uint8_t *exclude);
/* Same with the ids narrowed to u32 (always possible: n_items < 2^32; for the Rust side a one-line change of
 * `ItemIdSize`, src/util.rs:15): half the bytes over PCIe.
 * Both builds stream the table in 16 Mi-step chunks through two device staging buffers, the upload of chunk k + 1
 * overlapping the scatter kernel of chunk k.  If `items` is page-locked host memory (pgx_host_alloc, cudaHostAlloc,
 * cudaHostRegister) it is read by DMA where it lies; pageable memory is staged through pinned buffers of the handle,
 * u64 ids being narrowed to u32 by host threads on the way.  The bits are ORed into the bitmap (pgx_abacus_clear
 * first to rebuild from scratch). */
int pgx_abacus_build_u32(pgx_abacus *a, const uint32_t *items, uint64_t n_steps, const uint64_t *id_prefsum,
                         uint64_t n_paths, const int64_t *path_group, const uint8_t *exclude);
/* Page-locked host memory for tables and result buffers handed to this library (cudaHostAlloc, portable): uploads from
 * it and downloads into it are single DMA transfers.  Every entry point also accepts ordinary (pageable) memory. */
int pgx_host_alloc(void **out, size_t bytes);
void pgx_host_free(void *p);
int pgx_abacus_clear(pgx_abacus *a);
/* Device-to-device copy of an item range (bitmap rows + weights when the source has them): items
 * src_first_item .. src_first_item + dst.n_items - 1 of `src` become items 1 .. dst.n_items of `dst` (same n_groups; the
 * handles may live on different GPUs: NVLink peer copy).  Cuts an abacus built once into per-GPU item-range shards. */
int pgx_abacus_copy_rows(pgx_abacus *dst, pgx_abacus *src, uint64_t src_first_item);
/* AbacusByGroup's CSR {r, c, v} (abacus.rs:790-799) as built by compute_row_storage_space (abacus.rs:859-899) and
 * compute_column_values (abacus.rs:901-986, report_values = true, src/graph_broker.rs:382), derived on the device:
 *   r[0 .. n_items+1]  row offsets (n_items + 2 entries, r[0] = r[1] = 0: item 0 is the empty dummy row); from the
 *                      row popcounts of the bitmap + a device-wide exclusive scan.  *nnz = r[n_items + 1].
 *   c[0 .. nnz)        group ids of each row, ascending (the reference's GroupSize = u64): the set bits of the row
 *   v[0 .. nnz)        occurrence count of each (item, group) incidence (CountSize = u32): one atomic add per ItemTable
 *                      step; needs the same tables the bitmap was built from (pgx_abacus_build).  v may be NULL
 *                      (then the table arguments are ignored), c may be NULL.
 * Two calls because nnz is only known after the first.  Consumer in the reference: AbacusByGroup::to_tsv
 * (abacus.rs:1056-1178), i.e. the `table` analysis (src/analyses/table.rs:14-35). */
int pgx_abacus_csr_rows(pgx_abacus *a, uint64_t *r, uint64_t *nnz);
int pgx_abacus_csr_fill(pgx_abacus *a, const uint64_t *items, uint64_t n_steps, const uint64_t *id_prefsum,
                        uint64_t n_paths, const int64_t *path_group, const uint8_t *exclude, uint64_t *c, uint32_t *v);
/* Device -> host copy of the bitmap in the packed host layout (host_row_words per row). */
int pgx_abacus_download(pgx_abacus *a, uint64_t *bitmap, uint32_t host_row_words);

/* hot path -------------------------------------------------------------------------------------- */
/* Coverage histogram.  Replaces AbacusByTotal::coverage + construct_hist / construct_hist_bps
 * (abacus.rs:719-787) as called from Hist::from_abacus (src/graph_broker/hist.rs:39-49,
 * src/graph_broker.rs:353-362), before the sparse uncovered_bps patch (abacus.rs:779-785), which
 * stays on the host.
 *   hist_count[c]  = #items (1..=N) contained in exactly c groups        (G+1 entries, or NULL)
 *   hist_weight[c] = sum of weight[i] over those items                    (G+1 entries, or NULL)
 *   countable[i]   = number of groups containing item i; countable[0] = UINT32_MAX
 *                    (N+1 entries, or NULL) */
int pgx_hist(pgx_abacus *a, uint64_t *hist_count, uint64_t *hist_weight, uint32_t *countable);

/* Ordered growth.  Replaces AbacusByGroup::calc_growth (abacus.rs:989-1032) as called from
 * analyses/ordered_histgrowth.rs:184-186 and io.rs:575, for n_thresholds (coverage, quorum) pairs
 * in one pass over the bitmap.
 *   cov_abs[t]        = max(1, t_coverage.to_absolute(G))                          (abacus.rs:997)
 *   quorum_thr[t*G+g] = ceil((g as f64 + 1.0) * max(0, q_t)) as usize, computed by the host in f64
 *                       (abacus.rs:998,1010); NULL = all zero (q = 0 for every pair)
 *   col_order         = NULL: growth in group-index order 0..G-1.  Otherwise a permutation of
 *                       0..G-1: position j of the curve adds group col_order[j] (the reference gets
 *                       the same by re-building the abacus under `--order`, abacus.rs:324-326); when
 *                       given, quorum_thr is indexed by position j.
 *   weighted          = 0: count items (node / edge); 1: sum weight[i] (bp)
 *   curve[t*G+j]      = exact integer value of res[j]; the reference's f64 is `curve as f64`. */
int pgx_ordered_growth(pgx_abacus *a, uint32_t n_thresholds, const uint32_t *cov_abs,
                       const uint32_t *quorum_thr, const uint32_t *col_order, int weighted,
                       uint64_t *curve);

/* hist + ordered growth in a single pass over the bitmap (what `histgrowth`-style runs need);
 * any of hist_count / hist_weight may be NULL. */
int pgx_hist_ordered_growth(pgx_abacus *a, uint64_t *hist_count, uint64_t *hist_weight,
                            uint32_t n_thresholds, const uint32_t *cov_abs, const uint32_t *quorum_thr,
                            int weighted, uint64_t *curve);

/* Ordered growth under n_orders different group orders (the permutation-sampled growth estimator
 * of BASELINE.json config 3; each order is one `--order` run of the reference).
 *   orders[p*G + j]  = group added at position j of order p (each row a permutation of 0..G-1)
 *   curves[(p*T + t)*G + j] */
int pgx_permuted_growth(pgx_abacus *a, uint32_t n_orders, const uint32_t *orders, uint32_t n_thresholds,
                        const uint32_t *cov_abs, const uint32_t *quorum_thr, int weighted, uint64_t *curves);

/* Integer part of Similarity::set_table (src/analyses/similarity.rs:125-150) for group rows
 * [row_begin, row_end):
 *   inter[(x-row_begin)*G + y] = sum over items of w * [x in item] * [y in item]
 *   len[x]                     = sum over items of w * [x in item]          (all G groups)
 * with w = 1 (weighted = 0) or weight[i] (weighted = 1).  The f32 Jaccard division
 * (similarity.rs:153-163) and the clustering stay on the host. */
int pgx_similarity(pgx_abacus *a, int weighted, uint32_t row_begin, uint32_t row_end, uint64_t *inter,
                   uint64_t *len);
/* Same for the part of the rows on and right of the diagonal: only inter[(x-row_begin)*G + y] with y >= row_begin is
 * computed, the columns left of row_begin are returned as zero.  The matrix is symmetric, so the ranks of a multi-GPU
 * run can split the upper triangle instead of whole rows (panacus_b200/sharding.py: two folded row blocks per rank)
 * and mirror it after the all-gather -- about half the work of row blocks. */
int pgx_similarity_upper(pgx_abacus *a, int weighted, uint32_t row_begin, uint32_t row_end, uint64_t *inter,
                         uint64_t *len);

/* asynchronous / device-resident variants (used for kernel-only timing and multi-GPU reduction) -- */
/* Enqueues the fused pass on the handle's stream and leaves the raw accumulators in device memory:
 *   d_out[0 .. G]                 hist_count
 *   d_out[G+1 .. 2G+1]            hist_weight
 *   d_out[2G+2 + t*G + j]         first-difference of curve t at column j (curve = prefix sum)
 * d_out must hold pgx_fused_out_words(G, T) u64.  No synchronisation is performed. */
size_t pgx_fused_out_words(uint32_t n_groups, uint32_t n_thresholds);
int pgx_fused_pass_async(pgx_abacus *a, int want_hist_count, int want_hist_weight, uint32_t n_thresholds,
                         const uint32_t *cov_abs, const uint32_t *quorum_thr, int weighted,
                         uint64_t *d_out);
/* multi-GPU: fused exchange over NVLink peer memory --------------------------------------------------
 * Item-range sharding (one process per GPU, each handle holds the rows of its item range): after
 * pgx_exchange_connect, every fused pass (pgx_hist, pgx_ordered_growth without col_order,
 * pgx_hist_ordered_growth, pgx_fused_pass_async) ends with an in-kernel all-reduce of the result
 * vector: the last CTA of each rank stores its partial sums into every rank's exchange buffer with peer
 * stores, signals, waits for all ranks and adds them up -- one kernel does the scan and the exchange.
 * The calls are collective: all ranks must issue the same sequence of passes.
 *   pgx_exchange_export  writes an opaque PGX_EXCHANGE_HANDLE_BYTES handle of this process' buffer
 *                        (a cudaIpcMemHandle_t); gather the handles of all ranks with any transport.
 *   pgx_exchange_connect maps the peers' buffers (world <= 8 GPUs of one NVLink domain). */
#define PGX_EXCHANGE_HANDLE_BYTES 64
int pgx_exchange_export(pgx_abacus *a, void *handle_out);
int pgx_exchange_connect(pgx_abacus *a, uint32_t rank, uint32_t world, const void *all_handles);
int pgx_exchange_disconnect(pgx_abacus *a);
/* The synchronous fused passes return PGX_ERR_EXCHANGE themselves when the in-kernel watchdog (10 s) saw a peer
 * missing.  After pgx_fused_pass_async call this once the work is due: it synchronises the handle's stream and
 * reports (and clears) the watchdog flag; the device result of a flagged pass is a partial sum and must be dropped. */
int pgx_exchange_status(pgx_abacus *a);

/* multi-GPU: NCCL communicator + sharded entry points --------------------------------------------------------
 * One pgx_comm per GPU (one process per GPU, or one thread per GPU in a single process).  NCCL is loaded at run time
 * (dlopen of libnccl.so.2: the copy PyTorch already loaded when inside a torch process); without it these calls return
 * PGX_ERR_NCCL and everything single-GPU keeps working.  All *_sharded calls and pgx_abacus_broadcast are COLLECTIVE:
 * every rank of the communicator calls them with the same arguments (its own handle), in the same order; the work is
 * enqueued on the handle's stream and the result arrives on every rank.
 *   pgx_comm_unique_id   rank 0 obtains an id (ncclGetUniqueId) and hands it to the other ranks by any transport
 *   pgx_comm_create      ncclCommInitRank on `device`
 *   pgx_comm_create_all  single-process variant (ncclCommInitAll): out[i] drives devices[i]; use one host thread per
 *                        communicator for the collective calls */
#define PGX_COMM_ID_BYTES 128
typedef struct pgx_comm pgx_comm;
int pgx_comm_unique_id(void *id_out /* PGX_COMM_ID_BYTES */);
int pgx_comm_create(pgx_comm **out, int device, uint32_t rank, uint32_t world, const void *unique_id);
int pgx_comm_create_all(pgx_comm **out /* n */, uint32_t n, const int *devices);
void pgx_comm_destroy(pgx_comm *c);
int pgx_comm_info(const pgx_comm *c, uint32_t *rank, uint32_t *world, int *device);
/* Replicates rank `root`'s device bitmap (and weights when with_weights != 0) into every rank's handle over NVLink
 * (ncclBroadcast) -- one H2D upload per node instead of one per GPU.  All handles must have the same shape. */
int pgx_abacus_broadcast(pgx_abacus *a, pgx_comm *c, uint32_t root, int with_weights);
/* pgx_exchange_export + all-gather of the handles over the communicator + pgx_exchange_connect (multi-process only:
 * CUDA IPC handles cannot be opened in the process that exported them). */
int pgx_exchange_connect_comm(pgx_abacus *a, pgx_comm *c);
/* Item-range sharding (each rank's handle holds the rows of its own item range): pgx_hist_ordered_growth followed by
 * the path's one exchange, an ncclAllReduce (u64 sum) of the KB-sized result vector; results of the whole graph on
 * every rank.  The NCCL twin of the in-kernel exchange above (which must not be connected at the same time). */
int pgx_hist_ordered_growth_sharded(pgx_abacus *a, pgx_comm *c, uint64_t *hist_count, uint64_t *hist_weight,
                                    uint32_t n_thresholds, const uint32_t *cov_abs, const uint32_t *quorum_thr,
                                    int weighted, uint64_t *curve);
/* Work-item sharding (every rank's handle holds the WHOLE bitmap):
 *   pgx_permuted_growth_sharded  order p is computed by rank p % world; the curves stay on the device, are
 *                                all-gathered (ncclAllGather) and copied to the host once; `curves` as in
 *                                pgx_permuted_growth, complete on every rank.  The reference's parallel axis on this
 *                                path is the threshold pairs only (src/analyses/ordered_histgrowth.rs:174-188).
 *   pgx_similarity_sharded       the sum runs over the items, so the items are split: rank r computes the whole matrix
 *                                (upper triangle + mirror) over its share of the 64-item words, one ncclAllReduce (u64
 *                                sum) of the partial matrices, one copy to the host: inter (G x G) and len (G, = the
 *                                diagonal) complete on every rank.  Similarity::set_table
 *                                (src/analyses/similarity.rs:119-163) is serial in the reference. */
int pgx_permuted_growth_sharded(pgx_abacus *a, pgx_comm *c, uint32_t n_orders, const uint32_t *orders,
                                uint32_t n_thresholds, const uint32_t *cov_abs, const uint32_t *quorum_thr, int weighted,
                                uint64_t *curves);
int pgx_similarity_sharded(pgx_abacus *a, pgx_comm *c, int weighted, uint64_t *inter, uint64_t *len);
/* Host-only helper for callers that shard the similarity matrix by ROWS themselves with pgx_similarity_upper (the
 * torch.distributed twin in panacus_b200/sharding.py does): 2 * world + 1 tile-aligned boundaries; rank r takes blocks r and
 * 2 * world - 1 - r, each from its diagonal rightwards -- an equal share of the pair work. */
int pgx_similarity_shard_bounds(uint32_t n_groups, uint32_t world, uint32_t *bounds);

/* Kernel timing for measurements: while enabled, every hot-path kernel section a call launches (k_scan, the group-major
 * growth kernels, k_gm_similarity) is bracketed by CUDA events on the handle's stream.  pgx_kernel_time_ms waits for
 * them, returns the summed device time of the sections recorded since the last query and forgets them. */
int pgx_abacus_set_timing(pgx_abacus *a, int enable);
int pgx_kernel_time_ms(pgx_abacus *a, float *total_ms, uint32_t *n_sections);

/* Number of kernel launches issued through this handle so far (for bench accounting). */
uint64_t pgx_launch_count(const pgx_abacus *a);
/* Name and launch geometry of the last hot-path kernel (for logs). */
int pgx_last_launch_info(const pgx_abacus *a, char *buf, size_t buflen);

#ifdef __cplusplus
}
#endif
#endif /* PANACUS_B200_H */
