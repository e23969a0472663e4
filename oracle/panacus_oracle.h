/*
 * panacus_oracle.h -- CPU restatement of panacus's hist / growth / ordered-histgrowth /
 * similarity arithmetic (reference: marschall-lab/panacus @ 395ba41, v0.4.1).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  Nothing under panacus_b200/
 * links, imports or executes it.
 *
 * The reference is Rust and cannot be compiled in this image (no cargo/rustc), so there is no
 * oracle/_ref build.  Pinning status of this restatement:
 *   - closed-form growth (union/core/quorum, choose): PINNED by the exact-f64 KATs in
 *     src/graph_broker/hist.rs:341-398 and by the chr22 hist->growth golden embedded in
 *     docs/chr22.hprc-v1.0-pggb.histgrowth.html:267-274 (tests/golden/).
 *   - coverage / hist: PINNED by src/graph_broker/abacus.rs:1525,1579,1630 (chrM node/edge/bp
 *     hists) and tests/test_files/t_groups.hist.tsv.
 *   - ordered growth (AbacusByGroup::calc_growth), CSR build and similarity: the reference holds
 *     no golden vector for them -> "parity unpinned" (only the curve end point is pinned through
 *     the hist goldens: last value for c=1,q=0 equals sum(hist[1..])).
 *
 * Integer widths follow src/util.rs:15-17: ItemIdSize = u64, CountSize = u32, GroupSize = u64.
 */
#ifndef PANACUS_ORACLE_H
#define PANACUS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Threshold (src/util.rs:327-364). kind 0 = Relative(f64), 1 = Absolute(usize). */
typedef struct {
    int kind;
    double rel;
    uint64_t abs_;
} po_threshold;

uint64_t po_threshold_to_absolute(po_threshold t, uint64_t n); /* src/util.rs:351-356 */
double po_threshold_to_relative(po_threshold t, uint64_t n);   /* src/util.rs:358-363 */

/* AbacusByTotal::coverage over a whole path order (src/graph_broker/abacus.rs:539-586, 719-744).
 * items/id_prefsum: ItemTable (src/util.rs:80-93).  order_path[k], order_group[k]: output of
 * get_path_order with the running group index.  exclude: N+1 flags or NULL.
 * countable: N+1 u32, [0] = u32::MAX on return. */
void po_abacus_by_total(uint64_t n_items, const uint64_t *items, const uint64_t *id_prefsum,
                        const uint64_t *order_path, const uint64_t *order_group, uint64_t n_order,
                        const uint8_t *exclude, uint32_t *countable);

/* AbacusByTotal::construct_hist (abacus.rs:746-763). hist: G+1 entries. */
void po_construct_hist(const uint32_t *countable, uint64_t n_items, uint64_t n_groups, uint64_t *hist);

/* AbacusByTotal::construct_hist_bps (abacus.rs:765-787). unc_ids/unc_vals: uncovered_bps map. */
void po_construct_hist_bps(const uint32_t *countable, const uint32_t *node_lens, uint64_t n_items,
                           uint64_t n_groups, const uint64_t *unc_ids, const uint64_t *unc_vals,
                           uint64_t n_unc, uint64_t *hist);

/* AbacusByGroup CSR build: compute_row_storage_space + compute_column_values
 * (abacus.rs:859-986), literal two-pass transcription with the in-row cursor trick.
 * r: N+2 entries (caller allocated).  *c_out / *v_out: malloc'ed, nnz entries; free with po_free. */
int po_csr_build(uint64_t n_items, const uint64_t *items, const uint64_t *id_prefsum,
                 const uint64_t *order_path, const uint64_t *order_group, uint64_t n_order,
                 const uint8_t *exclude, uint64_t *r, uint64_t **c_out, uint32_t **v_out);
void po_free(void *p);

/* AbacusByGroup::calc_growth (abacus.rs:989-1032).  count_bp != 0 selects the Bp arm.
 * uncovered: dense N+1 array of uncovered bps (0 where absent) or NULL.  res: G doubles. */
void po_calc_growth(const uint64_t *r, const uint64_t *c, uint64_t n_items, uint64_t n_groups,
                    po_threshold t_coverage, po_threshold t_quorum, int count_bp,
                    const uint32_t *node_lens, const uint64_t *uncovered, double *res);

/* Closed-form growth (src/graph_broker/hist.rs:21-187). hist has n+1 entries; out has n. */
double po_choose(uint64_t n, uint64_t k);
void po_growth_union(const uint64_t *hist, uint64_t n, po_threshold t_cov, double *out);
void po_growth_core(const uint64_t *hist, uint64_t n, po_threshold t_cov, double *out);
void po_growth_quorum(const uint64_t *hist, uint64_t n, po_threshold t_cov, po_threshold t_quorum,
                      double *out);
/* Hist::calc_growth dispatch (hist.rs:51-66); returns number of values written (n or 0). */
uint64_t po_hist_calc_growth(const uint64_t *hist, uint64_t n, po_threshold t_cov,
                             po_threshold t_quorum, double *out);

/* Similarity::set_table integer part + f32 Jaccard (src/analyses/similarity.rs:119-163).
 * inter: G*G u64 (dense stand-in for the HashMap<u128,usize>), len: G u64, table: G*G f32
 * (may be NULL).  Returns -1 if a group has no entry in path_lens (reference would panic). */
int po_similarity(const uint64_t *r, const uint64_t *c, uint64_t n_items, uint64_t n_groups,
                  int count_bp, const uint32_t *node_lens, uint64_t *inter, uint64_t *len,
                  float *table);

#ifdef __cplusplus
}
#endif
#endif
