/*
 * panacus_oracle.c -- see panacus_oracle.h.  TEST INFRASTRUCTURE ONLY (parity oracle / CPU baseline).
 *
 * Every function follows the cited reference lines statement by statement, keeping the
 * reference's loop structure, integer widths and floating-point operation order, because the
 * printed TSV floors f64 values (src/io.rs:484,512) and a 1-ulp difference can flip an integer.
 * log2/exp2 are glibc's, as for a gnu-target Rust build.
 */
#include "panacus_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- Threshold (src/util.rs:351-363) ---------------------------------------------------- */

uint64_t po_threshold_to_absolute(po_threshold t, uint64_t n) {
    if (t.kind == 1) return t.abs_;
    double v = ceil((double)n * t.rel);
    if (!(v > 0.0)) return 0; /* Rust `as usize` saturates at 0 (and maps NaN to 0) */
    return (uint64_t)v;
}

double po_threshold_to_relative(po_threshold t, uint64_t n) {
    if (t.kind == 0) return t.rel;
    return (double)t.abs_ / (double)n;
}

/* ---- AbacusByTotal (src/graph_broker/abacus.rs:539-586, 719-744) ------------------------ */

static void coverage_one_path(uint32_t *countable, uint64_t *last, const uint64_t *items,
                              const uint64_t *id_prefsum, const uint8_t *exclude, uint64_t path_id,
                              uint64_t group_id) {
    /* abacus.rs:719-744 */
    uint64_t start = id_prefsum[path_id];
    uint64_t end = id_prefsum[path_id + 1];
    for (uint64_t j = start; j < end; j++) {
        uint64_t sid = items[j];
        if (last[sid] != group_id && (exclude == NULL || !exclude[sid])) {
            countable[sid] += 1;
            last[sid] = group_id;
        }
    }
}

void po_abacus_by_total(uint64_t n_items, const uint64_t *items, const uint64_t *id_prefsum,
                        const uint64_t *order_path, const uint64_t *order_group, uint64_t n_order,
                        const uint8_t *exclude, uint32_t *countable) {
    /* abacus.rs:548-569 */
    uint64_t *last = (uint64_t *)malloc((n_items + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i <= n_items; i++) {
        countable[i] = 0;
        last[i] = UINT64_MAX;
    }
    countable[0] = UINT32_MAX;
    for (uint64_t k = 0; k < n_order; k++)
        coverage_one_path(countable, last, items, id_prefsum, exclude, order_path[k], order_group[k]);
    free(last);
}

void po_construct_hist(const uint32_t *countable, uint64_t n_items, uint64_t n_groups, uint64_t *hist) {
    /* abacus.rs:746-763 */
    for (uint64_t g = 0; g <= n_groups; g++) hist[g] = 0;
    for (uint64_t i = 0; i <= n_items; i++) {
        uint64_t cov = countable[i];
        if (cov >= n_groups + 1) continue; /* ignored (drops the u32::MAX sentinel at index 0) */
        hist[cov] += 1;
    }
}

void po_construct_hist_bps(const uint32_t *countable, const uint32_t *node_lens, uint64_t n_items,
                           uint64_t n_groups, const uint64_t *unc_ids, const uint64_t *unc_vals,
                           uint64_t n_unc, uint64_t *hist) {
    /* abacus.rs:765-787 */
    for (uint64_t g = 0; g <= n_groups; g++) hist[g] = 0;
    for (uint64_t id = 0; id <= n_items; id++) {
        uint64_t cov = countable[id];
        if (cov >= n_groups + 1) continue;
        hist[cov] += (uint64_t)node_lens[id];
    }
    for (uint64_t k = 0; k < n_unc; k++) {
        hist[countable[unc_ids[k]]] -= unc_vals[k];
        hist[0] += unc_vals[k];
    }
}

/* ---- AbacusByGroup CSR build (abacus.rs:859-986) ---------------------------------------- */

int po_csr_build(uint64_t n_items, const uint64_t *items, const uint64_t *id_prefsum,
                 const uint64_t *order_path, const uint64_t *order_group, uint64_t n_order,
                 const uint8_t *exclude, uint64_t *r, uint64_t **c_out, uint32_t **v_out) {
    /* compute_row_storage_space, abacus.rs:859-899 */
    uint64_t *last = (uint64_t *)malloc((n_items + 1) * sizeof(uint64_t));
    if (!last) return -1;
    for (uint64_t i = 0; i <= n_items; i++) last[i] = UINT64_MAX;
    for (uint64_t i = 0; i < n_items + 2; i++) r[i] = 0;
    for (uint64_t k = 0; k < n_order; k++) {
        uint64_t path_id = order_path[k], group_id = order_group[k];
        uint64_t start = id_prefsum[path_id], end = id_prefsum[path_id + 1];
        for (uint64_t j = start; j < end; j++) {
            uint64_t sid = items[j];
            if (last[sid] != group_id && (exclude == NULL || !exclude[sid])) {
                r[sid] += 1;
                last[sid] = group_id;
            }
        }
    }
    free(last);
    uint64_t acc = 0;
    for (uint64_t i = 0; i < n_items + 2; i++) {
        uint64_t tmp = r[i];
        r[i] = acc;
        acc += tmp;
    }

    /* compute_column_values, abacus.rs:901-986 (report_values = true, graph_broker.rs:382) */
    uint64_t n = r[n_items + 1];
    uint64_t *c = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    uint32_t *v = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
    if (!c || !v) {
        free(c);
        free(v);
        return -1;
    }
    for (uint64_t i = 0; i < n; i++) c[i] = UINT64_MAX;
    for (uint64_t k = 0; k < n_order; k++) {
        uint64_t path_id = order_path[k], group_id = order_group[k];
        uint64_t start = id_prefsum[path_id], end = id_prefsum[path_id + 1];
        for (uint64_t j = start; j < end; j++) {
            uint64_t sid = items[j];
            uint64_t cv_start = r[sid];
            uint64_t cv_end = r[sid + 1];
            if (cv_end != cv_start) {
                if (cv_end - 1 > n) cv_end = n - 1; /* abacus.rs:936-942, never taken */
                uint64_t p = c[cv_end - 1];
                if (c[cv_end - 1] == UINT64_MAX) {
                    c[cv_start] = group_id;
                    if (cv_start < cv_end - 1) c[cv_end - 1] = 0;
                    v[cv_start] += 1;
                } else if (cv_start + p < cv_end - 1) {
                    if (c[cv_start + p] < group_id) {
                        c[cv_end - 1] += 1;
                        p += 1;
                        c[cv_start + p] = group_id;
                    }
                    v[cv_start + p] += 1;
                } else {
                    v[cv_end - 1] += 1;
                }
            }
        }
    }
    *c_out = c;
    *v_out = v;
    return 0;
}

void po_free(void *p) { free(p); }

/* ---- AbacusByGroup::calc_growth (abacus.rs:989-1032) ------------------------------------ */

void po_calc_growth(const uint64_t *r, const uint64_t *c, uint64_t n_items, uint64_t n_groups,
                    po_threshold t_coverage, po_threshold t_quorum, int count_bp,
                    const uint32_t *node_lens, const uint64_t *uncovered, double *res) {
    for (uint64_t j = 0; j < n_groups; j++) res[j] = 0.0;
    uint64_t cthr = po_threshold_to_absolute(t_coverage, n_groups);
    if (cthr < 1) cthr = 1;
    double q = po_threshold_to_relative(t_quorum, n_groups);
    if (!(q > 0.0)) q = 0.0; /* f64::max(0.0, q) */

    for (uint64_t i = 1; i <= n_items; i++) { /* windows (r[i], r[i+1]); first entry ignored */
        uint64_t start = r[i], end = r[i + 1];
        if (end - start >= cthr) {
            uint64_t k = start;
            for (uint64_t j = c[start]; j < n_groups; j++) {
                if (k < end - 1 && c[k + 1] <= j) k += 1;
                double need = ceil(((double)c[k] + 1.0) * q);
                uint64_t need_u = need > 0.0 ? (uint64_t)need : 0;
                if (k - start + 1 >= need_u) {
                    if (!count_bp) {
                        res[j] += 1.0;
                    } else {
                        uint64_t unc = uncovered ? uncovered[i] : 0;
                        uint64_t covered = node_lens[i];
                        if (unc > covered) {
                            /* reference logs an error and adds nothing */
                        } else {
                            res[j] += (double)(covered - unc);
                        }
                    }
                }
            }
        }
    }
}

/* ---- closed-form growth (src/graph_broker/hist.rs:21-187) ------------------------------- */

double po_choose(uint64_t n, uint64_t k) {
    /* hist.rs:21-36 */
    double res = 0.0;
    if (k > n) return 0.0;
    if (k > n - k) k = n - k;
    double nf = (double)n;
    for (uint64_t i = 0; i < k; i++) {
        res += log2(nf - (double)i);
        res -= log2((double)i + 1.0);
    }
    return res;
}

void po_growth_union(const uint64_t *hist, uint64_t n, po_threshold t_cov, double *out) {
    /* hist.rs:89-114 */
    uint64_t c = po_threshold_to_absolute(t_cov, n);
    if (c < 1) c = 1;
    double n_fall_m = 0.0;
    uint64_t tot_u = 0;
    for (uint64_t i = c; i <= n; i++) tot_u += hist[i];
    double tot = (double)tot_u;
    double *perc_mult = (double *)calloc(n + 1, sizeof(double));
    for (uint64_t m = 1; m < n + 1; m++) {
        double y = 0.0;
        n_fall_m += log2((double)n - (double)m + 1.0);
        for (uint64_t i = c; i < n - m + 1; i++) {
            perc_mult[i] += log2((double)n - (double)m - (double)i + 1.0);
            y += exp2(log2((double)hist[i]) + perc_mult[i] - n_fall_m);
        }
        out[m - 1] = tot - y;
    }
    free(perc_mult);
}

void po_growth_core(const uint64_t *hist, uint64_t n, po_threshold t_cov, double *out) {
    /* hist.rs:116-138; note to_absolute(n + 1) */
    uint64_t c = po_threshold_to_absolute(t_cov, n + 1);
    if (c < 1) c = 1;
    double n_fall_m = 0.0;
    double *perc_mult = (double *)calloc(n + 1, sizeof(double));
    for (uint64_t m = 1; m < n + 1; m++) {
        double y = 0.0;
        n_fall_m += log2((double)n - (double)m + 1.0);
        for (uint64_t i = (m > c ? m : c); i < n + 1; i++) {
            perc_mult[i] += log2((double)i - (double)m + 1.0);
            y += exp2(log2((double)hist[i]) + perc_mult[i] - n_fall_m);
        }
        out[m - 1] = y;
    }
    free(perc_mult);
}

void po_growth_quorum(const uint64_t *hist, uint64_t n, po_threshold t_cov, po_threshold t_quorum,
                      double *out) {
    /* hist.rs:140-187 */
    uint64_t c = po_threshold_to_absolute(t_cov, n);
    if (c < 1) c = 1;
    double quorum = po_threshold_to_relative(t_quorum, n);
    double n_fall_m = 0.0;
    double m_fact = 0.0;
    double *perc_mult = (double *)calloc(n + 1, sizeof(double));
    double *q = (double *)calloc((n + 1) * (n + 1), sizeof(double));
    for (uint64_t m = 1; m < n + 1; m++) {
        m_fact += log2((double)m);
        double mq = ceil((double)m * quorum);
        uint64_t m_quorum = mq > 0.0 ? (uint64_t)mq : 0;

        double yl = 0.0;
        n_fall_m += log2((double)n - (double)m + 1.0);
        for (uint64_t i = (m > c ? m : c); i < n + 1; i++) {
            perc_mult[i] += log2((double)i - (double)m + 1.0);
            yl += exp2(log2((double)hist[i]) + perc_mult[i] - n_fall_m);
        }

        double yr = 0.0;
        for (uint64_t i = m_quorum; i < n; i++) {
            double sum_q = 0.0;
            int add = 0;
            double *qi = q + i * (n + 1);
            for (uint64_t j = (m_quorum > c ? m_quorum : c); j < m; j++) {
                if (n + j + 1 > i + m && j <= i) {
                    if (qi[j] == 0.0) qi[j] = po_choose(i, j);
                    qi[j] += log2((double)n - (double)i - (double)m + 1.0 + (double)j);
                    qi[j] -= log2((double)m - (double)j);
                    sum_q += exp2(qi[j] + m_fact - n_fall_m);
                    add = 1;
                }
            }
            if (add) yr += exp2(log2((double)hist[i]) + log2(sum_q));
        }
        out[m - 1] = yl + yr;
    }
    free(perc_mult);
    free(q);
}

uint64_t po_hist_calc_growth(const uint64_t *hist, uint64_t n, po_threshold t_cov,
                             po_threshold t_quorum, double *out) {
    /* hist.rs:51-66 */
    if (n == 0) return 0;
    uint64_t quorum = po_threshold_to_absolute(t_quorum, n);
    if (quorum < 1) quorum = 1;
    if (quorum == 1)
        po_growth_union(hist, n, t_cov, out);
    else if (quorum >= n)
        po_growth_core(hist, n, t_cov, out);
    else
        po_growth_quorum(hist, n, t_cov, t_quorum, out);
    return n;
}

/* ---- Similarity::set_table (src/analyses/similarity.rs:119-163) ------------------------- */

int po_similarity(const uint64_t *r, const uint64_t *c, uint64_t n_items, uint64_t n_groups,
                  int count_bp, const uint32_t *node_lens, uint64_t *inter, uint64_t *len,
                  float *table) {
    uint64_t G = n_groups;
    uint8_t *seen = (uint8_t *)calloc(G ? G : 1, 1);
    memset(inter, 0, G * G * sizeof(uint64_t));
    memset(len, 0, G * sizeof(uint64_t));
    /* tuple_windows over r, enumerate from 0 (item 0 has an empty row), similarity.rs:125-150 */
    for (uint64_t index = 0; index <= n_items; index++) {
        uint64_t node_length = count_bp ? (uint64_t)node_lens[index] : 1;
        for (uint64_t a = r[index]; a < r[index + 1]; a++) {
            uint64_t x = c[a];
            len[x] += node_length;
            seen[x] = 1;
            for (uint64_t b = r[index]; b < r[index + 1]; b++) inter[x * G + c[b]] += node_length;
        }
    }
    int rc = 0;
    for (uint64_t i = 0; i < G; i++)
        if (!seen[i]) rc = -1; /* path_lens[&i] would panic (similarity.rs:161) */
    if (table && rc == 0) {
        for (uint64_t i = 0; i < G; i++)
            for (uint64_t j = 0; j < G; j++) {
                uint64_t intersection = inter[i * G + j];
                table[i * G + j] = (float)intersection / (float)(len[i] + len[j] - intersection);
            }
    }
    free(seen);
    return rc;
}
