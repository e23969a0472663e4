// pgx_api.cu -- the C ABI of libpanacus_b200 (see include/panacus_b200.h): handle management,
// host<->device staging, threshold routing and the launch sequences of the hot-path kernels.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "pgx_handle.h"

namespace pgx {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }
int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

}  // namespace pgx

using namespace pgx;

namespace pgx {

KernelTimer::KernelTimer(pgx_abacus *h) : a(h) {
    if (!a->timing) return;
    if (a->ev_used == a->ev_pool.size()) {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        a->ev_pool.emplace_back(e0, e1);
    }
    auto &pr = a->ev_pool[a->ev_used++];
    cudaEventRecord(pr.first, a->stream);
    stop = pr.second;
}
KernelTimer::~KernelTimer() {
    if (stop) cudaEventRecord(stop, a->stream);
}

int ensure_pinned(pgx_abacus *a, size_t words) {
    if (a->pinned_words >= words) return PGX_OK;
    if (a->h_pinned) cudaFreeHost(a->h_pinned);
    a->h_pinned = nullptr;
    a->pinned_words = 0;
    PGX_CUDA(cudaMallocHost(reinterpret_cast<void **>(&a->h_pinned), words * sizeof(uint64_t)));
    a->pinned_words = words;
    return PGX_OK;
}

int copy_to_host(pgx_abacus *a, uint64_t *dst, const uint64_t *d_src, size_t words) {
    if (words == 0) return PGX_OK;
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {  // the caller handed us page-locked memory: DMA straight into it
        PGX_CUDA(cudaMemcpyAsync(dst, d_src, words * 8u, cudaMemcpyDeviceToHost, a->stream));
        PGX_CUDA(cudaStreamSynchronize(a->stream));
        return PGX_OK;
    }
    const int rc = ensure_pinned(a, words);
    if (rc) return rc;
    PGX_CUDA(cudaMemcpyAsync(a->h_pinned, d_src, words * 8u, cudaMemcpyDeviceToHost, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    std::memcpy(dst, a->h_pinned, words * 8u);
    return PGX_OK;
}

void invalidate_derived(pgx_abacus *a) {
    a->countable_valid = false;
    a->gm_valid = false;
    a->gm_c_valid = false;
    a->planes_valid = false;
    a->csr_valid = false;
}

int check_handle(const pgx_abacus *a) {
    if (!a) return fail(PGX_ERR_INVALID, "null handle");
    if (!a->d_bitmap) return fail(PGX_ERR_STATE, "abacus has no bitmap");
    return PGX_OK;
}

// After a fused pass with the multi-GPU exchange connected and the stream synchronised: did the in-kernel watchdog
// fire (a peer never delivered its partial sums)?  The result buffer then holds a partial sum and must not be used.
int check_exchange(pgx_abacus *a) {
    if (a->x.world <= 1u) return PGX_OK;
    unsigned int err = 0;
    PGX_CUDA(cudaMemcpyAsync(&err, a->d_err + 1, 4, cudaMemcpyDeviceToHost, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    if (err) {
        PGX_CUDA(cudaMemsetAsync(a->d_err + 1, 0, 4, a->stream));
        return fail(PGX_ERR_EXCHANGE, "exchange timeout: a peer rank never delivered its partial sums (result discarded)");
    }
    return PGX_OK;
}

bool all_zero(const uint32_t *thr, uint32_t G) {
    if (!thr) return true;
    for (uint32_t g = 0; g < G; ++g)
        if (thr[g]) return false;
    return true;
}

int upload_thr(pgx_abacus *a, const std::vector<uint32_t> &thr) {
    if (thr.empty()) return PGX_OK;
    if (a->thr_cache == thr && a->d_thr) return PGX_OK;
    int rc = ensure_dev(&a->d_thr, &a->thr_cap, thr.size());
    if (rc) return rc;
    PGX_CUDA(cudaMemcpyAsync(a->d_thr, thr.data(), thr.size() * 4u, cudaMemcpyHostToDevice, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));  // thr is a stack/vector buffer: finish before it goes away
    a->thr_cache = thr;
    return PGX_OK;
}

int validate_thresholds(const pgx_abacus *a, uint32_t T, const uint32_t *cov) {
    if (T && !cov) return fail(PGX_ERR_INVALID, "cov_abs is null");
    for (uint32_t t = 0; t < T; ++t)
        if (cov[t] == 0) return fail(PGX_ERR_INVALID, "cov_abs entries must be >= 1 (abacus.rs:997 clamps)");
    (void)a;
    return PGX_OK;
}


// General-quorum thresholds in group order can run on either layout: k_scan<quorum> (node-major, one
// thread per item, divergent) or k_gm_growth on the group-major copy (bit-sliced over 64 items per
// thread, ~10x faster on large tables but needs the transposed copy).  PGX_QUORUM_PATH=scan|gm overrides.
bool quorum_via_gm(const pgx_abacus *a) {
    if (a->x.world > 1u) return false;  // the fused multi-GPU exchange lives in k_scan
    const char *env = getenv("PGX_QUORUM_PATH");
    if (env && !strcmp(env, "scan")) return false;
    if (env && !strcmp(env, "gm")) return true;
    return a->gm_valid || (size_t)a->n_rows * a->Wp * 8u >= (size_t)8 << 20;
}

// Launches k_scan for the thresholds ts[i0 .. i0+n) (global indices into cov/thr); n is reduced until
// accumulators + pipeline fit in shared memory.  Returns the number of thresholds handled in *n_done.
int scan_launch(pgx_abacus *a, bool quorum, uint32_t flags, const std::vector<uint32_t> &ts, size_t i0,
                const uint32_t *cov, const uint32_t *thr, uint32_t *d_countable, uint64_t *d_out, size_t *n_done) {
    const uint32_t G = a->G;
    size_t n = std::min<size_t>(kMaxThresholds, ts.size() - i0);
    ScanParams p;
    int grid = 0;
    for (;;) {
        std::memset(&p, 0, sizeof(p));
        p.bitmap = a->d_bitmap;
        p.weight = (flags & (kWeighted | kHistWeight)) ? a->d_weight : nullptr;  // count modes never stage the weights
        p.acc = a->d_acc;
        p.ticket = a->d_ticket;
        p.out = d_out;
        p.countable = d_countable;
        p.n_rows = a->n_rows;
        p.G = G;
        p.W = a->W;
        p.Wp = a->Wp;
        p.flags = flags;
        p.x = a->x;
        p.T = (uint32_t)n;
        for (size_t k = 0; k < n; ++k) {
            p.cov[k] = cov[ts[i0 + k]];
            p.slot[k] = ts[i0 + k];
        }
        const int rc = plan_scan(p, quorum, a->sm_count, &grid);
        if (rc == PGX_OK) break;
        const size_t n_min = quorum ? 1 : 0;  // the fast kernel may run with hist only
        if (rc != PGX_ERR_UNSUPPORTED || n <= n_min) return rc;
        --n;
    }
    if (quorum) {
        std::vector<uint32_t> packed(n * (size_t)G);
        for (size_t k = 0; k < n; ++k) std::memcpy(packed.data() + k * G, thr + (size_t)ts[i0 + k] * G, (size_t)G * 4u);
        const int rc = upload_thr(a, packed);
        if (rc) return rc;
        p.thr = a->d_thr;
    }
    if (a->x.world > 1u) {
        p.x.epoch = ++a->epoch;  // collective sequence number (never 0)
    } else if (!getenv("PGX_SCAN_TICKET")) {  // (PGX_SCAN_TICKET=1: the round-1 ticket + snapshot epilogue, for measurements)
        if (++a->scan_epoch == 0u) ++a->scan_epoch;
        p.zero_epoch = a->scan_epoch;
    }
    p.sched_dynamic = getenv("PGX_SCAN_STATIC") ? 0u : 1u;  // (PGX_SCAN_STATIC=1: round-robin tiles, for measurements)
    p.sched_parity = a->scan_launches++ & 1u;
    uint64_t *d_ts = nullptr;
    if (getenv("PGX_SCAN_TS")) {  // measurement aid: per-CTA phase time stamps of this launch, summarised on stderr
        if (cudaMalloc(reinterpret_cast<void **>(&d_ts), (size_t)grid * 64u) == cudaSuccess) {
            cudaMemsetAsync(d_ts, 0, (size_t)grid * 64u, a->stream);
            p.dbg_ts = d_ts;
        }
    }
    int rc;
    {
        KernelTimer kt(a);
        rc = launch_scan(p, quorum, grid, a->stream);
    }
    if (d_ts) {
        std::vector<uint64_t> ts((size_t)grid * 8u);
        cudaStreamSynchronize(a->stream);
        cudaMemcpy(ts.data(), d_ts, ts.size() * 8u, cudaMemcpyDeviceToHost);
        cudaFree(d_ts);
        uint64_t t0 = ~0ull;
        for (int b = 0; b < grid; ++b)
            if (ts[(size_t)b * 8u]) t0 = std::min(t0, ts[(size_t)b * 8u]);
        static const char *names[8] = {"cta_start", "init_done", "first_tile", "loop_done", "epilogue_start", "reds_issued", "fold_start", "fold_done"};
        fprintf(stderr, "PGX_SCAN_TS grid=%d (ns after the first CTA's start: min / median / max over CTAs)", grid);
        for (int k = 0; k < 8; ++k) {
            std::vector<uint64_t> v;
            for (int b = 0; b < grid; ++b)
                if (ts[(size_t)b * 8u + k]) v.push_back(ts[(size_t)b * 8u + k] - t0);
            if (v.empty()) continue;
            std::sort(v.begin(), v.end());
            fprintf(stderr, " | %s %llu / %llu / %llu", names[k], (unsigned long long)v.front(), (unsigned long long)v[v.size() / 2],
                    (unsigned long long)v.back());
        }
        fprintf(stderr, "\n");
    }
    if (rc) return rc;
    a->launches++;
    char buf[320];
    if (p.flags & kVertical)
        snprintf(buf, sizeof buf, "k_scan_vert<hist=%u,D=%u> grid=%d block=%d smem=%u tile_items=%u stages=%u tiles=%u T=%u planes=%u row_words=%u",
                 (p.flags & kHistCount) ? 1u : 0u, p.n_classes, grid, kScanThreads, p.L.total, p.tile_items, p.stages, p.n_tiles, p.T, p.L.vert_planes, p.Wp);
    else if (p.flags & kPrivate)
        snprintf(buf, sizeof buf, "k_scan_priv<u%u%s> grid=%d block=%d smem=%u tile_items=%u stages=%u tiles=%u T=%u classes=%u bins=%u",
                 p.L.priv_cw * 8u, (p.flags & kPrivGrowthAtomics) ? ",hist-only" : "", grid, kScanThreads, p.L.total, p.tile_items, p.stages,
                 p.n_tiles, p.T, p.n_classes, p.L.priv_bins);
    else
        snprintf(buf, sizeof buf, "k_scan<%s%s> grid=%d block=%d smem=%u tile_items=%u stages=%u tiles=%u T=%u", quorum ? "quorum" : "fast",
                 (p.flags & kJoint) ? ",joint" : "", grid, kScanThreads, p.L.total, p.tile_items, p.stages, p.n_tiles, p.T);
    a->last_launch = buf;
    *n_done = n;
    return PGX_OK;
}

// Fused pass for any number of thresholds; results in device memory `d_out` (fused layout with T
// thresholds).  q = 0 thresholds ride along with the histogram in k_scan<fast>; the others go to
// k_scan<quorum>.  Each launch takes as many thresholds as shared memory allows.
int fused_pass(pgx_abacus *a, bool want_cnt, bool want_w, uint32_t T, const uint32_t *cov, const uint32_t *thr,
               int weighted, uint32_t *d_countable, uint64_t *d_out) {
    int rc = validate_thresholds(a, T, cov);
    if (rc) return rc;
    const uint32_t G = a->G;
    std::vector<uint32_t> fast_t, gen_t;
    for (uint32_t t = 0; t < T; ++t) (all_zero(thr ? thr + (size_t)t * G : nullptr, G) ? fast_t : gen_t).push_back(t);
    const uint32_t wflag = weighted ? kWeighted : 0u;

    bool hist_pending = want_cnt || want_w || (d_countable && gen_t.empty());
    size_t i = 0;
    while (hist_pending || i < fast_t.size()) {
        uint32_t flags = wflag | (hist_pending && want_cnt ? kHistCount : 0u) | (hist_pending && want_w ? kHistWeight : 0u);
        size_t n = 0;
        rc = scan_launch(a, false, flags, fast_t, i, cov, thr, hist_pending || gen_t.empty() ? d_countable : nullptr, d_out, &n);
        if (rc == PGX_ERR_UNSUPPORTED && hist_pending && want_cnt && want_w) {
            // very large G: both histograms do not fit in shared memory together -> one pass each
            flags = wflag | kHistCount;
            if ((rc = scan_launch(a, false, flags, fast_t, fast_t.size(), cov, thr, d_countable, d_out, &n))) return rc;
            d_countable = nullptr;
            flags = wflag | kHistWeight;
            rc = scan_launch(a, false, flags, fast_t, i, cov, thr, nullptr, d_out, &n);
        }
        if (rc) return rc;
        if (!hist_pending && n == 0) return fail(PGX_ERR_UNSUPPORTED, "n_groups too large: no threshold fits in shared memory");
        if (hist_pending) d_countable = nullptr;  // written once
        hist_pending = false;
        i += n;
    }
    if (!gen_t.empty() && quorum_via_gm(a)) {
        bool need_cov = false;
        for (uint32_t t : gen_t) need_cov |= cov[t] > 1u;
        if (need_cov && (rc = ensure_countable(a))) return rc;
        if ((rc = ensure_gm(a))) return rc;
        if (!a->d_identity) {
            std::vector<uint32_t> id(G);
            for (uint32_t g = 0; g < G; ++g) id[g] = g;
            PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_identity), (size_t)G * 4u));
            PGX_CUDA(cudaMemcpyAsync(a->d_identity, id.data(), (size_t)G * 4u, cudaMemcpyHostToDevice, a->stream));
            PGX_CUDA(cudaStreamSynchronize(a->stream));
        }
        uint64_t *base = d_out + 2u * ((size_t)G + 1u);
        for (uint32_t t : gen_t) PGX_CUDA(cudaMemsetAsync(base + (size_t)t * G, 0, (size_t)G * 8u, a->stream));
        if (d_countable && a->d_countable && d_countable != a->d_countable)
            PGX_CUDA(cudaMemcpyAsync(d_countable, a->d_countable, a->n_rows * 4u, cudaMemcpyDeviceToDevice, a->stream));
        return gm_growth_launch(a, 1, a->d_identity, gen_t, cov, thr, weighted, base, 0);
    }
    i = 0;
    while (i < gen_t.size()) {
        size_t n = 0;
        rc = scan_launch(a, true, wflag, gen_t, i, cov, thr, d_countable, d_out, &n);
        if (rc) return rc;
        d_countable = nullptr;
        i += n;
    }
    return PGX_OK;
}

int ensure_countable(pgx_abacus *a) {
    if (a->countable_valid) return PGX_OK;
    if (!a->d_countable) PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_countable), a->n_rows * 4u));
    // the by-product histogram goes to its own buffer: d_scratch may hold a caller's partial results
    if (!a->d_hist_tmp) PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_hist_tmp), pgx_fused_out_words(a->G, 0) * 8u));
    int rc = fused_pass(a, true, false, 0, nullptr, nullptr, 0, a->d_countable, a->d_hist_tmp);
    if (rc) return rc;
    a->countable_valid = true;
    return PGX_OK;
}

int ensure_gm(pgx_abacus *a) {
    if (a->gm_valid) return PGX_OK;
    a->gm_stride = gm_stride_words(a->n_rows);
    if (!a->d_gm) PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_gm), (size_t)a->G * a->gm_stride * 8u));
    int rc;
    {
        KernelTimer kt(a);
        rc = launch_transpose(a->d_bitmap, a->n_rows, a->G, a->Wp, a->d_gm, a->gm_stride, nullptr, a->stream);
    }
    if (rc) return rc;
    a->launches++;
    a->gm_valid = true;
    return PGX_OK;
}

// Coverage-sorted group-major copy (items by descending coverage; bit position i of every row = item d_perm_c[i]).
int ensure_gm_cov(pgx_abacus *a) {
    if (a->gm_c_valid) return PGX_OK;
    int rc = ensure_countable(a);
    if (rc) return rc;
    a->gm_stride = gm_stride_words(a->n_rows);
    if (!a->d_gm_c) PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_gm_c), (size_t)a->G * a->gm_stride * 8u));
    if (!a->d_perm_c) PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_perm_c), a->n_rows * 4u));
    uint32_t *d_keys_out = nullptr;  // the sorted coverages themselves are not needed afterwards
    PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_keys_out), a->n_rows * 4u));
    {
        KernelTimer kt(a);
        rc = sort_items_by_weight(a->d_countable, a->n_rows, a->d_perm_c, d_keys_out, a->stream);
        if (!rc) rc = launch_transpose(a->d_bitmap, a->n_rows, a->G, a->Wp, a->d_gm_c, a->gm_stride, a->d_perm_c, a->stream);
    }
    cudaStreamSynchronize(a->stream);
    cudaFree(d_keys_out);
    if (rc) return rc;
    a->launches += 2;
    a->gm_c_valid = true;
    return PGX_OK;
}

__global__ void k_max_u32(const uint32_t *w, uint64_t n, unsigned int *out) {
    unsigned int m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        m = max(m, w[i]);
    m = __reduce_max_sync(0xFFFFFFFFu, m);
    if ((threadIdx.x & 31u) == 0) atomicMax(out, m);
}

// largest weight of the table (one reduction kernel when the weights were adopted on the device)
int ensure_max_weight(pgx_abacus *a) {
    if (a->max_weight_known || !a->d_weight) return PGX_OK;
    PGX_CUDA(cudaMemsetAsync(a->d_err, 0, 4, a->stream));
    k_max_u32<<<296, 256, 0, a->stream>>>(a->d_weight + 1, a->n_items, a->d_err);
    PGX_CUDA(cudaGetLastError());
    unsigned int m = 0;
    PGX_CUDA(cudaMemcpyAsync(&m, a->d_err, 4, cudaMemcpyDeviceToHost, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    PGX_CUDA(cudaMemsetAsync(a->d_err, 0, 4, a->stream));
    a->max_weight = m;
    a->max_weight_known = true;
    return PGX_OK;
}

int ensure_planes(pgx_abacus *a) {
    int rc_mw;
    if (a->planes_valid) return PGX_OK;
    if (!a->d_weight) {
        a->n_planes = 0;
        a->planes_valid = true;
        return PGX_OK;
    }
    if (a->n_rows > 0x7FFFFFFFull)  // cub::DeviceRadixSort is called with an int item count
        return fail(PGX_ERR_UNSUPPORTED, "weight-sorted copy (bp-weighted similarity / permuted growth) needs n_items < 2^31 - 1");
    if ((rc_mw = ensure_max_weight(a))) return rc_mw;
    uint32_t np = 0;
    while (np < 32u && (a->max_weight >> np)) ++np;
    a->gm_stride = gm_stride_words(a->n_rows);
    cudaFree(a->d_planes);
    a->d_planes = nullptr;
    a->n_planes = np;
    const size_t gm_bytes = (size_t)a->G * a->gm_stride * 8u;
    if (!a->d_gm_w) PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_gm_w), gm_bytes));
    if (!a->d_perm) PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_perm), a->n_rows * 4u));
    if (!a->d_sorted_w) PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_sorted_w), a->n_rows * 4u));
    if (!a->d_uniform_w) PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_uniform_w), a->gm_stride * 8u));
    if (!a->d_plane_mask) PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_plane_mask), a->gm_stride * 4u));
    PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_planes), std::max<size_t>((size_t)np, 1) * a->gm_stride * 8u));
    // among equal weights the items are ordered by coverage when that is known: half the nodes of a pangenome are 1 bp
    // long, and inside such a run k_gm_quorum's warps then hold items of one coverage class (see ensure_gm_cov)
    const char *cs = getenv("PGX_GM_COVSORT");
    a->planes_cov_order = a->countable_valid && a->d_countable != nullptr && !(cs && !strcmp(cs, "0"));
    int rc = sort_items_by_weight(a->d_weight, a->n_rows, a->d_perm, a->d_sorted_w, a->stream,
                                  a->planes_cov_order ? a->d_countable : nullptr);
    if (rc) return rc;
    if ((rc = launch_transpose(a->d_bitmap, a->n_rows, a->G, a->Wp, a->d_gm_w, a->gm_stride, a->d_perm, a->stream))) return rc;
    if ((rc = launch_weight_planes(a->d_sorted_w, a->n_rows, a->d_planes, a->gm_stride, np, a->d_uniform_w, a->d_plane_mask, 0,
                                   a->stream)))
        return rc;
    a->launches += 4;
    a->planes_valid = true;
    return PGX_OK;
}

bool gm_grid_col_fastest() {
    const char *env = getenv("PGX_GM_GRID");
    return env && !strcmp(env, "col");
}

int is_permutation(const uint32_t *order, uint32_t G) {
    std::vector<uint8_t> seen(G, 0);
    for (uint32_t j = 0; j < G; ++j) {
        if (order[j] >= G || seen[order[j]]) return 0;
        seen[order[j]] = 1;
    }
    return 1;
}

// Enqueues k_gm_growth for the thresholds `ts` (indices into cov / thr) under n_orders device-resident orders.
// First differences of threshold ts[k], order o land at d_out_base + o*out_order_stride + ts[k]*G (pre-zeroed
// by the caller).  Thresholds are taken in chunks that fit in shared memory.  q = 0 thresholds are HBM-bound here;
// general ones only use this kernel when k_gm_quorum cannot (see gm_growth_launch).
int gm_growth_launch_legacy(pgx_abacus *a, uint32_t n_orders, const uint32_t *d_orders, const std::vector<uint32_t> &ts,
                     const uint32_t *cov, const uint32_t *thr, int weighted, uint64_t *d_out_base, uint64_t out_order_stride) {
    const uint32_t G = a->G;
    int rc;
    for (size_t i0 = 0; i0 < ts.size();) {
        size_t n = std::min<size_t>(kMaxThresholds, ts.size() - i0);
        auto is_general = [&](size_t k) { return !all_zero(thr ? thr + (size_t)ts[i0 + k] * G : nullptr, G); };
        auto any_general = [&](size_t cnt) {
            for (size_t k = 0; k < cnt; ++k)
                if (is_general(k)) return true;
            return false;
        };
        while (n > 1 && gm_growth_smem_bytes(G, (uint32_t)n, any_general(n)) > 200u * 1024u) --n;
        bool direct = false;
        if (gm_growth_smem_bytes(G, (uint32_t)n, any_general(n)) > 200u * 1024u) {  // very large G: no smem staging of deltas
            direct = true;
            n = std::min<size_t>(kMaxThresholds, ts.size() - i0);
            while (n > 1 && gm_growth_smem_bytes(G, (uint32_t)n, any_general(n), true) > 200u * 1024u) --n;
        }
        GmGrowthParams p;
        std::memset(&p, 0, sizeof(p));
        p.direct_out = direct ? 1u : 0u;
        p.gm = a->d_gm;
        p.gm_stride = a->gm_stride;
        p.n_words = (a->n_rows + 63u) / 64u;
        p.n_rows = a->n_rows;
        p.weight = weighted ? a->d_weight : nullptr;
        p.countable = a->d_countable;
        p.G = G;
        p.T = (uint32_t)n;
        p.weighted = weighted ? 1 : 0;
        p.out_order_stride = out_order_stride;
        p.col_fastest = gm_grid_col_fastest() ? 1u : 0u;
        for (size_t k = 0; k < n; ++k) {
            p.cov[k] = cov[ts[i0 + k]];
            p.slot[k] = ts[i0 + k];
            if (is_general(k)) p.general_mask |= 1u << k;
        }
        if (p.general_mask) {
            std::vector<uint32_t> packed(n * (size_t)G, 0u);
            for (size_t k = 0; k < n; ++k)
                if ((p.general_mask >> k) & 1u)
                    std::memcpy(packed.data() + k * G, thr + (size_t)ts[i0 + k] * G, (size_t)G * 4u);
            if ((rc = upload_thr(a, packed))) return rc;
            p.thr = a->d_thr;
        }
        const uint32_t kBatch = 4096;  // orders per launch (gridDim.y)
        for (uint32_t o0 = 0; o0 < n_orders; o0 += kBatch) {
            p.n_orders = std::min<uint32_t>(kBatch, n_orders - o0);
            p.order = d_orders + (size_t)o0 * G;
            p.out = d_out_base + (size_t)o0 * out_order_stride;
            if ((rc = launch_gm_growth(p, a->sm_count, a->stream))) return rc;
            a->launches++;
        }
        char buf[160];
        snprintf(buf, sizeof buf, "k_gm_growth orders=%u T=%zu general_mask=0x%x", n_orders, n, p.general_mask);
        a->last_launch = buf;
        i0 += n;
    }
    return PGX_OK;
}

// Growth on the group-major copy for the thresholds `ts` under n_orders device-resident orders (results as above),
// on k_gm_quorum: general thresholds (q > 0) up to kGmQuorumMaxT per launch, one q = 0 threshold riding along with the
// first of those launches, any further q = 0 thresholds in T = 0 launches of up to four.  bp-weighted runs use the
// weight-sorted group-major copy (ensure_planes).  Shapes whose tables do not fit in shared memory (very large G) and
// PGX_GM_QUORUM=old use the first-generation k_gm_growth for everything.
int gm_growth_launch(pgx_abacus *a, uint32_t n_orders, const uint32_t *d_orders, const std::vector<uint32_t> &ts,
                     const uint32_t *cov, const uint32_t *thr, int weighted, uint64_t *d_out_base, uint64_t out_order_stride) {
    const uint32_t G = a->G;
    const bool w = weighted != 0;
    std::vector<uint32_t> fast, gen;
    for (uint32_t t : ts) (all_zero(thr ? thr + (size_t)t * G : nullptr, G) ? fast : gen).push_back(t);
    const char *env = getenv("PGX_GM_QUORUM");
    const bool fits = gen.empty() ? gm_quorum_smem_bytes(G, 0, 1, w) <= kGmQuorumSmemMax
                                  : (gm_quorum_planes(G) > 0 && gm_quorum_smem_bytes(G, 1, 1, w) <= kGmQuorumSmemMax);
    if ((env && !strcmp(env, "old")) || !fits || ts.empty())
        return gm_growth_launch_legacy(a, n_orders, d_orders, ts, cov, thr, weighted, d_out_base, out_order_stride);
    int rc;
    const bool sorted = w && a->d_weight != nullptr;  // weight-sorted item order: uniform-weight columns are the rule
    if (sorted) {
        // bp sums with coverage cutoffs over several orders: rebuild the weight-sorted copy once with the coverage as the tie
        // order (it may have been derived before the coverages were known, e.g. by a weighted similarity call)
        bool any_cut = false;
        for (uint32_t t : gen) any_cut |= cov[t] > 1u;
        const char *cs = getenv("PGX_GM_COVSORT");
        if (a->planes_valid && !a->planes_cov_order && any_cut && a->countable_valid && n_orders >= 4u && !(cs && !strcmp(cs, "0")))
            a->planes_valid = false;
        if ((rc = ensure_planes(a))) return rc;
    }
    // counting, general thresholds with coverage cutoffs > 1, several orders: run on the coverage-sorted copy, where whole
    // warps hold only items below a cutoff and skip that threshold's rank comparison (or the ranks altogether).  One sort +
    // one permuted transpose per graph, so only when the work is repeated (PGX_GM_COVSORT=0 / 1: never / always).
    bool cov_sorted = false;
    if (!w && !gen.empty()) {
        bool any_cut = false;
        for (uint32_t t : gen) any_cut |= cov[t] > 1u;
        const char *cs = getenv("PGX_GM_COVSORT");
        cov_sorted = any_cut && !(cs && !strcmp(cs, "0")) && (n_orders >= 4u || a->gm_c_valid || (cs && !strcmp(cs, "1")));
        if (cov_sorted && (rc = ensure_gm_cov(a))) return rc;
    }
    GmGrowthParams base;
    std::memset(&base, 0, sizeof(base));
    base.gm = sorted ? a->d_gm_w : (cov_sorted ? a->d_gm_c : a->d_gm);
    base.gm_stride = a->gm_stride;
    base.n_words = (a->n_rows + 63u) / 64u;
    base.n_rows = a->n_rows;
    base.weight = sorted ? a->d_sorted_w : nullptr;
    base.perm = sorted ? a->d_perm : (cov_sorted ? a->d_perm_c : nullptr);
    base.uniform_w = sorted ? a->d_uniform_w : nullptr;
    base.countable = a->d_countable;
    base.G = G;
    base.weighted = w ? 1 : 0;
    base.out_order_stride = out_order_stride;
    base.col_fastest = gm_grid_col_fastest() ? 1u : 0u;
    auto run = [&](GmGrowthParams &p) -> int {
        KernelTimer kt(a);
        const uint32_t kBatch = 4096;  // orders per launch
        for (uint32_t o0 = 0; o0 < n_orders; o0 += kBatch) {
            p.n_orders = std::min<uint32_t>(kBatch, n_orders - o0);
            p.order = d_orders + (size_t)o0 * G;
            p.out = d_out_base + (size_t)o0 * out_order_stride;
            const int r = launch_gm_quorum(p, a->stream);
            if (r) return r;
            a->launches++;
        }
        char buf[192];
        if (p.T == 0)
            snprintf(buf, sizeof buf, "k_gm_union orders=%u T=0 q0=%u%s", n_orders, p.n_fast, sorted ? " weight-sorted" : (cov_sorted ? " coverage-sorted" : ""));
        else
            snprintf(buf, sizeof buf, "k_gm_quorum<P=%d> orders=%u T=%u q0=%u%s smem=%zu", p.T ? gm_quorum_planes(G) : 0, n_orders,
                     p.T, p.n_fast, sorted ? " weight-sorted" : (cov_sorted ? " coverage-sorted" : ""), gm_quorum_smem_bytes(G, p.T, gm_quorum_fast_slots(p.T, p.n_fast), w));
        a->last_launch = buf;
        return PGX_OK;
    };
    size_t f0 = 0;  // q = 0 thresholds already placed
    for (size_t i0 = 0; i0 < gen.size();) {
        size_t n = std::min<size_t>(kGmQuorumMaxT, gen.size() - i0);
        // two CTAs per SM when possible (<= 112 KB of shared memory each), otherwise whatever still fits
        while (n > 1 && gm_quorum_smem_bytes(G, (uint32_t)n, 1, w) > 112u * 1024u) --n;
        GmGrowthParams p = base;
        p.T = (uint32_t)n;
        p.general_mask = (1u << n) - 1u;
        std::vector<uint32_t> packed(n * (size_t)G);
        for (size_t k = 0; k < n; ++k) {
            p.cov[k] = cov[gen[i0 + k]];
            p.slot[k] = gen[i0 + k];
            std::memcpy(packed.data() + k * G, thr + (size_t)gen[i0 + k] * G, (size_t)G * 4u);
        }
        if (f0 < fast.size() && f0 == 0) {
            p.n_fast = 1u;
            p.cov[n] = cov[fast[0]];
            p.slot[n] = fast[0];
            f0 = 1;
        }
        if ((rc = upload_thr(a, packed))) return rc;
        p.thr = a->d_thr;
        if ((rc = run(p))) return rc;
        i0 += n;
    }
    while (f0 < fast.size()) {  // q = 0 thresholds that did not ride along: up to four per pass
        size_t n = std::min<size_t>(4, fast.size() - f0);
        while (n > 1 && gm_quorum_smem_bytes(G, 0, gm_quorum_fast_slots(0, (uint32_t)n), w) > kGmQuorumSmemMax) --n;
        GmGrowthParams p = base;
        p.T = 0;
        p.n_fast = (uint32_t)n;
        for (size_t k = 0; k < n; ++k) {
            p.cov[k] = cov[fast[f0 + k]];
            p.slot[k] = fast[f0 + k];
        }
        if ((rc = run(p))) return rc;
        f0 += n;
    }
    return PGX_OK;
}

// Growth under explicit host orders on the group-major copy: uploads the orders, runs the kernels and leaves the
// CURVES (prefix-summed on the device) of order o, threshold t at d_out + (o * T + t) * G.  No synchronisation.
int gm_growth_device(pgx_abacus *a, uint32_t n_orders, const uint32_t *orders, uint32_t T, const uint32_t *cov,
                     const uint32_t *thr, int weighted, uint64_t *d_out) {
    const uint32_t G = a->G;
    int rc = validate_thresholds(a, T, cov);
    if (rc) return rc;
    if (T == 0 || n_orders == 0) return PGX_OK;
    if (!orders) return fail(PGX_ERR_INVALID, "orders is null");
    for (uint32_t o = 0; o < n_orders; ++o)
        if (!is_permutation(orders + (size_t)o * G, G)) return fail(PGX_ERR_INVALID, "order is not a permutation of 0..G-1");
    bool need_cov = false;
    for (uint32_t t = 0; t < T; ++t) need_cov |= cov[t] > 1u;
    if (need_cov && (rc = ensure_countable(a))) return rc;
    if ((rc = ensure_gm(a))) return rc;
    if ((rc = ensure_dev(&a->d_order, &a->order_cap, (size_t)n_orders * G))) return rc;
    PGX_CUDA(cudaMemcpyAsync(a->d_order, orders, (size_t)n_orders * G * 4u, cudaMemcpyHostToDevice, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));  // `orders` is the caller's buffer

    const size_t out_words = (size_t)n_orders * T * G;
    PGX_CUDA(cudaMemsetAsync(d_out, 0, out_words * 8u, a->stream));
    std::vector<uint32_t> ts(T);
    for (uint32_t t = 0; t < T; ++t) ts[t] = t;
    if ((rc = gm_growth_launch(a, n_orders, a->d_order, ts, cov, thr, weighted, d_out, (uint64_t)T * G))) return rc;
    // first differences -> curves (wrapping u64 prefix sums) where they are
    if ((rc = launch_prefix_curves(d_out, (uint64_t)n_orders * T, G, a->stream))) return rc;
    a->launches++;
    return PGX_OK;
}

// ... -> curves in host memory
int gm_growth(pgx_abacus *a, uint32_t n_orders, const uint32_t *orders, uint32_t T, const uint32_t *cov,
              const uint32_t *thr, int weighted, uint64_t *curves /*host*/) {
    const size_t out_words = (size_t)n_orders * T * a->G;
    if (out_words == 0) return validate_thresholds(a, T, cov);
    int rc = ensure_dev(&a->d_scratch, &a->scratch_cap, out_words);
    if (rc) return rc;
    if ((rc = gm_growth_device(a, n_orders, orders, T, cov, thr, weighted, a->d_scratch))) return rc;
    return copy_to_host(a, curves, a->d_scratch, out_words);
}

// Integer part of Similarity::set_table for the group rows [row_begin, row_end) x the columns >= col_begin, into the
// zeroed device buffer d_inter ((row_end - row_begin) x G).  No synchronisation.
int sim_rows_device(pgx_abacus *a, int weighted, uint32_t row_begin, uint32_t row_end, uint32_t col_begin, uint64_t *d_inter,
                    bool upper_only, uint64_t word_begin, uint64_t word_end) {
    int rc;
    const bool use_planes = weighted && a->d_weight;
    if (use_planes) {
        if ((rc = ensure_planes(a))) return rc;  // weight-sorted group-major copy + planes
    } else if ((rc = ensure_gm(a))) {
        return rc;
    }
    if (row_end == row_begin) return PGX_OK;
    GmSimParams p;
    std::memset(&p, 0, sizeof(p));
    // items [64 word_begin, 64 word_end) only (a rank's share of a sharded run): every per-word array starts at word_begin
    const uint64_t all_words = (a->n_rows + 63u) / 64u;
    if (word_end > all_words) word_end = all_words;
    if (word_begin >= word_end) return PGX_OK;
    p.gm = (use_planes ? a->d_gm_w : a->d_gm) + word_begin;
    p.gm_stride = a->gm_stride;
    p.n_words = word_end - word_begin;
    p.planes = use_planes ? a->d_planes + word_begin : nullptr;
    p.uniform_w = use_planes ? a->d_uniform_w + word_begin : nullptr;
    p.plane_mask = use_planes ? a->d_plane_mask + word_begin : nullptr;
    p.n_planes = use_planes ? a->n_planes : 0;
    p.G = a->G;
    p.row_begin = row_begin;
    p.row_end = row_end;
    p.col_begin = col_begin;
    p.inter = d_inter;
    p.upper_only = upper_only ? 1u : 0u;
    // PGX_SIM: "mma" = tensor cores (tcgen05 kind::i8 on the expanded bits), "csa" = carry-save AND / POPC pairs,
    // "plain" = one POPC per item word and pair.  Default: mma for unweighted tables with >= 256 groups.
    const char *env = getenv("PGX_SIM");
    const bool want_mma = env ? !strcmp(env, "mma") : a->G >= 256u;  // (a 128 x 256 tile is mostly padding below that)
    p.csa = (!use_planes && !(env && !strcmp(env, "plain"))) ? 1u : 0u;
    if (!use_planes && want_mma) {
        p.triangular = (row_begin == 0u && row_end == a->G && col_begin == 0u) ? 1u : 0u;
        if (const char *dbg = getenv("PGX_SIM_DEBUG")) p.csa = (uint32_t)atoi(dbg);  // timing experiments only (see k_sim_mma)
        {
            KernelTimer kt(a);
            rc = launch_sim_mma(p, a->sm_count, a->stream);
            if (!rc && p.triangular) rc = launch_sim_mirror(d_inter, a->G, a->stream);
        }
        if (rc) return rc;
        a->launches++;
        a->last_launch = "k_sim_mma (tcgen05.mma kind::i8)";
        return PGX_OK;
    }
    {
        KernelTimer kt(a);
        rc = launch_gm_similarity(p, a->sm_count, a->stream);
    }
    if (rc) return rc;
    a->launches++;
    a->last_launch = use_planes ? "k_gm_similarity<weighted>" : p.csa ? "k_gm_similarity<csa>" : "k_gm_similarity<plain>";
    return PGX_OK;
}

// len[g] = sum of w over the items of group g, for the groups [g_begin, g_end) -> d_len[0 .. g_end - g_begin)
int sim_len_device(pgx_abacus *a, int weighted, uint32_t g_begin, uint32_t g_end, uint64_t *d_len) {
    int rc;
    const bool use_planes = weighted && a->d_weight;
    if (use_planes) {
        if ((rc = ensure_planes(a))) return rc;
    } else if ((rc = ensure_gm(a))) {
        return rc;
    }
    if (g_end == g_begin) return PGX_OK;
    const uint64_t *gm = (use_planes ? a->d_gm_w : a->d_gm) + (uint64_t)g_begin * a->gm_stride;
    if ((rc = launch_gm_rowsum(gm, a->gm_stride, (a->n_rows + 63u) / 64u, use_planes ? a->d_planes : nullptr,
                               use_planes ? a->n_planes : 0, use_planes ? a->d_uniform_w : nullptr, g_end - g_begin, d_len, a->stream)))
        return rc;
    a->launches++;
    return PGX_OK;
}

}  // namespace pgx

extern "C" {

const char *pgx_version(void) { return "panacus_b200 0.1.0 (sm_100a)"; }
const char *pgx_last_error(void) { return g_last_error.c_str(); }

int pgx_device_count(int *n) {
    if (!n) return fail(PGX_ERR_INVALID, "null pointer");
    *n = 0;
    PGX_CUDA(cudaGetDeviceCount(n));
    return PGX_OK;
}

int pgx_device_warmup(int device) {
    int ndev = 0;
    PGX_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PGX_ERR_INVALID, "no such CUDA device");
    DeviceGuard guard(device);
    PGX_CUDA(cudaFree(nullptr));  // creates the primary context (seconds on a cold process); idempotent
    return PGX_OK;
}

uint32_t pgx_row_words(uint32_t n_groups) {
    const uint32_t w = (n_groups + 63u) / 64u;
    return w <= 1u ? 1u : (w + 1u) / 2u * 2u;
}

size_t pgx_fused_out_words(uint32_t n_groups, uint32_t n_thresholds) {
    return 2u * ((size_t)n_groups + 1u) + (size_t)n_thresholds * n_groups;
}

int pgx_abacus_create(pgx_abacus **out, int device, uint64_t n_items, uint32_t n_groups) {
    if (!out) return fail(PGX_ERR_INVALID, "null out pointer");
    *out = nullptr;
    if (n_groups == 0 || n_groups > (1u << 20)) return fail(PGX_ERR_INVALID, "n_groups must be in 1..2^20");
    if (n_items >= 0xFFFFFFFFull - 1) return fail(PGX_ERR_UNSUPPORTED, "n_items must be < 2^32 - 2");
    int ndev = 0;
    PGX_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PGX_ERR_INVALID, "no such CUDA device");
    pgx_abacus *a = new (std::nothrow) pgx_abacus();
    if (!a) return fail(PGX_ERR_NOMEM, "out of host memory");
    a->device = device;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) a->sm_count = prop.multiProcessorCount;
    a->n_items = n_items;
    a->n_rows = n_items + 1;
    a->G = n_groups;
    a->W = (n_groups + 63u) / 64u;
    a->Wp = pgx_row_words(n_groups);
    a->acc_words = pgx_fused_out_words(n_groups, kMaxThresholds);
    auto bail = [&](int rc) {
        pgx_abacus_destroy(a);
        return rc;
    };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&a->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(fail(PGX_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)));
    a->stream = a->own_stream;
    const size_t bm_bytes = (size_t)a->n_rows * a->Wp * 8u;
    if ((e = cudaMalloc(reinterpret_cast<void **>(&a->d_bitmap), bm_bytes)) != cudaSuccess)
        return bail(fail(PGX_ERR_NOMEM, std::string("cudaMalloc bitmap: ") + cudaGetErrorString(e)));
    a->own_bitmap = true;
    if ((e = cudaMemsetAsync(a->d_bitmap, 0, bm_bytes, a->stream)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&a->d_acc), a->acc_words * 8u)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&a->d_ticket), 16)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&a->d_err), 8)) != cudaSuccess ||
        (e = cudaMemsetAsync(a->d_acc, 0, a->acc_words * 8u, a->stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(a->d_ticket, 0, 16, a->stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(a->d_err, 0, 8, a->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(a->stream)) != cudaSuccess)
        return bail(fail(PGX_ERR_CUDA, std::string("abacus setup: ") + cudaGetErrorString(e)));
    *out = a;
    return PGX_OK;
}

void pgx_abacus_destroy(pgx_abacus *a) {
    if (!a) return;
    DeviceGuard guard(a->device);
    if (a->stream) cudaStreamSynchronize(a->stream);
    pgx_exchange_disconnect(a);
    cudaFree(a->d_xchg);
    if (a->own_bitmap && a->d_bitmap) cudaFree(a->d_bitmap);
    if (a->own_weight && a->d_weight) cudaFree(a->d_weight);
    cudaFree(a->d_countable);
    cudaFree(a->d_hist_tmp);
    cudaFree(a->d_csr_r);
    cudaFree(a->d_gm);
    cudaFree(a->d_planes);
    cudaFree(a->d_gm_w);
    cudaFree(a->d_gm_c);
    cudaFree(a->d_perm_c);
    cudaFree(a->d_perm);
    cudaFree(a->d_sorted_w);
    cudaFree(a->d_uniform_w);
    cudaFree(a->d_plane_mask);
    cudaFree(a->d_acc);
    cudaFree(a->d_ticket);
    cudaFree(a->d_err);
    cudaFree(a->d_thr);
    cudaFree(a->d_order);
    cudaFree(a->d_identity);
    cudaFree(a->d_scratch);
    if (a->h_pinned) cudaFreeHost(a->h_pinned);
    for (int b = 0; b < 2; ++b) {
        cudaFree(a->d_stage[b]);
        if (a->h_stage[b]) cudaFreeHost(a->h_stage[b]);
        if (a->ev_ready[b]) cudaEventDestroy(a->ev_ready[b]);
        if (a->ev_free[b]) cudaEventDestroy(a->ev_free[b]);
    }
    cudaFree(a->d_prefsum);
    cudaFree(a->d_path_group);
    cudaFree(a->d_exclude);
    for (auto &pr : a->ev_pool) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    if (a->copy_stream) cudaStreamDestroy(a->copy_stream);
    if (a->own_stream) cudaStreamDestroy(a->own_stream);
    cudaGetLastError();
    delete a;
}

int pgx_abacus_set_stream(pgx_abacus *a, void *cuda_stream) {
    if (!a) return fail(PGX_ERR_INVALID, "null handle");
    DeviceGuard guard(a->device);
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    a->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : a->own_stream;
    return PGX_OK;
}

int pgx_abacus_shape(const pgx_abacus *a, uint64_t *n_items, uint32_t *n_groups, uint32_t *row_words) {
    if (!a) return fail(PGX_ERR_INVALID, "null handle");
    if (n_items) *n_items = a->n_items;
    if (n_groups) *n_groups = a->G;
    if (row_words) *row_words = a->Wp;
    return PGX_OK;
}

int pgx_abacus_upload(pgx_abacus *a, const uint64_t *bitmap, uint32_t host_row_words, const uint32_t *weight) {
    if (!a) return fail(PGX_ERR_INVALID, "null handle");
    DeviceGuard guard(a->device);
    if (bitmap) {
        if (host_row_words < a->W) return fail(PGX_ERR_INVALID, "host_row_words < ceil(n_groups/64)");
        if (!a->own_bitmap) {
            a->d_bitmap = nullptr;
            PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_bitmap), (size_t)a->n_rows * a->Wp * 8u));
            a->own_bitmap = true;
        }
        if (host_row_words != a->Wp)
            PGX_CUDA(cudaMemsetAsync(a->d_bitmap, 0, (size_t)a->n_rows * a->Wp * 8u, a->stream));
        const size_t copy_w = std::min<uint32_t>(host_row_words, a->Wp);
        PGX_CUDA(cudaMemcpy2DAsync(a->d_bitmap, (size_t)a->Wp * 8u, bitmap, (size_t)host_row_words * 8u, copy_w * 8u,
                                   a->n_rows, cudaMemcpyHostToDevice, a->stream));
        invalidate_derived(a);
    }
    if (weight) {
        if (!a->own_weight || !a->d_weight) {
            a->d_weight = nullptr;
            PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_weight), a->n_rows * 4u));
            a->own_weight = true;
        }
        PGX_CUDA(cudaMemcpyAsync(a->d_weight, weight, a->n_rows * 4u, cudaMemcpyHostToDevice, a->stream));
        uint32_t m = 0;
        for (uint64_t i = 1; i < a->n_rows; ++i) m = std::max(m, weight[i]);
        a->max_weight = m;
        a->max_weight_known = true;
        a->planes_valid = false;
    }
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    return PGX_OK;
}

int pgx_abacus_adopt_device(pgx_abacus *a, uint64_t *d_bitmap, uint32_t *d_weight) {
    if (!a) return fail(PGX_ERR_INVALID, "null handle");
    if (!d_bitmap) return fail(PGX_ERR_INVALID, "d_bitmap is null");
    if ((reinterpret_cast<uintptr_t>(d_bitmap) & 15u) || (reinterpret_cast<uintptr_t>(d_weight) & 15u))
        return fail(PGX_ERR_INVALID, "device buffers must be 16-byte aligned");
    DeviceGuard guard(a->device);
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    if (a->own_bitmap && a->d_bitmap) cudaFree(a->d_bitmap);
    if (a->own_weight && a->d_weight) cudaFree(a->d_weight);
    a->d_bitmap = d_bitmap;
    a->own_bitmap = false;
    a->d_weight = d_weight;
    a->own_weight = false;
    a->max_weight_known = d_weight == nullptr;
    a->max_weight = 1;
    invalidate_derived(a);
    return PGX_OK;
}

int pgx_abacus_copy_rows(pgx_abacus *dst, pgx_abacus *src, uint64_t src_first_item) {
    int rc = check_handle(dst);
    if (rc || (rc = check_handle(src))) return rc;
    if (dst == src) return fail(PGX_ERR_INVALID, "source and destination are the same handle");
    if (dst->G != src->G) return fail(PGX_ERR_INVALID, "handles differ in n_groups");
    if (src_first_item == 0 || src_first_item + dst->n_items > src->n_items + 1u)
        return fail(PGX_ERR_INVALID, "item range outside the source abacus");
    {
        DeviceGuard g(src->device);
        PGX_CUDA(cudaStreamSynchronize(src->stream));  // the source rows must be complete
    }
    DeviceGuard guard(dst->device);
    const size_t row_bytes = (size_t)dst->Wp * 8u;
    // rows 1..n of dst <- rows first..first+n-1 of src; row 0 stays the (zero) dummy item
    PGX_CUDA(cudaMemcpyPeerAsync(dst->d_bitmap + dst->Wp, dst->device, src->d_bitmap + src_first_item * src->Wp, src->device,
                                 dst->n_items * row_bytes, dst->stream));
    PGX_CUDA(cudaMemsetAsync(dst->d_bitmap, 0, row_bytes, dst->stream));
    if (src->d_weight) {
        if (!dst->own_weight || !dst->d_weight) {
            dst->d_weight = nullptr;
            PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&dst->d_weight), dst->n_rows * 4u));
            dst->own_weight = true;
        }
        PGX_CUDA(cudaMemcpyPeerAsync(dst->d_weight + 1, dst->device, src->d_weight + src_first_item, src->device,
                                     dst->n_items * 4u, dst->stream));
        PGX_CUDA(cudaMemsetAsync(dst->d_weight, 0, 4u, dst->stream));
        dst->max_weight = src->max_weight;
        dst->max_weight_known = src->max_weight_known;
    }
    PGX_CUDA(cudaStreamSynchronize(dst->stream));
    invalidate_derived(dst);
    return PGX_OK;
}

int pgx_abacus_clear(pgx_abacus *a) {
    int rc = check_handle(a);
    if (rc) return rc;
    DeviceGuard guard(a->device);
    PGX_CUDA(cudaMemsetAsync(a->d_bitmap, 0, (size_t)a->n_rows * a->Wp * 8u, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    invalidate_derived(a);
    return PGX_OK;
}

int pgx_abacus_scatter(pgx_abacus *a, const uint64_t *items, uint64_t n_steps, uint32_t group_id,
                       const uint8_t *exclude) {
    int rc = check_handle(a);
    if (rc) return rc;
    if (group_id >= a->G) return fail(PGX_ERR_INVALID, "group_id >= n_groups");
    if (n_steps && !items) return fail(PGX_ERR_INVALID, "items is null");
    DeviceGuard guard(a->device);
    struct Staging {  // freed on every exit path
        uint64_t *items = nullptr;
        uint8_t *ex = nullptr;
        ~Staging() {
            cudaFree(items);
            cudaFree(ex);
        }
    } st;
    if (exclude) {
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&st.ex), a->n_rows));
        PGX_CUDA(cudaMemcpyAsync(st.ex, exclude, a->n_rows, cudaMemcpyHostToDevice, a->stream));
    }
    const uint64_t kChunk = 1ull << 24;  // 16 Mi steps = 128 MB per staging copy
    PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&st.items), std::min<uint64_t>(kChunk, std::max<uint64_t>(n_steps, 1)) * 8u));
    invalidate_derived(a);
    for (uint64_t s0 = 0; s0 < n_steps; s0 += kChunk) {
        const uint64_t n = std::min<uint64_t>(kChunk, n_steps - s0);
        PGX_CUDA(cudaMemcpyAsync(st.items, items + s0, n * 8u, cudaMemcpyHostToDevice, a->stream));
        if ((rc = launch_scatter(a->d_bitmap, a->Wp, a->n_rows, st.items, n, group_id, st.ex, a->d_err, a->stream))) return rc;
        a->launches++;
        PGX_CUDA(cudaStreamSynchronize(a->stream));  // the staging buffer is reused by the next chunk
    }
    unsigned int err = 0;
    PGX_CUDA(cudaMemcpyAsync(&err, a->d_err, 4, cudaMemcpyDeviceToHost, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    if (err) {
        PGX_CUDA(cudaMemsetAsync(a->d_err, 0, 4, a->stream));
        return fail(PGX_ERR_INVALID, "item id out of range 1..=n_items in scatter");
    }
    return PGX_OK;
}

namespace {

bool is_pinned_host(const void *p) {
    cudaPointerAttributes at;
    const bool pinned = cudaPointerGetAttributes(&at, p) == cudaSuccess && (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    return pinned;
}

// dst[i] = (u32)src[i] on up to 8 host threads (memory bound: 8 B read + 4 B written per step)
void narrow_ids(uint32_t *dst, const uint64_t *src, size_t n) {
    unsigned nt = std::thread::hardware_concurrency();
    nt = std::max(1u, std::min(8u, nt));
    if (n < (1u << 16)) nt = 1;
    auto work = [=](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) dst[i] = (uint32_t)src[i];
    };
    if (nt == 1) return work(0, n);
    std::vector<std::thread> th;
    const size_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        const size_t lo = std::min(n, t * per), hi = std::min(n, lo + per);
        if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto &x : th) x.join();
}

int ensure_build_pipeline(pgx_abacus *a, size_t stage_bytes, bool need_host_stage) {
    if (!a->copy_stream) {
        PGX_CUDA(cudaStreamCreateWithFlags(&a->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            PGX_CUDA(cudaEventCreateWithFlags(&a->ev_ready[b], cudaEventDisableTiming));
            PGX_CUDA(cudaEventCreateWithFlags(&a->ev_free[b], cudaEventDisableTiming));
        }
    }
    if (a->d_stage_bytes < stage_bytes) {
        for (int b = 0; b < 2; ++b) {
            cudaFree(a->d_stage[b]);
            a->d_stage[b] = nullptr;
        }
        a->d_stage_bytes = 0;
        for (int b = 0; b < 2; ++b) PGX_CUDA(cudaMalloc(&a->d_stage[b], stage_bytes));
        a->d_stage_bytes = stage_bytes;
    }
    if (need_host_stage && a->h_stage_bytes < stage_bytes) {
        for (int b = 0; b < 2; ++b) {
            if (a->h_stage[b]) cudaFreeHost(a->h_stage[b]);
            a->h_stage[b] = nullptr;
        }
        a->h_stage_bytes = 0;
        for (int b = 0; b < 2; ++b) PGX_CUDA(cudaMallocHost(&a->h_stage[b], stage_bytes));
        a->h_stage_bytes = stage_bytes;
    }
    return PGX_OK;
}

// ItemTable -> bitmap.  The table is streamed in chunks through two device staging buffers: the copy stream uploads chunk
// k + 1 while k_build scatters chunk k.  A page-locked caller buffer is read by DMA where it lies (u32 ids: 4 bytes per
// step over PCIe, u64: 8); a pageable one goes through two pinned host buffers, and its u64 ids (always < 2^32: n_items is)
// are narrowed to u32 on the way by host threads, which halves the PCIe bytes.
int build_pipeline(pgx_abacus *a, const void *items, int id_bytes, uint64_t n_steps, const uint64_t *id_prefsum, uint64_t n_paths,
                   const int64_t *path_group, const uint8_t *exclude) {
    int rc = check_handle(a);
    if (rc) return rc;
    if (!id_prefsum || !path_group || (n_steps && !items)) return fail(PGX_ERR_INVALID, "null table pointer");
    if (n_paths == 0) return n_steps ? fail(PGX_ERR_INVALID, "steps without paths") : (int)PGX_OK;
    if (id_prefsum[0] != 0 || id_prefsum[n_paths] != n_steps) return fail(PGX_ERR_INVALID, "id_prefsum does not span the items");
    for (uint64_t p = 0; p < n_paths; ++p)
        if (id_prefsum[p] > id_prefsum[p + 1]) return fail(PGX_ERR_INVALID, "id_prefsum is not monotone");
    DeviceGuard guard(a->device);
    const bool pinned = n_steps == 0 || is_pinned_host(items);
    const int wire_bytes = pinned ? id_bytes : 4;  // pageable tables are narrowed while they are staged
    const uint64_t kChunk = 1ull << 24;            // 16 Mi steps per chunk: 64 MB of u32 ids, ~1.2 ms of PCIe time
    const uint64_t chunk = std::max<uint64_t>(std::min(kChunk, n_steps), 1);
    if ((rc = ensure_build_pipeline(a, chunk * (size_t)wire_bytes, !pinned))) return rc;
    if ((rc = ensure_dev(&a->d_prefsum, &a->prefsum_cap, n_paths + 1))) return rc;
    if ((rc = ensure_dev(&a->d_path_group, &a->path_group_cap, n_paths))) return rc;
    PGX_CUDA(cudaMemcpyAsync(a->d_prefsum, id_prefsum, (n_paths + 1) * 8u, cudaMemcpyHostToDevice, a->stream));
    PGX_CUDA(cudaMemcpyAsync(a->d_path_group, path_group, n_paths * 8u, cudaMemcpyHostToDevice, a->stream));
    const uint8_t *d_ex = nullptr;
    if (exclude) {
        if ((rc = ensure_dev(&a->d_exclude, &a->exclude_cap, a->n_rows))) return rc;
        PGX_CUDA(cudaMemcpyAsync(a->d_exclude, exclude, a->n_rows, cudaMemcpyHostToDevice, a->stream));
        d_ex = a->d_exclude;
    }
    invalidate_derived(a);
    uint64_t k = 0;
    for (uint64_t s0 = 0; s0 < n_steps; s0 += chunk, ++k) {
        const uint64_t n = std::min<uint64_t>(chunk, n_steps - s0);
        const int b = (int)(k & 1u);
        const void *src = static_cast<const unsigned char *>(items) + s0 * (size_t)id_bytes;
        if (!pinned) {
            PGX_CUDA(cudaEventSynchronize(a->ev_ready[b]));  // the upload that last read h_stage[b] is done
            if (id_bytes == 8)
                narrow_ids(static_cast<uint32_t *>(a->h_stage[b]), static_cast<const uint64_t *>(src), n);
            else
                std::memcpy(a->h_stage[b], src, n * 4u);
            src = a->h_stage[b];
        }
        PGX_CUDA(cudaStreamWaitEvent(a->copy_stream, a->ev_free[b], 0));  // k_build is done with d_stage[b]
        PGX_CUDA(cudaMemcpyAsync(a->d_stage[b], src, n * (size_t)wire_bytes, cudaMemcpyHostToDevice, a->copy_stream));
        PGX_CUDA(cudaEventRecord(a->ev_ready[b], a->copy_stream));
        PGX_CUDA(cudaStreamWaitEvent(a->stream, a->ev_ready[b], 0));
        if ((rc = launch_build(a->d_bitmap, a->Wp, a->n_rows, a->G, a->d_stage[b], wire_bytes, s0, n, a->d_prefsum, n_paths,
                               a->d_path_group, d_ex, a->d_err, a->stream)))
            return rc;
        PGX_CUDA(cudaEventRecord(a->ev_free[b], a->stream));
        a->launches++;
    }
    unsigned int err = 0;
    PGX_CUDA(cudaMemcpyAsync(&err, a->d_err, 4, cudaMemcpyDeviceToHost, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    if (err) {
        PGX_CUDA(cudaMemsetAsync(a->d_err, 0, 4, a->stream));
        return fail(PGX_ERR_INVALID, (err & 1u) ? "item id out of range 1..=n_items in the ItemTable" : "path_group entry >= n_groups");
    }
    char buf[160];
    snprintf(buf, sizeof buf, "k_build<u%d> chunks=%llu source=%s", wire_bytes * 8, (unsigned long long)k,
             pinned ? "pinned (direct DMA)" : "pageable (staged, narrowed to u32)");
    a->last_launch = buf;
    return PGX_OK;
}

}  // namespace

int pgx_abacus_build(pgx_abacus *a, const uint64_t *items, uint64_t n_steps, const uint64_t *id_prefsum,
                     uint64_t n_paths, const int64_t *path_group, const uint8_t *exclude) {
    return build_pipeline(a, items, 8, n_steps, id_prefsum, n_paths, path_group, exclude);
}

int pgx_abacus_build_u32(pgx_abacus *a, const uint32_t *items, uint64_t n_steps, const uint64_t *id_prefsum,
                         uint64_t n_paths, const int64_t *path_group, const uint8_t *exclude) {
    return build_pipeline(a, items, 4, n_steps, id_prefsum, n_paths, path_group, exclude);
}

int pgx_host_alloc(void **out, size_t bytes) {
    if (!out) return fail(PGX_ERR_INVALID, "null pointer");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(PGX_ERR_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    }
    return PGX_OK;
}

void pgx_host_free(void *p) {
    if (p) cudaFreeHost(p);
    cudaGetLastError();
}

int pgx_abacus_csr_rows(pgx_abacus *a, uint64_t *r, uint64_t *nnz) {
    int rc = check_handle(a);
    if (rc) return rc;
    DeviceGuard guard(a->device);
    if (!a->csr_valid) {
        if (!a->d_csr_r) PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_csr_r), (a->n_rows + 1u) * 8u));
        if ((rc = launch_csr_rows(a->d_bitmap, a->n_rows, a->G, a->W, a->Wp, a->d_csr_r, a->stream))) return rc;
        a->launches += 2;
        a->csr_valid = true;
    }
    if (r) PGX_CUDA(cudaMemcpyAsync(r, a->d_csr_r, (a->n_rows + 1u) * 8u, cudaMemcpyDeviceToHost, a->stream));
    uint64_t total = 0;
    PGX_CUDA(cudaMemcpyAsync(&total, a->d_csr_r + a->n_rows, 8u, cudaMemcpyDeviceToHost, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    if (nnz) *nnz = total;
    a->last_launch = "k_csr_row_len + scan";
    return PGX_OK;
}

int pgx_abacus_csr_fill(pgx_abacus *a, const uint64_t *items, uint64_t n_steps, const uint64_t *id_prefsum, uint64_t n_paths,
                        const int64_t *path_group, const uint8_t *exclude, uint64_t *c, uint32_t *v) {
    int rc = check_handle(a);
    if (rc) return rc;
    if (v && (!id_prefsum || !path_group || (n_steps && !items))) return fail(PGX_ERR_INVALID, "null table pointer");
    if (v && n_paths == 0 && n_steps) return fail(PGX_ERR_INVALID, "steps without paths");
    if (v && n_paths && (id_prefsum[0] != 0 || id_prefsum[n_paths] != n_steps))
        return fail(PGX_ERR_INVALID, "id_prefsum does not span the items");
    uint64_t nnz = 0;
    if ((rc = pgx_abacus_csr_rows(a, nullptr, &nnz))) return rc;
    DeviceGuard guard(a->device);
    struct Staging {  // freed on every exit path
        uint64_t *c = nullptr, *items = nullptr, *prefsum = nullptr;
        uint32_t *v = nullptr;
        int64_t *group = nullptr;
        uint8_t *ex = nullptr;
        ~Staging() {
            cudaFree(c);
            cudaFree(items);
            cudaFree(prefsum);
            cudaFree(v);
            cudaFree(group);
            cudaFree(ex);
        }
    } st;
    if (c && nnz) {
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&st.c), nnz * 8u));
        if ((rc = launch_csr_cols(a->d_bitmap, a->n_rows, a->G, a->W, a->Wp, a->d_csr_r, st.c, a->stream))) return rc;
        a->launches++;
        PGX_CUDA(cudaMemcpyAsync(c, st.c, nnz * 8u, cudaMemcpyDeviceToHost, a->stream));
        PGX_CUDA(cudaStreamSynchronize(a->stream));
    }
    if (v && nnz && n_paths) {
        const uint64_t kChunk = 1ull << 25;  // 32 Mi steps = 256 MB per staging copy
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&st.v), nnz * 4u));
        PGX_CUDA(cudaMemsetAsync(st.v, 0, nnz * 4u, a->stream));
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&st.prefsum), (n_paths + 1) * 8u));
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&st.group), n_paths * 8u));
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&st.items), std::max<uint64_t>(std::min(kChunk, n_steps), 1) * 8u));
        PGX_CUDA(cudaMemcpyAsync(st.prefsum, id_prefsum, (n_paths + 1) * 8u, cudaMemcpyHostToDevice, a->stream));
        PGX_CUDA(cudaMemcpyAsync(st.group, path_group, n_paths * 8u, cudaMemcpyHostToDevice, a->stream));
        if (exclude) {
            PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&st.ex), a->n_rows));
            PGX_CUDA(cudaMemcpyAsync(st.ex, exclude, a->n_rows, cudaMemcpyHostToDevice, a->stream));
        }
        for (uint64_t s0 = 0; s0 < n_steps; s0 += kChunk) {
            const uint64_t n = std::min<uint64_t>(kChunk, n_steps - s0);
            PGX_CUDA(cudaMemcpyAsync(st.items, items + s0, n * 8u, cudaMemcpyHostToDevice, a->stream));
            if ((rc = launch_csr_vals(a->d_bitmap, a->n_rows, a->G, a->Wp, st.items, s0, n, st.prefsum, n_paths, st.group, st.ex,
                                      a->d_csr_r, st.v, a->d_err, a->stream)))
                return rc;
            a->launches++;
            PGX_CUDA(cudaStreamSynchronize(a->stream));  // the staging buffer is reused by the next chunk
        }
        unsigned int err = 0;
        PGX_CUDA(cudaMemcpyAsync(&err, a->d_err, 4, cudaMemcpyDeviceToHost, a->stream));
        PGX_CUDA(cudaMemcpyAsync(v, st.v, nnz * 4u, cudaMemcpyDeviceToHost, a->stream));
        PGX_CUDA(cudaStreamSynchronize(a->stream));
        if (err) {
            PGX_CUDA(cudaMemsetAsync(a->d_err, 0, 4, a->stream));
            return fail(PGX_ERR_INVALID, (err & 8u)   ? "ItemTable step without its bit in the bitmap: the abacus was not built from this table"
                                         : (err & 1u) ? "item id out of range 1..=n_items in the ItemTable"
                                                      : "path_group entry >= n_groups");
        }
    }
    a->last_launch = "k_csr_cols + k_csr_vals";
    return PGX_OK;
}

int pgx_abacus_download(pgx_abacus *a, uint64_t *bitmap, uint32_t host_row_words) {
    int rc = check_handle(a);
    if (rc) return rc;
    if (!bitmap || host_row_words < a->W) return fail(PGX_ERR_INVALID, "bad download buffer");
    DeviceGuard guard(a->device);
    if (host_row_words > a->Wp)
        for (uint64_t i = 0; i < a->n_rows; ++i)
            std::memset(bitmap + i * host_row_words + a->Wp, 0, (size_t)(host_row_words - a->Wp) * 8u);
    const size_t copy_w = std::min<uint32_t>(host_row_words, a->Wp);
    PGX_CUDA(cudaMemcpy2DAsync(bitmap, (size_t)host_row_words * 8u, a->d_bitmap, (size_t)a->Wp * 8u, copy_w * 8u,
                               a->n_rows, cudaMemcpyDeviceToHost, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    return PGX_OK;
}

int pgx_hist_ordered_growth(pgx_abacus *a, uint64_t *hist_count, uint64_t *hist_weight, uint32_t n_thresholds,
                            const uint32_t *cov_abs, const uint32_t *quorum_thr, int weighted, uint64_t *curve) {
    int rc = check_handle(a);
    if (rc) return rc;
    if (n_thresholds && !curve) return fail(PGX_ERR_INVALID, "curve is null");
    DeviceGuard guard(a->device);
    const uint32_t G = a->G, G1 = G + 1u;
    const size_t words = pgx_fused_out_words(G, n_thresholds);
    if ((rc = ensure_dev(&a->d_scratch, &a->scratch_cap, words))) return rc;
    if ((rc = ensure_pinned(a, words))) return rc;
    rc = fused_pass(a, hist_count != nullptr, hist_weight != nullptr, n_thresholds, cov_abs, quorum_thr, weighted, nullptr,
                    a->d_scratch);
    if (rc) return rc;
    PGX_CUDA(cudaMemcpyAsync(a->h_pinned, a->d_scratch, words * 8u, cudaMemcpyDeviceToHost, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    if ((rc = check_exchange(a))) return rc;
    if (hist_count) std::memcpy(hist_count, a->h_pinned, (size_t)G1 * 8u);
    if (hist_weight) std::memcpy(hist_weight, a->h_pinned + G1, (size_t)G1 * 8u);
    for (uint32_t t = 0; t < n_thresholds; ++t) {
        uint64_t run = 0;  // first differences -> curve (two's-complement wrapping sums)
        const uint64_t *d = a->h_pinned + 2u * G1 + (size_t)t * G;
        for (uint32_t j = 0; j < G; ++j) {
            run += d[j];
            curve[(size_t)t * G + j] = run;
        }
    }
    return PGX_OK;
}

int pgx_hist(pgx_abacus *a, uint64_t *hist_count, uint64_t *hist_weight, uint32_t *countable) {
    int rc = check_handle(a);
    if (rc) return rc;
    DeviceGuard guard(a->device);
    const uint32_t G1 = a->G + 1u;
    const size_t words = pgx_fused_out_words(a->G, 0);
    if ((rc = ensure_dev(&a->d_scratch, &a->scratch_cap, words))) return rc;
    if ((rc = ensure_pinned(a, words))) return rc;
    if (countable && !a->d_countable)
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_countable), a->n_rows * 4u));
    rc = fused_pass(a, hist_count != nullptr || (!hist_weight && !countable), hist_weight != nullptr, 0, nullptr, nullptr, 0,
                    countable ? a->d_countable : nullptr, a->d_scratch);
    if (rc) return rc;
    PGX_CUDA(cudaMemcpyAsync(a->h_pinned, a->d_scratch, words * 8u, cudaMemcpyDeviceToHost, a->stream));
    if (countable) {
        PGX_CUDA(cudaMemcpyAsync(countable, a->d_countable, a->n_rows * 4u, cudaMemcpyDeviceToHost, a->stream));
        a->countable_valid = true;
    }
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    if ((rc = check_exchange(a))) return rc;
    if (hist_count) std::memcpy(hist_count, a->h_pinned, (size_t)G1 * 8u);
    if (hist_weight) std::memcpy(hist_weight, a->h_pinned + G1, (size_t)G1 * 8u);
    return PGX_OK;
}

int pgx_ordered_growth(pgx_abacus *a, uint32_t n_thresholds, const uint32_t *cov_abs, const uint32_t *quorum_thr,
                       const uint32_t *col_order, int weighted, uint64_t *curve) {
    int rc = check_handle(a);
    if (rc) return rc;
    if (n_thresholds && !curve) return fail(PGX_ERR_INVALID, "curve is null");
    if (!col_order) return pgx_hist_ordered_growth(a, nullptr, nullptr, n_thresholds, cov_abs, quorum_thr, weighted, curve);
    DeviceGuard guard(a->device);
    return gm_growth(a, 1, col_order, n_thresholds, cov_abs, quorum_thr, weighted, curve);
}

int pgx_permuted_growth(pgx_abacus *a, uint32_t n_orders, const uint32_t *orders, uint32_t n_thresholds,
                        const uint32_t *cov_abs, const uint32_t *quorum_thr, int weighted, uint64_t *curves) {
    int rc = check_handle(a);
    if (rc) return rc;
    if (n_thresholds && n_orders && !curves) return fail(PGX_ERR_INVALID, "curves is null");
    DeviceGuard guard(a->device);
    return gm_growth(a, n_orders, orders, n_thresholds, cov_abs, quorum_thr, weighted, curves);
}

namespace {
int similarity_rows(pgx_abacus *a, int weighted, uint32_t row_begin, uint32_t row_end, uint32_t col_begin, uint64_t *inter,
                    uint64_t *len);
}

int pgx_similarity(pgx_abacus *a, int weighted, uint32_t row_begin, uint32_t row_end, uint64_t *inter, uint64_t *len) {
    return similarity_rows(a, weighted, row_begin, row_end, 0u, inter, len);
}

int pgx_similarity_upper(pgx_abacus *a, int weighted, uint32_t row_begin, uint32_t row_end, uint64_t *inter, uint64_t *len) {
    return similarity_rows(a, weighted, row_begin, row_end, row_begin, inter, len);
}

namespace {
int similarity_rows(pgx_abacus *a, int weighted, uint32_t row_begin, uint32_t row_end, uint32_t col_begin, uint64_t *inter,
                    uint64_t *len) {
    int rc = check_handle(a);
    if (rc) return rc;
    if (row_begin > row_end || row_end > a->G) return fail(PGX_ERR_INVALID, "bad row range");
    DeviceGuard guard(a->device);
    const uint32_t G = a->G;
    const uint32_t rows = inter ? row_end - row_begin : 0u;
    const size_t inter_words = (size_t)rows * G;
    const size_t words = inter_words + G;
    if ((rc = ensure_dev(&a->d_scratch, &a->scratch_cap, words))) return rc;
    PGX_CUDA(cudaMemsetAsync(a->d_scratch, 0, words * 8u, a->stream));
    if (rows && (rc = sim_rows_device(a, weighted, row_begin, row_end, col_begin, a->d_scratch, false, 0, ~0ull))) return rc;
    // len[g] = the diagonal entry inter[g][g]: the full square carries it, row blocks need the row sums of all G groups
    const bool full_square = rows == G && row_begin == 0u && col_begin == 0u;
    if (len && !full_square && (rc = sim_len_device(a, weighted, 0, G, a->d_scratch + inter_words))) return rc;
    if (rows && (rc = copy_to_host(a, inter, a->d_scratch, inter_words))) return rc;
    if (len && full_square) {
        for (uint32_t g = 0; g < G; ++g) len[g] = inter[(size_t)g * G + g];
    } else if (len && (rc = copy_to_host(a, len, a->d_scratch + inter_words, G))) {
        return rc;
    }
    return PGX_OK;
}
}  // namespace

int pgx_fused_pass_async(pgx_abacus *a, int want_hist_count, int want_hist_weight, uint32_t n_thresholds,
                         const uint32_t *cov_abs, const uint32_t *quorum_thr, int weighted, uint64_t *d_out) {
    int rc = check_handle(a);
    if (rc) return rc;
    if (!d_out) return fail(PGX_ERR_INVALID, "d_out is null");
    DeviceGuard guard(a->device);
    return fused_pass(a, want_hist_count != 0, want_hist_weight != 0, n_thresholds, cov_abs, quorum_thr, weighted, nullptr,
                      d_out);
}

static size_t xchg_data_bytes(const pgx_abacus *a) { return (size_t)2 * kMaxRanks * a->acc_words * 16u; }  // 2 packets per word

int pgx_exchange_export(pgx_abacus *a, void *handle_out) {
    if (!a || !handle_out) return fail(PGX_ERR_INVALID, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == PGX_EXCHANGE_HANDLE_BYTES, "handle size");
    DeviceGuard guard(a->device);
    if (!a->d_xchg) {
        const size_t bytes = xchg_data_bytes(a);
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_xchg), bytes));
        PGX_CUDA(cudaMemset(a->d_xchg, 0, bytes));
    }
    cudaIpcMemHandle_t h;
    PGX_CUDA(cudaIpcGetMemHandle(&h, a->d_xchg));
    std::memcpy(handle_out, &h, sizeof(h));
    return PGX_OK;
}

int pgx_exchange_connect(pgx_abacus *a, uint32_t rank, uint32_t world, const void *all_handles) {
    if (!a || !all_handles) return fail(PGX_ERR_INVALID, "bad arguments");
    if (world < 1 || world > (uint32_t)kMaxRanks || rank >= world) return fail(PGX_ERR_INVALID, "bad rank / world (<= 8)");
    if (!a->d_xchg) return fail(PGX_ERR_STATE, "call pgx_exchange_export first");
    DeviceGuard guard(a->device);
    pgx_exchange_disconnect(a);
    Exchange x = {};
    x.world = world;
    x.rank = rank;
    x.stride = (uint32_t)a->acc_words;
    x.err = a->d_err + 1;
    for (uint32_t r = 0; r < world; ++r) {
        unsigned char *base = nullptr;
        if (r == rank) {
            base = a->d_xchg;
        } else {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, static_cast<const unsigned char *>(all_handles) + (size_t)r * sizeof(h), sizeof(h));
            void *ptr = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                pgx_exchange_disconnect(a);
                return fail(PGX_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
            }
            a->peer_base[r] = ptr;
            base = static_cast<unsigned char *>(ptr);
        }
        x.data[r] = reinterpret_cast<uint64_t *>(base);
    }
    a->x = x;
    return PGX_OK;
}

int pgx_exchange_disconnect(pgx_abacus *a) {
    if (!a) return fail(PGX_ERR_INVALID, "null handle");
    DeviceGuard guard(a->device);
    if (a->stream) cudaStreamSynchronize(a->stream);
    for (int r = 0; r < kMaxRanks; ++r) {
        if (a->peer_base[r]) cudaIpcCloseMemHandle(a->peer_base[r]);
        a->peer_base[r] = nullptr;
    }
    a->x = Exchange{};
    cudaGetLastError();
    return PGX_OK;
}

int pgx_exchange_status(pgx_abacus *a) {
    if (!a) return fail(PGX_ERR_INVALID, "null handle");
    DeviceGuard guard(a->device);
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    return check_exchange(a);
}

int pgx_abacus_set_timing(pgx_abacus *a, int enable) {
    if (!a) return fail(PGX_ERR_INVALID, "null handle");
    a->timing = enable != 0;
    a->ev_used = 0;
    return PGX_OK;
}

int pgx_kernel_time_ms(pgx_abacus *a, float *total_ms, uint32_t *n_sections) {
    if (!a || !total_ms) return fail(PGX_ERR_INVALID, "bad arguments");
    DeviceGuard guard(a->device);
    float sum = 0.f;
    for (size_t i = 0; i < a->ev_used; ++i) {
        float ms = 0.f;
        PGX_CUDA(cudaEventSynchronize(a->ev_pool[i].second));
        PGX_CUDA(cudaEventElapsedTime(&ms, a->ev_pool[i].first, a->ev_pool[i].second));
        sum += ms;
    }
    *total_ms = sum;
    if (n_sections) *n_sections = (uint32_t)a->ev_used;
    a->ev_used = 0;
    return PGX_OK;
}

uint64_t pgx_launch_count(const pgx_abacus *a) { return a ? a->launches : 0; }

int pgx_last_launch_info(const pgx_abacus *a, char *buf, size_t buflen) {
    if (!a || !buf || !buflen) return fail(PGX_ERR_INVALID, "bad arguments");
    snprintf(buf, buflen, "%s", a->last_launch.c_str());
    return PGX_OK;
}

}  // extern "C"
