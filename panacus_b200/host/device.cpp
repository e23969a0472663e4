// device.cpp -- RAII wrapper over the C ABI (include/panacus_b200.h).  All counting happens in
// libpanacus_b200.so; there is no host fallback.
#include <cmath>
#include <exception>
#include <mutex>
#include <thread>

#include "../../include/panacus_b200.h"
#include "panacus_host.hpp"

namespace panacus {

namespace {
void check(int rc, const char *what) {
    if (rc != PGX_OK) throw Error(std::string(what) + ": " + pgx_last_error());
}
pgx_abacus *H(void *p) { return static_cast<pgx_abacus *>(p); }
}  // namespace

DeviceAbacus::DeviceAbacus(uint64_t n_items, uint32_t n_groups, int device) : n_items_(n_items), n_groups_(n_groups) {
    pgx_abacus *h = nullptr;
    check(pgx_abacus_create(&h, device, n_items, n_groups), "pgx_abacus_create");
    h_ = h;
}

DeviceAbacus::~DeviceAbacus() { pgx_abacus_destroy(H(h_)); }

void DeviceAbacus::build(const ItemTables &t, const std::vector<std::pair<uint64_t, std::string>> &path_order,
                         std::vector<std::string> &group_names) {
    // group id = rank of first appearance in counting order (abacus.rs:555-569, 816-829)
    group_names.clear();
    const uint8_t *ex = t.exclude.empty() ? nullptr : t.exclude.data();
    std::vector<int64_t> path_group(t.id_prefsum.size() - 1, -1);  // -1: path not counted
    for (auto &po : path_order) {
        if (group_names.empty() || group_names.back() != po.second) group_names.push_back(po.second);
        path_group[po.first] = (int64_t)group_names.size() - 1;
    }
    if (t.items32)  // lean parse: u32 ids, half the bytes on the wire and no narrowing pass
        check(pgx_abacus_build_u32(H(h_), t.items32, t.n_steps, t.id_prefsum.data(), path_group.size(), path_group.data(), ex),
              "pgx_abacus_build_u32");
    else
        check(pgx_abacus_build(H(h_), t.items.data(), t.items.size(), t.id_prefsum.data(), path_group.size(), path_group.data(), ex),
              "pgx_abacus_build");
}

void DeviceAbacus::csr(const ItemTables &t, const std::vector<std::pair<uint64_t, std::string>> &path_order,
                       std::vector<uint64_t> &r, std::vector<uint64_t> &c, std::vector<uint32_t> &v, bool want_v) {
    r.assign(n_items_ + 2, 0);
    uint64_t nnz = 0;
    check(pgx_abacus_csr_rows(H(h_), r.data(), &nnz), "pgx_abacus_csr_rows");
    c.assign(nnz, 0);
    v.clear();
    if (!want_v) {
        check(pgx_abacus_csr_fill(H(h_), nullptr, 0, nullptr, 0, nullptr, nullptr, c.data(), nullptr), "pgx_abacus_csr_fill");
        return;
    }
    v.assign(nnz, 0);
    std::vector<int64_t> path_group(t.id_prefsum.size() - 1, -1);  // same numbering as build()
    int64_t gid = -1;
    const std::string *last = nullptr;
    for (auto &po : path_order) {
        if (!last || *last != po.second) ++gid;
        last = &po.second;
        path_group[po.first] = gid;
    }
    check(pgx_abacus_csr_fill(H(h_), t.items.data(), t.items.size(), t.id_prefsum.data(), path_group.size(), path_group.data(),
                              t.exclude.empty() ? nullptr : t.exclude.data(), c.data(), v.data()),
          "pgx_abacus_csr_fill");
}

void DeviceAbacus::set_weights(const std::vector<uint32_t> &w) {
    if (w.size() != n_items_ + 1) throw Error("weight vector must have n_items + 1 entries");
    check(pgx_abacus_upload(H(h_), nullptr, 0, w.data()), "pgx_abacus_upload(weights)");
}

void DeviceAbacus::download(std::vector<uint64_t> &bitmap) const {
    const uint32_t W = (n_groups_ + 63u) / 64u;
    bitmap.assign((size_t)(n_items_ + 1) * W, 0);
    check(pgx_abacus_download(H(h_), bitmap.data(), W), "pgx_abacus_download");
}

void DeviceAbacus::upload(const std::vector<uint64_t> &bitmap) {
    const uint32_t W = (n_groups_ + 63u) / 64u;
    if (bitmap.size() != (size_t)(n_items_ + 1) * W) throw Error("bitmap size does not match the abacus shape");
    check(pgx_abacus_upload(H(h_), bitmap.data(), W, nullptr), "pgx_abacus_upload");
}

void DeviceAbacus::hist(std::vector<uint64_t> *count, std::vector<uint64_t> *weight, std::vector<uint32_t> *countable) {
    if (count) count->assign(n_groups_ + 1, 0);
    if (weight) weight->assign(n_groups_ + 1, 0);
    if (countable) countable->assign(n_items_ + 1, 0);
    check(pgx_hist(H(h_), count ? count->data() : nullptr, weight ? weight->data() : nullptr,
                   countable ? countable->data() : nullptr),
          "pgx_hist");
}

namespace {
void growth_cutoffs(const ThresholdContainer &aux, uint32_t G, std::vector<uint32_t> &cov, std::vector<uint32_t> &thr) {
    const uint32_t T = (uint32_t)aux.coverage.size();
    cov.assign(T, 0);
    thr.assign((size_t)T * G, 0);
    for (uint32_t t = 0; t < T; ++t) {
        cov[t] = (uint32_t)std::max<uint64_t>(1, aux.coverage[t].to_absolute(G));  // abacus.rs:997
        const double q = std::max(0.0, aux.quorum[t].to_relative(G));              // abacus.rs:998
        for (uint32_t g = 0; g < G; ++g) {
            const double need = std::ceil(((double)g + 1.0) * q);  // abacus.rs:1010, same f64 expression
            thr[(size_t)t * G + g] = need > 0.0 ? (uint32_t)need : 0u;
        }
    }
}
pgx_comm *CH(DeviceComm &c) { return static_cast<pgx_comm *>(c.handle()); }
}  // namespace

std::vector<std::vector<double>> DeviceAbacus::calc_growth(const ThresholdContainer &aux, bool weighted) {
    const uint32_t G = n_groups_, T = (uint32_t)aux.coverage.size();
    std::vector<uint32_t> cov, thr;
    growth_cutoffs(aux, G, cov, thr);
    std::vector<uint64_t> curve((size_t)T * G);
    check(pgx_ordered_growth(H(h_), T, cov.data(), thr.data(), nullptr, weighted ? 1 : 0, curve.data()), "pgx_ordered_growth");
    std::vector<std::vector<double>> out(T, std::vector<double>(G));
    for (uint32_t t = 0; t < T; ++t)
        for (uint32_t g = 0; g < G; ++g) out[t][g] = (double)curve[(size_t)t * G + g];  // exact below 2^53
    return out;
}

void device_warmup_async(int n_devices) {
    // context creation overlaps the GFA parse; errors surface later, at the first real device call
    for (int d = 0; d < n_devices; ++d) std::thread([d] { pgx_device_warmup(d); }).detach();
}

int device_count() {
    int n = 0;
    check(pgx_device_count(&n), "pgx_device_count");
    return n;
}

std::vector<std::unique_ptr<DeviceComm>> DeviceComm::create_all(const std::vector<int> &devices) {
    std::vector<pgx_comm *> raw(devices.size(), nullptr);
    check(pgx_comm_create_all(raw.data(), (uint32_t)devices.size(), devices.data()), "pgx_comm_create_all");
    std::vector<std::unique_ptr<DeviceComm>> out;
    for (size_t i = 0; i < raw.size(); ++i) out.emplace_back(new DeviceComm(raw[i], (uint32_t)i, (uint32_t)raw.size()));
    return out;
}

DeviceComm::~DeviceComm() { pgx_comm_destroy(static_cast<pgx_comm *>(h_)); }

void run_on_devices(uint32_t n, const std::function<void(uint32_t)> &fn) {
    std::vector<std::thread> th;
    std::exception_ptr first;
    std::mutex mu;
    for (uint32_t r = 0; r < n; ++r)
        th.emplace_back([&, r] {
            try {
                fn(r);
            } catch (...) {
                std::lock_guard<std::mutex> lk(mu);
                if (!first) first = std::current_exception();
            }
        });
    for (auto &t : th) t.join();
    if (first) std::rethrow_exception(first);
}

std::pair<uint64_t, uint64_t> item_range(uint64_t n_items, uint32_t rank, uint32_t world) {
    const uint64_t base = n_items / world, rem = n_items % world;
    const uint64_t lo = 1 + rank * base + std::min<uint64_t>(rank, rem);
    return {lo, lo + base + (rank < rem ? 1 : 0)};
}

void DeviceAbacus::copy_rows_from(DeviceAbacus &src, uint64_t first_item) {
    check(pgx_abacus_copy_rows(H(h_), H(src.h_), first_item), "pgx_abacus_copy_rows");
}

void DeviceAbacus::broadcast(DeviceComm &comm, uint32_t root, bool with_weights) {
    check(pgx_abacus_broadcast(H(h_), CH(comm), root, with_weights ? 1 : 0), "pgx_abacus_broadcast");
}

std::vector<std::vector<double>> DeviceAbacus::calc_growth_sharded(DeviceComm &comm, const ThresholdContainer &aux, bool weighted) {
    const uint32_t G = n_groups_, T = (uint32_t)aux.coverage.size();
    std::vector<uint32_t> cov, thr;
    growth_cutoffs(aux, G, cov, thr);
    std::vector<uint64_t> curve((size_t)T * G);
    check(pgx_hist_ordered_growth_sharded(H(h_), CH(comm), nullptr, nullptr, T, cov.data(), thr.data(), weighted ? 1 : 0, curve.data()),
          "pgx_hist_ordered_growth_sharded");
    std::vector<std::vector<double>> out(T, std::vector<double>(G));
    for (uint32_t t = 0; t < T; ++t)
        for (uint32_t g = 0; g < G; ++g) out[t][g] = (double)curve[(size_t)t * G + g];
    return out;
}

void DeviceAbacus::hist_sharded(DeviceComm &comm, std::vector<uint64_t> *count, std::vector<uint64_t> *weight) {
    if (count) count->assign(n_groups_ + 1, 0);
    if (weight) weight->assign(n_groups_ + 1, 0);
    check(pgx_hist_ordered_growth_sharded(H(h_), CH(comm), count ? count->data() : nullptr, weight ? weight->data() : nullptr, 0,
                                          nullptr, nullptr, 0, nullptr),
          "pgx_hist_ordered_growth_sharded");
}

void DeviceAbacus::similarity_sharded(DeviceComm &comm, bool weighted, std::vector<uint64_t> &inter, std::vector<uint64_t> &len) {
    inter.assign((size_t)n_groups_ * n_groups_, 0);
    len.assign(n_groups_, 0);
    check(pgx_similarity_sharded(H(h_), CH(comm), weighted ? 1 : 0, inter.data(), len.data()), "pgx_similarity_sharded");
}

void DeviceAbacus::similarity(bool weighted, std::vector<uint64_t> &inter, std::vector<uint64_t> &len) {
    inter.assign((size_t)n_groups_ * n_groups_, 0);
    len.assign(n_groups_, 0);
    check(pgx_similarity(H(h_), weighted ? 1 : 0, 0, n_groups_, inter.data(), len.data()), "pgx_similarity");
}

}  // namespace panacus
