// pgx_comm.cu -- multi-GPU side of the C ABI: an NCCL communicator handle (pgx_comm) and the sharded entry points
// (north_star: "one permutation / similarity row block per GPU with a final NCCL all-gather").
//
//   pgx_permuted_growth_sharded     order p -> rank p % world; curves stay on the device, one ncclAllGather on the
//                                   handle's stream, one copy to the host.  The reference's only parallel axis here is
//                                   across threshold pairs (src/analyses/ordered_histgrowth.rs:174-188).
//   pgx_similarity_sharded          every rank sums the (mirrored upper-triangular) matrix over its share of the ITEMS,
//                                   one ncclAllReduce (u64 sum) of the G x G partial matrices, one copy to the host.
//                                   Similarity::set_table is serial in the reference (src/analyses/similarity.rs:130-150).
//   pgx_hist_ordered_growth_sharded item-range shards, ncclAllReduce(u64 sum) of the KB-sized fused result vector (the
//                                   NCCL twin of the in-kernel NVLink exchange of pgx_exchange_connect).
//   pgx_abacus_broadcast            replicate a rank's bitmap + weights over NVLink instead of one H2D upload per GPU.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2"): inside a PyTorch process that is the NCCL torch already
// loaded, in the C++ CLI the system library; libpanacus_b200.so itself has no link-time NCCL dependency and every
// single-GPU entry point works without it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "pgx_handle.h"

using namespace pgx;

struct pgx_comm {
    ncclComm_t comm = nullptr;
    int device = 0;
    uint32_t rank = 0, world = 1;
};

namespace {

struct Nccl {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

Nccl g_nccl;
std::once_flag g_nccl_once;

template <typename F>
bool sym(void *lib, const char *name, F &fn, std::string &err) {
    fn = reinterpret_cast<F>(dlsym(lib, name));
    if (!fn) err = std::string("NCCL symbol missing: ") + name;
    return fn != nullptr;
}

const Nccl *nccl() {
    std::call_once(g_nccl_once, [] {
        Nccl &n = g_nccl;
        // NCCL_DEBUG output goes to stdout by default; results (TSV) own stdout, so send the log to stderr unless the user chose a file
        setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
        const char *env = getenv("PGX_NCCL_LIB");  // explicit path override
        const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char *name : names) {
            if (!name || !*name) continue;
            if ((n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL))) break;
        }
        if (!n.lib) {
            n.error = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
            return;
        }
        const bool ok = sym(n.lib, "ncclGetUniqueId", n.GetUniqueId, n.error) && sym(n.lib, "ncclCommInitRank", n.CommInitRank, n.error) &&
                        sym(n.lib, "ncclCommInitAll", n.CommInitAll, n.error) && sym(n.lib, "ncclCommDestroy", n.CommDestroy, n.error) &&
                        sym(n.lib, "ncclAllGather", n.AllGather, n.error) && sym(n.lib, "ncclAllReduce", n.AllReduce, n.error) &&
                        sym(n.lib, "ncclBroadcast", n.Broadcast, n.error) && sym(n.lib, "ncclGroupStart", n.GroupStart, n.error) &&
                        sym(n.lib, "ncclGroupEnd", n.GroupEnd, n.error) && sym(n.lib, "ncclGetErrorString", n.GetErrorString, n.error);
        if (!ok) n.lib = nullptr;
    });
    return g_nccl.lib ? &g_nccl : nullptr;
}

#define PGX_NCCL(call)                                                                                         \
    do {                                                                                                       \
        ncclResult_t r__ = (call);                                                                             \
        if (r__ != ncclSuccess) return fail(PGX_ERR_NCCL, std::string(#call) + ": " + N->GetErrorString(r__)); \
    } while (0)

int need_nccl(const Nccl **out) {
    *out = nccl();
    if (!*out) return fail(PGX_ERR_NCCL, g_nccl.error.empty() ? "NCCL unavailable" : g_nccl.error);
    return PGX_OK;
}

int check_comm(const pgx_abacus *a, const pgx_comm *c) {
    if (!c || !c->comm) return fail(PGX_ERR_INVALID, "null communicator");
    if (a && c->device != a->device) return fail(PGX_ERR_INVALID, "communicator and abacus live on different devices");
    return PGX_OK;
}

}  // namespace

namespace pgx {

// Row-block boundaries of the upper-triangle sharding: 2 * world blocks, tile aligned; rank r owns blocks r and
// 2 * world - 1 - r (block b costs ~ rows x (G - start_b), so every rank's pair adds up to the same share).
void sim_block_bounds(uint32_t G, uint32_t world, uint32_t *bounds /* 2 * world + 1 */) {
    const uint32_t nb = 2u * world, tile = 64u, tiles = (G + tile - 1u) / tile;
    for (uint32_t k = 0; k <= nb; ++k) {
        const uint64_t b = (uint64_t)k * tiles / nb * tile;
        bounds[k] = (uint32_t)(b < G ? b : G);
    }
    bounds[nb] = G;
}

}  // namespace pgx

extern "C" {

int pgx_comm_unique_id(void *id_out) {
    if (!id_out) return fail(PGX_ERR_INVALID, "null pointer");
    const Nccl *N;
    int rc = need_nccl(&N);
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == PGX_COMM_ID_BYTES, "unique id size");
    ncclUniqueId id;
    PGX_NCCL(N->GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof(id));
    return PGX_OK;
}

int pgx_comm_create(pgx_comm **out, int device, uint32_t rank, uint32_t world, const void *unique_id) {
    if (!out || !unique_id) return fail(PGX_ERR_INVALID, "null pointer");
    *out = nullptr;
    if (world == 0 || rank >= world) return fail(PGX_ERR_INVALID, "bad rank / world");
    if (world > (uint32_t)kMaxRanks) return fail(PGX_ERR_UNSUPPORTED, "world > 8 (one NVSwitch domain)");
    const Nccl *N;
    int rc = need_nccl(&N);
    if (rc) return rc;
    DeviceGuard guard(device);
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    pgx_comm *c = new (std::nothrow) pgx_comm();
    if (!c) return fail(PGX_ERR_NOMEM, "out of host memory");
    c->device = device;
    c->rank = rank;
    c->world = world;
    ncclResult_t r = N->CommInitRank(&c->comm, (int)world, id, (int)rank);
    if (r != ncclSuccess) {
        delete c;
        return fail(PGX_ERR_NCCL, std::string("ncclCommInitRank: ") + N->GetErrorString(r));
    }
    *out = c;
    return PGX_OK;
}

int pgx_comm_create_all(pgx_comm **out, uint32_t n, const int *devices) {
    if (!out || !devices) return fail(PGX_ERR_INVALID, "null pointer");
    for (uint32_t i = 0; i < n; ++i) out[i] = nullptr;
    if (n == 0) return fail(PGX_ERR_INVALID, "no devices");
    if (n > (uint32_t)kMaxRanks) return fail(PGX_ERR_UNSUPPORTED, "more than 8 devices (one NVSwitch domain)");
    const Nccl *N;
    int rc = need_nccl(&N);
    if (rc) return rc;
    ncclComm_t comms[kMaxRanks];
    PGX_NCCL(N->CommInitAll(comms, (int)n, devices));
    for (uint32_t i = 0; i < n; ++i) {
        pgx_comm *c = new (std::nothrow) pgx_comm();
        if (!c) {
            for (uint32_t k = 0; k < n; ++k) {
                if (k >= i) N->CommDestroy(comms[k]);
                pgx_comm_destroy(out[k]);
                out[k] = nullptr;
            }
            return fail(PGX_ERR_NOMEM, "out of host memory");
        }
        c->comm = comms[i];
        c->device = devices[i];
        c->rank = i;
        c->world = n;
        out[i] = c;
    }
    return PGX_OK;
}

void pgx_comm_destroy(pgx_comm *c) {
    if (!c) return;
    const Nccl *N = nccl();
    if (N && c->comm) {
        DeviceGuard guard(c->device);
        N->CommDestroy(c->comm);
    }
    delete c;
}

int pgx_comm_info(const pgx_comm *c, uint32_t *rank, uint32_t *world, int *device) {
    if (!c) return fail(PGX_ERR_INVALID, "null communicator");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (device) *device = c->device;
    return PGX_OK;
}

int pgx_similarity_shard_bounds(uint32_t n_groups, uint32_t world, uint32_t *bounds) {
    if (!bounds || n_groups == 0 || world == 0 || world > (uint32_t)kMaxRanks) return fail(PGX_ERR_INVALID, "bad arguments");
    sim_block_bounds(n_groups, world, bounds);
    return PGX_OK;
}

int pgx_abacus_broadcast(pgx_abacus *a, pgx_comm *c, uint32_t root, int with_weights) {
    int rc = check_handle(a);
    if (rc || (rc = check_comm(a, c))) return rc;
    if (root >= c->world) return fail(PGX_ERR_INVALID, "root >= world");
    const Nccl *N;
    if ((rc = need_nccl(&N))) return rc;
    DeviceGuard guard(a->device);
    if (with_weights && !a->d_weight) {  // the receivers need a buffer; the root must already hold its weights
        if (c->rank == root) return fail(PGX_ERR_STATE, "root has no weights to broadcast");
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&a->d_weight), a->n_rows * 4u));
        a->own_weight = true;
    }
    PGX_NCCL(N->GroupStart());
    PGX_NCCL(N->Broadcast(a->d_bitmap, a->d_bitmap, (size_t)a->n_rows * a->Wp, ncclUint64, (int)root, c->comm, a->stream));
    if (with_weights) PGX_NCCL(N->Broadcast(a->d_weight, a->d_weight, a->n_rows, ncclUint32, (int)root, c->comm, a->stream));
    PGX_NCCL(N->GroupEnd());
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    if (c->rank != root) {
        invalidate_derived(a);
        if (with_weights) {
            a->max_weight_known = false;
            a->planes_valid = false;
        }
    }
    return PGX_OK;
}

int pgx_exchange_connect_comm(pgx_abacus *a, pgx_comm *c) {
    int rc = check_handle(a);
    if (rc || (rc = check_comm(a, c))) return rc;
    const Nccl *N;
    if ((rc = need_nccl(&N))) return rc;
    DeviceGuard guard(a->device);
    unsigned char mine[PGX_EXCHANGE_HANDLE_BYTES];
    if ((rc = pgx_exchange_export(a, mine))) return rc;
    const size_t bytes = (size_t)c->world * PGX_EXCHANGE_HANDLE_BYTES;
    if ((rc = ensure_dev(&a->d_scratch, &a->scratch_cap, (bytes + PGX_EXCHANGE_HANDLE_BYTES + 7u) / 8u))) return rc;
    unsigned char *d_all = reinterpret_cast<unsigned char *>(a->d_scratch), *d_mine = d_all + bytes;
    PGX_CUDA(cudaMemcpyAsync(d_mine, mine, sizeof mine, cudaMemcpyHostToDevice, a->stream));
    PGX_NCCL(N->AllGather(d_mine, d_all, sizeof mine, ncclUint8, c->comm, a->stream));
    std::vector<unsigned char> all(bytes);
    PGX_CUDA(cudaMemcpyAsync(all.data(), d_all, bytes, cudaMemcpyDeviceToHost, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    return pgx_exchange_connect(a, c->rank, c->world, all.data());
}

int pgx_hist_ordered_growth_sharded(pgx_abacus *a, pgx_comm *c, uint64_t *hist_count, uint64_t *hist_weight, uint32_t n_thresholds,
                                    const uint32_t *cov_abs, const uint32_t *quorum_thr, int weighted, uint64_t *curve) {
    int rc = check_handle(a);
    if (rc || (rc = check_comm(a, c))) return rc;
    if (n_thresholds && !curve) return fail(PGX_ERR_INVALID, "curve is null");
    if (a->x.world > 1u) return fail(PGX_ERR_STATE, "the fused NVLink exchange is connected: its passes are already summed over the ranks");
    const Nccl *N;
    if ((rc = need_nccl(&N))) return rc;
    DeviceGuard guard(a->device);
    const uint32_t G = a->G, G1 = G + 1u;
    const size_t words = pgx_fused_out_words(G, n_thresholds);
    if ((rc = ensure_dev(&a->d_scratch, &a->scratch_cap, words))) return rc;
    if ((rc = ensure_pinned(a, words))) return rc;
    // the fused layout is written sparsely (only the requested parts): clear it so that every rank sums defined words
    PGX_CUDA(cudaMemsetAsync(a->d_scratch, 0, words * 8u, a->stream));
    if ((rc = fused_pass(a, hist_count != nullptr, hist_weight != nullptr, n_thresholds, cov_abs, quorum_thr, weighted, nullptr,
                         a->d_scratch)))
        return rc;
    // the path's one exchange: wrapping u64 sums of the per-shard histograms and first differences (order independent)
    PGX_NCCL(N->AllReduce(a->d_scratch, a->d_scratch, words, ncclUint64, ncclSum, c->comm, a->stream));
    PGX_CUDA(cudaMemcpyAsync(a->h_pinned, a->d_scratch, words * 8u, cudaMemcpyDeviceToHost, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    if (hist_count) std::memcpy(hist_count, a->h_pinned, (size_t)G1 * 8u);
    if (hist_weight) std::memcpy(hist_weight, a->h_pinned + G1, (size_t)G1 * 8u);
    for (uint32_t t = 0; t < n_thresholds; ++t) {
        uint64_t run = 0;
        const uint64_t *d = a->h_pinned + 2u * G1 + (size_t)t * G;
        for (uint32_t j = 0; j < G; ++j) {
            run += d[j];
            curve[(size_t)t * G + j] = run;
        }
    }
    return PGX_OK;
}

int pgx_permuted_growth_sharded(pgx_abacus *a, pgx_comm *c, uint32_t n_orders, const uint32_t *orders, uint32_t n_thresholds,
                                const uint32_t *cov_abs, const uint32_t *quorum_thr, int weighted, uint64_t *curves) {
    int rc = check_handle(a);
    if (rc || (rc = check_comm(a, c))) return rc;
    if (n_thresholds && n_orders && (!curves || !orders)) return fail(PGX_ERR_INVALID, "orders / curves is null");
    const Nccl *N;
    if ((rc = need_nccl(&N))) return rc;
    if ((rc = validate_thresholds(a, n_thresholds, cov_abs))) return rc;
    if (n_orders == 0 || n_thresholds == 0) return PGX_OK;
    DeviceGuard guard(a->device);
    const uint32_t G = a->G, T = n_thresholds, world = c->world, rank = c->rank;
    // order p -> rank p % world; every rank sends the same (padded) number of curves
    const uint32_t max_local = (n_orders + world - 1u) / world;
    const uint32_t n_local = n_orders > rank ? (n_orders - rank + world - 1u) / world : 0u;
    std::vector<uint32_t> mine((size_t)n_local * G);
    for (uint32_t k = 0; k < n_local; ++k)
        std::memcpy(mine.data() + (size_t)k * G, orders + (size_t)(rank + (size_t)k * world) * G, (size_t)G * 4u);
    const size_t per_rank = (size_t)max_local * T * G;
    // scratch: [ gathered: world x per_rank | send: per_rank ]
    if ((rc = ensure_dev(&a->d_scratch, &a->scratch_cap, per_rank * ((size_t)world + 1u)))) return rc;
    uint64_t *d_all = a->d_scratch, *d_send = a->d_scratch + per_rank * world;
    if (n_local < max_local) PGX_CUDA(cudaMemsetAsync(d_send, 0, per_rank * 8u, a->stream));
    if (n_local && (rc = gm_growth_device(a, n_local, mine.data(), T, cov_abs, quorum_thr, weighted, d_send))) return rc;
    PGX_NCCL(N->AllGather(d_send, d_all, per_rank, ncclUint64, c->comm, a->stream));
    // one copy to the host, then the round-robin interleave while copying out of the staging buffer
    if ((rc = ensure_pinned(a, per_rank * world))) return rc;
    PGX_CUDA(cudaMemcpyAsync(a->h_pinned, d_all, per_rank * world * 8u, cudaMemcpyDeviceToHost, a->stream));
    PGX_CUDA(cudaStreamSynchronize(a->stream));
    const size_t curve_words = (size_t)T * G;
    for (uint32_t p = 0; p < n_orders; ++p)
        std::memcpy(curves + (size_t)p * curve_words, a->h_pinned + ((size_t)(p % world) * max_local + p / world) * curve_words,
                    curve_words * 8u);
    return PGX_OK;
}

int pgx_similarity_sharded(pgx_abacus *a, pgx_comm *c, int weighted, uint64_t *inter, uint64_t *len) {
    int rc = check_handle(a);
    if (rc || (rc = check_comm(a, c))) return rc;
    if (!inter || !len) return fail(PGX_ERR_INVALID, "inter / len is null");
    const Nccl *N;
    if ((rc = need_nccl(&N))) return rc;
    DeviceGuard guard(a->device);
    const uint32_t G = a->G, world = c->world, rank = c->rank;
    // The contraction runs over the items, so the ITEMS are split: rank r sums the whole (upper-triangular, mirrored)
    // matrix over its share of the 64-item words, then one ncclAllReduce (u64 sum) of the G x G partial matrices.  Any G
    // and any world size balance perfectly (row blocks -- the first version, kept as pgx_similarity_upper for callers
    // that shard themselves -- leave the 128 x 256 tensor-core tiles half empty at 8 ranks x 64 rows).
    const uint64_t n_words = (a->n_rows + 63u) / 64u;
    uint64_t per = (n_words + world - 1u) / world;
    per = (per + 1u) & ~1ull;  // even: keeps the 16-byte operand loads of k_sim_mma aligned
    const uint64_t wb = std::min<uint64_t>((uint64_t)rank * per, n_words), we = std::min<uint64_t>(wb + per, n_words);
    const size_t full = (size_t)G * G;
    if ((rc = ensure_dev(&a->d_scratch, &a->scratch_cap, full))) return rc;
    uint64_t *d_full = a->d_scratch;
    // with timing enabled: device time of the phases (compute | all-reduce | copy to the host) in last_launch_info
    cudaEvent_t ev[4] = {};
    auto mark = [&](int k) {
        if (a->timing && cudaEventCreate(&ev[k]) == cudaSuccess) cudaEventRecord(ev[k], a->stream);
    };
    mark(0);
    PGX_CUDA(cudaMemsetAsync(d_full, 0, full * 8u, a->stream));
    if ((rc = sim_rows_device(a, weighted, 0, G, 0, d_full, false, wb, we))) return rc;
    mark(1);
    PGX_NCCL(N->AllReduce(d_full, d_full, full, ncclUint64, ncclSum, c->comm, a->stream));
    mark(2);
    rc = copy_to_host(a, inter, d_full, full);
    mark(3);
    if (!rc)
        for (uint32_t g = 0; g < G; ++g) len[g] = inter[(size_t)g * G + g];  // len[g] = the diagonal entry
    if (a->timing && ev[0] && ev[1] && ev[2] && ev[3]) {
        float t[3] = {};
        cudaEventSynchronize(ev[3]);
        for (int k = 0; k < 3; ++k) cudaEventElapsedTime(&t[k], ev[k], ev[k + 1]);
        char buf[200];
        snprintf(buf, sizeof buf, "pgx_similarity_sharded phases ms: compute %.3f | all-reduce %.3f | to host %.3f (%s)", t[0], t[1], t[2],
                 a->last_launch.c_str());
        a->last_launch = buf;
    }
    for (auto e : ev)
        if (e) cudaEventDestroy(e);
    return rc;
}

}  // extern "C"
