// panacus_host.hpp -- C++ host layer above the C ABI (include/panacus_b200.h).
//
// The reference's host code is Rust; this image has no Rust toolchain, so the host side of the
// drop-in is C++ with the reference's names and semantics for everything the hot path needs around
// the GPU calls: GFA front end, path grouping / ordering, subset / exclude bookkeeping, thresholds,
// the closed-form growth formulas (f64, host only) and the TSV writers.  Reference citations are
// relative to marschall-lab/panacus @ 395ba41.
#pragma once

#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace panacus {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---- src/util.rs ------------------------------------------------------------------------------------
enum class CountType { Node, Bp, Edge, All };  // util.rs:44-49
std::string to_string(CountType c);
CountType count_type_from_str(const std::string &s);  // case-insensitive, like clap's ignore_case

struct Threshold {  // util.rs:327-364
    bool is_relative = false;
    double rel = 0.0;
    uint64_t abs = 0;
    static Threshold Relative(double v) { return {true, v, 0}; }
    static Threshold Absolute(uint64_t v) { return {false, 0.0, v}; }
    uint64_t to_absolute(uint64_t n) const;  // util.rs:351-356
    double to_relative(uint64_t n) const;    // util.rs:358-363
    std::string get_string() const;          // util.rs:344-349
};

struct ThresholdContainer {  // graph_broker/hist.rs:207-323
    std::vector<Threshold> coverage, quorum;
    static ThresholdContainer parse_params(const std::string &quorum, const std::string &coverage);
};

std::string format_f64(double v);  // Rust `{}` for f64 (shortest round trip, no exponent)
std::string format_f32(float v);   // Rust `{}` for f32

// ---- src/graph_broker/graph.rs ------------------------------------------------------------------------
struct PathSegment {  // graph.rs:469-616
    std::string sample;
    std::optional<std::string> haplotype, seqid;
    std::optional<uint64_t> start, end;

    static PathSegment from_str(const std::string &s);                                  // graph.rs:495-549
    static PathSegment from_str_start_end(const std::string &s, uint64_t a, uint64_t b);  // graph.rs:551-556
    std::string id() const;                                                             // graph.rs:558-579
    PathSegment clear_coords() const;                                                   // graph.rs:581-589
    std::optional<std::pair<uint64_t, uint64_t>> coords() const;
    std::string to_string() const;  // Display: id[:start-end]
    bool operator<(const PathSegment &o) const;
    bool operator==(const PathSegment &o) const;
};

struct Step {
    uint32_t node;  // item id (1-based)
    bool forward;
};

// worker threads of the host front end (GFA step parsing); 0 = hardware concurrency.  The CLI's -t / --threads.
void set_host_threads(int n);
unsigned host_thread_budget();  // -t N, or one per core

// canonical edge key -> edge id: open addressing with linear probing (one cache line per lookup instead of the node
// chasing of std::unordered_map; one lookup per path step when edges are counted).  Keys are never 0 (ids are 1-based).
class EdgeMap {
  public:
    // inserts (key, value) unless the key is present; returns the stored value
    uint32_t emplace(uint64_t key, uint32_t value) {
        if ((size_ + 1) * 10 > keys_.size() * 7) grow();
        size_t i = slot(key);
        while (keys_[i] != 0 && keys_[i] != key) i = (i + 1) & (keys_.size() - 1);
        if (keys_[i] == 0) {
            keys_[i] = key;
            vals_[i] = value;
            ++size_;
        }
        return vals_[i];
    }
    uint32_t find(uint64_t key) const {  // 0 = absent
        if (keys_.empty()) return 0;
        size_t i = slot(key);
        while (keys_[i] != 0 && keys_[i] != key) i = (i + 1) & (keys_.size() - 1);
        return keys_[i] == key ? vals_[i] : 0u;
    }
    size_t size() const { return size_; }
    template <typename F>
    void for_each(F f) const {
        for (size_t i = 0; i < keys_.size(); ++i)
            if (keys_[i]) f(keys_[i], vals_[i]);
    }

  private:
    // locality-preserving: the slot follows the first node of the edge, so the edges a path walks one after the other
    // (node ids largely follow path order in pangenome graphs) sit in neighbouring cache lines; collisions are probed
    size_t slot(uint64_t key) const { return (size_t)((key >> 32) * 2u) & (keys_.size() - 1); }
    void grow() {
        std::vector<uint64_t> ok;
        std::vector<uint32_t> ov;
        ok.swap(keys_);
        ov.swap(vals_);
        keys_.assign(ok.empty() ? 1024 : ok.size() * 2, 0);
        vals_.assign(keys_.size(), 0);
        size_ = 0;
        for (size_t i = 0; i < ok.size(); ++i)
            if (ok[i]) emplace(ok[i], ov[i]);
    }
    std::vector<uint64_t> keys_;
    std::vector<uint32_t> vals_;
    size_t size_ = 0;
};

struct GraphStorage {  // graph.rs:150-375
    std::vector<uint32_t> node_lens;  // [0] = 0; node ids are 1..=node_count() in S-line order (graph.rs:323-340)
    std::vector<PathSegment> path_segments;
    std::vector<std::vector<Step>> path_steps;  // steps of every P / W line, file order (empty when `lean`)
    // lean parse (counting nodes / bp without subset or exclude lists): the node ids of all paths in one flat u32 array,
    // path k = flat_nodes[flat_prefsum[k] .. flat_prefsum[k + 1]) -- the wire format of pgx_abacus_build_u32
    bool lean = false;
    std::unique_ptr<uint32_t[]> flat_nodes;
    std::vector<uint64_t> flat_prefsum;
    uint64_t step_count() const;  // total path steps, either representation
    // lean parse + a subset list that takes every path entirely or not at all (the usual list of path names, or a regex):
    // paths outside the subset lose their steps -- an empty id range, as the general table build leaves them
    // (util.rs:276-291).  False (nothing changed) if some path is only partly inside: parse again, the general way.
    bool lean_apply_subset(const struct GraphMask &mask);
    bool lean_subset_applied = false;
    // canonical edge (graph.rs:142-148) packed as ((u << 1 | fwd_u) << 32) | (v << 1 | fwd_v) -> id (1-based)
    EdgeMap edge2id;
    bool has_edges = false;
    // segment names by id ([0] = ""), kept only when asked for (the `table` writer, abacus.rs:1067-1072)
    std::vector<std::string> node_names;

    uint64_t node_count() const { return node_lens.size() - 1; }
    uint64_t edge_count() const { return edge2id.size(); }
    static uint64_t edge_key(uint32_t u, bool fu, uint32_t v, bool fv);
    static GraphStorage from_gfa(const std::string &path, bool with_edges, bool with_names = false, bool lean = false);
};

// ---- src/graph_broker/abacus.rs: GraphMask ---------------------------------------------------------------
struct GraphMaskParameters {
    std::string groupby;  // file
    bool groupby_haplotype = false, groupby_sample = false;
    std::string positive_list, negative_list;  // --subset / --exclude: file or regex
    std::optional<std::string> order;
};

struct GraphMask {  // abacus.rs:25-383
    std::map<PathSegment, std::string> groups;
    std::optional<std::vector<PathSegment>> include_coords, exclude_coords, order;

    static GraphMask from_graph(const GraphStorage &g, const GraphMaskParameters &p);  // abacus.rs:55-150
    // (path index, group name) in counting order; abacus.rs:310-347
    std::vector<std::pair<uint64_t, std::string>> get_path_order(const std::vector<PathSegment> &paths) const;
};

std::vector<PathSegment> parse_bed_to_path_segments(const std::string &path, bool use_block_info);  // io.rs:35-115

// ---- ItemTable + subset / exclude bookkeeping (src/util.rs:80-310, graph_broker/util.rs:208-790) -------------
struct ItemTables {
    std::vector<uint64_t> items;       // ItemTable.items
    const uint32_t *items32 = nullptr;  // lean parse: the same ids as u32, owned by the GraphStorage (then `items` is empty)
    uint64_t n_steps = 0;               // number of ids in items / items32
    std::vector<uint64_t> id_prefsum;  // ItemTable.id_prefsum (P + 1)
    std::vector<uint8_t> exclude;      // ActiveTable.items (N + 1) or empty
    std::map<uint64_t, uint64_t> uncovered_bps;  // abacus.rs:1187-1229
    uint64_t n_items = 0;
};
ItemTables build_item_tables(const GraphStorage &g, const GraphMask &mask, CountType count);

// ---- closed-form growth (src/graph_broker/hist.rs:21-187), f64 on the host ------------------------------------
struct Hist {
    CountType count = CountType::Node;
    std::vector<uint64_t> coverage;
    std::vector<double> calc_growth(const Threshold &t_coverage, const Threshold &t_quorum) const;  // hist.rs:51-66
    std::vector<std::vector<double>> calc_all_growths(const ThresholdContainer &aux) const;         // hist.rs:69-87
    unsigned growth_threads_ = 1;  // threads one quorum pair may use (set by calc_all_growths from the -t budget)
};
double choose(uint64_t n, uint64_t k);  // hist.rs:21-36

// ---- TSV (src/io.rs:460-604, analyses/{hist,growth,similarity}.rs) ----------------------------------------------
std::string write_table(const std::vector<std::vector<std::string>> &headers,
                        const std::vector<std::vector<double>> &columns, uint64_t start_index = 0);
std::string write_ordered_table(const std::vector<std::vector<std::string>> &headers,
                                const std::vector<std::vector<double>> &columns, const std::vector<std::string> &index);
std::string write_metadata_comments(const std::string &argv_joined, bool with_version);
// AbacusByGroup::to_tsv (abacus.rs:1056-1178): the per-item coverage table of the `table` analysis
std::string abacus_by_group_to_tsv(const GraphStorage &g, CountType count, bool total, const std::vector<std::string> &groups,
                                   const std::vector<uint64_t> &r, const std::vector<uint64_t> &c, const std::vector<uint32_t> &v,
                                   const std::map<uint64_t, uint64_t> &uncovered_bps);
// hierarchical clustering of the similarity table (cluster.cpp; kodama::linkage in the reference, similarity.rs:165-181):
// the observations in the order the dendrogram's steps name them (get_order_from_dendrogram, similarity.rs:206-219)
std::vector<size_t> cluster_leaf_order(const std::vector<std::vector<float>> &table, const std::string &method);
std::string debug_linkage(const std::vector<double> &condensed, size_t n, const std::string &method, bool f32);
// hist-only TSV re-ingestion for `growth <file.tsv>` (io.rs:152-290)
std::vector<Hist> parse_hists(const std::string &path, std::vector<std::string> &comments);

// ---- device side (RAII over include/panacus_b200.h) ----------------------------------------------------------
int device_count();  // pgx_device_count
void device_warmup_async(int n_devices);  // CUDA context creation on side threads (overlaps the GFA parse)

// One GPU's NCCL communicator (pgx_comm); create_all = the single-process form, one communicator per device, each to be
// driven from its own host thread (run_on_devices).
class DeviceComm {
  public:
    static std::vector<std::unique_ptr<DeviceComm>> create_all(const std::vector<int> &devices);
    ~DeviceComm();
    DeviceComm(const DeviceComm &) = delete;
    DeviceComm &operator=(const DeviceComm &) = delete;
    void *handle() const { return h_; }
    uint32_t rank() const { return rank_; }
    uint32_t world() const { return world_; }

  private:
    DeviceComm(void *h, uint32_t rank, uint32_t world) : h_(h), rank_(rank), world_(world) {}
    void *h_;
    uint32_t rank_, world_;
};
// fn(r) on one host thread per rank (collective calls of rank r); the first exception is rethrown after all joined
void run_on_devices(uint32_t n, const std::function<void(uint32_t)> &fn);
// items [lo, hi) (ids from 1) of rank r when n_items are cut into `world` item ranges
std::pair<uint64_t, uint64_t> item_range(uint64_t n_items, uint32_t rank, uint32_t world);

class DeviceAbacus {
  public:
    DeviceAbacus(uint64_t n_items, uint32_t n_groups, int device = 0);
    ~DeviceAbacus();
    DeviceAbacus(const DeviceAbacus &) = delete;
    DeviceAbacus &operator=(const DeviceAbacus &) = delete;
    // a5/a6 replacement: one scatter per path in counting order
    void build(const ItemTables &t, const std::vector<std::pair<uint64_t, std::string>> &path_order,
               std::vector<std::string> &group_names);
    void set_weights(const std::vector<uint32_t> &w);
    // packed host copies of the bitmap ((n_items + 1) x ceil(G/64) u64), for the on-disk abacus cache
    void download(std::vector<uint64_t> &bitmap) const;
    void upload(const std::vector<uint64_t> &bitmap);
    void hist(std::vector<uint64_t> *count, std::vector<uint64_t> *weight, std::vector<uint32_t> *countable);
    // AbacusByGroup::calc_growth for all threshold pairs (abacus.rs:989-1032); f64 like the reference
    std::vector<std::vector<double>> calc_growth(const ThresholdContainer &aux, bool weighted);
    void similarity(bool weighted, std::vector<uint64_t> &inter, std::vector<uint64_t> &len);
    // ---- multi-GPU (collective: every rank's thread calls with its own abacus and communicator) ----
    // this abacus <- items first_item .. first_item + n_items() - 1 of `src` (device-to-device, any two GPUs)
    void copy_rows_from(DeviceAbacus &src, uint64_t first_item);
    void broadcast(DeviceComm &comm, uint32_t root, bool with_weights);  // replicate root's bitmap (+ weights) over NVLink
    // item-range shards + ncclAllReduce: results of the whole graph on every rank
    std::vector<std::vector<double>> calc_growth_sharded(DeviceComm &comm, const ThresholdContainer &aux, bool weighted);
    void hist_sharded(DeviceComm &comm, std::vector<uint64_t> *count, std::vector<uint64_t> *weight);
    // replicated bitmap, upper-triangle row blocks per rank + ncclAllGather
    void similarity_sharded(DeviceComm &comm, bool weighted, std::vector<uint64_t> &inter, std::vector<uint64_t> &len);
    // AbacusByGroup {r, c, v} (abacus.rs:790-799) of the abacus built from `t` under `path_order`; want_v = false skips
    // the occurrence counts (the --total table never reads them, abacus.rs:1093-1096)
    void csr(const ItemTables &t, const std::vector<std::pair<uint64_t, std::string>> &path_order, std::vector<uint64_t> &r,
             std::vector<uint64_t> &c, std::vector<uint32_t> &v, bool want_v);
    uint32_t n_groups() const { return n_groups_; }
    uint64_t n_items() const { return n_items_; }

  private:
    void *h_ = nullptr;
    uint64_t n_items_;
    uint32_t n_groups_;
};

// ---- measurement helpers (not part of the reference CLI) ----------------------------------------------------------
// pangenome-shaped GFA text for a given coverage histogram (synth.cpp); returns the number of path steps written
uint64_t synth_gfa(const std::string &path, uint64_t n_nodes, uint32_t samples, uint32_t haps, uint32_t contigs,
                   const std::vector<double> &node_hist, const std::vector<double> &bp_hist, double mean_len, uint64_t seed);

// ---- packed-abacus cache file (".pabm"): skip GFA parsing on repeated runs (the analogue of the
// reference's hist-TSV reuse, io.rs:244-290) -----------------------------------------------------------------------
struct AbacusFile {
    CountType count = CountType::Node;
    uint64_t n_items = 0;
    std::vector<std::string> groups;          // counting order
    std::vector<uint32_t> weights;            // node_lens (node / bp) or all ones (edge), n_items + 1
    std::map<uint64_t, uint64_t> uncovered;   // uncovered_bps (bp with --subset)
    std::vector<uint64_t> bitmap;             // (n_items + 1) x ceil(G/64), node-major
    void save(const std::string &path) const;
    static AbacusFile load(const std::string &path);
};

}  // namespace panacus
