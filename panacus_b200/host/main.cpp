// main.cpp -- `panacus` command line for the hot-path subcommands, with the reference's flags:
//   hist | growth | histgrowth | ordered-histgrowth | similarity | table (+ coverage-line, report)
// (src/commands/{hist,growth,histgrowth,ordered_histgrowth,similarity}.rs).  Counting runs on the GPU
// through libpanacus_b200.so; parsing, grouping, thresholds, closed-form growth and TSV stay here.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>

#include "panacus_host.hpp"

using namespace panacus;

namespace {

struct Args {
    std::string sub;
    std::vector<std::string> positional;
    std::map<std::string, std::string> opt;  // long name -> value ("1" for flags)
    bool has(const std::string &k) const { return opt.count(k) != 0; }
    std::string get(const std::string &k, const std::string &dflt = "") const {
        auto it = opt.find(k);
        return it == opt.end() ? dflt : it->second;
    }
};

const std::map<char, std::string> kShort = {{'s', "subset"},  {'e', "exclude"},  {'g', "groupby"}, {'H', "groupby-haplotype"},
                                             {'S', "groupby-sample"}, {'c', "count"}, {'l', "coverage"}, {'q', "quorum"},
                                             {'a', "hist"},    {'O', "order"},    {'m', "method"},  {'t', "threads"},
                                             {'v', "verbose"}};
const std::set<std::string> kFlags = {"groupby-haplotype", "groupby-sample", "hist", "verbose", "total", "dry-run", "json", "names", "no-cluster", "timing", "lean"};

Args parse_args(int argc, char **argv) {
    Args a;
    for (int i = 1; i < argc; ++i) {
        std::string tok = argv[i];
        std::string name, value;
        bool have_value = false;
        if (tok.rfind("--", 0) == 0) {
            const size_t eq = tok.find('=');
            name = tok.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            if (eq != std::string::npos) {
                value = tok.substr(eq + 1);
                have_value = true;
            }
        } else if (tok.size() >= 2 && tok[0] == '-' && !(tok[1] >= '0' && tok[1] <= '9')) {
            auto it = kShort.find(tok[1]);
            if (it == kShort.end()) throw Error("unknown option " + tok);
            name = it->second;
            if (tok.size() > 2) {
                if (kFlags.count(name)) {  // bundled flags, e.g. -SH
                    a.opt[name] = "1";
                    for (size_t k = 2; k < tok.size(); ++k) {
                        auto jt = kShort.find(tok[k]);
                        if (jt == kShort.end() || !kFlags.count(jt->second)) throw Error("unknown option in " + tok);
                        a.opt[jt->second] = "1";
                    }
                    continue;
                }
                value = tok.substr(2);
                have_value = true;
            }
        } else {
            if (a.sub.empty())
                a.sub = tok;
            else
                a.positional.push_back(tok);
            continue;
        }
        if (kFlags.count(name)) {
            a.opt[name] = "1";
        } else {
            if (!have_value) {
                if (i + 1 >= argc) throw Error("option --" + name + " needs a value");
                value = argv[++i];
            }
            a.opt[name] = value;
        }
    }
    return a;
}

std::string argv_joined(int argc, char **argv) {
    std::string s;
    for (int i = 0; i < argc; ++i) {
        if (i) s += ' ';
        s += argv[i];
    }
    return s;
}

// --timing: wall time of the phases of a run, one JSON object on stderr at exit (stdout carries the table)
struct PhaseTimer {
    bool on = false;
    std::vector<std::pair<std::string, double>> ms;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const std::string &name) {
        const auto t1 = std::chrono::steady_clock::now();
        const double d = std::chrono::duration<double, std::milli>(t1 - t0).count();
        t0 = t1;
        for (auto &kv : ms)
            if (kv.first == name) {
                kv.second += d;
                return;
            }
        ms.emplace_back(name, d);
    }
    void report() const {
        if (!on) return;
        double total = 0;
        std::cerr << "{\"phases_ms\": {";
        for (size_t i = 0; i < ms.size(); ++i) {
            std::cerr << (i ? ", " : "") << "\"" << ms[i].first << "\": " << ms[i].second;
            total += ms[i].second;
        }
        std::cerr << "}, \"total_ms\": " << total << "}\n";
    }
};
PhaseTimer g_phase;

struct Run {
    GraphStorage graph;
    GraphMask mask;
    std::vector<std::pair<uint64_t, std::string>> path_order;
};

// lean_ok: the caller only feeds the ItemTable to the device build (hist / growth / ordered growth / similarity); then,
// without edge counting and subset / exclude lists, the parser writes the u32 table directly (GraphStorage::lean)
Run load(const Args &a, const std::vector<CountType> &counts, bool with_order, bool with_names = false, bool lean_ok = false) {
    bool edges = false;
    for (auto c : counts) edges = edges || c == CountType::Edge;
    Run r;
    g_phase.lap("other");
    const bool lean = lean_ok && !edges && a.get("exclude").empty() && !getenv("PGX_NO_LEAN_PARSE");
    r.graph = GraphStorage::from_gfa(a.positional.at(0), edges, with_names, lean);
    g_phase.lap("gfa_parse");
    GraphMaskParameters p;
    p.groupby = a.get("groupby");
    p.groupby_sample = a.has("groupby-sample");  // commands/hist.rs:44-50: -S wins over -H, both over -g
    p.groupby_haplotype = !p.groupby_sample && a.has("groupby-haplotype");
    if (p.groupby_sample || p.groupby_haplotype) p.groupby.clear();
    p.positive_list = a.get("subset");
    p.negative_list = a.get("exclude");
    if (with_order && a.has("order")) p.order = a.get("order");
    r.mask = GraphMask::from_graph(r.graph, p);
    r.path_order = r.mask.get_path_order(r.graph.path_segments);
    g_phase.lap("grouping_order");
    if (lean && !r.graph.lean_apply_subset(r.mask)) {  // a path only partly inside the subset (BED intervals): the general parse
        r.graph = GraphStorage::from_gfa(a.positional.at(0), edges, with_names, false);
        g_phase.lap("gfa_parse");
    }
    return r;
}

std::vector<CountType> expand(CountType c) {
    if (c == CountType::All) return {CountType::Node, CountType::Bp, CountType::Edge};
    return {c};
}

uint32_t count_groups(const std::vector<std::pair<uint64_t, std::string>> &po) {
    uint32_t n = 0;
    const std::string *last = nullptr;
    for (auto &x : po) {
        if (!last || *last != x.second) ++n;
        last = &x.second;
    }
    return n;
}

// --gpus N (default 1): GPUs of this node the counting is sharded over (one host thread + one NCCL communicator each)
uint32_t n_gpus(const Args &a) {
    if (!a.has("gpus")) return 1;
    const int n = std::atoi(a.get("gpus").c_str());
    if (n < 1 || n > 8) throw Error("--gpus must be in 1..8");
    if (n > 1 && n > device_count()) throw Error("--gpus " + std::to_string(n) + ": only " + std::to_string(device_count()) + " CUDA devices visible");
    return (uint32_t)n;
}

std::vector<std::unique_ptr<DeviceComm>> make_comms(uint32_t n) {
    std::vector<int> devs(n);
    for (uint32_t r = 0; r < n; ++r) devs[r] = (int)r;
    return DeviceComm::create_all(devs);
}

// Item-range shards of `full` (which lives on GPU 0), one per GPU, cut on the device (NVLink peer copies)
std::vector<std::unique_ptr<DeviceAbacus>> shard_items(DeviceAbacus &full, uint32_t n) {
    std::vector<std::unique_ptr<DeviceAbacus>> shards;
    for (uint32_t r = 0; r < n; ++r) {
        const auto range = item_range(full.n_items(), r, n);
        shards.push_back(std::make_unique<DeviceAbacus>(range.second - range.first, full.n_groups(), (int)r));
        if (range.second > range.first) shards.back()->copy_rows_from(full, range.first);
    }
    return shards;
}

// One counted abacus on the device plus the host-side facts the analyses need around it; built from a GFA
// (Run) or read back from a packed-abacus cache file.
struct Counted {
    CountType count = CountType::Node;
    std::unique_ptr<DeviceAbacus> ab;
    std::vector<std::string> groups;          // counting order
    std::vector<uint32_t> weights;            // node_lens (node / bp), ones (edge)
    std::map<uint64_t, uint64_t> uncovered;   // uncovered_bps
};

struct Source {
    bool cached = false;
    std::string path;
    Run run;
};

bool ends_with(const std::string &s, const std::string &suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

Source open_source(const Args &a, const std::vector<CountType> &counts, bool with_order, bool with_names = false, bool lean_ok = false) {
    Source src;
    src.path = a.positional.at(0);
    src.cached = ends_with(src.path, ".pabm");
    if (src.cached) {
        for (const char *k : {"subset", "exclude", "groupby", "groupby-sample", "groupby-haplotype", "order"})
            if (a.has(k)) throw Error(std::string("--") + k + " cannot be combined with a packed-abacus (.pabm) input: grouping, "
                                      "subset and order are fixed when the cache is written");
    } else {
        src.run = load(a, counts, with_order, with_names, lean_ok);
    }
    return src;
}

Counted get_counted(const Source &src, CountType c, const Args &a) {
    Counted k;
    k.count = c;
    if (src.cached) {
        AbacusFile f = AbacusFile::load(src.path);
        if (f.count != c)
            throw Error("packed abacus " + src.path + " holds count type '" + to_string(f.count) + "', not '" + to_string(c) + "'");
        k.groups = std::move(f.groups);
        k.weights = std::move(f.weights);
        k.uncovered = std::move(f.uncovered);
        k.ab = std::make_unique<DeviceAbacus>(f.n_items, (uint32_t)k.groups.size());
        k.ab->upload(f.bitmap);
        return k;
    }
    const Run &r = src.run;
    g_phase.lap("other");
    ItemTables t = build_item_tables(r.graph, r.mask, c);
    g_phase.lap("item_table");
    const uint32_t G = count_groups(r.path_order);
    if (G == 0) throw Error("no path left to count (check --subset / --exclude)");
    k.ab = std::make_unique<DeviceAbacus>(t.n_items, G);
    g_phase.lap("device_init");
    k.ab->build(t, r.path_order, k.groups);
    g_phase.lap("h2d_build");
    k.weights = c == CountType::Edge ? std::vector<uint32_t>(t.n_items + 1, 1u) : r.graph.node_lens;
    if (c == CountType::Edge) k.weights[0] = 0;
    k.uncovered = std::move(t.uncovered_bps);
    if (a.has("save-abacus")) {  // <prefix>.<count>.pabm
        AbacusFile f;
        f.count = c;
        f.n_items = t.n_items;
        f.groups = k.groups;
        f.weights = k.weights;
        f.uncovered = k.uncovered;
        k.ab->download(f.bitmap);
        f.save(a.get("save-abacus") + "." + to_string(c) + ".pabm");
    }
    return k;
}

// Hist::from_abacus (graph_broker/hist.rs:39-49): coverage histogram of one count type
Hist device_hist(const Source &src, CountType c, const Args &a) {
    struct Teardown {
        ~Teardown() { g_phase.lap("device_teardown"); }
    } teardown;  // (declared first: laps after `k` released its handle -- cudaFree x 25, page-locked staging buffers)
    Counted k = get_counted(src, c, a);
    struct Lap {
        ~Lap() { g_phase.lap("kernels"); }
    } lap;
    Hist h;
    h.count = c;
    const uint32_t gpus = n_gpus(a);
    const bool bp = c == CountType::Bp;
    if (bp) k.ab->set_weights(k.weights);
    if (gpus > 1 && k.uncovered.empty()) {  // item ranges per GPU, ncclAllReduce of the per-shard histograms
        auto shards = shard_items(*k.ab, gpus);
        auto comms = make_comms(gpus);
        std::vector<std::vector<uint64_t>> part(gpus);
        run_on_devices(gpus, [&](uint32_t r) { shards[r]->hist_sharded(*comms[r], bp ? nullptr : &part[r], bp ? &part[r] : nullptr); });
        h.coverage = part[0];
    } else if (bp) {  // (with uncovered bps the patch below needs the per-item coverage: one GPU)
        std::vector<uint32_t> countable;
        k.ab->hist(nullptr, &h.coverage, k.uncovered.empty() ? nullptr : &countable);
        for (auto &kv : k.uncovered) {  // abacus.rs:779-785
            h.coverage[countable[kv.first]] -= kv.second;
            h.coverage[0] += kv.second;
        }
    } else {
        k.ab->hist(&h.coverage, nullptr, nullptr);
    }
    return h;
}

std::vector<double> as_f64(const std::vector<uint64_t> &v) { return std::vector<double>(v.begin(), v.end()); }

int cmd_hist(const Args &a, const std::string &cmdline, std::ostream &os) {
    const CountType count = count_type_from_str(a.get("count", "node"));
    const auto counts = expand(count);
    const Source src = open_source(a, counts, false, false, true);
    std::vector<std::vector<std::string>> headers = {{"panacus", "count", "", ""}};
    std::vector<std::vector<double>> cols;
    for (auto c : counts) {
        const Hist h = device_hist(src, c, a);
        cols.push_back(as_f64(h.coverage));
        headers.push_back({"hist", to_string(c), "", ""});
    }
    os << write_metadata_comments(cmdline, true) << write_table(headers, cols) << "\n";
    return 0;
}

std::string growth_table(const std::vector<Hist> &hists, const ThresholdContainer &aux, bool add_hist) {
    // analyses/growth.rs:33-102
    std::vector<std::vector<std::string>> headers = {{"panacus", "count", "coverage", "quorum"}};
    std::vector<std::vector<double>> cols;
    if (add_hist)
        for (auto &h : hists) {
            cols.push_back(as_f64(h.coverage));
            headers.push_back({"hist", to_string(h.count), "", ""});
        }
    g_phase.lap("other");
    for (auto &h : hists) {
        for (auto &g : h.calc_all_growths(aux)) cols.push_back(g);
        g_phase.lap("closed_form_growth");
        for (size_t k = 0; k < aux.coverage.size(); ++k)
            headers.push_back({"growth", to_string(h.count), aux.coverage[k].get_string(), aux.quorum[k].get_string()});
    }
    std::string out = write_table(headers, cols);
    g_phase.lap("tsv");
    return out;
}

int cmd_growth(const Args &a, const std::string &cmdline, bool histgrowth, std::ostream &os) {
    const ThresholdContainer aux = ThresholdContainer::parse_params(a.get("quorum", "0"), a.get("coverage", "1"));
    const std::string file = a.positional.at(0);
    if (!histgrowth && ends_with(file, ".tsv")) {  // src/lib.rs:160-190: growth from a hist table, no graph
        if (a.has("subset") || a.has("exclude") || a.has("groupby") || a.has("groupby-sample") || a.has("groupby-haplotype"))
            throw Error("subset, exclude and groupby can only be used in graph mode (with a .gfa or .gfa.gz file)");
        std::vector<std::string> comments;
        const std::vector<Hist> hists = parse_hists(file, comments);
        for (auto &c : comments) os << c << "\n";
        os << "# " << cmdline << "\n" << growth_table(hists, aux, a.has("hist")) << "\n";
        return 0;
    }
    // `growth <gfa>` has no --count: node (graph_broker.rs:158-160); histgrowth takes -c
    const CountType count = histgrowth ? count_type_from_str(a.get("count", "node")) : CountType::Node;
    const auto counts = expand(count);
    const Source src = open_source(a, counts, false, false, true);
    std::vector<Hist> hists;
    for (auto c : counts) hists.push_back(device_hist(src, c, a));
    os << "# " << cmdline << "\n" << growth_table(hists, aux, a.has("hist")) << "\n";
    return 0;
}

int cmd_ordered(const Args &a, const std::string &cmdline, std::ostream &os) {
    const CountType count = count_type_from_str(a.get("count", "node"));
    if (count == CountType::All) throw Error("ordered-histgrowth does not accept count type 'all'");
    const ThresholdContainer aux = ThresholdContainer::parse_params(a.get("quorum", "0"), a.get("coverage", "1"));
    const Source src = open_source(a, {count}, true, false, true);
    Counted k = get_counted(src, count, a);
    const std::vector<std::string> &groups = k.groups;
    if (count == CountType::Bp) {  // weights = node_lens - uncovered_bps (abacus.rs:1016-1023)
        std::vector<uint32_t> w = k.weights;
        for (auto &kv : k.uncovered) w[kv.first] = kv.second > w[kv.first] ? 0u : w[kv.first] - (uint32_t)kv.second;
        k.ab->set_weights(w);
    }
    std::vector<std::vector<double>> cols;
    g_phase.lap("other");
    const uint32_t gpus = n_gpus(a);
    if (gpus > 1) {  // item ranges per GPU; the path's one exchange is an ncclAllReduce of the KB-sized result vector
        auto shards = shard_items(*k.ab, gpus);
        auto comms = make_comms(gpus);
        std::vector<std::vector<std::vector<double>>> part(gpus);
        run_on_devices(gpus, [&](uint32_t r) { part[r] = shards[r]->calc_growth_sharded(*comms[r], aux, count == CountType::Bp); });
        cols = part[0];
    } else {
        cols = k.ab->calc_growth(aux, count == CountType::Bp);
    }
    g_phase.lap("kernels");
    for (auto &c : cols) c.insert(c.begin(), std::nan(""));  // io.rs:580-583
    std::vector<std::vector<std::string>> headers = {{"panacus", "count", "coverage", "quorum"}};
    for (size_t k = 0; k < aux.coverage.size(); ++k)
        headers.push_back({"ordered-growth", to_string(count), aux.coverage[k].get_string(), aux.quorum[k].get_string()});
    os << write_metadata_comments(cmdline, true) << write_ordered_table(headers, cols, groups) << "\n";
    return 0;
}

// ---- similarity: Jaccard + hierarchical clustering order (analyses/similarity.rs:119-254) ---------------------------
// The clustering itself (kodama 0.3.0 in the reference) lives in cluster.cpp.

int cmd_similarity(const Args &a, const std::string &cmdline, std::ostream &os) {
    const CountType count = count_type_from_str(a.get("count", "node"));
    if (count == CountType::All) throw Error("similarity does not accept count type 'all'");
    std::string method = a.get("method", "centroid");
    std::transform(method.begin(), method.end(), method.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    const Source src = open_source(a, {count}, false, false, true);
    Counted k = get_counted(src, count, a);
    const std::vector<std::string> &groups = k.groups;
    if (count == CountType::Bp) k.ab->set_weights(k.weights);  // no uncovered_bps correction here (similarity.rs:130-150)
    std::vector<uint64_t> inter, len;
    const uint32_t gpus = n_gpus(a);
    if (gpus > 1) {  // bitmap replicated over NVLink, upper-triangle row blocks per GPU, ncclAllGather
        const bool bp = count == CountType::Bp;
        std::vector<std::unique_ptr<DeviceAbacus>> rep(gpus);
        for (uint32_t r = 1; r < gpus; ++r) rep[r] = std::make_unique<DeviceAbacus>(k.ab->n_items(), k.ab->n_groups(), (int)r);
        auto comms = make_comms(gpus);
        std::vector<std::vector<uint64_t>> pi(gpus), pl(gpus);
        run_on_devices(gpus, [&](uint32_t r) {
            DeviceAbacus &ab = r == 0 ? *k.ab : *rep[r];
            ab.broadcast(*comms[r], 0, bp);
            ab.similarity_sharded(*comms[r], bp, pi[r], pl[r]);
        });
        inter.swap(pi[0]);
        len.swap(pl[0]);
    } else {
        k.ab->similarity(count == CountType::Bp, inter, len);
    }
    const size_t G = groups.size();
    std::vector<std::vector<float>> table(G, std::vector<float>(G));
    for (size_t i = 0; i < G; ++i)
        for (size_t j = 0; j < G; ++j)  // similarity.rs:153-163
            table[i][j] = (float)inter[i * G + j] / (float)(len[i] + len[j] - inter[i * G + j]);
    const std::vector<size_t> order = a.has("no-cluster") ? std::vector<size_t>() : cluster_leaf_order(table, method);
    std::vector<size_t> perm(G);
    for (size_t i = 0; i < G; ++i) perm[i] = i;
    if (order.size() == G) perm = order;
    std::string out = write_metadata_comments(cmdline, true);
    out += "group";
    for (size_t i = 0; i < G; ++i) out += "\t" + groups[perm[i]];
    out += "\n";
    for (size_t i = 0; i < G; ++i) {
        out += groups[perm[i]];
        for (size_t j = 0; j < G; ++j) out += "\t" + format_f32(table[perm[i]][perm[j]]);
        out += "\n";
    }
    os << out << "\n";
    return 0;
}

// `table` (src/commands/table.rs, analyses/table.rs:14-35 -> AbacusByGroup::to_tsv, abacus.rs:1056-1178): one row per
// node / edge, one column per group (occurrence count x bp length) or the number of groups containing it (--total).
// The reference registers the subcommand but has its CLI dispatch commented out (src/lib.rs:199-201); it is reachable
// through the YAML report (`!Table`), where `order` is never applied (analysis_parameter.rs:251-253) -- the CLI here
// honours -O as the subcommand's help text describes.
int cmd_table(const Args &a, const std::string &cmdline, bool with_order, std::ostream &os) {
    const CountType count = count_type_from_str(a.get("count", "node"));
    if (count == CountType::All) throw Error("table does not accept count type 'all'");
    if (ends_with(a.positional.at(0), ".pabm"))
        throw Error("table needs the graph itself (segment names and per-path occurrence counts), not a packed-abacus cache");
    const bool total = a.has("total") || a.has("hist");  // -a is --total here (commands/table.rs:18), --hist elsewhere
    const Run run = load(a, {count}, with_order, true);
    ItemTables t = build_item_tables(run.graph, run.mask, count);
    const uint32_t G = count_groups(run.path_order);
    if (G == 0) throw Error("no path left to count (check --subset / --exclude)");
    DeviceAbacus ab(t.n_items, G);
    std::vector<std::string> groups;
    ab.build(t, run.path_order, groups);
    std::vector<uint64_t> r, c;
    std::vector<uint32_t> v;
    ab.csr(t, run.path_order, r, c, v, !total);
    os << write_metadata_comments(cmdline, true)
       << abacus_by_group_to_tsv(run.graph, count, total, groups, r, c, v, t.uncovered_bps) << "\n";
    return 0;
}

// `panacus debug-parse <gfa> [-c count] [grouping / subset / exclude flags]`: wall time of the host front-end stages
// (GFA parse, grouping / ordering, ItemTable) without touching the device -- SURVEY 8f-2 is host work here.
int cmd_debug_parse(const Args &a, std::ostream &os) {
    const CountType count = count_type_from_str(a.get("count", "node"));
    if (count == CountType::All) throw Error("debug-parse takes one count type");
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto t0, auto t1) { return std::chrono::duration<double, std::milli>(t1 - t0).count(); };
    const auto t0 = now();
    Args b = a;
    const bool lean = a.has("lean") && count != CountType::Edge && a.get("exclude").empty();
    GraphStorage g = GraphStorage::from_gfa(a.positional.at(0), count == CountType::Edge, a.has("names"), lean);
    const auto t1 = now();
    GraphMaskParameters p;
    p.groupby = a.get("groupby");
    p.groupby_sample = a.has("groupby-sample");
    p.groupby_haplotype = !p.groupby_sample && a.has("groupby-haplotype");
    if (p.groupby_sample || p.groupby_haplotype) p.groupby.clear();
    p.positive_list = a.get("subset");
    p.negative_list = a.get("exclude");
    const GraphMask mask = GraphMask::from_graph(g, p);
    const auto order = mask.get_path_order(g.path_segments);
    if (lean && !g.lean_apply_subset(mask)) g = GraphStorage::from_gfa(a.positional.at(0), false, a.has("names"), false);
    const auto t2 = now();
    const ItemTables t = build_item_tables(g, mask, count);
    const auto t3 = now();
    const uint64_t steps = g.step_count();
    os << "nodes\t" << g.node_count() << "\nedges\t" << g.edge_count() << "\npaths\t" << g.path_segments.size() << "\nsteps\t" << steps
       << "\ngroups\t" << count_groups(order) << "\nitems\t" << t.n_steps << "\nparse_ms\t" << ms(t0, t1) << "\nmask_ms\t"
       << ms(t1, t2) << "\nitem_table_ms\t" << ms(t2, t3) << "\n";
    return 0;
}

// `panacus debug-table-tsv <gfa> --csr FILE [table flags]`: the to_tsv WRITER alone, fed with r / c / v read from a
// text file (three lines of TAB separated integers) instead of the device -- a test hook like debug-tables, so that the
// CPU test-suite can check the writer against the oracle's restatement without a GPU.  Nothing is counted here.
int cmd_debug_table_tsv(const Args &a, std::ostream &os) {
    const CountType count = count_type_from_str(a.get("count", "node"));
    const Run run = load(a, {count}, true, true);
    const ItemTables t = build_item_tables(run.graph, run.mask, count);
    std::vector<std::string> groups;
    for (auto &po : run.path_order)
        if (groups.empty() || groups.back() != po.second) groups.push_back(po.second);
    std::ifstream in(a.get("csr"));
    if (!in) throw Error("cannot open " + a.get("csr"));
    std::vector<std::vector<uint64_t>> rows(3);
    std::string line;
    for (int k = 0; k < 3 && std::getline(in, line); ++k) {
        std::stringstream ss(line);
        uint64_t x;
        while (ss >> x) rows[k].push_back(x);
    }
    const std::vector<uint32_t> v(rows[2].begin(), rows[2].end());
    os << abacus_by_group_to_tsv(run.graph, count, a.has("total"), groups, rows[0], rows[1], v, t.uncovered_bps);
    return 0;
}

// `panacus debug-growth <hist.tsv> -l .. -q ..`: the closed-form growth values of every hist column as C hex floats
// (%a: every bit of the f64), one line per threshold pair -- a test hook: the TSV prints floor(), which would hide a
// last-bit difference between the threaded host code and the oracle's statement-by-statement port.
// `panacus debug-linkage <file> -m method [--f32=1]`: the dendrogram of a condensed distance vector (first line: n, then
// the n (n - 1) / 2 distances, whitespace separated) as "cluster1 cluster2 height" lines (tests/test_cluster.py)
int cmd_debug_linkage(const Args &a, std::ostream &os) {
    std::ifstream in(a.positional.at(0));
    if (!in) throw Error("cannot open " + a.positional.at(0));
    size_t n = 0;
    in >> n;
    std::vector<double> cond;
    double x;
    while (in >> x) cond.push_back(x);
    if (n < 1 || cond.size() != n * (n - 1) / 2) throw Error("debug-linkage: expected n and n (n - 1) / 2 distances");
    os << debug_linkage(cond, n, a.get("method", "centroid"), a.has("f32"));
    return 0;
}

int cmd_debug_growth(const Args &a, std::ostream &os) {
    const ThresholdContainer aux = ThresholdContainer::parse_params(a.get("quorum", "0"), a.get("coverage", "1"));
    std::vector<std::string> comments;
    const std::vector<Hist> hists = parse_hists(a.positional.at(0), comments);
    char buf[64];
    for (auto &h : hists)
        for (auto &g : h.calc_all_growths(aux)) {
            os << to_string(h.count);
            for (size_t k = 1; k < g.size(); ++k) {
                snprintf(buf, sizeof buf, "\t%a", g[k]);
                os << buf;
            }
            os << "\n";
        }
    return 0;
}

// `panacus debug-synth-gfa <out.gfa> --nodes N [--samples 44 --haps 2 --contigs 22 --seed S --mean-len L --hist-file FILE]`:
// writes a pangenome-shaped GFA (synth.cpp).  FILE: two lines of samples + 1 numbers, the node and the bp coverage
// histogram to imitate (second line optional).
int cmd_debug_synth_gfa(const Args &a, std::ostream &os) {
    const uint64_t nodes = std::strtoull(a.get("nodes", "100000").c_str(), nullptr, 10);
    const uint32_t samples = (uint32_t)std::atoi(a.get("samples", "44").c_str());
    const uint32_t haps = (uint32_t)std::atoi(a.get("haps", "2").c_str());
    const uint32_t contigs = (uint32_t)std::atoi(a.get("contigs", "22").c_str());
    std::vector<double> node_hist, bp_hist;
    if (a.has("hist-file")) {  // (--hist is the -a flag of the growth subcommands)
        std::ifstream in(a.get("hist-file"));
        if (!in) throw Error("cannot open " + a.get("hist-file"));
        std::string line;
        for (int k = 0; k < 2 && std::getline(in, line); ++k) {
            std::stringstream ss(line);
            double x;
            while (ss >> x) (k ? bp_hist : node_hist).push_back(x);
        }
    } else {  // U shape: many private nodes, a long flat shell, a core peak
        node_hist.assign(samples + 1u, 1.0);
        node_hist[0] = 0.3;
        node_hist[1] = 0.25 * samples;
        if (samples >= 2) node_hist[2] = 0.12 * samples;
        node_hist[samples] = 0.2 * samples;
    }
    const uint64_t steps = synth_gfa(a.positional.at(0), nodes, samples, haps, contigs, node_hist, bp_hist,
                                     std::atof(a.get("mean-len", "100").c_str()), std::strtoull(a.get("seed", "22").c_str(), nullptr, 10));
    os << "nodes\t" << nodes << "\nsamples\t" << samples << "\npaths\t<= " << (uint64_t)samples * haps * contigs << "\nsteps\t" << steps << "\n";
    return 0;
}

// `panacus debug-dump-tables <gfa> --out PREFIX [-c count] [grouping flags]`: what the host front end hands to the device as
// raw little-endian files -- PREFIX.items.u64, PREFIX.prefsum.u64, PREFIX.path_group.i64, PREFIX.node_lens.u32 -- plus the
// group names on stdout.  Feeds the CPU oracle with the same tables at sizes where the text form of debug-tables is too big.
int cmd_debug_dump_tables(const Args &a, std::ostream &os) {
    const CountType count = count_type_from_str(a.get("count", "node"));
    if (count == CountType::All) throw Error("debug-dump-tables takes one count type");
    const auto t_start = std::chrono::steady_clock::now();
    const Run r = load(a, {count}, true, false, a.has("lean"));  // --lean: the direct u32 table (what the GPU commands use)
    const ItemTables t = build_item_tables(r.graph, r.mask, count);
    // front end = everything up to the ItemTable, without writing the dump (the CPU baseline of bench.py --workload c5)
    const double front_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    std::vector<std::string> groups;
    std::vector<int64_t> path_group(t.id_prefsum.size() - 1, -1);
    for (auto &po : r.path_order) {
        if (groups.empty() || groups.back() != po.second) groups.push_back(po.second);
        path_group[po.first] = (int64_t)groups.size() - 1;
    }
    auto dump = [&](const std::string &suffix, const void *p, size_t bytes) {
        std::ofstream o(a.get("out") + suffix, std::ios::binary);
        if (!o) throw Error("cannot write " + a.get("out") + suffix);
        o.write(static_cast<const char *>(p), (std::streamsize)bytes);
    };
    if (t.items32)
        dump(".items.u32", t.items32, t.n_steps * 4);
    else
        dump(".items.u64", t.items.data(), t.items.size() * 8);
    dump(".prefsum.u64", t.id_prefsum.data(), t.id_prefsum.size() * 8);
    dump(".path_group.i64", path_group.data(), path_group.size() * 8);
    dump(".node_lens.u32", r.graph.node_lens.data(), r.graph.node_lens.size() * 4);
    os << "n_items\t" << t.n_items << "\nsteps\t" << t.n_steps << "\nfront_ms\t" << front_ms << "\nitems_dtype\t" << (t.items32 ? "u32" : "u64")
       << "\ngroups";
    for (auto &g : groups) os << "\t" << g;
    os << "\n";
    return 0;
}

// CoverageLine (analyses/coverage_line.rs:23-57): the run's histograms without row 0, index starting at 1
int cmd_coverage_line(const Args &a, const std::string &cmdline, std::ostream &os) {
    const CountType count = count_type_from_str(a.get("count", "node"));
    const auto counts = expand(count);
    const Source src = open_source(a, counts, false, false, true);
    std::vector<std::vector<std::string>> headers = {{"panacus", "count", "", ""}};
    std::vector<std::vector<double>> cols;
    for (auto c : counts) {  // the reference iterates a HashMap here: column order across count types is unspecified
        const Hist h = device_hist(src, c, a);
        cols.emplace_back(h.coverage.begin() + 1, h.coverage.end());
        headers.push_back({"hist", to_string(c), "", ""});
    }
    os << write_metadata_comments(cmdline, true) << write_table(headers, cols, 1) << "\n";
    return 0;
}

// `panacus debug-tables <gfa> [-c count] [grouping / subset / exclude / order flags]`: dumps what the host front
// end hands to the device (group order, ItemTable, exclude flags, uncovered bps, node lengths) as text.  No GPU
// needed; the CPU test-suite compares it with the oracle's restatement of the reference front end.
int cmd_debug_tables(const Args &a, std::ostream &os) {
    const CountType count = count_type_from_str(a.get("count", "node"));
    if (count == CountType::All) throw Error("debug-tables takes one count type");
    const Run r = load(a, {count}, true, false, a.has("lean"));  // --lean: the parser's direct u32 table where it applies
    const ItemTables t = build_item_tables(r.graph, r.mask, count);
    std::vector<std::string> groups;
    os << "path_order";
    for (auto &po : r.path_order) {
        if (groups.empty() || groups.back() != po.second) groups.push_back(po.second);
        os << "\t" << po.first << ":" << groups.size() - 1;
    }
    os << "\ngroups";
    for (auto &g : groups) os << "\t" << g;
    os << "\nn_items\t" << t.n_items << "\nid_prefsum";
    for (auto v : t.id_prefsum) os << "\t" << v;
    os << "\nitems";
    if (t.items32)
        for (uint64_t k = 0; k < t.n_steps; ++k) os << "\t" << t.items32[k];
    else
        for (auto v : t.items) os << "\t" << v;
    os << "\nexclude";
    for (size_t i = 0; i < t.exclude.size(); ++i)
        if (t.exclude[i]) os << "\t" << i;
    os << "\nuncovered";
    for (auto &kv : t.uncovered_bps) os << "\t" << kv.first << ":" << kv.second;
    os << "\nnode_lens";
    for (auto v : r.graph.node_lens) os << "\t" << v;
    os << "\npaths";
    for (auto &p : r.graph.path_segments) os << "\t" << p.to_string();
    os << "\n";
    return 0;
}

// ---- report: the YAML front end (src/commands/report.rs, src/analysis_parameter.rs:82-258) -------------------------
// The reference renders an HTML / JSON report; rendering is outside the hot path, so this front end runs the
// same analyses on the same graph state and emits each analysis' TSV table under a "## run ... analysis ..." line.
struct AnalysisSpec {
    std::string type;  // Hist | Growth | OrderedGrowth | Similarity | (others: reported as unsupported)
    std::map<std::string, std::string> kv;
};
struct RunSpec {
    std::map<std::string, std::string> kv;  // graph, name, subset, exclude, grouping, grouping_file, nice
    std::vector<AnalysisSpec> analyses;
};

std::string yaml_scalar(std::string v) {
    size_t hash = std::string::npos;
    bool in_s = false, in_d = false;
    for (size_t i = 0; i < v.size(); ++i) {
        if (v[i] == '\'' && !in_d) in_s = !in_s;
        if (v[i] == '"' && !in_s) in_d = !in_d;
        if (v[i] == '#' && !in_s && !in_d && (i == 0 || v[i - 1] == ' ')) {
            hash = i;
            break;
        }
    }
    if (hash != std::string::npos) v = v.substr(0, hash);
    size_t a = 0, b = v.size();
    while (a < b && std::isspace((unsigned char)v[a])) ++a;
    while (b > a && std::isspace((unsigned char)v[b - 1])) --b;
    v = v.substr(a, b - a);
    if (v.size() >= 2 && ((v.front() == '"' && v.back() == '"') || (v.front() == '\'' && v.back() == '\''))) v = v.substr(1, v.size() - 2);
    if (v == "null" || v == "~") v.clear();
    return v;
}

void yaml_flow_map(const std::string &body, std::map<std::string, std::string> &kv) {  // {a: 1, b: "x,y"}
    std::string cur;
    std::vector<std::string> parts;
    bool in_q = false;
    for (char c : body) {
        if (c == '"' || c == '\'') in_q = !in_q;
        if (c == ',' && !in_q) {
            parts.push_back(cur);
            cur.clear();
        } else {
            cur += c;
        }
    }
    if (!cur.empty()) parts.push_back(cur);
    for (auto &part : parts) {
        const size_t colon = part.find(':');
        if (colon == std::string::npos) continue;
        kv[yaml_scalar(part.substr(0, colon))] = yaml_scalar(part.substr(colon + 1));
    }
}

std::vector<RunSpec> parse_report_yaml(const std::string &path) {
    std::ifstream in(path);
    if (!in) throw Error("cannot open " + path);
    std::vector<RunSpec> runs;
    std::string line;
    bool in_analyses = false;
    size_t lineno = 0;
    while (std::getline(in, line)) {
        ++lineno;
        while (!line.empty() && (line.back() == '\r')) line.pop_back();
        size_t indent = 0;
        while (indent < line.size() && line[indent] == ' ') ++indent;
        std::string body = line.substr(indent);
        if (body.empty() || body[0] == '#') continue;
        bool item = false;
        if (body.rfind("- ", 0) == 0 || body == "-") {
            item = true;
            body = body.size() > 2 ? body.substr(2) : "";
            indent += 2;
        }
        if (item && indent == 2) {  // a new run
            runs.emplace_back();
            in_analyses = false;
        }
        if (runs.empty()) throw Error("report YAML must be a list of runs (line " + std::to_string(lineno) + ")");
        RunSpec &run = runs.back();
        if (in_analyses && item) {  // "- !Hist" or "- !Hist {count_type: Bp}"
            if (body.empty() || body[0] != '!') throw Error("expected an analysis tag like !Hist at line " + std::to_string(lineno));
            AnalysisSpec spec;
            const size_t sp = body.find_first_of(" {");
            spec.type = body.substr(1, sp == std::string::npos ? std::string::npos : sp - 1);
            const size_t lb = body.find('{'), rb = body.rfind('}');
            if (lb != std::string::npos && rb != std::string::npos && rb > lb) yaml_flow_map(body.substr(lb + 1, rb - lb - 1), spec.kv);
            run.analyses.push_back(spec);
            continue;
        }
        const size_t colon = body.find(':');
        if (colon == std::string::npos) throw Error("cannot parse YAML line " + std::to_string(lineno) + ": " + line);
        const std::string key = yaml_scalar(body.substr(0, colon));
        const std::string val = yaml_scalar(body.substr(colon + 1));
        if (in_analyses && indent > 4 && !run.analyses.empty()) {  // block-style parameters of the last analysis
            run.analyses.back().kv[key] = val;
            continue;
        }
        in_analyses = false;
        if (key == "analyses") {
            in_analyses = true;
        } else if (key == "grouping" && val.rfind("!Custom", 0) == 0) {
            run.kv["grouping"] = "Custom";
            run.kv["grouping_file"] = yaml_scalar(val.substr(7));
        } else {
            run.kv[key] = val;
        }
    }
    return runs;
}

std::string lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return s;
}

int cmd_report(const Args &a0, std::ostream &os) {
    const std::vector<RunSpec> runs = parse_report_yaml(a0.positional.at(0));
    const bool dry = a0.has("dry-run");
    int rc = 0;
    for (size_t ri = 0; ri < runs.size(); ++ri) {
        const RunSpec &run = runs[ri];
        auto get = [&](const std::string &k) {
            auto it = run.kv.find(k);
            return it == run.kv.end() ? std::string() : it->second;
        };
        if (get("graph").empty()) throw Error("run " + std::to_string(ri) + " has no graph");
        // count type of the run's histograms: union of the !Hist requests; two or more -> all (graph_broker.rs:150-160)
        std::set<std::string> hist_counts;
        for (auto &an : run.analyses)
            if (an.type == "Hist") {
                auto it = an.kv.find("count_type");
                hist_counts.insert(lower(it == an.kv.end() ? "node" : it->second));
            }
        std::string run_count = "node";
        if (hist_counts.size() == 1) run_count = *hist_counts.begin();
        if (hist_counts.size() > 1) run_count = "all";
        Args base;
        base.positional.push_back(get("graph"));
        if (!get("subset").empty()) base.opt["subset"] = get("subset");
        if (!get("exclude").empty()) base.opt["exclude"] = get("exclude");
        if (get("grouping") == "Sample") base.opt["groupby-sample"] = "1";
        if (get("grouping") == "Haplotype") base.opt["groupby-haplotype"] = "1";
        if (get("grouping") == "Custom") base.opt["groupby"] = get("grouping_file");
        const std::string run_name = get("name").empty() ? get("graph") : get("name");
        for (auto &an : run.analyses) {
            Args a = base;
            auto kv = [&](const std::string &k, const std::string &d) {
                auto it = an.kv.find(k);
                return it == an.kv.end() || it->second.empty() ? d : it->second;
            };
            os << "## run " << run_name << " analysis " << an.type << "\n";
            const std::string cmdline = "panacus report " + a0.positional.at(0);
            if (an.type == "Hist") {
                a.sub = "hist";
                a.opt["count"] = run_count;
                if (!dry) rc |= cmd_hist(a, cmdline, os);
            } else if (an.type == "Growth") {
                a.sub = "histgrowth";
                a.opt["count"] = run_count;
                a.opt["coverage"] = kv("coverage", "1");
                a.opt["quorum"] = kv("quorum", "0");
                if (lower(kv("add_hist", "false")) == "true") a.opt["hist"] = "1";
                if (!dry) rc |= cmd_growth(a, cmdline, true, os);
            } else if (an.type == "OrderedGrowth") {
                a.sub = "ordered-histgrowth";
                a.opt["count"] = lower(kv("count_type", "node"));
                a.opt["coverage"] = kv("coverage", "1");
                a.opt["quorum"] = kv("quorum", "0");
                if (!kv("order", "").empty()) a.opt["order"] = kv("order", "");
                if (!dry) rc |= cmd_ordered(a, cmdline, os);
            } else if (an.type == "Table") {
                a.sub = "table";
                a.opt["count"] = lower(kv("count_type", "node"));
                if (lower(kv("total", "false")) == "true") a.opt["total"] = "1";
                if (!dry) rc |= cmd_table(a, cmdline, false, os);  // `order` is not applied by the reference here
            } else if (an.type == "CoverageLine") {
                a.sub = "coverage-line";
                a.opt["count"] = run_count;  // prints the run's histograms (coverage_line.rs:39-47)
                if (!dry) rc |= cmd_coverage_line(a, cmdline, os);
            } else if (an.type == "Similarity") {
                a.sub = "similarity";
                a.opt["count"] = lower(kv("count_type", "node"));
                a.opt["method"] = lower(kv("cluster_method", "centroid"));
                if (!dry) rc |= cmd_similarity(a, cmdline, os);
            } else {
                os << "# analysis " << an.type << " is outside the accelerated hot path: not supported by this build\n";
                continue;
            }
            if (dry) {
                os << "# would run: " << a.sub << " " << a.positional[0];
                for (auto &o : a.opt) os << " --" << o.first << (o.second == "1" && kFlags.count(o.first) ? "" : " " + o.second);
                os << "\n";
            }
        }
    }
    return rc;
}

void usage() {
    std::cerr << "panacus (B200 hot path) -- usage: panacus <hist|growth|histgrowth|ordered-histgrowth|similarity|table|coverage-line> <GFA_FILE> [options]\n"
                 "                                  panacus report <config.yaml> [--dry-run]   (tables as TSV; no HTML rendering)\n"
                 "  -s, --subset FILE   -e, --exclude FILE   -g, --groupby FILE   -H, --groupby-haplotype   -S, --groupby-sample\n"
                 "  -c, --count node|bp|edge|all   -l, --coverage LIST   -q, --quorum LIST   -a, --hist   -O, --order FILE\n"
                 "  -m, --method single|complete|average|weighted|ward|centroid|median (similarity)   -t, --threads N\n"
                 "  --gpus N   shard the counting over N GPUs of this node (hist / ordered-histgrowth: item ranges; similarity: row blocks)\n"
                 "  --save-abacus PREFIX   write PREFIX.<count>.pabm (packed abacus); a .pabm file is accepted in place of the GFA\n";
}

int dispatch(int argc, char **argv, std::ostream &os) {
    const Args a = parse_args(argc, argv);
    if (a.sub.empty() || a.positional.empty()) {
        usage();
        return 2;
    }
    const std::string cmdline = argv_joined(argc, argv);
    if (a.has("threads")) set_host_threads(std::atoi(a.get("threads").c_str()));
    g_phase.on = a.has("timing");
    struct Report {
        ~Report() {
            g_phase.lap("other");
            g_phase.report();
        }
    } report;
    {  // the subcommands that count on the GPU: start creating the CUDA context(s) now, while the GFA is parsed
        static const std::set<std::string> kDevice = {"hist", "histgrowth", "ordered-histgrowth", "similarity", "table", "coverage-line", "report"};
        const bool from_tsv = a.sub == "growth" && ends_with(a.positional.at(0), ".tsv");
        if ((kDevice.count(a.sub) || (a.sub == "growth" && !from_tsv)) && !a.has("dry-run"))
            device_warmup_async(a.has("gpus") ? std::max(1, std::atoi(a.get("gpus").c_str())) : 1);
    }
    if (a.sub == "hist") return cmd_hist(a, cmdline, os);
    if (a.sub == "growth") return cmd_growth(a, cmdline, false, os);
    if (a.sub == "histgrowth") return cmd_growth(a, cmdline, true, os);
    if (a.sub == "ordered-histgrowth") return cmd_ordered(a, cmdline, os);
    if (a.sub == "similarity") return cmd_similarity(a, cmdline, os);
    if (a.sub == "table") return cmd_table(a, cmdline, true, os);
    if (a.sub == "coverage-line") return cmd_coverage_line(a, cmdline, os);
    if (a.sub == "debug-table-tsv") return cmd_debug_table_tsv(a, os);
    if (a.sub == "debug-parse") return cmd_debug_parse(a, os);
    if (a.sub == "debug-growth") return cmd_debug_growth(a, os);
    if (a.sub == "debug-linkage") return cmd_debug_linkage(a, os);
    if (a.sub == "debug-synth-gfa") return cmd_debug_synth_gfa(a, os);
    if (a.sub == "debug-dump-tables") return cmd_debug_dump_tables(a, os);
    if (a.sub == "debug-tables") return cmd_debug_tables(a, os);
    if (a.sub == "report") return cmd_report(a, os);
    usage();
    return 2;
}

// `panacus batch <file>`: one command line per row of <file> (whitespace separated, no quoting), all in
// one process so the CUDA context is created once.  Output: "## batch <i> rc=<code>" then the command's
// stdout.  Not part of the reference CLI; used by the test-suite and for scripted runs.
int run_batch(const char *prog, const std::string &file) {
    std::ifstream in(file);
    if (!in) throw Error("cannot open " + file);
    std::string line;
    int idx = 0, worst = 0;
    while (std::getline(in, line)) {
        std::vector<std::string> tok = {prog};
        std::stringstream ss(line);
        std::string t;
        while (ss >> t) tok.push_back(t);
        if (tok.size() == 1) continue;
        std::vector<char *> av;
        for (auto &x : tok) av.push_back(const_cast<char *>(x.c_str()));
        std::ostringstream out;
        int rc;
        try {
            rc = dispatch((int)av.size(), av.data(), out);
        } catch (const std::exception &e) {
            rc = 1;
            out << "error: " << e.what() << "\n";
        }
        std::cout << "## batch " << idx++ << " rc=" << rc << "\n" << out.str();
        worst = std::max(worst, rc);
    }
    return worst;
}

}  // namespace

int main(int argc, char **argv) {
    int rc;
    try {
        if (argc == 3 && std::string(argv[1]) == "batch") return run_batch(argv[0], argv[2]);
        rc = dispatch(argc, argv, std::cout);
    } catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << "\n";
        std::cerr.flush();
        std::_Exit(1);  // (a context-creation side thread may still be running: no static teardown under it)
    }
    // results are out: skip the teardown of the CUDA context and of GB-sized host structures (fractions of a second each)
    std::cout.flush();
    std::cerr.flush();
    std::_Exit(rc);
}
