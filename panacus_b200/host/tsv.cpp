// tsv.cpp -- table writers and the hist-TSV reader (src/io.rs:152-290, 460-518, 546-555).
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>

#include "panacus_host.hpp"

namespace panacus {

namespace {

// format!("{:0}", x.floor()) of an f64 (io.rs:484, 512): integral values print without a fraction
std::string floor_str(double v) {
    if (std::isnan(v)) return "NaN";
    const double f = std::floor(v);
    if (std::isinf(f)) return f > 0 ? "inf" : "-inf";
    char buf[400];
    snprintf(buf, sizeof buf, "%.0f", f);
    return buf;
}

void write_headers(std::string &res, const std::vector<std::vector<std::string>> &headers) {
    const size_t n = headers.empty() ? 0 : headers[0].size();
    for (size_t i = 0; i < n; ++i) {
        for (size_t j = 0; j < headers.size(); ++j) {
            if (j) res += '\t';
            res += headers[j][i];
        }
        res += '\n';
    }
}

}  // namespace

std::string write_table(const std::vector<std::vector<std::string>> &headers,
                        const std::vector<std::vector<double>> &columns, uint64_t start_index) {  // io.rs:464-489
    std::string res;
    write_headers(res, headers);
    const size_t n = columns.empty() ? 0 : columns[0].size();
    for (size_t i = 0; i < n; ++i) {
        res += std::to_string(i + start_index);
        for (auto &col : columns) {
            res += '\t';
            res += floor_str(col[i]);
        }
        res += '\n';
    }
    return res;
}

std::string write_ordered_table(const std::vector<std::vector<std::string>> &headers,
                                const std::vector<std::vector<double>> &columns,
                                const std::vector<std::string> &index) {  // io.rs:491-518: row 0 (NaN) is skipped
    std::string res;
    write_headers(res, headers);
    const size_t n = columns.empty() ? 0 : columns[0].size();
    for (size_t i = 1; i < n; ++i) {
        res += index[i - 1];
        for (auto &col : columns) {
            res += '\t';
            res += floor_str(col[i]);
        }
        res += '\n';
    }
    return res;
}

std::string write_metadata_comments(const std::string &argv_joined, bool with_version) {  // io.rs:546-555
    std::string res = "# " + argv_joined + "\n";
    if (with_version) res += "# version 0.4.1\n";
    return res;
}

// AbacusByGroup::to_tsv (abacus.rs:1056-1178)
std::string abacus_by_group_to_tsv(const GraphStorage &g, CountType count, bool total, const std::vector<std::string> &groups,
                                   const std::vector<uint64_t> &r, const std::vector<uint64_t> &c, const std::vector<uint32_t> &v,
                                   const std::map<uint64_t, uint64_t> &uncovered_bps) {
    if (g.node_names.size() != g.node_lens.size()) throw Error("table: the graph was loaded without segment names");
    const uint64_t G = groups.size();
    std::string out;
    auto header = [&](const char *first) {
        out += first;
        if (total) {
            out += "\ttotal";
        } else {
            for (auto &grp : groups) {
                out += '\t';
                out += grp;
            }
        }
        out += '\n';
    };
    if (count == CountType::Node || count == CountType::Bp) {
        header("node");
        for (uint64_t i = 1; i + 1 < r.size(); ++i) {  // windows of r, first entry ignored (abacus.rs:1086-1088)
            const uint64_t start = r[i], end = r[i + 1];
            uint64_t bp = 1;
            if (count == CountType::Bp) {
                auto it = uncovered_bps.find(i);
                bp = (uint64_t)g.node_lens[i] - (it == uncovered_bps.end() ? 0 : it->second);
            }
            out += g.node_names[i];
            if (total) {
                out += '\t';
                out += std::to_string(end - start);
                out += '\n';
                continue;
            }
            uint64_t k = start;
            for (uint64_t j = 0; j < G; ++j) {
                if (k == end || j < c[k]) {
                    out += "\t0";
                } else if (j == c[k]) {
                    out += '\t';
                    out += std::to_string(v.empty() ? bp : (uint64_t)v[k] * bp);
                    ++k;
                }
            }
            out += '\n';
        }
        return out;
    }
    if (count != CountType::Edge) throw Error("inadmissible count type");
    if (!g.has_edges) return out;  // abacus.rs:1121: no edge2id, nothing is written
    std::vector<uint64_t> id2edge(g.edge_count() + 1, 0);
    g.edge2id.for_each([&](uint64_t key, uint32_t id) { id2edge[id] = key; });
    header("edge");
    for (uint64_t i = 1; i + 1 < r.size(); ++i) {
        const uint64_t start = r[i], end = r[i + 1];
        const uint64_t key = id2edge[i];
        const uint32_t a = (uint32_t)(key >> 32), b = (uint32_t)key;
        out += (a & 1u) ? '>' : '<';
        out += g.node_names[a >> 1];
        out += (b & 1u) ? '>' : '<';
        out += g.node_names[b >> 1];
        if (total) {
            out += '\t';
            out += std::to_string(end - start);
            out += '\n';
            continue;
        }
        uint64_t k = start;
        for (uint64_t j = 0; j < G; ++j) {
            if (k == end || j < c[k]) {
                out += "\t0";
            } else if (j == c[k]) {
                // the reference indexes the occurrence counts by the GROUP index here (`v[j as usize]`, abacus.rs:1166),
                // not by the row cursor k; reproduced as is, including the out-of-bounds panic
                if (v.empty()) {
                    out += "\t1";
                } else {
                    if (j >= v.size()) throw Error("index out of bounds: the len is " + std::to_string(v.size()) + " but the index is " + std::to_string(j));
                    out += '\t';
                    out += std::to_string(v[j]);
                }
                ++k;
            }
        }
        out += '\n';
    }
    return out;
}

std::vector<Hist> parse_hists(const std::string &path, std::vector<std::string> &comments) {  // io.rs:152-290
    std::ifstream in(path);
    if (!in) throw Error("cannot open " + path);
    std::vector<std::vector<std::string>> table;
    std::string line;
    while (std::getline(in, line)) {
        while (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        std::vector<std::string> row;
        std::stringstream ss(line);
        std::string cell;
        while (std::getline(ss, cell, '\t')) row.push_back(cell);
        if (!line.empty() && line.back() == '\t') row.emplace_back();
        if (row.empty()) continue;
        if (!row[0].empty() && row[0][0] == '#') {
            comments.push_back(line);
            continue;
        }
        bool all_empty = true;
        for (auto &c : row) all_empty = all_empty && c.empty();
        if (all_empty) continue;  // header rows 3-4 of a hist-only table
        table.push_back(row);
    }
    if (table.size() < 3 || table[0].empty() || table[0][0] != "panacus")
        throw Error("error in line " + std::to_string(comments.size()) + ": table appears not to be generated by panacus");
    const size_t ncol = table[0].size();
    auto parse_col = [&](size_t col) {
        std::vector<uint64_t> out;
        for (size_t r = 2; r < table.size(); ++r) {
            const std::string &e = col < table[r].size() ? table[r][col] : std::string();
            if (e.empty() || e.find_first_not_of("0123456789") != std::string::npos)
                throw Error("error in line " + std::to_string(r + 1 + comments.size()) + ": value must be integer, but is '" + e + "'");
            out.push_back(std::stoull(e));
        }
        return out;
    };
    const std::vector<uint64_t> index = parse_col(0);
    uint64_t mx = 0;
    for (auto v : index) mx = std::max(mx, v);
    std::vector<Hist> res;
    for (size_t col = 1; col < ncol; ++col) {
        if (table[0][col] != "hist") continue;
        Hist h;
        h.count = count_type_from_str(col < table[1].size() ? table[1][col] : "");
        h.coverage.assign(mx + 1, 0);
        const std::vector<uint64_t> vals = parse_col(col);
        for (size_t k = 0; k < index.size(); ++k) h.coverage[index[k]] = vals[k];
        res.push_back(std::move(h));
    }
    if (res.empty()) throw Error("table does not contain hist columns");
    return res;
}

// ---- .pabm cache file ---------------------------------------------------------------------------------------------
namespace {
const char kMagic[8] = {'P', 'A', 'B', 'M', '1', '\n', 0, 0};
template <typename T>
void put(std::ofstream &o, const T &v) {
    o.write(reinterpret_cast<const char *>(&v), sizeof(T));
}
template <typename T>
T get(std::ifstream &i) {
    T v;
    i.read(reinterpret_cast<char *>(&v), sizeof(T));
    if (!i) throw Error("truncated abacus cache file");
    return v;
}
}  // namespace

void AbacusFile::save(const std::string &path) const {
    std::ofstream o(path, std::ios::binary);
    if (!o) throw Error("cannot write " + path);
    o.write(kMagic, 8);
    put<uint32_t>(o, (uint32_t)count);
    put<uint64_t>(o, n_items);
    put<uint32_t>(o, (uint32_t)groups.size());
    put<uint64_t>(o, (uint64_t)uncovered.size());
    for (auto &g : groups) {
        put<uint32_t>(o, (uint32_t)g.size());
        o.write(g.data(), (std::streamsize)g.size());
    }
    o.write(reinterpret_cast<const char *>(weights.data()), (std::streamsize)(weights.size() * 4));
    for (auto &kv : uncovered) {
        put<uint64_t>(o, kv.first);
        put<uint64_t>(o, kv.second);
    }
    o.write(reinterpret_cast<const char *>(bitmap.data()), (std::streamsize)(bitmap.size() * 8));
    if (!o) throw Error("write error on " + path);
}

AbacusFile AbacusFile::load(const std::string &path) {
    std::ifstream i(path, std::ios::binary);
    if (!i) throw Error("cannot open " + path);
    char magic[8];
    i.read(magic, 8);
    if (!i || std::string(magic, 6) != std::string(kMagic, 6)) throw Error(path + " is not a packed-abacus (.pabm) file");
    AbacusFile f;
    const uint32_t c = get<uint32_t>(i);
    if (c > 2) throw Error("bad count type in " + path);
    f.count = (CountType)c;
    f.n_items = get<uint64_t>(i);
    const uint32_t G = get<uint32_t>(i);
    const uint64_t n_unc = get<uint64_t>(i);
    if (!i || G == 0 || G > (1u << 20) || f.n_items >= (1ull << 32) - 2 || n_unc > f.n_items)
        throw Error("bad shape in " + path);
    for (uint32_t g = 0; g < G; ++g) {
        const uint32_t len = get<uint32_t>(i);
        if (!i || len > (1u << 16)) throw Error("bad group name in " + path);  // never trust a length from the file
        std::string name(len, '\0');
        i.read(name.data(), len);
        if (!i) throw Error("truncated abacus cache file " + path);
        f.groups.push_back(std::move(name));
    }
    f.weights.resize(f.n_items + 1);
    i.read(reinterpret_cast<char *>(f.weights.data()), (std::streamsize)(f.weights.size() * 4));
    if (!i) throw Error("truncated abacus cache file " + path);
    for (uint64_t k = 0; k < n_unc; ++k) {
        const uint64_t id = get<uint64_t>(i), v = get<uint64_t>(i);
        // the ids index per-item arrays later (hist patch, weight correction): they must name an item of this table
        if (!i || id == 0 || id > f.n_items) throw Error("bad uncovered-bps entry in " + path);
        f.uncovered[id] = v;
    }
    f.bitmap.resize((size_t)(f.n_items + 1) * ((G + 63u) / 64u));
    i.read(reinterpret_cast<char *>(f.bitmap.data()), (std::streamsize)(f.bitmap.size() * 8));
    if (!i) throw Error("truncated abacus cache file " + path);
    return f;
}

}  // namespace panacus
