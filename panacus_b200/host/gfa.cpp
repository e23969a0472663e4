// gfa.cpp -- GFA front end, PanSN path names, grouping / ordering, subset / exclude bookkeeping.
// Semantics follow the reference (marschall-lab/panacus @ 395ba41); every function cites what it mirrors.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <regex>
#include <set>
#include <sstream>
#include <memory>
#include <mutex>
#include <string_view>
#include <thread>

#include "panacus_host.hpp"

namespace panacus {

namespace {

constexpr uint64_t kUsizeMax = ~0ull;

// The whole file as one read-only byte range.  Plain files are mapped (no copy: the page cache is the buffer; reading
// 1.25 GB into a zero-filled std::string was 0.5-1.2 s of the chr22-sized parse); gzip files are inflated into memory
// (transparently, like io.rs:23-33) -- bgzip-written ones block by block on the worker threads, see load_file.
struct FileData {
    const char *ptr = nullptr;
    size_t len = 0;
    std::string owned;
    std::unique_ptr<char[]> heap;  // inflated BGZF blocks (not zero-filled first)
    void *map = nullptr;
    FileData() = default;
    FileData(const FileData &) = delete;
    FileData &operator=(const FileData &) = delete;
    ~FileData() {
        if (map) munmap(map, len);
    }
};

std::vector<std::string> split(const std::string &s, char sep) {
    std::vector<std::string> out;
    size_t i = 0;
    for (;;) {
        const size_t j = s.find(sep, i);
        if (j == std::string::npos) {
            out.push_back(s.substr(i));
            break;
        }
        out.push_back(s.substr(i, j - i));
        i = j + 1;
    }
    return out;
}

bool all_digits(const std::string &s) { return !s.empty() && std::all_of(s.begin(), s.end(), [](char c) { return c >= '0' && c <= '9'; }); }

// ^(.+):([0-9]+)-([0-9]+)$  (graph.rs:17)
bool match_coords(const std::string &s, std::string &name, uint64_t &a, uint64_t &b) {
    const size_t colon = s.rfind(':');
    if (colon == std::string::npos || colon == 0) return false;
    const std::string tail = s.substr(colon + 1);
    const size_t dash = tail.find('-');
    if (dash == std::string::npos) return false;
    const std::string x = tail.substr(0, dash), y = tail.substr(dash + 1);
    if (!all_digits(x) || !all_digits(y)) return false;
    name = s.substr(0, colon);
    a = std::stoull(x);
    b = std::stoull(y);
    return true;
}

// ^([^#]+)(#[^#]+)?(#[^#].*)?$  (graph.rs:16) -> the matched groups, or empty if no match
std::vector<std::string> match_pansn(const std::string &s) {
    std::vector<std::string> segs;
    size_t i = 0;
    while (i < s.size() && s[i] != '#') ++i;
    if (i == 0) return {};
    segs.push_back(s.substr(0, i));
    if (i == s.size()) return segs;
    // group 2: '#' + one or more non-'#'
    size_t j = i + 1;
    while (j < s.size() && s[j] != '#') ++j;
    if (j == i + 1) return {};  // "##" or trailing '#': neither group 2 nor group 3 can start here
    segs.push_back(s.substr(i, j - i));
    if (j == s.size()) return segs;
    // group 3: '#' + a non-'#' + anything
    if (j + 1 >= s.size() || s[j + 1] == '#') return {};
    segs.push_back(s.substr(j));
    return segs;
}

void merge_intervals(std::vector<std::pair<uint64_t, uint64_t>> &v) {
    std::sort(v.begin(), v.end());
    size_t i = 1;
    while (i < v.size()) {
        if (v[i - 1].second >= v[i].first) {
            v[i - 1].second = std::max(v[i - 1].second, v[i].second);
            v.erase(v.begin() + (long)i);
        } else {
            ++i;
        }
    }
}

using Intervals = std::vector<std::pair<uint64_t, uint64_t>>;

bool intersects(const Intervals &v, std::pair<uint64_t, uint64_t> el) {  // util.rs:370-383
    for (auto &iv : v)
        if (iv.first <= el.second && iv.second >= el.first) return true;
    return false;
}
bool is_contained(const Intervals &v, std::pair<uint64_t, uint64_t> el) {  // util.rs:385-398
    for (auto &iv : v)
        if (iv.first <= el.first && iv.second >= el.second) return true;
    return false;
}

struct IntervalContainer {  // util.rs:199-310
    std::map<uint64_t, Intervals> map;
    void add(uint64_t id, uint64_t a, uint64_t b) {
        auto &v = map[id];
        v.emplace_back(a, b);
        merge_intervals(v);
    }
    const Intervals *get(uint64_t id) const {
        auto it = map.find(id);
        return it == map.end() ? nullptr : &it->second;
    }
    bool contains(uint64_t id) const { return map.count(id) != 0; }
    void remove(uint64_t id) { map.erase(id); }
    uint64_t total_coverage(uint64_t id, const Intervals *exclude) const {  // util.rs:257-298 (wrapping usize)
        const Intervals *iv = get(id);
        if (!iv) return 0;
        uint64_t res = 0;
        if (!exclude) {
            for (auto &x : *iv) res += x.second - x.first;
            return res;
        }
        size_t i = 0;
        for (auto &x : *iv) {
            const uint64_t start = x.first, end = x.second;
            while (i < exclude->size() && (*exclude)[i].second <= start) ++i;
            if (i < exclude->size() && (*exclude)[i].first < end) {
                res += std::min((*exclude)[i].first - 1, end) - start;
                if ((*exclude)[i].second < end) res += end - (*exclude)[i].second + 1;
            } else {
                res += end - start;
            }
        }
        return res;
    }
};

struct ActiveTable {  // util.rs:117-197
    std::vector<uint8_t> items;
    bool with_annotation;
    IntervalContainer annotation;
    ActiveTable(size_t n, bool ann) : items(n, 0), with_annotation(ann) {}
    void activate(uint64_t id) { items[id] = 1; }
    void activate_n_annotate(uint64_t id, uint64_t item_len, uint64_t start, uint64_t end) {
        if (end - start == item_len) {
            items[id] = 1;
            annotation.remove(id);
        } else {
            if (start <= end) annotation.add(id, start, end);
            const Intervals *iv = annotation.get(id);
            if (iv && !iv->empty() && (*iv)[0] == std::make_pair<uint64_t, uint64_t>(0, (uint64_t)item_len)) {
                annotation.remove(id);
                items[id] = 1;
            }
        }
    }
    Intervals get_active_intervals(uint64_t id, uint64_t item_len) const {
        if (items[id]) return {{0, item_len}};
        if (with_annotation) {
            const Intervals *iv = annotation.get(id);
            if (iv) return *iv;
        }
        return {};
    }
};

std::map<std::string, Intervals> build_subpath_map(const std::vector<PathSegment> &segs) {  // abacus.rs:354-383
    std::map<std::string, std::set<std::pair<uint64_t, uint64_t>>> tmp;
    for (auto &x : segs) {
        auto c = x.coords();
        tmp[x.id()].insert(c ? *c : std::make_pair<uint64_t, uint64_t>(0, (uint64_t)kUsizeMax));
    }
    std::map<std::string, Intervals> out;
    for (auto &kv : tmp) {
        Intervals v(kv.second.begin(), kv.second.end());
        merge_intervals(v);
        out[kv.first] = v;
    }
    return out;
}

uint64_t canonical_edge(uint32_t u, bool f1, uint32_t v, bool f2) {  // graph.rs:142-148
    if (u > v || (u == v && !f1)) return GraphStorage::edge_key(v, !f2, u, !f1);
    return GraphStorage::edge_key(u, f1, v, f2);
}

}  // namespace

// ---- CountType / PathSegment ---------------------------------------------------------------------------

std::string to_string(CountType c) {
    switch (c) {
        case CountType::Node: return "node";
        case CountType::Bp: return "bp";
        case CountType::Edge: return "edge";
        default: return "all";
    }
}

CountType count_type_from_str(const std::string &s0) {
    std::string s = s0;
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    if (s == "node") return CountType::Node;
    if (s == "bp") return CountType::Bp;
    if (s == "edge") return CountType::Edge;
    if (s == "all") return CountType::All;
    throw Error("invalid count type '" + s0 + "' (expected node, bp, edge or all)");
}

PathSegment PathSegment::from_str(const std::string &s) {  // graph.rs:495-549
    PathSegment p;
    p.sample = s;
    const std::vector<std::string> segs = match_pansn(s);
    std::string name;
    uint64_t a, b;
    if (segs.size() == 3) {
        p.sample = segs[0];
        p.haplotype = segs[1].substr(1);
        const std::string rest = segs[2].substr(1);
        if (match_coords(rest, name, a, b)) {
            p.seqid = name;
            p.start = a;
            p.end = b;
        } else {
            p.seqid = rest;
        }
    } else if (segs.size() == 2) {
        p.sample = segs[0];
        const std::string rest = segs[1].substr(1);
        if (match_coords(rest, name, a, b)) {
            p.haplotype = name;
            p.start = a;
            p.end = b;
        } else {
            p.haplotype = rest;
        }
    } else if (segs.size() == 1) {
        if (match_coords(segs[0], name, a, b)) {
            p.sample = name;
            p.start = a;
            p.end = b;
        }
    }
    return p;
}

PathSegment PathSegment::from_str_start_end(const std::string &s, uint64_t a, uint64_t b) {
    PathSegment p = from_str(s);
    p.start = a;
    p.end = b;
    return p;
}

std::string PathSegment::id() const {  // graph.rs:558-579
    if (haplotype) return sample + "#" + *haplotype + (seqid ? "#" + *seqid : "");
    if (seqid) return sample + "#*#" + *seqid;
    return sample;
}

PathSegment PathSegment::clear_coords() const {
    PathSegment p = *this;
    p.start.reset();
    p.end.reset();
    return p;
}

std::optional<std::pair<uint64_t, uint64_t>> PathSegment::coords() const {
    if (start && end) return std::make_pair(*start, *end);
    return std::nullopt;
}

std::string PathSegment::to_string() const {
    auto c = coords();
    if (c) return id() + ":" + std::to_string(c->first) + "-" + std::to_string(c->second);
    return id();
}

bool PathSegment::operator<(const PathSegment &o) const {
    return std::tie(sample, haplotype, seqid, start, end) < std::tie(o.sample, o.haplotype, o.seqid, o.start, o.end);
}
bool PathSegment::operator==(const PathSegment &o) const {
    return std::tie(sample, haplotype, seqid, start, end) == std::tie(o.sample, o.haplotype, o.seqid, o.start, o.end);
}

// ---- GraphStorage::from_gfa (graph.rs:195-375, util.rs:368-410, 916-931, 1093-1142) ---------------------------

uint64_t GraphStorage::edge_key(uint32_t u, bool fu, uint32_t v, bool fv) {
    return ((uint64_t)((u << 1) | (fu ? 1u : 0u)) << 32) | (uint64_t)((v << 1) | (fv ? 1u : 0u));
}

namespace {

int g_host_threads = 0;  // 0 = hardware concurrency

// fn(k) for k in [0, n) on `nthreads` threads, work handed out through an atomic counter (the units differ a lot in
// size); the first exception stops the hand-out and is rethrown on the caller's thread
template <typename F>
void parallel_for(size_t n, unsigned nthreads, F fn) {
    nthreads = std::max(1u, std::min<unsigned>({nthreads, 32u, (unsigned)std::max<size_t>(n, 1)}));
    if (nthreads == 1) {
        for (size_t k = 0; k < n; ++k) fn(k);
        return;
    }
    std::atomic<size_t> next{0};
    std::mutex err_mu;
    std::string err;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; ++t)
        pool.emplace_back([&] {
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= n) return;
                try {
                    fn(k);
                } catch (const std::exception &e) {
                    std::lock_guard<std::mutex> lock(err_mu);
                    if (err.empty()) err = e.what();
                    next.store(n);
                    return;
                }
            }
        });
    for (auto &th : pool) th.join();
    if (!err.empty()) throw Error(err);
}

// -t N is taken literally; the default is one thread per core for inputs worth the thread start-up
unsigned host_threads(bool big_input) {
    return g_host_threads > 0 ? (unsigned)g_host_threads : (big_input ? std::max(1u, std::thread::hardware_concurrency()) : 1u);
}

// ---- input: mapped as it is, or inflated into memory ---------------------------------------------------------------------
// A gzip file written by bgzip (BGZF: independent members of <= 64 KiB, each carrying its compressed size in a "BC" extra
// field and its inflated size in the trailer) is inflated block by block on the worker threads: the headers are walked
// first (no inflation), which gives every block its place in the output.  Any other gzip stream is one serial inflate.

struct BgzfBlock {
    size_t in_off, in_len;  // raw deflate data
    size_t out_off;
    uint32_t out_len, crc;
};

// block list of a BGZF file, or false if the data is not (entirely) BGZF
bool bgzf_index(const unsigned char *z, size_t n, std::vector<BgzfBlock> &blocks, size_t &total) {
    size_t o = 0;
    total = 0;
    while (o < n) {
        if (n - o < 18 + 8 || z[o] != 0x1f || z[o + 1] != 0x8b || z[o + 2] != 8 || !(z[o + 3] & 4)) return false;
        const size_t xlen = (size_t)z[o + 10] | ((size_t)z[o + 11] << 8);
        if (z[o + 3] & ~4u) return false;  // (bgzip sets FEXTRA only: no name / comment / header crc to skip)
        if (o + 12 + xlen > n) return false;
        size_t bsize = 0;
        for (size_t x = o + 12; x + 4 <= o + 12 + xlen;) {
            const size_t slen = (size_t)z[x + 2] | ((size_t)z[x + 3] << 8);
            if (z[x] == 'B' && z[x + 1] == 'C' && slen == 2 && x + 6 <= o + 12 + xlen) bsize = ((size_t)z[x + 4] | ((size_t)z[x + 5] << 8)) + 1;
            x += 4 + slen;
        }
        if (bsize < 12 + xlen + 8 || o + bsize > n) return false;
        const unsigned char *t = z + o + bsize - 8;
        BgzfBlock b;
        b.in_off = o + 12 + xlen;
        b.in_len = bsize - (12 + xlen) - 8;
        b.crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        b.out_len = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
        b.out_off = total;
        total += b.out_len;
        if (b.out_len) blocks.push_back(b);  // (the end-of-file marker is an empty block)
        o += bsize;
    }
    return !blocks.empty() || n > 0;
}

void inflate_bgzf(const std::string &path, const unsigned char *z, const std::vector<BgzfBlock> &blocks, size_t total, FileData &f) {
    f.heap.reset(new char[total ? total : 1]);
    char *out = f.heap.get();
    // a few blocks per hand-out: a block is ~64 KiB of output, ~50 us of work
    const size_t kBatch = 16, n_batches = (blocks.size() + kBatch - 1) / kBatch;
    parallel_for(n_batches, host_threads(total >= (8u << 20)), [&](size_t bi) {
        z_stream zs;
        std::memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) throw Error("zlib: inflateInit2 failed");
        for (size_t k = bi * kBatch; k < std::min(blocks.size(), (bi + 1) * kBatch); ++k) {
            const BgzfBlock &b = blocks[k];
            zs.next_in = const_cast<unsigned char *>(z + b.in_off);
            zs.avail_in = (uInt)b.in_len;
            zs.next_out = reinterpret_cast<unsigned char *>(out + b.out_off);
            zs.avail_out = b.out_len;
            const int rc = inflate(&zs, Z_FINISH);
            const bool ok = rc == Z_STREAM_END && zs.avail_out == 0 &&
                            (uint32_t)crc32(crc32(0L, Z_NULL, 0), reinterpret_cast<const unsigned char *>(out + b.out_off), b.out_len) == b.crc;
            if (!ok) {
                inflateEnd(&zs);
                throw Error("corrupt BGZF block in " + path);
            }
            inflateReset(&zs);
        }
        inflateEnd(&zs);
    });
    f.ptr = out;
    f.len = total;
}

void load_file(const std::string &path, FileData &f) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) throw Error("cannot open " + path);
    unsigned char magic[2] = {0, 0};
    const ssize_t got = pread(fd, magic, 2, 0);
    struct stat st;
    std::memset(&st, 0, sizeof(st));
    const bool gz = got == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) {
        if (st.st_size == 0) {
            close(fd);
            return;
        }
        void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m != MAP_FAILED) {
            madvise(m, (size_t)st.st_size, MADV_WILLNEED);
            if (!gz) {
                close(fd);
                f.map = m;
                f.ptr = static_cast<const char *>(m);
                f.len = (size_t)st.st_size;
                return;
            }
            std::vector<BgzfBlock> blocks;
            size_t total = 0;
            const bool bgzf = bgzf_index(static_cast<const unsigned char *>(m), (size_t)st.st_size, blocks, total);
            try {
                if (bgzf) inflate_bgzf(path, static_cast<const unsigned char *>(m), blocks, total, f);
            } catch (...) {
                munmap(m, (size_t)st.st_size);
                close(fd);
                throw;
            }
            munmap(m, (size_t)st.st_size);
            if (bgzf) {
                close(fd);
                return;
            }
        }
    }
    close(fd);
    gzFile z = gzopen(path.c_str(), "rb");  // (also reads plain data: pipes, files that could not be mapped)
    if (!z) throw Error("cannot open " + path);
    gzbuffer(z, 1u << 20);
    if (gz && S_ISREG(st.st_mode)) f.owned.reserve((size_t)st.st_size * 4u);  // GFA text deflates ~3-4x: one or no regrowth
    std::vector<char> buf(4u << 20);
    int n;
    while ((n = gzread(z, buf.data(), (unsigned)buf.size())) > 0) f.owned.append(buf.data(), (size_t)n);
    const bool bad = n < 0;
    gzclose(z);
    if (bad) throw Error("read error in " + path);
    f.ptr = f.owned.data();
    f.len = f.owned.size();
}

// Segment name -> item id (1-based, S-line order).  Real pangenome GFAs (pggb, minigraph-cactus) name their segments
// with decimal integers: then the lookup is a direct table indexed by the value, filled and read without hashing
// (one hash probe per path step is what dominated the parse: ~260 ns per step in a 1M-entry map).  Any other naming
// scheme falls back to a hash map keyed by views into the file buffer.
struct NodeIndex {
    bool numeric = false;
    std::vector<uint32_t> num2id;                             // numeric: value -> id, 0 = absent
    std::unordered_map<std::string_view, uint32_t> name2id;  // otherwise

    // canonical decimal (no sign, no leading zero, <= 18 digits) -> value
    static bool parse_canonical(const char *b, const char *e, uint64_t &v) {
        const size_t n = (size_t)(e - b);
        if (n == 0 || n > 18 || (n > 1 && *b == '0')) return false;
        uint64_t x = 0;
        for (const char *p = b; p < e; ++p) {
            const unsigned d = (unsigned)(*p - '0');
            if (d > 9u) return false;
            x = x * 10u + d;
        }
        v = x;
        return true;
    }

    void build(const std::vector<std::string_view> &names, unsigned nthreads) {
        const size_t n = names.size();
        numeric = n != 0;
        // every name parsed once, in slices on the worker threads: value (or "not a canonical decimal") and the maximum
        std::vector<uint64_t> vals(n);
        const size_t n_slices = std::max<size_t>(1, std::min<size_t>((size_t)nthreads * 4u, n / 65536u + 1u));
        const size_t per = (n + n_slices - 1) / std::max<size_t>(n_slices, 1);
        std::vector<uint64_t> slice_max(n_slices, 0);
        std::vector<uint8_t> slice_ok(n_slices, 1);
        parallel_for(n_slices, nthreads, [&](size_t c) {
            uint64_t mx = 0;
            for (size_t i = c * per; i < std::min(n, (c + 1) * per); ++i) {
                uint64_t v;
                if (!parse_canonical(names[i].data(), names[i].data() + names[i].size(), v)) {
                    slice_ok[c] = 0;
                    return;
                }
                vals[i] = v;
                mx = std::max(mx, v);
            }
            slice_max[c] = mx;
        });
        uint64_t mx = 0;
        for (size_t c = 0; c < n_slices; ++c) {
            numeric = numeric && slice_ok[c];
            mx = std::max(mx, slice_max[c]);
        }
        if (numeric && mx > 8ull * n + (1ull << 20)) numeric = false;  // sparse numbering: the table would be mostly holes
        if (numeric) {
            num2id.assign(mx + 1, 0u);
            // distinct values hit distinct slots, so the slices can fill the table concurrently; a duplicate name is two
            // writers of one slot, one of which loses: the verification pass finds the loser
            parallel_for(n_slices, nthreads, [&](size_t c) {
                for (size_t i = c * per; i < std::min(n, (c + 1) * per); ++i) num2id[vals[i]] = (uint32_t)i + 1u;
            });
            parallel_for(n_slices, nthreads, [&](size_t c) {
                for (size_t i = c * per; i < std::min(n, (c + 1) * per); ++i)
                    if (num2id[vals[i]] != (uint32_t)i + 1u)
                        throw Error("Segment with ID " + std::string(names[i]) + " occurs multiple times in GFA");
            });
        } else {
            name2id.reserve(n * 2);
            for (size_t i = 0; i < n; ++i)
                if (!name2id.emplace(names[i], (uint32_t)i + 1u).second)
                    throw Error("Segment with ID " + std::string(names[i]) + " occurs multiple times in GFA");
        }
    }

    uint32_t find(const char *b, const char *e) const {  // 0 = unknown
        if (numeric) {
            uint64_t v;
            return (parse_canonical(b, e, v) && v < num2id.size()) ? num2id[v] : 0u;
        }
        auto it = name2id.find(std::string_view(b, (size_t)(e - b)));
        return it == name2id.end() ? 0u : it->second;
    }

    uint32_t get(const char *b, const char *e) const {
        const uint32_t id = find(b, e);
        if (!id) throw Error("unknown node " + std::string(b, (size_t)(e - b)));
        return id;
    }
};

// steps of a P line's segment list "12+,13-,..." (util.rs:1093-1142)
void parse_path_steps(const NodeIndex &idx, const char *b, const char *e, std::vector<Step> &steps) {
    steps.reserve((size_t)std::count(b, e, ',') + 1);
    const char *s = b;
    while (s < e) {
        if (idx.numeric) {  // digits, then the orientation; anything else is reported by the generic branch below
            uint64_t v = 0;
            const char *q = s;
            while (q < e && (unsigned)(*q - '0') <= 9u && q - s < 18) v = v * 10u + (unsigned)(*q++ - '0');
            if (q > s && q < e && (*q == '+' || *q == '-') && !(q - s > 1 && *s == '0') && (q + 1 == e || q[1] == ',')) {
                const uint32_t id = v < idx.num2id.size() ? idx.num2id[v] : 0u;
                if (!id) throw Error("unknown node " + std::string(s, (size_t)(q - s)));
                steps.push_back({id, *q == '+'});
                s = q + 2;
                continue;
            }
        }
        const void *t = memchr(s, ',', (size_t)(e - s));
        const char *te = t ? (const char *)t : e;
        if (te > s) steps.push_back({idx.get(s, te - 1), te[-1] == '+'});
        s = te + 1;
    }
}

// steps of a W line's walk ">12<13..." (util.rs:916-931)
void parse_walk_steps(const NodeIndex &idx, const char *b, const char *e, std::vector<Step> &steps) {
    const char *s = b;
    while (s < e) {
        const bool fwd = *s == '>';
        const char *q = s + 1;
        while (q < e && *q != '>' && *q != '<') ++q;
        if (q > s + 1) steps.push_back({idx.get(s + 1, q), fwd});
        s = q;
    }
}

// ---- lean variants: node ids only (same tokenisation as the two functions above) -------------------------------------
uint64_t count_path_steps(const char *b, const char *e) {  // non-empty comma separated tokens
    uint64_t n = 0;
    bool in_token = false;
    for (const char *s = b; s < e; ++s) {
        const bool sep = *s == ',';
        n += (!sep && !in_token) ? 1u : 0u;
        in_token = !sep;
    }
    return n;
}

uint64_t count_walk_steps(const char *b, const char *e) {  // '>' / '<' followed by a non-empty name
    uint64_t n = 0;
    for (const char *s = b; s < e;) {
        const char *q = s + 1;
        while (q < e && *q != '>' && *q != '<') ++q;
        if (q > s + 1) ++n;
        s = q;
    }
    return n;
}

uint64_t parse_path_nodes(const NodeIndex &idx, const char *b, const char *e, uint32_t *out) {
    uint64_t n = 0;
    const char *s = b;
    while (s < e) {
        if (idx.numeric) {
            uint64_t v = 0;
            const char *q = s;
            while (q < e && (unsigned)(*q - '0') <= 9u && q - s < 18) v = v * 10u + (unsigned)(*q++ - '0');
            if (q > s && q < e && (*q == '+' || *q == '-') && !(q - s > 1 && *s == '0') && (q + 1 == e || q[1] == ',')) {
                const uint32_t id = v < idx.num2id.size() ? idx.num2id[v] : 0u;
                if (!id) throw Error("unknown node " + std::string(s, (size_t)(q - s)));
                out[n++] = id;
                s = q + 2;
                continue;
            }
        }
        const void *t = memchr(s, ',', (size_t)(e - s));
        const char *te = t ? (const char *)t : e;
        if (te > s) out[n++] = idx.get(s, te - 1);
        s = te + 1;
    }
    return n;
}

uint64_t parse_walk_nodes(const NodeIndex &idx, const char *b, const char *e, uint32_t *out) {
    uint64_t n = 0;
    const char *s = b;
    while (s < e) {
        const char *q = s + 1;
        while (q < e && *q != '>' && *q != '<') ++q;
        if (q > s + 1) out[n++] = idx.get(s + 1, q);
        s = q;
    }
    return n;
}

}  // namespace

void set_host_threads(int n) { g_host_threads = n > 0 ? n : 0; }
unsigned host_thread_budget() { return host_threads(true); }

GraphStorage GraphStorage::from_gfa(const std::string &path, bool with_edges, bool with_names, bool lean) {
    GraphStorage g;
    g.node_lens.push_back(0);
    g.has_edges = with_edges;
    const bool tm = getenv("PGX_PARSE_TIMING") != nullptr;  // stage times on stderr (measurement aid)
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!tm) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "from_gfa %s: %.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    FileData file;
    load_file(path, file);
    const char *base = file.ptr;
    const size_t size = file.len;
    const unsigned nt = host_threads(size >= (8u << 20));
    lap("open / map file");
    // line index: every thread collects the newlines of its slice of the file; the [begin, end) pairs (without the
    // newline, empty lines dropped) are then strung together in file order
    std::vector<std::pair<size_t, size_t>> lines;
    {
        const size_t n_slices = std::max<size_t>(1, std::min<size_t>((size_t)nt * 4u, size / (1u << 20) + 1u));
        const size_t slice = (size + n_slices - 1) / n_slices;
        std::vector<std::vector<size_t>> nls(n_slices);
        parallel_for(n_slices, nt, [&](size_t c) {
            const size_t lo = c * slice, hi = std::min(size, lo + slice);
            for (size_t i = lo; i < hi;) {
                const void *nl = memchr(base + i, '\n', hi - i);
                if (!nl) break;
                const size_t j = (size_t)((const char *)nl - base);
                nls[c].push_back(j);
                i = j + 1;
            }
        });
        size_t total = 0;
        for (auto &v : nls) total += v.size();
        lines.reserve(total + 1);
        size_t prev = 0;
        for (auto &v : nls)
            for (size_t j : v) {
                if (j > prev) lines.emplace_back(prev, j);
                prev = j + 1;
            }
        if (prev < size) lines.emplace_back(prev, size);
    }
    lap("line index");
    auto field = [&](size_t b, size_t e, int k, size_t &fb, size_t &fe) -> bool {  // k-th TAB separated field
        size_t s = b;
        for (int c = 0; c < k; ++c) {
            const void *t = memchr(base + s, '\t', e - s);
            if (!t) return false;
            s = (size_t)((const char *)t - base) + 1;
        }
        const void *t = memchr(base + s, '\t', e - s);
        fb = s;
        fe = t ? (size_t)((const char *)t - base) : e;
        while (fe > fb && base[fe - 1] == '\r') --fe;
        return true;
    };
    // pass 1: segments and path names (graph.rs:308-375); ids follow the S-line order.  Blocks of lines are scanned on
    // the worker threads and their finds concatenated in file order.
    std::vector<std::string_view> names;
    std::vector<size_t> path_lines;  // indices into `lines` of the P / W lines, file order
    {
        struct Found {
            std::vector<std::string_view> names;
            std::vector<uint32_t> lens;
            std::vector<size_t> path_lines;
            std::vector<PathSegment> segs;
        };
        const size_t n_blocks = std::max<size_t>(1, std::min<size_t>((size_t)nt * 8u, lines.size() / 4096u + 1u));
        const size_t per = (lines.size() + n_blocks - 1) / n_blocks;
        std::vector<Found> found(n_blocks);
        parallel_for(n_blocks, nt, [&](size_t blk) {
            Found &f = found[blk];
            const size_t lo = blk * per, hi = std::min(lines.size(), lo + per);
            for (size_t li = lo; li < hi; ++li) {
                const auto &ln = lines[li];
                const char tag = base[ln.first];
                size_t fb, fe;
                if (tag == 'S') {
                    if (!field(ln.first, ln.second, 1, fb, fe)) throw Error("malformed S line");
                    f.names.emplace_back(base + fb, fe - fb);
                    size_t sb, se;
                    uint32_t len = 0;
                    if (field(ln.first, ln.second, 2, sb, se)) len = (uint32_t)(se - sb);
                    f.lens.push_back(len);
                } else if (tag == 'P') {
                    if (!field(ln.first, ln.second, 1, fb, fe)) throw Error("malformed P line");
                    f.segs.push_back(PathSegment::from_str(std::string(base + fb, fe - fb)));
                    f.path_lines.push_back(li);
                } else if (tag == 'W') {
                    std::string w[6];
                    for (int k = 1; k <= 5; ++k) {
                        if (!field(ln.first, ln.second, k, fb, fe)) throw Error("malformed W line");
                        w[k] = std::string(base + fb, fe - fb);
                    }
                    PathSegment ps;  // util.rs:368-398
                    ps.sample = w[1];
                    ps.haplotype = w[2];
                    ps.seqid = w[3];
                    if (w[4] != "*") ps.start = std::stoull(w[4]);
                    if (w[5] != "*") ps.end = std::stoull(w[5]);
                    f.segs.push_back(ps);
                    f.path_lines.push_back(li);
                }
            }
        });
        size_t n_seg = 0, n_path = 0;
        for (auto &f : found) n_seg += f.names.size(), n_path += f.path_lines.size();
        names.reserve(n_seg);
        g.node_lens.reserve(n_seg + 1);
        path_lines.reserve(n_path);
        g.path_segments.reserve(n_path);
        for (auto &f : found) {
            names.insert(names.end(), f.names.begin(), f.names.end());
            g.node_lens.insert(g.node_lens.end(), f.lens.begin(), f.lens.end());
            path_lines.insert(path_lines.end(), f.path_lines.begin(), f.path_lines.end());
            for (auto &x : f.segs) g.path_segments.push_back(std::move(x));
        }
    }
    lap("pass 1 (S / P / W lines)");
    if (names.size() >= (1u << 31)) throw Error("more than 2^31 segments are not supported");
    NodeIndex idx;
    idx.build(names, nt);
    lap("segment index");
    if (with_names) {
        g.node_names.reserve(names.size() + 1);
        g.node_names.emplace_back();
        for (auto &nm : names) g.node_names.emplace_back(nm);
    }
    // pass 1b: links (graph.rs:276-306)
    if (with_edges) {
        for (auto &ln : lines) {
            if (base[ln.first] != 'L') continue;
            size_t b1, e1, b2, e2, b3, e3, b4, e4;
            if (!field(ln.first, ln.second, 1, b1, e1) || !field(ln.first, ln.second, 2, b2, e2) ||
                !field(ln.first, ln.second, 3, b3, e3) || !field(ln.first, ln.second, 4, b4, e4))
                throw Error("malformed L line");
            const uint64_t e = canonical_edge(idx.get(base + b1, base + e1), base[b2] == '+', idx.get(base + b3, base + e3), base[b4] == '+');
            g.edge2id.emplace(e, (uint32_t)g.edge2id.size() + 1);  // no-op if the edge is already known
        }
        lap("links");
    }
    // pass 2: steps.  Every P / W line is independent and the index is read-only: the lines are handed out to the
    // worker threads through an atomic counter (the lines of a pangenome differ a lot in length).
    auto steps_field = [&](size_t k, size_t &fb, size_t &fe) -> bool {  // the step list of path line k; true: P line
        const auto &ln = lines[path_lines[k]];
        const bool is_p = base[ln.first] == 'P';
        if (!field(ln.first, ln.second, is_p ? 2 : 6, fb, fe)) throw Error(is_p ? "malformed P line" : "malformed W line");
        return is_p;
    };
    if (lean) {
        // Counting nodes / bp without subset or exclude lists only needs WHICH nodes a path visits: the ids go straight
        // into one flat u32 array (the wire format of pgx_abacus_build_u32) -- no per-path vectors of (node, orientation),
        // no u64 ItemTable copy, no narrowing pass: 4 bytes written per step instead of 8 + 8 + 4.
        g.lean = true;
        const size_t P = path_lines.size();
        g.flat_prefsum.assign(P + 1, 0);
        parallel_for(P, nt, [&](size_t k) {
            size_t fb, fe;
            const bool is_p = steps_field(k, fb, fe);
            g.flat_prefsum[k + 1] = is_p ? count_path_steps(base + fb, base + fe) : count_walk_steps(base + fb, base + fe);
        });
        for (size_t k = 0; k < P; ++k) g.flat_prefsum[k + 1] += g.flat_prefsum[k];
        g.flat_nodes.reset(new uint32_t[std::max<uint64_t>(1, g.flat_prefsum[P])]);  // (uninitialised: every word is written below)
        parallel_for(P, nt, [&](size_t k) {
            size_t fb, fe;
            const bool is_p = steps_field(k, fb, fe);
            uint32_t *out = g.flat_nodes.get() + g.flat_prefsum[k];
            const uint64_t n = g.flat_prefsum[k + 1] - g.flat_prefsum[k];
            const uint64_t got = is_p ? parse_path_nodes(idx, base + fb, base + fe, out) : parse_walk_nodes(idx, base + fb, base + fe, out);
            if (got != n) throw Error("internal error: step count mismatch while parsing a path");
        });
        lap("pass 2 (path steps -> flat u32 ids)");
        return g;
    }
    g.path_steps.resize(path_lines.size());
    auto parse_line = [&](size_t k) {
        size_t fb, fe;
        if (steps_field(k, fb, fe))
            parse_path_steps(idx, base + fb, base + fe, g.path_steps[k]);
        else
            parse_walk_steps(idx, base + fb, base + fe, g.path_steps[k]);
    };
    parallel_for(path_lines.size(), nt, parse_line);
    lap("pass 2 (path steps)");
    return g;
}

// ---- BED / group files (io.rs:35-147) ------------------------------------------------------------------------------

std::vector<PathSegment> parse_bed_to_path_segments(const std::string &path, bool use_block_info) {
    std::ifstream in(path);
    if (!in) throw Error("cannot open " + path);
    std::vector<PathSegment> segs;
    std::string line;
    size_t i = 0;
    while (std::getline(in, line)) {
        ++i;
        while (!line.empty() && (line.back() == '\r')) line.pop_back();
        if (line.empty()) continue;
        const std::vector<std::string> f = split(line, '\t');
        const std::string &name = f[0];
        if (name.rfind("browser ", 0) == 0 || name.rfind("track ", 0) == 0 || name[0] == '#') continue;
        if (f.size() == 1) {
            segs.push_back(PathSegment::from_str(name));
        } else if (f.size() >= 3) {
            const uint64_t start = std::stoull(f[1]), end = std::stoull(f[2]);
            if (use_block_info && f.size() == 12) {
                std::vector<uint64_t> sizes, starts;
                for (auto &s : split(f[10], ','))
                    if (all_digits(s)) sizes.push_back(std::stoull(s));
                for (auto &s : split(f[11], ','))
                    if (all_digits(s)) starts.push_back(std::stoull(s));
                for (size_t k = 0; k < std::min(sizes.size(), starts.size()); ++k)
                    segs.push_back(PathSegment::from_str_start_end(name, start + starts[k], start + starts[k] + sizes[k]));
            } else {
                segs.push_back(PathSegment::from_str_start_end(name, start, end));
            }
        } else {
            throw Error("error in line " + std::to_string(i) + ": row must have either 1, 3, or 12 columns, but has 2");
        }
    }
    return segs;
}

// ---- GraphMask ------------------------------------------------------------------------------------------------------

namespace {

std::map<PathSegment, std::string> load_groups(const GraphStorage &g, const GraphMaskParameters &p) {  // abacus.rs:242-308
    std::map<PathSegment, std::string> res;
    if (p.groupby_haplotype) {
        for (auto &x : g.path_segments) res[x.clear_coords()] = x.sample + "#" + (x.haplotype ? *x.haplotype : "");
        return res;
    }
    if (p.groupby_sample) {
        for (auto &x : g.path_segments) res[x.clear_coords()] = x.sample;
        return res;
    }
    if (!p.groupby.empty()) {
        std::ifstream in(p.groupby);
        if (!in) throw Error("cannot open " + p.groupby);
        std::string line;
        size_t i = 0;
        while (std::getline(in, line)) {
            ++i;
            while (!line.empty() && line.back() == '\r') line.pop_back();
            if (line.empty()) continue;
            const auto cols = split(line, '\t');
            if (cols.size() != 2) throw Error("error in line " + std::to_string(i) + ": table must have exactly two columns");
            const PathSegment seg = PathSegment::from_str(cols[0]).clear_coords();
            auto it = res.find(seg);
            if (it != res.end() && it->second != cols[1])
                throw Error("error in line " + std::to_string(i) + ": path " + seg.to_string() + " cannot be assigned to more than one group");
            res.emplace(seg, cols[1]);
        }
        for (auto &x : g.path_segments) res.emplace(x.clear_coords(), x.id());
        return res;
    }
    for (auto &x : g.path_segments) res[x.clear_coords()] = x.id();
    return res;
}

std::optional<std::vector<PathSegment>> load_coord_list(const std::string &text, const std::vector<PathSegment> &paths) {
    // abacus.rs:212-240: a file (BED / 1-column list) or, failing that, a regex over the path names
    if (text.empty()) return std::nullopt;
    std::ifstream probe(text);
    if (probe.good()) return parse_bed_to_path_segments(text, true);
    std::regex rx(text);
    std::vector<PathSegment> out;
    for (auto &p : paths)
        if (std::regex_search(p.to_string(), rx)) out.push_back(p);
    return out;
}

std::optional<std::vector<PathSegment>> complement_with_group_assignments(
    const std::optional<std::vector<PathSegment>> &coords, const std::map<PathSegment, std::string> &groups) {  // abacus.rs:152-201
    if (!coords) return std::nullopt;
    std::map<std::string, std::vector<PathSegment>> group2paths;
    for (auto &kv : groups) group2paths[kv.second].push_back(kv.first);
    std::vector<PathSegment> out;
    for (auto &p : *coords) {
        if (groups.count(p.clear_coords())) {
            out.push_back(p);
        } else {
            auto it = group2paths.find(p.id());
            if (it != group2paths.end()) {
                if (p.coords())
                    throw Error("invalid coordinate \"" + p.to_string() + "\": group identifiers are not allowed to have start/stop information!");
                out.insert(out.end(), it->second.begin(), it->second.end());
            }
            // unknown path / group: dropped (the reference only logs)
        }
    }
    return out;
}

}  // namespace

GraphMask GraphMask::from_graph(const GraphStorage &g, const GraphMaskParameters &p) {
    GraphMask m;
    m.groups = load_groups(g, p);
    m.include_coords = complement_with_group_assignments(load_coord_list(p.positive_list, g.path_segments), m.groups);
    m.exclude_coords = complement_with_group_assignments(load_coord_list(p.negative_list, g.path_segments), m.groups);
    if (p.order && !p.order->empty()) {
        m.order = complement_with_group_assignments(parse_bed_to_path_segments(*p.order, true), m.groups);
        if (m.order && !m.order->empty()) {  // groups must not be fragmented (abacus.rs:114-127)
            std::set<std::string> visited;
            std::string cur = m.groups.at((*m.order)[0].clear_coords());
            for (auto &seg : *m.order) {
                const std::string &grp = m.groups.at(seg.clear_coords());
                if (cur != grp && !visited.insert(grp).second)
                    throw Error("order of paths contains fragmented groups: path " + seg.to_string() +
                                " belongs to group that is interspersed by one or more other groups");
                cur = grp;
            }
        }
    }
    return m;
}

std::vector<std::pair<uint64_t, std::string>> GraphMask::get_path_order(const std::vector<PathSegment> &paths) const {
    // abacus.rs:310-347: all paths of a group are emitted together at the group's first position
    std::map<std::string, std::vector<std::pair<uint64_t, std::string>>> group_to_paths;
    for (size_t i = 0; i < paths.size(); ++i) {
        const std::string &grp = groups.at(paths[i].clear_coords());
        group_to_paths[grp].emplace_back((uint64_t)i, grp);
    }
    std::vector<PathSegment> ord;
    if (order) {
        ord = *order;
    } else if (include_coords) {
        ord = *include_coords;
    } else {
        std::set<PathSegment> ex;
        if (exclude_coords) ex.insert(exclude_coords->begin(), exclude_coords->end());
        for (auto &p : paths)
            if (!ex.count(p)) ord.push_back(p);
    }
    std::vector<std::pair<uint64_t, std::string>> out;
    for (auto &p : ord) {
        auto git = groups.find(p.clear_coords());
        if (git == groups.end()) continue;
        auto it = group_to_paths.find(git->second);
        if (it == group_to_paths.end()) continue;
        out.insert(out.end(), it->second.begin(), it->second.end());
        group_to_paths.erase(it);
    }
    return out;
}

// ---- ItemTable construction (graph_broker/util.rs:208-366, 569-790, 1186-1248; abacus.rs:1187-1229) ------------------

namespace {

void update_tables(const GraphStorage &g, const std::vector<Step> &steps, const Intervals &include, const Intervals &exclude,
                   uint64_t offset, std::vector<uint64_t> &items, IntervalContainer *subset_covered, ActiveTable *exclude_table) {
    // graph_broker/util.rs:569-721
    size_t i = 0, j = 0;
    uint64_t p = offset;
    for (auto &st : steps) {
        const uint64_t sid = st.node, l = g.node_lens[sid];
        bool stop_here = false;
        while (i < include.size() && include[i].first < p + l && !stop_here) {
            if (include[i].second > p) {
                uint64_t a = include[i].first > p ? include[i].first - p : 0, b;
                if (include[i].second < p + l) {
                    ++i;
                    b = include[i - 1].second - p;
                } else {
                    stop_here = true;
                    b = l;
                }
                if (!st.forward) {
                    const uint64_t na = l - b, nb = l - a;
                    a = na;
                    b = nb;
                }
                items.push_back(sid);
                if (subset_covered) {
                    if (b - a == l) {
                        if (subset_covered->contains(sid)) subset_covered->remove(sid);
                    } else {
                        subset_covered->add(sid, a, b);
                    }
                }
            } else {
                ++i;
            }
        }
        stop_here = false;
        while (j < exclude.size() && exclude[j].first < p + l && !stop_here) {
            if (exclude[j].second > p) {
                uint64_t a = exclude[j].first > p ? exclude[j].first - p : 0, b;
                if (exclude[j].second < p + l) {
                    ++j;
                    b = exclude[j - 1].second - p;
                } else {
                    stop_here = true;
                    b = l;
                }
                if (!st.forward) {
                    const uint64_t na = l - b, nb = l - a;
                    a = na;
                    b = nb;
                }
                if (exclude_table) {
                    if (exclude_table->with_annotation)
                        exclude_table->activate_n_annotate(sid, l, a, b);
                    else
                        exclude_table->activate(sid);
                }
            } else {
                ++j;
            }
        }
        if (i >= include.size() && j >= exclude.size()) break;
        p += l;
    }
}

void update_tables_edgecount(const GraphStorage &g, const std::vector<Step> &steps, const Intervals &include,
                             const Intervals &exclude, uint64_t offset, std::vector<uint64_t> &items, ActiveTable *exclude_table) {
    // graph_broker/util.rs:723-790
    size_t i = 0, j = 0;
    uint64_t p = offset;
    if (!steps.empty()) p += g.node_lens[steps[0].node];
    for (size_t k = 0; k + 1 < steps.size(); ++k) {
        const Step &s1 = steps[k], &s2 = steps[k + 1];
        while (i < include.size() && include[i].second <= p) ++i;
        while (j < exclude.size() && exclude[j].second <= p) ++j;
        const uint64_t l = g.node_lens[s2.node];
        const uint64_t eid = g.edge2id.find(canonical_edge(s1.node, s1.forward, s2.node, s2.forward));
        if (!eid) throw Error("path uses an edge that has no L line");
        if (i < include.size() && include[i].first < p + l) items.push_back(eid);
        if (exclude_table && j < exclude.size() && exclude[j].first < p + l)
            exclude_table->activate(eid);
        else if (i >= include.size() && j >= exclude.size())
            break;
        p += l;
    }
}

}  // namespace

uint64_t GraphStorage::step_count() const {
    if (lean) return flat_prefsum.empty() ? 0 : flat_prefsum.back();
    uint64_t n = 0;
    for (auto &v : path_steps) n += v.size();
    return n;
}

bool GraphStorage::lean_apply_subset(const GraphMask &mask) {
    if (!lean || !mask.include_coords) return true;
    const std::map<std::string, Intervals> include_map = build_subpath_map(*mask.include_coords);
    const Intervals none = {};
    const size_t P = path_segments.size();
    std::vector<uint8_t> keep(P, 0);
    for (size_t pi = 0; pi < P; ++pi) {  // the same three cases as build_item_tables below
        const PathSegment &seg = path_segments[pi];
        auto it = include_map.find(seg.id());
        const Intervals &inc = it == include_map.end() ? none : it->second;
        const auto c = seg.coords();
        const std::pair<uint64_t, uint64_t> whole = {c ? c->first : 0, c ? c->second : kUsizeMax};
        if (!intersects(inc, whole)) continue;
        if (!is_contained(inc, whole)) return false;
        keep[pi] = 1;
    }
    uint64_t w = 0;
    std::vector<uint64_t> prefsum(P + 1, 0);
    for (size_t pi = 0; pi < P; ++pi) {
        const uint64_t b = flat_prefsum[pi], n = flat_prefsum[pi + 1] - b;
        if (keep[pi]) {
            if (w != b) std::memmove(flat_nodes.get() + w, flat_nodes.get() + b, n * sizeof(uint32_t));
            w += n;
        }
        prefsum[pi + 1] = w;
    }
    flat_prefsum.swap(prefsum);
    lean_subset_applied = true;
    return true;
}

ItemTables build_item_tables(const GraphStorage &g, const GraphMask &mask, CountType count) {
    ItemTables t;
    const bool edge = count == CountType::Edge;
    t.n_items = edge ? g.edge_count() : g.node_count();
    if (g.lean) {  // the parser already wrote the table (from_gfa): every step of every (subset) path counts its node
        if (edge || mask.exclude_coords || (mask.include_coords && !g.lean_subset_applied))
            throw Error("internal error: lean parse used with edge counting, an exclude list or an unapplied subset list");
        t.items32 = g.flat_nodes.get();
        t.id_prefsum = g.flat_prefsum;
        t.n_steps = g.flat_prefsum.empty() ? 0 : g.flat_prefsum.back();
        return t;
    }
    std::optional<IntervalContainer> subset_covered;
    if (count == CountType::Bp && mask.include_coords) subset_covered.emplace();
    std::optional<ActiveTable> exclude_table;
    if (mask.exclude_coords) exclude_table.emplace(t.n_items + 1, count == CountType::Bp);
    std::map<std::string, Intervals> include_map, exclude_map;
    if (mask.include_coords) include_map = build_subpath_map(*mask.include_coords);
    if (mask.exclude_coords) exclude_map = build_subpath_map(*mask.exclude_coords);
    const Intervals complete = {{0, kUsizeMax}}, none = {};
    size_t total_steps = 0;
    for (auto &v : g.path_steps) total_steps += v.size();
    if (!mask.include_coords && !mask.exclude_coords) {
        // no subset / exclude list: every path contributes all of its items and nothing is shared between paths, so
        // the paths are translated in parallel (edge counting does one hash lookup per step) and concatenated
        const size_t P = g.path_segments.size();
        const unsigned nt = host_threads(total_steps >= (1u << 20));
        std::vector<std::vector<uint64_t>> per(edge ? P : 0);
        t.id_prefsum.assign(P + 1, 0);
        if (edge) {
            parallel_for(P, nt, [&](size_t pi) {
                const auto c = g.path_segments[pi].coords();
                per[pi].reserve(g.path_steps[pi].size());
                update_tables_edgecount(g, g.path_steps[pi], complete, none, c ? c->first : 0, per[pi], nullptr);
            });
            for (size_t pi = 0; pi < P; ++pi) t.id_prefsum[pi + 1] = t.id_prefsum[pi] + per[pi].size();
        } else {
            for (size_t pi = 0; pi < P; ++pi) t.id_prefsum[pi + 1] = t.id_prefsum[pi] + g.path_steps[pi].size();
        }
        t.items.resize(t.id_prefsum[P]);
        parallel_for(P, nt, [&](size_t pi) {
            uint64_t *dst = t.items.data() + t.id_prefsum[pi];
            if (edge) {
                std::copy(per[pi].begin(), per[pi].end(), dst);
                std::vector<uint64_t>().swap(per[pi]);
            } else {
                const std::vector<Step> &steps = g.path_steps[pi];
                for (size_t k = 0; k < steps.size(); ++k) dst[k] = steps[k].node;
            }
        });
        t.n_steps = t.items.size();
        return t;
    }
    t.items.reserve(total_steps);
    t.id_prefsum.reserve(g.path_segments.size() + 1);
    t.id_prefsum.push_back(0);
    for (size_t pi = 0; pi < g.path_segments.size(); ++pi) {
        const PathSegment &seg = g.path_segments[pi];
        const std::vector<Step> &steps = g.path_steps[pi];
        const Intervals *inc = &complete, *exc = &none;
        if (mask.include_coords) {
            auto it = include_map.find(seg.id());
            inc = it == include_map.end() ? &none : &it->second;
        }
        if (mask.exclude_coords) {
            auto it = exclude_map.find(seg.id());
            exc = it == exclude_map.end() ? &none : &it->second;
        }
        const auto c = seg.coords();
        const uint64_t start = c ? c->first : 0, end = c ? c->second : kUsizeMax;
        if (mask.include_coords && !intersects(*inc, {start, end}) && !intersects(*exc, {start, end})) {
            t.id_prefsum.push_back(t.items.size());  // path neither subset nor excluded: skipped (util.rs:276-291)
            continue;
        }
        if (!edge && (!mask.include_coords || is_contained(*inc, {start, end})) &&
            (!mask.exclude_coords || is_contained(*exc, {start, end}))) {
            // parse_path_seq_update_tables (util.rs:1186-1248): every step counts; if the path is on the
            // exclude list, all of its items are flagged
            const size_t first = t.items.size();
            for (auto &s : steps) t.items.push_back(s.node);
            if (!exc->empty() && exclude_table)
                for (size_t k = first; k < t.items.size(); ++k) exclude_table->items[t.items[k]] = 1;
        } else if (edge) {
            update_tables_edgecount(g, steps, *inc, *exc, start, t.items, exclude_table ? &*exclude_table : nullptr);
        } else {
            update_tables(g, steps, *inc, *exc, start, t.items, subset_covered ? &*subset_covered : nullptr,
                          exclude_table ? &*exclude_table : nullptr);
        }
        t.id_prefsum.push_back(t.items.size());
    }
    if (subset_covered) {  // quantify_uncovered_bps, abacus.rs:1187-1229
        for (auto &kv : subset_covered->map) {
            const uint64_t sid = kv.first;
            if (exclude_table && exclude_table->items[sid]) continue;
            const uint64_t l = g.node_lens[sid];
            uint64_t covered;
            if (exclude_table) {
                const Intervals ex = exclude_table->get_active_intervals(sid, l);
                covered = subset_covered->total_coverage(sid, &ex);
            } else {
                covered = subset_covered->total_coverage(sid, nullptr);
            }
            if (covered <= l) t.uncovered_bps[sid] = l - covered;
        }
    }
    if (exclude_table) t.exclude = exclude_table->items;
    t.n_steps = t.items.size();
    return t;
}

}  // namespace panacus
