// synth.cpp -- generator of pangenome-shaped GFA text for measurements (`panacus debug-synth-gfa`): a graph with the
// size and coverage statistics of a real one when the real file is not at hand (BASELINE.json configs[4]: chr22 of the
// HPRC v1.0 pggb graph is not available offline).  Not part of the reference CLI; nothing here is on the counting path.
//
// Shape: numeric segment ids 1..N in S-line order (pggb style); `samples` x `haps` haplotypes named sample#hap#contig
// (PanSN), each cut into `contigs` P lines over consecutive id ranges; every node draws its sample coverage c from a
// given coverage histogram (e.g. the node histogram the reference documents for chr22,
// docs/chr22.hprc-v1.0-pggb.histgrowth.html:268), then c distinct samples, then for each of them haplotype 1, 2 or
// both; its length follows the mean length of its coverage class (bp histogram / node histogram), half the nodes
// being 1 bp long like in pggb graphs.  L lines chain consecutive ids.  Deterministic for a seed (splitmix64).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>

#include "panacus_host.hpp"

namespace panacus {

namespace {
struct Rng {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return next() % n; }
};

struct Out {
    FILE *f;
    std::vector<char> buf;
    explicit Out(const std::string &path) : f(fopen(path.c_str(), "wb")) {
        if (!f) throw Error("cannot write " + path);
        buf.reserve(1u << 24);
    }
    ~Out() {
        flush();
        fclose(f);
    }
    void flush() {
        if (!buf.empty() && fwrite(buf.data(), 1, buf.size(), f) != buf.size()) throw Error("short write");
        buf.clear();
    }
    void put(const char *p, size_t n) {
        if (buf.size() + n > buf.capacity()) flush();
        buf.insert(buf.end(), p, p + n);
    }
    void put(const std::string &s) { put(s.data(), s.size()); }
    void num(uint64_t v) {
        char tmp[24];
        int n = 0;
        do tmp[n++] = (char)('0' + v % 10);
        while (v /= 10);
        char rev[24];
        for (int i = 0; i < n; ++i) rev[i] = tmp[n - 1 - i];
        put(rev, (size_t)n);
    }
};
}  // namespace

// node_hist / bp_hist: samples + 1 entries each (coverage 0 .. samples); bp_hist may be empty (then lengths are
// 1 + Exp(2 * mean_len - 2) for every class).  Returns the number of path steps written.
uint64_t synth_gfa(const std::string &path, uint64_t n_nodes, uint32_t samples, uint32_t haps, uint32_t contigs,
                   const std::vector<double> &node_hist, const std::vector<double> &bp_hist, double mean_len, uint64_t seed) {
    if (samples == 0 || haps == 0 || contigs == 0 || n_nodes == 0) throw Error("synth-gfa: empty shape");
    if (node_hist.size() != samples + 1u) throw Error("synth-gfa: the coverage histogram needs samples + 1 entries");
    Rng rng{seed};
    // cumulative distribution of the coverage classes and mean length per class
    std::vector<double> cdf(samples + 1u);
    const double tot = std::accumulate(node_hist.begin(), node_hist.end(), 0.0);
    double run = 0;
    for (uint32_t c = 0; c <= samples; ++c) cdf[c] = (run += node_hist[c] / tot);
    std::vector<double> mlen(samples + 1u, mean_len);
    if (bp_hist.size() == node_hist.size())
        for (uint32_t c = 0; c <= samples; ++c)
            if (node_hist[c] > 0) mlen[c] = std::max(1.0, bp_hist[c] / node_hist[c]);
    const uint32_t H = samples * haps;
    std::vector<std::vector<uint32_t>> steps(H);  // node ids per haplotype, ascending
    std::vector<uint32_t> pick(samples);
    Out out(path);
    out.put("H\tVN:Z:1.0\n");
    static const char kBase[4] = {'A', 'C', 'G', 'T'};
    std::string seq;
    for (uint64_t id = 1; id <= n_nodes; ++id) {
        const double u = rng.uniform();
        const uint32_t c = (uint32_t)(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin());
        const uint32_t cov = std::min(c, samples);
        // length: half the nodes 1 bp, the others 1 + Exp so that the class mean is mlen[cov]
        uint64_t len = 1;
        if (rng.uniform() >= 0.5) {
            const double m = std::max(0.0, 2.0 * mlen[cov] - 2.0);
            len = 1 + (uint64_t)std::min(99999.0, std::floor(-std::log(1.0 - rng.uniform()) * m));
        }
        out.put("S\t", 2);
        out.num(id);
        out.put("\t", 1);
        seq.resize(len);
        for (uint64_t k = 0; k < len; k += 32) {
            uint64_t r = rng.next();
            for (uint64_t j = k; j < std::min(len, k + 32); ++j, r >>= 2) seq[j] = kBase[r & 3u];
        }
        out.put(seq);
        out.put("\n", 1);
        // cov distinct samples (partial Fisher-Yates), each through haplotype 1, 2, ... or several
        std::iota(pick.begin(), pick.end(), 0u);
        for (uint32_t k = 0; k < cov; ++k) {
            const uint32_t j = k + (uint32_t)rng.below(samples - k);
            std::swap(pick[k], pick[j]);
            const uint32_t s = pick[k];
            const uint64_t r = rng.next();
            bool any = false;
            for (uint32_t h = 0; h < haps; ++h)
                if ((r >> h) & 1u) {
                    steps[s * haps + h].push_back((uint32_t)id);
                    any = true;
                }
            if (!any) steps[s * haps + (uint32_t)((r >> 32) % haps)].push_back((uint32_t)id);
        }
    }
    for (uint64_t id = 1; id < n_nodes; ++id) {  // a chain of links (the counting path never reads them for node / bp)
        out.put("L\t", 2);
        out.num(id);
        out.put("\t+\t", 3);
        out.num(id + 1);
        out.put("\t+\t0M\n", 6);
    }
    uint64_t total_steps = 0;
    for (uint32_t s = 0; s < samples; ++s)
        for (uint32_t h = 0; h < haps; ++h) {
            const auto &v = steps[s * haps + h];
            for (uint32_t k = 0; k < contigs; ++k) {  // contig k: the haplotype's steps inside the k-th id range
                const uint64_t lo = 1 + n_nodes * k / contigs, hi = 1 + n_nodes * (k + 1) / contigs;
                auto b = std::lower_bound(v.begin(), v.end(), (uint32_t)lo), e = std::lower_bound(v.begin(), v.end(), (uint32_t)hi);
                if (b == e) continue;
                out.put("P\tS", 3);
                out.num(s + 1);
                out.put("#", 1);
                out.num(h + 1);
                out.put("#ctg", 4);
                out.num(k + 1);
                out.put("\t", 1);
                for (auto it = b; it != e; ++it) {
                    if (it != b) out.put(",", 1);
                    out.num(*it);
                    out.put("+", 1);
                }
                out.put("\t*\n", 3);
                total_steps += (uint64_t)(e - b);
            }
        }
    return total_steps;
}

}  // namespace panacus
