// growth.cpp -- thresholds and the closed-form growth expectations (host only, f64).
//
// The printed TSV floors these values (src/io.rs:484), so a 1-ulp difference can flip a printed
// integer: every loop keeps the reference's operation order (src/graph_broker/hist.rs:21-187) and
// uses glibc log2/exp2 like a gnu-target Rust build; compile with -ffp-contract=off.
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <sstream>
#include <thread>

#include "panacus_host.hpp"

namespace panacus {

// ---- Threshold (src/util.rs:327-364) -------------------------------------------------------------------

uint64_t Threshold::to_absolute(uint64_t n) const {
    if (!is_relative) return abs;
    const double v = std::ceil((double)n * rel);
    return v > 0.0 ? (uint64_t)v : 0;  // Rust `as usize` saturates
}

double Threshold::to_relative(uint64_t n) const {
    if (is_relative) return rel;
    return (double)abs / (double)n;
}

std::string format_f64(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[512];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);  // shortest round-trip, no exponent
    return std::string(buf, r.ptr);
}

std::string format_f32(float v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[256];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

std::string Threshold::get_string() const { return is_relative ? format_f64(rel) : std::to_string(abs); }

// ---- ThresholdContainer::parse_params (src/graph_broker/hist.rs:207-323) -------------------------------------

namespace {

std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

std::vector<Threshold> parse_threshold_cli(const std::string &s, bool require_absolute) {
    std::vector<Threshold> out;
    std::stringstream ss(s);
    std::string el;
    size_t i = 0;
    while (std::getline(ss, el, ',')) {
        ++i;
        el = trim(el);
        if (require_absolute) {
            if (el.empty() || !std::all_of(el.begin(), el.end(), [](char c) { return c >= '0' && c <= '9'; }))
                throw Error("threshold \"" + s + "\" (" + std::to_string(i) + ". element in list) is required to be integer, but isn't.");
            out.push_back(Threshold::Absolute(std::stoull(el)));
        } else {
            double t;
            try {
                size_t pos = 0;
                t = std::stod(el, &pos);
                if (pos != el.size()) throw 1;
            } catch (...) {
                throw Error("threshold \"" + s + "\" (" + std::to_string(i) + ". element in list) is required to be float, but isn't.");
            }
            if (!(t >= 0.0 && t <= 1.0))
                throw Error("relative threshold \"" + s + "\" (" + std::to_string(i) + ". element in list) must be within [0,1].");
            out.push_back(Threshold::Relative(t));
        }
    }
    return out;
}

}  // namespace

ThresholdContainer ThresholdContainer::parse_params(const std::string &quorum, const std::string &coverage) {
    ThresholdContainer c;
    if (quorum.empty()) throw Error("quorum threshold setting requires at least one element, but none is given");
    c.quorum = parse_threshold_cli(quorum, false);
    if (coverage.empty()) throw Error("coverage threshold setting requires at least one element, but none is given");
    c.coverage = parse_threshold_cli(coverage, true);
    if (c.quorum.size() != c.coverage.size()) {
        if (c.quorum.size() == 1)
            c.quorum.assign(c.coverage.size(), c.quorum[0]);
        else if (c.coverage.size() == 1)
            c.coverage.assign(c.quorum.size(), c.coverage[0]);
        else
            throw Error("number of coverage and quorum threshold must match, or either one must have a single value");
    }
    return c;
}

// ---- closed-form growth ---------------------------------------------------------------------------------------

double choose(uint64_t n, uint64_t k) {  // hist.rs:21-36 (log2 of the binomial coefficient)
    double res = 0.0;
    if (k > n) return 0.0;
    const uint64_t kk = std::min(k, n - k);
    const double nf = (double)n;
    for (uint64_t i = 0; i < kk; ++i) {
        res += std::log2(nf - (double)i);
        res -= std::log2((double)i + 1.0);
    }
    return res;
}

namespace {

std::vector<double> growth_union(const std::vector<uint64_t> &h, const Threshold &t_cov) {  // hist.rs:89-114
    const uint64_t n = h.size() - 1;
    const uint64_t c = std::max<uint64_t>(1, t_cov.to_absolute(n));
    double n_fall_m = 0.0;
    uint64_t tot_u = 0;
    for (uint64_t i = c; i <= n; ++i) tot_u += h[i];
    const double tot = (double)tot_u;
    std::vector<double> perc_mult(n + 1, 0.0), out(n, 0.0);
    for (uint64_t m = 1; m <= n; ++m) {
        double y = 0.0;
        n_fall_m += std::log2((double)n - (double)m + 1.0);
        for (uint64_t i = c; i < n - m + 1; ++i) {
            perc_mult[i] += std::log2((double)n - (double)m - (double)i + 1.0);
            y += std::exp2(std::log2((double)h[i]) + perc_mult[i] - n_fall_m);
        }
        out[m - 1] = tot - y;
    }
    return out;
}

std::vector<double> growth_core(const std::vector<uint64_t> &h, const Threshold &t_cov) {  // hist.rs:116-138
    const uint64_t n = h.size() - 1;
    const uint64_t c = std::max<uint64_t>(1, t_cov.to_absolute(n + 1));
    double n_fall_m = 0.0;
    std::vector<double> perc_mult(n + 1, 0.0), out(n, 0.0);
    for (uint64_t m = 1; m <= n; ++m) {
        double y = 0.0;
        n_fall_m += std::log2((double)n - (double)m + 1.0);
        for (uint64_t i = std::max(m, c); i <= n; ++i) {
            perc_mult[i] += std::log2((double)i - (double)m + 1.0);
            y += std::exp2(std::log2((double)h[i]) + perc_mult[i] - n_fall_m);
        }
        out[m - 1] = y;
    }
    return out;
}

// fn(k) for k in [0, n) on up to `nthreads` threads (dynamic hand-out: the rows differ in cost)
template <typename F>
void parallel_rows(size_t n, unsigned nthreads, F fn) {
    nthreads = std::max(1u, std::min<unsigned>(nthreads, (unsigned)std::max<size_t>(n, 1)));
    if (nthreads == 1) {
        for (size_t k = 0; k < n; ++k) fn(k);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; ++t)
        pool.emplace_back([&] {
            for (size_t k; (k = next.fetch_add(1)) < n;) fn(k);
        });
    for (auto &th : pool) th.join();
}

// hist.rs:140-187.  The reference walks m = 1..n and, inside, every coverage class i with its memo row q[i][*]; rows
// of different i never touch each other.  Here each row i is carried through all m on its own (threads over i, the
// per-(i, j) update sequence and every f64 expression unchanged), leaving its contribution to yr in term[i][m]; the
// sums over i are then taken serially in the reference's order (i ascending), so every intermediate rounding is the
// same and the result is bit-identical to the serial loop -- and to a one-thread run.
std::vector<double> growth_quorum(const std::vector<uint64_t> &h, const Threshold &t_cov, const Threshold &t_quorum,
                                  unsigned nthreads) {
    const uint64_t n = h.size() - 1;
    const uint64_t c = std::max<uint64_t>(1, t_cov.to_absolute(n));
    const double quorum = t_quorum.to_relative(n);
    std::vector<double> perc_mult(n + 1, 0.0), out(n, 0.0);
    // the m-only recurrences, with the reference's operation order
    std::vector<double> m_fact(n + 1, 0.0), n_fall_m(n + 1, 0.0);
    std::vector<uint64_t> m_quorum(n + 1, 0);
    for (uint64_t m = 1; m <= n; ++m) {
        m_fact[m] = m_fact[m - 1] + std::log2((double)m);
        n_fall_m[m] = n_fall_m[m - 1] + std::log2((double)n - (double)m + 1.0);
        const double mq = std::ceil((double)m * quorum);
        m_quorum[m] = mq > 0.0 ? (uint64_t)mq : 0;
    }
    // log2 of every integer the inner loop can ask for, from the same libm call (the arguments are exact integers in f64,
    // so lg[k] is bit for bit what std::log2((double)k) returns there); two of the three transcendental calls per step
    std::vector<double> lg(2 * n + 3);
    for (uint64_t k = 0; k < lg.size(); ++k) lg[k] = std::log2((double)k);
    auto choose_lg = [&](uint64_t nn, uint64_t k) {  // choose() of hist.rs:21-36 on the table
        double res = 0.0;
        if (k > nn) return 0.0;
        const uint64_t kk = std::min(k, nn - k);
        for (uint64_t i = 0; i < kk; ++i) {
            res += lg[nn - i];
            res -= lg[i + 1];
        }
        return res;
    };
    // yr contributions: term[i * n + (m - 1)], NaN = "no j qualified" (the reference's `add` flag stays false)
    std::vector<double> term((size_t)n * n, std::nan(""));
    parallel_rows(n, nthreads, [&](size_t i) {
        std::vector<double> qi(n + 1, 0.0);
        double *ti = term.data() + i * n;
        const double log2_hi = std::log2((double)h[i]);
        for (uint64_t m = 1; m <= n; ++m) {
            if (i < m_quorum[m]) continue;  // the reference's loop over i starts at m_quorum
            double sum_q = 0.0;
            bool add = false;
            for (uint64_t j = std::max(m_quorum[m], c); j < m; ++j) {
                if (n + j + 1 > i + m && j <= i) {
                    if (qi[j] == 0.0) qi[j] = choose_lg(i, j);
                    qi[j] += lg[n - i - m + 1 + j];  // log2(n - i - m + 1 + j): n + j + 1 > i + m keeps the index >= 1
                    qi[j] -= lg[m - j];
                    sum_q += std::exp2(qi[j] + m_fact[m] - n_fall_m[m]);
                    add = true;
                }
            }
            if (add) ti[m - 1] = std::exp2(log2_hi + std::log2(sum_q));
        }
    });
    for (uint64_t m = 1; m <= n; ++m) {
        double yl = 0.0;
        for (uint64_t i = std::max(m, c); i <= n; ++i) {
            perc_mult[i] += std::log2((double)i - (double)m + 1.0);
            yl += std::exp2(std::log2((double)h[i]) + perc_mult[i] - n_fall_m[m]);
        }
        double yr = 0.0;
        for (uint64_t i = m_quorum[m]; i < n; ++i) {
            const double t = term[(size_t)i * n + (m - 1)];
            if (t == t) yr += t;  // skip the "not added" marker; a computed term is never NaN (h[i] = 0 gives exp2(-inf) = 0)
        }
        out[m - 1] = yl + yr;
    }
    return out;
}

}  // namespace

std::vector<double> Hist::calc_growth(const Threshold &t_coverage, const Threshold &t_quorum) const {  // hist.rs:51-66
    const uint64_t n = coverage.size() - 1;
    if (n == 0) return {};
    const uint64_t quorum = std::max<uint64_t>(1, t_quorum.to_absolute(n));
    if (quorum == 1) return growth_union(coverage, t_coverage);
    if (quorum >= n) return growth_core(coverage, t_coverage);
    return growth_quorum(coverage, t_coverage, t_quorum, growth_threads_);
}

std::vector<std::vector<double>> Hist::calc_all_growths(const ThresholdContainer &aux) const {  // hist.rs:69-87
    // the reference fans the pairs out on rayon (par_iter over coverage.zip(quorum)); here the pairs run on threads and
    // each quorum pair spreads its coverage classes over its share of the thread budget (see growth_quorum)
    const size_t T = aux.coverage.size();
    std::vector<std::vector<double>> out(T);
    // quorum pairs (O(n^3)) take the whole budget one after the other; with many cheap pairs the pairs share it
    const unsigned budget = host_thread_budget();
    const unsigned outer = T > 4 * (size_t)budget ? budget : 1u;
    Hist worker = *this;
    worker.growth_threads_ = std::max(1u, budget / outer);
    parallel_rows(T, outer, [&](size_t k) {
        std::vector<double> g = worker.calc_growth(aux.coverage[k], aux.quorum[k]);
        g.insert(g.begin(), std::nan(""));
        out[k] = std::move(g);
    });
    return out;
}

}  // namespace panacus
