// cluster.cpp -- hierarchical clustering of the similarity table (reference src/analyses/similarity.rs:165-181).
//
// The reference calls `kodama::linkage` (kodama 0.3.0, Cargo.toml:34; the crate is NOT in the reference tree and cannot
// be built here).  kodama documents itself as a port of Daniel Muellner's fastcluster, "Modern hierarchical,
// agglomerative clustering algorithms" (arXiv:1109.2378); what is restated here is that published scheme, with the
// dispatch and the data flow kodama documents:
//   * single            -> MST-LINKAGE (Prim order: vertex k joins the cluster of vertex k - 1 at its key)
//   * complete, average, weighted, ward -> NN-CHAIN (reciprocal nearest neighbours, chain kept across merges)
//   * centroid, median  -> GENERIC-LINKAGE (nearest-neighbour candidates in a binary min-heap, lazily repaired)
//   * ward / centroid / median run on SQUARED dissimilarities (the step heights are square-rooted afterwards)
//   * MST and NN-chain steps are sorted by height (stable) and then labelled with a union-find: the cluster created by
//     sorted step i is n + i, and a step lists the smaller label first; the generic algorithm already emits its steps
//     in merge order with those labels.
// Arithmetic is done in the table's own type (f32 in the product, like `linkage::<f32>`); the Lance-Williams updates are
// written in the plain textbook form (sizes converted to the float type, one division).  Tie-breaking (strict `<` when
// scanning for minima, lower index first) follows the published pseudo-code.  What cannot be checked here is kodama's
// own source: the row ORDER of the similarity TSV stays "parity unpinned" (DESIGN.md section 5); the merge structure is
// cross-checked against scipy.cluster.hierarchy.linkage on tie-free inputs (tests/test_cluster.py).
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>

#include "panacus_host.hpp"

namespace panacus {

namespace {

enum class Method { Single, Complete, Average, Weighted, Ward, Centroid, Median };

Method parse_method(const std::string &m) {
    if (m == "single") return Method::Single;
    if (m == "complete") return Method::Complete;
    if (m == "average") return Method::Average;
    if (m == "weighted") return Method::Weighted;
    if (m == "ward") return Method::Ward;
    if (m == "centroid") return Method::Centroid;
    if (m == "median") return Method::Median;
    throw Error("unknown cluster method '" + m + "' (single, complete, average, weighted, ward, centroid, median)");
}

template <typename T>
struct Step {
    size_t c1, c2;
    T d;
};

// condensed upper triangle, row-major: (i, j) with i < j
template <typename T>
struct Condensed {
    std::vector<T> &v;
    size_t n;
    T &operator()(size_t i, size_t j) { return v[i * n - i * (i + 1) / 2 + (j - i - 1)]; }  // i < j
    T &sym(size_t i, size_t j) { return i < j ? (*this)(i, j) : (*this)(j, i); }
};

// the active observations as a doubly linked list over 0..n (n = end sentinel)
struct ActiveList {
    std::vector<size_t> succ, pred;
    size_t start = 0;
    explicit ActiveList(size_t n) : succ(n + 1), pred(n + 1) {
        for (size_t i = 0; i <= n; ++i) {
            succ[i] = i + 1;
            pred[i] = i ? i - 1 : 0;
        }
    }
    void remove(size_t i) {
        if (i == start)
            start = succ[i];
        else {
            succ[pred[i]] = succ[i];
            pred[succ[i]] = pred[i];
        }
    }
};

// Lance-Williams updates: new dissimilarity between the merged cluster (a + b) and x, written over d(b, x)
template <typename T>
T lw_update(Method m, T dax, T dbx, T dab, size_t na, size_t nb, size_t nx) {
    const T sa = (T)na, sb = (T)nb, sx = (T)nx;
    switch (m) {
        case Method::Single: return std::min(dax, dbx);
        case Method::Complete: return std::max(dax, dbx);
        case Method::Average: return (sa * dax + sb * dbx) / (sa + sb);
        case Method::Weighted: return (T)0.5 * (dax + dbx);
        case Method::Ward: return ((sx + sa) * dax + (sx + sb) * dbx - sx * dab) / (sa + sb + sx);
        case Method::Centroid: {
            const T sab = sa + sb;
            return (sa * dax + sb * dbx) / sab - (sa * sb * dab) / (sab * sab);
        }
        default: return (T)0.5 * (dax + dbx) - dab / (T)4;  // median
    }
}

// sort (stable) + union-find labelling: cluster of sorted step i = n + i, smaller label first
template <typename T>
void sort_and_label(std::vector<Step<T>> &steps, size_t n) {
    std::stable_sort(steps.begin(), steps.end(), [](const Step<T> &x, const Step<T> &y) { return x.d < y.d; });
    std::vector<size_t> parent(2 * n, (size_t)-1);
    size_t next = n;
    auto find = [&](size_t x) {
        size_t r = x;
        while (parent[r] != (size_t)-1) r = parent[r];
        while (parent[x] != (size_t)-1) {  // path compression
            const size_t p = parent[x];
            parent[x] = r;
            x = p;
        }
        return r;
    };
    for (auto &s : steps) {
        const size_t a = find(s.c1), b = find(s.c2);
        parent[a] = parent[b] = next++;
        s.c1 = std::min(a, b);
        s.c2 = std::max(a, b);
    }
}

template <typename T>
void mst_linkage(Condensed<T> D, std::vector<Step<T>> &steps) {
    const size_t n = D.n;
    ActiveList active(n);
    std::vector<T> d(n, std::numeric_limits<T>::infinity());
    size_t idx2 = 1;
    T min = std::numeric_limits<T>::infinity();
    for (size_t i = 1; i < n; ++i) {
        d[i] = D(0, i);
        if (d[i] < min) {
            min = d[i];
            idx2 = i;
        }
    }
    steps.push_back({0, idx2, min});
    for (size_t j = 1; j + 1 < n; ++j) {
        const size_t prev = idx2;
        active.remove(prev);
        idx2 = active.succ[0];
        min = d[idx2];
        size_t i;
        for (i = idx2; i < prev; i = active.succ[i]) {
            const T t = D(i, prev);
            if (t < d[i]) d[i] = t;
            if (d[i] < min) {
                min = d[i];
                idx2 = i;
            }
        }
        for (; i < n; i = active.succ[i]) {
            const T t = D(prev, i);
            if (d[i] > t) d[i] = t;
            if (d[i] < min) {
                min = d[i];
                idx2 = i;
            }
        }
        steps.push_back({prev, idx2, min});
    }
    sort_and_label(steps, n);
}

template <typename T>
void nn_chain(Condensed<T> D, Method m, std::vector<Step<T>> &steps) {
    const size_t n = D.n;
    ActiveList active(n);
    std::vector<size_t> members(n, 1), chain(n);
    size_t tip = 0, idx1 = 0, idx2 = 0;
    T min = 0;
    for (size_t j = 0; j + 1 < n; ++j) {
        if (tip <= 3) {  // (re)start the chain at the first active observation
            chain[0] = idx1 = active.start;
            tip = 1;
            idx2 = active.succ[idx1];
            min = D(idx1, idx2);
            for (size_t i = active.succ[idx2]; i < n; i = active.succ[i])
                if (D(idx1, i) < min) {
                    min = D(idx1, i);
                    idx2 = i;
                }
        } else {  // the two clusters merged last sat at the tip: resume two elements below them
            tip -= 3;
            idx1 = chain[tip - 1];
            idx2 = chain[tip];
            min = D.sym(idx1, idx2);
        }
        do {  // follow nearest neighbours until a reciprocal pair is found
            chain[tip] = idx2;
            for (size_t i = active.start; i < idx2; i = active.succ[i])
                if (D(i, idx2) < min) {
                    min = D(i, idx2);
                    idx1 = i;
                }
            for (size_t i = active.succ[idx2]; i < n; i = active.succ[i])
                if (D(idx2, i) < min) {
                    min = D(idx2, i);
                    idx1 = i;
                }
            idx2 = idx1;
            idx1 = chain[tip++];
        } while (idx2 != chain[tip - 2]);
        steps.push_back({idx1, idx2, min});
        if (idx1 > idx2) std::swap(idx1, idx2);
        const size_t na = members[idx1], nb = members[idx2];
        members[idx2] += members[idx1];
        active.remove(idx1);  // the merged cluster lives on in the larger index
        size_t i;
        for (i = active.start; i < idx1; i = active.succ[i]) D(i, idx2) = lw_update(m, D(i, idx1), D(i, idx2), min, na, nb, members[i]);
        for (; i < idx2; i = active.succ[i]) D(i, idx2) = lw_update(m, D(idx1, i), D(i, idx2), min, na, nb, members[i]);
        for (i = active.succ[idx2]; i < n; i = active.succ[i]) D(idx2, i) = lw_update(m, D(idx1, i), D(idx2, i), min, na, nb, members[i]);
    }
    sort_and_label(steps, n);
}

// binary min-heap over A[0 .. size) with position <-> element index maps (I: heap position -> element, R: inverse)
template <typename T>
struct MinHeap {
    std::vector<T> &A;
    size_t size;
    std::vector<size_t> I, R;
    MinHeap(std::vector<T> &a, size_t n) : A(a), size(n), I(n), R(n) {
        std::iota(I.begin(), I.end(), (size_t)0);
        std::iota(R.begin(), R.end(), (size_t)0);
    }
    T H(size_t i) const { return A[I[i]]; }
    void swap_(size_t i, size_t j) {
        std::swap(I[i], I[j]);
        R[I[i]] = i;
        R[I[j]] = j;
    }
    void up(size_t i) {
        for (size_t j; i > 0 && H(i) < H(j = (i - 1) >> 1); i = j) swap_(i, j);
    }
    void down(size_t i) {
        for (size_t j; (j = 2 * i + 1) < size; i = j) {
            if (H(j) >= H(i)) {
                ++j;
                if (j >= size || H(j) >= H(i)) break;
            } else if (j + 1 < size && H(j + 1) < H(j)) {
                ++j;
            }
            swap_(i, j);
        }
    }
    void heapify() {
        for (size_t idx = size >> 1; idx > 0;) down(--idx);
    }
    size_t argmin() const { return I[0]; }
    void pop() {
        --size;
        I[0] = I[size];
        R[I[0]] = 0;
        down(0);
    }
    void update_leq(size_t idx, T val) {
        A[idx] = val;
        up(R[idx]);
    }
    void update_geq(size_t idx, T val) {
        A[idx] = val;
        down(R[idx]);
    }
    void update(size_t idx, T val) {
        if (val <= A[idx])
            update_leq(idx, val);
        else
            update_geq(idx, val);
    }
};

template <typename T>
void generic_linkage(Condensed<T> D, Method m, std::vector<Step<T>> &steps) {
    const size_t n = D.n, n1 = n - 1;
    std::vector<size_t> nghbr(n1), row_repr(n), members(n, 1);
    std::vector<T> mindist(n1);
    ActiveList active(n);
    std::iota(row_repr.begin(), row_repr.end(), (size_t)0);
    for (size_t i = 0; i < n1; ++i) {  // nearest neighbour among the higher indices
        T min = std::numeric_limits<T>::infinity();
        size_t idx = i + 1;
        for (size_t j = i + 1; j < n; ++j)
            if (D(i, j) < min) {
                min = D(i, j);
                idx = j;
            }
        mindist[i] = min;
        nghbr[i] = idx;
    }
    MinHeap<T> heap(mindist, n1);
    heap.heapify();
    for (size_t it = 0; it < n1; ++it) {
        // mindist[i] is a lower bound of min_{j > i} D(i, j); repair the smallest candidate until it is exact
        size_t idx1 = heap.argmin();
        while (mindist[idx1] < D(idx1, nghbr[idx1])) {
            size_t j = active.succ[idx1];
            nghbr[idx1] = j;
            T min = D(idx1, j);
            for (j = active.succ[j]; j < n; j = active.succ[j])
                if (D(idx1, j) < min) {
                    min = D(idx1, j);
                    nghbr[idx1] = j;
                }
            heap.update_geq(idx1, min);
            idx1 = heap.argmin();
        }
        heap.pop();
        const size_t idx2 = nghbr[idx1];
        const size_t node1 = row_repr[idx1], node2 = row_repr[idx2];
        const size_t na = members[idx1], nb = members[idx2];
        members[idx2] += members[idx1];
        const T dab = mindist[idx1];
        steps.push_back({std::min(node1, node2), std::max(node1, node2), dab});
        active.remove(idx1);
        row_repr[idx2] = n + it;
        // centroid / median: distances may shrink below both old ones, so every row is checked against its candidate
        size_t j;
        for (j = active.start; j < idx1; j = active.succ[j]) {
            D(j, idx2) = lw_update(m, D(j, idx1), D(j, idx2), dab, na, nb, members[j]);
            if (D(j, idx2) < mindist[j]) {
                heap.update_leq(j, D(j, idx2));
                nghbr[j] = idx2;
            } else if (nghbr[j] == idx1) {
                nghbr[j] = idx2;
            }
        }
        for (; j < idx2; j = active.succ[j]) {
            D(j, idx2) = lw_update(m, D(idx1, j), D(j, idx2), dab, na, nb, members[j]);
            if (D(j, idx2) < mindist[j]) {
                heap.update_leq(j, D(j, idx2));
                nghbr[j] = idx2;
            }
        }
        if (idx2 < n1) {
            j = active.succ[idx2];
            if (j < n) {
                nghbr[idx2] = j;
                D(idx2, j) = lw_update(m, D(idx1, j), D(idx2, j), dab, na, nb, members[j]);
                T min = D(idx2, j);
                for (j = active.succ[j]; j < n; j = active.succ[j]) {
                    D(idx2, j) = lw_update(m, D(idx1, j), D(idx2, j), dab, na, nb, members[j]);
                    if (D(idx2, j) < min) {
                        min = D(idx2, j);
                        nghbr[idx2] = j;
                    }
                }
                heap.update(idx2, min);
            }
        }
    }
}

template <typename T>
std::vector<Step<T>> linkage(std::vector<T> cond, size_t n, Method m) {
    std::vector<Step<T>> steps;
    if (n < 2) return steps;
    const bool squares = m == Method::Ward || m == Method::Centroid || m == Method::Median;
    if (squares)
        for (auto &x : cond) x = x * x;
    Condensed<T> D{cond, n};
    if (m == Method::Single)
        mst_linkage(D, steps);
    else if (m == Method::Centroid || m == Method::Median)
        generic_linkage(D, m, steps);
    else
        nn_chain(D, m, steps);
    if (squares)
        for (auto &s : steps) s.d = std::sqrt(s.d);
    return steps;
}

}  // namespace

// euclidean distances between the table's rows, f32, accumulated left to right like similarity.rs:238-253
// (`(v1 - v2).powf(2.0)`: x * x is the correctly rounded square either way)
static std::vector<float> condensed_distances(const std::vector<std::vector<float>> &table) {
    const size_t n = table.size();
    std::vector<float> cond;
    cond.reserve(n * (n - 1) / 2);
    for (size_t i = 0; i + 1 < n; ++i)
        for (size_t j = i + 1; j < n; ++j) {
            float s = 0.f;
            for (size_t k = 0; k < table[i].size(); ++k) {
                const float d = table[i][k] - table[j][k];
                s += d * d;
            }
            cond.push_back(std::sqrt(s));
        }
    return cond;
}

// leaves in dendrogram-step order (get_order_from_dendrogram, similarity.rs:206-219)
std::vector<size_t> cluster_leaf_order(const std::vector<std::vector<float>> &table, const std::string &method) {
    const size_t n = table.size();
    std::vector<size_t> leaves;
    if (n < 2) {
        for (size_t i = 0; i < n; ++i) leaves.push_back(i);
        return leaves;
    }
    for (const auto &s : linkage<float>(condensed_distances(table), n, parse_method(method))) {
        if (s.c1 < n) leaves.push_back(s.c1);
        if (s.c2 < n) leaves.push_back(s.c2);
    }
    return leaves;
}

// debug / test hook: the dendrogram of a condensed distance vector in f64 or f32, one "c1 c2 height" line per step
std::string debug_linkage(const std::vector<double> &cond, size_t n, const std::string &method, bool f32) {
    std::string out;
    char buf[96];
    if (f32) {
        std::vector<float> c(cond.begin(), cond.end());
        for (const auto &s : linkage<float>(c, n, parse_method(method))) {
            snprintf(buf, sizeof buf, "%zu %zu %.9g\n", s.c1, s.c2, (double)s.d);
            out += buf;
        }
    } else {
        for (const auto &s : linkage<double>(cond, n, parse_method(method))) {
            snprintf(buf, sizeof buf, "%zu %zu %.17g\n", s.c1, s.c2, s.d);
            out += buf;
        }
    }
    return out;
}

}  // namespace panacus
