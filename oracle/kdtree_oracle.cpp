// TEST INFRASTRUCTURE ONLY -- the CPU oracle ("port") for the nearest-neighbour hot path.
// Nothing under point-cloud-utils_b200/ may import, link or call this file; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// A from-scratch restatement (no nanoflann include, no Eigen, no pybind11) of what the reference
// computes on this path, so that it can run on the GPU box where /root/reference does not exist.
// Parity status: PINNED -- tests/test_oracle.py checks it bit-for-bit (indices, distances, tie
// order, padding) against oracle/_ref (the reference's own nanoflann header compiled in place)
// and against the golden vectors in tests/golden/ that oracle/make_golden.py generated from
// oracle/_ref.
//
// Reference locations restated here (all under /root/reference):
//   tree build    external/nanoflann/nanoflann.hpp:1363-1375 (buildIndex), :1491-1536 (init_vind,
//                 computeBoundingBox), :1001-1059 (divideTree), :1061-1110 (middleSplit_),
//                 :1121-1162 (planeSplit), :986-999 (computeMinMax)
//   search        :1394-1418 (findNeighbors), :1164-1187 (computeInitialDistances),
//                 :1545-1624 (searchLevel), :496-507 (L2_Simple evalMetric), :509-513 (accum_dist)
//   result set    :157-230 (KNNResultSet: init / addPoint / worstDist)
//   driver        src/point_cloud_distance.cpp:21-99 (shortest_distances_nanoflann),
//                 :186-234 (one_sided_hausdorff_distance), src/common/common.h:182-212
//
// Arithmetic contract (what makes indices bit-exact): every distance is
// ((qx-px)^2 + (qy-py)^2) + (qz-pz)^2 with each operation rounded in the input precision, query
// minus data, no FMA (compile with -ffp-contract=off); equal distances keep kd-tree visit order.
#include <cstdint>
#include <cstddef>
#include <cmath>
#include <limits>
#include <thread>
#include <utility>
#include <vector>
#include <algorithm>
#if defined(_OPENMP)
#include <omp.h>
#endif

namespace {

template <typename T>
struct KdOracle {
    struct Node {
        int32_t feat;        // split dimension, -1 for a leaf
        T div_lo, div_hi;    // max of the left subtree / min of the right subtree along feat
        int64_t first, last; // leaf: range [first, last) into order[]
        int32_t kid[2];
    };

    const T* pts;
    int64_t m;
    int64_t leaf_cap;
    std::vector<int64_t> order;  // the permutation nanoflann calls vAcc
    std::vector<Node> nodes;
    T root_lo[3], root_hi[3];

    T at(int64_t slot, int dim) const { return pts[3 * order[slot] + dim]; }

    // nanoflann.hpp:986-999
    void span_of(int64_t first, int64_t count, int dim, T& lo, T& hi) const {
        lo = hi = at(first, dim);
        for (int64_t i = 1; i < count; ++i) {
            const T v = at(first + i, dim);
            if (v < lo) lo = v;
            if (v > hi) hi = v;
        }
    }

    // nanoflann.hpp:1121-1162.  Two sweeps of a two-pointer exchange partition; the index type
    // there is unsigned, hence the explicit "right pointer reached slot 0" exits.
    void partition_about(int64_t first, uint64_t count, int dim, T cut, uint64_t& n_less, uint64_t& n_less_eq) {
        uint64_t lo = 0, hi = count - 1;
        for (;;) {
            while (lo <= hi && at(first + lo, dim) < cut) ++lo;
            while (hi != 0 && lo <= hi && at(first + hi, dim) >= cut) --hi;
            if (lo > hi || hi == 0) break;
            std::swap(order[first + lo], order[first + hi]);
            ++lo;
            --hi;
        }
        n_less = lo;
        hi = count - 1;
        for (;;) {
            while (lo <= hi && at(first + lo, dim) <= cut) ++lo;
            while (hi != 0 && lo <= hi && at(first + hi, dim) > cut) --hi;
            if (lo > hi || hi == 0) break;
            std::swap(order[first + lo], order[first + hi]);
            ++lo;
            --hi;
        }
        n_less_eq = lo;
    }

    // nanoflann.hpp:1061-1110
    void choose_split(int64_t first, uint64_t count, const T lo[3], const T hi[3], uint64_t& left_count,
                      int& feat, T& cut) {
        const T eps = static_cast<T>(0.00001);
        T widest = hi[0] - lo[0];
        for (int d = 1; d < 3; ++d) {
            const T w = hi[d] - lo[d];
            if (w > widest) widest = w;
        }
        T best_spread = -1;
        feat = 0;
        for (int d = 0; d < 3; ++d) {
            const T w = hi[d] - lo[d];
            if (w > (1 - eps) * widest) {
                T a, b;
                span_of(first, (int64_t)count, d, a, b);
                const T spread = b - a;
                if (spread > best_spread) {
                    feat = d;
                    best_spread = spread;
                }
            }
        }
        const T mid = (lo[feat] + hi[feat]) / 2;
        T a, b;
        span_of(first, (int64_t)count, feat, a, b);
        if (mid < a) cut = a;
        else if (mid > b) cut = b;
        else cut = mid;

        uint64_t n_less, n_less_eq;
        partition_about(first, count, feat, cut, n_less, n_less_eq);
        if (n_less > count / 2) left_count = n_less;
        else if (n_less_eq < count / 2) left_count = n_less_eq;
        else left_count = count / 2;
    }

    // nanoflann.hpp:1001-1059.  lo/hi: on entry the loose box inherited from the parent, on exit
    // the tight box of the points below this node.
    int32_t grow(int64_t first, int64_t last, T lo[3], T hi[3]) {
        const int32_t me = (int32_t)nodes.size();
        nodes.push_back(Node());
        if ((last - first) <= leaf_cap) {
            Node nd;
            nd.feat = -1; nd.div_lo = nd.div_hi = 0; nd.first = first; nd.last = last; nd.kid[0] = nd.kid[1] = -1;
            nodes[me] = nd;
            for (int d = 0; d < 3; ++d) lo[d] = hi[d] = at(first, d);
            for (int64_t s = first + 1; s < last; ++s)
                for (int d = 0; d < 3; ++d) {
                    const T v = at(s, d);
                    if (lo[d] > v) lo[d] = v;
                    if (hi[d] < v) hi[d] = v;
                }
            return me;
        }
        uint64_t left_count; int feat; T cut;
        choose_split(first, (uint64_t)(last - first), lo, hi, left_count, feat, cut);

        T llo[3], lhi[3], rlo[3], rhi[3];
        for (int d = 0; d < 3; ++d) { llo[d] = rlo[d] = lo[d]; lhi[d] = rhi[d] = hi[d]; }
        lhi[feat] = cut;
        rlo[feat] = cut;
        const int32_t k0 = grow(first, first + (int64_t)left_count, llo, lhi);
        const int32_t k1 = grow(first + (int64_t)left_count, last, rlo, rhi);
        Node nd;
        nd.feat = feat; nd.div_lo = lhi[feat]; nd.div_hi = rlo[feat]; nd.first = first; nd.last = last;
        nd.kid[0] = k0; nd.kid[1] = k1;
        nodes[me] = nd;
        for (int d = 0; d < 3; ++d) { lo[d] = std::min(llo[d], rlo[d]); hi[d] = std::max(lhi[d], rhi[d]); }
        return me;
    }

    // nanoflann.hpp:1363-1375 + :1491-1536
    KdOracle(const T* points, int64_t count, int64_t leaf) : pts(points), m(count), leaf_cap(leaf) {
        order.resize(m);
        for (int64_t i = 0; i < m; ++i) order[i] = i;
        if (m == 0) return;
        for (int d = 0; d < 3; ++d) root_lo[d] = root_hi[d] = at(0, d);
        for (int64_t s = 1; s < m; ++s)
            for (int d = 0; d < 3; ++d) {
                const T v = at(s, d);
                if (v < root_lo[d]) root_lo[d] = v;
                if (v > root_hi[d]) root_hi[d] = v;
            }
        nodes.reserve((size_t)(2 * m / std::max<int64_t>(1, leaf_cap) + 16));
        grow(0, m, root_lo, root_hi);
    }

    // nanoflann.hpp:157-230
    struct Best {
        int64_t* id; T* d2; int64_t cap; int64_t have;
        Best(int64_t* ids, T* ds, int64_t k) : id(ids), d2(ds), cap(k), have(0) {
            if (cap) d2[cap - 1] = (std::numeric_limits<T>::max)();
        }
        T worst() const { return d2[cap - 1]; }
        void offer(T dist, int64_t index) {
            int64_t i = have;
            for (; i > 0; --i) {
                if (d2[i - 1] > dist) {       // strict: equal distances keep arrival order
                    if (i < cap) { d2[i] = d2[i - 1]; id[i] = id[i - 1]; }
                } else break;
            }
            if (i < cap) { d2[i] = dist; id[i] = index; }
            if (have < cap) ++have;
        }
    };

    // nanoflann.hpp:496-507
    T metric(const T* q, int64_t p) const {
        T acc = T();
        for (int d = 0; d < 3; ++d) {
            const T diff = q[d] - pts[3 * p + d];
            acc += diff * diff;
        }
        return acc;
    }

    // nanoflann.hpp:1545-1624
    void descend(Best& best, const T* q, int32_t at_node, T bound, T off[3]) const {
        const Node& nd = nodes[at_node];
        if (nd.feat < 0) {
            const T worst_on_entry = best.worst();  // cached for the whole leaf (:1555)
            for (int64_t s = nd.first; s < nd.last; ++s) {
                const int64_t p = order[s];
                const T d = metric(q, p);
                if (d < worst_on_entry) best.offer(d, p);
            }
            return;
        }
        const int f = nd.feat;
        const T v = q[f];
        const T d1 = v - nd.div_lo;
        const T d2 = v - nd.div_hi;
        int32_t near_kid, far_kid;
        T cut;
        if ((d1 + d2) < 0) { near_kid = nd.kid[0]; far_kid = nd.kid[1]; cut = (v - nd.div_hi) * (v - nd.div_hi); }
        else               { near_kid = nd.kid[1]; far_kid = nd.kid[0]; cut = (v - nd.div_lo) * (v - nd.div_lo); }
        descend(best, q, near_kid, bound, off);
        const T saved = off[f];
        bound = bound + cut - saved;
        off[f] = cut;
        const float eps_error = 1.0f;  // 1 + SearchParams().eps (:667, :1405)
        if (bound * eps_error <= best.worst()) descend(best, q, far_kid, bound, off);
        off[f] = saved;
    }

    // nanoflann.hpp:1429-1439 + :1394-1418 + :1164-1187
    int64_t knn(const T* q, int64_t k, int64_t* ids, T* d2) const {
        Best best(ids, d2, k);
        if (m == 0) return 0;
        T off[3] = {0, 0, 0};
        T bound = T();
        for (int d = 0; d < 3; ++d) {
            if (q[d] < root_lo[d]) { off[d] = (q[d] - root_lo[d]) * (q[d] - root_lo[d]); bound += off[d]; }
            if (q[d] > root_hi[d]) { off[d] = (q[d] - root_hi[d]) * (q[d] - root_hi[d]); bound += off[d]; }
        }
        descend(best, q, 0, bound, off);
        return best.have;
    }
};

int thread_policy(int64_t n, int num_threads) {
    // src/point_cloud_distance.cpp:29-31 + src/common/common.h:194-199
    const bool parallel = n >= 100000 && num_threads != 0;
    if (!parallel) return 1;
    if (num_threads < 0) return std::max(1, (int)std::thread::hardware_concurrency());
    return num_threads;
}

// src/point_cloud_distance.cpp:21-99 (single tree build: the reference's three builds produce the
// same tree)
template <typename T>
void shortest_distances(const T* query, int64_t n, const T* dataset, int64_t m, int k, int squared, int leaf,
                        int num_threads, T* out_d, int64_t* out_i) {
    const KdOracle<T> tree(dataset, m, leaf);
    const int nthr = thread_policy(n, num_threads);
    (void)nthr;
#if defined(_OPENMP)
#pragma omp parallel num_threads(nthr) if (nthr > 1)
#endif
    {
        std::vector<int64_t> ids(k);
        std::vector<T> d2(k);
#if defined(_OPENMP)
#pragma omp for
#endif
        for (int64_t i = 0; i < n; ++i) {
            const int64_t found = tree.knn(query + 3 * i, k, ids.data(), d2.data());
            for (int64_t c = 0; c < found; ++c) {
                out_i[i * k + c] = ids[c];
                out_d[i * k + c] = squared ? d2[c] : std::sqrt(d2[c]);
            }
            for (int64_t c = found; c < k; ++c) { out_i[i * k + c] = -1; out_d[i * k + c] = (T)-1.0; }
        }
    }
}

// src/point_cloud_distance.cpp:219-225
template <typename T>
void one_sided(const T* src, int64_t n, const T* dst, int64_t m, int squared, int leaf, T* out_max,
               int64_t* out_i, int64_t* out_j) {
    std::vector<T> dist(n);
    std::vector<int64_t> corr(n);
    shortest_distances<T>(src, n, dst, m, 1, squared, leaf, 0, dist.data(), corr.data());
    int64_t arg = 0;
    for (int64_t i = 1; i < n; ++i)
        if (dist[i] > dist[arg]) arg = i;  // first maximum
    *out_max = dist[arg]; *out_i = arg; *out_j = corr[arg];
}

// Exposes the built tree so that GPU-side replicas of the build can be checked slot for slot.
template <typename T>
int64_t dump_tree(const T* dataset, int64_t m, int leaf, int64_t* order_out, int64_t node_cap, int32_t* feat,
                  T* div_lo, T* div_hi, int64_t* first, int64_t* last, int32_t* kid0, int32_t* kid1) {
    const KdOracle<T> tree(dataset, m, leaf);
    for (int64_t i = 0; i < m; ++i) order_out[i] = tree.order[i];
    const int64_t nn = (int64_t)tree.nodes.size();
    for (int64_t i = 0; i < nn && i < node_cap; ++i) {
        feat[i] = tree.nodes[i].feat; div_lo[i] = tree.nodes[i].div_lo; div_hi[i] = tree.nodes[i].div_hi;
        first[i] = tree.nodes[i].first; last[i] = tree.nodes[i].last;
        kid0[i] = tree.nodes[i].kid[0]; kid1[i] = tree.nodes[i].kid[1];
    }
    return nn;
}

}  // namespace

extern "C" {

int pcu_oracle_knn_f32(const float* q, int64_t n, const float* d, int64_t m, int k, int squared, int leaf,
                       int num_threads, float* out_d, int64_t* out_i) {
    if (k <= 0 || n <= 0 || m <= 0 || leaf < 1) return 1;
    shortest_distances<float>(q, n, d, m, k, squared, leaf, num_threads, out_d, out_i);
    return 0;
}
int pcu_oracle_knn_f64(const double* q, int64_t n, const double* d, int64_t m, int k, int squared, int leaf,
                       int num_threads, double* out_d, int64_t* out_i) {
    if (k <= 0 || n <= 0 || m <= 0 || leaf < 1) return 1;
    shortest_distances<double>(q, n, d, m, k, squared, leaf, num_threads, out_d, out_i);
    return 0;
}
int pcu_oracle_one_sided_hausdorff_f32(const float* s, int64_t n, const float* t, int64_t m, int squared, int leaf,
                                       float* out_max, int64_t* out_i, int64_t* out_j) {
    if (n <= 0 || m <= 0 || leaf < 1) return 1;
    one_sided<float>(s, n, t, m, squared, leaf, out_max, out_i, out_j);
    return 0;
}
int pcu_oracle_one_sided_hausdorff_f64(const double* s, int64_t n, const double* t, int64_t m, int squared, int leaf,
                                       double* out_max, int64_t* out_i, int64_t* out_j) {
    if (n <= 0 || m <= 0 || leaf < 1) return 1;
    one_sided<double>(s, n, t, m, squared, leaf, out_max, out_i, out_j);
    return 0;
}
int64_t pcu_oracle_tree_f32(const float* d, int64_t m, int leaf, int64_t* order_out, int64_t node_cap, int32_t* feat,
                            float* div_lo, float* div_hi, int64_t* first, int64_t* last, int32_t* kid0, int32_t* kid1) {
    return dump_tree<float>(d, m, leaf, order_out, node_cap, feat, div_lo, div_hi, first, last, kid0, kid1);
}
int64_t pcu_oracle_tree_f64(const double* d, int64_t m, int leaf, int64_t* order_out, int64_t node_cap, int32_t* feat,
                            double* div_lo, double* div_hi, int64_t* first, int64_t* last, int32_t* kid0, int32_t* kid1) {
    return dump_tree<double>(d, m, leaf, order_out, node_cap, feat, div_lo, div_hi, first, last, kid0, kid1);
}
int pcu_oracle_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
