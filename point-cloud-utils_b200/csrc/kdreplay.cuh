// kdreplay.cuh -- GPU replica of the reference's kd-tree, used ONLY for queries whose answer
// depends on how exactly-equal distances are ordered.
//
// nanoflann breaks distance ties by kd-tree visit order (strict comparisons at
// external/nanoflann/nanoflann.hpp:206 and :1563 of the reference; NANOFLANN_FIRST_MATCH is not
// defined), so bit-identical indices for tied queries need the very tree the reference builds:
// the same split dimension / value at every node and the same permutation of points inside every
// leaf.  The grid search flags such queries (search.cuh); this file rebuilds that tree on the GPU
// and re-answers just those queries by walking it exactly like nanoflann does.
//
//   build   : level-synchronous restatement of divideTree / middleSplit_ / planeSplit
//             (nanoflann.hpp:1001-1059, :1061-1110, :1121-1162).  planeSplit's two-pointer exchange
//             is deterministic -- the j-th misplaced element from the left trades places with the
//             j-th misplaced element from the right -- so it is reproduced with prefix sums.
//   search  : findNeighbors / searchLevel / KNNResultSet (nanoflann.hpp:1394-1418, :1545-1624,
//             :157-230) with an explicit stack; one thread per flagged query.
//
// The build costs a few hundred small launches and is only run when at least one query is
// flagged (uniform random clouds: none for k = 1, a handful per million queries for k = 16).
#pragma once
#include <atomic>
#include <cfloat>
#include "common.cuh"
#include "grid.cuh"
#include "host_util.h"
#include "../../include/pcu_b200.h"

namespace pcu {

template <typename T>
struct KdNode {
    int feat;            // split dimension; -1 = leaf; -2 = not decided yet
    int first, last;     // slot range [first, last) in order[]
    int kid0, kid1;
    int parent;          // parent node (-1 for the root); side: 0 = left child, 1 = right child
    int side;
    int n_less, n_less_eq;
    T cut;
    T div_lo, div_hi;
    T loose_lo[3], loose_hi[3];
    typename Real<T>::bits_t tight_lo[3], tight_hi[3];   // order-preserving integer images (atomics)
};

struct KdCounters {
    int n_nodes;        // nodes allocated so far
    int level_begin;    // first node id of the current level
    int level_end;      // one past the last node id of the current level
    int n_split;        // nodes of the current level that were split
};

template <typename T>
struct KdReplayBuffers {
    long long capacity = 0;   // points
    int* order = nullptr;
    int* node_of = nullptr;
    unsigned* prefix = nullptr;       // capacity + 1
    unsigned* scan_partial = nullptr;
    int* left_pos = nullptr;
    int* right_pos = nullptr;
    KdNode<T>* nodes = nullptr;       // 2 * capacity
    KdCounters* counters = nullptr;
    long long* one_row = nullptr;     // scratch for the single-query (witness) replay
    T* one_dist = nullptr;
    long long* one_idx = nullptr;

    void carve(Carver& cv, long long points) {
        capacity = points;
        order = cv.take<int>((size_t)points);
        node_of = cv.take<int>((size_t)points);
        prefix = cv.take<unsigned>((size_t)points + 1);
        scan_partial = cv.take<unsigned>(((size_t)points + 1 + kScanTile - 1) / kScanTile + 1);
        left_pos = cv.take<int>((size_t)points);
        right_pos = cv.take<int>((size_t)points);
        nodes = cv.take<KdNode<T>>((size_t)2 * points + 2);
        counters = cv.take<KdCounters>(1);
        one_row = cv.take<long long>(1);
        one_dist = cv.take<T>(1);
        one_idx = cv.take<long long>(1);
    }
};

__global__ void widen_counter_kernel(const unsigned* src, long long* dst) { *dst = (long long)*src; }

// ---- build kernels -------------------------------------------------------------------------------
template <typename T>
__global__ void kd_init_kernel(KdReplayBuffers<T> b, int m) {
    using R = Real<T>;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < m) { b.order[s] = s; b.node_of[s] = 0; }
    if (s == 0) {
        KdNode<T> nd{};
        nd.feat = -2; nd.first = 0; nd.last = m; nd.kid0 = nd.kid1 = -1; nd.parent = -1; nd.side = 0;
        for (int d = 0; d < 3; ++d) { nd.tight_lo[d] = ordered<T>(R::inf()); nd.tight_hi[d] = ordered<T>(-R::inf()); }
        b.nodes[0] = nd;
        KdCounters c; c.n_nodes = 1; c.level_begin = 0; c.level_end = 1; c.n_split = 0;
        *b.counters = c;
    }
}

// tight bounding box of every node of the current level (element-parallel, atomics on the ordered
// integer image; a warp whose lanes all sit in the same node reduces first)
template <typename T>
__global__ void kd_tight_box_kernel(KdReplayBuffers<T> b, const T* __restrict__ pts, int m) {
    using bits_t = typename Real<T>::bits_t;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int node = s < m ? b.node_of[s] : -1;
    bits_t v[3] = {0, 0, 0};
    if (node >= 0) {
        const long long p = b.order[s];
        for (int d = 0; d < 3; ++d) v[d] = ordered<T>(pts[3 * p + d]);
    }
    const int lead = __shfl_sync(0xffffffffu, node, 0);
    const bool uniform = __all_sync(0xffffffffu, node == lead);
    if (uniform) {
        if (lead < 0) return;
        bits_t lo[3], hi[3];
        for (int d = 0; d < 3; ++d) {
            lo[d] = hi[d] = v[d];
            for (int o = 16; o > 0; o >>= 1) {
                const bits_t a = __shfl_xor_sync(0xffffffffu, lo[d], o), c = __shfl_xor_sync(0xffffffffu, hi[d], o);
                lo[d] = a < lo[d] ? a : lo[d];
                hi[d] = c > hi[d] ? c : hi[d];
            }
        }
        if ((threadIdx.x & 31) == 0)
            for (int d = 0; d < 3; ++d) { atomicMin(&b.nodes[lead].tight_lo[d], lo[d]); atomicMax(&b.nodes[lead].tight_hi[d], hi[d]); }
    } else if (node >= 0) {
        for (int d = 0; d < 3; ++d) { atomicMin(&b.nodes[node].tight_lo[d], v[d]); atomicMax(&b.nodes[node].tight_hi[d], v[d]); }
    }
}

// leaf-or-split decision + split plane of every node of the current level (node-parallel);
// also hands this node's tight extent along the parent's split axis up to the parent
// (divlow / divhigh, nanoflann.hpp:1047-1048).
template <typename T>
__global__ void kd_decide_kernel(KdReplayBuffers<T> b, int leaf_cap) {
    using R = Real<T>;
    const KdCounters c = *b.counters;
    for (int id = c.level_begin + blockIdx.x * blockDim.x + threadIdx.x; id < c.level_end; id += gridDim.x * blockDim.x) {
        KdNode<T>& nd = b.nodes[id];
        T tlo[3], thi[3];
        for (int d = 0; d < 3; ++d) { tlo[d] = unordered<T>(nd.tight_lo[d]); thi[d] = unordered<T>(nd.tight_hi[d]); }
        if (nd.parent >= 0) {
            KdNode<T>& par = b.nodes[nd.parent];
            if (nd.side == 0) par.div_lo = thi[par.feat]; else par.div_hi = tlo[par.feat];
        } else {
            for (int d = 0; d < 3; ++d) { nd.loose_lo[d] = tlo[d]; nd.loose_hi[d] = thi[d]; }  // root: computeBoundingBox
        }
        const int count = nd.last - nd.first;
        if (count <= leaf_cap) { nd.feat = -1; continue; }
        // middleSplit_ (nanoflann.hpp:1061-1096)
        const T eps = (T)0.00001;
        T widest = R::sub(nd.loose_hi[0], nd.loose_lo[0]);
        for (int d = 1; d < 3; ++d) {
            const T w = R::sub(nd.loose_hi[d], nd.loose_lo[d]);
            if (w > widest) widest = w;
        }
        const T gate = R::mul(R::sub((T)1, eps), widest);
        T best_spread = (T)-1;
        int feat = 0;
        for (int d = 0; d < 3; ++d) {
            const T w = R::sub(nd.loose_hi[d], nd.loose_lo[d]);
            if (w > gate) {
                const T spread = R::sub(thi[d], tlo[d]);
                if (spread > best_spread) { feat = d; best_spread = spread; }
            }
        }
        const T mid = R::mul(R::add(nd.loose_lo[feat], nd.loose_hi[feat]), (T)0.5);
        T cut;
        if (mid < tlo[feat]) cut = tlo[feat];
        else if (mid > thi[feat]) cut = thi[feat];
        else cut = mid;
        nd.feat = feat;
        nd.cut = cut;
    }
}

// sweep 1: flag = (value < cut); sweep 2: flag = (value <= cut) on the part right of n_less.
template <typename T, int kSweep>
__global__ void kd_flag_kernel(KdReplayBuffers<T> b, const T* __restrict__ pts, int m) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > m) return;
    unsigned f = 0;
    if (s < m) {
        const int node = b.node_of[s];
        if (node >= 0) {
            const KdNode<T>& nd = b.nodes[node];
            if (nd.feat >= 0) {
                const T v = pts[3 * (long long)b.order[s] + nd.feat];
                if (kSweep == 1) f = v < nd.cut ? 1u : 0u;
                else f = (s >= nd.first + nd.n_less && v <= nd.cut) ? 1u : 0u;
            }
        }
    }
    b.prefix[s] = f;   // entry m is the sentinel that becomes the grand total
}

// exclusive scan of prefix[0 .. m] (three phases, same scheme as grid.cuh)
template <typename T>
__global__ void __launch_bounds__(kScanThreads) kd_scan_reduce_kernel(KdReplayBuffers<T> b, int count) {
    const long long base = (long long)blockIdx.x * kScanTile;
    unsigned s = 0;
    for (int k = 0; k < kScanItems; ++k) {
        const long long i = base + (long long)k * kScanThreads + threadIdx.x;
        if (i < count) s += b.prefix[i];
    }
    unsigned total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) b.scan_partial[blockIdx.x] = total;
}
template <typename T>
__global__ void __launch_bounds__(kScanThreads) kd_scan_partials_kernel(KdReplayBuffers<T> b, int nb) {
    unsigned carry = 0;
    for (int base = 0; base < nb; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const unsigned v = i < nb ? b.scan_partial[i] : 0u;
        unsigned total;
        const unsigned ex = block_exclusive_scan(v, &total);
        if (i < nb) b.scan_partial[i] = carry + ex;
        carry += total;
    }
}
template <typename T>
__global__ void __launch_bounds__(kScanThreads) kd_scan_apply_kernel(KdReplayBuffers<T> b, int count) {
    const long long base = (long long)blockIdx.x * kScanTile;
    unsigned v[kScanItems];
    unsigned s = 0;
    const long long first = base + (long long)threadIdx.x * kScanItems;
    for (int k = 0; k < kScanItems; ++k) { v[k] = (first + k) < count ? b.prefix[first + k] : 0u; s += v[k]; }
    unsigned total;
    unsigned run = block_exclusive_scan(s, &total) + b.scan_partial[blockIdx.x];
    for (int k = 0; k < kScanItems; ++k) {
        if ((first + k) < count) b.prefix[first + k] = run;
        run += v[k];
    }
}

// Who trades places with whom (nanoflann.hpp:1125-1160): inside a node, with F = number of flagged
// slots of the swept range [lo, last), the flagged slots must end up in [lo, lo + F).  The j-th
// unflagged slot of [lo, lo + F) (ascending) exchanges with the j-th flagged slot of [lo + F, last)
// (descending).
template <typename T, int kSweep>
__global__ void kd_partner_kernel(KdReplayBuffers<T> b, int m) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m) return;
    const int node = b.node_of[s];
    if (node < 0) return;
    KdNode<T>& nd = b.nodes[node];
    if (nd.feat < 0) return;
    const int lo = kSweep == 1 ? nd.first : nd.first + nd.n_less;
    const unsigned at_lo = b.prefix[lo], at_last = b.prefix[nd.last];
    const int F = (int)(at_last - at_lo);
    if (s == nd.first) {
        if (kSweep == 1) nd.n_less = F; else nd.n_less_eq = nd.n_less + F;
    }
    if (s < lo) return;
    const bool flagged = b.prefix[s + 1] != b.prefix[s];
    const int r = s - lo;
    if (r < F && !flagged) {
        const int j = r - (int)(b.prefix[s] - at_lo);
        b.left_pos[lo + j] = s;
    } else if (r >= F && flagged) {
        const int j = (int)(at_last - b.prefix[s + 1]);
        b.right_pos[lo + j] = s;
    }
}

template <typename T, int kSweep>
__global__ void kd_exchange_kernel(KdReplayBuffers<T> b, int m) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m) return;
    const int node = b.node_of[s];
    if (node < 0) return;
    const KdNode<T>& nd = b.nodes[node];
    if (nd.feat < 0) return;
    const int lo = kSweep == 1 ? nd.first : nd.first + nd.n_less;
    if (s < lo) return;
    const int F = (int)(b.prefix[nd.last] - b.prefix[lo]);
    const int misplaced = F - (int)(b.prefix[lo + F] - b.prefix[lo]);   // unflagged slots inside [lo, lo + F)
    const int j = s - lo;
    if (j >= misplaced) return;
    const int a = b.left_pos[lo + j], c = b.right_pos[lo + j];
    const int t = b.order[a];
    b.order[a] = b.order[c];
    b.order[c] = t;
}

// children of every split node of the current level (nanoflann.hpp:1098-1110, :1033-1045)
template <typename T>
__global__ void kd_children_kernel(KdReplayBuffers<T> b) {
    using R = Real<T>;
    const KdCounters c = *b.counters;
    for (int id = c.level_begin + blockIdx.x * blockDim.x + threadIdx.x; id < c.level_end; id += gridDim.x * blockDim.x) {
        KdNode<T>& nd = b.nodes[id];
        if (nd.feat < 0) continue;
        const int count = nd.last - nd.first;
        int left;
        if (nd.n_less > count / 2) left = nd.n_less;
        else if (nd.n_less_eq < count / 2) left = nd.n_less_eq;
        else left = count / 2;
        const int k0 = atomicAdd(&b.counters->n_nodes, 2);
        atomicAdd(&b.counters->n_split, 1);
        nd.kid0 = k0; nd.kid1 = k0 + 1;
        for (int side = 0; side < 2; ++side) {
            KdNode<T> ch{};
            ch.feat = -2;
            ch.first = side == 0 ? nd.first : nd.first + left;
            ch.last = side == 0 ? nd.first + left : nd.last;
            ch.kid0 = ch.kid1 = -1; ch.parent = id; ch.side = side;
            for (int d = 0; d < 3; ++d) {
                ch.loose_lo[d] = nd.loose_lo[d]; ch.loose_hi[d] = nd.loose_hi[d];
                ch.tight_lo[d] = ordered<T>(R::inf()); ch.tight_hi[d] = ordered<T>(-R::inf());
            }
            if (side == 0) ch.loose_hi[nd.feat] = nd.cut; else ch.loose_lo[nd.feat] = nd.cut;
            b.nodes[k0 + side] = ch;
        }
    }
}

// slots move down to the child that now owns them; slots of leaves retire
template <typename T>
__global__ void kd_descend_kernel(KdReplayBuffers<T> b, int m) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m) return;
    const int node = b.node_of[s];
    if (node < 0) return;
    const KdNode<T>& nd = b.nodes[node];
    if (nd.feat < 0) { b.node_of[s] = -1; return; }
    b.node_of[s] = s < b.nodes[nd.kid0].last ? nd.kid0 : nd.kid1;
}

template <typename T>
__global__ void kd_next_level_kernel(KdReplayBuffers<T> b) {
    KdCounters c = *b.counters;
    c.level_begin = c.level_end;
    c.level_end = c.n_nodes;
    c.n_split = 0;
    *b.counters = c;
}

// ---- search ---------------------------------------------------------------------------------------
// nanoflann.hpp:157-230 with the list stored in the caller's output row (squared distances while
// searching).
template <typename T>
struct KdBest {
    T* d2; long long* id; int cap; int have;
    __device__ void init(T* d, long long* i, int k) {
        d2 = d; id = i; cap = k; have = 0;
        d2[cap - 1] = sizeof(T) == 4 ? (T)FLT_MAX : (T)DBL_MAX;
    }
    __device__ T worst() const { return d2[cap - 1]; }
    __device__ void offer(T dist, long long index) {
        int i = have;
        for (; i > 0; --i) {
            if (d2[i - 1] > dist) {
                if (i < cap) { d2[i] = d2[i - 1]; id[i] = id[i - 1]; }
            } else break;
        }
        if (i < cap) { d2[i] = dist; id[i] = index; }
        if (have < cap) ++have;
    }
};

constexpr int kKdStack = 96;

template <typename T>
__device__ void kd_search_one(const KdReplayBuffers<T>& b, const T* __restrict__ pts, const T q[3], int k, bool squared,
                              T* out_d, long long* out_i) {
    using R = Real<T>;
    KdBest<T> best;
    best.init(out_d, out_i, k);
    struct Frame { int node; T bound; T off[3]; bool far; };
    Frame stack[kKdStack];
    int top = 0;
    {   // computeInitialDistances (nanoflann.hpp:1164-1187)
        const KdNode<T>& root = b.nodes[0];
        Frame f; f.node = 0; f.bound = (T)0; f.far = false;
        for (int d = 0; d < 3; ++d) {
            f.off[d] = (T)0;
            const T lo = unordered<T>(root.tight_lo[d]), hi = unordered<T>(root.tight_hi[d]);
            if (q[d] < lo) { f.off[d] = sq_gap<T>(q[d], lo); f.bound = R::add(f.bound, f.off[d]); }
            if (q[d] > hi) { f.off[d] = sq_gap<T>(q[d], hi); f.bound = R::add(f.bound, f.off[d]); }
        }
        stack[top++] = f;
    }
    while (top > 0) {
        Frame f = stack[--top];
        // the far-side test is made when the near side has been searched completely (:1609)
        if (f.far && !(f.bound <= best.worst())) continue;
        for (;;) {
            const KdNode<T>& nd = b.nodes[f.node];
            if (nd.feat < 0) {
                const T worst_on_entry = best.worst();   // cached for the whole leaf (:1555)
                for (int s = nd.first; s < nd.last; ++s) {
                    const long long p = b.order[s];
                    const T d = dist2<T>(q[0], q[1], q[2], pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]);
                    if (d < worst_on_entry) best.offer(d, p);
                }
                break;
            }
            const int ft = nd.feat;
            const T v = q[ft];
            const T d1 = R::sub(v, nd.div_lo), d2 = R::sub(v, nd.div_hi);
            int near_kid, far_kid;
            T cut;
            if (R::add(d1, d2) < (T)0) { near_kid = nd.kid0; far_kid = nd.kid1; cut = sq_gap<T>(v, nd.div_hi); }
            else                       { near_kid = nd.kid1; far_kid = nd.kid0; cut = sq_gap<T>(v, nd.div_lo); }
            Frame g = f;
            g.node = far_kid;
            g.bound = R::sub(R::add(f.bound, cut), f.off[ft]);
            g.off[ft] = cut;
            g.far = true;
            if (top < kKdStack) stack[top++] = g;
            f.node = near_kid;
        }
    }
    for (int c = 0; c < best.have; ++c) if (!squared) out_d[c] = R::root(out_d[c]);
    for (int c = best.have; c < k; ++c) { out_d[c] = (T)-1; out_i[c] = -1; }
}

template <typename T>
__global__ void kd_replay_kernel(KdReplayBuffers<T> b, const T* __restrict__ query, const T* __restrict__ pts, int k,
                                 int squared, const long long* __restrict__ rows, const unsigned* __restrict__ n_rows,
                                 T* out_dist, long long* out_idx) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *n_rows) return;
    const long long row = rows[t];
    const T q[3] = {query[3 * row], query[3 * row + 1], query[3 * row + 2]};
    kd_search_one<T>(b, pts, q, k, squared != 0, out_dist + row * k, out_idx + row * k);
}

// single query taken from a stats record (the Hausdorff witness)
template <typename T>
__global__ void kd_witness_kernel(KdReplayBuffers<T> b, const T* __restrict__ query, const T* __restrict__ pts,
                                  pcu_b200_nn_stats* stats) {
    const long long row = stats->argmax_query;
    const T q[3] = {query[3 * row], query[3 * row + 1], query[3 * row + 2]};
    kd_search_one<T>(b, pts, q, 1, true, b.one_dist, b.one_idx);
    stats->argmax_data = *b.one_idx;
    stats->witness_tied = 0;
}

// ---- host-side sequencing ---------------------------------------------------------------------------
#define KD_LAUNCH(kernel, grid, block, stream, ...)                                   \
    do {                                                                              \
        kernel<<<grid, block, 0, stream>>>(__VA_ARGS__);                              \
        launches.fetch_add(1, std::memory_order_relaxed);                             \
        if (cudaGetLastError() != cudaSuccess) return PCU_B200_CUDA_ERROR;            \
    } while (0)

// Builds the replica of the reference's tree for `pts` (m points).  Synchronises once per level.
template <typename T>
int build_kd_replica(KdReplayBuffers<T>& b, const T* pts, long long m_ll, int leaf_cap, cudaStream_t stream,
                     std::atomic<long long>& launches) {
    if (m_ll > b.capacity || m_ll >= 0x7fffffffLL) return PCU_B200_INTERNAL;
    const int m = (int)m_ll;
    const unsigned eb = (unsigned)((m + 1 + kThreads - 1) / kThreads);
    const int scan_count = m + 1;
    const unsigned sb = (unsigned)((scan_count + kScanTile - 1) / kScanTile);
    KD_LAUNCH(kd_init_kernel<T>, eb, kThreads, stream, b, m);
    for (int level = 0; level < 4096; ++level) {
        KD_LAUNCH(kd_tight_box_kernel<T>, eb, kThreads, stream, b, pts, m);
        KD_LAUNCH(kd_decide_kernel<T>, 256, kThreads, stream, b, leaf_cap);
        // sweep 1: strictly less than the cut to the front
        KD_LAUNCH((kd_flag_kernel<T, 1>), eb, kThreads, stream, b, pts, m);
        KD_LAUNCH(kd_scan_reduce_kernel<T>, sb, kScanThreads, stream, b, scan_count);
        KD_LAUNCH(kd_scan_partials_kernel<T>, 1, kScanThreads, stream, b, (int)sb);
        KD_LAUNCH(kd_scan_apply_kernel<T>, sb, kScanThreads, stream, b, scan_count);
        KD_LAUNCH((kd_partner_kernel<T, 1>), eb, kThreads, stream, b, m);
        KD_LAUNCH((kd_exchange_kernel<T, 1>), eb, kThreads, stream, b, m);
        // sweep 2: equal to the cut next
        KD_LAUNCH((kd_flag_kernel<T, 2>), eb, kThreads, stream, b, pts, m);
        KD_LAUNCH(kd_scan_reduce_kernel<T>, sb, kScanThreads, stream, b, scan_count);
        KD_LAUNCH(kd_scan_partials_kernel<T>, 1, kScanThreads, stream, b, (int)sb);
        KD_LAUNCH(kd_scan_apply_kernel<T>, sb, kScanThreads, stream, b, scan_count);
        KD_LAUNCH((kd_partner_kernel<T, 2>), eb, kThreads, stream, b, m);
        KD_LAUNCH((kd_exchange_kernel<T, 2>), eb, kThreads, stream, b, m);
        KD_LAUNCH(kd_children_kernel<T>, 256, kThreads, stream, b);
        KD_LAUNCH(kd_descend_kernel<T>, eb, kThreads, stream, b, m);
        KdCounters h;
        if (cudaMemcpyAsync(&h, b.counters, sizeof h, cudaMemcpyDeviceToHost, stream) != cudaSuccess) return PCU_B200_CUDA_ERROR;
        if (cudaStreamSynchronize(stream) != cudaSuccess) return PCU_B200_CUDA_ERROR;
        if (h.n_split == 0) return PCU_B200_OK;
        KD_LAUNCH(kd_next_level_kernel<T>, 1, 1, stream, b);
    }
    return PCU_B200_INTERNAL;
}

// Re-answers the rows listed in tie_list with the reference's own tie order.  Reads the list length
// back (one synchronisation); does nothing more when it is zero.
template <typename T>
int enqueue_tie_replay(KdReplayBuffers<T>& b, const T* query, const T* dataset, long long m, int k, int squared,
                       int leaf_cap, const long long* tie_list, const unsigned* tie_count, T* out_dist,
                       long long* out_idx, cudaStream_t stream, std::atomic<long long>& launches) {
    unsigned h_count = 0;
    if (cudaMemcpyAsync(&h_count, tie_count, sizeof h_count, cudaMemcpyDeviceToHost, stream) != cudaSuccess) return PCU_B200_CUDA_ERROR;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return PCU_B200_CUDA_ERROR;
    if (h_count == 0) return PCU_B200_OK;
    const int st = build_kd_replica<T>(b, dataset, m, leaf_cap, stream, launches);
    if (st != PCU_B200_OK) return st;
    KD_LAUNCH(kd_replay_kernel<T>, (h_count + 127) / 128, 128, stream, b, query, dataset, k, squared, tie_list,
              tie_count, out_dist, out_idx);
    return PCU_B200_OK;
}

// Replays the single query stats->argmax_query (already known to be tie-dependent).
template <typename T>
int enqueue_witness_replay(KdReplayBuffers<T>& b, const T* query, const T* dataset, long long m, int leaf_cap,
                           pcu_b200_nn_stats* stats, cudaStream_t stream, std::atomic<long long>& launches) {
    const int st = build_kd_replica<T>(b, dataset, m, leaf_cap, stream, launches);
    if (st != PCU_B200_OK) return st;
    KD_LAUNCH(kd_witness_kernel<T>, 1, 1, stream, b, query, dataset, stats);
    return PCU_B200_OK;
}

}  // namespace pcu
