// search.cuh -- exact nearest-neighbour sweeps over the cell-sorted clouds.
//
// Replaces nanoflann's per-query kd-tree descent (external/nanoflann/nanoflann.hpp:1545-1624,
// KNNResultSet :157-230) and the reductions the reference performs afterwards on the CPU
// (dists.maxCoeff, src/point_cloud_distance.cpp:221-225; norm(...).mean(),
// point_cloud_utils/__init__.py:112-113).
//
// Exactness argument (DESIGN.md "why the grid search is exact"): a query examines whole cells; a
// cell that was NOT examined only contains points whose coordinate along some axis lies beyond a
// wall (grid.cuh), and because fl(q - p), fl(d*d) and fl(a + b) are all monotone, the squared gap
// to that wall, computed with the very same rounded operations, is a lower bound of the
// reference-rounded distance to every such point.  A query is final only when its current k-th
// distance is STRICTLY below that bound, so exact ties are never hidden either.
#pragma once
#include "common.cuh"
#include "grid.cuh"
#include "../../include/pcu_b200.h"

namespace pcu {

template <typename T>
struct SweepPartial {
    double sum, sumsq;
    T max_d2;
    long long arg_q, arg_d;
    unsigned n_tied;
    unsigned tie_at_max;
};

// One query-cloud -> dataset-cloud direction.
template <typename T>
struct Sweep {
    int qcloud, dcloud;
    int k;
    int squared;
    int main_blocks;             // partial slots [0, main_blocks) are written by the main kernel
    int pad;
    T* out_dist;                 // (n, k) or null
    long long* out_idx;          // (n, k) or null
    SweepPartial<T>* partial;    // main_blocks + far blocks slots, or null
    unsigned* far_list;          // sorted-order positions of queries the one-ring pass could not settle
    unsigned* counters;          // [0] far queries, [1] tied queries
    long long* tie_list;         // caller-order rows whose answer depends on tie order
    pcu_b200_nn_stats* stats;    // or null
};

template <typename T>
struct Best1 {
    T d;
    typename Real<T>::index_t i;
    bool tie;
};

template <typename T> __device__ __forceinline__ typename Real<T>::index_t no_index();
template <> __device__ __forceinline__ int32_t no_index<float>() { return 0x7fffffff; }
template <> __device__ __forceinline__ long long no_index<double>() { return 0x7fffffffffffffffLL; }

template <typename T>
__device__ __forceinline__ void offer1(Best1<T>& b, T d, typename Real<T>::index_t i) {
    if (d <= b.d) {
        if (d == b.d) { b.tie = true; b.i = i < b.i ? i : b.i; }
        else { b.d = d; b.i = i; b.tie = false; }
    }
}

template <typename T>
__device__ __forceinline__ void scan_run1(const Pt<T>* __restrict__ pts, unsigned a, unsigned b, T qx, T qy, T qz,
                                          Best1<T>& best) {
    for (unsigned j = a; j < b; ++j) {
        const Pt<T> p = load_pt<T>(pts + j);
        offer1<T>(best, dist2<T>(qx, qy, qz, p.x, p.y, p.z), p.i);
    }
}

// Visits the cells of the dataset grid around (qx,qy,qz) ring by ring (Chebyshev distance in cell
// units), starting at ring r0.  `visit(a, b, bound)` receives a contiguous run [a, b) of sorted
// points and a lower bound of the distance to all of them; `settled(lb)` is asked after each ring
// with the lower bound for everything not examined so far.
template <typename T, typename Visit, typename Settled>
__device__ __forceinline__ void expand_rings(const GridHeader<T>& g, const T* __restrict__ wall_lo,
                                             const T* __restrict__ wall_hi, const unsigned* __restrict__ cell_start,
                                             T qx, T qy, T qz, int r0, Visit&& visit, Settled&& settled) {
    using R = Real<T>;
    const int cx = cell_of<T>(qx, g.origin[0], g.inv_h, g.dim[0]);
    const int cy = cell_of<T>(qy, g.origin[1], g.inv_h, g.dim[1]);
    const int cz = cell_of<T>(qz, g.origin[2], g.inv_h, g.dim[2]);
    const int st = g.stride;
    const T* lo_x = wall_lo;           const T* hi_x = wall_hi;
    const T* lo_y = wall_lo + st;      const T* hi_y = wall_hi + st;
    const T* lo_z = wall_lo + 2 * st;  const T* hi_z = wall_hi + 2 * st;
    for (int r = r0;; ++r) {
        const int xa = max(cx - r, 0), xb = min(cx + r, g.dim[0] - 1);
        const int ya = max(cy - r, 0), yb = min(cy + r, g.dim[1] - 1);
        const int za = max(cz - r, 0), zb = min(cz + r, g.dim[2] - 1);
        for (int z = za; z <= zb; ++z) {
            const T bz = z < cz ? sq_gap<T>(qz, lo_z[z + 1]) : (z > cz ? sq_gap<T>(qz, hi_z[z]) : (T)0);
            for (int y = ya; y <= yb; ++y) {
                const T by = y < cy ? sq_gap<T>(qy, lo_y[y + 1]) : (y > cy ? sq_gap<T>(qy, hi_y[y]) : (T)0);
                const unsigned row = (unsigned)((z * g.dim[1] + y) * g.dim[0]);
                const bool shell_row = (z == cz - r) || (z == cz + r) || (y == cy - r) || (y == cy + r);
                if (shell_row || r == 0) {
                    visit(cell_start[row + xa], cell_start[row + xb + 1], R::add(by, bz));
                } else {
                    if (cx - r >= 0) {
                        const T bx = sq_gap<T>(qx, lo_x[cx - r + 1]);
                        visit(cell_start[row + cx - r], cell_start[row + cx - r + 1], R::add(R::add(bx, by), bz));
                    }
                    if (cx + r <= g.dim[0] - 1) {
                        const T bx = sq_gap<T>(qx, hi_x[cx + r]);
                        visit(cell_start[row + cx + r], cell_start[row + cx + r + 1], R::add(R::add(bx, by), bz));
                    }
                }
            }
        }
        T lb = sq_gap<T>(qx, lo_x[xa]);
        lb = R::vmin(lb, sq_gap<T>(qx, hi_x[xb + 1]));
        lb = R::vmin(lb, sq_gap<T>(qy, lo_y[ya]));
        lb = R::vmin(lb, sq_gap<T>(qy, hi_y[yb + 1]));
        lb = R::vmin(lb, sq_gap<T>(qz, lo_z[za]));
        lb = R::vmin(lb, sq_gap<T>(qz, hi_z[zb + 1]));
        if (settled(lb)) return;
        if (xa == 0 && ya == 0 && za == 0 && xb == g.dim[0] - 1 && yb == g.dim[1] - 1 && zb == g.dim[2] - 1) return;
    }
}

// ---------------------------------------------------------------------------------------------
// block-level reduction of the fused statistics
template <typename T>
struct MaxCand {
    T d2;
    long long q, d;
    unsigned tie;
};
template <typename T>
__device__ __forceinline__ void take_max(MaxCand<T>& a, const MaxCand<T>& b) {
    // larger distance wins; equal distance -> lower query row (Eigen maxCoeff: first maximum)
    if (b.d2 > a.d2 || (b.d2 == a.d2 && b.q < a.q)) a = b;
}

template <typename T>
__device__ __forceinline__ void block_reduce_stats(double sum, double sumsq, MaxCand<T> mc, unsigned ties,
                                                   SweepPartial<T>* out) {
    __shared__ double s_sum[kThreads / 32], s_sq[kThreads / 32];
    __shared__ MaxCand<T> s_mc[kThreads / 32];
    __shared__ unsigned s_t[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
        ties += __shfl_xor_sync(0xffffffffu, ties, o);
        MaxCand<T> other;
        other.d2 = __shfl_xor_sync(0xffffffffu, mc.d2, o);
        other.q = __shfl_xor_sync(0xffffffffu, mc.q, o);
        other.d = __shfl_xor_sync(0xffffffffu, mc.d, o);
        other.tie = __shfl_xor_sync(0xffffffffu, mc.tie, o);
        take_max<T>(mc, other);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s_sum[w] = sum; s_sq[w] = sumsq; s_mc[w] = mc; s_t[w] = ties; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < kThreads / 32; ++i) {   // fixed order: deterministic
            sum += s_sum[i]; sumsq += s_sq[i]; ties += s_t[i];
            take_max<T>(mc, s_mc[i]);
        }
        SweepPartial<T> p;
        p.sum = sum; p.sumsq = sumsq; p.max_d2 = mc.d2; p.arg_q = mc.q; p.arg_d = mc.d; p.n_tied = ties;
        p.tie_at_max = mc.tie;
        *out = p;
    }
}

template <typename T, bool kOut, bool kStats>
__device__ __forceinline__ void finish_query1(const Sweep<T>& sw, bool have, const Best1<T>& best, long long row,
                                              double& sum, double& sumsq, MaxCand<T>& mc, unsigned& ties) {
    using R = Real<T>;
    if (!have) return;
    const bool found = best.i != no_index<T>();
    const long long di = found ? (long long)best.i : -1;
    const T root = R::root(best.d);
    if (kOut) {
        sw.out_idx[row] = di;
        sw.out_dist[row] = found ? (sw.squared ? best.d : root) : (T)-1;
        if (best.tie) sw.tie_list[atomicAdd(sw.counters + 1, 1u)] = row;
    }
    if (kStats) {
        sum += (double)root;
        sumsq += (double)best.d;
        MaxCand<T> c; c.d2 = best.d; c.q = row; c.d = di; c.tie = best.tie ? 1u : 0u;
        take_max<T>(mc, c);
        ties += best.tie ? 1u : 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// k = 1 main pass: one thread per (cell-sorted) query, 3x3x3 neighbourhood, rows pruned by their
// wall bounds.  grid (ceil(max_n / kThreads), nsweeps).
template <typename T, bool kOut, bool kStats>
__global__ void __launch_bounds__(kThreads) nn1_kernel(const Cloud<T>* __restrict__ clouds,
                                                       const Sweep<T>* __restrict__ sweeps) {
    using R = Real<T>;
    const Sweep<T> sw = sweeps[blockIdx.y];
    const Cloud<T> qc = clouds[sw.qcloud];
    const Cloud<T> dc = clouds[sw.dcloud];
    if ((long long)blockIdx.x * blockDim.x >= qc.n) return;   // blocks beyond this sweep's queries
    __shared__ GridHeader<T> g;
    if (threadIdx.x == 0) g = *dc.grid;
    __syncthreads();

    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < qc.n;
    Best1<T> best; best.d = R::inf(); best.i = no_index<T>(); best.tie = false;
    bool settled = false;
    long long row = -1;
    if (active) {
        const Pt<T> q = load_pt<T>(qc.sorted + t);
        row = (long long)q.i;
        const int st = g.stride;
        const T* lo_x = dc.wall_lo;           const T* hi_x = dc.wall_hi;
        const T* lo_y = dc.wall_lo + st;      const T* hi_y = dc.wall_hi + st;
        const T* lo_z = dc.wall_lo + 2 * st;  const T* hi_z = dc.wall_hi + 2 * st;
        const int cx = cell_of<T>(q.x, g.origin[0], g.inv_h, g.dim[0]);
        const int cy = cell_of<T>(q.y, g.origin[1], g.inv_h, g.dim[1]);
        const int cz = cell_of<T>(q.z, g.origin[2], g.inv_h, g.dim[2]);
        const int xa = max(cx - 1, 0), xb = min(cx + 1, g.dim[0] - 1);
        // gaps to the walls of the query's own cell along y and z (row pruning)
        const T gy[3] = {(T)0, sq_gap<T>(q.y, __ldg(lo_y + cy)), sq_gap<T>(q.y, __ldg(hi_y + cy + 1))};
        const T gz[3] = {(T)0, sq_gap<T>(q.z, __ldg(lo_z + cz)), sq_gap<T>(q.z, __ldg(hi_z + cz + 1))};
        // (dy, dz) as indices into {0: same, 1: minus one, 2: plus one}; nearest rows first
        const int order_y[9] = {0, 1, 2, 0, 0, 1, 2, 1, 2};
        const int order_z[9] = {0, 0, 0, 1, 2, 1, 1, 2, 2};
#pragma unroll
        for (int s = 0; s < 9; ++s) {
            const int oy = order_y[s], oz = order_z[s];
            const int y = cy + (oy == 1 ? -1 : (oy == 2 ? 1 : 0));
            const int z = cz + (oz == 1 ? -1 : (oz == 2 ? 1 : 0));
            if (y < 0 || y >= g.dim[1] || z < 0 || z >= g.dim[2]) continue;
            const T bound = R::add(gy[oy], gz[oz]);
            if (bound > best.d) continue;   // every point of the row is strictly farther
            const unsigned base = (unsigned)((z * g.dim[1] + y) * g.dim[0]);
            const unsigned a = __ldg(dc.cell_start + base + xa), b = __ldg(dc.cell_start + base + xb + 1);
            scan_run1<T>(dc.sorted, a, b, q.x, q.y, q.z, best);
        }
        const int ya = max(cy - 1, 0), yb = min(cy + 1, g.dim[1] - 1);
        const int za = max(cz - 1, 0), zb = min(cz + 1, g.dim[2] - 1);
        T lb = sq_gap<T>(q.x, __ldg(lo_x + xa));
        lb = R::vmin(lb, sq_gap<T>(q.x, __ldg(hi_x + xb + 1)));
        lb = R::vmin(lb, sq_gap<T>(q.y, __ldg(lo_y + ya)));
        lb = R::vmin(lb, sq_gap<T>(q.y, __ldg(hi_y + yb + 1)));
        lb = R::vmin(lb, sq_gap<T>(q.z, __ldg(lo_z + za)));
        lb = R::vmin(lb, sq_gap<T>(q.z, __ldg(hi_z + zb + 1)));
        settled = best.d < lb;
        if (!settled) sw.far_list[atomicAdd(sw.counters, 1u)] = (unsigned)t;
    }
    double sum = 0.0, sumsq = 0.0;
    unsigned ties = 0;
    MaxCand<T> mc; mc.d2 = (T)-1; mc.q = 0x7fffffffffffffffLL; mc.d = -1; mc.tie = 0;
    finish_query1<T, kOut, kStats>(sw, active && settled, best, row, sum, sumsq, mc, ties);
    if (kStats) block_reduce_stats<T>(sum, sumsq, mc, ties, sw.partial + blockIdx.x);
}

// k = 1 slow pass for the queries the one-ring pass could not settle (empty neighbourhoods,
// queries outside the dataset's box): ring expansion until the wall bound closes.
// grid (ceil(max_n / kThreads), nsweeps); blocks beyond the far count exit immediately.
template <typename T, bool kOut, bool kStats>
__global__ void __launch_bounds__(kThreads) nn1_far_kernel(const Cloud<T>* __restrict__ clouds,
                                                           const Sweep<T>* __restrict__ sweeps) {
    using R = Real<T>;
    const Sweep<T> sw = sweeps[blockIdx.y];
    const unsigned n_far = sw.counters[0];
    if ((unsigned long long)blockIdx.x * blockDim.x >= n_far) return;
    const Cloud<T> qc = clouds[sw.qcloud];
    const Cloud<T> dc = clouds[sw.dcloud];
    __shared__ GridHeader<T> g;
    if (threadIdx.x == 0) g = *dc.grid;
    __syncthreads();
    const unsigned f = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = f < n_far;
    Best1<T> best; best.d = R::inf(); best.i = no_index<T>(); best.tie = false;
    long long row = -1;
    if (active) {
        const Pt<T> q = load_pt<T>(qc.sorted + sw.far_list[f]);
        row = (long long)q.i;
        expand_rings<T>(g, dc.wall_lo, dc.wall_hi, dc.cell_start, q.x, q.y, q.z, 0,
            [&](unsigned a, unsigned b, T bound) {
                if (bound > best.d) return;
                scan_run1<T>(dc.sorted, a, b, q.x, q.y, q.z, best);
            },
            [&](T lb) { return best.d < lb; });
    }
    double sum = 0.0, sumsq = 0.0;
    unsigned ties = 0;
    MaxCand<T> mc; mc.d2 = (T)-1; mc.q = 0x7fffffffffffffffLL; mc.d = -1; mc.tie = 0;
    finish_query1<T, kOut, kStats>(sw, active, best, row, sum, sumsq, mc, ties);
    if (kStats) block_reduce_stats<T>(sum, sumsq, mc, ties, sw.partial + sw.main_blocks + blockIdx.x);
}

// Combines the per-block partials of one sweep into its pcu_b200_nn_stats.  grid (1, nsweeps).
template <typename T>
__global__ void __launch_bounds__(kThreads) stats_finalize_kernel(const Cloud<T>* __restrict__ clouds,
                                                                  const Sweep<T>* __restrict__ sweeps) {
    const Sweep<T> sw = sweeps[blockIdx.y];
    const long long n = clouds[sw.qcloud].n;
    const unsigned n_far = sw.counters[0];
    const int main_used = (int)((n + kThreads - 1) / kThreads);
    const int far_used = (int)((n_far + kThreads - 1) / kThreads);
    double sum = 0.0, sumsq = 0.0;
    unsigned ties = 0;
    MaxCand<T> mc; mc.d2 = (T)-1; mc.q = 0x7fffffffffffffffLL; mc.d = -1; mc.tie = 0;
    const int total = main_used + far_used;
    for (int s = threadIdx.x; s < total; s += blockDim.x) {
        const SweepPartial<T> p = sw.partial[s < main_used ? s : sw.main_blocks + (s - main_used)];
        sum += p.sum; sumsq += p.sumsq; ties += p.n_tied;
        MaxCand<T> c; c.d2 = p.max_d2; c.q = p.arg_q; c.d = p.arg_d; c.tie = p.tie_at_max;
        take_max<T>(mc, c);
    }
    __shared__ SweepPartial<T> result;
    block_reduce_stats<T>(sum, sumsq, mc, ties, &result);
    __syncthreads();
    if (threadIdx.x == 0) {
        pcu_b200_nn_stats s;
        s.sum_dist = result.sum;
        s.sum_sq_dist = result.sumsq;
        s.max_sq_dist = (double)result.max_d2;
        s.argmax_query = result.arg_q;
        s.argmax_data = result.arg_d;
        s.n_queries = n;
        s.n_tied = result.n_tied;
        s.n_far = n_far;
        s.witness_tied = result.tie_at_max ? 1 : 0;
        *sw.stats = s;
    }
}

// chamfer = mean_x |x - NN_y(x)| + mean_y |y - NN_x(y)|  (point_cloud_utils/__init__.py:112-115)
// stats: 2 per pair ([2p] = x->y, [2p+1] = y->x).  One block; pairs strided over its threads.
template <typename T>
__global__ void __launch_bounds__(kThreads) chamfer_value_kernel(const pcu_b200_nn_stats* __restrict__ stats,
                                                                 long long npairs, T* __restrict__ out_value,
                                                                 double* __restrict__ out_sum) {
    double acc = 0.0;
    for (long long p = threadIdx.x; p < npairs; p += blockDim.x) {
        const pcu_b200_nn_stats a = stats[2 * p], b = stats[2 * p + 1];
        const double v = a.sum_dist / (double)a.n_queries + b.sum_dist / (double)b.n_queries;
        const T vt = (T)v;
        if (out_value) out_value[p] = vt;
        acc += (double)vt;
    }
    if (out_sum == nullptr) return;
    __shared__ double s[kThreads];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_sum = s[0];
}

// ---------------------------------------------------------------------------------------------
// 2 <= k <= 32: one warp per query; lane j holds the j-th best (distance, index) pair, candidates
// are evaluated 32 at a time and inserted with shuffles.  Order inside the list is (distance, index)
// ascending, which is deterministic; queries whose answer depends on how the reference orders equal
// distances are reported in tie_list and re-answered by the kd-tree replay.
// grid (ceil(max_n * 32 / kThreads), nsweeps).
template <typename T>
__global__ void __launch_bounds__(kThreads) knn_warp_kernel(const Cloud<T>* __restrict__ clouds,
                                                            const Sweep<T>* __restrict__ sweeps) {
    using R = Real<T>;
    using index_t = typename R::index_t;
    const Sweep<T> sw = sweeps[blockIdx.y];
    const Cloud<T> qc = clouds[sw.qcloud];
    const Cloud<T> dc = clouds[sw.dcloud];
    const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= qc.n) return;   // warp-uniform
    const int lane = threadIdx.x & 31;
    const int k = sw.k;
    const GridHeader<T> g = *dc.grid;
    const Pt<T> q = load_pt<T>(qc.sorted + t);

    T dl = R::inf();
    index_t il = no_index<T>();
    T worst = R::inf();
    index_t worst_i = no_index<T>();
    T rej = R::inf();   // smallest distance that was turned away or pushed out (uniform across lanes)
    const unsigned kmask = k >= 32 ? 0xffffffffu : ((1u << k) - 1u);

    auto visit = [&](unsigned a, unsigned b, T bound) {
        if (bound > worst) return;
        for (unsigned base = a; base < b; base += 32) {
            const unsigned j = base + lane;
            const bool valid = j < b;
            T d = R::inf();
            index_t pi = no_index<T>();
            if (valid) {
                const Pt<T> p = load_pt<T>(dc.sorted + j);
                d = dist2<T>(q.x, q.y, q.z, p.x, p.y, p.z);
                pi = p.i;
            }
            const bool pass = valid && (d < worst || (d == worst && pi < worst_i));
            // distances that never enter the list still matter for the tie flag
            T turned = (valid && !pass) ? d : R::inf();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) turned = R::vmin(turned, __shfl_xor_sync(0xffffffffu, turned, o));
            rej = R::vmin(rej, turned);
            unsigned mask = __ballot_sync(0xffffffffu, pass);
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const T cd = __shfl_sync(0xffffffffu, d, src);
                const index_t ci = __shfl_sync(0xffffffffu, pi, src);
                if (!(cd < worst || (cd == worst && ci < worst_i))) { rej = R::vmin(rej, cd); continue; }
                const bool before = (dl < cd) || (dl == cd && il < ci);
                const int pos = __popc(__ballot_sync(0xffffffffu, before) & kmask);
                const T pushed = __shfl_sync(0xffffffffu, dl, k - 1);
                const T up_d = __shfl_up_sync(0xffffffffu, dl, 1);
                const index_t up_i = __shfl_up_sync(0xffffffffu, il, 1);
                if (lane < k) {
                    if (lane > pos) { dl = up_d; il = up_i; }
                    else if (lane == pos) { dl = cd; il = ci; }
                }
                rej = R::vmin(rej, pushed);
                worst = __shfl_sync(0xffffffffu, dl, k - 1);
                worst_i = __shfl_sync(0xffffffffu, il, k - 1);
            }
        }
    };
    expand_rings<T>(g, dc.wall_lo, dc.wall_hi, dc.cell_start, q.x, q.y, q.z, 0, visit,
                    [&](T lb) { return worst < lb; });

    const long long row = (long long)q.i;
    if (lane < k) {
        const bool found = il != no_index<T>();
        sw.out_idx[row * k + lane] = found ? (long long)il : -1;
        sw.out_dist[row * k + lane] = found ? (sw.squared ? dl : R::root(dl)) : (T)-1;
    }
    const T next_d = __shfl_down_sync(0xffffffffu, dl, 1);
    const bool dup = lane < k - 1 && dl == next_d && il != no_index<T>();
    const bool edge = (rej == worst) && (worst_i != no_index<T>());
    const unsigned any = __ballot_sync(0xffffffffu, dup || edge);
    if (any && lane == 0) sw.tie_list[atomicAdd(sw.counters + 1, 1u)] = row;
}

// k > 32: one thread per query, the (distance, index)-sorted list lives in the caller's output rows
// (squared distances while searching).  Generic and slow; large k is not a hot configuration.
// grid (ceil(max_n / kThreads), nsweeps).
template <typename T>
__global__ void __launch_bounds__(kThreads) knn_big_kernel(const Cloud<T>* __restrict__ clouds,
                                                           const Sweep<T>* __restrict__ sweeps) {
    using R = Real<T>;
    const Sweep<T> sw = sweeps[blockIdx.y];
    const Cloud<T> qc = clouds[sw.qcloud];
    const Cloud<T> dc = clouds[sw.dcloud];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= qc.n) return;
    const GridHeader<T> g = *dc.grid;
    const Pt<T> q = load_pt<T>(qc.sorted + t);
    const int k = sw.k;
    const long long row = (long long)q.i;
    T* ld = sw.out_dist + row * k;
    long long* li = sw.out_idx + row * k;
    int have = 0;
    T worst = R::inf();
    long long worst_i = 0x7fffffffffffffffLL;
    T rej = R::inf();
    auto visit = [&](unsigned a, unsigned b, T bound) {
        if (bound > worst) return;
        for (unsigned j = a; j < b; ++j) {
            const Pt<T> p = load_pt<T>(dc.sorted + j);
            const T d = dist2<T>(q.x, q.y, q.z, p.x, p.y, p.z);
            const long long pi = (long long)p.i;
            if (!(d < worst || (d == worst && pi < worst_i))) { rej = R::vmin(rej, d); continue; }
            if (have == k) rej = R::vmin(rej, ld[k - 1]);
            int s = have < k ? have : k - 1;
            while (s > 0 && (ld[s - 1] > d || (ld[s - 1] == d && li[s - 1] > pi))) {
                ld[s] = ld[s - 1]; li[s] = li[s - 1]; --s;
            }
            ld[s] = d; li[s] = pi;
            if (have < k) ++have;
            if (have == k) { worst = ld[k - 1]; worst_i = li[k - 1]; }
        }
    };
    expand_rings<T>(g, dc.wall_lo, dc.wall_hi, dc.cell_start, q.x, q.y, q.z, 0, visit,
                    [&](T lb) { return worst < lb; });
    bool tie = have == k && rej == worst;
    for (int s = 0; s + 1 < have; ++s) tie = tie || (ld[s] == ld[s + 1]);
    if (!sw.squared) for (int s = 0; s < have; ++s) ld[s] = R::root(ld[s]);
    for (int s = have; s < k; ++s) { ld[s] = (T)-1; li[s] = -1; }
    if (tie) sw.tie_list[atomicAdd(sw.counters + 1, 1u)] = row;
}

}  // namespace pcu
