// topk.cuh -- k > 1 nearest-neighbour sweeps (warp top-k for k <= 32, generic list for larger k).
//
// Replaces nanoflann's KNNResultSet-driven search for k > 1 (external/nanoflann/nanoflann.hpp:157-230,
// :1545-1624 in the reference).  Shares the ring walk and its exactness argument with search.cuh.
#pragma once
#include "search.cuh"

namespace pcu {

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
