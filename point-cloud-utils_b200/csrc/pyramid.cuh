// pyramid.cuh -- robust slow path for queries that are far from every dataset point.
//
// The ring walk (search.cuh / nn1.cuh) visits every cell of a ring whether it holds points or not, so
// a query at distance D from the data costs O((D/h)^3) cell visits -- fine for the stray query next to
// an empty patch, hopeless when the two clouds do not overlap at all (every query is "far").  Queries
// still unsettled after a few rings are therefore answered by a best-first-ish descent of an occupancy
// pyramid over the dataset grid (level l halves the resolution l times; a node stores how many points
// lie below it): empty space is skipped in O(1) per node and the search cost becomes logarithmic in
// the grid size.  Exactness is the same wall argument as everywhere else: a node is skipped only when
// the lower bound of the reference-rounded distance to its cell box is STRICTLY above the current best.
//
// The pyramid is built by the last CTA of the far pass (nn1.cuh), and only when that sweep has such
// queries: uniform random clouds never pay for it.
#pragma once
#include "search.cuh"
#include "nn1.cuh"

namespace pcu {

// Fills the levels of the dataset's occupancy pyramid; executed by ONE CTA (the far pass's last one).
template <typename T>
__device__ void build_pyramid(const Cloud<T>& dc) {
    __shared__ PyramidShape ps;
    for (int w = threadIdx.x; w < (int)(sizeof(PyramidShape) / sizeof(int)); w += blockDim.x)
        reinterpret_cast<int*>(&ps)[w] = reinterpret_cast<const int*>(dc.shape)[w];
    __syncthreads();
    for (int l = 1; l <= ps.levels; ++l) {
        const int nx = ps.lvl_dim[l][0], ny = ps.lvl_dim[l][1], nz = ps.lvl_dim[l][2];
        const int cx = ps.lvl_dim[l - 1][0], cy = ps.lvl_dim[l - 1][1], cz = ps.lvl_dim[l - 1][2];
        unsigned* out = dc.pyramid + ps.lvl_off[l];
        const unsigned* below = dc.pyramid + ps.lvl_off[l - 1];   // unused for l == 1
        for (int node = threadIdx.x; node < nx * ny * nz; node += blockDim.x) {
            const int x = node % nx, y = (node / nx) % ny, z = node / (nx * ny);
            unsigned total = 0;
            for (int dz = 0; dz < 2; ++dz)
                for (int dy = 0; dy < 2; ++dy) {
                    const int yy = 2 * y + dy, zz = 2 * z + dz;
                    if (yy >= cy || zz >= cz) continue;
                    const int row = (zz * cy + yy) * cx;
                    const int x0 = 2 * x, x1 = min(2 * x + 2, cx);
                    if (l == 1) total += dc.cell_start[row + x1] - dc.cell_start[row + x0];   // x-adjacent cells are contiguous
                    else for (int xx = x0; xx < x1; ++xx) total += __ldcg(below + row + xx);
                }
            out[node] = total;
        }
        __threadfence();
        __syncthreads();
    }
}

// One thread per very-far query: depth-first descent, nearer children first.
// grid (sw.far_blocks, nsweeps), thread-stride loop over the very-far list.
template <typename T, bool kOut, bool kStats>
__global__ void __launch_bounds__(kThreads) nn1_vfar_kernel(const Cloud<T>* __restrict__ clouds,
                                                            const Sweep<T>* __restrict__ sweeps) {
    using R = Real<T>;
    const Sweep<T> sw = sweeps[blockIdx.y];
    const unsigned n_vfar = sw.counters[2];
    const Cloud<T> qc = clouds[sw.qcloud];
    const Cloud<T> dc = clouds[sw.dcloud];
    double sum = 0.0, sumsq = 0.0;
    unsigned ties = 0;
    MaxCand<T> mc; mc.d2 = (T)-1; mc.q = 0x7fffffffffffffffLL; mc.d = -1; mc.tie = 0;
    if (n_vfar > 0) {
        __shared__ GridHeader<T> g;
        __shared__ PyramidShape ps;
        if (threadIdx.x == 0) g = *dc.grid;
        for (int w = threadIdx.x; w < (int)(sizeof(PyramidShape) / sizeof(int)); w += blockDim.x)
            reinterpret_cast<int*>(&ps)[w] = reinterpret_cast<const int*>(dc.shape)[w];
        __syncthreads();
        const int st = g.stride;
        const T* lo[3] = {dc.wall_lo, dc.wall_lo + st, dc.wall_lo + 2 * st};
        const T* hi[3] = {dc.wall_hi, dc.wall_hi + st, dc.wall_hi + 2 * st};
        const unsigned step = gridDim.x * blockDim.x;
        for (unsigned f = blockIdx.x * blockDim.x + threadIdx.x; f < n_vfar; f += step) {
            const Pt<T> q = load_pt<T>(qc.sorted + sw.vfar_list[f]);
            const T qv[3] = {q.x, q.y, q.z};
            int qc3[3];
            for (int a = 0; a < 3; ++a) qc3[a] = cell_of<T>(qv[a], g.origin[a], g.inv_h, g.dim[a]);
            Best1<T> best; best.d = R::inf(); best.i = no_index<T>(); best.tie = false;
            // node = level (4 bits) | x (12) | y (12) | z (12)
            unsigned long long stack[8 * kMaxLevels + 8];
            int top = 0;
            stack[top++] = (unsigned long long)ps.levels << 36;
            while (top > 0) {
                const unsigned long long nd = stack[--top];
                const int l = (int)(nd >> 36), x = (int)((nd >> 24) & 0xfff), y = (int)((nd >> 12) & 0xfff), z = (int)(nd & 0xfff);
                const int c0[3] = {x << l, y << l, z << l};
                T bound = (T)0;
                int pref = 0;   // bit a set: the query lies towards the upper half of the node along axis a
                {
                    T gap[3];
                    for (int a = 0; a < 3; ++a) {
                        const int c1 = min(((c0[a] >> l) + 1 << l) - 1, g.dim[a] - 1);
                        gap[a] = qc3[a] < c0[a] ? sq_gap<T>(qv[a], hi[a][c0[a]])
                                                : (qc3[a] > c1 ? sq_gap<T>(qv[a], lo[a][c1 + 1]) : (T)0);
                        if (l > 0 && qc3[a] >= c0[a] + (1 << (l - 1))) pref |= 1 << a;
                    }
                    bound = R::add(R::add(gap[0], gap[1]), gap[2]);
                }
                if (bound > best.d) continue;   // everything below is strictly farther
                if (l == 0) {
                    const unsigned lin = (unsigned)((z * g.dim[1] + y) * g.dim[0] + x);
                    scan_run1<T>(dc.sorted, dc.cell_start[lin], dc.cell_start[lin + 1], q.x, q.y, q.z, best);
                    continue;
                }
                const int cl = l - 1;
                const int nx = ps.lvl_dim[cl][0], ny = ps.lvl_dim[cl][1], nz = ps.lvl_dim[cl][2];
                const unsigned* lvl = dc.pyramid + ps.lvl_off[cl];
                // children in order of increasing Hamming distance from the preferred octant; pushed in
                // reverse so that the nearest is popped first
                const int order[8] = {7, 6, 5, 3, 4, 2, 1, 0};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int ch = order[j] ^ pref;
                    const int xx = 2 * x + (ch & 1), yy = 2 * y + ((ch >> 1) & 1), zz = 2 * z + ((ch >> 2) & 1);
                    if (xx >= nx || yy >= ny || zz >= nz) continue;
                    const unsigned lin = (unsigned)((zz * ny + yy) * nx + xx);
                    const unsigned count = cl == 0 ? dc.cell_start[lin + 1] - dc.cell_start[lin] : lvl[lin];
                    if (count == 0) continue;
                    stack[top++] = ((unsigned long long)cl << 36) | ((unsigned long long)xx << 24) |
                                   ((unsigned long long)yy << 12) | (unsigned long long)zz;
                }
            }
            finish_query1<T, kOut, kStats>(sw, true, best, (long long)q.i, sum, sumsq, mc, ties);
        }
    }
    if (kStats) {
        block_reduce_stats<T>(sum, sumsq, mc, ties, sw.partial + sw.main_blocks + sw.far_blocks + blockIdx.x);
        // The CTA that finishes last folds all partials of this sweep (main pass + this pass) into the
        // caller's statistics record; for a bidirectional call the sweep that finishes second also
        // writes the Chamfer value.  No separate finalize launch.
        __shared__ bool s_last;
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = atomicAdd(sw.counters + 3, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            finalize_sweep<T>(sw, qc.n);
            if (sw.value_out != nullptr && threadIdx.x == 0) {
                __threadfence();
                if (atomicAdd(sw.pair_ticket, 1u) == 1u) {
                    __threadfence();
                    const volatile pcu_b200_nn_stats* ps = sw.pair_stats;
                    pcu_b200_nn_stats a, b;
                    a.sum_dist = ps[0].sum_dist; a.n_queries = ps[0].n_queries;
                    b.sum_dist = ps[1].sum_dist; b.n_queries = ps[1].n_queries;
                    *sw.value_out = chamfer_of<T>(a, b);
                }
            }
        }
    }
}

}  // namespace pcu
