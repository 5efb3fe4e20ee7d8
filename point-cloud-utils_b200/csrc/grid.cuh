// grid.cuh -- device-side uniform-grid build (counting sort of a cloud into cells).
//
// Replaces what the reference does before it can answer a query: nanoflann's serial kd-tree build
// (external/nanoflann/nanoflann.hpp:1363-1375, divideTree :1001-1059), which the reference runs three
// times per call.  Here: bounding box -> grid shape -> per-cell histogram (atomics
// that also hand out each point's rank in its cell) -> exclusive scan -> scatter into cell order.
// Every kernel takes an array of cloud descriptors and uses blockIdx.y as the cloud index, so the
// two clouds of a Chamfer call -- or the 2*B clouds of a batch -- are binned by the same launches.
// Nothing here synchronises with the host: the grid shape is chosen on the device.
#pragma once
#include "common.cuh"

namespace pcu {

// ---------------------------------------------------------------------------------------------
// 1. partial bounding boxes: grid (kBBoxBlocks, nclouds)
template <typename T>
__global__ void __launch_bounds__(kThreads) bbox_partial_kernel(const Cloud<T>* __restrict__ clouds) {
    using R = Real<T>;
    const Cloud<T> c = clouds[blockIdx.y];
    T lo[3] = {R::inf(), R::inf(), R::inf()};
    T hi[3] = {-R::inf(), -R::inf(), -R::inf()};
    // flat, fully coalesced walk over the 3n scalars; the axis of element e is e % 3
    const long long total = 3 * c.n;
    const long long step = (long long)gridDim.x * blockDim.x;
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int axis = (int)(e % 3);
    const int axis_step = (int)(step % 3);
    for (; e < total; e += step) {
        const T v = __ldg(c.raw + e);
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (a == axis) { lo[a] = R::vmin(lo[a], v); hi[a] = R::vmax(hi[a], v); }
        axis += axis_step;
        if (axis >= 3) axis -= 3;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = R::vmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = R::vmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    __shared__ T s[kThreads / 32][6];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { s[w][a] = lo[a]; s[w][3 + a] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        T v = s[0][threadIdx.x];
        for (int i = 1; i < kThreads / 32; ++i)
            v = threadIdx.x < 3 ? R::vmin(v, s[i][threadIdx.x]) : R::vmax(v, s[i][threadIdx.x]);
        c.bbox_partial[blockIdx.x * 6 + threadIdx.x] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// 2. grid shape + wall tables: grid (1, nclouds), kThreads threads
template <typename T>
__global__ void __launch_bounds__(kThreads) grid_setup_kernel(const Cloud<T>* __restrict__ clouds) {
    using R = Real<T>;
    using bits_t = typename R::bits_t;
    const Cloud<T> c = clouds[blockIdx.y];
    __shared__ T box[6];
    __shared__ GridHeader<T> hdr;
    if (threadIdx.x < 6) {
        T v = c.bbox_partial[threadIdx.x];
        for (int i = 1; i < kBBoxBlocks; ++i) {
            const T u = c.bbox_partial[i * 6 + threadIdx.x];
            v = threadIdx.x < 3 ? R::vmin(v, u) : R::vmax(v, u);
        }
        box[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int maxdim = c.stride - 1;
        double ext[3], emax = 0.0;
        for (int a = 0; a < 3; ++a) {
            ext[a] = (double)box[3 + a] - (double)box[a];
            if (!(ext[a] > 0.0)) ext[a] = 0.0;      // also swallows NaN
            if (!(ext[a] < 1e300)) ext[a] = 1e300;  // +inf input: keep the arithmetic finite
            emax = fmax(emax, ext[a]);
        }
        double h = 1.0;
        if (emax > 0.0) {
            auto cells_at = [&](double hh) {
                double p = 1.0;
                for (int a = 0; a < 3; ++a) {
                    double d = ceil(ext[a] / hh);
                    d = fmin(fmax(d, 1.0), (double)maxdim);
                    p *= d;
                }
                return p;
            };
            double lo_h = emax / (double)maxdim, hi_h = emax;   // cells_at(hi_h) == 1
            const double cap = (double)c.cell_cap;
            if (cells_at(lo_h) <= cap) {
                h = lo_h;
            } else {
                for (int it = 0; it < 64; ++it) {
                    const double mid = 0.5 * (lo_h + hi_h);
                    if (cells_at(mid) <= cap) hi_h = mid; else lo_h = mid;
                }
                h = hi_h;
            }
        }
        long long nc = 1;
        for (int a = 0; a < 3; ++a) {
            double d = emax > 0.0 ? ceil(ext[a] / h) : 1.0;
            d = fmin(fmax(d, 1.0), (double)maxdim);
            hdr.dim[a] = (int)d;
            nc *= (long long)d;
            hdr.origin[a] = box[a];
        }
        if (nc > (long long)c.cell_cap) {  // cannot happen (cells_at(h) <= cap); belt and braces
            hdr.dim[0] = hdr.dim[1] = hdr.dim[2] = 1; nc = 1; h = emax > 0.0 ? emax : 1.0;
        }
        hdr.ncells = (int)nc;
        hdr.h = (T)h;
        hdr.inv_h = (T)(1.0 / h);
        hdr.stride = c.stride;
        hdr.pad = 0;
        *c.grid = hdr;
    }
    __syncthreads();
    // wall tables: bisection over the ordered-integer image of the reals, using the very cell
    // function the binning kernels use, so the walls are exact by construction.
    const int stride = c.stride;
    for (int t = threadIdx.x; t < 3 * stride; t += blockDim.x) {
        const int a = t / stride, j = t - a * stride;
        const int dim = hdr.dim[a];
        T wl, wh;
        if (j == 0) { wl = -R::inf(); wh = -R::inf(); }
        else if (j >= dim) { wl = R::inf(); wh = R::inf(); }
        else {
            bits_t lo_u = ordered<T>(hdr.origin[a]);   // cell 0 < j
            bits_t hi_u = ordered<T>(R::inf());        // clamps to dim-1 >= j
            while (hi_u - lo_u > 1) {
                const bits_t mid = lo_u + (hi_u - lo_u) / 2;
                if (cell_of<T>(unordered<T>(mid), hdr.origin[a], hdr.inv_h, dim) >= j) hi_u = mid; else lo_u = mid;
            }
            wl = unordered<T>(lo_u);
            wh = unordered<T>(hi_u);
        }
        c.wall_lo[t] = wl;
        c.wall_hi[t] = wh;
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int linear_cell(const GridHeader<T>& g, T x, T y, T z) {
    const int cx = cell_of<T>(x, g.origin[0], g.inv_h, g.dim[0]);
    const int cy = cell_of<T>(y, g.origin[1], g.inv_h, g.dim[1]);
    const int cz = cell_of<T>(z, g.origin[2], g.inv_h, g.dim[2]);
    return (cz * g.dim[1] + cy) * g.dim[0] + cx;
}

// 3. histogram; the atomic's return value is the point's rank inside its cell.
//    grid (ceil(max_n / kThreads), nclouds)
template <typename T>
__global__ void __launch_bounds__(kThreads) cell_count_kernel(const Cloud<T>* __restrict__ clouds) {
    const Cloud<T> c = clouds[blockIdx.y];
    __shared__ GridHeader<T> g;
    if (threadIdx.x == 0) g = *c.grid;
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    const T x = __ldg(c.raw + 3 * i), y = __ldg(c.raw + 3 * i + 1), z = __ldg(c.raw + 3 * i + 2);
    const int lin = linear_cell<T>(g, x, y, z);
    c.rank[i] = atomicAdd(c.cell_start + lin, 1u);
}

// ---------------------------------------------------------------------------------------------
// 4. exclusive scan of cell_start[0 .. cell_cap] (three phases, in place)
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* total) {
    // kScanThreads threads; returns the exclusive prefix of v, *total = block sum
    __shared__ unsigned warp_sum[kScanThreads / 32];
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
        if (l >= o) inc += u;
    }
    if (l == 31) warp_sum[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned s = l < kScanThreads / 32 ? warp_sum[l] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, s, o);
            if (l >= o) s += u;
        }
        if (l < kScanThreads / 32) warp_sum[l] = s;   // inclusive over warps
    }
    __syncthreads();
    const unsigned before = w ? warp_sum[w - 1] : 0u;
    *total = warp_sum[kScanThreads / 32 - 1];
    __syncthreads();
    return before + inc - v;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const Cloud<T>* __restrict__ clouds) {
    const Cloud<T> c = clouds[blockIdx.y];
    const long long count = (long long)c.cell_cap + 1;
    const long long base = (long long)blockIdx.x * kScanTile;
    if (base >= count) return;
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const long long i = base + (long long)k * kScanThreads + threadIdx.x;
        if (i < count) s += c.cell_start[i];
    }
    unsigned total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) c.scan_partial[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_partials_kernel(const Cloud<T>* __restrict__ clouds) {
    const Cloud<T> c = clouds[blockIdx.y];
    const long long count = (long long)c.cell_cap + 1;
    const int nb = (int)((count + kScanTile - 1) / kScanTile);
    unsigned carry = 0;
    for (int base = 0; base < nb; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const unsigned v = i < nb ? c.scan_partial[i] : 0u;
        unsigned total;
        const unsigned ex = block_exclusive_scan(v, &total);
        if (i < nb) c.scan_partial[i] = carry + ex;
        carry += total;
    }
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const Cloud<T>* __restrict__ clouds) {
    const Cloud<T> c = clouds[blockIdx.y];
    const long long count = (long long)c.cell_cap + 1;
    const long long base = (long long)blockIdx.x * kScanTile;
    if (base >= count) return;
    // each thread owns kScanItems consecutive entries
    unsigned v[kScanItems];
    unsigned s = 0;
    const long long first = base + (long long)threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (first + k) < count ? c.cell_start[first + k] : 0u;
        s += v[k];
    }
    unsigned total;
    unsigned run = block_exclusive_scan(s, &total) + c.scan_partial[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if ((first + k) < count) c.cell_start[first + k] = run;
        run += v[k];
    }
}

// ---------------------------------------------------------------------------------------------
// 5. scatter into cell order: grid (ceil(max_n / kThreads), nclouds)
template <typename T>
__global__ void __launch_bounds__(kThreads) scatter_kernel(const Cloud<T>* __restrict__ clouds) {
    const Cloud<T> c = clouds[blockIdx.y];
    __shared__ GridHeader<T> g;
    if (threadIdx.x == 0) g = *c.grid;
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    const T x = __ldg(c.raw + 3 * i), y = __ldg(c.raw + 3 * i + 1), z = __ldg(c.raw + 3 * i + 2);
    const int lin = linear_cell<T>(g, x, y, z);
    const unsigned pos = c.cell_start[lin] + c.rank[i];
    store_pt<T>(c.sorted + pos, x, y, z, i);
}

}  // namespace pcu
