// normals.cuh -- point-cloud normals from k nearest neighbours.
//
// Replaces estimate_local_normal_knn + estimate_normals of the reference
// (src/point_cloud_normals.cpp:115-173, :175-300; binding :375-411): for every point, the k nearest points of
// the SAME cloud (the point itself included, as nanoflann's knnSearch returns it), the plane fitted to them
// and, optionally, orientation towards / filtering by a per-point view direction.
//
// The neighbour search is the exact top-k path of this library (topk.cuh + tie replay), so the neighbour SETS
// are the reference's, bit for bit.  The plane fit is where parity is a tolerance: the reference takes the
// right singular vector of the smallest singular value of the (k, 3) matrix of neighbour offsets from Eigen's
// JacobiSVD in double precision (Eigen is fetched at build time and is not in the reference tree); here it is
// the eigenvector of the smallest eigenvalue of the 3 x 3 scatter matrix A^T A, accumulated in fp64 and
// diagonalised by cyclic Jacobi rotations -- the same subspace, equal up to rounding (and up to sign: without
// view directions the reference's sign is whatever its SVD happens to return).
#pragma once
#include "common.cuh"

namespace pcu {

// One Jacobi rotation annihilating a[p][q] of the symmetric 3 x 3 matrix a; v accumulates the rotations.
template <int P, int Q>
__device__ __forceinline__ void jacobi_rotate(double (&a)[3][3], double (&v)[3][3]) {
    const double apq = a[P][Q];
    if (apq == 0.0) return;
    const double theta = (a[Q][Q] - a[P][P]) / (2.0 * apq);
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
    constexpr int R = 3 - P - Q;   // the third index
    const double app = a[P][P], aqq = a[Q][Q], arp = a[R][P], arq = a[R][Q];
    a[P][P] = app - t * apq;
    a[Q][Q] = aqq + t * apq;
    a[P][Q] = a[Q][P] = 0.0;
    a[R][P] = a[P][R] = c * arp - s * arq;
    a[R][Q] = a[Q][R] = s * arp + c * arq;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double vip = v[i][P], viq = v[i][Q];
        v[i][P] = c * vip - s * viq;
        v[i][Q] = s * vip + c * viq;
    }
}

// Unit eigenvector of the smallest eigenvalue of the symmetric positive semi-definite matrix
// [[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]].
__device__ __forceinline__ void smallest_eigenvector(double xx, double xy, double xz, double yy, double yz, double zz,
                                                     double (&n)[3]) {
    double a[3][3] = {{xx, xy, xz}, {xy, yy, yz}, {xz, yz, zz}};
    double v[3][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}};
    const double scale = fabs(xx) + fabs(yy) + fabs(zz);
    for (int sweep = 0; sweep < 16; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (!(off > 1e-300) || off <= 1e-22 * scale) break;
        jacobi_rotate<0, 1>(a, v);
        jacobi_rotate<0, 2>(a, v);
        jacobi_rotate<1, 2>(a, v);
    }
    const double e0 = a[0][0], e1 = a[1][1], e2 = a[2][2];
    const int which = (e0 <= e1 && e0 <= e2) ? 0 : (e1 <= e2 ? 1 : 2);
    double nx = which == 0 ? v[0][0] : (which == 1 ? v[0][1] : v[0][2]);
    double ny = which == 0 ? v[1][0] : (which == 1 ? v[1][1] : v[1][2]);
    double nz = which == 0 ? v[2][0] : (which == 1 ? v[2][1] : v[2][2]);
    const double len = sqrt(nx * nx + ny * ny + nz * nz);
    if (len > 0.0) { nx /= len; ny /= len; nz /= len; }
    n[0] = nx; n[1] = ny; n[2] = nz;
}

// Orientation towards / filtering by the view direction (:160-171 and :103-112, the same code in both estimators).
__device__ __forceinline__ bool orient_by_view(double (&nrm)[3], double vx, double vy, double vz, double drop_angle_threshold) {
    const double d = nrm[0] * vx + nrm[1] * vy + nrm[2] * vz;
    const double sgn = (double)((0.0 < d) - (d < 0.0));       // sign(): 0 when the dot product is 0 (:21-23)
    nrm[0] *= sgn; nrm[1] *= sgn; nrm[2] *= sgn;
    const double angle = acos(nrm[0] * vx + nrm[1] * vy + nrm[2] * vz);
    if (angle > drop_angle_threshold) { nrm[0] = nrm[1] = nrm[2] = 0.0; return false; }
    return true;
}

// One thread per point.  idx: (n, k) neighbour rows from the k-NN call (-1 padded when the cloud has fewer than k
// points: such points are dropped, src/point_cloud_normals.cpp:139-142).  view_dirs: (n, 3) or null.
// normals: (n, 3) dense, written for every point; keep: (n) 1 = the reference would return this point.
template <typename T>
__global__ void __launch_bounds__(kThreads) normals_knn_kernel(const T* __restrict__ points, long long n,
                                                               const long long* __restrict__ idx, int k,
                                                               const T* __restrict__ view_dirs, double drop_angle_threshold,
                                                               T* __restrict__ normals, unsigned char* __restrict__ keep) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long* row = idx + i * k;
    bool ok = row[k - 1] >= 0;
    double nrm[3] = {0.0, 0.0, 0.0};
    if (ok) {
        double xx = 0.0, xy = 0.0, xz = 0.0, yy = 0.0, yz = 0.0, zz = 0.0;
        for (int j = 0; j < k; ++j) {
            const long long p = row[j];
            // the reference subtracts in the cloud's precision and widens the difference (:150-154)
            const double dx = (double)(T)(points[3 * p] - points[3 * i]);
            const double dy = (double)(T)(points[3 * p + 1] - points[3 * i + 1]);
            const double dz = (double)(T)(points[3 * p + 2] - points[3 * i + 2]);
            xx += dx * dx; xy += dx * dy; xz += dx * dz; yy += dy * dy; yz += dy * dz; zz += dz * dz;
        }
        smallest_eigenvector(xx, xy, xz, yy, yz, zz, nrm);
        if (view_dirs != nullptr) {
            const double vx = (double)view_dirs[3 * i], vy = (double)view_dirs[3 * i + 1], vz = (double)view_dirs[3 * i + 2];
            ok = orient_by_view(nrm, vx, vy, vz, drop_angle_threshold);   // :160-171
        }
    }
    normals[3 * i] = (T)nrm[0]; normals[3 * i + 1] = (T)nrm[1]; normals[3 * i + 2] = (T)nrm[2];
    keep[i] = ok ? 1 : 0;
}

// ---- normals from all points in a ball (estimate_local_normal_rbf, src/point_cloud_normals.cpp:48-113; binding
// :303-370) ------------------------------------------------------------------------------------------------------
// The reference hands `ball_radius` to nanoflann's radiusSearch, whose radius is in the metric's units -- SQUARED
// distance for L2_Simple (nanoflann.hpp RadiusResultSet::addPoint: `if (dist < radius)`) -- so the neighbourhood is
//     { j : d2(p_i, p_j) < (T)ball_radius }        (the point itself included),
// while the weight function receives the true distance: w = wendland(sqrt(d2), ball_radius).  Both are reproduced as
// they are.  The neighbourhood test uses the reference-rounded d2 of common.cuh, so neighbour SETS are exact.
//
// max_pts_per_ball > 0: the reference shuffles the neighbours with std::shuffle(default_random_engine(rand())) and
// keeps the first max_pts_per_ball -- a uniformly random subset that depends on the C library's rand() state and on
// the thread that happens to process the point.  Here: a uniformly random subset of the same size by selection
// sampling (Knuth 3.4.2 S) over the neighbours in cell order, driven by a counter-based hash of (seed, row, position):
// deterministic for a given seed, the same distribution, not the same subset.
struct BallParams {
    double radius;
    double drop_angle_threshold;
    int min_pts, max_pts;
    int weight_kind;        // 0 constant, 1 Wendland ("rbf")
    unsigned seed;
};

__device__ __forceinline__ unsigned mix32(unsigned h) {
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}

template <typename T> __device__ __forceinline__ T sub_down(T a, T b);
template <> __device__ __forceinline__ float sub_down<float>(float a, float b) { return __fsub_rd(a, b); }
template <> __device__ __forceinline__ double sub_down<double>(double a, double b) { return __dsub_rd(a, b); }
template <typename T> __device__ __forceinline__ T add_up(T a, T b);
template <> __device__ __forceinline__ float add_up<float>(float a, float b) { return __fadd_ru(a, b); }
template <> __device__ __forceinline__ double add_up<double>(double a, double b) { return __dadd_ru(a, b); }

// Calls visit(point, d2) for every point of the binned cloud with d2(q, point) < r2, in cell order.
// Coverage: d2 >= fl(fl(q-p)^2) per axis and every rounding is monotone, so |q_a - p_a| <= sqrt(r2) (1 + 2 ulp) for
// every qualifying point; the box below is that interval, widened and rounded outwards, mapped through the monotone
// cell_of.  Rows (fixed y, z) whose wall gap already reaches r2 are skipped with the bound of search.cuh.
template <typename T, typename Visit>
__device__ __forceinline__ void walk_ball(const Cloud<T>& cl, const GridHeader<T>& g, const Pt<T>& q, T r2, Visit&& visit) {
    using R = Real<T>;
    const T eps = sizeof(T) == 4 ? (T)1.1920929e-7 : (T)2.220446049250313e-16;
    const T reach = R::root(r2) * ((T)1 + (T)8 * eps) + (sizeof(T) == 4 ? (T)1.1754944e-38 : (T)2.2250738585072014e-308);
    int lo[3], hi[3], c[3];
    const T qv[3] = {q.x, q.y, q.z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = cell_of<T>(sub_down<T>(qv[a], reach), g.origin[a], g.inv_h, g.dim[a]);
        hi[a] = cell_of<T>(add_up<T>(qv[a], reach), g.origin[a], g.inv_h, g.dim[a]);
        c[a] = cell_of<T>(qv[a], g.origin[a], g.inv_h, g.dim[a]);
    }
    const int st = g.stride;
    const T* lo_y = cl.wall_lo + st;      const T* hi_y = cl.wall_hi + st;
    const T* lo_z = cl.wall_lo + 2 * st;  const T* hi_z = cl.wall_hi + 2 * st;
    for (int z = lo[2]; z <= hi[2]; ++z) {
        const T bz = z < c[2] ? sq_gap<T>(q.z, lo_z[z + 1]) : (z > c[2] ? sq_gap<T>(q.z, hi_z[z]) : (T)0);
        if (!(bz < r2)) continue;
        for (int y = lo[1]; y <= hi[1]; ++y) {
            const T by = y < c[1] ? sq_gap<T>(q.y, lo_y[y + 1]) : (y > c[1] ? sq_gap<T>(q.y, hi_y[y]) : (T)0);
            if (!(R::add(by, bz) < r2)) continue;
            const unsigned row = (unsigned)((z * g.dim[1] + y) * g.dim[0]);
            const unsigned a = cl.cell_start[row + lo[0]], b = cl.cell_start[row + hi[0] + 1];
            for (unsigned j = a; j < b; ++j) {
                const Pt<T> p = load_pt<T>(cl.sorted + j);
                const T d2 = dist2<T>(q.x, q.y, q.z, p.x, p.y, p.z);
                if (d2 < r2) visit(p, d2);
            }
        }
    }
}

// The binning kernels place the points of a cell in arrival order (an atomic counter), which changes from run to run.
// The ball walk adds its neighbours up in cell order, so for reproducible sums -- and a reproducible random subset --
// the points of every cell are put in ascending row order first.  One thread per cell; cells hold a few points
// (insertion sort), the odd crowded cell is heap-sorted in place.
template <typename T>
__device__ __forceinline__ void sift_down(Pt<T>* p, unsigned root, unsigned m) {
    for (;;) {
        unsigned child = 2 * root + 1;
        if (child >= m) return;
        if (child + 1 < m && p[child].i < p[child + 1].i) ++child;
        if (!(p[root].i < p[child].i)) return;
        const Pt<T> t = p[root]; p[root] = p[child]; p[child] = t;
        root = child;
    }
}
template <typename T>
__global__ void __launch_bounds__(kThreads) cell_order_kernel(Cloud<T> cl) {
    grid_dependency_wait();
    const int ncells = cl.grid->ncells;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += (long long)gridDim.x * blockDim.x) {
        const unsigned a = cl.cell_start[c], b = cl.cell_start[c + 1];
        const unsigned m = b - a;
        if (m < 2) continue;
        Pt<T>* p = cl.sorted + a;
        if (m <= 24) {
            for (unsigned j = 1; j < m; ++j) {
                const Pt<T> key = p[j];
                unsigned t = j;
                while (t > 0 && p[t - 1].i > key.i) { p[t] = p[t - 1]; --t; }
                p[t] = key;
            }
        } else {
            for (unsigned r = m / 2; r-- > 0;) sift_down<T>(p, r, m);
            for (unsigned end = m - 1; end > 0; --end) {
                const Pt<T> t = p[0]; p[0] = p[end]; p[end] = t;
                sift_down<T>(p, 0, end);
            }
        }
    }
}

// One thread per point, in CELL order (neighbouring lanes walk overlapping boxes).  normals / keep are indexed by the
// caller's row.
template <typename T>
__global__ void __launch_bounds__(kThreads) normals_ball_kernel(Cloud<T> cl, const T* __restrict__ view_dirs, BallParams bp,
                                                                T* __restrict__ normals, unsigned char* __restrict__ keep) {
    grid_dependency_wait();
    using R = Real<T>;
    __shared__ GridHeader<T> g;
    if (threadIdx.x == 0) g = *cl.grid;
    __syncthreads();
    const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= cl.n) return;
    const Pt<T> q = load_pt<T>(cl.sorted + pos);
    const long long row = (long long)q.i;
    const T r2 = (T)bp.radius;                       // radiusSearch takes a DistanceType: the double is narrowed (:74)
    double xx = 0.0, xy = 0.0, xz = 0.0, yy = 0.0, yz = 0.0, zz = 0.0;
    auto fold = [&](const Pt<T>& p, T d2) {
        double w = 1.0;
        if (bp.weight_kind == 1) {                   // wendland_rbf (:334-339), in double like the reference
            const double r = sqrt((double)d2) / bp.radius;
            const double v1 = 1.0 - r, v2 = 4 * r + 1.0;
            w = v1 * v1 * v1 * v1 * v2;
        }
        // (points(nbr, j) - query[j]) * weight: the difference in the cloud's precision, the product in double (:89)
        const double dx = (double)R::sub(p.x, q.x) * w, dy = (double)R::sub(p.y, q.y) * w, dz = (double)R::sub(p.z, q.z) * w;
        xx += dx * dx; xy += dx * dy; xz += dx * dz; yy += dy * dy; yz += dy * dz; zz += dz * dz;
    };
    long long found = 0;
    walk_ball<T>(cl, g, q, r2, [&](const Pt<T>& p, T d2) { ++found; fold(p, d2); });
    bool ok = found >= (long long)bp.min_pts;        // :75-77 (before the cap)
    if (ok && bp.max_pts > 0 && found > (long long)bp.max_pts) {
        xx = xy = xz = yy = yz = zz = 0.0;
        unsigned long long remaining = (unsigned long long)found, needed = (unsigned long long)bp.max_pts;
        unsigned t = 0;
        const unsigned salt = mix32(bp.seed ^ mix32((unsigned)row * 0x9e3779b9u + 0x85ebca6bu));
        walk_ball<T>(cl, g, q, r2, [&](const Pt<T>& p, T d2) {
            const unsigned long long u = ((unsigned long long)mix32(salt + t * 0x9e3779b9u) * remaining) >> 32;   // uniform in [0, remaining)
            ++t;
            if (u < needed) { fold(p, d2); --needed; }
            --remaining;
        });
    }
    double nrm[3] = {0.0, 0.0, 0.0};
    if (ok) {
        smallest_eigenvector(xx, xy, xz, yy, yz, zz, nrm);
        if (view_dirs != nullptr)
            ok = orient_by_view(nrm, (double)view_dirs[3 * row], (double)view_dirs[3 * row + 1], (double)view_dirs[3 * row + 2],
                                bp.drop_angle_threshold);
    }
    normals[3 * row] = (T)nrm[0]; normals[3 * row + 1] = (T)nrm[1]; normals[3 * row + 2] = (T)nrm[2];
    keep[row] = ok ? 1 : 0;
}

// Order-preserving compaction of the kept points (the reference's single-thread loop appends in index order,
// :267-279): per-block counts, one block turns them into offsets, every block scatters its kept rows.
__global__ void __launch_bounds__(kThreads) keep_count_kernel(const unsigned char* __restrict__ keep, long long n, unsigned* __restrict__ block_count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned mine = i < n ? keep[i] : 0u;
    const unsigned total = __syncthreads_count((int)mine);
    if (threadIdx.x == 0) block_count[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) keep_offsets_kernel(unsigned* __restrict__ block_count, long long nblocks, long long* __restrict__ out_count) {
    __shared__ unsigned warp_sum[32];
    __shared__ unsigned carry_s;
    if (threadIdx.x == 0) carry_s = 0u;
    __syncthreads();
    for (long long base = 0; base < nblocks; base += 1024) {
        const long long b = base + threadIdx.x;
        const unsigned v = b < nblocks ? block_count[b] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += u;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned s = warp_sum[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_up_sync(0xffffffffu, s, o);
                if (threadIdx.x >= o) s += u;
            }
            warp_sum[threadIdx.x] = s;
        }
        __syncthreads();
        const unsigned before = (threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0u;
        const unsigned carry = carry_s;
        if (b < nblocks) block_count[b] = carry + before + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_count = (long long)carry_s;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) keep_scatter_kernel(const unsigned char* __restrict__ keep, long long n,
                                                                const unsigned* __restrict__ block_offset,
                                                                const T* __restrict__ normals, long long* __restrict__ out_idx,
                                                                T* __restrict__ out_normals) {
    __shared__ unsigned warp_sum[kThreads / 32];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned mine = i < n ? keep[i] : 0u;
    const unsigned ballot = __ballot_sync(0xffffffffu, mine != 0u);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_sum[w] = __popc(ballot);
    __syncthreads();
    unsigned before = block_offset[blockIdx.x];
    for (int j = 0; j < w; ++j) before += warp_sum[j];
    if (mine) {
        const unsigned pos = before + __popc(ballot & ((1u << lane) - 1u));
        out_idx[pos] = i;
        out_normals[3ll * pos] = normals[3 * i];
        out_normals[3ll * pos + 1] = normals[3 * i + 1];
        out_normals[3ll * pos + 2] = normals[3 * i + 2];
    }
}

}  // namespace pcu
