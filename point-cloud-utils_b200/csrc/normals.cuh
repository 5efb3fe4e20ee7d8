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
            const double d = nrm[0] * vx + nrm[1] * vy + nrm[2] * vz;
            const double sgn = (double)((0.0 < d) - (d < 0.0));       // sign(): 0 when the dot product is 0 (:21-23)
            nrm[0] *= sgn; nrm[1] *= sgn; nrm[2] *= sgn;
            const double angle = acos(nrm[0] * vx + nrm[1] * vy + nrm[2] * vz);
            if (angle > drop_angle_threshold) { ok = false; nrm[0] = nrm[1] = nrm[2] = 0.0; }   // :166-170
        }
    }
    normals[3 * i] = (T)nrm[0]; normals[3 * i + 1] = (T)nrm[1]; normals[3 * i + 2] = (T)nrm[2];
    keep[i] = ok ? 1 : 0;
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
