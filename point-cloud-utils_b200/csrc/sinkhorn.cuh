// sinkhorn.cuh -- dense pairwise distances and the log-domain Sinkhorn iteration.
//
// Replaces the numpy code of /root/reference/point_cloud_utils/_sinkhorn.py: pairwise_distances (:4-34),
// sinkhorn (:37-126; the iteration :104-118, the plan :120-122) and what earth_movers_distance (:129-156) adds on
// top, (P * M).sum().  Floating point: every element-wise expression is evaluated in the arrays' precision in the
// reference's order -- (-M + v) / eps, x - max, exp, log, eps * (log a - lse) -- and only the two reductions
// (the sum of the exponentials, the L1 change of the iterates) are accumulated in fp64 instead of numpy's pairwise
// fp32 / fp64 summation, so results agree to rounding, not bit for bit (tests state the tolerance).
#pragma once
#include "common.cuh"

namespace pcu {

// p-norm selector of np.linalg.norm(x, ord=p, axis=-1) for vectors
constexpr int kNorm2 = 0, kNorm1 = 1, kNormInf = 2, kNormNegInf = 3, kNorm0 = 4, kNormP = 5;

// out[b, i, j] = || a[b, i, :] - b[b, j, :] ||_p ; a: (nb, n, d), b: (nb, m, d).  grid (ceil(m / 256), n, nb)
template <typename T>
__global__ void __launch_bounds__(kThreads) pairwise_kernel(const T* __restrict__ a, const T* __restrict__ b, long long n,
                                                            long long m, int d, int kind, double p, T* __restrict__ out) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = blockIdx.y, bt = blockIdx.z;
    if (j >= m) return;
    const T* pa = a + (bt * n + i) * d;
    const T* pb = b + (bt * m + j) * d;
    using R = Real<T>;
    T acc = kind == kNormNegInf ? R::inf() : (T)0;
    for (int c = 0; c < d; ++c) {
        const T diff = R::sub(pa[c], pb[c]);
        const T ad = diff < (T)0 ? -diff : diff;
        if (kind == kNorm2) acc = R::add(acc, R::mul(diff, diff));   // no contraction: numpy rounds the product and the sum
        else if (kind == kNorm1) acc = R::add(acc, ad);
        else if (kind == kNormInf) acc = ad > acc ? ad : acc;
        else if (kind == kNormNegInf) acc = ad < acc ? ad : acc;
        else if (kind == kNorm0) acc += diff != (T)0 ? (T)1 : (T)0;
        else acc += (T)pow((double)ad, p);
    }
    if (kind == kNorm2) acc = R::root(acc);
    else if (kind == kNormP) acc = (T)pow((double)acc, 1.0 / p);
    out[(bt * n + i) * m + j] = acc;
}

template <typename T> __device__ __forceinline__ T exp_t(T x);
template <> __device__ __forceinline__ float exp_t<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double exp_t<double>(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ T log_t(T x);
template <> __device__ __forceinline__ float log_t<float>(float x) { return logf(x); }
template <> __device__ __forceinline__ double log_t<double>(double x) { return log(x); }

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];   // fixed order
    __syncthreads();
    return t;
}
template <typename T>
__device__ __forceinline__ T block_max(T v, T* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const T u = __shfl_xor_sync(0xffffffffu, v, o); v = u > v ? u : v; }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    T t = red[0];
    for (int w = 1; w < kThreads / 32; ++w) t = red[w] > t ? red[w] : t;
    __syncthreads();
    return t;
}

// One half-iteration (_sinkhorn.py:108-112): for every row r of the (nb, R, C) matrix view
//   pot_out[b, r] = eps * (log(w[b, r]) - logsumexp_c((-M[b, r, c] + pot_in[b, c]) / eps))
// and err[b] += |pot_out_old - pot_out_new|.  kTranspose: the row of the VIEW is a column of M (the v update walks
// M^T): a block then owns 32 consecutive view rows and reads M row-major, coalesced, each warp taking a slice of the
// c range.  Skipped entirely when *done != 0 (converged: the remaining launches of the call are no-ops).
// grid (R or ceil(R / 32), nb)
template <typename T, bool kTranspose>
__global__ void __launch_bounds__(kThreads) sinkhorn_half_kernel(const T* __restrict__ M, const T* __restrict__ w,
                                                                 const T* __restrict__ pot_in, T* __restrict__ pot_out,
                                                                 long long R, long long C, T eps, double* __restrict__ err,
                                                                 const int* __restrict__ done) {
    if (*done) return;
    const long long bt = blockIdx.y;
    __shared__ double red_d[kThreads / 32];
    __shared__ T red_t[kThreads / 32];
    if (!kTranspose) {
        const long long r = blockIdx.x;
        const T* row = M + (bt * R + r) * C;
        const T* pin = pot_in + bt * C;
        T mx = -Real<T>::inf();
        for (long long c = threadIdx.x; c < C; c += blockDim.x) {
            const T x = (-row[c] + pin[c]) / eps;
            mx = x > mx ? x : mx;
        }
        mx = block_max<T>(mx, red_t);
        double sum = 0.0;
        for (long long c = threadIdx.x; c < C; c += blockDim.x) {
            const T x = (-row[c] + pin[c]) / eps;
            sum += (double)exp_t<T>(x - mx);
        }
        sum = block_sum(sum, red_d);
        if (threadIdx.x == 0) {
            const T lse = log_t<T>((T)sum) + mx;
            const T nu = eps * (log_t<T>(w[bt * R + r]) - lse);
            const T old = pot_out[bt * R + r];
            pot_out[bt * R + r] = nu;
            const T diff = old - nu;
            atomicAdd(err + bt, (double)(diff < (T)0 ? -diff : diff));
        }
    } else {
        // 32 view rows (columns of M) per block; lane = column, warp w walks M rows w, w + 8, ...
        __shared__ T part_mx[kThreads / 32][32];
        __shared__ double part_sum[kThreads / 32][32];
        const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
        const long long r = (long long)blockIdx.x * 32 + lane;     // view row = column of M
        const bool ok = r < R;
        const T* base = M + bt * C * R;                            // M is (C, R) row-major here: C rows of length R
        const T* pin = pot_in + bt * C;
        T mx = -Real<T>::inf();
        for (long long c = wp; c < C; c += kThreads / 32) {
            if (ok) { const T x = (-base[c * R + r] + pin[c]) / eps; mx = x > mx ? x : mx; }
        }
        part_mx[wp][lane] = mx;
        __syncthreads();
        mx = part_mx[0][lane];
        for (int k = 1; k < kThreads / 32; ++k) mx = part_mx[k][lane] > mx ? part_mx[k][lane] : mx;
        double sum = 0.0;
        for (long long c = wp; c < C; c += kThreads / 32) {
            if (ok) { const T x = (-base[c * R + r] + pin[c]) / eps; sum += (double)exp_t<T>(x - mx); }
        }
        part_sum[wp][lane] = sum;
        __syncthreads();
        if (wp == 0 && ok) {
            double tot = 0.0;
            for (int k = 0; k < kThreads / 32; ++k) tot += part_sum[k][lane];
            const T lse = log_t<T>((T)tot) + mx;
            const T nu = eps * (log_t<T>(w[bt * R + r]) - lse);
            const T old = pot_out[bt * R + r];
            pot_out[bt * R + r] = nu;
            const T diff = old - nu;
            atomicAdd(err + bt, (double)(diff < (T)0 ? -diff : diff));
        }
    }
}

// End of an iteration (_sinkhorn.py:114-118): err_u = max_b sum_i |du|, err_v likewise; converged when both are
// below stop_thresh.  Also counts the iterations done and clears the accumulators.  One block.
__global__ void sinkhorn_check_kernel(double* __restrict__ err_u, double* __restrict__ err_v, long long nb, double stop_thresh,
                                      int* __restrict__ done, int* __restrict__ iters) {
    if (*done) return;
    __shared__ double su[kThreads], sv[kThreads];
    double mu = 0.0, mv = 0.0;
    for (long long b = threadIdx.x; b < nb; b += blockDim.x) {
        mu = fmax(mu, err_u[b]); mv = fmax(mv, err_v[b]);
        err_u[b] = 0.0; err_v[b] = 0.0;
    }
    su[threadIdx.x] = mu; sv[threadIdx.x] = mv;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { su[threadIdx.x] = fmax(su[threadIdx.x], su[threadIdx.x + o]); sv[threadIdx.x] = fmax(sv[threadIdx.x], sv[threadIdx.x + o]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *iters += 1;
        if (su[0] < stop_thresh && sv[0] < stop_thresh) *done = 1;
    }
}

// P[b, i, j] = exp((-M[b, i, j] + u[b, i] + v[b, j]) / eps)  (:120-122); optionally cost[b] += sum_ij P * M (fp64).
// grid (ceil(m / 256), n, nb)
template <typename T>
__global__ void __launch_bounds__(kThreads) sinkhorn_plan_kernel(const T* __restrict__ M, const T* __restrict__ u,
                                                                 const T* __restrict__ v, long long n, long long m, T eps,
                                                                 T* __restrict__ P, double* __restrict__ cost) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = blockIdx.y, bt = blockIdx.z;
    double mine = 0.0;
    if (j < m) {
        const long long at = (bt * n + i) * m + j;
        const T val = exp_t<T>(((-M[at] + u[bt * n + i]) + v[bt * m + j]) / eps);
        P[at] = val;
        mine = (double)(T)(val * M[at]);
    }
    if (cost != nullptr) {
        __shared__ double red[kThreads / 32];
        const double tot = block_sum(mine, red);
        if (threadIdx.x == 0 && tot != 0.0) atomicAdd(cost + bt, tot);
    }
}

}  // namespace pcu
