// morton.cuh -- 64-bit 3-D Morton codes and the window query over a sorted code array.
//
// Replaces src/common/morton_code.cpp (MortonCode64: constructor :43-63, decode :73-81, operator+ :131-146,
// Negate :118-129, operator- :160-163) and the five bindings of src/morton.cpp (morton_add :26-103,
// morton_subtract :106-183, morton_encode :185-239, morton_decode :253-310, morton_knn :324-414) of the reference.
// Integer work throughout: results are bit-identical.
// Layout of a code: bit 3 i + a holds bit i of coordinate a (a = 0: x); coordinates are 21-bit two's complement
// with the sign bit (bit 20 -> code bits 60 .. 62) stored INVERTED, so that unsigned order of the codes is the
// order of the signed coordinates along the curve.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcu {

constexpr unsigned long long kMortonSigns = 0x7000000000000000ull;
constexpr unsigned long long kMortonX = 0x1249249249249249ull;   // every third bit: the x coordinate

__host__ __device__ __forceinline__ unsigned long long morton_split21(int x) {   // morton_code.cpp:13-26
    unsigned long long r = (unsigned long long)(long long)x;   // `uint64_t r = x` sign-extends a negative int32 like this
    r = (r | r << 32) & 0x1f00000000ffffull;
    r = (r | r << 16) & 0x1f0000ff0000ffull;
    r = (r | r << 8) & 0x100f00f00f00f00full;
    r = (r | r << 4) & 0x10c30c30c30c30c3ull;
    r = (r | r << 2) & 0x1249249249249249ull;
    return r;
}
__host__ __device__ __forceinline__ int morton_compact21(unsigned long long x) {   // :28-41
    unsigned long long d = x & 0x1249249249249249ull;
    d = (d | d >> 2) & 0x10c30c30c30c30c3ull;
    d = (d | d >> 4) & 0x100f00f00f00f00full;
    d = (d | d >> 8) & 0x1f0000ff0000ffull;
    d = (d | d >> 16) & 0x1f00000000ffffull;
    d = (d | d >> 32);
    d = (d & 0x100000ull) ? (d | 0xffe00000ull) : d;   // sign extension
    return (int)d;
}
__host__ __device__ __forceinline__ unsigned long long morton_encode3(int x, int y, int z) {   // :48-63
    // sign bit to bit 20; in the reference `x & 0x80000000` is unsigned (the literal does not fit an int), so the shift is logical
    x = (int)(((unsigned)x & 0x80000000u) >> 11 | ((unsigned)x & 0x0fffffu));
    y = (int)(((unsigned)y & 0x80000000u) >> 11 | ((unsigned)y & 0x0fffffu));
    z = (int)(((unsigned)z & 0x80000000u) >> 11 | ((unsigned)z & 0x0fffffu));
    const unsigned long long data = morton_split21(x) | morton_split21(y) << 1 | morton_split21(z) << 2;
    return data ^ kMortonSigns;
}
__host__ __device__ __forceinline__ void morton_decode3(unsigned long long code, int& x, int& y, int& z) {   // :73-81
    const unsigned long long d = code ^ kMortonSigns;
    x = morton_compact21(d);
    y = morton_compact21(d >> 1);
    z = morton_compact21(d >> 2);
}
__host__ __device__ __forceinline__ unsigned long long morton_add2(unsigned long long a, unsigned long long b) {   // :131-146
    const unsigned long long c1 = a ^ kMortonSigns, c2 = b ^ kMortonSigns;
    const unsigned long long ym = kMortonX << 1, zm = kMortonX << 2;
    const unsigned long long xs = (c1 | ~kMortonX) + (c2 & kMortonX);
    const unsigned long long ys = (c1 | ~ym) + (c2 & ym);
    const unsigned long long zs = (c1 | ~zm) + (c2 & zm);
    return ((xs & kMortonX) | (ys & ym) | (zs & zm)) ^ kMortonSigns;
}
__host__ __device__ __forceinline__ unsigned long long morton_negate(unsigned long long a) {   // :118-129 (no sign-bit inversion there)
    const unsigned long long ym = kMortonX << 1, zm = kMortonX << 2;
    const unsigned long long d = ~a;
    const unsigned long long xs = (d | ~kMortonX) + 1, ys = (d | ~ym) + 1, zs = (d | ~zm) + 1;
    return (xs & kMortonX) | (ys & ym) | (zs & zm);
}

// pts: (n, 3) int32 or int64 (truncated to int32 like the reference's `int32_t px = pts(i, 0)`, morton.cpp:229)
template <typename I>
__global__ void morton_encode_kernel(const I* __restrict__ pts, long long n, unsigned long long* __restrict__ codes) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    codes[i] = morton_encode3((int)pts[3 * i], (int)pts[3 * i + 1], (int)pts[3 * i + 2]);
}
__global__ void morton_decode_kernel(const unsigned long long* __restrict__ codes, long long n, int* __restrict__ pts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x, y, z;
    morton_decode3(codes[i], x, y, z);
    pts[3 * i] = x; pts[3 * i + 1] = y; pts[3 * i + 2] = z;
}
// op 0: a + b; op 1: a - b = a + negate(b)  (morton_code.cpp:160-163)
__global__ void morton_addsub_kernel(const unsigned long long* __restrict__ a, const unsigned long long* __restrict__ b,
                                     long long n, int op, unsigned long long* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = morton_add2(a[i], op ? morton_negate(b[i]) : b[i]);
}

// morton_knn (morton.cpp:324-414): lower_bound of the query code in the SORTED codes, then the window of k
// consecutive positions around it (k / 2 above, the rest below, shifted back inside the array at the ends).
// k has already been clamped to n by the caller (:351).  sort_dist: order the window by squared distance to the
// query point (ties by position) -- the reference's comparator reads three uninitialised locals there (:381-398),
// so that ORDER is undefined in the reference; the window itself is exactly the reference's.
__global__ void morton_knn_kernel(const unsigned long long* __restrict__ codes, long long n,
                                  const unsigned long long* __restrict__ qcodes, long long m, int k, int sort_dist,
                                  long long* __restrict__ out_idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const unsigned long long q = qcodes[i];
    long long lo = 0, hi = n;                 // std::lower_bound
    while (lo < hi) {
        const long long mid = lo + (hi - lo) / 2;
        if (codes[mid] < q) lo = mid + 1; else hi = mid;
    }
    const int half_up = k / 2, half_down = k - half_up;
    long long upper = lo + half_up, lower = lo - half_down;
    if (upper >= n) { lower -= (upper - n); upper = n; }
    if (lower < 0) { upper += -lower; lower = 0; }
    long long* row = out_idx + i * k;
    const int count = (int)(upper - lower);
    if (!sort_dist) {
        for (int j = 0; j < count; ++j) row[j] = lower + j;
        return;
    }
    int qx, qy, qz;
    morton_decode3(q, qx, qy, qz);
    for (int j = 0; j < count; ++j) {         // insertion sort by (distance, position)
        const long long pos = lower + j;
        int x, y, z;
        morton_decode3(codes[pos], x, y, z);
        const double dx = (double)qx - x, dy = (double)qy - y, dz = (double)qz - z;
        const double d = dx * dx + dy * dy + dz * dz;
        int s = j;
        while (s > 0) {
            int px, py, pz;
            morton_decode3(codes[row[s - 1]], px, py, pz);
            const double ex = (double)qx - px, ey = (double)qy - py, ez = (double)qz - pz;
            if (ex * ex + ey * ey + ez * ez <= d) break;
            row[s] = row[s - 1];
            --s;
        }
        row[s] = pos;
    }
}

}  // namespace pcu
