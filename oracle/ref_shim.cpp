// TEST INFRASTRUCTURE ONLY -- never imported by the product package.
//
// oracle/_ref/libpcu_ref.so : the reference's OWN nearest-neighbour path, compiled from the
// vendored header where it lies (`-I/root/reference/external/nanoflann`, nothing is copied
// into this repo).  This file only supplies what the reference gets from numpyeigen/Eigen/
// pybind11 around that header:
//
//   * a stub for the two pybind11 symbols the reference's patch at nanoflann.hpp:1004 needs
//     (`PyErr_CheckSignals`, `pybind11::error_already_set`), so the .so has no Python dependency
//     and can be dlopen'ed with ctypes;
//   * a minimal row-major matrix type with the members `KDTreeEigenMatrixAdaptor` reads
//     (nanoflann.hpp:2251-2260, 2322-2337): Scalar, Index = ptrdiff_t, ColsAtCompileTime = -1,
//     rows(), cols(), coeff(i, j);
//   * a driver that performs the same sequence of steps as
//     `shortest_distances_nanoflann` (src/point_cloud_distance.cpp:21-99) and the two bindings
//     (`k_nearest_neighbors` :123-164, `one_sided_hausdorff_distance` :186-234): deep-copy both
//     inputs, construct the adaptor (tree builds #1 and #2, nanoflann.hpp:1357 and :2288), call
//     buildIndex() once more (#3, point_cloud_distance.cpp:42), sweep the queries (OpenMP only
//     when n >= 100000 and num_threads != 0, :29-30), sqrt unless squared (:84-88), pad with
//     -1 (:90-93); for Hausdorff: serial sweep (:219), first maximum in row order (:221-225).
//
// Flags (oracle/Makefile): -O3 -msse3 -ffp-contract=off -fopenmp -std=c++17, i.e. the x86 wheel's
// Release build without FMA (CMakeLists.txt:4,222-223).
#include <cstddef>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#include <array>
#include <functional>
#include <stdexcept>
#if defined(_OPENMP)
#include <omp.h>
#endif

namespace pybind11 { struct error_already_set {}; }
static inline int PyErr_CheckSignals() { return 0; }

#include "nanoflann.hpp"  // found through -I/root/reference/external/nanoflann

namespace {

template <typename T>
struct RowMajorPoints {
    using Scalar = T;
    using Index = std::ptrdiff_t;
    enum { ColsAtCompileTime = -1, RowsAtCompileTime = -1 };
    std::vector<T> storage;  // owned deep copy (point_cloud_distance.cpp:152-153)
    Index n_rows = 0;
    RowMajorPoints(const T* src, Index n) : storage(src, src + 3 * n), n_rows(n) {}
    Index rows() const { return n_rows; }
    Index cols() const { return 3; }
    T coeff(Index i, Index j) const { return storage[3 * i + j]; }
};

int resolve_threads(int num_threads) {
    // common.h:194-199: -1 -> hardware_concurrency, otherwise the explicit count
    if (num_threads < 0) return (int)std::thread::hardware_concurrency();
    return num_threads;
}

template <typename T>
void sweep(const T* query, int64_t n, const T* dataset, int64_t m, int k, int squared,
           int leaf, int num_threads, int faithful_builds, T* out_d, int64_t* out_i) {
    using Mat = RowMajorPoints<T>;
    Mat q(query, n), d(dataset, m);
    const bool run_parallel = n >= 100000 && num_threads != 0;
    const int nthr = run_parallel ? std::max(1, resolve_threads(num_threads)) : 1;
    (void)nthr;

    using Adaptor = nanoflann::KDTreeEigenMatrixAdaptor<Mat, 3, nanoflann::metric_L2_Simple>;
    Adaptor tree(3, std::cref(d), leaf);          // builds #1 and #2
    if (faithful_builds) tree.index->buildIndex();  // build #3

#if defined(_OPENMP)
#pragma omp parallel num_threads(nthr) if (run_parallel)
#endif
    {
        std::array<T, 3> p;
        std::vector<std::ptrdiff_t> idx(k);
        std::vector<T> d2(k);
#if defined(_OPENMP)
#pragma omp for
#endif
        for (int64_t i = 0; i < n; ++i) {
            for (int j = 0; j < 3; ++j) p[j] = q.coeff(i, j);
            const size_t found = tree.index->knnSearch(p.data(), (size_t)k, idx.data(), d2.data());
            for (size_t c = 0; c < found; ++c) {
                out_i[i * k + c] = idx[c];
                out_d[i * k + c] = squared ? d2[c] : std::sqrt(d2[c]);
            }
            for (int c = (int)found; c < k; ++c) {
                out_i[i * k + c] = -1;
                out_d[i * k + c] = (T)-1.0;
            }
        }
    }
}

template <typename T>
void one_sided(const T* src, int64_t n, const T* dst, int64_t m, int squared, int leaf,
               T* out_max, int64_t* out_i, int64_t* out_j) {
    std::vector<T> dist(n);
    std::vector<int64_t> corr(n);
    sweep<T>(src, n, dst, m, 1, squared, leaf, /*num_threads=*/0, 1, dist.data(), corr.data());
    int64_t best = 0;  // Eigen maxCoeff: first maximum in row order
    for (int64_t i = 1; i < n; ++i)
        if (dist[i] > dist[best]) best = i;
    *out_max = dist[best];
    *out_i = best;
    *out_j = corr[best];
}

}  // namespace

extern "C" {

int pcu_ref_knn_f32(const float* q, int64_t n, const float* d, int64_t m, int k, int squared, int leaf,
                    int num_threads, int faithful_builds, float* out_d, int64_t* out_i) {
    try { sweep<float>(q, n, d, m, k, squared, leaf, num_threads, faithful_builds, out_d, out_i); }
    catch (...) { return 1; }
    return 0;
}
int pcu_ref_knn_f64(const double* q, int64_t n, const double* d, int64_t m, int k, int squared, int leaf,
                    int num_threads, int faithful_builds, double* out_d, int64_t* out_i) {
    try { sweep<double>(q, n, d, m, k, squared, leaf, num_threads, faithful_builds, out_d, out_i); }
    catch (...) { return 1; }
    return 0;
}
int pcu_ref_one_sided_hausdorff_f32(const float* s, int64_t n, const float* t, int64_t m, int squared, int leaf,
                                    float* out_max, int64_t* out_i, int64_t* out_j) {
    try { one_sided<float>(s, n, t, m, squared, leaf, out_max, out_i, out_j); }
    catch (...) { return 1; }
    return 0;
}
int pcu_ref_one_sided_hausdorff_f64(const double* s, int64_t n, const double* t, int64_t m, int squared, int leaf,
                                    double* out_max, int64_t* out_i, int64_t* out_j) {
    try { one_sided<double>(s, n, t, m, squared, leaf, out_max, out_i, out_j); }
    catch (...) { return 1; }
    return 0;
}
int pcu_ref_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
