// TEST INFRASTRUCTURE ONLY -- C entry points over the reference's OWN MortonCode64 class.
// Built by oracle/Makefile together with /root/reference/src/common/morton_code.cpp (compiled where it lies; the
// header comes through -I/root/reference/src) into oracle/_ref/libpcu_ref_morton.so.  The loops restate the bodies
// of the bindings in /root/reference/src/morton.cpp (:84-86, :164-166, :229-231, :296-300, :351-407), which
// themselves need numpyeigen and cannot be compiled here.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "common/morton_code.h"

extern "C" {

void pcu_ref_morton_encode(const int32_t* pts, int64_t n, uint64_t* codes) {
    for (int64_t i = 0; i < n; ++i) {
        int32_t px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
        MortonCode64 code(px, py, pz);
        codes[i] = code.get_data();
    }
}
void pcu_ref_morton_decode(const uint64_t* codes, int64_t n, int32_t* pts) {
    for (int64_t i = 0; i < n; ++i) {
        int32_t px, py, pz;
        MortonCode64(codes[i]).decode(px, py, pz);
        pts[3 * i] = px; pts[3 * i + 1] = py; pts[3 * i + 2] = pz;
    }
}
void pcu_ref_morton_add(const uint64_t* a, const uint64_t* b, int64_t n, uint64_t* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = (MortonCode64(a[i]) + MortonCode64(b[i])).get_data();
}
void pcu_ref_morton_subtract(const uint64_t* a, const uint64_t* b, int64_t n, uint64_t* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = (MortonCode64(a[i]) - MortonCode64(b[i])).get_data();
}
// morton_knn with sort_dist = false (the sorted variant's comparator reads uninitialised variables in the reference)
void pcu_ref_morton_knn_window(const uint64_t* codes, int64_t n, const uint64_t* qcodes, int64_t m, int k, int64_t* out) {
    for (int64_t i = 0; i < m; ++i) {
        const uint64_t* code_ptr = std::lower_bound(codes, codes + n, qcodes[i]);
        std::ptrdiff_t idx = code_ptr - codes;
        const int half_k_up = k / 2;
        const int half_k_down = k - half_k_up;
        std::ptrdiff_t upper_bound = idx + half_k_up;
        std::ptrdiff_t lower_bound = idx - half_k_down;
        if (upper_bound >= n) { lower_bound -= (upper_bound - n); upper_bound = n; }
        if (lower_bound < 0) { upper_bound += -lower_bound; lower_bound = 0; }
        for (int j = 0; j < (upper_bound - lower_bound); j += 1) out[i * k + j] = lower_bound + j;
    }
}

}  // extern "C"
