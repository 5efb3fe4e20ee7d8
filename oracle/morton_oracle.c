/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's Morton-code path.
 *
 * Follows /root/reference/src/common/morton_code.cpp (SplitBy3Bits21 :13-26, CompactBy3Bits21 :28-41, the int32
 * constructor :48-63, decode :73-81, Negate :118-129, operator+ :131-146, operator- :160-163) and the loops of
 * /root/reference/src/morton.cpp (morton_add :84-86, morton_subtract :164-166, morton_encode :229-231,
 * morton_decode :296-300, morton_knn :351-407).  Only tests/ and bench.py's CPU legs may use it.
 * Parity status: PINNED against oracle/_ref/libpcu_ref_morton.so (the reference's own morton_code.cpp compiled in
 * place) by tests/test_morton.py and against the golden vectors generated from it. */
#include <stdint.h>
#include <stddef.h>

static uint64_t split21(int32_t x) {
    uint64_t r = (uint64_t)(int64_t)x;
    r = (r | r << 32) & 0x1f00000000ffffULL;
    r = (r | r << 16) & 0x1f0000ff0000ffULL;
    r = (r | r << 8) & 0x100f00f00f00f00fULL;
    r = (r | r << 4) & 0x10c30c30c30c30c3ULL;
    r = (r | r << 2) & 0x1249249249249249ULL;
    return r;
}
static int32_t compact21(uint64_t x) {
    uint64_t d = x & 0x1249249249249249ULL;
    d = (d | d >> 2) & 0x10c30c30c30c30c3ULL;
    d = (d | d >> 4) & 0x100f00f00f00f00fULL;
    d = (d | d >> 8) & 0x1f0000ff0000ffULL;
    d = (d | d >> 16) & 0x1f00000000ffffULL;
    d = (d | d >> 32);
    d = (d & 0x100000) ? (d | 0xffe00000) : d;
    return (int32_t)d;
}
static uint64_t encode3(int32_t x, int32_t y, int32_t z) {
    x = (int32_t)(((uint32_t)x & 0x80000000u) >> 11 | ((uint32_t)x & 0x0fffffu));
    y = (int32_t)(((uint32_t)y & 0x80000000u) >> 11 | ((uint32_t)y & 0x0fffffu));
    z = (int32_t)(((uint32_t)z & 0x80000000u) >> 11 | ((uint32_t)z & 0x0fffffu));
    return (split21(x) | split21(y) << 1 | split21(z) << 2) ^ 0x7000000000000000ULL;
}
static void decode3(uint64_t code, int32_t* x, int32_t* y, int32_t* z) {
    const uint64_t d = code ^ 0x7000000000000000ULL;
    *x = compact21(d); *y = compact21(d >> 1); *z = compact21(d >> 2);
}
static const uint64_t XM = 0x1249249249249249ULL;
static uint64_t add2(uint64_t a, uint64_t b) {
    const uint64_t c1 = a ^ 0x7000000000000000ULL, c2 = b ^ 0x7000000000000000ULL, ym = XM << 1, zm = XM << 2;
    const uint64_t xs = (c1 | ~XM) + (c2 & XM), ys = (c1 | ~ym) + (c2 & ym), zs = (c1 | ~zm) + (c2 & zm);
    return ((xs & XM) | (ys & ym) | (zs & zm)) ^ 0x7000000000000000ULL;
}
static uint64_t negate(uint64_t a) {
    const uint64_t ym = XM << 1, zm = XM << 2, d = ~a;
    const uint64_t xs = (d | ~XM) + 1, ys = (d | ~ym) + 1, zs = (d | ~zm) + 1;
    return (xs & XM) | (ys & ym) | (zs & zm);
}

void pcu_oracle_morton_encode(const int32_t* pts, int64_t n, uint64_t* codes) {
    for (int64_t i = 0; i < n; ++i) codes[i] = encode3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
}
void pcu_oracle_morton_decode(const uint64_t* codes, int64_t n, int32_t* pts) {
    for (int64_t i = 0; i < n; ++i) decode3(codes[i], pts + 3 * i, pts + 3 * i + 1, pts + 3 * i + 2);
}
void pcu_oracle_morton_add(const uint64_t* a, const uint64_t* b, int64_t n, uint64_t* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = add2(a[i], b[i]);
}
void pcu_oracle_morton_subtract(const uint64_t* a, const uint64_t* b, int64_t n, uint64_t* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = add2(a[i], negate(b[i]));
}
/* the window of morton_knn (k already clamped to n); unsorted, i.e. sort_dist = False */
void pcu_oracle_morton_knn_window(const uint64_t* codes, int64_t n, const uint64_t* qcodes, int64_t m, int k, int64_t* out) {
    for (int64_t i = 0; i < m; ++i) {
        int64_t lo = 0, hi = n;
        while (lo < hi) { const int64_t mid = lo + (hi - lo) / 2; if (codes[mid] < qcodes[i]) lo = mid + 1; else hi = mid; }
        const int half_up = k / 2, half_down = k - half_up;
        int64_t upper = lo + half_up, lower = lo - half_down;
        if (upper >= n) { lower -= (upper - n); upper = n; }
        if (lower < 0) { upper += -lower; lower = 0; }
        for (int64_t j = 0; j < upper - lower; ++j) out[i * k + j] = lower + j;
    }
}
