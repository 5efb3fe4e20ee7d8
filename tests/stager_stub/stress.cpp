// Stress test of csrc/staging.h against the stand-in runtime: many copies of random sizes (single bytes to several
// rings) back to back through one stager; every byte must arrive.  Exit code 0 = pass.
#include "staging.h"
#include <cstdio>
#include <random>
#include <vector>
int main(int argc, char** argv) {
    const int rounds = argc > 1 ? std::atoi(argv[1]) : 150;
    pcu::HostStager stager;
    std::mt19937 rng(12345);
    for (int it = 0; it < rounds; ++it) {
        size_t bytes = (size_t)(rng() % (70u << 20)) + 1;                 // up to ~2 rings
        if (it % 5 == 0) bytes = (size_t)(rng() % 9000) + 1;              // tiny
        if (it % 11 == 0) bytes = pcu::HostStager::kChunk * (1 + rng() % 70) + (rng() % 2);   // chunk multiples, +-1
        std::vector<unsigned char> src(bytes), dst(bytes, 0xee);
        for (size_t i = 0; i < bytes; i += 1021) src[i] = (unsigned char)(i * 131 + it);
        src[bytes - 1] = (unsigned char)it;
        if (stager.copy(dst.data(), src.data(), bytes, nullptr) != cudaSuccess) { std::printf("copy failed at round %d\n", it); return 2; }
        if (std::memcmp(src.data(), dst.data(), bytes) != 0) { std::printf("mismatch at round %d (%zu bytes)\n", it, bytes); return 1; }
    }
    std::printf("ok %d rounds\n", rounds);
    return 0;
}
