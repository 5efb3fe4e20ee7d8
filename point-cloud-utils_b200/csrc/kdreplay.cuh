// kdreplay.cuh -- GPU replica of the reference's kd-tree (build + traversal) for the queries whose
// answer depends on how equal distances are ordered.  (stub: filled in by the next milestone)
#pragma once
#include <atomic>
#include "common.cuh"
#include "host_util.h"
#include "../../include/pcu_b200.h"

namespace pcu {

template <typename T>
struct KdReplayBuffers {
    void carve(Carver&, long long) {}
};

__global__ void widen_counter_kernel(const unsigned* src, long long* dst) { *dst = (long long)*src; }

template <typename T>
int enqueue_tie_replay(KdReplayBuffers<T>&, const T*, const T*, long long, int, int, int, const long long*,
                       const unsigned*, T*, long long*, cudaStream_t, std::atomic<long long>&) {
    return PCU_B200_OK;
}

template <typename T>
int enqueue_witness_replay(KdReplayBuffers<T>&, const T*, const T*, long long, int, const unsigned*,
                           pcu_b200_nn_stats*, cudaStream_t, std::atomic<long long>&) {
    return PCU_B200_OK;
}

}  // namespace pcu
