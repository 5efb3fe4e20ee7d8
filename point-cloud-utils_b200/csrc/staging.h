// staging.h -- pageable host memory -> device through a pinned ring, filled by a few copy threads.
//
// cudaMemcpyAsync from pageable memory makes the driver stage the data itself, on the calling thread (measured
// here: 24 MB in 1.4 ms, 17 GB/s, against 0.44 ms from pinned memory; with this ring: 12 MB are in the ring after
// 0.2 ms of host time and the numpy-in Chamfer call of 2 x 10^6 points takes 0.82 instead of 1.6 ms).  The reference's callers hold ordinary numpy
// arrays, so that is the path a drop-in user is on.  The stager splits such a copy into 512 KB chunks: worker threads
// memcpy chunk c into slot c mod kSlots of a page-locked ring while the calling thread enqueues the H2D copy of
// every chunk as soon as it is filled, so the CPU copies (several cores) and the DMA overlap.
//
// Slot reuse is ordered by two per-slot generation counters that run on across jobs: `filled` (a worker has copied
// generation g into the slot) and `drained` (the H2D of generation g has completed: the calling thread waits for the
// event of the chunk kSlots / 2 behind the one it has just enqueued).  Dependencies only point to smaller chunk
// numbers, so nothing can deadlock; when a job returns, its last chunks may still be in flight on the stream -- their
// slots are released by the next job (or by the destructor, after the device has been synchronised by the owner).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#if defined(__linux__)
#include <sched.h>
#endif
#include <mutex>
#include <thread>

namespace pcu {

// While alive, the calling thread runs on the CPUs of the NUMA node the given GPU hangs off (sysfs: the PCI device's
// numa_node, that node's cpulist), so that page-locked memory allocated meanwhile is placed next to the GPU.  Where a
// page-locked buffer lives decides the copy rate on two-socket hosts: measured on these boxes, 24 MB H2D in 0.44 ms
// from the GPU's node against 0.6 - 1.5 ms from the other one (and fluctuating with the traffic on the socket link).
// Does nothing when the topology cannot be read (no sysfs, one node, numa_node = -1).
class NearDevice {
public:
    explicit NearDevice(int device) {
#if defined(__linux__)
        char bus[32] = {0};
        if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return; }
        for (char* c = bus; *c; ++c) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
        char path[128];
        std::snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
        int node = -1;
        if (std::FILE* f = std::fopen(path, "r")) { if (std::fscanf(f, "%d", &node) != 1) node = -1; std::fclose(f); }
        if (node < 0) return;
        std::snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
        std::FILE* f = std::fopen(path, "r");
        if (!f) return;
        char list[4096] = {0};
        const bool got = std::fgets(list, (int)sizeof list, f) != nullptr;
        std::fclose(f);
        if (!got || sched_getaffinity(0, sizeof old_, &old_) != 0) return;
        cpu_set_t want;
        CPU_ZERO(&want);
        int any = 0;
        for (char* p = list; *p;) {                    // "0-31,64-95"
            char* end = nullptr;
            const long a = std::strtol(p, &end, 10);
            if (end == p) break;
            long b = a;
            p = end;
            if (*p == '-') { b = std::strtol(p + 1, &end, 10); p = end; }
            for (long c = a; c <= b && c < CPU_SETSIZE; ++c)
                if (CPU_ISSET((int)c, &old_)) { CPU_SET((int)c, &want); ++any; }
            while (*p == ',' || *p == ' ' || *p == '\n') ++p;
        }
        if (any && sched_setaffinity(0, sizeof want, &want) == 0) bound_ = true;
#else
        (void)device;
#endif
    }
    ~NearDevice() {
#if defined(__linux__)
        if (bound_) sched_setaffinity(0, sizeof old_, &old_);
#endif
    }
    NearDevice(const NearDevice&) = delete;
    NearDevice& operator=(const NearDevice&) = delete;

private:
#if defined(__linux__)
    cpu_set_t old_;
#endif
    bool bound_ = false;
};

class HostStager {
public:
    static constexpr size_t kChunk = size_t(512) << 10;
    static constexpr int kSlots = 64;      // a 32 MB ring
    static constexpr int kWorkers = 8;
    static constexpr int kMaxRun = 8;      // chunks per DMA transfer, at most
    static constexpr size_t kMinBytes = size_t(4) << 20;   // below this the driver's own staging is as good

    HostStager() = default;
    HostStager(const HostStager&) = delete;
    HostStager& operator=(const HostStager&) = delete;

    // The owner has synchronised the device (no copy out of the ring is in flight).
    ~HostStager() {
        if (started_) {
            {
                std::lock_guard<std::mutex> lock(mu_);
                stop_ = true;
            }
            cv_.notify_all();
            for (auto& w : workers_) if (w.joinable()) w.join();
        }
        for (auto& e : sent_) if (e) cudaEventDestroy(e);
        if (ring_) cudaFreeHost(ring_);
    }

    // true: `src` is ordinary pageable memory and large enough for the ring to pay
    static bool wants(const void* src, size_t bytes) {
        if (bytes < kMinBytes) return false;
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, src) != cudaSuccess) { cudaGetLastError(); return false; }
        return attr.type == cudaMemoryTypeUnregistered;
    }

    // Enqueues dst[0, bytes) <- src[0, bytes) on `stream`; returns once every chunk's H2D copy has been enqueued
    // (src may be reused by the caller as soon as the call returns: all of it has been copied into the ring).
    cudaError_t copy(void* dst, const void* src, size_t bytes, cudaStream_t stream) {
        if (bytes == 0) return cudaSuccess;
        cudaError_t e = start();
        if (e != cudaSuccess) return e;
        const long long first = issued_;
        const long long count = (long long)((bytes + kChunk - 1) / kChunk);
        {
            std::lock_guard<std::mutex> lock(mu_);
            job_src_ = static_cast<const unsigned char*>(src);
            job_bytes_ = bytes;
            job_first_ = first;
            job_end_ = first + count;
            next_.store(first, std::memory_order_relaxed);
            active_.store(kWorkers, std::memory_order_relaxed);
            job_id_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        cudaError_t status = cudaSuccess;
        long long c = first;
        while (c < first + count) {
            const int s = (int)(c % kSlots);
            while (filled_[s].load(std::memory_order_acquire) < c / kSlots + 1) std::this_thread::yield();
            // chunks that are already filled and follow in the ring go out in the same transfer (few large DMA
            // transfers run at the link's rate, many small ones do not: 28 vs 55 GB/s measured at 512 KB)
            long long run = 1;
            while (run < kMaxRun && c + run < first + count && s + run < kSlots &&
                   filled_[s + run].load(std::memory_order_acquire) >= (c + run) / kSlots + 1) ++run;
            const size_t at = (size_t)(c - first) * kChunk;
            const size_t want = (size_t)run * kChunk;
            const size_t len = bytes - at < want ? bytes - at : want;
            if (status == cudaSuccess) status = cudaMemcpyAsync(static_cast<unsigned char*>(dst) + at, ring_ + (size_t)s * kChunk, len,
                                                                cudaMemcpyHostToDevice, stream);
            const int last = s + (int)run - 1;
            if (status == cudaSuccess) status = cudaEventRecord(sent_[last], stream);
            for (int i = s; i <= last; ++i) cover_[i] = last;
            c += run;
            issued_ = c;
            // release the slots half a ring behind (their copies have long been on the device)
            while (released_ + kSlots / 2 < issued_) release_one(status);
        }
        // the workers must have left this job before the next one rewrites its description
        while (active_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
        return status;
    }

private:
    void release_one(cudaError_t& status) {
        const int s = (int)(released_ % kSlots);
        // the event of the transfer that carried this slot; it cannot have been re-recorded since: the slot it belongs
        // to is released after this one, and only then filled and sent again
        if (status == cudaSuccess) status = cudaEventSynchronize(sent_[cover_[s]]);
        drained_[s].store(released_ / kSlots + 1, std::memory_order_release);   // even after an error: nobody may wait for ever
        ++released_;
    }

    cudaError_t start() {
        if (started_) return cudaSuccess;
        // write-combined: the CPU only ever streams into the ring and the GPU only reads it (numpy-in Chamfer of
        // 2 x 10^6 points 0.92 -> 0.82 ms against an ordinary page-locked ring)
        int device = 0;
        if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); device = 0; }
        cudaError_t e;
        {
            NearDevice near(device);       // the ring's pages next to the GPU that reads them
            e = cudaHostAlloc((void**)&ring_, kChunk * kSlots, cudaHostAllocWriteCombined);
            if (e == cudaSuccess) std::memset(ring_, 0, kChunk * kSlots);
        }
        if (e != cudaSuccess) return e;
        for (auto& ev : sent_) {
            e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        for (int s = 0; s < kSlots; ++s) { filled_[s].store(0); drained_[s].store(0); }
        for (auto& w : workers_) w = std::thread([this] { work(); });
        started_ = true;
        return cudaSuccess;
    }

    void work() {
        unsigned long long seen = 0;
        for (;;) {
            const unsigned char* src; size_t bytes; long long first, end;
            // calls tend to come in bursts (both clouds of a pair, a loop over pairs): look for the next job for a
            // moment before going to sleep on the condition variable (a wake-up costs tens of microseconds)
            for (int spin = 0; spin < 4000 && job_id_.load(std::memory_order_acquire) == seen; ++spin) std::this_thread::yield();
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_.wait(lock, [&] { return stop_ || job_id_.load(std::memory_order_relaxed) != seen; });
                if (stop_) return;
                seen = job_id_.load(std::memory_order_relaxed);
                src = job_src_; bytes = job_bytes_; first = job_first_; end = job_end_;
            }
            for (;;) {
                const long long c = next_.fetch_add(1, std::memory_order_relaxed);
                if (c >= end) break;
                const int s = (int)(c % kSlots);
                const long long gen = c / kSlots;
                while (drained_[s].load(std::memory_order_acquire) < gen) std::this_thread::yield();   // the slot's previous load is on the device
                const size_t at = (size_t)(c - first) * kChunk;
                const size_t len = bytes - at < kChunk ? bytes - at : kChunk;
                std::memcpy(ring_ + (size_t)s * kChunk, src + at, len);
                filled_[s].store(gen + 1, std::memory_order_release);
            }
            active_.fetch_sub(1, std::memory_order_release);
        }
    }

    unsigned char* ring_ = nullptr;
    cudaEvent_t sent_[kSlots] = {};
    int cover_[kSlots] = {};   // slot -> slot whose event was recorded behind the transfer that carried it
    std::thread workers_[kWorkers];
    bool started_ = false;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_ = false;
    std::atomic<unsigned long long> job_id_{0};
    const unsigned char* job_src_ = nullptr;
    size_t job_bytes_ = 0;
    long long job_first_ = 0, job_end_ = 0;
    std::atomic<long long> next_{0};
    std::atomic<int> active_{0};
    std::atomic<long long> filled_[kSlots];
    std::atomic<long long> drained_[kSlots];
    long long issued_ = 0;     // chunks whose H2D has been enqueued (calling thread only)
    long long released_ = 0;   // chunks whose slot has been handed back (calling thread only)
};

}  // namespace pcu
