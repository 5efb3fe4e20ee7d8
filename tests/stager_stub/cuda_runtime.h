// Stand-in for <cuda_runtime.h> so that csrc/staging.h (the pinned-ring stager of the host entry points) can be
// exercised on a machine without a GPU: page-locked memory is malloc, an "asynchronous" copy is a memcpy, events
// complete at once.  The ring's slot protocol (two generation counters per slot, copy threads, coalesced
// transfers) is what tests/test_host_stager.py stresses with it.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <cstring>
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef int* cudaEvent_t;
enum { cudaSuccess = 0 };
enum { cudaHostAllocDefault = 0, cudaHostAllocWriteCombined = 4, cudaEventDisableTiming = 2, cudaMemcpyHostToDevice = 1 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1 };
struct cudaPointerAttributes { int type; };
inline cudaError_t cudaHostAlloc(void** p, size_t n, int) { *p = std::malloc(n); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, int) { *e = new int(0); return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->type = cudaMemoryTypeUnregistered; return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaDeviceGetPCIBusId(char* bus, int len, int) { if (len > 12) std::strcpy(bus, "0000:00:00.0"); return 0; }
