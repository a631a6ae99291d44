// TEST INFRASTRUCTURE ONLY -- a host-memory stand-in for the handful of CUDA runtime calls the C++ host side of
// curvine_b200 makes (csrc/host/*.cu), so that the ingest pipeline's HOST logic (job planning, pinned ring, copy groups,
// registrar, fetch threads, verify batching, result harvesting) can run on a machine without a GPU, under ASan/TSan.
// "Device memory" is malloc'ed host memory, copies are memcpy at enqueue time, streams and events are trivially complete.
// Built only by tests/mock_cuda/build.py into a separate library under /tmp; the product library never sees this header
// (curvine_b200/build.py compiles with nvcc against the real runtime) and nothing under curvine_b200/ can load the mock.
#pragma once
#include <stddef.h>
#include <stdint.h>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorInvalidDevice = 101, cudaErrorHostMemoryAlreadyRegistered = 712,
       cudaErrorHostMemoryNotRegistered = 713, cudaErrorNotReady = 600, cudaErrorNotSupported = 801 };
typedef struct MockStream* cudaStream_t;
typedef struct MockEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes {
    enum cudaMemoryType type;
    int device;
    void* devicePointer;
    void* hostPointer;
};
#define cudaHostAllocDefault 0u
#define cudaHostRegisterDefault 0u
#define cudaHostRegisterReadOnly 8u
#define cudaStreamNonBlocking 1u
#define cudaEventDisableTiming 2u

extern "C" {
cudaError_t mock_cudaMalloc(void** p, size_t n);
cudaError_t mock_cudaHostAlloc(void** p, size_t n, unsigned flags);
cudaError_t mock_cudaMallocAsync(void** p, size_t n, cudaStream_t st);
cudaError_t cudaFree(void* p);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaFreeAsync(void* p, cudaStream_t st);
cudaError_t cudaHostRegister(void* p, size_t n, unsigned flags);
cudaError_t cudaHostUnregister(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, enum cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, enum cudaMemcpyKind kind, cudaStream_t st);
cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t st);
cudaError_t cudaSetDevice(int dev);
cudaError_t cudaGetDevice(int* dev);
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaDeviceSynchronize(void);
cudaError_t cudaDeviceGetPCIBusId(char* buf, int len, int dev);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t st);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t ev, unsigned flags);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* ev, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t ev);
cudaError_t cudaEventRecord(cudaEvent_t ev, cudaStream_t st);
cudaError_t cudaEventSynchronize(cudaEvent_t ev);
cudaError_t cudaEventQuery(cudaEvent_t ev);
cudaError_t cudaPointerGetAttributes(struct cudaPointerAttributes* a, const void* p);
cudaError_t cudaGetLastError(void);
const char* cudaGetErrorString(cudaError_t e);
// test hooks
void mock_cuda_set_last_error(int e);
void mock_cuda_set_register_supported(int on);   // 0: cudaHostRegister answers cudaErrorNotSupported (read-only fs, disk-backed mappings)
void mock_cuda_counters(uint64_t out[6]);         // memcpy calls, memcpy bytes, registered ranges now, register calls, live device allocs, live pinned allocs
}

// the typed overloads the real headers provide
template <typename T>
static inline cudaError_t cudaMalloc(T** p, size_t n) { return mock_cudaMalloc(reinterpret_cast<void**>(p), n); }
template <typename T>
static inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned flags) { return mock_cudaHostAlloc(reinterpret_cast<void**>(p), n, flags); }
template <typename T>
static inline cudaError_t cudaMallocAsync(T** p, size_t n, cudaStream_t st) { return mock_cudaMallocAsync(reinterpret_cast<void**>(p), n, st); }
