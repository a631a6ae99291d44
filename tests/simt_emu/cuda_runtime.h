// TEST INFRASTRUCTURE ONLY -- what csrc/kernels.cu sees as <cuda_runtime.h> when it is compiled for the SIMT shim: the host-memory
// runtime stand-in of tests/mock_cuda plus the few extra runtime calls the kernel launchers make, and the shim itself.
#pragma once
#include "../mock_cuda/cuda_runtime.h"
#include "simt_emu.h"

enum { cudaErrorInvalidConfiguration = 9, cudaErrorPeerAccessAlreadyEnabled = 704 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemAllocationType { cudaMemAllocationTypePinned = 1 };
enum cudaMemAllocationHandleType { cudaMemHandleTypeNone = 0 };
enum cudaMemLocationType { cudaMemLocationTypeDevice = 1 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
struct cudaMemLocation {
    enum cudaMemLocationType type;
    int id;
};
struct cudaMemPoolProps {
    enum cudaMemAllocationType allocType;
    enum cudaMemAllocationHandleType handleTypes;
    struct cudaMemLocation location;
};
typedef struct MockMemPool* cudaMemPool_t;

extern "C" {
cudaError_t cudaDeviceGetAttribute(int* value, enum cudaDeviceAttr attr, int device);
cudaError_t cudaMemPoolCreate(cudaMemPool_t* pool, const struct cudaMemPoolProps* props);
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t pool, enum cudaMemPoolAttr attr, void* value);
cudaError_t cudaMallocFromPoolAsync(void** p, size_t n, cudaMemPool_t pool, cudaStream_t st);
cudaError_t cudaEventCreate(cudaEvent_t* ev);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaDeviceEnablePeerAccess(int peer, unsigned flags);
void mock_cuda_set_last_error(int e);
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F*, enum cudaFuncAttribute, int bytes) {
    return bytes <= 232448 ? cudaSuccess : cudaErrorInvalidValue;
}
