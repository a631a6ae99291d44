// cvh_*: the host-side CUDA plumbing of include/curvine_b200_kernels.h (pinned buffers, device buffers, async copies, streams, events)
// for callers that link nothing but this library.  Thin wrappers; every call returns the runtime's error code.
#include <cuda_runtime.h>

#include "../../../include/curvine_b200_kernels.h"

extern "C" {

int cvh_pinned_alloc(size_t bytes, void** out) {
    if (!out) return int(cudaErrorInvalidValue);
    *out = nullptr;
    return int(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
}
int cvh_pinned_free(void* p) { return p ? int(cudaFreeHost(p)) : 0; }

int cvh_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return int(cudaErrorInvalidValue);
    return int(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
}
int cvh_host_unregister(void* p) { return p ? int(cudaHostUnregister(p)) : 0; }

int cvh_device_alloc(size_t bytes, void** out) {
    if (!out) return int(cudaErrorInvalidValue);
    *out = nullptr;
    return int(cudaMalloc(out, bytes ? bytes : 1));
}
int cvh_device_free(void* d_p) { return d_p ? int(cudaFree(d_p)) : 0; }

static int copy_async(void* dst, const void* src, size_t n, cudaMemcpyKind kind, cv_stream_t stream, cv_event_t done_event) {
    if (n && (!dst || !src)) return int(cudaErrorInvalidValue);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n) {
        const cudaError_t e = cudaMemcpyAsync(dst, src, n, kind, st);
        if (e != cudaSuccess) return int(e);
    }
    return done_event ? int(cudaEventRecord(static_cast<cudaEvent_t>(done_event), st)) : 0;
}
int cvh_h2d_async(void* d_dst, const void* h_src, size_t n, cv_stream_t copy_stream, cv_event_t done_event) {
    return copy_async(d_dst, h_src, n, cudaMemcpyHostToDevice, copy_stream, done_event);
}
int cvh_d2h_async(void* h_dst, const void* d_src, size_t n, cv_stream_t stream, cv_event_t done_event) {
    return copy_async(h_dst, d_src, n, cudaMemcpyDeviceToHost, stream, done_event);
}

int cvh_stream_create(cv_stream_t* out) {
    if (!out) return int(cudaErrorInvalidValue);
    cudaStream_t s = nullptr;
    const cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    *out = s;
    return int(e);
}
int cvh_stream_destroy(cv_stream_t s) { return s ? int(cudaStreamDestroy(static_cast<cudaStream_t>(s))) : 0; }
int cvh_stream_synchronize(cv_stream_t s) { return int(cudaStreamSynchronize(static_cast<cudaStream_t>(s))); }
int cvh_stream_wait_event(cv_stream_t s, cv_event_t e) {
    if (!e) return int(cudaErrorInvalidValue);
    return int(cudaStreamWaitEvent(static_cast<cudaStream_t>(s), static_cast<cudaEvent_t>(e), 0));
}

int cvh_event_create(cv_event_t* out) {
    if (!out) return int(cudaErrorInvalidValue);
    cudaEvent_t e = nullptr;
    const cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    *out = e;
    return int(rc);
}
int cvh_event_destroy(cv_event_t e) { return e ? int(cudaEventDestroy(static_cast<cudaEvent_t>(e))) : 0; }
int cvh_event_record(cv_event_t e, cv_stream_t s) {
    if (!e) return int(cudaErrorInvalidValue);
    return int(cudaEventRecord(static_cast<cudaEvent_t>(e), static_cast<cudaStream_t>(s)));
}
int cvh_event_synchronize(cv_event_t e) { return e ? int(cudaEventSynchronize(static_cast<cudaEvent_t>(e))) : int(cudaErrorInvalidValue); }
int cvh_event_query(cv_event_t e) { return e ? int(cudaEventQuery(static_cast<cudaEvent_t>(e))) : int(cudaErrorInvalidValue); }

}  // extern "C"
