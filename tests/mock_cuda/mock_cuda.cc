// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory).
#include "cuda_runtime.h"

#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>

namespace {
std::mutex g_mu;
std::map<uintptr_t, size_t> g_registered;   // cudaHostRegister'ed ranges
std::map<uintptr_t, size_t> g_dev, g_pinned;  // live allocations
std::atomic<uint64_t> g_memcpy_calls{0}, g_memcpy_bytes{0}, g_register_calls{0};
std::atomic<int> g_register_supported{1};
thread_local int t_device = 0;
struct MockStreamImpl { int id; };
struct MockEventImpl { int recorded; };
}  // namespace

extern "C" {

static cudaError_t alloc_into(std::map<uintptr_t, size_t>& m, void** p, size_t n) {
    void* q = nullptr;
    if (posix_memalign(&q, 4096, n ? n : 1) != 0) return cudaErrorMemoryAllocation;
    memset(q, 0xCD, n);  // device memory is never zero by accident
    std::lock_guard<std::mutex> lk(g_mu);
    m[reinterpret_cast<uintptr_t>(q)] = n;
    *p = q;
    return cudaSuccess;
}
static cudaError_t free_from(std::map<uintptr_t, size_t>& m, void* p) {
    if (!p) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = m.find(reinterpret_cast<uintptr_t>(p));
        if (it == m.end()) return cudaErrorInvalidValue;
        m.erase(it);
    }
    free(p);
    return cudaSuccess;
}
cudaError_t mock_cudaMalloc(void** p, size_t n) { return alloc_into(g_dev, p, n); }
cudaError_t mock_cudaMallocAsync(void** p, size_t n, cudaStream_t) { return alloc_into(g_dev, p, n); }
cudaError_t mock_cudaHostAlloc(void** p, size_t n, unsigned) { return alloc_into(g_pinned, p, n); }
cudaError_t cudaFree(void* p) { return free_from(g_dev, p); }
cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return free_from(g_dev, p); }
cudaError_t cudaFreeHost(void* p) { return free_from(g_pinned, p); }

cudaError_t cudaHostRegister(void* p, size_t n, unsigned flags) {
    g_register_calls++;
    if (!g_register_supported.load() || (flags & cudaHostRegisterReadOnly)) return cudaErrorNotSupported;
    // the real call pins the pages: touch every page so an invalid mapping faults here, like the driver would fail
    volatile const uint8_t* b = static_cast<const uint8_t*>(p);
    uint8_t acc = 0;
    for (size_t i = 0; i < n; i += 4096) acc ^= b[i];
    (void)acc;
    std::lock_guard<std::mutex> lk(g_mu);
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    auto it = g_registered.upper_bound(a);
    if (it != g_registered.begin()) {
        auto pr = std::prev(it);
        if (pr->first + pr->second > a) return cudaErrorHostMemoryAlreadyRegistered;
    }
    if (it != g_registered.end() && it->first < a + n) return cudaErrorHostMemoryAlreadyRegistered;
    g_registered[a] = n;
    return cudaSuccess;
}
cudaError_t cudaHostUnregister(void* p) {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_registered.erase(reinterpret_cast<uintptr_t>(p)) ? cudaSuccess : cudaErrorHostMemoryNotRegistered;
}

cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, enum cudaMemcpyKind) {
    g_memcpy_calls++, g_memcpy_bytes += n;
    memmove(dst, src, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, enum cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(dst, src, n, k); }
cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t) {
    memset(dst, v, n);
    return cudaSuccess;
}
cudaError_t cudaSetDevice(int dev) {
    if (dev < 0 || dev >= 8) return cudaErrorInvalidDevice;  // eight pretend devices, all the same host memory
    t_device = dev;
    return cudaSuccess;
}
cudaError_t cudaGetDevice(int* dev) {
    *dev = t_device;
    return cudaSuccess;
}
cudaError_t cudaGetDeviceCount(int* n) {
    *n = 8;
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaDeviceGetPCIBusId(char* buf, int len, int) {
    strncpy(buf, "0000:00:00.0", static_cast<size_t>(len));  // no such sysfs node: the ingest falls back to "no NUMA binding"
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned) {
    *st = reinterpret_cast<cudaStream_t>(new MockStreamImpl{1});
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t st) {
    delete reinterpret_cast<MockStreamImpl*>(st);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* ev, unsigned) {
    *ev = reinterpret_cast<cudaEvent_t>(new MockEventImpl{0});
    return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t ev) {
    delete reinterpret_cast<MockEventImpl*>(ev);
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(struct cudaPointerAttributes* a, const void* p) {
    // pinned allocations and registered ranges are host memory; everything else a caller hands in as a destination is "device"
    const uintptr_t x = reinterpret_cast<uintptr_t>(p);
    std::lock_guard<std::mutex> lk(g_mu);
    auto inside = [&](const std::map<uintptr_t, size_t>& m) {
        auto it = m.upper_bound(x);
        if (it == m.begin()) return false;
        --it;
        return x < it->first + it->second;
    };
    a->type = (inside(g_pinned) || inside(g_registered)) ? cudaMemoryTypeHost : cudaMemoryTypeDevice;
    a->device = t_device;  // "device memory" belongs to whichever device the asking thread has selected
    a->devicePointer = const_cast<void*>(p), a->hostPointer = nullptr;
    return cudaSuccess;
}
// sticky-until-read error of the calling thread (the SIMT shim of tests/simt_emu records invalid launch configurations here)
static thread_local int t_last_error = cudaSuccess;
void mock_cuda_set_last_error(int e) { t_last_error = e; }
cudaError_t cudaGetLastError(void) {
    const int e = t_last_error;
    t_last_error = cudaSuccess;
    return e;
}
const char* cudaGetErrorString(cudaError_t e) {
    switch (e) {
        case cudaSuccess: return "no error";
        case cudaErrorInvalidValue: return "invalid argument";
        case cudaErrorMemoryAllocation: return "out of memory";
        case cudaErrorInvalidDevice: return "invalid device ordinal";
        case cudaErrorHostMemoryAlreadyRegistered: return "part or all of the requested memory range is already mapped";
        case cudaErrorHostMemoryNotRegistered: return "pointer does not correspond to a registered memory region";
        case cudaErrorNotSupported: return "operation not supported";
        default: return "mock cuda error";
    }
}
void mock_cuda_set_register_supported(int on) { g_register_supported.store(on); }
void mock_cuda_counters(uint64_t out[6]) {
    std::lock_guard<std::mutex> lk(g_mu);
    out[0] = g_memcpy_calls.load(), out[1] = g_memcpy_bytes.load(), out[2] = g_registered.size(), out[3] = g_register_calls.load();
    out[4] = g_dev.size(), out[5] = g_pinned.size();
}
}

// ---- gds.h stand-in: no cuFile on the mock runtime; disk-tier blocks always take the pinned ring
#include "../../curvine_b200/csrc/host/gds.h"
namespace cv {
const GdsInfo& gds_info() {
    static GdsInfo g;
    g.detail = "mock runtime: no cuFile";
    return g;
}
Err gds_read(const std::string&, void*, int64_t, int64_t) { return Err(kUnsupported, "mock runtime: no cuFile"); }
void gds_forget(const std::string&) {}
std::string gds_last_refusal() { return ""; }
}  // namespace cv
