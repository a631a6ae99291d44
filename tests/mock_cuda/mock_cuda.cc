// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory).
#include "cuda_runtime.h"

#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <vector>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <set>
#include <thread>

namespace {
std::mutex g_mu;
std::map<uintptr_t, size_t> g_registered;   // cudaHostRegister'ed ranges
std::map<uintptr_t, size_t> g_dev, g_pinned;  // live allocations
std::atomic<uint64_t> g_memcpy_calls{0}, g_memcpy_bytes{0}, g_register_calls{0};
std::atomic<int> g_register_supported{1};
thread_local int t_device = 0;

// ---- streams and events.  Default: everything completes at enqueue time (the host logic is what is tested).  With
// MOCK_CUDA_ASYNC=1 every created stream is a FIFO with its own thread, operations run when their turn comes -- after a random pause of
// up to MOCK_CUDA_JITTER_US microseconds (copies pause longest) -- and the only ordering between streams is what events establish, as on
// the device.  A dependency the pipeline forgot (a kernel that consumes a buffer whose copy was issued on another stream, a result read
// before its D2H landed, a workspace freed under a running kernel) then shows up as wrong bytes in the parity tests instead of passing
// by luck.  The NULL stream stays synchronous (its operations run on the calling thread, after which they are complete): the tests read
// "device" memory directly after calls on it.  Asynchronous mode needs real launchers behind cvk_* that enqueue (tests/simt_emu); the
// plain-loop stand-ins of mock_cvk.cc run at call time.
struct MockStreamImpl {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    uint64_t enq = 0, done = 0;
    bool stop = false;
    std::thread th;
    std::minstd_rand rng{12345};
    void loop();
};
struct MockEventImpl {
    std::mutex mu;
    std::condition_variable cv;
    uint64_t recorded = 0, completed = 0;
};
typedef std::shared_ptr<MockEventImpl> EventRef;  // operations in flight keep their event alive past cudaEventDestroy

const bool g_async = [] { const char* e = getenv("MOCK_CUDA_ASYNC"); return e && atoi(e) != 0; }();
const int g_jitter_us = [] { const char* e = getenv("MOCK_CUDA_JITTER_US"); return e ? atoi(e) : 300; }();
std::mutex g_streams_mu;
std::map<MockStreamImpl*, std::shared_ptr<MockStreamImpl>> g_streams;  // owners: a device-wide sync may still hold a stream that is being destroyed
thread_local int t_op_weight = 1;  // set by the enqueuing call: copies pause longer than markers

void MockStreamImpl::loop() {
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
        cv.wait(lk, [&] { return stop || !q.empty(); });
        if (q.empty()) return;  // stop requested and everything queued has run
        std::function<void()> fn = std::move(q.front());
        q.pop_front();
        const int pause = g_jitter_us > 0 ? static_cast<int>(rng() % static_cast<unsigned>(g_jitter_us + 1)) : 0;
        lk.unlock();
        if (pause) std::this_thread::sleep_for(std::chrono::microseconds(pause));
        fn();
        lk.lock();
        done++;
        cv.notify_all();
    }
}
void enqueue(cudaStream_t st, std::function<void()> fn) {
    MockStreamImpl* s = reinterpret_cast<MockStreamImpl*>(st);
    if (!g_async || !s) {  // synchronous mode, or the NULL stream
        fn();
        return;
    }
    std::lock_guard<std::mutex> lk(s->mu);
    s->q.push_back(std::move(fn));
    s->enq++;
    s->cv.notify_all();
}
void sync_stream(MockStreamImpl* s) {
    if (!g_async || !s) return;
    std::unique_lock<std::mutex> lk(s->mu);
    const uint64_t target = s->enq;
    s->cv.wait(lk, [&] { return s->done >= target; });
}
void sync_device() {
    if (!g_async) return;
    std::vector<std::shared_ptr<MockStreamImpl>> all;
    {
        std::lock_guard<std::mutex> lk(g_streams_mu);
        for (auto& kv : g_streams) all.push_back(kv.second);
    }
    for (auto& s : all) sync_stream(s.get());
}
bool host_pinned(const void* p) {  // pinned allocation or registered range
    const uintptr_t x = reinterpret_cast<uintptr_t>(p);
    std::lock_guard<std::mutex> lk(g_mu);
    auto inside = [&](const std::map<uintptr_t, size_t>& m) {
        auto it = m.upper_bound(x);
        if (it == m.begin()) return false;
        --it;
        return x < it->first + it->second;
    };
    return inside(g_pinned) || inside(g_registered);
}
}  // namespace

// the SIMT shim's launches go through here (tests/simt_emu/simt_emu.cc): fn(arg) runs in stream order
extern "C" void mock_cuda_enqueue(cudaStream_t st, void (*fn)(void*), void* arg) {
    enqueue(st, [fn, arg] { fn(arg); });
}
extern "C" int mock_cuda_is_async(void) { return g_async ? 1 : 0; }

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
cudaError_t cudaFree(void* p) {  // synchronises the device, as the real call does
    sync_device();
    return free_from(g_dev, p);
}
cudaError_t cudaFreeAsync(void* p, cudaStream_t st) {  // stream-ordered: the memory goes away when the stream gets there
    if (!g_async || !st) return free_from(g_dev, p);
    enqueue(st, [p] { free_from(g_dev, p); });
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void* p) {
    sync_device();
    return free_from(g_pinned, p);
}

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
    sync_device();
    std::lock_guard<std::mutex> lk(g_mu);
    return g_registered.erase(reinterpret_cast<uintptr_t>(p)) ? cudaSuccess : cudaErrorHostMemoryNotRegistered;
}

cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, enum cudaMemcpyKind) {
    g_memcpy_calls++, g_memcpy_bytes += n;
    memmove(dst, src, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, enum cudaMemcpyKind k, cudaStream_t st) {
    if (!g_async || !st) return cudaMemcpy(dst, src, n, k);
    g_memcpy_calls++, g_memcpy_bytes += n;
    const bool h2d = k == cudaMemcpyHostToDevice || (k == cudaMemcpyDefault && !host_pinned(dst));
    const bool d2h = k == cudaMemcpyDeviceToHost;
    if (h2d && !host_pinned(src)) {  // pageable source: staged before the call returns, copied in stream order
        std::shared_ptr<std::vector<uint8_t>> stage = std::make_shared<std::vector<uint8_t>>(static_cast<const uint8_t*>(src), static_cast<const uint8_t*>(src) + n);
        enqueue(st, [dst, stage] { memcpy(dst, stage->data(), stage->size()); });
        return cudaSuccess;
    }
    enqueue(st, [dst, src, n] { memmove(dst, src, n); });
    if (d2h && !host_pinned(dst)) sync_stream(reinterpret_cast<MockStreamImpl*>(st));  // pageable destination: returns when the copy is done
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t st) {
    enqueue(st, [dst, v, n] { memset(dst, v, n); });
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
cudaError_t cudaDeviceSynchronize(void) {
    sync_device();
    return cudaSuccess;
}
cudaError_t cudaDeviceGetPCIBusId(char* buf, int len, int) {
    strncpy(buf, "0000:00:00.0", static_cast<size_t>(len));  // no such sysfs node: the ingest falls back to "no NUMA binding"
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned) {
    std::shared_ptr<MockStreamImpl> ref = std::make_shared<MockStreamImpl>();
    MockStreamImpl* s = ref.get();
    if (g_async) {
        s->rng.seed(static_cast<unsigned>(reinterpret_cast<uintptr_t>(s) >> 4) | 1u);
        s->th = std::thread([s] { s->loop(); });
    }
    std::lock_guard<std::mutex> lk(g_streams_mu);
    g_streams[s] = std::move(ref);
    *st = reinterpret_cast<cudaStream_t>(s);
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t st) {  // the real call returns at once and the stream goes away when its work is done: same result
    MockStreamImpl* s = reinterpret_cast<MockStreamImpl*>(st);
    if (!s) return cudaErrorInvalidValue;
    std::shared_ptr<MockStreamImpl> ref;
    {
        std::lock_guard<std::mutex> lk(g_streams_mu);
        auto it = g_streams.find(s);
        if (it == g_streams.end()) return cudaErrorInvalidValue;
        ref = std::move(it->second);
        g_streams.erase(it);
    }
    if (g_async) {
        {
            std::lock_guard<std::mutex> lk(s->mu);
            s->stop = true;
            s->cv.notify_all();
        }
        s->th.join();
    }
    return cudaSuccess;  // the object goes when the last holder (possibly a device-wide sync in another thread) lets go
}
cudaError_t cudaStreamSynchronize(cudaStream_t st) {
    sync_stream(reinterpret_cast<MockStreamImpl*>(st));
    return cudaSuccess;
}
static EventRef event_ref(cudaEvent_t ev) { return *reinterpret_cast<EventRef*>(ev); }
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t ev, unsigned) {
    EventRef e = event_ref(ev);
    uint64_t target;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        target = e->recorded;  // the most recent record at the time of this call; a later record does not move the wait
    }
    if (!target) return cudaSuccess;
    enqueue(st, [e, target] {
        std::unique_lock<std::mutex> lk(e->mu);
        e->cv.wait(lk, [&] { return e->completed >= target; });
    });
    return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* ev, unsigned) {
    *ev = reinterpret_cast<cudaEvent_t>(new EventRef(std::make_shared<MockEventImpl>()));
    return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t ev) {
    delete reinterpret_cast<EventRef*>(ev);
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t ev, cudaStream_t st) {
    EventRef e = event_ref(ev);
    uint64_t gen;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        gen = ++e->recorded;
    }
    enqueue(st, [e, gen] {
        std::lock_guard<std::mutex> lk(e->mu);
        if (e->completed < gen) e->completed = gen;
        e->cv.notify_all();
    });
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t ev) {
    EventRef e = event_ref(ev);
    std::unique_lock<std::mutex> lk(e->mu);
    const uint64_t target = e->recorded;
    e->cv.wait(lk, [&] { return e->completed >= target; });
    return cudaSuccess;
}
cudaError_t cudaEventQuery(cudaEvent_t ev) {
    EventRef e = event_ref(ev);
    std::lock_guard<std::mutex> lk(e->mu);
    return e->completed >= e->recorded ? cudaSuccess : cudaErrorNotReady;
}
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
