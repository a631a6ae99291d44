// TEST INFRASTRUCTURE ONLY -- known-answer kernels for the SIMT shim itself (tests/test_simt_emu.py): warp collectives, block
// barriers, early-returning threads, dynamic and static shared memory, atomics, multi-block grids, the PTX forms the product
// kernels use.  `selftest deadlock` runs a kernel whose barrier half the block never reaches: the scheduler must report it.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void scan_kernel(const uint32_t* in, uint32_t* out) {  // inclusive scan of 1024 values: shuffles + shared hop
    __shared__ uint32_t sums[32];
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    uint32_t v = in[blockIdx.x * 1024 + t];
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= uint32_t(d)) v += u;
    }
    if (lane == 31) sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = sums[lane];
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= uint32_t(d)) s += u;
        }
        sums[lane] = s;
    }
    __syncthreads();
    out[blockIdx.x * 1024 + t] = v + (warp ? sums[warp - 1] : 0u);
}

__global__ void early_exit_kernel(uint32_t n, uint32_t* out, uint32_t* count) {  // odd lanes leave before the collectives
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (threadIdx.x & 1)) return;
    const uint32_t m = __ballot_sync(0xffffffffu, (threadIdx.x & 3) == 0);
    const uint32_t x = __shfl_xor_sync(0xffffffffu, i, 2);
    const uint32_t dn = __shfl_down_sync(0xffffffffu, i, 2);
    __syncthreads();  // only the threads that are still alive take part
    out[i] = (m & 0xffffu) ^ (x << 16) ^ dn;
    if ((threadIdx.x & 31) == 0) atomicAdd(count, __popc(m));
}

__global__ void smem_ptx_kernel(const uint4* in, uint4* out) {  // cp.async into dynamic shared memory, ld.shared back, st.global
    extern __shared__ uint32_t smem[];
    const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const uint32_t mine = base + threadIdx.x * 16u, other = base + ((threadIdx.x + 1) % blockDim.x) * 16u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(mine), "l"(in + blockIdx.x * blockDim.x + threadIdx.x) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(0) : "memory");
    __syncthreads();
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(other) : "memory");
    uint32_t w;
    asm("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(mine));
    r.x = __funnelshift_r(r.x, r.y, 8) ^ __byte_perm(w, 0, 0x4441);
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + blockIdx.x * blockDim.x + threadIdx.x), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
}

__global__ void divergent_barrier_kernel(uint32_t* out) {
    // the low half of every warp waits for the whole warp, the high half for the whole block: neither collective can complete
    if ((threadIdx.x & 31) < 16) __syncwarp();
    else __syncthreads();
    out[0] = 1;
}

__global__ void oob_kernel(const uint4* in, uint4* out, int what) {
    // what 1: thread 4 stores one vector past a 4-vector buffer; what 2: it loads a granule that lies entirely outside its buffer;
    // what 3: every thread reads the granule that holds the last 3 bytes of a 51-byte buffer (the rule the kernels rely on: fine)
    uint4 v = make_uint4(1, 2, 3, 4);
    if (what == 2 && threadIdx.x == 4) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + 5));
    if (what == 3) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + 3));
    if (what != 1 && threadIdx.x >= 4) return;
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + threadIdx.x), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void neighbour_kernel(uint32_t* out, int barrier) {
    // every thread publishes a value in shared memory and reads its neighbour's: needs a barrier in between.
    // barrier 0: none (a race); 1: __syncthreads; 2: __syncwarp (enough: the neighbour t^1 is in the same warp)
    __shared__ uint32_t sh[128];
    sh[threadIdx.x] = threadIdx.x * 3u;
    if (barrier == 1) __syncthreads();
    if (barrier == 2) __syncwarp();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sh[threadIdx.x ^ 1u];
}
__device__ __forceinline__ uint32_t block_sum_no_barrier(uint32_t v, uint32_t* sh, int barrier) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) sh[warp] = v;
    if (barrier) __syncthreads();
    uint32_t t = sh[lane];
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    __syncthreads();
    return t;
}
__global__ void __launch_bounds__(1024) block_sum_kernel(uint32_t* out, int barrier) {  // the shape of the product's block_sum_1024
    __shared__ uint32_t sh[32];
    const uint32_t t = block_sum_no_barrier(threadIdx.x, sh, barrier);
    if (threadIdx.x == 0) out[blockIdx.x] = t;
}
__global__ void global_race_kernel(uint32_t* out, int atomic) {  // two BLOCKS update one global counter: plain add is a race, atomicAdd is not
    if (threadIdx.x == 0) {
        if (atomic) atomicAdd(out, 1u);
        else out[0] = out[0] + 1u;
    }
}

int main(int argc, char** argv) {
    if (argc > 2 && argv[1][0] == 'r') {  // selftest race 0|1|2|3|4: meaningful in the ThreadSanitizer build
        const int k = atoi(argv[2]);
        uint32_t* out = static_cast<uint32_t*>(calloc(256, 4));
        if (k >= 5) block_sum_kernel<<<3, 1024, 0, (cudaStream_t)0>>>(out, k == 6);
        else if (k <= 2) neighbour_kernel<<<2, 128, 0, (cudaStream_t)0>>>(out, k);
        else global_race_kernel<<<64, 32, 0, (cudaStream_t)0>>>(out, k == 4);
        printf("race kernel done %u\n", out[0]);
        return 0;
    }
    if (argc > 2 && argv[1][0] == 'o') {  // selftest oob 1|2|3: meaningful in the AddressSanitizer build
        uint4 *in = nullptr, *out = nullptr;
        if (posix_memalign(reinterpret_cast<void**>(&in), 16, 51) || posix_memalign(reinterpret_cast<void**>(&out), 16, 64)) return 2;
        memset(in, 7, 51);
        oob_kernel<<<1, 5, 0, (cudaStream_t)0>>>(in, out, atoi(argv[2]));
        printf("oob kernel done\n");
        return 0;
    }
    if (argc > 1 && argv[1][0] == 'd') {
        uint32_t* z = static_cast<uint32_t*>(calloc(4, 1));
        divergent_barrier_kernel<<<1, 128, 0, (cudaStream_t)0>>>(z);
        printf("not reached\n");
        return 0;
    }
    int bad = 0;
    {   // scan: 5 blocks
        const uint32_t nb = 5;
        uint32_t* in = new uint32_t[nb * 1024];
        uint32_t* out = new uint32_t[nb * 1024];
        for (uint32_t i = 0; i < nb * 1024; i++) in[i] = (i * 2654435761u) >> 20;
        scan_kernel<<<nb, 1024, 0, (cudaStream_t)0>>>(in, out);
        for (uint32_t b = 0; b < nb; b++) {
            uint32_t acc = 0;
            for (uint32_t t = 0; t < 1024; t++) {
                acc += in[b * 1024 + t];
                if (out[b * 1024 + t] != acc) bad++;
            }
        }
        printf("scan %s\n", bad ? "BAD" : "ok");
    }
    {   // early exit: n = 1000 over 4 blocks of 256 (the last block has threads beyond n)
        const uint32_t n = 1000;
        uint32_t* out = static_cast<uint32_t*>(calloc(1024, 4));
        uint32_t count = 0;
        early_exit_kernel<<<4, 256, 0, (cudaStream_t)0>>>(n, out, &count);
        uint32_t want_count = 0;
        int b2 = 0;
        for (uint32_t i = 0; i < 1024; i++) {
            const uint32_t t = i & 255, lane = t & 31, w0 = i - lane;
            uint32_t live = 0;  // even lanes below n
            for (uint32_t l = 0; l < 32; l += 2) if (w0 + l < n) live |= 1u << l;
            uint32_t want = 0;
            if (i < n && !(t & 1)) {
                uint32_t m = 0;
                for (uint32_t l = 0; l < 32; l += 4) if (live >> l & 1) m |= 1u << l;
                const uint32_t xl = lane ^ 2, dl = lane + 2;
                const uint32_t x = (live >> xl & 1) ? w0 + xl : i;  // a lane that has left contributes nothing: own value (the shim's rule)
                const uint32_t dn = dl < 32 && (live >> dl & 1) ? w0 + dl : i;
                want = (m & 0xffffu) ^ (x << 16) ^ dn;
                if (lane == 0) want_count += __builtin_popcount(m);
            }
            if (out[i] != want) b2++;
        }
        if (count != want_count) b2++;
        printf("early_exit %s\n", b2 ? "BAD" : "ok");
        bad += b2;
    }
    {   // shared memory + PTX forms
        const uint32_t nb = 3, nt = 128;
        uint4* in = static_cast<uint4*>(aligned_alloc(16, nb * nt * 16));
        uint4* out = static_cast<uint4*>(aligned_alloc(16, nb * nt * 16));
        for (uint32_t i = 0; i < nb * nt; i++) in[i] = make_uint4(i * 7u + 1, i * 0x01010101u, ~i, i << 9);
        smem_ptx_kernel<<<nb, nt, nt * 16, (cudaStream_t)0>>>(in, out);
        int b3 = 0;
        for (uint32_t b = 0; b < nb; b++)
            for (uint32_t t = 0; t < nt; t++) {
                const uint4 o = in[b * nt + (t + 1) % nt], me = in[b * nt + t];
                const uint32_t x = static_cast<uint32_t>((((uint64_t)o.y << 32) | o.x) >> 8) ^ ((me.x >> 8) & 0xffu);
                const uint4 g = out[b * nt + t];
                if (g.x != x || g.y != o.y || g.z != o.z || g.w != o.w) b3++;
            }
        printf("smem_ptx %s\n", b3 ? "BAD" : "ok");
        bad += b3;
    }
    {   // invalid configurations are launch errors, not crashes
        scan_kernel<<<0, 1024, 0, (cudaStream_t)0>>>(nullptr, nullptr);
        const int e1 = cudaGetLastError(), e2 = cudaGetLastError();
        scan_kernel<<<1, 2048, 0, (cudaStream_t)0>>>(nullptr, nullptr);
        const int e3 = cudaGetLastError();
        printf("launch_errors %s\n", e1 == 9 && e2 == 0 && e3 == 9 ? "ok" : "BAD");
        bad += !(e1 == 9 && e2 == 0 && e3 == 9);
    }
    printf(bad ? "selftest FAILED\n" : "selftest ok\n");
    return bad ? 1 : 0;
}
