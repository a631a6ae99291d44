// Box probe: host facts + pinned H2D bandwidth (sets the PCIe-ingest denominator).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
int main(){
  int n=0; CK(cudaGetDeviceCount(&n)); printf("devices=%d\n", n);
  for(int d=0; d<n; d++){
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,d));
    printf("dev%d %s sm=%d smem_optin=%zu l2=%d mem=%zu pci=%04x:%02x:%02x\n", d,p.name,p.multiProcessorCount,p.sharedMemPerBlockOptin,p.l2CacheSize,p.totalGlobalMem,p.pciDomainID,p.pciBusID,p.pciDeviceID);
  }
  CK(cudaSetDevice(0));
  size_t sz = (size_t)1<<30;
  void *h, *dv; CK(cudaHostAlloc(&h, sz, cudaHostAllocDefault)); memset(h, 1, sz);
  CK(cudaMalloc(&dv, sz));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  cudaEvent_t a,b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  size_t chunks[] = {128<<10, 1<<20, 4<<20, 16<<20, 64<<20, (size_t)1<<30};
  for(size_t c : chunks){
    for(int rep=0; rep<2; rep++){
      CK(cudaEventRecord(a,s));
      for(size_t o=0;o<sz;o+=c) CK(cudaMemcpyAsync((char*)dv+o,(char*)h+o,c,cudaMemcpyHostToDevice,s));
      CK(cudaEventRecord(b,s)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms,a,b));
      if(rep) printf("H2D chunk=%zuKiB %.2f GB/s\n", c>>10, sz/ms/1e6);
    }
  }
  // D2H
  CK(cudaEventRecord(a,s)); CK(cudaMemcpyAsync(h,dv,sz,cudaMemcpyDeviceToHost,s)); CK(cudaEventRecord(b,s)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms,a,b)); printf("D2H 1GiB %.2f GB/s\n", sz/ms/1e6);
  // two streams concurrently H2D
  cudaStream_t s2; CK(cudaStreamCreate(&s2));
  CK(cudaEventRecord(a,s));
  for(size_t o=0;o<sz/2;o+=4<<20){ CK(cudaMemcpyAsync((char*)dv+o,(char*)h+o,4<<20,cudaMemcpyHostToDevice,s)); CK(cudaMemcpyAsync((char*)dv+sz/2+o,(char*)h+sz/2+o,4<<20,cudaMemcpyHostToDevice,s2)); }
  CK(cudaStreamSynchronize(s2)); CK(cudaEventRecord(b,s)); CK(cudaEventSynchronize(b));
  CK(cudaEventElapsedTime(&ms,a,b)); printf("H2D 2-stream 4MiB %.2f GB/s\n", sz/ms/1e6);
  // host memcpy bandwidth (single thread) pageable->pinned
  void* src = malloc(sz); memset(src,2,sz);
  struct timespec t0,t1; clock_gettime(CLOCK_MONOTONIC,&t0); memcpy(h,src,sz); clock_gettime(CLOCK_MONOTONIC,&t1);
  printf("host memcpy 1 thread %.2f GB/s\n", sz/((t1.tv_sec-t0.tv_sec)+(t1.tv_nsec-t0.tv_nsec)*1e-9)/1e9);
  // cudaHostRegister cost
  clock_gettime(CLOCK_MONOTONIC,&t0); CK(cudaHostRegister(src, sz, cudaHostRegisterDefault)); clock_gettime(CLOCK_MONOTONIC,&t1);
  printf("cudaHostRegister 1GiB %.1f ms\n", ((t1.tv_sec-t0.tv_sec)+(t1.tv_nsec-t0.tv_nsec)*1e-9)*1e3);
  return 0;
}
