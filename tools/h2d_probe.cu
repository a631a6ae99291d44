// H2D ingest micro-probe: what limits pinned->HBM throughput on this box?
//   mode A: T threads, each cudaMemcpyAsync(4 MiB) from its own pinned slots, own stream / one shared stream
//   mode B: + pread of a tmpfs file into the slot before each copy
//   mode C: copies straight from cudaHostRegister'ed mmaps of the tmpfs files
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#define CK(x) do{cudaError_t e_=(x); if(e_!=cudaSuccess){printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} }while(0)
static double now(){ return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void bind_node0(){ cpu_set_t s; CPU_ZERO(&s); for(int c=0;c<32;c++){CPU_SET(c,&s); CPU_SET(c+64,&s);} sched_setaffinity(0,sizeof(s),&s); }
int main(int argc, char** argv){
  const size_t BS = 4<<20; const int NB = argc>1? atoi(argv[1]) : 2048; // 8 GiB
  const char* dir = argc>2? argv[2] : "/dev/shm/h2dprobe";
  CK(cudaSetDevice(0));
  uint8_t* dv; CK(cudaMalloc(&dv, (size_t)NB*BS));
  // files
  std::string d(dir); mkdir(d.c_str(), 0755);
  { std::vector<std::thread> ts; std::atomic<int> nx{0};
    for(int t=0;t<32;t++) ts.emplace_back([&]{ std::vector<uint8_t> b(BS); for(;;){ int i=nx++; if(i>=NB) break; memset(b.data(), i&255, BS); b[7]=i>>8; std::string p=d+"/b"+std::to_string(i); int fd=open(p.c_str(),O_WRONLY|O_CREAT|O_TRUNC,0644); if(write(fd,b.data(),BS)!=(ssize_t)BS) abort(); close(fd);} });
    for(auto&t:ts) t.join(); }
  const int SL = 64; uint8_t* ring; CK(cudaHostAlloc(&ring, (size_t)SL*BS, cudaHostAllocDefault)); memset(ring,1,(size_t)SL*BS);
  auto run = [&](const char* name, int T, bool shared_stream, int mode, int group){
    std::vector<cudaStream_t> ss(T); for(auto&s:ss) CK(cudaStreamCreateWithFlags(&s,cudaStreamNonBlocking));
    std::vector<cudaEvent_t> ev(SL); for(auto&e:ev) CK(cudaEventCreateWithFlags(&e,cudaEventDisableTiming));
    std::atomic<int> nx{0}; double t0=now();
    std::vector<std::thread> ts;
    for(int t=0;t<T;t++) ts.emplace_back([&,t]{ bind_node0(); cudaSetDevice(0); cudaStream_t s = shared_stream? ss[0]:ss[t];
      int slot_per = SL/T; int k=0;
      for(;;){ int i=nx.fetch_add(group); if(i>=NB) break; int n = std::min(group, NB-i);
        int slot = t*slot_per + (k++ % (slot_per/group>0?slot_per/group:1))*group; if(slot+n>SL) slot=0;
        CK(cudaEventSynchronize(ev[slot]));
        if(mode==1){ for(int j=0;j<n;j++){ std::string p=d+"/b"+std::to_string(i+j); int fd=open(p.c_str(),O_RDONLY); size_t got=0; while(got<BS){ ssize_t r=pread(fd, ring+(size_t)(slot+j)*BS+got, BS-got, got); if(r<=0) abort(); got+=r;} close(fd);} }
        CK(cudaMemcpyAsync(dv+(size_t)i*BS, ring+(size_t)slot*BS, (size_t)n*BS, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(ev[slot], s)); }
    });
    for(auto&t:ts) t.join(); CK(cudaDeviceSynchronize()); double dt=now()-t0;
    printf("%-34s T=%2d shared=%d group=%2d : %6.2f GB/s\n", name, T, (int)shared_stream, group, (double)NB*BS/dt/1e9);
    for(auto&s:ss) cudaStreamDestroy(s); for(auto&e:ev) cudaEventDestroy(e);
  };
  for(int rep=0;rep<2;rep++) run("A pinned ring, copy only", 1, false, 0, 1);
  run("A pinned ring, copy only", 16, false, 0, 1);
  run("A pinned ring, copy only", 16, true, 0, 1);
  run("A pinned ring, copy only", 32, false, 0, 1);
  run("A pinned ring, copy only", 4, false, 0, 4);
  run("A pinned ring, copy only", 16, false, 0, 4);
  run("B pread tmpfs -> ring -> copy", 16, false, 1, 1);
  run("B pread tmpfs -> ring -> copy", 16, true, 1, 1);
  run("B pread tmpfs -> ring -> copy", 12, false, 1, 1);
  run("B pread tmpfs -> ring -> copy", 20, false, 1, 1);
  run("B pread tmpfs -> ring -> copy", 16, false, 1, 2);
  // pread only (no GPU): CPU-side ceiling
  for(int T : {8,16,32,64}){ std::atomic<int> nx{0}; double t0=now(); std::vector<std::thread> ts;
    for(int t=0;t<T;t++) ts.emplace_back([&,t]{ bind_node0(); uint8_t* buf = ring + (size_t)(t%SL)*BS; for(;;){ int i=nx++; if(i>=NB) break; std::string p=d+"/b"+std::to_string(i); int fd=open(p.c_str(),O_RDONLY); size_t got=0; while(got<BS){ ssize_t r=pread(fd,buf+got,BS-got,got); if(r<=0) abort(); got+=r;} close(fd);} });
    for(auto&t:ts) t.join(); printf("pread only (node0-bound) T=%2d : %6.2f GB/s\n", T, (double)NB*BS/(now()-t0)/1e9); }
  // C: registered mmaps
  { std::vector<uint8_t*> maps(NB); double t0=now(); std::atomic<int> nx{0}; std::vector<std::thread> ts;
    for(int t=0;t<16;t++) ts.emplace_back([&]{ cudaSetDevice(0); for(;;){ int i=nx++; if(i>=NB) break; std::string p=d+"/b"+std::to_string(i); int fd=open(p.c_str(),O_RDWR); void* m=mmap(nullptr,BS,PROT_READ|PROT_WRITE,MAP_SHARED|MAP_POPULATE,fd,0); close(fd); cudaError_t e=cudaHostRegister(m,BS,cudaHostRegisterDefault); if(e!=cudaSuccess){ printf("register failed: %s\n", cudaGetErrorString(e)); exit(1);} maps[i]=(uint8_t*)m; } });
    for(auto&t:ts) t.join(); printf("C register %d x 4 MiB mmaps with 16 threads: %.2f s (%.2f GB/s)\n", NB, now()-t0, (double)NB*BS/(now()-t0)/1e9);
    for(int T : {1,4,16}) for(int rep=0; rep<2; rep++){ std::vector<cudaStream_t> ss(T); for(auto&s:ss) CK(cudaStreamCreateWithFlags(&s,cudaStreamNonBlocking)); std::atomic<int> n2{0}; double t1=now(); std::vector<std::thread> t2;
      for(int t=0;t<T;t++) t2.emplace_back([&,t]{ cudaSetDevice(0); for(;;){ int i=n2++; if(i>=NB) break; CK(cudaMemcpyAsync(dv+(size_t)i*BS, maps[i], BS, cudaMemcpyHostToDevice, ss[t])); } });
      for(auto&t:t2) t.join(); CK(cudaDeviceSynchronize()); printf("C copy from registered mmaps T=%2d : %6.2f GB/s\n", T, (double)NB*BS/(now()-t1)/1e9); for(auto&s:ss) cudaStreamDestroy(s); }
  }
  { // D: groups of 4 files mapped back to back, one registration, 16 MiB copies
    const int G=4; int NGp=NB/G; std::vector<uint8_t*> maps(NGp); double t0=now(); std::atomic<int> nx{0}; std::vector<std::thread> ts;
    for(int t=0;t<16;t++) ts.emplace_back([&]{ cudaSetDevice(0); for(;;){ int g=nx++; if(g>=NGp) break; uint8_t* base=(uint8_t*)mmap(nullptr,G*BS,PROT_NONE,MAP_PRIVATE|MAP_ANONYMOUS|MAP_NORESERVE,-1,0);
        for(int j=0;j<G;j++){ std::string p=d+"/b"+std::to_string(g*G+j); int fd=open(p.c_str(),O_RDWR); void* m=mmap(base+j*BS,BS,PROT_READ|PROT_WRITE,MAP_SHARED|MAP_FIXED|MAP_POPULATE,fd,0); close(fd); if(m==MAP_FAILED) abort(); }
        cudaError_t e=cudaHostRegister(base,G*BS,cudaHostRegisterDefault); if(e!=cudaSuccess){ printf("group register failed: %s\n", cudaGetErrorString(e)); exit(1);} maps[g]=base; } });
    for(auto&t:ts) t.join(); printf("D register %d x 16 MiB grouped mmaps with 16 threads: %.2f s (%.2f GB/s)\n", NGp, now()-t0, (double)NB*BS/(now()-t0)/1e9);
    for(int T : {1,8}) for(int rep=0; rep<2; rep++){ std::vector<cudaStream_t> ss(T); for(auto&s:ss) CK(cudaStreamCreateWithFlags(&s,cudaStreamNonBlocking)); std::atomic<int> n2{0}; double t1=now(); std::vector<std::thread> t2;
      for(int t=0;t<T;t++) t2.emplace_back([&,t]{ cudaSetDevice(0); for(;;){ int g=n2++; if(g>=NGp) break; CK(cudaMemcpyAsync(dv+(size_t)g*G*BS, maps[g], G*BS, cudaMemcpyHostToDevice, ss[t])); } });
      for(auto&t:t2) t.join(); CK(cudaDeviceSynchronize()); printf("D copy from grouped registered mmaps T=%2d : %6.2f GB/s\n", T, (double)NB*BS/(now()-t1)/1e9); for(auto&s:ss) cudaStreamDestroy(s); }
  }
  std::string rm = "rm -rf " + d; if(system(rm.c_str())){}
  return 0;
}
