#include <cstdio>
#include <cuda_runtime.h>
#include <chrono>
static double now(){ return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(){
  const size_t N=1<<19, S=4096; char *h,*d; cudaHostAlloc(&h,N*S,cudaHostAllocDefault); cudaMalloc(&d,N*S); memset(h,1,N*S);
  cudaStream_t s1,s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  auto t=[&](const char* name, auto f){ f(); cudaDeviceSynchronize(); double t0=now(); f(); cudaDeviceSynchronize(); double dt=now()-t0; printf("%-40s %.1f ms  %.1f GB/s\n", name, dt*1e3, N*S/dt/1e9); };
  t("H2D 1D 2GiB", [&]{ cudaMemcpyAsync(d,h,N*S,cudaMemcpyHostToDevice,s1); });
  t("D2H 1D 2GiB", [&]{ cudaMemcpyAsync(h,d,N*S,cudaMemcpyDeviceToHost,s1); });
  t("H2D+D2H concurrent 1D (each 1GiB)", [&]{ cudaMemcpyAsync(d,h,N*S/2,cudaMemcpyHostToDevice,s1); cudaMemcpyAsync(h+N*S/2,d+N*S/2,N*S/2,cudaMemcpyDeviceToHost,s2); });
  for (size_t w : {256, 512, 1024, 2048}) { char nm[64]; snprintf(nm,64,"H2D 2D width %zu B all columns", w);
    t(nm, [&]{ for(size_t c=0;c<S;c+=w) cudaMemcpy2DAsync(d+c,S,h+c,S,w,N,cudaMemcpyHostToDevice,s1); }); }
  for (size_t w : {512, 1024}) { char nm[64]; snprintf(nm,64,"D2H 2D width %zu B all columns", w);
    t(nm, [&]{ for(size_t c=0;c<S;c+=w) cudaMemcpy2DAsync(h+c,S,d+c,S,w,N,cudaMemcpyDeviceToHost,s1); }); }
  t("2D w1024 H2D(s1) || D2H(s2) halves", [&]{ for(size_t c=0;c<S/2;c+=1024){ cudaMemcpy2DAsync(d+c,S,h+c,S,1024,N,cudaMemcpyHostToDevice,s1); cudaMemcpy2DAsync(h+S/2+c,S,d+S/2+c,S,1024,N,cudaMemcpyDeviceToHost,s2);} });
  return 0; }
