#include "../fastecc_b200/csrc/gf.cuh"
extern "C" __global__ void k(uint32_t* out, const uint4* tw){
    uint32_t a[8], b[8];
    for(int i=0;i<8;i++){ a[i]=out[threadIdx.x+i*32]; b[i]=out[threadIdx.x+i*32+256]; }
    uint4 w = tw[threadIdx.x>>2];
    #pragma unroll
    for(int i=0;i<8;i++){ uint32_t v = gf::mul(b[i], w.x, w.y, w.z); uint32_t s=gf::addl(a[i],v), d=gf::subl(a[i],v); a[i]=s; b[i]=d; }
    for(int i=0;i<8;i++){ out[threadIdx.x+i*32]=a[i]; out[threadIdx.x+i*32+256]=b[i]; }
}
