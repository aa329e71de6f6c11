#include <cstdint>
#define P 0xFFF00001u
#define CC 0x000FFFFFu
__device__ __forceinline__ uint32_t addfix(uint32_t a, uint32_t v){
    uint32_t s;
    asm("{\n\t.reg .pred q;\n\t.reg .u32 c;\n\tadd.cc.u32 %0, %1, %2;\n\taddc.u32 c, 0, 0;\n\tsetp.ne.u32 q, c, 0;\n\t@q add.u32 %0, %0, 0xFFFFF;\n\t}" : "=r"(s) : "r"(a), "r"(v));
    return s;
}
__device__ __forceinline__ uint32_t subfix(uint32_t a, uint32_t v){
    uint32_t s;
    asm("{\n\t.reg .pred q;\n\t.reg .u32 c;\n\tsub.cc.u32 %0, %1, %2;\n\tsubc.u32 c, 0, 0;\n\tsetp.ne.u32 q, c, 0;\n\t@q sub.u32 %0, %0, 0xFFFFF;\n\t}" : "=r"(s) : "r"(a), "r"(v));
    return s;
}
#if VARIANT==0
// Barrett, plain C
__device__ __forceinline__ uint32_t mulw(uint32_t b, uint4 w, uint32_t z){
    uint32_t t = __umulhi(b, w.z);
    uint64_t Q = (uint64_t)b * w.y + t;
    return (uint32_t)(Q>>32) * CC + b * w.x;
}
#elif VARIANT==1
// Barrett with opaque zero high half
__device__ __forceinline__ uint32_t mulw(uint32_t b, uint4 w, uint32_t z){
    uint32_t t = __umulhi(b, w.z);
    uint64_t c64 = ((uint64_t)z << 32) | t;
    uint64_t Q = (uint64_t)b * w.y + c64;
    return (uint32_t)(Q>>32) * CC + b * w.x;
}
#elif VARIANT==2
// Montgomery' : w.x = w_m, w.y = w_m * Pinv mod 2^32  ; canonical result
__device__ __forceinline__ uint32_t mulw(uint32_t b, uint4 w, uint32_t z){
    uint32_t m = b * w.y;
    uint64_t Z = (uint64_t)b * w.x;
    uint32_t h = (uint32_t)(((uint64_t)m * CC + Z) >> 32);
    return subfix(h, m);   // h - m, +P if borrow  (subfix subtracts C on borrow == +P)
}
#endif
extern "C" __global__ void __launch_bounds__(256) k(uint32_t* __restrict__ out, const uint4* __restrict__ tw, int iters){
    const int NB=8;
    uint32_t a[NB], b[NB]; uint4 w[NB];
    uint32_t z; asm volatile("mov.u32 %0, 0;" : "=r"(z));
    #pragma unroll
    for(int i=0;i<NB;i++){ a[i]=out[threadIdx.x+i*256]; b[i]=out[threadIdx.x+i*256+4096]; w[i]=tw[(threadIdx.x>>2)+i*64]; }
    for(int it=0;it<iters;it++){
        #pragma unroll
        for(int i=0;i<NB;i++){
            uint32_t v = mulw(b[i], w[i], z);
            uint32_t s = addfix(a[i], v), d = subfix(a[i], v);
            a[i]=s; b[i]=d;
        }
        uint32_t t0=a[0];
        #pragma unroll
        for(int i=0;i<NB-1;i++) a[i]=a[i+1];
        a[NB-1]=t0;
    }
    #pragma unroll
    for(int i=0;i<NB;i++){ out[threadIdx.x+i*256]=a[i]; out[threadIdx.x+i*256+4096]=b[i]; }
}
