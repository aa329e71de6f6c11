#include <cstdint>
__device__ __forceinline__ uint32_t subfix(uint32_t a, uint32_t v){
    uint32_t s;
#if SV==0
    asm("{\n\t.reg .pred q;\n\t.reg .u32 c;\n\tsub.cc.u32 %0, %1, %2;\n\taddc.u32 c, 0, 0;\n\tsetp.ne.u32 q, c, 0;\n\t@q sub.u32 %0, %0, 0xFFFFF;\n\t}" : "=r"(s) : "r"(a), "r"(v));
#elif SV==1
    asm("{\n\t.reg .pred q;\n\t.reg .u32 c;\n\tsub.cc.u32 %0, %1, %2;\n\tsubc.u32 c, 0, 0;\n\tsetp.eq.u32 q, c, 0;\n\t@!q sub.u32 %0, %0, 0xFFFFF;\n\t}" : "=r"(s) : "r"(a), "r"(v));
#elif SV==2
    // a - v = a + ~v + 1 : two-step with explicit not
    asm("{\n\t.reg .pred q;\n\t.reg .u32 c, nv;\n\tnot.b32 nv, %2;\n\tadd.cc.u32 %0, %1, 1;\n\taddc.cc.u32 %0, %0, nv;\n\taddc.u32 c, 0, 0;\n\tsetp.eq.u32 q, c, 0;\n\t@q sub.u32 %0, %0, 0xFFFFF;\n\t}" : "=r"(s) : "r"(a), "r"(v));
#elif SV==3
    asm("{\n\t.reg .pred q;\n\tsub.u32 %0, %1, %2;\n\tsetp.lt.u32 q, %1, %2;\n\t@q sub.u32 %0, %0, 0xFFFFF;\n\t}" : "=r"(s) : "r"(a), "r"(v));
#elif SV==4
    asm("{\n\t.reg .pred q;\n\t.reg .u32 c;\n\tsub.cc.u32 %0, %1, %2;\n\tsubc.u32 c, 1, 0;\n\tsetp.eq.u32 q, c, 0;\n\t@q sub.u32 %0, %0, 0xFFFFF;\n\t}" : "=r"(s) : "r"(a), "r"(v));
#endif
    return s;
}
extern "C" __global__ void k(uint32_t* out){
    uint32_t a[8];
    for(int i=0;i<8;i++) a[i]=out[threadIdx.x+i*32];
    #pragma unroll
    for(int i=0;i<7;i++) a[i+1]=subfix(a[i+1],a[i]);
    for(int i=0;i<8;i++) out[threadIdx.x+i*32]=a[i];
}
