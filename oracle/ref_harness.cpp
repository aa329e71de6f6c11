// TEST INFRASTRUCTURE ONLY.  Thin extern "C" harness around the UNMODIFIED FastECC reference templates.
// Compiled by oracle/Makefile with -I$(REF) from the sources where they lie (/root/reference); nothing from the
// reference is copied into this repository.  Include set and order follow RS.cpp:2-18.
#include <iostream>
#include <algorithm>
#include <stdint.h>
#include <cstddef>
#include <string.h>
#include <math.h>
#include <cassert>
#include <utility>
#include <functional>
#include <vector>
#include <memory>

#include "wall_clock_timer.h"
#include "LargePages.cpp"
#include "GF(p).cpp"
#include "ntt.cpp"

typedef uint32_t T;
static const T Pm = 0xFFF00001;

extern "C" {
uint32_t ref_gf_add (uint32_t x, uint32_t y) { return GF_Add<T,Pm>(x,y); }
uint32_t ref_gf_sub (uint32_t x, uint32_t y) { return GF_Sub<T,Pm>(x,y); }
uint32_t ref_gf_mul (uint32_t x, uint32_t y) { return GF_Mul<T,Pm>(x,y); }
uint32_t ref_gf_pow (uint32_t x, uint32_t n) { return GF_Pow<T,Pm>(x,n); }
uint32_t ref_gf_root(uint32_t n)             { return GF_Root<T,Pm>(n); }
uint32_t ref_gf_inv (uint32_t x)             { return GF_Inv<T,Pm>(x); }

// main.cpp:203-212 (restated: main.cpp has its own main() and cannot be included)
uint32_t ref_hash (uint32_t** data, size_t N, size_t SIZE)
{
    uint32_t hash = 314159253;
    for (size_t i=0; i<N; i++) { uint32_t* p = data[i]; for (size_t k=0; k<SIZE; k++) hash = (hash+p[k])*123456791 + (hash>>17); }
    return hash;
}
void ref_mfa_ntt  (uint32_t** data, size_t N, size_t SIZE, int inv) { MFA_NTT<T,Pm>(data, N, SIZE, inv!=0); }
void ref_slow_ntt (uint32_t*  data, size_t N, size_t SIZE, int inv) { Slow_NTT<T,Pm>(data, N, SIZE, inv!=0); }

// The timed body of EncodeReedSolomon, RS.cpp:41-63, executed on a caller-supplied pointer table.
void ref_rs_encode (uint32_t** data, size_t N, size_t SIZE)
{
    MFA_NTT<T,Pm> (data, N, SIZE, true);
    T root_2N = GF_Root<T,Pm>(2*N),  inv_N = GF_Inv<T,Pm>(N);
    #pragma omp parallel for
    for (ptrdiff_t i=0; i<(ptrdiff_t)N; i++) {
        T root_i = GF_Mul<T,Pm> (inv_N, GF_Pow<T,Pm>(root_2N,i));
        T* __restrict__ block = data[i];
        for (size_t k=0; k<SIZE; k++) block[k] = GF_Mul<T,Pm> (block[k], root_i);
    }
    MFA_NTT<T,Pm> (data, N, SIZE, false);
}

// Flat-buffer convenience wrappers: build the T** table like RS.cpp:31-33, run, then gather results through
// data[i] back into natural order (MFA_NTT leaves the pointer table permuted).
static void run_flat(uint32_t* flat, size_t N, size_t SIZE, int what, int inv)
{
    std::vector<uint32_t*> tab(N);
    for (size_t i=0;i<N;i++) tab[i] = flat + i*SIZE;
    if (what==0) ref_mfa_ntt(tab.data(), N, SIZE, inv); else ref_rs_encode(tab.data(), N, SIZE);
    bool moved=false; for (size_t i=0;i<N;i++) if (tab[i] != flat+i*SIZE) { moved=true; break; }
    if (moved) {
        uint32_t* tmp = (uint32_t*) malloc(N*SIZE*sizeof(uint32_t));
        #pragma omp parallel for
        for (ptrdiff_t i=0;i<(ptrdiff_t)N;i++) memcpy(tmp+i*SIZE, tab[i], SIZE*sizeof(uint32_t));
        memcpy(flat, tmp, N*SIZE*sizeof(uint32_t)); free(tmp);
    }
}
void ref_mfa_ntt_flat  (uint32_t* flat, size_t N, size_t SIZE, int inv) { run_flat(flat,N,SIZE,0,inv); }
void ref_rs_encode_flat(uint32_t* flat, size_t N, size_t SIZE)          { run_flat(flat,N,SIZE,1,0); }
void ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int  ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
const char* ref_build_flavour(void) {
#if SIMD==AVX2
    return "avx2";
#elif SIMD==SSE2
    return "sse2";
#else
    return "scalar";
#endif
}
}
