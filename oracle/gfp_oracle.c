/* TEST INFRASTRUCTURE ONLY -- see gfp_oracle.h.  Plain C (gcc -O3 -fopenmp). */
#include "gfp_oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define P ORACLE_P

/* GF(p).cpp:37-42 */
uint32_t oracle_gf_sub(uint32_t x, uint32_t y) { uint32_t r = x - y; return r + (r > x ? P : 0); }
/* GF(p).cpp:44-48 : GF_Sub(X, P-Y) */
uint32_t oracle_gf_add(uint32_t x, uint32_t y) { return oracle_gf_sub(x, P - y); }
/* value of every GF_Mul variant is (X*Y) mod P (main.cpp:105-106 is the reference's own check) */
uint32_t oracle_gf_mul(uint32_t x, uint32_t y) { return (uint32_t)(((uint64_t)x * y) % P); }

/* GF(p).cpp:110-122, scalar form, restated literally so the Barrett constants are pinned too */
uint32_t oracle_gf_mul32(uint32_t x, uint32_t y)
{
    const uint64_t est = (((uint64_t)1 << 63) / P) << 1;
    const uint32_t invP32 = (uint32_t)((uint64_t)(est * P) > (uint64_t)((est + 1) * P) ? est : est + 1);  /* 0x1000FF */
    uint64_t res = (uint64_t)x * y;
    res -= ((res + (res >> 32) * invP32) >> 32) * P;
    return (uint32_t)(res >= P ? res - P : res);
}

/* GF(p).cpp:254-264 */
uint32_t oracle_gf_pow(uint32_t x, uint32_t n)
{
    uint32_t r = 1;
    for (; n; n /= 2) { if (n & 1) r = oracle_gf_mul(r, x); x = oracle_gf_mul(x, x); }
    return r;
}
/* GF(p).cpp:267-276 */
uint32_t oracle_gf_root(uint32_t n) { return oracle_gf_pow(19, (P - 1) / n); }
/* GF(p).cpp:293-297 */
uint32_t oracle_gf_inv(uint32_t x) { return oracle_gf_pow(x, P - 2); }

/* main.cpp:203-212 */
uint32_t oracle_hash(const uint32_t *data, size_t N, size_t size_words, size_t pitch_words)
{
    uint32_t h = 314159253u;
    for (size_t i = 0; i < N; i++) {
        const uint32_t *p = data + i * pitch_words;
        for (size_t k = 0; k < size_words; k++) h = (h + p[k]) * 123456791u + (h >> 17);
    }
    return h;
}

void oracle_fill_A(uint32_t *d, size_t n) { for (size_t i = 0; i < n; i++) d[i] = (uint32_t)(i % P); }
void oracle_fill_B(uint32_t *d, size_t n) { uint32_t x = 12345u; for (size_t i = 0; i < n; i++) { x = x * 1664525u + 1013904223u; d[i] = x % P; } }

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ntt.cpp:451-483 */
void oracle_slow_ntt(uint32_t *data, size_t N, size_t S, int inverse)
{
    uint32_t *out = (uint32_t *)malloc(N * S * sizeof(uint32_t));
    uint32_t root = oracle_gf_root((uint32_t)N);
    if (inverse) root = oracle_gf_inv(root);
    uint32_t dw = 1;
    for (size_t i = 0; i < N; i++) {
        #pragma omp parallel for
        for (long k = 0; k < (long)S; k++) {
            uint32_t t = 0, w = 1;
            for (size_t x = 0; x < N; x++) {
                t = oracle_gf_add(t, oracle_gf_mul(w, data[x * S + k]));
                w = oracle_gf_mul(w, dw);
            }
            out[i * S + k] = t;
        }
        dw = oracle_gf_mul(dw, root);
    }
    memcpy(data, out, N * S * sizeof(uint32_t));
    free(out);
}

static int is_pow2(size_t n) { return n && !(n & (n - 1)); }

/* Column-strip version of revbin_permute + IterativeNTT_Steps (ntt.cpp:251-318).  Works on a strip of W<=64
 * word-columns copied into a contiguous scratch [N][W] so that big transforms stay cache friendly; the arithmetic
 * and the stage/twiddle order are exactly the reference's (root of order 2h per stage, running root_i product). */
static void ntt_strip(uint32_t *buf, size_t N, size_t W, const uint32_t *roots, int nroots)
{
    /* bit reversal of rows (ntt.cpp:292-309 does it on the pointer table) */
    for (size_t m = 1, mr = 0; m < N; m++) {
        size_t l = N;
        do { l >>= 1; } while (mr + l >= N);
        mr = (mr & (l - 1)) + l;
        if (mr > m)
            for (size_t k = 0; k < W; k++) { uint32_t t = buf[m * W + k]; buf[m * W + k] = buf[mr * W + k]; buf[mr * W + k] = t; }
    }
    const uint32_t *root_ptr = roots + nroots;
    for (size_t h = 1; h < N; h *= 2) {
        uint32_t root = *--root_ptr;                                   /* primitive root of order 2h */
        for (size_t x = 0; x < N; x += 2 * h) {
            uint32_t root_i = 1;
            for (size_t i = 0; i < h; i++) {
                uint32_t *b1 = buf + (x + i) * W, *b2 = buf + (x + i + h) * W;
                for (size_t k = 0; k < W; k++) {
                    uint32_t u = b1[k];
                    uint32_t v = (i == 0) ? b2[k] : oracle_gf_mul(b2[k], root_i);   /* ntt.cpp:262-267 / 274-279 */
                    b1[k] = oracle_gf_add(u, v);
                    b2[k] = oracle_gf_sub(u, v);
                }
                root_i = oracle_gf_mul(root_i, root);
            }
        }
    }
}

int oracle_ntt(uint32_t *data, size_t N, size_t S, int inverse)
{
    if (!is_pow2(N) || N > ((size_t)1 << 20) || S == 0) return -1;
    if (N == 1) return 0;
    /* ntt.cpp:397-402: roots[] = {w, w^2, w^4, ...} until 1 */
    uint32_t roots[66]; int nroots = 0;
    uint32_t root = oracle_gf_root((uint32_t)N);
    if (inverse) root = oracle_gf_inv(root);
    while (root != 1) { roots[nroots++] = root; root = oracle_gf_mul(root, root); }

    const size_t W = 16;
    size_t nstrips = (S + W - 1) / W;
    #pragma omp parallel
    {
        uint32_t *buf = (uint32_t *)malloc(N * W * sizeof(uint32_t));
        #pragma omp for schedule(dynamic)
        for (long s = 0; s < (long)nstrips; s++) {
            size_t k0 = (size_t)s * W, w = (k0 + W <= S) ? W : S - k0;
            for (size_t i = 0; i < N; i++) memcpy(buf + i * w, data + i * S + k0, w * sizeof(uint32_t));
            ntt_strip(buf, N, w, roots, nroots);
            for (size_t i = 0; i < N; i++) memcpy(data + i * S + k0, buf + i * w, w * sizeof(uint32_t));
        }
        free(buf);
    }
    return 0;
}

/* RS.cpp:41-63 */
int oracle_rs_encode(uint32_t *data, size_t N, size_t S)
{
    if (!is_pow2(N) || N > ((size_t)1 << 19) || S == 0) return -1;
    if (oracle_ntt(data, N, S, 1)) return -1;                                     /* RS.cpp:41 */
    uint32_t root_2N = oracle_gf_root((uint32_t)(2 * N)), inv_N = oracle_gf_inv((uint32_t)N);   /* RS.cpp:51 */
    #pragma omp parallel for
    for (long i = 0; i < (long)N; i++) {
        uint32_t root_i = oracle_gf_mul(inv_N, oracle_gf_pow(root_2N, (uint32_t)i));              /* RS.cpp:54 */
        uint32_t *b = data + (size_t)i * S;
        for (size_t k = 0; k < S; k++) b[k] = oracle_gf_mul(b[k], root_i);                        /* RS.cpp:56-58 */
    }
    return oracle_ntt(data, N, S, 0);                                             /* RS.cpp:63 */
}

/* Lagrange-free definitional check: coefficients by inverse DFT definition, then evaluation at odd powers. */
void oracle_rs_encode_by_definition(const uint32_t *data, uint32_t *out, size_t N, size_t S)
{
    uint32_t r2n = oracle_gf_root((uint32_t)(2 * N));
    uint32_t w = oracle_gf_mul(r2n, r2n), winv = oracle_gf_inv(w), invN = oracle_gf_inv((uint32_t)N);
    uint32_t *coef = (uint32_t *)malloc(N * sizeof(uint32_t));
    for (size_t k = 0; k < S; k++) {
        for (size_t m = 0; m < N; m++) {                         /* c[m] = 1/N sum_i d[i] w^(-i m) */
            uint32_t acc = 0, step = oracle_gf_pow(winv, (uint32_t)m), x = 1;
            for (size_t i = 0; i < N; i++) { acc = oracle_gf_add(acc, oracle_gf_mul(data[i * S + k], x)); x = oracle_gf_mul(x, step); }
            coef[m] = oracle_gf_mul(acc, invN);
        }
        for (size_t j = 0; j < N; j++) {                         /* parity[j] = f(r2n^(2j+1)) */
            uint32_t pt = oracle_gf_pow(r2n, (uint32_t)(2 * j + 1)), x = 1, acc = 0;
            for (size_t m = 0; m < N; m++) { acc = oracle_gf_add(acc, oracle_gf_mul(coef[m], x)); x = oracle_gf_mul(x, pt); }
            out[j * S + k] = acc;
        }
    }
    free(coef);
}
