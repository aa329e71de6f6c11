/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the GF(0xFFF00001) NTT / Reed-Solomon encode hot path.
 *
 * This is a plain-C restatement of the algorithm in the FastECC reference (GF(p).cpp, ntt.cpp, RS.cpp,
 * main.cpp:hash).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may call it.  The product path (fastecc_b200/csrc) never links or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 *   (1) the golden values recorded from the unmodified reference (the one published pin
 *       Benchmarks.md:491-507), and
 *   (2) oracle/_ref/libfastecc_ref.so, the reference's own templates compiled from /root/reference by
 *       oracle/Makefile, when that library is present.
 */
#ifndef GFP_ORACLE_H
#define GFP_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_P 0xFFF00001u

uint32_t oracle_gf_add (uint32_t x, uint32_t y);              /* GF(p).cpp:44-48  */
uint32_t oracle_gf_sub (uint32_t x, uint32_t y);              /* GF(p).cpp:37-42  */
uint32_t oracle_gf_mul (uint32_t x, uint32_t y);              /* GF(p).cpp:110-127 (result == (x*y) % P, main.cpp:105) */
uint32_t oracle_gf_mul32(uint32_t x, uint32_t y);             /* literal GF_Mul32 Barrett form, GF(p).cpp:110-122 */
uint32_t oracle_gf_pow (uint32_t x, uint32_t n);              /* GF(p).cpp:254-264 */
uint32_t oracle_gf_root(uint32_t n);                          /* GF(p).cpp:267-276: 19^((P-1)/n) */
uint32_t oracle_gf_inv (uint32_t x);                          /* GF(p).cpp:293-297: x^(P-2) */

/* main.cpp:203-212 -- rolling hash over blocks data[0..N-1], each `size_words` u32 words, block i at data + i*pitch_words */
uint32_t oracle_hash (const uint32_t *data, size_t N, size_t size_words, size_t pitch_words);

/* Slow_NTT, ntt.cpp:451-483 -- O(N^2) definitional DFT of every word column, flat [N][size] array, in place. */
void oracle_slow_ntt (uint32_t *data, size_t N, size_t size_words, int inverse);

/* Same result as MFA_NTT<uint32_t,0xFFF00001> (ntt.cpp:382-447) seen through data[i]: natural order in and out,
 * inverse is unnormalised.  Implemented as revbin_permute + IterativeNTT_Steps (ntt.cpp:251-318) on the flat array
 * (rows physically permuted instead of the pointer table).  N must be a power of two <= 2^20. Returns 0 or -1. */
int  oracle_ntt (uint32_t *data, size_t N, size_t size_words, int inverse);

/* Body of EncodeReedSolomon, RS.cpp:41-63: iNTT, scale row i by inv_N*root_2N^i, NTT.  In place: parity overwrites
 * data.  N power of two <= 2^19.  Returns 0 or -1. */
int  oracle_rs_encode (uint32_t *data, size_t N, size_t size_words);

/* Definition-level check used by the tests: parity[j] = f(root_2N^(2j+1)) where f is the
 * degree<N polynomial with f(root_2N^(2i)) = data[i]; O(N^2) per column.  out gets N*size words. */
void oracle_rs_encode_by_definition (const uint32_t *data, uint32_t *out, size_t N, size_t size_words);

/* fills used by the reference drivers / the survey goldens */
void oracle_fill_A (uint32_t *data, size_t nwords);   /* data0[i] = i % P         RS.cpp:28-29, main.cpp:249-250 */
void oracle_fill_B (uint32_t *data, size_t nwords);   /* LCG x=12345; x=x*1664525+1013904223; data0[i]=x%P */

int  oracle_num_threads(void);
#ifdef __cplusplus
}
#endif
#endif
