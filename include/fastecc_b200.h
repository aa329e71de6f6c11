/* fastecc_b200 -- C ABI of the B200 (sm_100a) NTT Reed-Solomon encoder.
 *
 * This is the drop-in boundary for the one hot path of Bulat-Ziganshin/FastECC.  The reference has no FFI layer:
 * its programs textually include the template sources, so the boundary is created at the narrowest existing call
 * surface.  Each entry point below names the reference interface it replaces (file:line in the reference tree).
 * All functions return 0 on success and a negative FASTECC_B200_E* code otherwise; fastecc_b200_last_error()
 * returns a human-readable message for the calling thread.  There is NO CPU fallback: without a usable CUDA
 * device every compute entry point fails with FASTECC_B200_ECUDA.
 *
 * Concurrency: the context is process-wide.  The host (T**) entry points are serialised by an internal lock (they share one
 * staging buffer and three internal streams).  The *_dev entry points are asynchronous on the caller's stream and may be
 * called from several threads / on several streams: the shared scratch buffers and the lazily built twiddle tables are
 * ordered across streams with events (two transforms that need the same scratch buffer run one after the other).
 *
 * Data model (same as the reference): N blocks of SIZE 32-bit words, words are elements of
 * GF(P), P = 0xFFF00001, the transform runs across blocks independently for every word column.  Inputs may be
 * any 32-bit values (they are taken mod P); outputs are canonical residues in [0, P).
 */
#ifndef FASTECC_B200_H
#define FASTECC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FASTECC_B200_P        0xFFF00001u
#define FASTECC_B200_MAX_LOG_N        20      /* P-1 = 2^20 * 4095: largest power-of-two transform (main.cpp:326) */
#define FASTECC_B200_MAX_LOG_N_ENCODE 19      /* encode needs a root of order 2N (RS.cpp:51, README.md:29)        */

enum {
    FASTECC_B200_OK       =  0,
    FASTECC_B200_EINVAL   = -1,   /* N not a power of two / out of range, SIZE == 0, bad pitch, null pointer */
    FASTECC_B200_ECUDA    = -2,   /* CUDA runtime error (message has the cudaError string)                   */
    FASTECC_B200_ENOMEM   = -3,   /* device or pinned-host allocation failed                                 */
    FASTECC_B200_ENOINIT  = -4    /* fastecc_b200_init() has not succeeded in this process                   */
};

/* Process-wide context on one CUDA device: uploads the 16 MiB power table g^e (g = GF_Root(2^20),
 * GF(p).cpp:267-276) and caches scratch buffers.  Idempotent for the same device.  One process per GPU. */
int  fastecc_b200_init(int device);
void fastecc_b200_shutdown(void);
const char* fastecc_b200_last_error(void);
int  fastecc_b200_device(void);                 /* device ordinal of the live context, or -1 */
int  fastecc_b200_num_sms(void);

/* ---- reference call surface, host memory ---------------------------------------------------------------------
 * Replaces  template<T,P> void MFA_NTT(T** data, size_t N, size_t SIZE, bool InvNTT)      ntt.cpp:382-383
 * for <uint32_t,0xFFF00001>; call sites RS.cpp:41,63 and main.cpp:234,272,286.
 * data[i] -> SIZE words of block i (host memory, pageable or pinned).  In place: on return data[i] addresses
 * result block i; the pointer table itself is left untouched (the reference permutes it, callers only read
 * through data[i]).  The inverse transform is unnormalised, as in the reference.  1 <= N <= 2^20, power of two.
 * Beyond MFA_NTT (which needs a power of two): N = 3 * 2^k and 9 * 2^k with 2^k <= 2^20 are accepted too -- the reference's
 * order-3 / order-9 codelets (NTT3 ntt.cpp:26-46, NTT9 ntt.cpp:114-146) composed with the power-of-two transforms, result = the
 * DFT with GF_Root(N) like Slow_NTT (ntt.cpp:451-483) -- so that block counts need not be rounded up (README.md:176). */
int fastecc_b200_ntt_u32(uint32_t** data, size_t N, size_t SIZE_words, int inverse);

/* Opt-in for callers that keep one large block array alive and transform it repeatedly from PAGEABLE memory, the way the
 * reference's drivers do (RS.cpp:31-33, main.cpp:244-247: allocated once, never freed): the host entry points then page-lock
 * such an array in place on first sight (cudaHostRegister, about 0.2 s per GiB, once) so that its transfers pipeline like
 * pinned memory; registrations are dropped by fastecc_b200_pin_host_buffers(0) and at fastecc_b200_shutdown().  While it is enabled, arrays
 * that were passed to the host entry points must not be freed (a page-locked range that is unmapped and re-allocated would
 * be transferred from its old pages).  Off by default; fastecc_b200/shim/ntt.cpp turns it on. */
int fastecc_b200_pin_host_buffers(int enable);

/* Replaces the timed body of  template<T,P> void EncodeReedSolomon(size_t N, size_t SIZE)   RS.cpp:41-63
 * (iNTT, multiply block i by root_2N^i / N, NTT): N data blocks in, N parity blocks out, same buffers.
 * 1 <= N <= 2^19 (N = 2^20 is rejected: the reference silently computes garbage there, GF(p).cpp:274). */
int fastecc_b200_rs_encode(uint32_t** data, size_t N, size_t SIZE_words);

/* ---- device-resident variants (what bench.py times as `value`; PCIe excluded) ----------------------------------
 * d_blocks: device pointer to N rows of pitch_words words each (row i = block i), first SIZE_words of a row
 * are data.  The fast path needs d_blocks 16-byte aligned and pitch_words % 4 == 0 (columns are processed in
 * groups of 4 words, so up to 3 pad words per row are read and rewritten); other layouts go through an internal
 * repack.  stream: a cudaStream_t (NULL = default stream).  Asynchronous with respect to the host. */
/* (N = 3 * 2^k, 9 * 2^k: transforms only; the encoder needs a power of two.) */
int fastecc_b200_ntt_u32_dev  (uint32_t* d_blocks, size_t N, size_t SIZE_words, size_t pitch_words, int inverse, void* stream);
int fastecc_b200_rs_encode_dev(uint32_t* d_blocks, size_t N, size_t SIZE_words, size_t pitch_words, void* stream);

/* Measurement variant of fastecc_b200_rs_encode_dev (bench.py's per-kernel roofline): the same encode with a CUDA event
 * between the passes; synchronises the stream.  In: *n_passes = capacity of pass_ms / pass_kernel (may be NULL).  Out: the
 * duration of every pass kernel in ms, the name of the instantiation it ran as (static strings) and their number. */
int fastecc_b200_rs_encode_dev_timed(uint32_t* d_blocks, size_t N, size_t SIZE_words, size_t pitch_words, void* stream,
                                     float* pass_ms, const char** pass_kernel, int* n_passes);

/* ---- fewer parity blocks than data blocks ------------------------------------------------------
 * The reference only describes this (RS.cpp:65-66, NTT.md:46-49: "in order to compute only even-indexed points ...").
 * N data blocks in, M = N / 2^k parity blocks out: parity block j' is parity block (N/M)*j' of the full N -> N encode,
 * i.e. the value of the data polynomial at root_2N^(2*(N/M)*j' + 1).  On return data[j'], j' < M, hold the parity;
 * blocks M .. N-1 are undefined (host variant: left untouched).  1 <= M <= N <= 2^19, both powers of two.
 * The _dev variant needs a 16-byte aligned buffer and pitch_words % 4 == 0. */
int fastecc_b200_rs_encode_asym    (uint32_t** data, size_t N, size_t M, size_t SIZE_words);
int fastecc_b200_rs_encode_asym_dev(uint32_t* d_blocks, size_t N, size_t M, size_t SIZE_words, size_t pitch_words, void* stream);

/* ---- arbitrary bytes <-> words below P -----------------------------------------------------------
 * The reference describes, but does not implement, a bit-optimal recoding (GF.md:72-104, README.md:160-163): the top
 * 12 bits of the W words of a block are rewritten from base 4096 to base 4095 (no 0xFFF, hence every word < P) using
 * one extra bit, stored as word W of the block: 4096-byte blocks become 4100-byte (1025-word) blocks to encode.
 * d_bytes: n_blocks contiguous blocks of 4*W bytes; d_words: n_blocks rows of pitch_words words (>= W+1, % 4 == 0), words
 * 0..W of a row are meaningful.  W = words_per_block: multiple of 4, <= 1024.  Both buffers 16-byte aligned.
 * gfp_to_bytes(bytes_to_gfp(x)) == x for every x; parity rows produced by the encoder are NOT recoded data and are
 * stored as they are (GF.md:101-104). */
int fastecc_b200_bytes_to_gfp_dev(const void* d_bytes, uint32_t* d_words, size_t n_blocks, size_t words_per_block, size_t pitch_words, void* stream);
int fastecc_b200_gfp_to_bytes_dev(const uint32_t* d_words, void* d_bytes, size_t n_blocks, size_t words_per_block, size_t pitch_words, void* stream);

/* ---- element-wise helpers around the transforms (erasure decoding) --------------------------------
 * GF_Mul (GF(p).cpp:110-127) and GF_Inv (GF(p).cpp:293-297, x^(P-2); 0 -> 0) on device vectors of n words, and the shape
 * of the reference's scaling loop (RS.cpp:51-59) with an arbitrary constant per block: block i *= d_consts[i].
 * Inputs are taken mod P, results are canonical.  row_scale: 16-byte aligned buffer, pitch_words % 4 == 0. */
int fastecc_b200_gf_mul_dev(const uint32_t* d_a, const uint32_t* d_b, uint32_t* d_out, size_t n, void* stream);
int fastecc_b200_gf_inv_dev(const uint32_t* d_a, uint32_t* d_out, size_t n, void* stream);
int fastecc_b200_row_scale_dev(uint32_t* d_blocks, size_t n_rows, size_t SIZE_words, size_t pitch_words, const uint32_t* d_consts, void* stream);

/* ---- erasure decoding --------------------------------------------------------------------------------
 * The reference describes the decoder and does not implement it (README.md:88-119 "Fastest", RS.md:42-79, roadmap
 * README.md:173).  Code word of 2N rows: row 2i = data block i, row 2j+1 = parity block j of fastecc_b200_rs_encode
 * (RS.cpp:22-68).  Any set of at most N lost rows is recovered from the others by the formal-derivative method on two
 * order-2N transforms of the hot path.
 *   _pattern: everything that depends only on WHICH rows are lost (erasure locator by a product tree of batched transforms,
 *             its values and derivative values), built on the device once; erased = HOST array of ascending distinct row
 *             numbers < n_rows; n_rows = 2N = 2 .. 2^20, n_erased <= N.  Synchronises the stream.
 *   _recover: d_code = the 2N rows (pitch_words per row, first SIZE_words meaningful, 16-byte aligned, pitch % 4 == 0; content
 *             of the lost rows arbitrary); it is the workspace of the two transforms and is DESTROYED.  Row i of d_out
 *             (out_pitch_words per row) receives lost row erased[i], canonical residues.  Asynchronous on the stream.
 * A pattern may be used for any number of code words / columns.  Not pinned by reference code (there is none): checked
 * against interpolation by definition and by encode -> erase -> decode round trips (tests/test_decoder.py). */
typedef struct fastecc_b200_erasures fastecc_b200_erasures;
int    fastecc_b200_rs_decode_pattern(size_t n_rows, const uint32_t* erased, size_t n_erased, void* stream, fastecc_b200_erasures** pattern);
int    fastecc_b200_rs_decode_recover(const fastecc_b200_erasures* pattern, uint32_t* d_code, size_t SIZE_words, size_t pitch_words,
                                      uint32_t* d_out, size_t out_pitch_words, void* stream);
size_t fastecc_b200_rs_decode_count(const fastecc_b200_erasures* pattern);        /* n_erased */
void   fastecc_b200_rs_decode_free(fastecc_b200_erasures* pattern);

/* ---- one transform sharded over several GPUs (one process per GPU; BASELINE config 4) ---------------------------
 * Global block i = l*n_ranks + rank is local block l (data in, parity out).  An encode is: pass 0 on every rank,
 * all-to-all of whole blocks, pass 1, all-to-all, pass 2 -- the exchange is the caller's (fastecc_b200/sharded.py does it
 * with torch.distributed / NCCL); these are the three local passes of RS.cpp:41-63 with the sharded twists.
 * N = global number of data blocks, 2^11 .. 2^19; d_local holds N/n_ranks rows of pitch_words (16-byte aligned). */
int fastecc_b200_rs_encode_shard_pass(uint32_t* d_local, size_t N, int n_ranks, int rank, size_t SIZE_words, size_t pitch_words,
                                      int which, void* stream);

/* The same passes with the exchange FUSED into the kernels' stores (no all-to-all, no staging): every rank owns two
 * buffers of N/n_ranks rows, X (data in / parity out) and Y (intermediate), allocated with fastecc_b200_dev_alloc and
 * mapped into the other ranks with fastecc_b200_ipc_export / _open (CUDA IPC; peer access over NVLink).
 *   which 0: reads the local X, writes the rows of Y on their owners     d_src = X_local, d_peers[r] = rank r's Y
 *   which 1: reads the local Y, writes the rows of X on their owners     d_src = Y_local, d_peers[r] = rank r's X
 *   which 2: local, in place on X                                        d_src = X_local, d_peers[rank] = X_local
 * d_peers: HOST array of n_ranks device pointers (own buffer at [rank]).  The caller separates the passes with a
 * cross-rank barrier on the stream (fastecc_b200_shard_barrier below).  n_ranks = 2, 4 or 8; N = 2^12..2^19 with
 * both factors of N = N1*N2 (csrc/plan.h split_l1) at least 32*n_ranks, so that the 32 rows a thread stores go to one rank. */
int fastecc_b200_rs_encode_shard_pass_p2p(const uint32_t* d_src, uint32_t* const* d_peers, size_t N, int n_ranks, int rank,
                                          size_t SIZE_words, size_t pitch_words, int which, void* stream);
/* The cross-rank barrier between those passes, as a kernel on the stream (no NCCL, no host round trip): every rank owns an array
 * of FASTECC_B200_BARRIER_WORDS zero-initialised words from fastecc_b200_dev_alloc, mapped into the peers like X and Y;
 * d_flag_peers = HOST array of the n_ranks arrays (own at [rank]).  Rank r writes `epoch` into word [r] of every peer's array and
 * waits until its own words [0, n_ranks) have reached it; epoch = 1, 2, 3, ... in the same order on every rank.  The kernel runs
 * behind the pass whose peer stores it publishes.  A peer that never arrives makes it give up after about ten seconds and set
 * word [8] of the own array (check it after synchronising; fastecc_b200/sharded.py raises). */
#define FASTECC_B200_BARRIER_WORDS 16
int fastecc_b200_shard_barrier(uint32_t* const* d_flag_peers, int n_ranks, int rank, uint32_t epoch, void* stream);

/* One rank's share of a whole sharded encode / transform as ONE call: the passes above and the barriers between them, enqueued
 * on the stream.  d_x_peers / d_y_peers / d_flag_peers: HOST arrays of every rank's X, Y and barrier-flag buffers (own at [rank]);
 * *epoch: this rank's barrier counter, 0 before the first call (the same sequence of calls on every rank).  This is all a C++ host
 * needs for the multi-GPU path once the IPC handles are exchanged (integration/shard_example.cpp). */
int fastecc_b200_rs_encode_shard_p2p(uint32_t* const* d_x_peers, uint32_t* const* d_y_peers, uint32_t* const* d_flag_peers, uint32_t* epoch,
                                     size_t N, int n_ranks, int rank, size_t SIZE_words, size_t pitch_words, void* stream);
int fastecc_b200_ntt_shard_p2p(uint32_t* const* d_x_peers, uint32_t* const* d_y_peers, uint32_t* const* d_flag_peers, uint32_t* epoch,
                               size_t N, int n_ranks, int rank, size_t SIZE_words, size_t pitch_words, int inverse, void* stream);

/* ONE standalone transform (MFA_NTT, ntt.cpp:382-447; unnormalised inverse) sharded the same way: cyclic blocks in and out.
 *   which 0: reads the local X, stores into the Ys of the owners (the four-step transpose as peer stores)   d_src = X_local, d_peers[r] = rank r's Y
 *   which 1: local, Y -> X                                                                                  d_src = Y_local, d_peers[rank] = X_local
 * with one cross-rank barrier between them.  N = 2^11 .. 2^20, n_ranks = 2, 4 or 8, first tile height (N1 of csrc/plan.h) >= 32 * n_ranks. */
int fastecc_b200_ntt_shard_pass_p2p(const uint32_t* d_src, uint32_t* const* d_peers, size_t N, int n_ranks, int rank,
                                    size_t SIZE_words, size_t pitch_words, int inverse, int which, void* stream);

/* The decomposition N = N1 * N2 the passes use (csrc/plan.h split_l1, FASTECC_B200_SPLIT honoured), for callers that do the
 * exchange themselves; *fused_exchange_ok: bit 0 = the encode _p2p passes, bit 1 = the NTT _p2p passes support (N, n_ranks).  Fails if N cannot be sharded. */
int fastecc_b200_shard_geometry(size_t N, int n_ranks, size_t* N1, size_t* N2, int* fused_exchange_ok);
/* cudaMemcpy2DAsync between pinned host memory and a device buffer (column chunks of a block array), for pipelined
 * host <-> device staging around the sharded passes (fastecc_b200/sharded.py encode_host). */
int fastecc_b200_copy2d_async(void* dst, size_t dst_pitch_bytes, const void* src, size_t src_pitch_bytes, size_t width_bytes, size_t rows,
                              int to_device, void* stream);
void* fastecc_b200_dev_alloc(size_t bytes);                       /* cudaMalloc: exportable, unlike a caching-allocator block */
void  fastecc_b200_dev_free(void* d_ptr);
int   fastecc_b200_ipc_export(void* d_ptr, void* handle64);       /* 64-byte cudaIpcMemHandle_t */
int   fastecc_b200_ipc_open(const void* handle64, void** d_ptr);  /* in another process on the same node */
int   fastecc_b200_ipc_close(void* d_ptr);

/* Replaces  uint32_t hash(T** data, size_t N, size_t SIZE)   main.cpp:203-212: the rolling checksum the reference's
 * driver prints before and after a transform, walked through data[i] in block order (host memory).  Sequential by
 * construction (about one second for 2 GiB); used to compare results with the reference's published / golden hashes. */
uint32_t fastecc_b200_hash_u32(uint32_t* const* data, size_t N, size_t SIZE_words);

/* Counters for benchmarking: kernels launched by this library since init (all entry points). */
unsigned long long fastecc_b200_kernel_launches(void);

/* Pinned host memory helpers for callers that want the fast host path (cudaHostAlloc / cudaFreeHost). */
void* fastecc_b200_host_alloc(size_t bytes);
void  fastecc_b200_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
