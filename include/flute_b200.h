/*
 * flute_b200 -- C ABI of the B200 (sm_100a) LUT-quantized GEMM engine.
 *
 * This is the drop-in boundary for the reference's `flute/csrc`: plain pointers and sizes,
 * no torch types.  Every device pointer is a CUDA device pointer on `device`; `stream` is a
 * cudaStream_t passed as void*.  All calls are asynchronous on `stream`, allocate nothing,
 * never synchronise, and are CUDA-graph capturable.  Return value: 0 (FLUTE_B200_OK) or a
 * negative FLUTE_B200_ERR_* code; `flute_b200_last_error()` gives the message for the calling
 * thread.  Citations are into the reference tree (HanGuo97/flute @ v0.4.2).
 */
#ifndef FLUTE_B200_H_
#define FLUTE_B200_H_

#if defined(__GNUC__)
#define FLUTE_B200_API __attribute__((visibility("default")))
#else
#define FLUTE_B200_API
#endif

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLUTE_B200_VERSION 100 /* 0.1.0 */

/* dtype codes: the arithmetic type T of activations, scales, tables and outputs */
#define FLUTE_B200_F16 0
#define FLUTE_B200_BF16 1

/* flags */
#define FLUTE_B200_FLAG_PDL 1 /* launch with programmatic stream serialization (overlap with the previous kernel) */
#define FLUTE_B200_FLAG_STATIC_WEIGHTS 2 /* with PDL: Q, S and table2 are not written by earlier work still in flight on
                                           the stream, so the kernel may start streaming them before that work completes */

/* error codes (the reference raises AT_ERROR / launch-check failures, flute/csrc/qgemm.cpp:82,153,171) */
enum {
    FB_OK = 0,
    FB_ERR_BITS = -1,      /* num_bits not in {2,3,4}            (qgemm.cpp:171) */
    FB_ERR_GROUP = -2,     /* group_size not in {64,128,256}     (qgemm.cpp:153) */
    FB_ERR_DTYPE = -3,     /* dtype not fp16/bf16                (qgemm.cpp:176-193) */
    FB_ERR_SHAPE = -4,     /* K % 64, K % group, N % block, M < 0 ... (ops.py:17-49) */
    FB_ERR_TILE_P = -5,    /* tile_P not 32/64, or 64 with 3-bit (utils.py:138-139) */
    FB_ERR_WORKSPACE = -6, /* workspace too small for the Stream-K partials */
    FB_ERR_LAUNCH = -7,    /* CUDA launch failure                (qgemm.cpp:82) */
    FB_ERR_DRIVER = -8,    /* cuTensorMapEncodeTiled unavailable */
    FB_ERR_TENSORMAP = -9, /* tensor-map encode rejected (misaligned pointer / stride) */
    FB_ERR_NULL = -10,     /* null pointer */
    FB_ERR_DEVICE = -11,   /* not an sm_100 device / cudaSetDevice failed */
    FB_ERR_INTERNAL = -12,
    FB_ERR_HADAMARD = -13, /* hadamard size not a power of two <= 32768 (hadamard_transform.cpp:24-26) */
    FB_ERR_KERNEL = -14    /* a previous kernel trapped; see flute_b200_last_error() */
};
#define FLUTE_B200_OK FB_OK

/*
 * D[M,N] = A[M,K] . W_hat[K,N],  W_hat[k,n] = round_T(table2[code].{lo,hi} * S[n, k / group_size]).
 *
 * Numerics.  Prefill and general kernels (4-bit M > 16, 2-bit M > 4, all of 3-bit): exactly the formula above, fp32
 * accumulation -- the reference's arithmetic (packbits_utils.hpp:105,139).  Decode kernel (4-bit M <= 16, 2-bit M <= 4): the
 * group scale is applied to the fp32 partial sum of each group instead,
 * D[m,n] = sum_g S[n,g] * (sum_{k in g} A[m,k] * table2[code(k,n)]), i.e. the product table*S is NOT rounded to T first.
 * With a one-hot A this still yields round_T(table*S) exactly (bit-identical to the reference's identity reconstruction;
 * tested), for general A it differs from the other kernels' result by rounding only (bounded in
 * tests/test_qgemm_gpu.py::test_decode_vs_prefill_numerics_bound: < 1.0e-3 fp16 / 5.5e-3 bf16 relative, inside the
 * reference's own 2.0e-3 / 1.1e-2 acceptance bound).  Consequence: 4-bit results are not batch-invariant across M = 16 | 17.
 * Split-K partial sums are added in fp32 in arrival order, so results are reproducible to fp32 reduction-order noise only
 * (as with the reference's Stream-K fix-up).
 *
 * Replaces  torch.ops.flute.qgemm_raw_simple -> qgemm_raw<T,NumBits,GroupSize> -> _qgemm_raw -> qgemm_host
 *           (flute/csrc/qgemm.cpp:44-198, qgemm_kernel_raw_generated.cu:15-768, qgemm_kernel.hpp:841-939).
 *   A         [M, K]  T, row-major, contiguous                      (qgemm.cpp:71,110)
 *   Q         [N/16*num_bits, K] int16, the reference's packed wire format (flute/utils.py:59-253), tile_P as packed
 *   D         [M, N]  T, row-major (output)
 *   S         [N, K/group_size] T                                   (qgemm_kernel.hpp:137)
 *   table     [2^num_bits] T -- accepted for signature parity; like every instantiated reference
 *             template (codegen_utils.py:97-102) the kernel looks up through table2 only
 *   table2    [2^num_bits, 2^num_bits, 1] float32 = bit view of T pairs, low half = even k (flute/utils.py:15-33)
 *   workspace zero-initialised once by the caller, reused across calls on one stream; the kernel
 *             leaves every flag it touches zero again (contract of flute/utils.py:36-56)
 *   tile_P    32 or 64: the packing the weights were stored with (TEMPLATE_CONFIGS[(bits,id)]["TileP"])
 *   dtype     FLUTE_B200_F16 / FLUTE_B200_BF16
 */
FLUTE_B200_API int flute_b200_qgemm(const void* A, const void* Q, void* D, const void* S, const void* table, const void* table2,
                     void* workspace, size_t workspace_bytes, int M, int N, int K, int num_bits, int group_size,
                     int tile_P, int dtype, int flags, int device, void* stream);

/*
 * Same computation with HOST activation / output buffers: copies A host->device into
 * `A_dev_scratch`, runs the GEMM, copies D device->host, all on `stream` (pinned host memory makes
 * the copies asynchronous).  Weights, scales, tables and workspace stay device-resident, as they
 * do behind `FluteLinear.forward` (flute/integrations/base.py:277-301).  This is the call
 * bench.py times for its end-to-end number.
 */
FLUTE_B200_API int flute_b200_qgemm_host(const void* A_host, void* D_host, void* A_dev_scratch, void* D_dev_scratch, const void* Q,
                          const void* S, const void* table, const void* table2, void* workspace,
                          size_t workspace_bytes, int M, int N, int K, int num_bits, int group_size, int tile_P,
                          int dtype, int flags, int device, void* stream);

/*
 * out = in.reshape(rows, had_size) @ H / sqrt(had_size)  (Sylvester order), T in, T out, fp32 inside.
 * Replaces hadamard_transform(Tensor&, bool) (flute/csrc/hadamard_transform.cpp:17-57,
 * hadamard_transform_cuda.cu:92-748) as used by qgemm_raw_simple_hadamard (qgemm.cpp:201-244).
 * `in` may equal `out`.
 */
FLUTE_B200_API int flute_b200_hadamard(const void* in, void* out, long rows, int had_size, int dtype, int device, void* stream);

/*
 * W_hat[K, N] (T, row-major) from the packed weights: the dequantiser alone.  Replaces the
 * identity-matrix GEMM the reference uses to reconstruct / unpack weights
 * (flute/utils.py:347-407): same values, no K x K identity operand.
 */
FLUTE_B200_API int flute_b200_dequantize(const void* Q, const void* S, const void* table2, void* W_hat, int N, int K, int num_bits,
                          int group_size, int tile_P, int dtype, int device, void* stream);

/* Bytes of workspace `make_workspace_streamk` should allocate (flute/utils.py:36-45 formula). */
FLUTE_B200_API size_t flute_b200_workspace_bytes(int num_sms);

/* SM count of `device` (flute/utils.py:410-412), or a negative error code. */
FLUTE_B200_API int flute_b200_num_sms(int device);

/* Largest activation-row tile (MMA N) the engine uses for `num_bits`; informational. */
FLUTE_B200_API int flute_b200_max_batch_tile(int num_bits);

/*
 * Tensor-parallel column shard with the exchange fused into the GEMM (the reference's only
 * forward collective is the all-gather vLLM's ColumnParallelLinear adds around flute.qgemm_simple,
 * flute/integrations/vllm_utils.py:242-244,328-349).
 *
 * Rank r of tp owns output columns [r*N, (r+1)*N) of an n_total = tp*N wide linear (its Q / S row slices,
 * flute_b200/parallel.py).  flute_b200_qgemm_tp computes them like flute_b200_qgemm, but the kernel's epilogue stores
 * the slice straight into EVERY rank's gathered output -- peer-mapped device pointers (e.g. torch symmetric memory over
 * NVLink): out_peers[r] is rank r's plain image [M, n_total] in T (kept current only if write_plain), ll_peers[r] its
 * "word image": [M, n_total] 8-byte words {value, sequence number}, zero-initialised once.  A consumer call given
 * `in_ll` (the word image of the buffer its A lives in, at A's first element; row stride in_ll_stride words) reads A
 * from there and spins per word until it carries the expected sequence number: an aligned 8-byte store is single-copy
 * atomic, so no fence, no flag and no collective kernel sit between producer and consumer -- one NVLink one-way trip.
 * sequence = (*epoch - 1) * uses + call + 1, with `uses` = producing calls per step of that buffer, `call` = index of the
 * producing call within the step (out_* for this call's output, in_* for A) and *epoch the step number (a device word,
 * >= 1, advanced once per step by flute_b200_tp_advance on the same stream).  CUDA-graph capturable; every rank issues
 * the same sequence of calls.  M <= 4 only (2 / 4 bits).  Every rank uses ONE layout: ll_peers[r] -
 * out_peers[r] is the same for all r (one symmetric allocation per rank).
 *
 * Readers that are not qgemm_tp calls (the copy of the step's result, another library's kernel) take the plain image:
 * the producing call sets write_plain, then flute_b200_tp_publish (after it, same stream) makes this rank's stores
 * visible system-wide and bumps that output's arrival counter on every rank, and flute_b200_tp_wait holds the stream
 * until this rank's counter has seen all tp publishes of the step.
 */
typedef struct flute_b200_tp {
    int tp, rank;
    int n_total;                 /* columns of the gathered output = tp * N */
    void* out_peers[8];          /* every rank's plain image [M, n_total] of THIS output */
    void* ll_peers[8];           /* every rank's word image [M, n_total] x 8 bytes of THIS output */
    int write_plain;             /* 0: every reader of this output is a qgemm_tp call using in_ll */
    unsigned out_uses, out_call;
    const void* in_ll;           /* word image of A (NULL: A is an ordinary local tensor) */
    int in_ll_stride;
    unsigned in_uses, in_call;
    const unsigned* epoch;       /* device word holding the step number */
} flute_b200_tp;

FLUTE_B200_API int flute_b200_qgemm_tp(const void* A, const void* Q, const void* S, const void* table, const void* table2,
                        void* workspace, size_t workspace_bytes, int M, int N, int K, int num_bits, int group_size,
                        int tile_P, int dtype, int flags, int device, void* stream, const flute_b200_tp* tp);
/* ++*epoch on `stream` (one tiny kernel): call once at the start of every step, before the step's first qgemm_tp. */
FLUTE_B200_API int flute_b200_tp_advance(unsigned* epoch, int device, void* stream);
/* After the producing call(s) on `stream`: system-scope fence, then +1 on flag_peers[r] for r < tp. */
FLUTE_B200_API int flute_b200_tp_publish(unsigned* const* flag_peers, int tp, int device, void* stream);
/* Stream-ordered wait until `flag` has received (*epoch - 1) * per_step + offset arrivals (one publish per rank and use:
 * per_step = offset = tp for an output published once per step). */
FLUTE_B200_API int flute_b200_tp_wait(const unsigned* flag, unsigned per_step, unsigned offset, const unsigned* epoch, int device,
                       void* stream);

/* Name of the kernel the automatic dispatch of flute_b200_qgemm selects for (M, num_bits, dtype) -- reporting only
 * (bench.py's roofline.kernel); a static string. */
FLUTE_B200_API const char* flute_b200_dispatch_name(int M, int num_bits, int dtype);
/* CTAs a decode-shaped launch (M <= 16) uses for `total_stages` = column tiles x (K / 64) pipeline stages on a device with
 * `num_sms` SMs -- reporting / test hook of the host-side schedule (csrc/qgemm_sm100.h decode_grid_for): all SMs but four,
 * or fewer when a CTA's share of 2..8 stages can be aligned to tile boundaries.  No GPU needed. */
FLUTE_B200_API int flute_b200_decode_grid(long long total_stages, int k_iters, int num_sms, int num_bits);

FLUTE_B200_API const char* flute_b200_last_error(void);
FLUTE_B200_API const char* flute_b200_error_string(int code);
FLUTE_B200_API int flute_b200_version(void);

/*
 * Diagnostics.  The kernels bound every barrier wait (default 10 s; 0 disables) and trap with a
 * reason instead of hanging the GPU.  `flute_b200_set_timeout_ms` changes the bound for later
 * launches; `flute_b200_check(device)` returns FB_ERR_KERNEL and fills last_error if a kernel on
 * that device recorded a timeout since the last check.
 */
FLUTE_B200_API void flute_b200_set_timeout_ms(long ms);
FLUTE_B200_API int flute_b200_check(int device);

/* Test hook: when non-null, every CTA of later qgemm launches writes 8 globaltimer stamps (ns) to
 * device_ptr[blockIdx * 8 + i]: start, setup done, first tile landed, last MMA issued, accumulators
 * complete, epilogue done, fix-up done, exit.  Pass NULL to switch tracing off. */
FLUTE_B200_API void flute_b200_set_trace_buffer(void* device_ptr);

/* Test hook: kernel selection for later qgemm launches: -1 automatic (default: decode kernel for M <= 16 at 4 bits and
 * M <= 4 at 2 bits, prefill kernel for M > 16 at 4 bits, general kernel otherwise); 0 general kernel, LARGE footprint
 * (1 CTA/SM); 1 general kernel, SMALL footprint (2 CTAs/SM; M <= 16, 2/4 bits); 2 same as -1 (historical: the decode
 * kernel's opt-in for 5 <= M <= 16).  Bits 8..15: perf-ablation mask, bits 16..23: decode kernel's L2 prefetch distance + 1,
 * bits 24..30: SMs left idle per decode launch
 * (tools/microbench.py only). */
FLUTE_B200_API void flute_b200_set_variant(int variant);

/* Test hook: like flute_b200_qgemm with explicit tiling overrides (0 / -1 = engine's choice) and an
 * optional device buffer receiving the first dequantised TMEM chunk (128 x 128 uint32). */
FLUTE_B200_API int flute_b200_qgemm_debug(const void* A, const void* Q, void* D, const void* S, const void* table2, void* workspace,
                           size_t workspace_bytes, int M, int N, int K, int num_bits, int group_size, int tile_P,
                           int dtype, int flags, int device, void* stream, int force_mb, int force_stages,
                           int force_grid, int force_streamk, void* dbg_chunk);

#ifdef __cplusplus
}
#endif

#endif /* FLUTE_B200_H_ */
