// C ABI of flute_b200 (declared in include/flute_b200.h): argument validation, device / stream
// handling, error reporting.  No torch types, no allocation on the hot path, no host sync.
#include <cuda_runtime.h>
#include <mutex>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/flute_b200.h"
#include "aux_kernels.h"
#include "ptx.cuh"
#include "qgemm_sm100.h"

namespace {

thread_local char g_last_error[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

constexpr int kMaxDevices = 64;
struct DeviceState {
    bool probed = false;
    int num_sms = 0;
    int cc_major = 0, cc_minor = 0;
    fb::Diag* diag_host = nullptr;   // pinned + mapped
    fb::Diag* diag_dev = nullptr;
};
DeviceState g_dev[kMaxDevices];
std::mutex g_dev_mutex;                    // guards the lazy per-device initialisation below
long g_timeout_ms = 10000;
unsigned long long* g_trace = nullptr;
int g_ablate = 0;
int g_grid_slack = 0;                     // test hook: SMs left idle per launch (grid = num_sms - slack)
int g_l2_prefetch = -1;                   // test hook: decode kernel's L2 prefetch distance (-1 = engine's choice)
int g_variant = -1;                       // test hook: kernel footprint override   // test hook: per-CTA globaltimer stamps

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    int rc = FB_OK;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); rc = FB_ERR_DEVICE; return; }
        if (prev != device) {
            if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); rc = FB_ERR_DEVICE; return; }
            switched = true;
        }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

int probe_device(int device) {
    if (device < 0 || device >= kMaxDevices) return fail(FB_ERR_DEVICE, "device index %d out of range", device);
    DeviceState& d = g_dev[device];
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    if (d.probed) return FB_OK;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        cudaGetLastError();
        return fail(FB_ERR_DEVICE, "cudaGetDeviceProperties(%d) failed (no CUDA device?)", device);
    }
    d.num_sms = prop.multiProcessorCount;
    d.cc_major = prop.major;
    d.cc_minor = prop.minor;
    d.probed = true;
    return FB_OK;
}

// Lazily create the mapped diagnostics block -- never while a stream capture is in flight.
fb::Diag* diag_for(int device, cudaStream_t stream) {
    DeviceState& d = g_dev[device];
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    if (d.diag_dev != nullptr) return d.diag_dev;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (st != cudaStreamCaptureStatusNone) return nullptr;
    void* h = nullptr;
    if (cudaHostAlloc(&h, sizeof(fb::Diag), cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    memset(h, 0, sizeof(fb::Diag));
    void* dptr = nullptr;
    if (cudaHostGetDevicePointer(&dptr, h, 0) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(h); return nullptr; }
    d.diag_host = static_cast<fb::Diag*>(h);
    d.diag_dev = static_cast<fb::Diag*>(dptr);
    return d.diag_dev;
}

int validate_quant(int N, int K, int num_bits, int group_size, int tile_P, int dtype) {
    if (num_bits != 2 && num_bits != 3 && num_bits != 4) return fail(FB_ERR_BITS, "Unsupported `num_bits` %d", num_bits);
    if (group_size != 64 && group_size != 128 && group_size != 256)
        return fail(FB_ERR_GROUP, "Unsupported `group_size` %d", group_size);
    if (dtype != FLUTE_B200_F16 && dtype != FLUTE_B200_BF16) return fail(FB_ERR_DTYPE, "Unsupported dtype code %d", dtype);
    if (tile_P != 32 && tile_P != 64) return fail(FB_ERR_TILE_P, "tile_P must be 32 or 64, got %d", tile_P);
    if (num_bits == 3 && tile_P != 32) return fail(FB_ERR_TILE_P, "3-bit weights are packed with tile_P == 32 only");
    if (K <= 0 || K % 64 != 0) return fail(FB_ERR_SHAPE, "K = %d must be a positive multiple of 64", K);
    if (K % group_size != 0) return fail(FB_ERR_SHAPE, "K = %d is not a multiple of group_size = %d", K, group_size);
    const int block = (num_bits == 3) ? 512 : (16 / num_bits) * tile_P;
    if (N <= 0 || N % block != 0)
        return fail(FB_ERR_SHAPE, "N = %d must be a positive multiple of %d for %d-bit / tile_P = %d", N, block, num_bits, tile_P);
    return FB_OK;
}

int run_qgemm(const void* A, const void* Q, void* D, const void* S, const void* table2, void* workspace,
              size_t workspace_bytes, int M, int N, int K, int num_bits, int group_size, int tile_P, int dtype,
              int flags, int device, void* stream, int force_mb, int force_stages, int force_grid, int force_streamk,
              void* dbg, const flute_b200_tp* tp = nullptr) {
    int rc = validate_quant(N, K, num_bits, group_size, tile_P, dtype);
    if (rc != FB_OK) return rc;
    if (M < 0) return fail(FB_ERR_SHAPE, "M = %d is negative", M);
    if (M == 0) return FB_OK;   // empty batch: nothing to compute (qgemm.cpp:110 reshape({-1,K}) of an empty input)
    if (!A || !Q || !D || !S || !table2 || !workspace) return fail(FB_ERR_NULL, "null pointer argument");
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(Q) & 15))
        return fail(FB_ERR_TENSORMAP, "A and Q must be 16-byte aligned for TMA");
    if (force_mb != 0 && (force_mb % 16 != 0 || force_mb < 16 || force_mb > fb::qgemm_max_mb(num_bits)))
        return fail(FB_ERR_SHAPE, "force_mb %d invalid", force_mb);
    rc = probe_device(device);
    if (rc != FB_OK) return rc;
    if (g_dev[device].cc_major != 10)
        return fail(FB_ERR_DEVICE, "device %d is sm_%d%d; this engine is sm_100a only", device, g_dev[device].cc_major,
                    g_dev[device].cc_minor);
    DeviceGuard guard(device);
    if (guard.rc != FB_OK) return fail(guard.rc, "cudaSetDevice(%d) failed", device);

    fb::QgemmArgs a{};
    a.A = A; a.Q = Q; a.D = D; a.S = S; a.table2 = table2;
    a.workspace = workspace; a.workspace_bytes = workspace_bytes;
    a.M = M; a.N = N; a.K = K;
    a.num_bits = num_bits; a.group_size = group_size; a.tile_p = tile_P;
    a.bf16 = (dtype == FLUTE_B200_BF16);
    a.flags = flags;
    a.device = device;
    a.num_sms = g_dev[device].num_sms;
    a.diag = diag_for(device, static_cast<cudaStream_t>(stream));
    a.dbg = static_cast<uint32_t*>(dbg);
    a.trace = g_trace;
    a.variant = g_variant;
    a.ablate = g_ablate;
    a.l2_prefetch = g_l2_prefetch;
    a.tp = tp;
    if (tp != nullptr && !(M <= 4 && (num_bits == 4 || num_bits == 2)))
        return fail(FB_ERR_SHAPE, "tensor-parallel fused exchange: decode shapes only (M <= 4 at 2/4 bits), got M=%d bits=%d", M, num_bits);
    a.timeout_ns = (g_timeout_ms > 0) ? (uint64_t)g_timeout_ms * 1000000ull : 0ull;
    a.force_mb = force_mb; a.force_stages = force_stages; a.force_grid = force_grid; a.force_streamk = force_streamk;
    if (a.force_grid <= 0 && g_grid_slack > 0 && g_grid_slack < a.num_sms) a.force_grid = a.num_sms - g_grid_slack;
    rc = fb::qgemm_launch(a, static_cast<cudaStream_t>(stream));
    switch (rc) {
        case FB_OK: return FB_OK;
        case FB_ERR_WORKSPACE:
            return fail(rc, "workspace of %zu bytes is too small for M=%d N=%d K=%d (allocate flute_b200_workspace_bytes())",
                        workspace_bytes, M, N, K);
        case FB_ERR_LAUNCH: return fail(rc, "CUDA kernel launch failed (M=%d N=%d K=%d bits=%d)", M, N, K, num_bits);
        case FB_ERR_DRIVER: return fail(rc, "cuTensorMapEncodeTiled not available from the CUDA driver");
        case FB_ERR_TENSORMAP: return fail(rc, "tensor-map encoding failed (pointer / stride alignment)");
        default: return fail(rc, "qgemm launch failed with code %d", rc);
    }
}

}  // namespace

extern "C" {

int flute_b200_qgemm(const void* A, const void* Q, void* D, const void* S, const void* table, const void* table2,
                     void* workspace, size_t workspace_bytes, int M, int N, int K, int num_bits, int group_size,
                     int tile_P, int dtype, int flags, int device, void* stream) {
    (void)table;   // semantics live in table2 (every reference template is a Vectorized* mode, codegen_utils.py:97-102)
    return run_qgemm(A, Q, D, S, table2, workspace, workspace_bytes, M, N, K, num_bits, group_size, tile_P, dtype, flags,
                     device, stream, 0, 0, 0, -1, nullptr);
}

int flute_b200_qgemm_tp(const void* A, const void* Q, const void* S, const void* table, const void* table2, void* workspace,
                        size_t workspace_bytes, int M, int N, int K, int num_bits, int group_size, int tile_P, int dtype,
                        int flags, int device, void* stream, const flute_b200_tp* tp) {
    (void)table;
    if (tp == nullptr) return fail(FB_ERR_NULL, "null tensor-parallel descriptor");
    if (tp->tp < 1 || tp->tp > 8 || tp->rank < 0 || tp->rank >= tp->tp) return fail(FB_ERR_SHAPE, "tp=%d rank=%d out of range", tp->tp, tp->rank);
    if (tp->n_total != tp->tp * N) return fail(FB_ERR_SHAPE, "n_total=%d is not tp*N=%d", tp->n_total, tp->tp * N);
    if (tp->out_peers[tp->rank] == nullptr) return fail(FB_ERR_NULL, "null gathered-output pointer");
    // the local slice pointer doubles as D for validation; with tp == 1 it is simply the output
    void* D_local = static_cast<char*>(tp->out_peers[tp->rank]);
    return run_qgemm(A, Q, D_local, S, table2, workspace, workspace_bytes, M, N, K, num_bits, group_size, tile_P, dtype, flags,
                     device, stream, 0, 0, 0, -1, nullptr, tp);
}

int flute_b200_tp_advance(unsigned* epoch, int device, void* stream) {
    if (epoch == nullptr) return fail(FB_ERR_NULL, "null epoch pointer");
    int rc = probe_device(device);
    if (rc != FB_OK) return rc;
    DeviceGuard guard(device);
    if (guard.rc != FB_OK) return fail(guard.rc, "cudaSetDevice(%d) failed", device);
    rc = fb::tp_advance_launch(epoch, static_cast<cudaStream_t>(stream));
    return rc == FB_OK ? FB_OK : fail(rc, "tp_advance launch failed");
}

int flute_b200_tp_publish(unsigned* const* flag_peers, int tp, int device, void* stream) {
    if (flag_peers == nullptr || tp < 1 || tp > 8) return fail(FB_ERR_SHAPE, "tp_publish: tp=%d", tp);
    for (int r = 0; r < tp; ++r)
        if (flag_peers[r] == nullptr) return fail(FB_ERR_NULL, "null counter pointer");
    int rc = probe_device(device);
    if (rc != FB_OK) return rc;
    DeviceGuard guard(device);
    if (guard.rc != FB_OK) return fail(guard.rc, "cudaSetDevice(%d) failed", device);
    rc = fb::tp_publish_launch(flag_peers, tp, static_cast<cudaStream_t>(stream));
    return rc == FB_OK ? FB_OK : fail(rc, "tp_publish launch failed");
}

int flute_b200_tp_wait(const unsigned* flag, unsigned per_step, unsigned offset, const unsigned* epoch, int device, void* stream) {
    if (flag == nullptr || epoch == nullptr) return fail(FB_ERR_NULL, "null flag / epoch pointer");
    int rc = probe_device(device);
    if (rc != FB_OK) return rc;
    DeviceGuard guard(device);
    if (guard.rc != FB_OK) return fail(guard.rc, "cudaSetDevice(%d) failed", device);
    const uint64_t timeout_ns = (g_timeout_ms > 0) ? (uint64_t)g_timeout_ms * 1000000ull : 0ull;
    rc = fb::tp_wait_launch(flag, per_step, offset, epoch, timeout_ns, diag_for(device, static_cast<cudaStream_t>(stream)),
                            static_cast<cudaStream_t>(stream));
    return rc == FB_OK ? FB_OK : fail(rc, "tp_wait launch failed");
}

int flute_b200_qgemm_debug(const void* A, const void* Q, void* D, const void* S, const void* table2, void* workspace,
                           size_t workspace_bytes, int M, int N, int K, int num_bits, int group_size, int tile_P,
                           int dtype, int flags, int device, void* stream, int force_mb, int force_stages,
                           int force_grid, int force_streamk, void* dbg_chunk) {
    return run_qgemm(A, Q, D, S, table2, workspace, workspace_bytes, M, N, K, num_bits, group_size, tile_P, dtype, flags,
                     device, stream, force_mb, force_stages, force_grid, force_streamk, dbg_chunk);
}

int flute_b200_qgemm_host(const void* A_host, void* D_host, void* A_dev_scratch, void* D_dev_scratch, const void* Q,
                          const void* S, const void* table, const void* table2, void* workspace,
                          size_t workspace_bytes, int M, int N, int K, int num_bits, int group_size, int tile_P,
                          int dtype, int flags, int device, void* stream) {
    if (!A_host || !D_host || !A_dev_scratch || !D_dev_scratch) return fail(FB_ERR_NULL, "null host/scratch pointer");
    if (M <= 0) return (M == 0) ? FB_OK : fail(FB_ERR_SHAPE, "M = %d is negative", M);
    int rc = probe_device(device);
    if (rc != FB_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        DeviceGuard guard(device);
        if (guard.rc != FB_OK) return fail(guard.rc, "cudaSetDevice(%d) failed", device);
        if (cudaMemcpyAsync(A_dev_scratch, A_host, (size_t)M * K * 2, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            cudaGetLastError();
            return fail(FB_ERR_LAUNCH, "host->device copy of the activations failed");
        }
    }
    rc = flute_b200_qgemm(A_dev_scratch, Q, D_dev_scratch, S, table, table2, workspace, workspace_bytes, M, N, K,
                          num_bits, group_size, tile_P, dtype, flags & ~FLUTE_B200_FLAG_PDL, device, stream);
    if (rc != FB_OK) return rc;
    DeviceGuard guard(device);
    if (guard.rc != FB_OK) return fail(guard.rc, "cudaSetDevice(%d) failed", device);
    if (cudaMemcpyAsync(D_host, D_dev_scratch, (size_t)M * N * 2, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
        cudaGetLastError();
        return fail(FB_ERR_LAUNCH, "device->host copy of the output failed");
    }
    return FB_OK;
}

int flute_b200_hadamard(const void* in, void* out, long rows, int had_size, int dtype, int device, void* stream) {
    if (dtype != FLUTE_B200_F16 && dtype != FLUTE_B200_BF16) return fail(FB_ERR_DTYPE, "Only fp16 and bf16 supported currently");
    if (had_size <= 0 || (had_size & (had_size - 1)) || had_size > (1 << 15))
        return fail(FB_ERR_HADAMARD, "Only power of two Hadamard sizes up to 2^15 are supported, got %d", had_size);
    if (rows < 0) return fail(FB_ERR_SHAPE, "rows = %ld is negative", rows);
    if (rows == 0) return FB_OK;
    if (!in || !out) return fail(FB_ERR_NULL, "null pointer argument");
    int rc = probe_device(device);
    if (rc != FB_OK) return rc;
    DeviceGuard guard(device);
    if (guard.rc != FB_OK) return fail(guard.rc, "cudaSetDevice(%d) failed", device);
    rc = fb::hadamard_launch(in, out, rows, had_size, dtype == FLUTE_B200_BF16, static_cast<cudaStream_t>(stream));
    if (rc != FB_OK) return fail(rc, "hadamard launch failed (rows=%ld, h=%d)", rows, had_size);
    return FB_OK;
}

int flute_b200_dequantize(const void* Q, const void* S, const void* table2, void* W_hat, int N, int K, int num_bits,
                          int group_size, int tile_P, int dtype, int device, void* stream) {
    int rc = validate_quant(N, K, num_bits, group_size, tile_P, dtype);
    if (rc != FB_OK) return rc;
    if (!Q || !S || !table2 || !W_hat) return fail(FB_ERR_NULL, "null pointer argument");
    rc = probe_device(device);
    if (rc != FB_OK) return rc;
    DeviceGuard guard(device);
    if (guard.rc != FB_OK) return fail(guard.rc, "cudaSetDevice(%d) failed", device);
    rc = fb::dequantize_launch(Q, S, table2, W_hat, N, K, num_bits, group_size, tile_P, dtype == FLUTE_B200_BF16,
                               static_cast<cudaStream_t>(stream));
    if (rc != FB_OK) return fail(rc, "dequantize launch failed");
    return FB_OK;
}

size_t flute_b200_workspace_bytes(int num_sms) {
    // flute/utils.py:36-45: blocks_max(num_sms*4) * threads_max(256) * accum_size_max(4*64*8) + 4 * blocks_max
    if (num_sms <= 0) return 0;
    const size_t blocks_max = (size_t)num_sms * 4;
    return blocks_max * 256 * (4 * 64 * 8) + 4 * blocks_max;
}

int flute_b200_num_sms(int device) {
    int rc = probe_device(device);
    if (rc != FB_OK) return rc;
    return g_dev[device].num_sms;
}

int flute_b200_max_batch_tile(int num_bits) {
    if (num_bits != 2 && num_bits != 3 && num_bits != 4) return FB_ERR_BITS;
    return fb::qgemm_max_mb(num_bits);
}

const char* flute_b200_dispatch_name(int M, int num_bits, int dtype) {
    if (num_bits != 2 && num_bits != 3 && num_bits != 4) return "unsupported";
    return fb::qgemm_dispatch_name(M, num_bits, dtype == FLUTE_B200_BF16);
}

int flute_b200_decode_grid(long long total_stages, int k_iters, int num_sms, int num_bits) {
    if (total_stages < 1 || k_iters < 1 || num_sms < 1) return 0;
    const int max_grid = num_sms > 16 ? num_sms - 4 : num_sms;
    // 2 / 4 bits: the decode kernel (may halve single-stage shares); 3 bits: the general kernel
    return fb::decode_grid_for(total_stages, k_iters, max_grid, num_bits != 3);
}

const char* flute_b200_last_error(void) { return g_last_error; }

const char* flute_b200_error_string(int code) {
    switch (code) {
        case FB_OK: return "ok";
        case FB_ERR_BITS: return "unsupported num_bits";
        case FB_ERR_GROUP: return "unsupported group_size";
        case FB_ERR_DTYPE: return "unsupported dtype";
        case FB_ERR_SHAPE: return "invalid shape";
        case FB_ERR_TILE_P: return "invalid tile_P";
        case FB_ERR_WORKSPACE: return "workspace too small";
        case FB_ERR_LAUNCH: return "CUDA launch failure";
        case FB_ERR_DRIVER: return "driver entry point unavailable";
        case FB_ERR_TENSORMAP: return "tensor-map encode failure";
        case FB_ERR_NULL: return "null pointer";
        case FB_ERR_DEVICE: return "invalid or unsupported device";
        case FB_ERR_INTERNAL: return "internal error";
        case FB_ERR_HADAMARD: return "invalid hadamard size";
        case FB_ERR_KERNEL: return "kernel trapped";
        default: return "unknown error";
    }
}

int flute_b200_version(void) { return FLUTE_B200_VERSION; }

void flute_b200_set_timeout_ms(long ms) { g_timeout_ms = ms; }

void flute_b200_set_variant(int variant) {
    if (variant < 0) { g_variant = -1; g_ablate = 0; g_l2_prefetch = -1; g_grid_slack = 0; return; }
    g_variant = variant & 0xff;
    if (g_variant == 0xff) g_variant = -1;
    g_ablate = (variant >> 8) & 0xff;   // undocumented perf-ablation bits, tools/microbench.py only
    g_l2_prefetch = ((variant >> 16) & 0xff) - 1;   // tools only: 0 = engine's choice, n + 1 = prefetch n stages
    g_grid_slack = (variant >> 24) & 0x7f;          // tools only: SMs left idle per launch
}

void flute_b200_set_trace_buffer(void* device_ptr) { g_trace = static_cast<unsigned long long*>(device_ptr); }

int flute_b200_check(int device) {
    if (device < 0 || device >= kMaxDevices) return fail(FB_ERR_DEVICE, "device index %d out of range", device);
    fb::Diag* d = g_dev[device].diag_host;
    if (d == nullptr || d->code == 0) return FB_OK;
    // sites 1-10: general kernel (ptx.cuh DiagSite); 21-28: decode kernel (DSITE_*); 41-48: prefill kernel (PSITE_*)
    static const char* general[] = {"?", "producer:empty", "dequant:full", "dequant:a_empty", "dequant:scale",
                                    "dequant:acc_full", "mma:full", "mma:a_full", "mma:acc_empty", "scale:empty", "final"};
    static const char* decode[] = {"decode/dequant:full", "decode/dequant:a_empty", "decode/apply:group_sums",
                                   "decode/apply:scales_full", "decode/producer|activations:empty", "decode/scales:empty",
                                   "decode/mma:a_full", "decode/mma:p_empty"};
    static const char* prefill[] = {"prefill/dequant:full", "prefill/dequant:a_slot", "prefill/epilogue:acc_full",
                                    "prefill/dequant:scales_full", "prefill/producer:empty", "prefill/scales:empty",
                                    "prefill/mma:a_full", "prefill/mma:acc_empty"};
    int site = d->site;
    const char* name = "?";
    if (site >= 0 && site <= 10) name = general[site];
    else if (site >= 21 && site <= 28) name = decode[site - 21];
    else if (site >= 41 && site <= 48) name = prefill[site - 41];
    int rc = fail(FB_ERR_KERNEL, "kernel barrier timeout: block %d warp %d waiting at %s(site %d)[%d] parity %d iter %d", d->block,
                  d->warp, name, site, d->index, d->parity, d->iter);
    d->code = 0;
    return rc;
}

}  // extern "C"
