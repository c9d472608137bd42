// Decode-shaped LUT-quantized GEMM for sm_100a (M <= 16, 2- and 4-bit):
//   D[M,N] = A[M,K] . W_hat[K,N],   W_hat[k,n] = table2[code(k/2,n)].{lo,hi} * S[n, k/group]
//
// Same role as qgemm_sm100.cu (replaces flute/csrc/qgemm_kernel.hpp:24-939 for the reference's
// M <= 16 templates, tile_scheduler_utils.hpp:58-1058, packbits_utils.hpp:24-144) but organised
// around the two things that bound a memory-bound GEMV on this part: issue slots and shared-memory
// wavefronts per streamed byte.
//
//   * The group scale is NOT multiplied into every weight.  Raw table2 pairs go straight from the
//     lane-replicated LUT into tensor memory; tcgen05.mma forms the per-GROUP partial sums
//     P_g[n,m] = sum_{k in g} T[q[k,n]] * A[m,k] (fp32, in TMEM) and the scale is applied once per
//     (n, group) on the accumulator:  acc[n,m] += S[n,g] * P_g[n,m]  (fp32 FFMA in registers).
//     Per streamed pair that leaves 1 PRMT + 1 LDS (+ 1/4 STTM) instead of PRMT + LDS + HFMA2.
//     Rounding: the reference rounds T*S to T before the MMA (packbits_utils.hpp:105,139); here the
//     product stays exact in fp32.  With A = I the result is round_T(T*S) -- bit-identical to the
//     reference's reconstruction -- and for general A it differs by less than the reference's own
//     tolerance (tests/kernel.py: 2e-3 fp16 / 1.1e-2 bf16).  See DESIGN.md "Decode numerics".
//   * All dequantiser warps convert quarter-row pieces (8 k-pairs per call) of the stages in flight, so the
//     latency of a stage is one piece, not one 128x32 chunk per warp, and a CTA that owns only three or four
//     stages (4096x4096 over 148 SMs) still uses every warp.
//   * Scales ([tile columns] x [8 groups] blocks) and activation rows are copied with cp.async by their own
//     warps; only the packed weights use TMA (a second TMA box per stage costs 18 % of the streaming rate,
//     profiles/r01_probe_tma_stream.log).  Accumulators live in registers of the apply warps, so there is no
//     TMEM accumulator hand-off; one tcgen05.commit per stage signals "smem stage free", "TMEM A slot free" and
//     "group sums ready" at once.
//
// One CTA per SM (214 KB shared memory, 512 TMEM columns).  A half-SM footprint of this kernel (two CTAs of
// CONSECUTIVE launches co-resident under programmatic dependent launch, 32 KB folded LUT, 3-stage ring, 40 registers)
// was built and measured in round 2 and removed again: co-residency works, but 113 KB leaves a 3-stage ring, the
// running CTA then has ~1.5 weight tiles in flight and streams at 1.6 TB/s instead of 3.2 (gpurun r02a/r02b,
// DESIGN.md section 3.1).  What a launch can hide behind its predecessor is done inside this footprint instead:
// weights, scales and the LUT never wait for the previous kernel (static weights), the dequantisers fill all three
// TMEM A slots before the activations exist, and the split-K fix-up runs in its own warp off the streaming path.
//
// Warp roles (persistent over a contiguous Stream-K range of (tile, k) stages).  DQ = dequantiser warps
// (16); 832 threads for M <= 4, 928 for the 16-accumulator 5 <= M <= 16 variant:
//   warps 0..DQ-1    dequantisers: warp w owns TMEM lane quarter w%4 and a fixed set of 16-byte quads of its row
//   warp DQ          TMA producer (packed weights, one 128-row x 64-k box per stage; optional L2 prefetch ahead)
//   warps DQ+1,DQ+3  tcgen05.mma issuers (alternate scale groups; warp DQ+1 also allocates TMEM)
//   warp DQ+2        activation rows of every stage (16-byte cp.async, up to three stages ahead)
//   warps DQ+4..     4 (8 for M > 4) scale application (acc += S * P_g), epilogue and split-K fix-up
//   next warp        split-K fix-up: arrival counter, last-arriver conversion of the fp32 partial sums to D
//   last warp        scale blocks (cp.async); for M > 4 the activation warp does this
#include "ptx.cuh"
#include "qgemm_sm100.h"

#include <cuda.h>

namespace fb {
namespace dec {

// Optional per-role cycle accounting (profiling build only: python flute_b200/build.py --profile).
#ifdef FB_PROFILE
#define DPROF_DECL(...) long long __VA_ARGS__
#define DPROF_T0(t) long long t = clock64()
#define DPROF_ADD(acc, t) do { long long _n = clock64(); acc += _n - t; t = _n; } while (0)
#define DPROF_OUT(slot, v) do { if (p.trace != nullptr) p.trace[blockIdx.x * 48 + (slot)] = (unsigned long long)(v); } while (0)
#define DTRACE_ON(p) ((p).trace != nullptr)
#define DABLATE(p, bit) (((p).ablate & (bit)) != 0)
#else
#define DPROF_DECL(...)
#define DPROF_T0(t)
#define DPROF_ADD(acc, t)
#define DPROF_OUT(slot, v)
// (compiling the per-CTA time stamps and ablation switches out of the production build as well was measured: small shapes
// 1-2 % faster, large ones 3-5 % slower, bench 725 -> 707 tok/s -- gpurun r02ab2 vs r02ab4 -- so they stay run-time checks)
#define DTRACE_ON(p) ((p).trace != nullptr)
#define DABLATE(p, bit) (((p).ablate & (bit)) != 0)
#endif

// NJ      pair fields per 32-bit word (accumulated output columns per packed row)
// CK2     k-pairs per TMEM chunk (an A slot = NJ*CK2 columns);  CPS chunks per 64-k stage
// DQ      dequantiser warps;  DQG sets of them that convert alternate stages
// LUTB    bytes of the lane-replicated pair LUT
template <int BITS>
struct DCfg;
template <>
struct DCfg<4> {
    static constexpr int NJ = 4, CK2 = 32, CPS = 1, LUTN = 256, A_SLOTS = 3, P_SLOTS = 2;
    static constexpr int DQ = 16, DQG = 2, SC_SLOTS = 3, MAX_STAGES = 10, TMEM_COLS = 512, LUTB = 256 * 256;
};
template <>
struct DCfg<2> {
    static constexpr int NJ = 8, CK2 = 16, CPS = 2, LUTN = 16, A_SLOTS = 2, P_SLOTS = 2;
    static constexpr int DQ = 16, DQG = 1, SC_SLOTS = 3, MAX_STAGES = 10, TMEM_COLS = 512, LUTB = 16 * 256;
};

// apply warps: 4 (each all NJ fields) for M <= 4; 8 (two field halves) when a field needs 16 accumulators
__host__ __device__ constexpr int apply_warps(int mc) { return mc > 4 ? 8 : 4; }
// M <= 4: one more warp copies the scale blocks; with 8 apply warps the activation warp does
__host__ __device__ constexpr bool has_scale_warp(int mc) { return mc <= 4; }
// dequantisers + producer, 2 MMA issuers, activation warp + apply warps + fix-up warp (+ scale warp)
__host__ __device__ constexpr int threads_for(int dq, int mc) {
    return (dq + 4 + apply_warps(mc) + 1 + (has_scale_warp(mc) ? 1 : 0)) * 32;
}
constexpr uint32_t kSmemBudget = 232448u;
constexpr int kMaxStagesAny = 10;
constexpr int kMaxScSlots = 3;
constexpr int kWBytes = 128 * 128;   // packed-weight part of a stage: 128 rows x 64 k x 2 B
constexpr int kBBytes = 16 * 128;    // activation part: 16 rows x 64 k x 2 B
constexpr int kStageBytes = kWBytes + kBBytes;
constexpr int kMb = 16;              // MMA N

struct Ctl {
    uint64_t full[kMaxStagesAny];       // packed weights of the stage have landed (TMA)
    uint64_t act_full[kMaxStagesAny];   // activation rows of the stage have landed (cp.async)
    uint64_t empty[kMaxStagesAny];
    uint64_t a_full[3];
    uint64_t a_empty[3];
    uint64_t p_full[2];
    uint64_t p_empty[2];
    uint64_t sc_full[kMaxScSlots];
    uint64_t sc_empty[kMaxScSlots];
    uint64_t tmem_ready;     // the allocating warp arrives once the TMEM base address is in tmem_base
    uint64_t fix_full[2];    // partial-tile hand-over apply warps -> fix-up warp (a contiguous range has <= 2 partial segments)
    uint32_t tmem_base;
};

struct DecodeParams {
    const uint16_t* A;
    const uint8_t* Q;      // packed weights [P, K] int16 (also behind the tensor map; raw pointer for the entry prefetch)
    const uint16_t* S;
    const uint32_t* table2;
    uint16_t* D;
    uint8_t* workspace;
    Diag* diag;
    unsigned long long* trace;
    unsigned long long timeout_ns;   // barrier-wait bound (flute_b200_set_timeout_ms); 0 = unbounded
    int M, N, K, G;
    int P;               // packed rows = N / 16 * bits
    int tile_p;
    int gshift;          // log2(group_size / 64): stages per group
    int n_tiles, k_iters;
    int stages;
    int tma_scales;      // scale rows of a block are 16-byte aligned (G % 8 == 0): cp.async, else scalar loads
    int static_weights;
    int ablate;          // perf ablation (tools only): 1 no MMA issue, 2 no dequant pieces, 4 no scale/accumulate
    int l2_prefetch;     // stages the producer prefetches into L2 ahead of its shared-memory ring (0 = off)
    uint32_t partial_offset;
};

// Tensor-parallel extension of the parameters: a separate kernel argument that only the TP instantiation carries, so the
// single-GPU kernel's parameter block, code and registers are exactly what they are without it.  Measured on one box
// (gpurun r02ab / r02ab2 / r02ab3, us per launch qkv / o / gate_up / down): fields folded into DecodeParams behind run-time
// `if (p.tp > 1)` 9.86 / 8.92 / 19.00 / 12.63; this split 8.07 / 7.07 / 17.24 / 11.48; additionally squeezing DecodeParams
// under 256 bytes with byte-sized fields 8.15 / 7.13 / 17.67 / 11.67 (no gain: not kept).
struct TpParams {
    // Tensor-parallel column shard with the exchange fused in: the tile writer stores its [M, tile] slice into EVERY rank's
    // gathered buffer (peer-mapped pointers, NVLink) at column offset rank * N -- as a plain T image (for readers outside
    // this engine, optional) and as ONE 8-byte word {value, sequence number} per element (the NCCL "LL" idea): an aligned
    // 8-byte store is single-copy atomic, so a reader that sees the expected sequence number has the value -- no fence, no
    // separate flag, one NVLink one-way trip.  The activation warp of the consuming launch reads A from that word image
    // and spins per word.  sequence = (epoch - 1) * uses + call + 1 (grows for ever; the buffers start zeroed).
    // All accesses here are relaxed at GPU scope: peer stores travel over NVLink into the destination's L2, where the
    // destination's L1-bypassing loads find them.  System-scope fences / reductions (arrival counters for readers that
    // are not qgemm launches) live in the two tiny kernels below, not in this one.
    int tp, rank;
    int n_total;                  // row stride of the gathered output (= tp * N)
    uint16_t* out_peers[8];       // every rank's plain image [M, n_total]
    long long ll_delta;           // byte distance from a rank's plain image to its word image (same on every rank)
    int write_plain;              // also keep the plain image current (0: every reader is a qgemm_tp launch)
    unsigned out_uses, out_call;
    const uint2* in_ll;           // word image A is read from (nullptr: plain A through cp.async)
    int in_ll_stride;             // elements per row of that image
    unsigned in_uses, in_call;
    const unsigned* epoch;        // device word: step number (>= 1)
};
template <bool TP>
struct TpArg {
    TpParams v;
};
template <>
struct TpArg<false> {};

enum : int { DSITE_FULL = 21, DSITE_AEMPTY, DSITE_PFULL, DSITE_SCFULL, DSITE_EMPTY, DSITE_SCEMPTY, DSITE_AFULL, DSITE_PEMPTY };

static __device__ __noinline__ void wait_timeout(Diag* diag, int site, uint32_t bar, uint32_t parity, int iter) {
    if (diag != nullptr) {
        diag->block = blockIdx.x;
        diag->warp = threadIdx.x >> 5;
        diag->site = site;
        diag->index = (int)bar;
        diag->parity = (int)parity;
        diag->iter = iter;
        diag->code = 1;
        __threadfence_system();
    }
    __trap();
}

// Lean bounded wait: one try_wait on the fast path; the bound (p.timeout_ns, 0 = none) is checked out of line
// every 1024 failed probes (each failed try_wait already suspends the warp for a hardware-defined interval).
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity, const DecodeParams& p, int site, int iter = 0) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3ff) == 0 && p.timeout_ns != 0) {
            const uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > p.timeout_ns) wait_timeout(p.diag, site, bar, parity, iter);   // trap, don't hang
        }
    }
}

// One elected lane of a converged warp.  tcgen05.mma / commit / TMA are uniform-datapath instructions: under
// `if (lane == 0)` ptxas cannot prove uniformity and wraps each one in an ELECT ... BRA.U.ANY waterfall loop
// (~210 cycles per MMA measured, tools/mma_rate_probe.cu); under elect.sync it issues them directly.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// (byte J of w) << 8 | lane4  -- the LUT offset of pair-code byte J for this lane (256-byte entry stride)
template <int J>
__device__ __forceinline__ uint32_t code_lane(uint32_t w, uint32_t lane4) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(lane4), "n"(0x6504 + (J << 4)));
    return r;
}
__device__ __forceinline__ void tmem_st_x4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_x1(uint32_t taddr, uint32_t (&r)[1]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
template <int MC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[MC]) {
    if constexpr (MC == 1) tmem_ld_x1(taddr, r);
    else if constexpr (MC == 4) tmem_ld_x4(taddr, r);
    else tmem_ld_32x32b_x16(taddr, r);
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ int atom_add_acq_rel(int* addr, int v) {
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu_u32(const unsigned* addr) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_sys_add_u32(unsigned* addr, unsigned v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* addr) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
// Pull one packed-weight box into L2 without a shared-memory destination.
__device__ __forceinline__ void tma_prefetch_l2_2d(const void* tmap, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1) : "memory");
}

template <int BITS, int NJ>
__device__ __forceinline__ int n_local(int L, int j, int tile_p) {
    if (tile_p == 32) return (L >> 5) * (NJ * 32) + j * 32 + (L & 31);
    return (L >> 6) * (NJ * 64) + j * 64 + (L & 63);
}

struct Range {
    int it0, it1;
};
__device__ __forceinline__ Range cta_range(int total, int b, int grid) {
    const int base = total / grid, rem = total - base * grid;
    Range r;
    r.it0 = b * base + min(b, rem);
    r.it1 = r.it0 + base + (b < rem ? 1 : 0);
    return r;
}
__device__ __forceinline__ int cta_of(int total, int it, int grid) {
    const int base = total / grid, rem = total - base * grid;
    const int thr = rem * (base + 1);
    return it < thr ? it / (base + 1) : rem + (it - thr) / base;
}

// One piece of a dequantiser warp's work on row L of a stage.
template <int BITS>
struct Piece;
template <>
struct Piece<4> {
    // two 16-byte quads (8 consecutive k-pairs) of row L -> 4 fields x 8 TMEM columns, field blocks FSTRIDE columns apart
    template <int FSTRIDE>
    static __device__ __forceinline__ void run(uint32_t row, int pq0, int pq1, uint32_t lut, uint32_t lane4, uint32_t tcol) {
        const uint4 v0 = lds128(row + (uint32_t)(pq0 << 4));
        const uint4 v1 = lds128(row + (uint32_t)(pq1 << 4));
        const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t r[4][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            r[0][i] = lds32(lut + code_lane<0>(w[i], lane4));
            r[1][i] = lds32(lut + code_lane<1>(w[i], lane4));
            r[2][i] = lds32(lut + code_lane<2>(w[i], lane4));
            r[3][i] = lds32(lut + code_lane<3>(w[i], lane4));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) tmem_st_x8(tcol + j * FSTRIDE, r[j]);
    }
};
template <>
struct Piece<2> {
    static __device__ __forceinline__ void run(uint32_t row, int pq, uint32_t lut, uint32_t lane4, uint32_t tcol) {
        const uint4 v = lds128(row + (uint32_t)(pq << 4));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t r[8][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t lo = w[i] & 0x0f0f0f0fu;          // nibbles 0, 2, 4, 6 as bytes
            const uint32_t hi = (w[i] >> 4) & 0x0f0f0f0fu;   // nibbles 1, 3, 5, 7
            r[0][i] = lds32(lut + code_lane<0>(lo, lane4));
            r[1][i] = lds32(lut + code_lane<0>(hi, lane4));
            r[2][i] = lds32(lut + code_lane<1>(lo, lane4));
            r[3][i] = lds32(lut + code_lane<1>(hi, lane4));
            r[4][i] = lds32(lut + code_lane<2>(lo, lane4));
            r[5][i] = lds32(lut + code_lane<2>(hi, lane4));
            r[6][i] = lds32(lut + code_lane<3>(lo, lane4));
            r[7][i] = lds32(lut + code_lane<3>(hi, lane4));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) tmem_st_x4(tcol + j * 16, r[j][0], r[j][1], r[j][2], r[j][3]);
    }
};

template <bool BF16>
__device__ __forceinline__ float scale_to_f32(uint32_t s16) {
    if constexpr (BF16) return __uint_as_float(s16 << 16);
    else return __half2float(__ushort_as_half((unsigned short)s16));
}

template <int BITS, bool BF16, int MC, bool TP>
__global__ void __launch_bounds__(threads_for(DCfg<BITS>::DQ, MC), 1)
qgemm_decode_kernel(const __grid_constant__ CUtensorMap tmap_w, const DecodeParams p, const TpArg<TP> tpa) {
    using F = DCfg<BITS>;
    constexpr int NJ = F::NJ, CK2 = F::CK2, CPS = F::CPS;
    constexpr int AS = F::A_SLOTS, PS = F::P_SLOTS;
    constexpr int kDqWarps = F::DQ;
    constexpr int kProducerWarp = kDqWarps;
    constexpr int kMmaWarp = kDqWarps + 1;
    constexpr int kActWarp = kDqWarps + 2;          // activation rows (cp.async)
    constexpr int kMmaWarpB = kDqWarps + 3;         // second MMA issuer: flush groups alternate between the two
    constexpr int kApplyWarp0 = kDqWarps + 4;       // scale/accumulate/epilogue warps, lane quarter = warp & 3
    constexpr int kScSlots = F::SC_SLOTS;
    constexpr int TN = NJ * 128;
    constexpr uint32_t kScBytes = TN * 16;
    constexpr uint32_t kACols = NJ * CK2;          // TMEM columns of one A slot
    constexpr uint32_t kPCol0 = AS * kACols;       // P slots sit after the A slots
    constexpr uint32_t kPCols = NJ * kMb;
    constexpr int kApplyWarps = apply_warps(MC);
    constexpr bool kScaleWarpExists = has_scale_warp(MC);
    constexpr int kFixWarp = kApplyWarp0 + kApplyWarps;     // split-K fix-up, off the streaming path
    constexpr int kScaleWarp = kFixWarp + 1;                // exists iff kScaleWarpExists
    constexpr int NFA = NJ / (kApplyWarps / 4);   // fields per apply warp
    static_assert(kPCol0 + PS * kPCols <= (uint32_t)F::TMEM_COLS, "TMEM budget");
    static_assert((kApplyWarp0 & 3) == 0, "apply warp w must own TMEM lane quarter w & 3");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t ring = smem_base;
    const uint32_t sc_smem = ring + p.stages * kStageBytes;
    const uint32_t lut = sc_smem + kScSlots * kScBytes;
    Ctl* ctl = reinterpret_cast<Ctl*>(smem_gen + (lut + F::LUTB - smem_base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int grid = gridDim.x;
    const int total = p.n_tiles * p.k_iters;
    const Range rg = cta_range(total, blockIdx.x, grid);
    // (a {group 64, tile_P 32, static weights, aligned scales} instantiation with these four folded to constants measured 1 %
    // faster -- gpurun r02ab5 -- and was not kept)
    const int r_gshift = p.gshift, r_tile_p = p.tile_p, r_static = p.static_weights, r_tma_scales = p.tma_scales;
    const int spg_mask = (1 << r_gshift) - 1;

    if (DTRACE_ON(p) && threadIdx.x == 0) p.trace[blockIdx.x * 48 + 0] = globaltimer_ns();
    if (!r_static) pdl_wait_prior_grids();
    // Entry prefetch: the first TMA box cannot be requested before the barriers exist and the tensor map has been
    // fetched (~1 us after launch), and then pays a cold DRAM + page-walk latency on top.  The row addresses are plain
    // arithmetic, so every dequantiser thread asks L2 for one 128-byte row piece of the CTA's first ring-full of
    // stages right away; the TMA loads then hit L2.
    if (warp < kDqWarps && rg.it1 > rg.it0) {
        const int npf = min(rg.it1 - rg.it0, p.stages);
        const int r = threadIdx.x & 127;
        for (int s0 = threadIdx.x >> 7; s0 < npf; s0 += kDqWarps / 4) {
            const int it = rg.it0 + s0;
            const int tile = it / p.k_iters;
            const int k = it - tile * p.k_iters;
            const int prow = tile * 128 + r;
            if (prow < p.P) {
                const uint8_t* addr = p.Q + ((size_t)prow * p.K + (size_t)k * 64) * 2;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(addr) : "memory");
            }
        }
    }

    if (warp == kProducerWarp && lane == 0) {
        tma_prefetch_desc(&tmap_w);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&ctl->full[s]), 1);       // TMA (expect_tx): what the dequantisers wait for
            mbar_init(smem_u32(&ctl->act_full[s]), 1);   // the activation warp: what the MMA issuers wait for
            mbar_init(smem_u32(&ctl->empty[s]), 1);
        }
        for (int s = 0; s < AS; ++s) {
            mbar_init(smem_u32(&ctl->a_full[s]), kDqWarps / F::DQG);
            mbar_init(smem_u32(&ctl->a_empty[s]), 1);
        }
        for (int s = 0; s < PS; ++s) {
            mbar_init(smem_u32(&ctl->p_full[s]), 1);
            mbar_init(smem_u32(&ctl->p_empty[s]), kApplyWarps);
        }
        for (int s = 0; s < kScSlots; ++s) {
            mbar_init(smem_u32(&ctl->sc_full[s]), 32);   // one (deferred) arrival per lane of the activation/scale warp
            mbar_init(smem_u32(&ctl->sc_empty[s]), kApplyWarps);
        }
        mbar_init(smem_u32(&ctl->tmem_ready), 1);
        for (int s = 0; s < 2; ++s) mbar_init(smem_u32(&ctl->fix_full[s]), kApplyWarps);
        mbar_fence_init();
    }
    // First sync: barriers visible.  The producer starts streaming right after it, the dequant warps build the LUT;
    // tensor memory is allocated AFTER it by one warp, and only the TMEM users wait for the address (tmem_ready):
    // should another kernel's CTA still hold this SM's columns the allocation blocks until it exits, and everything
    // that does not touch TMEM proceeds meanwhile.
    __syncthreads();
    pdl_launch_dependents();
    if (warp == kMmaWarp) {
        tmem_alloc(smem_u32(&ctl->tmem_base), F::TMEM_COLS);
        tmem_relinquish();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ctl->tmem_ready));
    }
    auto tmem_base_when_ready = [&]() -> uint32_t {
        wait(smem_u32(&ctl->tmem_ready), 0u, p, DSITE_AFULL, -1);
        tc_fence_after();
        return *reinterpret_cast<volatile uint32_t*>(&ctl->tmem_base);
    };

    // D[m, n] = v -- locally, or (tensor parallel) into every rank's gathered buffer at this rank's column offset, plain
    // and as a {value, sequence} word for the low-latency readers
    auto store_out = [&](int m, int n, uint16_t v, unsigned seq) {
        if constexpr (!TP) {
            p.D[(size_t)m * p.N + n] = v;
        } else {
            const TpParams& t = tpa.v;
            const size_t off = (size_t)m * t.n_total + (size_t)t.rank * p.N + n;
#pragma unroll 1
            for (int r = 0; r < t.tp; ++r) {
                // the plain image is for readers outside this engine; calls whose output only feeds other qgemm_tp calls
                // (write_plain == 0) keep just the word image and halve their NVLink stores
                if (t.write_plain) t.out_peers[r][off] = v;
                asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(reinterpret_cast<uint2*>(reinterpret_cast<char*>(t.out_peers[r]) + t.ll_delta) + off),
                             "r"((uint32_t)v), "r"(seq) : "memory");
            }
        }
    };
    auto out_sequence = [&]() -> unsigned {
        if constexpr (TP) return (ld_relaxed_gpu_u32(tpa.v.epoch) - 1u) * tpa.v.out_uses + tpa.v.out_call + 1u;
        else return 0u;
    };
    auto tp_in_ll = [&]() -> const uint2* {
        if constexpr (TP) return tpa.v.in_ll;
        else return nullptr;
    };
    auto tp_in_ll_stride = [&]() -> int {
        if constexpr (TP) return tpa.v.in_ll_stride;
        else return 0;
    };
    auto tp_expected_sequence = [&]() -> unsigned {
        if constexpr (TP) return (ld_relaxed_gpu_u32(tpa.v.epoch) - 1u) * tpa.v.in_uses + tpa.v.in_call + 1u;
        else return 0u;
    };

    // One step of the scale-block schedule (called once per stage, in stage order, by ONE warp): when stage (tile, k)
    // starts a new block of 8 groups, copy [tile columns] x [8 groups] (16 bytes per row) into the next scale slot.
    // cp.async, not TMA: as a TMA box the 512 sixteen-byte rows cost the producer ~2000 cycles of issue time per
    // block (measured), during which no weight tile could be requested.
    auto scale_step = [&](int tile, int k, int& nb, int& last_blk) {
        const int blk = (k >> r_gshift) >> 3;
        if (blk == last_blk) return;
        const int slot = nb % kScSlots;
        const uint32_t par = ((nb / kScSlots) & 1) ^ 1u;
        wait(smem_u32(&ctl->sc_empty[slot]), par, p, DSITE_SCEMPTY);
        const uint32_t dst = sc_smem + slot * kScBytes;
        if (r_tma_scales) {
#pragma unroll 4
            for (int r = lane; r < TN; r += 32) {
                const int n = tile * TN + r;
                if (n < p.N) {
                    const uint16_t* src = p.S + (size_t)n * p.G + blk * 8;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + r * 16), "l"(src) : "memory");
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&ctl->sc_full[slot])) : "memory");
        } else {
            uint16_t* d16 = reinterpret_cast<uint16_t*>(smem_gen + (dst - smem_base));
            for (int r = lane; r < TN; r += 32) {
                const int n = tile * TN + r;
#pragma unroll
                for (int gi = 0; gi < 8; ++gi) {
                    const int g = blk * 8 + gi;
                    d16[r * 8 + gi] = (n < p.N && g < p.G) ? __ldg(p.S + (size_t)n * p.G + g) : (uint16_t)0;
                }
            }
            mbar_arrive(smem_u32(&ctl->sc_full[slot]));
        }
        last_blk = blk;
        ++nb;
    };

    if (warp == kProducerWarp) {
        // =============================== TMA producer ===============================
        // Packed weights by TMA, one 128-row x 64-k box per stage.  Static data: nothing here waits for the previous
        // kernel.  With l2_prefetch = PF > 0 the boxes of the next PF stages beyond the ring are pulled into L2 first
        // (no shared-memory destination); off by default (no gain with a 9-stage ring, gpurun r02a), kept as a tuning
        // knob for short rings (force_stages).
        if (rg.it1 > rg.it0) {
            const uint64_t pol_w = policy_evict_first();
            const int n_it = rg.it1 - rg.it0;
            int tile = rg.it0 / p.k_iters;
            int k = rg.it0 - tile * p.k_iters;
            int stage = 0;
            uint32_t ephase = 1;           // parity to wait on `empty` (first pass: free)
            // L2 prefetch cursor: stage index pf_i with coordinates (pf_tile, pf_k); runs p.stages + PF ahead of i
            const int pf_ahead = p.l2_prefetch > 0 ? p.stages + p.l2_prefetch : 0;
            int pf_i = min(n_it, p.stages);
            int pf_tile = tile, pf_k = k;
            if (pf_ahead > 0) {
                pf_k += pf_i;
                while (pf_k >= p.k_iters) { pf_k -= p.k_iters; ++pf_tile; }
            }
            DPROF_DECL(pw_sc = 0, pw_empty = 0, pw_w = 0, pw_a = 0);
            DPROF_T0(pt);
            for (int i = 0; i < n_it; ++i) {
                DPROF_ADD(pw_sc, pt);
                wait(smem_u32(&ctl->empty[stage]), ephase, p, DSITE_EMPTY);
                DPROF_ADD(pw_empty, pt);
                if (elect_one()) {
                    const uint32_t bar = smem_u32(&ctl->full[stage]);
                    mbar_arrive_expect_tx(bar, kWBytes);
                    tma_load_2d(ring + stage * kStageBytes, &tmap_w, bar, k * 64, tile * 128, pol_w);
                    const int pf_end = pf_ahead > 0 ? min(n_it, i + 1 + pf_ahead) : 0;
                    while (pf_i < pf_end) {       // first iterations: catch up; steady state: one box per stage
                        tma_prefetch_l2_2d(&tmap_w, pf_k * 64, pf_tile * 128);
                        ++pf_i;
                        if (++pf_k == p.k_iters) { pf_k = 0; ++pf_tile; }
                    }
                }
                __syncwarp();
                DPROF_ADD(pw_w, pt);
                DPROF_ADD(pw_a, pt);
                if (++stage == p.stages) { stage = 0; ephase ^= 1u; }
                if (++k == p.k_iters) { k = 0; ++tile; }
            }
            if (lane == 0) { DPROF_OUT(8, pw_sc); DPROF_OUT(9, pw_empty); DPROF_OUT(10, pw_w); DPROF_OUT(11, pw_a); DPROF_OUT(12, n_it); }
        }
    } else if (warp == kMmaWarp || warp == kMmaWarpB) {
        // =============================== MMA issuers ================================
        // Two issuing warps take alternate flush groups (all stages of one scale group go to the same warp, so
        // the first MMA of a group overwrites and the rest accumulate in issue order).  Group f uses P slot
        // f & 1, i.e. each warp owns one P slot.  A single issuer spends as long waiting on barriers and
        // committing as the tensor pipe spends on the 16 N=16 MMAs of a stage (~16 cycles each, measured:
        // tools/mma_rate_probe.cu), which made it the slowest role of the pipeline.
        if (rg.it1 > rg.it0) {
            const int mine = (warp == kMmaWarp) ? 0 : 1;
            const uint32_t tmem = tmem_base_when_ready();
            const uint32_t idesc = make_idesc_f16(BF16, 128, kMb);
            int stage = 0;
            uint32_t sphase = 0;           // parity of the ring pass `stage` is in
            int aslot = 0;
            uint32_t aphase = 0;
            int f = 0;                     // flush groups started so far
            bool grp_first = true;
            DPROF_DECL(mw_afull = 0, mw_pempty = 0, mw_issue = 0);
            DPROF_T0(mt);
            for (int it = rg.it0; it < rg.it1;) {
                const int tile = it / p.k_iters;
                const int kb = it - tile * p.k_iters;
                const int ke = min(p.k_iters, kb + (rg.it1 - it));
                for (int k = kb; k < ke; ++k) {
                    const bool my_group = ((f & 1) == mine);
                    if (my_group) {
                        // The dequantisers only need the weights (full[stage]), so the first A slots fill while the
                        // previous kernel is still running; the activation rows are what the MMAs themselves wait for.
                        wait(smem_u32(&ctl->act_full[stage]), sphase, p, DSITE_FULL, 1);
                        const uint64_t bdesc = make_smem_desc_sw128(ring + stage * kStageBytes + kWBytes);
                        const int pslot = mine;
#pragma unroll
                        for (int c = 0; c < CPS; ++c) {
                            wait(smem_u32(&ctl->a_full[aslot]), aphase, p, DSITE_AFULL);
                            DPROF_ADD(mw_afull, mt);
                            if (grp_first) wait(smem_u32(&ctl->p_empty[pslot]), (((uint32_t)f >> 1) & 1u) ^ 1u, p, DSITE_PEMPTY);
                            DPROF_ADD(mw_pempty, mt);
                            tc_fence_after();
                            if (DTRACE_ON(p) && lane == 0 && mine == 0 && f == 0 && c == 0 && k == kb) p.trace[blockIdx.x * 48 + 3] = globaltimer_ns();
                            if (elect_one()) {
                                const uint32_t a_base = tmem + aslot * kACols;
                                const uint32_t d_base = tmem + kPCol0 + pslot * kPCols;
                                if (!DABLATE(p, 1))
#pragma unroll
                                for (int j = 0; j < NJ; ++j) {
#pragma unroll
                                    for (int kk = 0; kk < CK2 / 8; ++kk) {
                                        tc_mma_ts(d_base + j * kMb, a_base + j * CK2 + kk * 8,
                                                  bdesc + (uint64_t)(((c * CK2 + kk * 8) * 4) >> 4), idesc,
                                                  (grp_first && kk == 0) ? 0u : 1u);
                                    }
                                }
                                // one completion signal per stage: frees the smem stage (producer, activation warp), the
                                // TMEM A slot (dequantisers, AS/CPS stages later) and publishes the group sums (apply warps)
                                if (c == CPS - 1) tc_commit(smem_u32(&ctl->empty[stage]));
                                else tc_commit(smem_u32(&ctl->a_empty[aslot]));
                            }
                            __syncwarp();
                            grp_first = false;
                            if (++aslot == AS) { aslot = 0; aphase ^= 1u; }
                        }
                        DPROF_ADD(mw_issue, mt);
                    } else {
                        // Not mine, but stay in step with every a_full phase: a parity wait is only meaningful for
                        // the NEXT completion of a barrier, so this warp must not run a whole phase ahead of it.
#pragma unroll
                        for (int c = 0; c < CPS; ++c) {
                            wait(smem_u32(&ctl->a_full[aslot]), aphase, p, DSITE_AFULL);
                            if (++aslot == AS) { aslot = 0; aphase ^= 1u; }
                        }
                    }
                    const bool flush = (((k + 1) & spg_mask) == 0) || (k == ke - 1);
                    if (flush) { grp_first = true; ++f; }
                    if (++stage == p.stages) { stage = 0; sphase ^= 1u; }
                }
                it += ke - kb;
            }
            if (lane == 0 && mine == 0) { DPROF_OUT(14, mw_afull); DPROF_OUT(15, mw_pempty); DPROF_OUT(16, mw_issue); }
        }
    } else if (warp == kActWarp) {
        // =============================== activation rows ============================
        // The activation tile of a stage (M rows x 64 k inside a 16-row, 128-byte-swizzled K-major tile) is written
        // with 16-byte cp.async copies, D stages ahead, instead of a second TMA box: measured on B200
        // (tools/tma_stream_probe.cu) every cp.async.bulk.tensor costs the TMA unit ~160 cycles + ~2.7 per
        // 128-byte row, and a second box per 64-k stage (even with one in-bounds row) costs 18 % of the streaming
        // rate.  Rows >= M of every stage's tile are zeroed once and never written again.
        if (rg.it1 > rg.it0) {
            // Copies run D stages ahead of the `full` arrivals; the ring slot of stage i is only free once stage
            // i - stages has been consumed, which needs its arrival, so D must stay below the ring depth.
            const int D = min(3, p.stages - 1);
            const int n_it = rg.it1 - rg.it0;
            for (int s2 = 0; s2 < p.stages; ++s2) {
                uint4* z = reinterpret_cast<uint4*>(smem_gen + (ring + s2 * kStageBytes + kWBytes - smem_base));
#pragma unroll
                for (int i = 0; i < kBBytes / 16 / 32; ++i) z[i * 32 + lane] = make_uint4(0, 0, 0, 0);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            int tile = rg.it0 / p.k_iters;
            int k = rg.it0 - tile * p.k_iters;
            int nb = 0;                    // scale blocks issued
            int last_blk = -1;
            // Scales are static like the weights: when this warp also copies them, the first block goes out before
            // the wait for the previous kernel.
            if (!kScaleWarpExists && r_static) scale_step(tile, k, nb, last_blk);
            // Activations come from the previous kernel -- unless they are read from a {value, sequence} word image
            // (tensor parallel): each word says by itself whether it is this hand-over's, so the launch does not wait for
            // the producing grid to complete and flush (~1 us per launch), it consumes the words as they land.  Reuse of
            // an image is safe without the wait: a launch can only write its output once it has ALL of its input, i.e.
            // once every CTA of every rank's producing launch has finished reading that launch's own input.
            const bool a_from_words = TP && tp_in_ll() != nullptr;
            if (r_static && !a_from_words) pdl_wait_prior_grids();
            if (DTRACE_ON(p) && lane == 0) p.trace[blockIdx.x * 48 + 2] = globaltimer_ns();
            int stage = 0, astage = 0;
            uint32_t ephase = 1;
            if (TP && tp_in_ll() != nullptr) {
                // Low-latency path (tensor parallel): A is read from the {value, sequence} image of a gathered buffer.  Lane l
                // owns k-pair l of the stage: one 16-byte volatile load = two words; it spins until both carry this
                // hand-over's sequence number, then writes the pair into the swizzled tile.  No cp.async, no look-ahead:
                // the words arrive from the peers while the dequantisers are already filling TMEM.
                const unsigned expected = tp_expected_sequence();
                for (int i = 0; i < n_it; ++i) {
                    if (!kScaleWarpExists) scale_step(tile, k, nb, last_blk);
                    wait(smem_u32(&ctl->empty[stage]), ephase, p, DSITE_EMPTY);
                    const uint32_t bt = ring + stage * kStageBytes + kWBytes;
                    for (int r = 0; r < p.M; ++r) {
                        const uint2* src = tp_in_ll() + (size_t)r * tp_in_ll_stride() + (size_t)k * 64 + 2 * lane;
                        uint32_t d0, f0, d1, f1;
                        uint64_t t0 = 0;
                        uint32_t spins = 0;
                        for (;;) {
                            asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(d0), "=r"(f0), "=r"(d1), "=r"(f1) : "l"(src) : "memory");
                            if (f0 == expected && f1 == expected) break;
                            if ((++spins & 0xff) == 0 && p.timeout_ns != 0) {
                                const uint64_t now = globaltimer_ns();
                                if (t0 == 0) t0 = now;
                                else if (now - t0 > p.timeout_ns) wait_timeout(p.diag, DSITE_FULL, f0, expected, -4);
                            }
                        }
                        const uint32_t dst = bt + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
                        sts32(dst, (d0 & 0xffffu) | (d1 << 16));
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&ctl->act_full[stage]));
                    if (++stage == p.stages) { stage = 0; ephase ^= 1u; }
                    if (++k == p.k_iters) { k = 0; ++tile; last_blk = -1; }
                }
            } else {
            const int r0 = lane >> 3, c16 = lane & 7;          // this lane's row (mod 4) and 16-byte chunk
            const uint8_t* a_lane = reinterpret_cast<const uint8_t*>(p.A) + c16 * 16;
            for (int i = 0; i < n_it; ++i) {
                if (!kScaleWarpExists) scale_step(tile, k, nb, last_blk);
                wait(smem_u32(&ctl->empty[stage]), ephase, p, DSITE_EMPTY);
                const uint32_t bt = ring + stage * kStageBytes + kWBytes;
#pragma unroll
                for (int m0 = 0; m0 < MC; m0 += 4) {
                    const int r = m0 + r0;
                    if (r < p.M) {
                        const uint32_t dst = bt + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
                        const uint8_t* src = a_lane + ((size_t)r * p.K + (size_t)k * 64) * 2;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (i >= D) {
                    if (D == 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
                    else if (D == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
                    else asm volatile("cp.async.wait_group 1;" ::: "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&ctl->act_full[astage]));
                    if (++astage == p.stages) astage = 0;
                }
                if (++stage == p.stages) { stage = 0; ephase ^= 1u; }
                if (++k == p.k_iters) { k = 0; ++tile; last_blk = -1; }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            for (int i = max(0, n_it - D); i < n_it; ++i) {
                if (lane == 0) mbar_arrive(smem_u32(&ctl->act_full[astage]));
                if (++astage == p.stages) astage = 0;
            }
            }
        }
    } else if (kScaleWarpExists && warp == kScaleWarp) {
        // =============================== scale blocks ===============================
        if (rg.it1 > rg.it0) {
            const int n_it = rg.it1 - rg.it0;
            int tile = rg.it0 / p.k_iters;
            int k = rg.it0 - tile * p.k_iters;
            int nb = 0, last_blk = -1;
            for (int i = 0; i < n_it; ++i) {
                scale_step(tile, k, nb, last_blk);
                if (++k == p.k_iters) { k = 0; ++tile; last_blk = -1; }
            }
        }
    } else if (warp == kFixWarp) {
        // =============================== split-K fix-up ==============================
        // For every partial segment of this CTA's range, in order: once the apply warps have issued their reductions,
        // bump the tile's arrival counter (release: their red.adds happen-before it through the mbarrier hand-over);
        // the CTA that arrives last reads the completed fp32 sums back, zeroes the scratch (workspace contract:
        // zero between launches) and writes the tile in T.  A contiguous Stream-K range has at most two partial
        // segments (its first and its last tile), one mbarrier each.
        // (Measured alternative, gpurun r02h2 / r02h3: {sum, arrivals} pairs bumped by one returning packed atomic per
        // contributor -- atom.add.v2.f32, no counter, no fix-up warp.  Correct (151 tests), but a returning atomic on an
        // address 18 CTAs hit takes ~1.3 us, and the volume scales with M: M = 1 7.07 -> 6.88 us, M = 2 7.66 -> 12.3,
        // M = 16 16.2 -> 40.2 on 4096x4096.  Fire-and-forget reductions + one counter per tile stay.)
        if (rg.it1 > rg.it0) {
            int n_fix = 0;
            bool synced = false;
            for (int it = rg.it0; it < rg.it1;) {
                const int tile = it / p.k_iters;
                const int kb = it - tile * p.k_iters;
                const int ke = min(p.k_iters, kb + (rg.it1 - it));
                if (!((kb == 0) && (ke == p.k_iters))) {
                    wait(smem_u32(&ctl->fix_full[n_fix]), 0u, p, DSITE_PFULL, 2);
                    ++n_fix;
                    if (!synced) { pdl_wait_prior_grids(); synced = true; }
                    if (DTRACE_ON(p) && lane == 0) p.trace[blockIdx.x * 48 + 44] = globaltimer_ns();
                    const int tile_it0 = tile * p.k_iters;
                    const int first_cta = cta_of(total, tile_it0, grid);
                    const int contributors = cta_of(total, tile_it0 + p.k_iters - 1, grid) - first_cta + 1;
                    int last = 0;
                    if (lane == 0) {
                        const int old = atom_add_acq_rel(reinterpret_cast<int*>(p.workspace) + tile, 1);
                        last = (old == contributors - 1) ? 1 : 0;
                        if (last) reinterpret_cast<int*>(p.workspace)[tile] = 0;   // self-resetting
                    }
                    last = __shfl_sync(0xffffffffu, last, 0);
                    if (DTRACE_ON(p) && lane == 0) {
                        p.trace[blockIdx.x * 48 + 45] = globaltimer_ns();
                        p.trace[blockIdx.x * 48 + 47] = (unsigned long long)((it - rg.it0) << 8 | last | (contributors << 20));
                    }
                    if (last) {
                        const unsigned seq = out_sequence();
                        __threadfence();      // every lane: order its reads after lane 0's acquire
                        float* accum = reinterpret_cast<float*>(p.workspace + p.partial_offset) + (size_t)tile * (NJ * kMb * 128);
                        const int n_base = tile * TN;
#pragma unroll 1
                        for (int m = 0; m < p.M; ++m) {
                            float v[NJ][4];
#pragma unroll
                            for (int j = 0; j < NJ; ++j)
#pragma unroll
                                for (int qq = 0; qq < 4; ++qq) v[j][qq] = __ldcg(accum + (j * kMb + m) * 128 + qq * 32 + lane);
#pragma unroll
                            for (int j = 0; j < NJ; ++j)
#pragma unroll
                                for (int qq = 0; qq < 4; ++qq) {
                                    accum[(j * kMb + m) * 128 + qq * 32 + lane] = 0.f;
                                    const int n = n_base + n_local<BITS, NJ>(qq * 32 + lane, j, r_tile_p);
                                    if (n < p.N) store_out(m, n, f32_to_t<BF16>(v[j][qq]), seq);
                                }
                        }
                    }
                    if (DTRACE_ON(p) && lane == 0) p.trace[blockIdx.x * 48 + 46] = globaltimer_ns();
                }
                it += ke - kb;
            }
        }
    } else if (warp >= kApplyWarp0) {
        // ===================== scale + accumulate + epilogue ========================
        const int q = warp & 3;
        const int L = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const int fset = (warp - kApplyWarp0) >> 2;        // which NFA-field subset this warp owns
        int nloc[NFA];
#pragma unroll
        for (int j = 0; j < NFA; ++j) nloc[j] = n_local<BITS, NJ>(L, fset * NFA + j, r_tile_p);
        const bool do_apply = !DABLATE(p, 4);
        const uint32_t tmem = (rg.it1 > rg.it0) ? tmem_base_when_ready() : 0u;
        int pslot = 0;
        int stage = 0;                     // ring slot / phase of the stage the loop is at
        uint32_t dphase = 0;
        int sc_idx = 0;
        uint32_t sc_par = 0;
        bool synced = false;
        int n_fix = 0;                     // partial segments handed to the fix-up warp so far (<= 2)
        DPROF_DECL(aw_pfull = 0, aw_sc = 0, aw_work = 0, aw_epi = 0);
        DPROF_T0(at);
        for (int it = rg.it0; it < rg.it1;) {
            const int tile = it / p.k_iters;
            const int kb = it - tile * p.k_iters;
            const int ke = min(p.k_iters, kb + (rg.it1 - it));
            float acc[NFA][MC];
#pragma unroll
            for (int j = 0; j < NFA; ++j)
#pragma unroll
                for (int m = 0; m < MC; ++m) acc[j][m] = 0.f;
            int cur_blk = -1;
            auto release_scales = [&]() {
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl->sc_empty[sc_idx]));
                if (++sc_idx == kScSlots) { sc_idx = 0; sc_par ^= 1u; }
            };
            for (int k = kb; k < ke; ++k) {
                const bool flush = (((k + 1) & spg_mask) == 0) || (k == ke - 1);
                const int fstage = stage;
                const uint32_t fpar = dphase;
                if (++stage == p.stages) { stage = 0; dphase ^= 1u; }
                if (!flush) continue;
                const int g = k >> r_gshift;
                const int blk = g >> 3;
                if (blk != cur_blk) {
                    if (cur_blk >= 0) release_scales();
                    cur_blk = blk;
                    wait(smem_u32(&ctl->sc_full[sc_idx]), sc_par, p, DSITE_SCFULL);
                }
                DPROF_ADD(aw_sc, at);
                const uint32_t sc_base = sc_smem + sc_idx * kScBytes + (g & 7) * 2;
                float sc[NFA];
#pragma unroll
                for (int j = 0; j < NFA; ++j) sc[j] = scale_to_f32<BF16>(lds16(sc_base + nloc[j] * 16));
                wait(smem_u32(&ctl->empty[fstage]), fpar, p, DSITE_PFULL);   // every MMA of the group has completed
                DPROF_ADD(aw_pfull, at);
                tc_fence_after();
                const uint32_t pcol = tmem + lane_sel + kPCol0 + pslot * kPCols;
                if (do_apply) {
                    if constexpr (MC <= 4) {
                        uint32_t r[NFA][MC];
#pragma unroll
                        for (int j = 0; j < NFA; ++j) tmem_ld_cols<MC>(pcol + (fset * NFA + j) * kMb, r[j]);
                        tc_wait_ld();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&ctl->p_empty[pslot]));
#pragma unroll
                        for (int j = 0; j < NFA; ++j)
#pragma unroll
                            for (int m = 0; m < MC; ++m) acc[j][m] = fmaf(sc[j], __uint_as_float(r[j][m]), acc[j][m]);
                    } else {
#pragma unroll
                        for (int j = 0; j < NFA; ++j) {
                            uint32_t r[MC];
                            tmem_ld_cols<MC>(pcol + (fset * NFA + j) * kMb, r);
                            tc_wait_ld();
#pragma unroll
                            for (int m = 0; m < MC; ++m) acc[j][m] = fmaf(sc[j], __uint_as_float(r[m]), acc[j][m]);
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&ctl->p_empty[pslot]));
                    }
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&ctl->p_empty[pslot]));
                }
                if (++pslot == PS) pslot = 0;
                DPROF_ADD(aw_work, at);
            }
            if (cur_blk >= 0) release_scales();

            // ------------------------------- epilogue ------------------------------------
            if (DTRACE_ON(p) && warp == kApplyWarp0 && lane == 0 && it == rg.it0) p.trace[blockIdx.x * 48 + 4] = globaltimer_ns();
            if (!synced) { pdl_wait_prior_grids(); synced = true; }
            const bool full_k = (kb == 0) && (ke == p.k_iters);
            const int n_base = tile * TN;
            if (full_k) {
                const unsigned seq = out_sequence();
#pragma unroll
                for (int j = 0; j < NFA; ++j) {
                    const int n = n_base + nloc[j];
                    if (n < p.N) {
#pragma unroll
                        for (int m = 0; m < MC; ++m)
                            if (m < p.M) store_out(m, n, f32_to_t<BF16>(acc[j][m]), seq);
                    }
                }
            } else {
                // Partial K range: fire-and-forget fp32 reductions into the tile's scratch (zero on entry, left zero
                // on exit), then hand over to the fix-up warp (arrival counter, last-arriver conversion) and carry on
                // with the next segment: nothing on the streaming path waits for a global-memory round trip.
                float* accum = reinterpret_cast<float*>(p.workspace + p.partial_offset) + (size_t)tile * (NJ * kMb * 128);
#pragma unroll
                for (int j = 0; j < NFA; ++j)
#pragma unroll
                    for (int m = 0; m < MC; ++m)
                        if (m < p.M) red_add_f32(accum + ((fset * NFA + j) * kMb + m) * 128 + L, acc[j][m]);
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl->fix_full[n_fix]));
                ++n_fix;
            }
            if (DTRACE_ON(p) && warp == kApplyWarp0 && lane == 0 && it == rg.it0) p.trace[blockIdx.x * 48 + 5] = globaltimer_ns();
            DPROF_ADD(aw_epi, at);
            it += ke - kb;
        }
        if (DTRACE_ON(p) && warp == kApplyWarp0 && lane == 0) p.trace[blockIdx.x * 48 + 6] = globaltimer_ns();
        if (lane == 0 && warp == kApplyWarp0) { DPROF_OUT(40, aw_sc); DPROF_OUT(41, aw_pfull); DPROF_OUT(42, aw_work); DPROF_OUT(43, aw_epi); }
    } else {
        // ================================ dequantisers ==================================
        const int q = warp & 3;                 // TMEM lane quarter
        const int sw = warp >> 2;               // 0..DQ/4-1: piece column set
        const int L = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        {   // lane-replicated LUT (bank == lane, conflict free): every warp fills whole entries, one conflict-free
            // 128-byte store per entry (value broadcast from the lane that loaded it).
            constexpr int kEntriesPerWarp = (F::LUTN + kDqWarps - 1) / kDqWarps;
            static_assert(kEntriesPerWarp <= 32, "one table2 word per lane");
            const int e0 = warp * kEntriesPerWarp;
            const int e_mine = e0 + lane;
            const uint32_t v = (lane < kEntriesPerWarp && e_mine < F::LUTN) ? __ldg(p.table2 + e_mine) : 0u;
#pragma unroll
            for (int i = 0; i < kEntriesPerWarp; ++i) {
                const uint32_t vi = __shfl_sync(0xffffffffu, v, i);
                if (e0 + i < F::LUTN) sts32(lut + (uint32_t)(e0 + i) * 256u + lane * 4, vi);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kDqWarps * 32) : "memory");
        }
        if (DTRACE_ON(p) && threadIdx.x == 0) p.trace[blockIdx.x * 48 + 1] = globaltimer_ns();
        const uint32_t tmem = (rg.it1 > rg.it0) ? tmem_base_when_ready() : 0u;
        const uint32_t lane4 = (uint32_t)lane * 4;

        const uint32_t wrow = (uint32_t)L * 128;
        const int xq = L & 7;
        const bool do_dq = !DABLATE(p, 2);
        // DQG stage groups: with DQG == 2 the 16 warps split into two sets of 8 that convert alternate stages
        // (each warp: 4 quads = 16 k-pairs of its row).  During the conversion of a stage the shared-memory
        // crossbar is the limiter (16 B of packed words + 32 x 4 B of LUT reads per lane = 640 wavefronts per
        // stage); with every warp on the SAME stage the barrier waits of all warps coincide and the crossbar
        // idles ~1/3 of the time.  Two sets, one stage apart, fill each other's gaps.
        constexpr int DQG = F::DQG;
        const int grp = (DQG == 2) ? (sw >> 1) : 0;
        const int hw = (DQG == 2) ? (sw & 1) : sw;          // position inside the set
        DPROF_DECL(dw_full = 0, dw_aempty = 0, dw_piece = 0, dw_st = 0);
        DPROF_T0(dt);
        const int n_it = rg.it1 - rg.it0;
        const int S = p.stages;
        // ring slot / parity of stage i, and of stage t = i - AS/CPS whose completion frees this stage's last A slot
        int stage = grp % S;
        uint32_t fphase = (uint32_t)(grp / S) & 1u;
        int tstage = 0;
        uint32_t tphase = 0;
        {
            const int t0 = grp - AS / CPS + ((grp < AS / CPS) ? DQG * ((AS / CPS - grp + DQG - 1) / DQG) : 0);   // first t >= 0
            tstage = t0 % S;
            tphase = (uint32_t)(t0 / S) & 1u;
        }
        int chunk = grp * CPS;             // global chunk index of (stage i, c = 0)
        int aslot = chunk % AS;
        uint32_t aphase = ((uint32_t)(chunk / AS) & 1u) ^ 1u;
        for (int i = grp; i < n_it; i += DQG) {
            wait(smem_u32(&ctl->full[stage]), fphase, p, DSITE_FULL);
            DPROF_ADD(dw_full, dt);
            const uint32_t row = ring + stage * kStageBytes + wrow;
#pragma unroll
            for (int c = 0; c < CPS; ++c) {
                if (c == CPS - 1) {
                    if (i >= AS / CPS) {
                        wait(smem_u32(&ctl->empty[tstage]), tphase, p, DSITE_AEMPTY);
                        tstage += DQG;
                        if (tstage >= S) { tstage -= S; tphase ^= 1u; }
                    }
                } else {
                    wait(smem_u32(&ctl->a_empty[aslot]), aphase, p, DSITE_AEMPTY, i * 1000 + n_it);
                }
                DPROF_ADD(dw_aempty, dt);
                tc_fence_after();
                const uint32_t tcol = tmem + lane_sel + aslot * kACols;
                if (do_dq) {
                    if constexpr (BITS == 4) {
                        // this warp's 16 k-pairs of the stage: 16-byte quads 4*hw .. 4*hw + 3 of row L
                        Piece<4>::run<CK2>(row, (4 * hw) ^ xq, (4 * hw + 1) ^ xq, lut, lane4, tcol + hw * 16);
                        Piece<4>::run<CK2>(row, (4 * hw + 2) ^ xq, (4 * hw + 3) ^ xq, lut, lane4, tcol + hw * 16 + 8);
                    } else {
                        // half stage c: quad c*4 + sw
                        Piece<2>::run(row, (c * 4 + hw) ^ xq, lut, lane4, tcol + hw * 4);
                    }
                }
                DPROF_ADD(dw_piece, dt);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl->a_full[aslot]));
                if (++aslot == AS) { aslot = 0; aphase ^= 1u; }
                DPROF_ADD(dw_st, dt);
            }
            // skip the chunks of the stages the other set converts
#pragma unroll
            for (int x = 0; x < (DQG - 1) * CPS; ++x)
                if (++aslot == AS) { aslot = 0; aphase ^= 1u; }
            stage += DQG;
            if (stage >= S) { stage -= S; fphase ^= 1u; }
        }
#ifdef FB_PROFILE
        if (lane == 0 && (warp == 0 || warp == 5)) {
            const int o = (warp == 0) ? 24 : 32;
            DPROF_OUT(o + 0, dw_full); DPROF_OUT(o + 1, dw_aempty); DPROF_OUT(o + 2, dw_piece); DPROF_OUT(o + 3, dw_st);
        }
#endif
    }

    // ---- teardown ------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (DTRACE_ON(p) && threadIdx.x == 0) p.trace[blockIdx.x * 48 + 7] = globaltimer_ns();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(*reinterpret_cast<volatile uint32_t*>(&ctl->tmem_base), F::TMEM_COLS);
    }
}

template <int BITS, bool BF16, int MC, bool TP>
static int launch_t(const QgemmArgs& a, cudaStream_t stream) {
    using F = DCfg<BITS>;
    constexpr int TN = F::NJ * 128;
    DecodeParams p{};
    p.A = static_cast<const uint16_t*>(a.A);
    p.Q = static_cast<const uint8_t*>(a.Q);
    p.S = static_cast<const uint16_t*>(a.S);
    p.table2 = static_cast<const uint32_t*>(a.table2);
    p.D = static_cast<uint16_t*>(a.D);
    p.workspace = static_cast<uint8_t*>(a.workspace);
    p.diag = a.diag;
    p.trace = a.trace;
    p.timeout_ns = a.timeout_ns;
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.G = a.K / a.group_size;
    p.P = a.N / 16 * BITS;
    p.tile_p = a.tile_p;
    p.gshift = (a.group_size == 64) ? 0 : (a.group_size == 128) ? 1 : 2;
    p.n_tiles = (a.N + TN - 1) / TN;
    p.k_iters = a.K / 64;
    p.static_weights = (a.flags & FB_FLAG_STATIC_WEIGHTS) ? 1 : 0;
    p.ablate = a.ablate;
    p.tma_scales = ((p.G % 8) == 0 && (reinterpret_cast<uintptr_t>(a.S) & 15) == 0) ? 1 : 0;   // 16-byte scale rows
    p.l2_prefetch = a.l2_prefetch >= 0 ? a.l2_prefetch : 0;
    TpArg<TP> tpa{};
    if constexpr (TP) {
        TpParams& t = tpa.v;
        if (a.tp->tp < 1 || a.tp->tp > 8 || a.tp->rank < 0 || a.tp->rank >= a.tp->tp || a.tp->n_total != a.tp->tp * a.N || a.tp->epoch == nullptr)
            return FB_ERR_SHAPE;
        t.tp = a.tp->tp; t.rank = a.tp->rank; t.n_total = a.tp->n_total;
        // one allocation per rank with the same layout (symmetric memory): the word image sits at the same distance from
        // the plain image on every rank, so the kernel argument carries tp pointers + one distance
        if (a.tp->out_peers[0] == nullptr || a.tp->ll_peers[0] == nullptr) return FB_ERR_NULL;
        t.ll_delta = static_cast<char*>(a.tp->ll_peers[0]) - static_cast<char*>(a.tp->out_peers[0]);
        for (int r = 0; r < t.tp; ++r) {
            if (a.tp->out_peers[r] == nullptr || a.tp->ll_peers[r] == nullptr) return FB_ERR_NULL;
            t.out_peers[r] = static_cast<uint16_t*>(a.tp->out_peers[r]);
            if (static_cast<char*>(a.tp->ll_peers[r]) - static_cast<char*>(a.tp->out_peers[r]) != t.ll_delta) return FB_ERR_SHAPE;
        }
        t.epoch = a.tp->epoch;
        t.write_plain = a.tp->write_plain;
        t.out_uses = a.tp->out_uses; t.out_call = a.tp->out_call;
        t.in_ll = static_cast<const uint2*>(a.tp->in_ll); t.in_ll_stride = a.tp->in_ll_stride;
        t.in_uses = a.tp->in_uses; t.in_call = a.tp->in_call;
        if (t.in_ll != nullptr && ((reinterpret_cast<uintptr_t>(t.in_ll) & 15) != 0 || (t.in_ll_stride & 1) != 0)) return FB_ERR_SHAPE;
    }

    const uint32_t fixed = F::SC_SLOTS * TN * 16 + F::LUTB + sizeof(Ctl) + 1024 /*alignment slack*/;
    int stages = (int)((kSmemBudget - fixed) / kStageBytes);
    if (stages > F::MAX_STAGES) stages = F::MAX_STAGES;
    if (a.force_stages > 0 && a.force_stages < stages) stages = a.force_stages;
    // The dequantisers wait for the MMAs of stage i - A_SLOTS/CPS on that stage's ring barrier; a shallower ring would let
    // the barrier run two phases ahead of that wait (parity aliasing -> deadlock), so the ring is at least that deep.
    constexpr int kMinStages = (F::A_SLOTS / F::CPS > 2) ? F::A_SLOTS / F::CPS : 2;
    if (stages < kMinStages) stages = kMinStages;
    p.stages = stages;
    const uint32_t smem_bytes = stages * kStageBytes + fixed;

    const long long total = (long long)p.n_tiles * p.k_iters;
    if (total > 0x3fffffffLL) return FB_ERR_SHAPE;
    // A few SMs are left out of every launch: the CTAs that finalise a split-K tile finish ~2 us after the others, and
    // with all SMs in use the next launch's last CTAs inherit exactly that delay (they can only start where a CTA has
    // exited), launch after launch.  With kGridSlack spare SMs the late finishers are simply not waited for
    // (measured r02e/r02f: qkv 9.5 -> 9.0 us, o 8.5 -> 8.1, down 13.1 -> 12.3, gate_up unchanged, for 2.7 % fewer SMs).
    constexpr int kGridSlack = 4;
    int grid = a.num_sms > 4 * kGridSlack ? a.num_sms - kGridSlack : a.num_sms;
    if (a.force_grid > 0) {
        grid = a.force_grid;
        if (grid > total) grid = (int)total;
    } else {
        grid = decode_grid_for(total, p.k_iters, grid, true);
    }

    constexpr size_t kCounterBytes = 65536;
    p.partial_offset = (uint32_t)kCounterBytes;
    const size_t need = kCounterBytes + (size_t)p.n_tiles * F::NJ * kMb * 128 * 4;
    if ((size_t)p.n_tiles * 4 > kCounterBytes || need + prefill_scratch_bytes(a.num_sms) > a.workspace_bytes) return FB_ERR_WORKSPACE;

    CUtensorMap tm_w;
    const uint64_t P = (uint64_t)a.N / 16 * BITS;
    int rc = make_tmap_2d(&tm_w, CU_TENSOR_MAP_DATA_TYPE_UINT16, a.Q, (uint64_t)a.K, P, (uint64_t)a.K * 2, 64, 128,
                          CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FB_OK) return rc;
    auto kern = qgemm_decode_kernel<BITS, BF16, MC, TP>;
    static PerDeviceOnce attr_set;      // one per template instantiation
    if (!attr_set.done(a.device)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget) != cudaSuccess) {
            cudaGetLastError();
            return FB_ERR_LAUNCH;
        }
        attr_set.mark(a.device);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads_for(F::DQ, MC));
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    int nattr = 0;
    if (a.flags & FB_FLAG_PDL) {
        attrs[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[nattr].val.programmaticStreamSerializationAllowed = 1;
        ++nattr;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = nattr;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tm_w, p, tpa);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return FB_ERR_LAUNCH;
    }
    return FB_OK;
}

template <int BITS, bool BF16>
static int launch_mc(const QgemmArgs& a, cudaStream_t stream) {
    if (a.tp != nullptr) {        // tensor-parallel fused exchange (tp >= 1): its own instantiations (M <= 4)
        if (a.M == 1) return launch_t<BITS, BF16, 1, true>(a, stream);
        if (a.M <= 4) return launch_t<BITS, BF16, 4, true>(a, stream);
        return FB_ERR_SHAPE;
    }
    if (a.M == 1) return launch_t<BITS, BF16, 1, false>(a, stream);
    if (a.M <= 4) return launch_t<BITS, BF16, 4, false>(a, stream);
    if constexpr (BITS == 4) return launch_t<BITS, BF16, 16, false>(a, stream);
    return FB_ERR_INTERNAL;
}

}  // namespace dec


namespace dec {
__global__ void tp_advance_kernel(unsigned* epoch) { *epoch += 1u; }
// System-scope side of the exchange, for readers that are NOT qgemm_tp launches: `publish` runs after the launches whose
// output it announces (stream order: their stores are complete), makes them visible system-wide and bumps the arrival
// counter of that output on every rank; `wait` spins until this rank's counter has seen every rank's publish of the step.
struct TpFlags {
    unsigned* flag[8];
};
__global__ void tp_publish_kernel(TpFlags f, int tp) {
    __threadfence_system();
    for (int r = 0; r < tp; ++r) red_release_sys_add_u32(f.flag[r], 1u);
}
__global__ void tp_wait_kernel(const unsigned* flag, unsigned per_step, unsigned offset, const unsigned* epoch, uint64_t timeout_ns,
                               Diag* diag) {
    const unsigned expected = (ld_acquire_sys_u32(epoch) - 1u) * per_step + offset;
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while ((int)(ld_acquire_sys_u32(flag) - expected) < 0) {
        if ((++spins & 0xff) == 0 && timeout_ns != 0) {
            const uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > timeout_ns) wait_timeout(diag, DSITE_FULL, 0u, expected, -3);
        }
    }
}
}  // namespace dec

int tp_advance_launch(unsigned* epoch, cudaStream_t stream) {
    dec::tp_advance_kernel<<<1, 1, 0, stream>>>(epoch);
    return cudaGetLastError() == cudaSuccess ? FB_OK : FB_ERR_LAUNCH;
}
int tp_publish_launch(unsigned* const* flags, int tp, cudaStream_t stream) {
    dec::TpFlags f{};
    for (int r = 0; r < tp && r < 8; ++r) f.flag[r] = flags[r];
    dec::tp_publish_kernel<<<1, 1, 0, stream>>>(f, tp);
    return cudaGetLastError() == cudaSuccess ? FB_OK : FB_ERR_LAUNCH;
}
int tp_wait_launch(const unsigned* flag, unsigned per_step, unsigned offset, const unsigned* epoch, uint64_t timeout_ns, Diag* diag,
                   cudaStream_t stream) {
    dec::tp_wait_kernel<<<1, 1, 0, stream>>>(flag, per_step, offset, epoch, timeout_ns, diag);
    return cudaGetLastError() == cudaSuccess ? FB_OK : FB_ERR_LAUNCH;
}

bool qgemm_decode_supported(const QgemmArgs& a) {
    // 4-bit: M <= 16 (16 accumulators per field from M = 5 on).  In round 1 the 16-accumulator instantiation lost to the
    // general kernel (gate_up M = 16: 48 vs 40 us); with this round's changes it wins at every M and shape measured
    // (gpurun r02p2, us, general -> this kernel: qkv M = 5 15.9 -> 12.0, M = 16 20.4 -> 17.8; gate_up M = 16 40.2 -> 30.7;
    // down M = 8 22.6 -> 18.3).  2-bit: eight accumulated fields per lane, so the register-resident accumulators stop at
    // M = 4.
    if (a.M < 1) return false;
    const int m_max = (a.num_bits == 4) ? 16 : 4;
    return (a.num_bits == 4 || a.num_bits == 2) && a.M <= m_max;
}

int qgemm_decode_launch(const QgemmArgs& a, cudaStream_t stream) {
    if (a.num_bits == 4) return a.bf16 ? dec::launch_mc<4, true>(a, stream) : dec::launch_mc<4, false>(a, stream);
    if (a.num_bits == 2) return a.bf16 ? dec::launch_mc<2, true>(a, stream) : dec::launch_mc<2, false>(a, stream);
    return FB_ERR_BITS;
}

}  // namespace fb
