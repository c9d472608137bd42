// LUT-quantized GEMM for sm_100a:  D[M,N] = A[M,K] . W_hat[K,N],
//   W_hat[k,n] = round_T(table2[code(k/2,n)].{lo,hi} * S[n, k/group]).
//
// Replaces the reference's qgemm_device / qgemm_host (flute/csrc/qgemm_kernel.hpp:24-939),
// its register dequantiser (flute/csrc/packbits_utils.hpp:24-427), its Stream-K scheduler
// and fix-up (flute/csrc/tile_scheduler_utils.hpp:58-1058) with one warp-specialised
// kernel, runtime-parameterised instead of template-enumerated.  Consumes the reference's
// packed-weight wire format unchanged (flute/utils.py:59-253).
//
// Mapping ("swap-AB"): the WEIGHTS are the tcgen05 "A" operand and live in TENSOR MEMORY;
// the activations are the "B" operand in shared memory.  One CTA tile is 128 packed rows
// (TMEM lanes) x one 64-wide K stage.  A packed 32-bit word holds NJ pair-codes that belong to
// NJ different output columns (NJ = 4 / 8 / 16 for 4 / 2 / 3 bits), so the tile owns NJ
// accumulators D_j[128 lanes x Mb] and the thread on lane L uses every field of every word
// it loads: field j of word (L, k2) is dequantised into column k2 of "A_j".  The N
// permutation the wire format bakes in is undone for free in the epilogue's store address.
//
// Warp roles (384 threads, persistent over a contiguous Stream-K range):
//   warps 0-7   dequantisers (two groups of four; warp%4 = TMEM lane quarter) + epilogue
//   warp  8     TMA producer (packed-weight tiles + activation tiles, mbarrier ring)
//   warp  9     tcgen05.mma issuer (warp-uniform loop, one elected lane issues), TMEM allocator
//   warps 10-11 scale loaders (global -> smem, group-major so reads are conflict-free)
//
// Two footprints of the same kernel:
//   LARGE  1 CTA/SM, 512 TMEM columns, up to 8 stages, Mb up to 64      (M > 16, and 3-bit)
//   SMALL  2 CTAs/SM, 256 TMEM columns, 3 stages, Mb = 16                (decode, M <= 16)
// SMALL leaves half of every SM free so that, with programmatic dependent launch, the NEXT
// kernel in the stream is already resident and has its first weight tiles in flight while this
// one drains: back-to-back decode GEMMs keep HBM busy across kernel boundaries.
#include "ptx.cuh"
#include "qgemm_sm100.h"

#include <cuda.h>
#include <stdio.h>

namespace fb {

// Optional per-role cycle accounting (profiling build only: python flute_b200/build.py --profile).
#ifdef FB_PROFILE
#define PROF_DECL(...) long long __VA_ARGS__
#define PROF_T0(t) long long t = clock64()
#define PROF_ADD(acc, t) do { long long _n = clock64(); acc += _n - t; t = _n; } while (0)
#define PROF_OUT(slot, v) do { if (p.trace != nullptr) p.trace[blockIdx.x * 48 + (slot)] = (unsigned long long)(v); } while (0)
#else
#define PROF_DECL(...)
#define PROF_T0(t)
#define PROF_ADD(acc, t)
#define PROF_OUT(slot, v)
#endif
constexpr int kTraceStride = 48;

// Pin a kernel parameter in a register for the lifetime of a role's loop.  Without this ptxas re-reads it
// from the constant bank (LDCU, ~50-100 cycles, serially dependent on the uniform datapath) every iteration;
// profiling showed ~800 cycles per pipeline stage going to that alone.
template <typename T>
__device__ __forceinline__ T pin(T v) {
    static_assert(sizeof(T) == 4, "pin: 32-bit values");
    asm volatile("" : "+r"(v));
    return v;
}

// ----------------------------------------------------------------------------------------
// Per-format / per-footprint constants
// ----------------------------------------------------------------------------------------
template <int BITS, bool SMALL>
struct Cfg;
// NJ    pair fields per 32-bit word == accumulators per tile
// NG    dequantiser groups (4 warps each, one per TMEM lane quarter); chunk c belongs to group c % NG
// CPS   TMEM chunks per 64-k stage;  K2C  k-pairs per chunk (TMEM columns per field)
// ROWS  smem rows per stage;  LUTN  table2 entries;  SCH  scale groups per smem scale chunk
// LUT_STRIDE  bytes between LUT entries: 256 lets ONE prmt build (code << 8) | lane*4 for 4-bit codes
template <>
struct Cfg<4, false> {
    static constexpr int NJ = 4, NG = 4, CPS = 2, K2C = 16, ROWS = 128, LUTN = 256, SCH = 8;
    static constexpr int TMEM_COLS = 512, MIN_BLOCKS = 1, LUT_STRIDE = 256;
};
template <>
struct Cfg<4, true> {
    static constexpr int NJ = 4, NG = 2, CPS = 2, K2C = 16, ROWS = 128, LUTN = 256, SCH = 8;
    static constexpr int TMEM_COLS = 256, MIN_BLOCKS = 2, LUT_STRIDE = 128;
};
template <>
struct Cfg<2, false> {
    static constexpr int NJ = 8, NG = 4, CPS = 4, K2C = 8, ROWS = 128, LUTN = 16, SCH = 8;
    static constexpr int TMEM_COLS = 512, MIN_BLOCKS = 1, LUT_STRIDE = 128;
};
template <>
struct Cfg<2, true> {
    static constexpr int NJ = 8, NG = 2, CPS = 4, K2C = 8, ROWS = 128, LUTN = 16, SCH = 8;
    static constexpr int TMEM_COLS = 256, MIN_BLOCKS = 2, LUT_STRIDE = 128;
};
template <>
struct Cfg<3, false> {
    static constexpr int NJ = 16, NG = 2, CPS = 4, K2C = 8, ROWS = 384, LUTN = 64, SCH = 4;
    static constexpr int TMEM_COLS = 512, MIN_BLOCKS = 1, LUT_STRIDE = 128;
};
template <int BITS, bool SMALL>
struct Roles {
    using F = Cfg<BITS, SMALL>;
    static constexpr int kDequantWarps = F::NG * 4;
    static constexpr int kProducerWarp = kDequantWarps;
    static constexpr int kMmaWarp = kDequantWarps + 1;
    static constexpr int kScaleWarp0 = kDequantWarps + 2;
    static constexpr int kThreads = (kDequantWarps + 4) * 32;
    // dequant warps that read one smem stage (each arrives on its `empty` barrier once)
    static constexpr int kStageReaders = (F::CPS < F::NG ? F::CPS : F::NG) * 4;
};

constexpr int kScaleWarps = 2;
constexpr int kMaxStages = 8;
constexpr int kMaxChunkSlots = 8;
constexpr int kStageK = 64;

struct SmemCtl {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t a_full[kMaxChunkSlots];
    uint64_t a_empty[kMaxChunkSlots];
    uint64_t sc_full[2];
    uint64_t sc_empty[2];
    uint64_t acc_full;
    uint64_t acc_empty;
    uint32_t tmem_base;
    int is_last;
};

// ----------------------------------------------------------------------------------------
// Work partition (identical arithmetic in every role)
// ----------------------------------------------------------------------------------------
struct Range {
    int it0, it1;
};

__device__ __forceinline__ Range cta_range(const QgemmParams& p, int b, int grid) {
    Range r;
    if (p.streamk) {
        int total = p.n_tiles * p.m_tiles * p.k_iters;
        int base = total / grid, rem = total - base * grid;
        r.it0 = b * base + min(b, rem);
        r.it1 = r.it0 + base + (b < rem ? 1 : 0);
    } else {
        int tiles = p.n_tiles * p.m_tiles;
        int base = tiles / grid, rem = tiles - base * grid;
        int t0 = b * base + min(b, rem);
        int t1 = t0 + base + (b < rem ? 1 : 0);
        r.it0 = t0 * p.k_iters;
        r.it1 = t1 * p.k_iters;
    }
    return r;
}

__device__ __forceinline__ int streamk_cta_of(const QgemmParams& p, int it, int grid) {
    int total = p.n_tiles * p.m_tiles * p.k_iters;
    int base = total / grid, rem = total - base * grid;
    int thr = rem * (base + 1);
    return it < thr ? it / (base + 1) : rem + (it - thr) / base;
}

// column of N that TMEM lane L / field j of this tile maps to (relative to the tile's first column)
template <int BITS, int NJ>
__device__ __forceinline__ int n_local(int L, int j, int tile_p) {
    if (BITS == 3 || tile_p == 32) return (L >> 5) * (NJ * 32) + j * 32 + (L & 31);
    return (L >> 6) * (NJ * 64) + j * 64 + (L & 63);
}

// ----------------------------------------------------------------------------------------
// Dequantise one TMEM chunk for lane L:  packed words (smem, 128B-swizzled rows) -> table2 pair
// lookup (32x lane-replicated LUT: bank == lane, conflict free) -> one mul.{f16,bf16}x2 by the
// group scale -> tcgen05.st.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// byte J of w, zero extended (one PRMT)
template <int J>
__device__ __forceinline__ uint32_t byte_of(uint32_t w) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0u), "n"(0x4440 + J));
    return r;
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ int atom_add_acq_rel(int* addr, int v) {
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

template <int N>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t (&r)[N]) {
    static_assert(N == 16 || N == 32, "chunk store width");
    if constexpr (N == 32) tmem_st_32x32b_x32(taddr, r);
    else tmem_st_32x32b_x16(taddr, r);
}

template <int BITS, int K2C, bool BF16>
struct Dequant;

// 4-bit: byte j of word k2 -> column j*K2C + k2.  LUT address of code c for lane l:
//   STRIDE 256:  lut + ((c << 8) | l*4)         one PRMT merges byte J of w with the lane byte
//   STRIDE 128:  lut + ((c << 7) | l*4)         PRMT + SHF (ALU pipe) and PRMT + IMAD (FMA pipe) alternate,
//                                               so neither pipe carries the whole address arithmetic
template <int J>
__device__ __forceinline__ uint32_t prmt_code_lane(uint32_t w, uint32_t lane_byte) {
    uint32_t r;   // bytes: [0] = lane_byte.b0, [1] = w.b<J>, [2] = lane_byte.b1 (0), [3] = lane_byte.b2 (0)
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(lane_byte), "n"(0x6504 + (J << 4)));
    return r;
}
template <int K2C, bool BF16, int STRIDE>
struct Dequant4 {
    template <int J>
    static __device__ __forceinline__ void field(const uint32_t (&w)[K2C], uint32_t lut, uint32_t lane, uint32_t sc,
                                                 uint32_t nz, uint32_t taddr) {
        uint32_t r[K2C];
        if constexpr (STRIDE == 256) {
            const uint32_t lane4 = lane * 4;
#ifndef FB_ABL
#define FB_ABL 0
#endif
            // compile-time ablations for perf studies (tools only): 4 no TMEM store, 8 no LUT load, 16 no multiply
#pragma unroll
            for (int i = 0; i < K2C; ++i) {
                uint32_t a = prmt_code_lane<J>(w[i], lane4);
                uint32_t v = (FB_ABL & 8) ? a : lds32(a + lut);
                r[i] = (FB_ABL & 16) ? (v ^ sc) : mul2<BF16>(v, sc, nz);
            }
            if (FB_ABL & 4) {
                uint32_t x = 0;
#pragma unroll
                for (int i = 0; i < K2C; ++i) x ^= r[i];
                if (x == 0x12345678u) tmem_st_cols(taddr + J * K2C, r);   // keeps r[] live, (almost) never stores
                return;
            }
        } else {
            const uint32_t lane8 = lane * 8, lut_lane = lut + lane * 4;
#pragma unroll
            for (int i = 0; i < K2C; ++i) {
                uint32_t a;
                if (i & 1) a = (prmt_code_lane<J>(w[i], lane8) >> 1) + lut;
                else a = byte_of<J>(w[i]) * 128u + lut_lane;
                r[i] = mul2<BF16>(lds32(a), sc, nz);
            }
        }
        tmem_st_cols(taddr + J * K2C, r);
    }
    static __device__ __forceinline__ void run(uint32_t wstage, int sub, int L, uint32_t lut, uint32_t lane,
                                               const uint32_t* sc, uint32_t nz, uint32_t tchunk) {
        const uint32_t row = wstage + L * 128;
        const int x = L & 7;
        uint32_t w[K2C];
#pragma unroll
        for (int c = 0; c < K2C / 4; ++c) {
            uint4 v = lds128(row + (((sub * (K2C / 4) + c) ^ x) << 4));
            w[4 * c + 0] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
        }
        field<0>(w, lut, lane, sc[0], nz, tchunk);
        field<1>(w, lut, lane, sc[1], nz, tchunk);
        field<2>(w, lut, lane, sc[2], nz, tchunk);
        field<3>(w, lut, lane, sc[3], nz, tchunk);
    }
};
template <int K2C, bool BF16>
struct Dequant<4, K2C, BF16> {
    template <int STRIDE>
    static __device__ __forceinline__ void run(uint32_t wstage, int sub, int L, uint32_t lut, uint32_t lane,
                                               const uint32_t* sc, uint32_t nz, uint32_t tchunk) {
        Dequant4<K2C, BF16, STRIDE>::run(wstage, sub, L, lut, lane, sc, nz, tchunk);
    }
};

// 2-bit: nibble j of word k2 -> column j*K2C + k2; stores cover 32 columns (32/K2C fields)
template <int K2C, bool BF16>
struct Dequant<2, K2C, BF16> {
    template <int STRIDE>
    static __device__ __forceinline__ void run(uint32_t wstage, int sub, int L, uint32_t lut, uint32_t lane,
                                               const uint32_t* sc, uint32_t nz, uint32_t tchunk) {
        static_assert(STRIDE == 128, "2-bit LUT stride");
        const uint32_t lut_lane = lut + lane * 4;
        const uint32_t row = wstage + L * 128;
        const int x = L & 7;
        uint32_t w[K2C];
#pragma unroll
        for (int c = 0; c < K2C / 4; ++c) {
            uint4 v = lds128(row + (((sub * (K2C / 4) + c) ^ x) << 4));
            w[4 * c + 0] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
        }
        constexpr int FPS = 32 / K2C;   // fields per 32-column store
#pragma unroll
        for (int jj = 0; jj < 8 / FPS; ++jj) {
            uint32_t r[32];
#pragma unroll
            for (int h = 0; h < FPS; ++h) {
                const int j = FPS * jj + h;
#pragma unroll
                for (int i = 0; i < K2C; ++i) {
                    const uint32_t code = (w[i] >> (4 * j)) & 0xfu;
                    r[h * K2C + i] = mul2<BF16>(lds32(code * 128u + lut_lane), sc[j], nz);
                }
            }
            tmem_st_32x32b_x32(tchunk + jj * 32, r);
        }
    }
};

// 3-bit (K2C = 8): three words per k2 (plane 0 rows [0,128); planes 1/2 rows [128,384): 64 per
// 32-row block, second plane +32); 6-bit field j -> column j*8 + k2
template <bool BF16>
struct Dequant<3, 8, BF16> {
    template <int STRIDE>
    static __device__ __forceinline__ void run(uint32_t wstage, int sub, int L, uint32_t lut, uint32_t lane,
                                               const uint32_t* sc, uint32_t nz, uint32_t tchunk) {
        static_assert(STRIDE == 128, "3-bit LUT stride");
        const uint32_t lut_lane = lut + lane * 4;
        const uint32_t row0 = wstage + L * 128;
        const uint32_t row1 = wstage + (128 + (L >> 5) * 64 + (L & 31)) * 128;
        const uint32_t row2 = row1 + 32 * 128;
        const int x = L & 7;
        uint32_t w0[8], w1[8], w2[8];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const uint32_t off = ((2 * sub + c) ^ x) << 4;
            uint4 a = lds128(row0 + off), b = lds128(row1 + off), d = lds128(row2 + off);
            w0[4 * c + 0] = a.x; w0[4 * c + 1] = a.y; w0[4 * c + 2] = a.z; w0[4 * c + 3] = a.w;
            w1[4 * c + 0] = b.x; w1[4 * c + 1] = b.y; w1[4 * c + 2] = b.z; w1[4 * c + 3] = b.w;
            w2[4 * c + 0] = d.x; w2[4 * c + 1] = d.y; w2[4 * c + 2] = d.z; w2[4 * c + 3] = d.w;
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            uint32_t r[32];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int j = 4 * jj + h;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint32_t code;
                    if (j < 15) {
                        const uint32_t w = (j % 3 == 0) ? w0[i] : (j % 3 == 1) ? w1[i] : w2[i];
                        code = (w >> (6 * (j / 3))) & 0x3fu;
                    } else {
                        code = (w0[i] >> 30) | ((w1[i] >> 30) << 2) | ((w2[i] >> 30) << 4);
                    }
                    r[h * 8 + i] = mul2<BF16>(lds32(code * 128u + lut_lane), sc[j], nz);
                }
            }
            tmem_st_32x32b_x32(tchunk + jj * 32, r);
        }
    }
};

// ----------------------------------------------------------------------------------------
// The kernel
// ----------------------------------------------------------------------------------------
template <int BITS, bool BF16, bool SMALL>
__global__ void __launch_bounds__(Roles<BITS, SMALL>::kThreads, Cfg<BITS, SMALL>::MIN_BLOCKS)
qgemm_sm100_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
                   const QgemmParams p) {
    using F = Cfg<BITS, SMALL>;
    using R = Roles<BITS, SMALL>;
    constexpr int NJ = F::NJ;
    constexpr int NG = F::NG;
    constexpr int CPS = F::CPS;
    constexpr int K2C = F::K2C;
    constexpr int CC = NJ * K2C;          // TMEM columns per chunk
    constexpr int kDequantWarps = R::kDequantWarps;
    constexpr int kProducerWarp = R::kProducerWarp;
    constexpr int kMmaWarp = R::kMmaWarp;
    constexpr int kScaleWarp0 = R::kScaleWarp0;
    constexpr int kThreads = R::kThreads;
    constexpr int TN = NJ * 128;
    constexpr int SCH = F::SCH;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const uint32_t ring = smem_base;
    const uint32_t lut = ring + p.stages * p.stage_bytes;
    const uint32_t sc_smem = lut + F::LUTN * F::LUT_STRIDE;
    uint16_t* sc_gen = reinterpret_cast<uint16_t*>(smem_gen + (sc_smem - smem_base));
    constexpr uint32_t kScSlotElems = SCH * TN;
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem_gen + (sc_smem + 2 * kScSlotElems * 2 - smem_base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int grid = gridDim.x;
    const Range rg = cta_range(p, blockIdx.x, grid);

    // ---- one-time setup -------------------------------------------------------------
    if (p.trace != nullptr && threadIdx.x == 0) p.trace[blockIdx.x * kTraceStride + 0] = globaltimer_ns();
    // Under programmatic dependent launch this grid may start while earlier kernels of the stream still run.
    // Unless the caller promises that Q / S / table2 are static (FLUTE_B200_FLAG_STATIC_WEIGHTS), wait for them
    // before reading anything; with the promise only activations, outputs and workspace are ordered (below).
    if (!p.static_weights) pdl_wait_prior_grids();
    if (warp == kProducerWarp && lane == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_a);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&ctl->full[s]), 1);
            mbar_init(smem_u32(&ctl->empty[s]), R::kStageReaders + 1);
        }
        for (int c = 0; c < kMaxChunkSlots; ++c) {
            mbar_init(smem_u32(&ctl->a_full[c]), 4);
            mbar_init(smem_u32(&ctl->a_empty[c]), 1);
        }
        for (int c = 0; c < 2; ++c) {
            mbar_init(smem_u32(&ctl->sc_full[c]), kScaleWarps);
            mbar_init(smem_u32(&ctl->sc_empty[c]), kDequantWarps);
        }
        mbar_init(smem_u32(&ctl->acc_full), 1);
        mbar_init(smem_u32(&ctl->acc_empty), kDequantWarps);
        mbar_fence_init();
    }
    if (warp == kMmaWarp) {
        tmem_alloc(smem_u32(&ctl->tmem_base), F::TMEM_COLS);
        tmem_relinquish();
    }
    {   // lane-replicated LUT: entry e of lane l at lut + e*LUT_STRIDE + l*4  (weights-only data: no PDL wait).
        // One global load per entry, staged through the (still unused) scale buffer, then replicated.
        uint32_t* stage = reinterpret_cast<uint32_t*>(sc_gen);
        for (int i = threadIdx.x; i < F::LUTN; i += kThreads) stage[i] = __ldg(p.table2 + i);
        __syncthreads();
        uint32_t* lut_gen = reinterpret_cast<uint32_t*>(smem_gen + (lut - smem_base));
        for (int i = threadIdx.x; i < F::LUTN * 32; i += kThreads)
            lut_gen[(i >> 5) * (F::LUT_STRIDE / 4) + (i & 31)] = stage[i >> 5];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = ctl->tmem_base;
    pdl_launch_dependents();
    if (p.trace != nullptr && threadIdx.x == 0) p.trace[blockIdx.x * kTraceStride + 1] = globaltimer_ns();

    const uint32_t acc_col = p.nchunk * CC;   // accumulators sit after the A chunk slots

    if (warp == kProducerWarp) {
        // =============================== TMA producer ===============================
        if (lane == 0 && rg.it1 > rg.it0) {
            const uint64_t pol_w = policy_evict_first();
            const uint64_t pol_a = policy_evict_last();
            const int n_it = rg.it1 - rg.it0;
            const int r_stages = pin(p.stages), r_k_iters = pin(p.k_iters), r_m_tiles = pin(p.m_tiles), r_mb = pin(p.mb);
            const uint32_t r_stage_bytes = pin(p.stage_bytes), r_w_bytes = pin(p.w_bytes);
            const uint32_t r_tx_bytes = pin(p.w_bytes + p.b_bytes), r_plane1 = pin(p.plane1_row0);
            const int npro = min(r_stages, n_it);
            // (nt, mt, k) of an iteration, advanced incrementally: no divisions on the per-stage path
            struct Cur { int nt, mt, k; };
            auto advance = [&](Cur& c) {
                if (++c.k == r_k_iters) {
                    c.k = 0;
                    if (++c.mt == r_m_tiles) { c.mt = 0; ++c.nt; }
                }
            };
            Cur c0;
            {
                const int tile = rg.it0 / p.k_iters;
                c0.k = rg.it0 - tile * p.k_iters;
                c0.nt = tile / p.m_tiles;
                c0.mt = tile - c0.nt * p.m_tiles;
            }
            auto load_w = [&](const Cur& c, int s) {
                const uint32_t dst = ring + s * r_stage_bytes;
                const uint32_t bar = smem_u32(&ctl->full[s]);
                if (BITS == 3) {
                    tma_load_2d(dst, &tmap_w, bar, c.k * kStageK, c.nt * 128, pol_w);
                    tma_load_2d(dst + 128 * 128, &tmap_w, bar, c.k * kStageK, r_plane1 + c.nt * 256, pol_w);
                    tma_load_2d(dst + 256 * 128, &tmap_w, bar, c.k * kStageK, r_plane1 + c.nt * 256 + 128, pol_w);
                } else {
                    tma_load_2d(dst, &tmap_w, bar, c.k * kStageK, c.nt * 128, pol_w);
                }
            };
            auto load_a = [&](const Cur& c, int s) {
                tma_load_2d(ring + s * r_stage_bytes + r_w_bytes, &tmap_a, smem_u32(&ctl->full[s]), c.k * kStageK,
                            c.mt * r_mb, pol_a);
            };
            // Weights never depend on the previous kernel in the stream: start streaming them
            // before the programmatic-dependency wait, activations after it.
            Cur cw = c0;
            for (int i = 0; i < npro; ++i) {
                mbar_arrive_expect_tx(smem_u32(&ctl->full[i]), r_tx_bytes);
                load_w(cw, i);
                advance(cw);
            }
            pdl_wait_prior_grids();
            Cur ca = c0;
            for (int i = 0; i < npro; ++i) { load_a(ca, i); advance(ca); }
            int stage = (npro == r_stages) ? 0 : npro;
            uint32_t phase = (npro == r_stages) ? 1u : 0u;
            PROF_DECL(pw_empty = 0, pw_issue = 0);
            PROF_T0(pt);
            for (int i = npro; i < n_it; ++i) {
                mbar_wait(smem_u32(&ctl->empty[stage]), phase ^ 1u, p.diag, p.timeout_ns, SITE_PROD_EMPTY, stage, i);
                PROF_ADD(pw_empty, pt);
                mbar_arrive_expect_tx(smem_u32(&ctl->full[stage]), r_tx_bytes);
                load_w(cw, stage);
                load_a(cw, stage);
                advance(cw);
                if (++stage == r_stages) { stage = 0; phase ^= 1u; }
                PROF_ADD(pw_issue, pt);
            }
            PROF_OUT(8, pw_empty); PROF_OUT(9, pw_issue); PROF_OUT(10, n_it - npro);
        }
    } else if (warp == kMmaWarp) {
        // =============================== MMA issuer =================================
        // The whole warp runs the loop (so the address arithmetic stays on the uniform datapath);
        // one elected lane issues the tcgen05 instructions.
        if (rg.it1 > rg.it0) {
            const uint32_t idesc = pin(make_idesc_f16(BF16, 128, p.mb));
            const uint32_t d_base = tmem + acc_col;
            const int r_stages = pin(p.stages), r_nchunk = pin(p.nchunk), r_mb = pin(p.mb);
            const uint32_t r_stage_bytes = pin(p.stage_bytes), r_w_bytes = pin(p.w_bytes);
            int stage = 0;
            uint32_t phase = 0;
            int slot = 0;
            uint32_t aphase = 0;
            int seg = 0;
            bool stamped = false;
            PROF_DECL(mw_full = 0, mw_afull = 0, mw_issue = 0, mw_acc = 0);
            PROF_T0(mt_);
            for (int it = rg.it0; it < rg.it1;) {
                const int tile = it / p.k_iters;
                const int kb = it - tile * p.k_iters;
                const int ke = min(p.k_iters, kb + (rg.it1 - it));
                mbar_wait(smem_u32(&ctl->acc_empty), (seg & 1) ^ 1u, p.diag, p.timeout_ns, SITE_MMA_ACCEMPTY, 0, seg);
                tc_fence_after();
                PROF_ADD(mw_acc, mt_);
                uint32_t acc_flag = 0;   // first MMA of every accumulator overwrites
                for (int k = kb; k < ke; ++k) {
                    mbar_wait(smem_u32(&ctl->full[stage]), phase, p.diag, p.timeout_ns, SITE_MMA_FULL, stage, it);
                    PROF_ADD(mw_full, mt_);
                    if (p.trace != nullptr && !stamped && lane == 0) { p.trace[blockIdx.x * kTraceStride + 2] = globaltimer_ns(); stamped = true; }
                    const uint64_t bdesc = make_smem_desc_sw128(ring + stage * r_stage_bytes + r_w_bytes);
#pragma unroll
                    for (int sub = 0; sub < CPS; ++sub) {
                        mbar_wait(smem_u32(&ctl->a_full[slot]), aphase, p.diag, p.timeout_ns, SITE_MMA_AFULL, slot, k);
                        tc_fence_after();
                        PROF_ADD(mw_afull, mt_);
                        const uint32_t a_base = tmem + slot * CC;
                        if (elect_one()) {
                            if (!(p.ablate & 1))
#pragma unroll
                            for (int j = 0; j < NJ; ++j) {
#pragma unroll
                                for (int kk = 0; kk < K2C / 8; ++kk) {
                                    tc_mma_ts(d_base + j * r_mb, a_base + j * K2C + kk * 8,
                                              bdesc + (uint64_t)((sub * K2C * 4 + kk * 32) >> 4), idesc,
                                              kk == 0 ? acc_flag : 1u);
                                }
                            }
                            tc_commit(smem_u32(&ctl->a_empty[slot]));
                        }
                        __syncwarp();
                        acc_flag = 1;
                        if (++slot == r_nchunk) { slot = 0; aphase ^= 1u; }
                    }
                    if (elect_one()) tc_commit(smem_u32(&ctl->empty[stage]));
                    __syncwarp();
                    if (++stage == r_stages) { stage = 0; phase ^= 1u; }
                    PROF_ADD(mw_issue, mt_);
                }
                if (elect_one()) tc_commit(smem_u32(&ctl->acc_full));
                __syncwarp();
                if (p.trace != nullptr && lane == 0) p.trace[blockIdx.x * kTraceStride + 3] = globaltimer_ns();
                it += ke - kb;
                ++seg;
            }
            if (lane == 0) { PROF_OUT(11, mw_full); PROF_OUT(12, mw_afull); PROF_OUT(13, mw_issue); PROF_OUT(14, mw_acc); }
        }
    } else if (warp >= kScaleWarp0) {
        // =============================== scale loaders ==============================
        const int tid = (warp - kScaleWarp0) * 32 + lane;
        int n_sc = 0;
        PROF_DECL(lw_empty = 0, lw_load = 0);
        PROF_T0(lt_);
        const bool vec_ok = ((p.G % SCH) == 0) && ((reinterpret_cast<uintptr_t>(p.S) & 15) == 0);
        const int r_gshift = pin(p.group_shift);
        for (int it = rg.it0; it < rg.it1;) {
            const int tile = it / p.k_iters;
            const int kb = it - tile * p.k_iters;
            const int ke = min(p.k_iters, kb + (rg.it1 - it));
            const int nt = tile / p.m_tiles;
            const int c_first = (kb >> r_gshift) / SCH;
            const int c_last = ((ke - 1) >> r_gshift) / SCH;
            for (int c = c_first; c <= c_last; ++c, ++n_sc) {
                const int slot = n_sc & 1;
                const uint32_t par = (n_sc >> 1) & 1;
                mbar_wait(smem_u32(&ctl->sc_empty[slot]), par ^ 1u, p.diag, p.timeout_ns, SITE_SC_EMPTY, slot, n_sc);
                PROF_ADD(lw_empty, lt_);
                uint16_t* dst = sc_gen + slot * kScSlotElems;
                // every row this thread owns is fetched before anything is stored: one memory latency per
                // chunk instead of one per row
                constexpr int RPT = TN / (kScaleWarps * 32);   // rows per thread
                if (vec_ok) {
                    uint4 q[RPT];
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        const int n = nt * TN + tid + r * (kScaleWarps * 32);
                        q[r] = make_uint4(0, 0, 0, 0);
                        if (n < p.N) {
                            const uint16_t* src = p.S + (size_t)n * p.G + c * SCH;
                            if constexpr (SCH == 8) {
                                q[r] = __ldg(reinterpret_cast<const uint4*>(src));
                            } else {
                                const uint2 h = __ldg(reinterpret_cast<const uint2*>(src));
                                q[r].x = h.x; q[r].y = h.y;
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        const int nl = tid + r * (kScaleWarps * 32);
                        const uint32_t w4[4] = {q[r].x, q[r].y, q[r].z, q[r].w};
#pragma unroll
                        for (int g = 0; g < SCH; ++g) dst[g * TN + nl] = (uint16_t)(w4[g >> 1] >> ((g & 1) * 16));
                    }
                } else {
                    for (int r0 = 0; r0 < RPT; r0 += 4) {
                        uint16_t v[4][SCH];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int n = nt * TN + tid + (r0 + r) * (kScaleWarps * 32);
#pragma unroll
                            for (int g = 0; g < SCH; ++g)
                                v[r][g] = (n < p.N && c * SCH + g < p.G) ? __ldg(p.S + (size_t)n * p.G + c * SCH + g) : (uint16_t)0;
                        }
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int nl = tid + (r0 + r) * (kScaleWarps * 32);
#pragma unroll
                            for (int g = 0; g < SCH; ++g) dst[g * TN + nl] = v[r][g];
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl->sc_full[slot]));
                PROF_ADD(lw_load, lt_);
            }
            it += ke - kb;
        }
        if (tid == 0) { PROF_OUT(40, lw_empty); PROF_OUT(41, lw_load); PROF_OUT(42, n_sc); }
    } else {
        // ========================= dequantisers + epilogue ==========================
        const int group = warp >> 2;            // 0 .. NG-1
        const int q = warp & 3;                 // TMEM lane quarter
        const int L = q * 32 + lane;            // TMEM lane == packed row within the tile
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        // Per-warp pipeline state.  A group steps directly from one of ITS chunks to the next (chunk ids
        // group, group + NG, ...): no per-chunk iteration over the other groups' work.
        int my_il = 0;                          // CTA-local stage index of my next chunk
        int my_sub = group;                     // sub-chunk (0..CPS-1) of my next chunk inside that stage
        while (my_sub >= CPS) { my_sub -= CPS; ++my_il; }
        int stage = my_il;                      // ring slot of stage my_il (my_il < NG <= stages here)
        uint32_t sphase = 0;
        int slot = group;                       // TMEM chunk slot of my next chunk (group < NG <= ... handled below)
        uint32_t apar = 0;
        int seg_il0 = 0;                        // CTA-local index of the segment's first stage
        int sc_seg0 = 0;                        // global index of the segment's first scale chunk
        int sc_cur = -1;                        // global index of the scale chunk currently held (-1: none)
        int sc_done = 0;                        // scale chunks released so far (every warp releases every chunk once)
        int seg = 0;
        bool synced = false;
        bool dbg_done = false;
        const int r_stages = pin(p.stages), r_nchunk = pin(p.nchunk), r_gshift = pin(p.group_shift);
        const uint32_t r_stage_bytes = pin(p.stage_bytes);
        while (stage >= r_stages) stage -= r_stages;   // (only if stages < NG)
        while (slot >= r_nchunk) { slot -= r_nchunk; apar ^= 1u; }
        const uint32_t r_nz = pin(p.neg_zero2);
        PROF_DECL(dw_scw = 0, dw_sc = 0, dw_full = 0, dw_aempty = 0, dw_run = 0, dw_st = 0, dw_epi = 0, dw_chunks = 0);
        PROF_T0(dt_);
        int nloc[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) nloc[j] = n_local<BITS, NJ>(L, j, p.tile_p);

        // Release scale chunks [sc_done, upto): the one this warp holds and any it never needed.  A chunk that
        // was never acquired is first waited for, so no warp can run more than one buffer generation ahead of
        // the others (each sc_empty phase must collect exactly one arrival per warp).
        auto release_scales_upto = [&](int upto) {
            while (sc_done < upto) {
                if (sc_done != sc_cur)
                    mbar_wait(smem_u32(&ctl->sc_full[sc_done & 1]), (sc_done >> 1) & 1, p.diag, p.timeout_ns,
                              SITE_DQ_SCALE, sc_done & 1, sc_done);
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl->sc_empty[sc_done & 1]));
                ++sc_done;
            }
        };

        for (int it = rg.it0; it < rg.it1;) {
            const int tile = it / p.k_iters;
            const int kb = it - tile * p.k_iters;
            const int ke = min(p.k_iters, kb + (rg.it1 - it));
            const int nt = tile / p.m_tiles;
            const int mt = tile - nt * p.m_tiles;
            const int seg_il1 = seg_il0 + (ke - kb);
            const int scid0 = (kb >> r_gshift) / SCH;                       // first scale chunk id of the segment
            const int sc_count = ((ke - 1) >> r_gshift) / SCH - scid0 + 1;  // scale chunks the loader makes for it
            while (my_il < seg_il1) {
                const int k = kb + (my_il - seg_il0);
                // ---- scale chunk for this stage ----
                const int g = k >> r_gshift;
                const int sc_need = sc_seg0 + (g / SCH - scid0);
                if (sc_need != sc_cur) {
                    release_scales_upto(sc_need);       // the one I held and any I skipped over
                    sc_cur = sc_need;
                    PROF_ADD(dw_sc, dt_);
                    mbar_wait(smem_u32(&ctl->sc_full[sc_need & 1]), (sc_need >> 1) & 1, p.diag, p.timeout_ns,
                              SITE_DQ_SCALE, sc_need & 1, sc_need);
                    PROF_ADD(dw_scw, dt_);
                }
                PROF_ADD(dw_sc, dt_);
                mbar_wait(smem_u32(&ctl->full[stage]), sphase, p.diag, p.timeout_ns, SITE_DQ_FULL, stage, k);
                PROF_ADD(dw_full, dt_);
                // group scales for this lane's NJ columns, replicated into both halves
                uint32_t sc[NJ];
                {
                    const uint16_t* src = sc_gen + (sc_need & 1) * kScSlotElems + (g % SCH) * TN;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        uint32_t s = src[nloc[j]];
                        sc[j] = s | (s << 16);
                    }
                }
                mbar_wait(smem_u32(&ctl->a_empty[slot]), apar ^ 1u, p.diag, p.timeout_ns, SITE_DQ_AEMPTY, slot, k);
                tc_fence_after();
                PROF_ADD(dw_aempty, dt_);
                if (!(p.ablate & 2))
                    Dequant<BITS, K2C, BF16>::template run<F::LUT_STRIDE>(ring + stage * r_stage_bytes, my_sub, L, lut,
                                                                          (uint32_t)lane, sc, r_nz, tmem + lane_sel + slot * CC);
                if (my_sub + NG >= CPS) {   // this warp's last read of the stage
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&ctl->empty[stage]));
                }
                PROF_ADD(dw_run, dt_);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl->a_full[slot]));
                PROF_ADD(dw_st, dt_);
#ifdef FB_PROFILE
                ++dw_chunks;
#endif
                if (p.dbg != nullptr && blockIdx.x == 0 && !dbg_done && group == 0) {
                    dbg_done = true;
                    // debug: read the first chunk back out of TMEM (128 lanes x CC columns, row pitch 128)
                    for (int c4 = 0; c4 < CC / 32; ++c4) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tmem + lane_sel + slot * CC + c4 * 32, r);
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) p.dbg[L * 128 + c4 * 32 + i] = r[i];
                    }
                }
                // ---- advance to my next chunk: chunk id += NG ----
                my_sub += NG;
                while (my_sub >= CPS) {
                    my_sub -= CPS;
                    ++my_il;
                    if (++stage == r_stages) { stage = 0; sphase ^= 1u; }
                }
                slot += NG;
                while (slot >= r_nchunk) { slot -= r_nchunk; apar ^= 1u; }
            }
            // every warp releases every scale chunk of the segment exactly once
            release_scales_upto(sc_seg0 + sc_count);
            sc_cur = -1;
            sc_seg0 += sc_count;
            seg_il0 = seg_il1;

            // ------------------------------- epilogue ------------------------------------
            PROF_ADD(dw_sc, dt_);
            mbar_wait(smem_u32(&ctl->acc_full), seg & 1, p.diag, p.timeout_ns, SITE_DQ_ACCFULL, 0, seg);
            tc_fence_after();
            if (p.trace != nullptr && threadIdx.x == 0 && seg == 0) p.trace[blockIdx.x * kTraceStride + 4] = globaltimer_ns();
            if (!synced) { pdl_wait_prior_grids(); synced = true; }   // D / workspace may be in use by the prior grid
            const bool full_k = (kb == 0) && (ke == p.k_iters);
            const int m_base = mt * p.mb;
            const int n_base = nt * TN;
            const int rows_valid = min(p.mb, p.M - m_base);
            float* accum = nullptr;
            int contributors = 1;
            if (!full_k) {
                // Partial K range: accumulate into the tile's fp32 scratch (zero on entry, left zero on exit)
                // with fire-and-forget reductions; the CTA that arrives last converts and writes the tile.
                const int tile_it0 = tile * p.k_iters;
                const int first_cta = streamk_cta_of(p, tile_it0, grid);
                contributors = streamk_cta_of(p, tile_it0 + p.k_iters - 1, grid) - first_cta + 1;
                accum = reinterpret_cast<float*>(p.workspace + p.partial_offset) + (size_t)tile * (NJ * p.mb * 128);
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if ((j % NG) != group) continue;
                const int n = n_base + nloc[j];
                for (int mc = 0; mc < rows_valid; mc += 16) {
                    uint32_t r[16];
                    tmem_ld_32x32b_x16(tmem + lane_sel + acc_col + j * p.mb + mc, r);
                    tc_wait_ld();
                    if (full_k) {
                        if (n < p.N) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (mc + i < rows_valid)
                                    p.D[(size_t)(m_base + mc + i) * p.N + n] = f32_to_t<BF16>(__uint_as_float(r[i]));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (mc + i < rows_valid) red_add_f32(accum + (j * p.mb + mc + i) * 128 + L, __uint_as_float(r[i]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&ctl->acc_empty));
            if (p.trace != nullptr && threadIdx.x == 0 && seg == 0) p.trace[blockIdx.x * kTraceStride + 5] = globaltimer_ns();

            if (!full_k) {
                asm volatile("bar.sync 1, %0;" ::"n"(kDequantWarps * 32) : "memory");
                if (threadIdx.x == 0) {
                    // release: publishes this CTA's reductions (cumulative through the barrier above);
                    // acquire: the last arriver sees everyone else's.
                    const int old = atom_add_acq_rel(reinterpret_cast<int*>(p.workspace) + tile, 1);
                    const int last = (old == contributors - 1) ? 1 : 0;
                    if (last) reinterpret_cast<int*>(p.workspace)[tile] = 0;   // self-resetting
                    ctl->is_last = last;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kDequantWarps * 32) : "memory");
                if (ctl->is_last) {
                    // one row at a time, every field of this warp in flight together (one L2 round trip per row, not
                    // one per element: the decode shapes have a single row)
                    for (int mi = 0; mi < rows_valid; ++mi) {
                        float v[NJ];
#pragma unroll
                        for (int j = 0; j < NJ; ++j)
                            if ((j % NG) == group) v[j] = __ldcg(accum + (j * p.mb + mi) * 128 + L);
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            if ((j % NG) != group) continue;
                            accum[(j * p.mb + mi) * 128 + L] = 0.f;
                            const int n = n_base + nloc[j];
                            if (n < p.N) p.D[(size_t)(m_base + mi) * p.N + n] = f32_to_t<BF16>(v[j]);
                        }
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kDequantWarps * 32) : "memory");   // is_last is reused by the next segment
                if (p.trace != nullptr && threadIdx.x == 0 && seg == 0) p.trace[blockIdx.x * kTraceStride + 6] = globaltimer_ns();
            }
            PROF_ADD(dw_epi, dt_);
            it += ke - kb;
            ++seg;
        }
#ifdef FB_PROFILE
        if (lane == 0 && (warp == 0 || warp == 4)) {
            const int o = (warp == 0) ? 16 : 24;
            PROF_OUT(o + 0, dw_sc); PROF_OUT(o + 1, dw_full); PROF_OUT(o + 2, dw_aempty); PROF_OUT(o + 3, dw_run);
            PROF_OUT(o + 4, dw_st); PROF_OUT(o + 5, dw_epi); PROF_OUT(o + 6, dw_chunks); PROF_OUT(o + 7, dw_scw);
        }
#endif
    }

    // ---- teardown ------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (p.trace != nullptr && threadIdx.x == 0) p.trace[blockIdx.x * kTraceStride + 7] = globaltimer_ns();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, F::TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------------------
// Host side
// ----------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled lookup_encode_fn() {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        return reinterpret_cast<PFN_encodeTiled>(ptr);
    return nullptr;
}
static PFN_encodeTiled get_encode_fn() {
    static const PFN_encodeTiled fn = lookup_encode_fn();   // C++11 magic static: initialised once, thread-safe
    return fn;
}

// Eager callers (one qgemm per Python call) would otherwise pay a cuTensorMapEncodeTiled per launch; the descriptor
// is a pure function of these eight values, so a small per-thread direct-mapped cache returns it.
namespace {
struct TmapKey {
    const void* base;
    uint64_t inner, outer, row_bytes;
    uint32_t box_inner, box_outer;
    int dt, swizzle;
    bool operator==(const TmapKey& o) const {
        return base == o.base && inner == o.inner && outer == o.outer && row_bytes == o.row_bytes &&
               box_inner == o.box_inner && box_outer == o.box_outer && dt == o.dt && swizzle == o.swizzle;
    }
};
struct TmapSlot {
    TmapKey key;
    CUtensorMap map;
    bool valid;
};
constexpr int kTmapSlots = 256;
thread_local TmapSlot t_tmap_cache[kTmapSlots];
}  // namespace

int make_tmap_2d(CUtensorMap* tm, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t outer,
                 uint64_t row_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle) {
    const TmapKey key{base, inner, outer, row_bytes, box_inner, box_outer, (int)dt, (int)swizzle};
    const uint64_t h = (reinterpret_cast<uint64_t>(base) >> 8) * 0x9E3779B97F4A7C15ull + outer * 31 + box_outer;
    TmapSlot& slot = t_tmap_cache[(h >> 40) & (kTmapSlots - 1)];
    if (slot.valid && slot.key == key) {
        *tm = slot.map;
        return FB_OK;
    }
    PFN_encodeTiled enc = get_encode_fn();
    if (enc == nullptr) return FB_ERR_DRIVER;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return FB_ERR_TENSORMAP;
    slot.key = key;
    slot.map = *tm;
    slot.valid = true;
    return FB_OK;
}

int make_tmap_3d(CUtensorMap* tm, CUtensorMapDataType dt, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                 CUtensorMapSwizzle swizzle) {
    PFN_encodeTiled enc = get_encode_fn();
    if (enc == nullptr) return FB_ERR_DRIVER;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? FB_OK : FB_ERR_TENSORMAP;
}

int qgemm_max_mb(int bits) { return bits == 4 ? 64 : bits == 2 ? 32 : 16; }

template <int BITS, bool BF16, bool SMALL>
static int launch_t(const QgemmArgs& a, cudaStream_t stream) {
    using F = Cfg<BITS, SMALL>;
    constexpr int TN = F::NJ * 128;
    constexpr int CC = F::NJ * F::K2C;
    QgemmParams p{};
    p.S = static_cast<const uint16_t*>(a.S);
    p.table2 = static_cast<const uint32_t*>(a.table2);
    p.D = static_cast<uint16_t*>(a.D);
    p.workspace = static_cast<uint8_t*>(a.workspace);
    p.diag = a.diag;
    p.dbg = a.dbg;
    p.trace = a.trace;
    p.ablate = a.ablate;
    p.static_weights = (a.flags & FB_FLAG_STATIC_WEIGHTS) ? 1 : 0;
    p.neg_zero2 = 0x80008000u;   // (-0, -0) in fp16 and bf16 alike; see ptx.cuh mul2()
    p.timeout_ns = a.timeout_ns;
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.group_size = a.group_size;
    p.G = a.K / a.group_size;
    p.group_shift = (a.group_size == 64) ? 0 : (a.group_size == 128) ? 1 : 2;   // stages (64 k) per group, log2
    p.tile_p = a.tile_p;
    int mb_max = SMALL ? 16 : qgemm_max_mb(BITS);
    int mb = ((a.M + 15) / 16) * 16;
    if (mb > mb_max) mb = mb_max;
    if (a.force_mb > 0 && a.force_mb <= mb_max) mb = a.force_mb;
    p.mb = mb;
    p.n_tiles = (a.N + TN - 1) / TN;
    p.m_tiles = (a.M + mb - 1) / mb;
    p.k_iters = a.K / kStageK;
    p.nchunk = (F::TMEM_COLS - F::NJ * mb) / CC;
    if (p.nchunk > kMaxChunkSlots) p.nchunk = kMaxChunkSlots;
    if (p.nchunk < 2) return FB_ERR_INTERNAL;
    p.plane1_row0 = (BITS == 3) ? a.N / 16 : 0;

    p.w_bytes = F::ROWS * 128;
    p.b_bytes = mb * 128;
    p.stage_bytes = (p.w_bytes + p.b_bytes + 1023u) & ~1023u;
    const uint32_t fixed = F::LUTN * F::LUT_STRIDE + 2 * F::SCH * F::NJ * 128 * 2 + sizeof(SmemCtl) + 1024 /*alignment slack*/;
    // 227 KB per CTA alone on an SM; (228 KB - 2 x 1 KB reserved) / 2 = 113 KB when two must fit
    const uint32_t smem_budget = SMALL ? 115712u : 232448u;
    int stages = (int)((smem_budget - fixed) / p.stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (a.force_stages > 0 && a.force_stages < stages) stages = a.force_stages;
    if (stages < 2) return FB_ERR_INTERNAL;
    p.stages = stages;
    const uint32_t smem_bytes = stages * p.stage_bytes + fixed;

    const long long tiles = (long long)p.n_tiles * p.m_tiles;
    const long long total = tiles * p.k_iters;
    if (total > 0x3fffffffLL) return FB_ERR_SHAPE;
    int grid = a.num_sms;
    if (a.force_grid > 0) grid = a.force_grid;
    p.streamk = (tiles < 4LL * grid) ? 1 : 0;
    if (a.force_streamk >= 0) p.streamk = a.force_streamk;
    if (p.streamk) { if (grid > total) grid = (int)total; }
    else           { if (grid > tiles) grid = (int)tiles; }
    // decode-shaped launches: a few SMs stay free for the next launch's late starters, shares aligned to tiles
    if (!SMALL && p.streamk && a.force_grid <= 0 && a.M <= 16 && p.m_tiles == 1)
        grid = decode_grid_for(total, p.k_iters, a.num_sms > 16 ? a.num_sms - 4 : a.num_sms, false);

    // workspace: [tile counters: fixed 64 KB][fp32 tile accumulators, one per output tile].  Both regions are
    // zero on entry (the caller zero-initialises the workspace once) and every kernel leaves what it touched
    // zero again, so calls of any shape can follow each other on a stream (contract of flute/utils.py:36-56).
    constexpr size_t kCounterBytes = 65536;
    p.partial_offset = (uint32_t)kCounterBytes;
    if (p.streamk) {
        const size_t need = kCounterBytes + (size_t)tiles * F::NJ * mb * 128 * 4;
        if ((size_t)tiles * 4 > kCounterBytes || need + prefill_scratch_bytes(a.num_sms) > a.workspace_bytes) return FB_ERR_WORKSPACE;
    }

    CUtensorMap tm_w, tm_a;
    const uint64_t P = (uint64_t)a.N / 16 * BITS;
    int rc = make_tmap_2d(&tm_w, CU_TENSOR_MAP_DATA_TYPE_UINT16, a.Q, (uint64_t)a.K, P, (uint64_t)a.K * 2, kStageK, 128,
                          CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FB_OK) return rc;
    rc = make_tmap_2d(&tm_a, BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, a.A,
                      (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.K * 2, kStageK, (uint32_t)mb, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FB_OK) return rc;

    auto kern = qgemm_sm100_kernel<BITS, BF16, SMALL>;
    static PerDeviceOnce attr_set;
    if (!attr_set.done(a.device)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_budget) != cudaSuccess) {
            cudaGetLastError();
            return FB_ERR_LAUNCH;
        }
        if (SMALL) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set.mark(a.device);
    }

    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(Roles<BITS, SMALL>::kThreads);
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tm_w, tm_a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return FB_ERR_LAUNCH;
    }
    return FB_OK;
}

int qgemm_launch(const QgemmArgs& a, cudaStream_t stream) {
    // a.variant: -1 auto, 0 LARGE, 1 SMALL.  Measured on B200 (profiles/r01_*): the dequantiser is bound by
    // shared-memory LUT wavefronts, not by residency, and LARGE (16 dequant warps, 8-stage ring) beats SMALL
    // (2 CTAs/SM, 3-stage ring) at every Llama shape, so automatic selection is LARGE.  SMALL stays reachable
    // for the cross-kernel-overlap experiments (flute_b200_set_variant(1)).
    // M <= 16, 2/4-bit: the decode kernel (qgemm_decode_sm100.cu) unless a test pins the general kernel
    // (variant 0 / 1, or explicit tiling overrides that only the general kernel understands).
    if ((a.variant < 0 || a.variant >= 2) && a.force_mb == 0 && qgemm_decode_supported(a)) return qgemm_decode_launch(a, stream);
    // M > 16, 4-bit: the prefill kernel (qgemm_prefill_sm100.cu)
    if ((a.variant < 0 || a.variant >= 2) && a.force_mb == 0 && qgemm_prefill_supported(a)) return qgemm_prefill_launch(a, stream);
    bool small = false;
    if (a.variant == 1 && (a.num_bits == 4 || a.num_bits == 2) && a.M <= 16) small = true;
    switch (a.num_bits * 4 + (a.bf16 ? 2 : 0) + (small ? 1 : 0)) {
        case 4 * 4 + 0: return launch_t<4, false, false>(a, stream);
        case 4 * 4 + 1: return launch_t<4, false, true>(a, stream);
        case 4 * 4 + 2: return launch_t<4, true, false>(a, stream);
        case 4 * 4 + 3: return launch_t<4, true, true>(a, stream);
        case 2 * 4 + 0: return launch_t<2, false, false>(a, stream);
        case 2 * 4 + 1: return launch_t<2, false, true>(a, stream);
        case 2 * 4 + 2: return launch_t<2, true, false>(a, stream);
        case 2 * 4 + 3: return launch_t<2, true, true>(a, stream);
        case 3 * 4 + 0: return launch_t<3, false, false>(a, stream);
        case 3 * 4 + 2: return launch_t<3, true, false>(a, stream);
    }
    return FB_ERR_BITS;
}


// Which kernel qgemm_launch picks for a problem (automatic selection; reporting only).
const char* qgemm_dispatch_name(int M, int num_bits, bool bf16) {
    QgemmArgs a{};
    a.M = M; a.num_bits = num_bits; a.bf16 = bf16; a.variant = -1;
    if (M >= 1 && qgemm_decode_supported(a)) {
        if (num_bits == 4 && M > 4) return bf16 ? "fb::dec::qgemm_decode_kernel<4,bf16,MC=16>" : "fb::dec::qgemm_decode_kernel<4,f16,MC=16>";
        if (num_bits == 4) return M == 1 ? (bf16 ? "fb::dec::qgemm_decode_kernel<4,bf16,MC=1>" : "fb::dec::qgemm_decode_kernel<4,f16,MC=1>")
                                         : (bf16 ? "fb::dec::qgemm_decode_kernel<4,bf16,MC=4>" : "fb::dec::qgemm_decode_kernel<4,f16,MC=4>");
        return M == 1 ? (bf16 ? "fb::dec::qgemm_decode_kernel<2,bf16,MC=1>" : "fb::dec::qgemm_decode_kernel<2,f16,MC=1>")
                      : (bf16 ? "fb::dec::qgemm_decode_kernel<2,bf16,MC=4>" : "fb::dec::qgemm_decode_kernel<2,f16,MC=4>");
    }
    if (qgemm_prefill_supported(a)) return bf16 ? "fb::pre::qgemm_prefill_kernel<bf16>" : "fb::pre::qgemm_prefill_kernel<f16>";
    switch (num_bits) {
        case 4: return bf16 ? "fb::qgemm_sm100_kernel<4,bf16,LARGE>" : "fb::qgemm_sm100_kernel<4,f16,LARGE>";
        case 3: return bf16 ? "fb::qgemm_sm100_kernel<3,bf16,LARGE>" : "fb::qgemm_sm100_kernel<3,f16,LARGE>";
        case 2: return bf16 ? "fb::qgemm_sm100_kernel<2,bf16,LARGE>" : "fb::qgemm_sm100_kernel<2,f16,LARGE>";
    }
    return "unsupported";
}

}  // namespace fb
