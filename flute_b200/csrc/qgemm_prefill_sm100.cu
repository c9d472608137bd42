// Prefill-shaped LUT-quantized GEMM for sm_100a (M > 16, 4-bit):
//   D[M,N] = A[M,K] . W_hat[K,N],   W_hat[k,n] = round_T(table2[code(k/2,n)].{lo,hi} * S[n, k/group])
//
// Same role as the reference's large-M templates (flute/csrc/qgemm_kernel.hpp:24-939, config.hpp:238-325) and
// the same dequantisation arithmetic (one multiply in T, rounded to T, before the MMA:
// packbits_utils.hpp:105,139).  Structure shared with the decode kernel (qgemm_decode_sm100.cu): weights are
// the tcgen05 "A" operand in TENSOR MEMORY (TMEM lane = packed row), activations the "B" operand in shared
// memory, all 16 dequantiser warps convert quarter-row pieces of the same 64-k stage.  Differences:
//   * the group scale is multiplied in while dequantising (one fma.rn.{f16,bf16}x2 per pair): bit-identical
//     to the reference's W_hat;
//   * a tile is 128 packed rows x TWO of the four pair fields (256 output columns) x 128 activation rows.
//     Dequantisation is the scarce resource (measured ~540 cycles per 64-k stage for four fields against
//     ~520 cycles of tensor work), so every dequantised weight is used for 128 rows instead of 64 and a stage
//     converts two fields instead of four; the other field pair of the same packed rows is another tile (the
//     packed words are fetched twice, from L2);
//   * fp32 accumulators (2 fields x 128 columns) live in TMEM across the K loop; tcgen05.mma N = 128 runs at the
//     full tensor rate with the A operand in TMEM (64.6 cycles per 128x128x16, tools/mma_rate_probe.cu);
//   * activations arrive by TMA (128 rows x 64 k per stage);
//   * a K range that does not cover a whole tile (Stream-K) is written as plain fp32 vectors to a per-CTA
//     scratch slot and summed by the last CTA to arrive: 16K atomics per partial tile cost ~40k cycles.
//
// Warp roles (608 threads, 1 CTA/SM, persistent):
//   warps 0-15   dequantisers: two sets of 8 (w/8) convert alternate 64-k stages, one stage apart -- 4 quads of the row
//                per warp and turn, so the per-stage fixed cost (barrier probes, scale loads, loop) is paid once per 32
//                look-ups instead of once per 16 -- four TMEM A slots ahead of the MMAs; at the end of a tile the same
//                warps run the epilogue: warp w owns field (w/4)&1, activation-row half w/8, lane quarter w%4
//   warp  16     TMA producer (weights + activations per stage)
//   warp  17     tcgen05.mma issuer, TMEM allocator
//   warp  18     scale blocks ([tile columns] x [8 groups]) by cp.async
#include "ptx.cuh"
#include "qgemm_sm100.h"

#include <cuda.h>

namespace fb {
namespace pre {

#ifdef FB_PROFILE
#define PPROF_DECL(...) long long __VA_ARGS__
#define PPROF_T0(t) long long t = clock64()
#define PPROF_ADD(acc, t) do { long long _n = clock64(); acc += _n - t; t = _n; } while (0)
#define PPROF_OUT(slot, v) do { if (p.trace != nullptr) p.trace[blockIdx.x * 48 + (slot)] = (unsigned long long)(v); } while (0)
#else
#define PPROF_DECL(...)
#define PPROF_T0(t)
#define PPROF_ADD(acc, t)
#define PPROF_OUT(slot, v)
#endif

constexpr int NJ = 4;                // pair fields per packed word (4-bit)
constexpr int NF = 2;                // fields per tile
constexpr int kMb = 128;             // activation rows per tile == MMA N
constexpr int kDqWarps = 16;
constexpr int kProducerWarp = 16;
constexpr int kMmaWarp = 17;
constexpr int kScaleWarp = 18;
constexpr int kThreads = 19 * 32;
constexpr int kMaxStages = 4;
constexpr int kScSlots = 3;
constexpr int kLutStride = 256;
constexpr int kWBytes = 128 * 128;
constexpr int kBBytes = kMb * 128;
constexpr int kStageBytes = kWBytes + kBBytes;
constexpr int TN = NJ * 128;
constexpr uint32_t kScBytes = TN * 16;
constexpr int AS = 4;                // TMEM A slots (NF * 32 = 64 columns each)
constexpr uint32_t kACols = NF * 32;
constexpr uint32_t kDCol0 = AS * kACols;
constexpr int DQG = 2;               // dequantiser sets: 8 warps each, alternate stages (one stage apart)
constexpr size_t kPartBytes = (size_t)NF * kMb * 128 * 4;   // one partial tile: fp32, float4 index = rows/4 * 512 + thread (16 warps x 32 lanes)

struct Ctl {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t a_full[AS];
    uint64_t acc_full;
    uint64_t acc_empty;
    uint64_t sc_full[kScSlots];
    uint64_t sc_empty[kScSlots];
    uint32_t tmem_base;
    int is_last;
};

struct Params {
    const uint16_t* S;
    const uint32_t* table2;
    uint16_t* D;
    uint8_t* workspace;
    Diag* diag;
    unsigned long long* trace;
    unsigned long long timeout_ns;   // barrier-wait bound (flute_b200_set_timeout_ms); 0 = unbounded
    int M, N, K, G;
    int tile_p;
    int gshift;
    int n_tiles, m_tiles, k_iters;
    int stages;
    int streamk;
    int dp_tiles;            // Stream-K: the first dp_tiles tiles are taken whole (dp_tiles / grid per CTA), only the rest is split
    int tma_scales;
    size_t scratch_offset;   // per-CTA partial-tile slots (2 per CTA) at the end of the workspace
};

enum : int { PSITE_FULL = 41, PSITE_ASLOT, PSITE_ACCFULL, PSITE_SCFULL, PSITE_EMPTY, PSITE_SCEMPTY, PSITE_AFULL, PSITE_ACCEMPTY };

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
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity, const Params& p, int site, int iter = 0) {
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
// keep a kernel parameter in a register (ptxas otherwise re-reads the constant bank on every use)
__device__ __forceinline__ int pin(int v) {
    asm volatile("" : "+r"(v));
    return v;
}
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
template <int J>
__device__ __forceinline__ uint32_t code_lane(uint32_t w, uint32_t lane4) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(lane4), "n"(0x6504 + (J << 4)));
    return r;
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ int atom_add_acq_rel(int* addr, int v) {
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void tma_load_3d_nohint(uint32_t dst_smem, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ int n_local(int L, int j, int tile_p) {
    if (tile_p == 32) return (L >> 5) * (NJ * 32) + j * 32 + (L & 31);
    return (L >> 6) * (NJ * 64) + j * 64 + (L & 63);
}

struct Range {
    int it0, it1;
};
// tile -> (packed-row block nt, field pair fp, activation-row block mt); mt fastest so that consecutive tiles share W
struct TileCoord {
    int nt, fp, mt;
};
__device__ __forceinline__ TileCoord tile_coord(const Params& p, int tile) {
    TileCoord c;
    const int per_nt = 2 * p.m_tiles;
    c.nt = tile / per_nt;
    const int r = tile - c.nt * per_nt;
    c.fp = r / p.m_tiles;
    c.mt = r - c.fp * p.m_tiles;
    return c;
}
// Work of CTA b: part 0 = whole tiles, part 1 = its share of the Stream-K region (the tiles after dp_tiles, split at
// stage granularity).  Whole tiles first keeps the split -- and its fix-up through global scratch -- to the last
// wave-and-a-bit: at 4096 x 4096, M = 4096 (512 tiles on 148 CTAs) 296 tiles need no fix-up and 216 are split with
// ~1.5 tiles of work per CTA, instead of every CTA boundary of the whole problem falling inside a tile.
__device__ __forceinline__ Range cta_part(const Params& p, int b, int grid, int part) {
    Range r;
    const int tiles = p.n_tiles * 2 * p.m_tiles;
    if (!p.streamk) {
        if (part == 1) { r.it0 = r.it1 = 0; return r; }
        const int base = tiles / grid, rem = tiles - base * grid;
        const int t0 = b * base + min(b, rem);
        r.it0 = t0 * p.k_iters;
        r.it1 = (t0 + base + (b < rem ? 1 : 0)) * p.k_iters;
        return r;
    }
    if (part == 0) {
        const int w = p.dp_tiles / grid;
        r.it0 = b * w * p.k_iters;
        r.it1 = (b + 1) * w * p.k_iters;
        return r;
    }
    const int u0 = p.dp_tiles * p.k_iters;
    const int total = (tiles - p.dp_tiles) * p.k_iters;
    const int base = total / grid, rem = total - base * grid;
    r.it0 = u0 + b * base + min(b, rem);
    r.it1 = r.it0 + base + (b < rem ? 1 : 0);
    return r;
}
// which CTA's Stream-K share holds stage-unit `it` (it >= dp_tiles * k_iters)
__device__ __forceinline__ int cta_of(const Params& p, int it, int grid) {
    const int tiles = p.n_tiles * 2 * p.m_tiles;
    const int u0 = p.dp_tiles * p.k_iters;
    const int total = (tiles - p.dp_tiles) * p.k_iters;
    const int base = total / grid, rem = total - base * grid;
    const int thr = rem * (base + 1);
    const int x = it - u0;
    return x < thr ? x / (base + 1) : rem + (x - thr) / base;
}

// two 16-byte quads (8 consecutive k-pairs) of row L -> fields 2*FP, 2*FP+1 x 8 TMEM columns, scaled
template <bool BF16, int FP>
__device__ __forceinline__ void piece(uint32_t row, int pq0, int pq1, uint32_t lut, uint32_t lane4, const uint32_t (&sc)[2],
                                      uint32_t nz, uint32_t tcol) {
    const uint4 v0 = lds128(row + (uint32_t)(pq0 << 4));
    const uint4 v1 = lds128(row + (uint32_t)(pq1 << 4));
    const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t r[2][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r[0][i] = mul2<BF16>(lds32(lut + code_lane<2 * FP>(w[i], lane4)), sc[0], nz);
        r[1][i] = mul2<BF16>(lds32(lut + code_lane<2 * FP + 1>(w[i], lane4)), sc[1], nz);
    }
    tmem_st_x8(tcol, r[0]);
    tmem_st_x8(tcol + 32, r[1]);
}

template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1)
qgemm_prefill_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t ring = smem_base;
    const uint32_t sc_smem = ring + p.stages * kStageBytes;
    const uint32_t lut = sc_smem + kScSlots * kScBytes;
    Ctl* ctl = reinterpret_cast<Ctl*>(smem_gen + (lut + 256 * kLutStride - smem_base));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int grid = gridDim.x;
    const Range rg_parts[2] = {cta_part(p, blockIdx.x, grid, 0), cta_part(p, blockIdx.x, grid, 1)};
    const int spg_mask = (1 << p.gshift) - 1;
    const int S = p.stages;

    if (p.trace != nullptr && threadIdx.x == 0) p.trace[blockIdx.x * 48 + 0] = globaltimer_ns();
    pdl_wait_prior_grids();   // activations (and possibly weights) come from earlier work on the stream

    if (warp == kProducerWarp && lane == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_a);
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&ctl->full[s]), 1);
            mbar_init(smem_u32(&ctl->empty[s]), 1);
        }
        for (int s = 0; s < AS; ++s) mbar_init(smem_u32(&ctl->a_full[s]), kDqWarps / DQG);
        mbar_init(smem_u32(&ctl->acc_full), 1);
        mbar_init(smem_u32(&ctl->acc_empty), kDqWarps);
        for (int s = 0; s < kScSlots; ++s) {
            mbar_init(smem_u32(&ctl->sc_full[s]), 32);
            mbar_init(smem_u32(&ctl->sc_empty[s]), kDqWarps);
        }
        mbar_fence_init();
    }
    if (warp == kMmaWarp) {
        tmem_alloc(smem_u32(&ctl->tmem_base), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = ctl->tmem_base;
    pdl_launch_dependents();

    if (warp == kProducerWarp) {
        // =============================== TMA producer ===============================
        {
            const uint64_t pol_w = policy_evict_last();     // weight tiles are re-read by the other m-tiles
            const uint64_t pol_a = policy_evict_last();
            int stage = 0;
            uint32_t ephase = 1;
            int n_total = 0;
            PPROF_DECL(pw_sc = 0, pw_empty = 0, pw_issue = 0);
            PPROF_T0(pt);
            for (int part = 0; part < 2; ++part) {
                const Range rg = rg_parts[part];
                if (rg.it1 <= rg.it0) continue;
                const int n_it = rg.it1 - rg.it0;
                n_total += n_it;
                int tile = rg.it0 / p.k_iters;
                int k = rg.it0 - tile * p.k_iters;
                TileCoord tc = tile_coord(p, tile);
                for (int i = 0; i < n_it; ++i) {
                    PPROF_ADD(pw_sc, pt);
                    wait(smem_u32(&ctl->empty[stage]), ephase, p, PSITE_EMPTY);
                    PPROF_ADD(pw_empty, pt);
                    if (elect_one()) {
                        const uint32_t bar = smem_u32(&ctl->full[stage]);
                        mbar_arrive_expect_tx(bar, kStageBytes);
                        tma_load_2d(ring + stage * kStageBytes, &tmap_w, bar, k * 64, tc.nt * 128, pol_w);
                        tma_load_2d(ring + stage * kStageBytes + kWBytes, &tmap_a, bar, k * 64, tc.mt * kMb, pol_a);
                    }
                    __syncwarp();
                    PPROF_ADD(pw_issue, pt);
                    if (++stage == S) { stage = 0; ephase ^= 1u; }
                    if (++k == p.k_iters) {
                        k = 0;
                        tc = tile_coord(p, ++tile);
                    }
                }
            }
            if (lane == 0) { PPROF_OUT(8, pw_sc); PPROF_OUT(9, pw_empty); PPROF_OUT(10, pw_issue); PPROF_OUT(12, n_total); }
        }
    } else if (warp == kMmaWarp) {
        // =============================== MMA issuer =================================
        {
            const uint32_t idesc = make_idesc_f16(BF16, 128, kMb);
            int stage = 0;
            int aslot = 0;
            uint32_t aphase = 0;
            int seg = 0;
            PPROF_DECL(mw_acc = 0, mw_afull = 0, mw_issue = 0);
            PPROF_T0(mt_);
            for (int part = 0; part < 2; ++part) {
                const Range rg = rg_parts[part];
                for (int it = rg.it0; it < rg.it1;) {
                    const int tile = it / p.k_iters;
                    const int kb = it - tile * p.k_iters;
                    const int ke = min(p.k_iters, kb + (rg.it1 - it));
                    wait(smem_u32(&ctl->acc_empty), (uint32_t)(seg & 1) ^ 1u, p, PSITE_ACCEMPTY);
                    PPROF_ADD(mw_acc, mt_);
                    for (int k = kb; k < ke; ++k) {
                        const uint64_t bdesc = make_smem_desc_sw128(ring + stage * kStageBytes + kWBytes);
                        wait(smem_u32(&ctl->a_full[aslot]), aphase, p, PSITE_AFULL);
                        PPROF_ADD(mw_afull, mt_);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t a_base = tmem + aslot * kACols;
                            const uint32_t d_base = tmem + kDCol0;
                            const uint32_t first = (k == kb) ? 0u : 1u;
#pragma unroll
                            for (int f = 0; f < NF; ++f) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    tc_mma_ts(d_base + f * kMb, a_base + f * 32 + kk * 8, bdesc + (uint64_t)((kk * 32) >> 4), idesc,
                                              kk == 0 ? first : 1u);
                            }
                            tc_commit(smem_u32(&ctl->empty[stage]));
                            if (k == ke - 1) tc_commit(smem_u32(&ctl->acc_full));
                        }
                        __syncwarp();
                        if (++aslot == AS) { aslot = 0; aphase ^= 1u; }
                        if (++stage == S) stage = 0;
                        PPROF_ADD(mw_issue, mt_);
                    }
                    it += ke - kb;
                    ++seg;
                }
            }
            if (lane == 0) { PPROF_OUT(13, mw_acc); PPROF_OUT(14, mw_afull); PPROF_OUT(16, mw_issue); }
        }
    } else if (warp == kScaleWarp) {
        // =============================== scale blocks ===============================
        int nb = 0;
        for (int part = 0; part < 2; ++part) {
            const Range rg = rg_parts[part];
            if (rg.it1 <= rg.it0) continue;
            const int n_it = rg.it1 - rg.it0;
            int tile = rg.it0 / p.k_iters;
            int k = rg.it0 - tile * p.k_iters;
            int nt = tile_coord(p, tile).nt;
            int last_blk = -1;
            for (int i = 0; i < n_it; ++i) {
                const int blk = (k >> p.gshift) >> 3;
                if (blk != last_blk) {
                    const int slot = nb % kScSlots;
                    const uint32_t par = ((nb / kScSlots) & 1) ^ 1u;
                    wait(smem_u32(&ctl->sc_empty[slot]), par, p, PSITE_SCEMPTY);
                    const uint32_t dst = sc_smem + slot * kScBytes;
                    if (p.tma_scales) {
#pragma unroll 4
                        for (int r = lane; r < TN; r += 32) {
                            const int n = nt * TN + r;
                            if (n < p.N) {
                                const uint16_t* src = p.S + (size_t)n * p.G + blk * 8;
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + r * 16), "l"(src) : "memory");
                            }
                        }
                        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&ctl->sc_full[slot])) : "memory");
                    } else {
                        uint16_t* d16 = reinterpret_cast<uint16_t*>(smem_gen + (dst - smem_base));
                        for (int r = lane; r < TN; r += 32) {
                            const int n = nt * TN + r;
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
                }
                if (++k == p.k_iters) {
                    k = 0;
                    last_blk = -1;
                    nt = tile_coord(p, ++tile).nt;
                }
            }
        }
    } else {
        // ========================= dequantisers + epilogue ==========================
        const int q = warp & 3;
        const int sw = warp >> 2;               // quads 2*sw, 2*sw+1 of the row; field sw in the epilogue
        const int L = q * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const uint32_t lane4 = (uint32_t)lane * 4;
        {   // lane-replicated LUT: entry e for lane l at lut + e*256 + l*4
            uint32_t* lut_gen = reinterpret_cast<uint32_t*>(smem_gen + (lut - smem_base));
            const int t = threadIdx.x;      // 0..511
            const uint32_t v = __ldg(p.table2 + (t & 255));
            const int half = t >> 8;
#pragma unroll
            for (int l = 0; l < 16; ++l) lut_gen[(t & 255) * (kLutStride / 4) + half * 16 + l] = v;
            asm volatile("bar.sync 1, %0;" ::"n"(kDqWarps * 32) : "memory");
        }
        const uint32_t wrow = (uint32_t)L * 128;
        const int xq = L & 7;
        const uint32_t nz = neg_zero2();
        const int ef = sw & 1, emh = sw >> 1;          // epilogue: field of the pair, activation-row half

        const int set = sw >> 1;                // which stages (global stage parity) this warp converts
        const int hw = sw & 1;                  // its quads of the row: 4*hw .. 4*hw + 3
        const int r_gshift = pin(p.gshift), r_k_iters = pin(p.k_iters);
        int sc_idx = 0;
        uint32_t sc_par = 0;
        int gs = 0;                             // stages since the CTA started (ring slot gs % S, A slot gs % AS)
        int stage = 0;
        uint32_t fphase = 0;
        int aslot = 0;
        int seg = 0;
        PPROF_DECL(dw_sc = 0, dw_full = 0, dw_aslot = 0, dw_piece = 0, dw_st = 0, dw_epiw = 0, dw_epi = 0);
        PPROF_T0(dt);
        for (int part = 0; part < 2; ++part) {
        const Range rg = rg_parts[part];
        for (int it = rg.it0; it < rg.it1;) {
            const int tile = it / r_k_iters;
            const int kb = it - tile * r_k_iters;
            const int ke = min(r_k_iters, kb + (rg.it1 - it));
            const TileCoord tc = tile_coord(p, tile);
            const int nl0 = n_local(L, 2 * tc.fp, p.tile_p), nl1 = n_local(L, 2 * tc.fp + 1, p.tile_p);
            int cur_blk = -1;
            int sc_g = -1;
            uint32_t sc[2] = {0, 0};
            for (int k = kb; k < ke; ++k, ++gs) {
                const int g = k >> r_gshift;
                // scale-block bookkeeping on EVERY stage (also the other set's): each block must be released by all 16 warps
                if ((g >> 3) != cur_blk) {
                    if (cur_blk >= 0) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&ctl->sc_empty[sc_idx]));
                        if (++sc_idx == kScSlots) { sc_idx = 0; sc_par ^= 1u; }
                    }
                    cur_blk = g >> 3;
                    wait(smem_u32(&ctl->sc_full[sc_idx]), sc_par, p, PSITE_SCFULL);
                }
                if ((gs & (DQG - 1)) == set) {
                    if (g != sc_g) {
                        const uint32_t sc_base = sc_smem + sc_idx * kScBytes + (g & 7) * 2;
                        const uint32_t s0 = lds16(sc_base + nl0 * 16), s1 = lds16(sc_base + nl1 * 16);
                        sc[0] = s0 | (s0 << 16);
                        sc[1] = s1 | (s1 << 16);
                        sc_g = g;
                    }
                    PPROF_ADD(dw_sc, dt);
                    // (the stage's A slot is free as soon as its data has landed: with as many A slots as ring stages the
                    // producer could only refill this stage after the MMAs that read slot `aslot` four stages ago completed)
                    wait(smem_u32(&ctl->full[stage]), fphase, p, PSITE_FULL);
                    PPROF_ADD(dw_full, dt);
                    tc_fence_after();
                    const uint32_t row = ring + stage * kStageBytes + wrow;
                    const uint32_t tcol = tmem + lane_sel + aslot * kACols + hw * 16;
                    if (tc.fp == 0) {
                        piece<BF16, 0>(row, (4 * hw) ^ xq, (4 * hw + 1) ^ xq, lut, lane4, sc, nz, tcol);
                        piece<BF16, 0>(row, (4 * hw + 2) ^ xq, (4 * hw + 3) ^ xq, lut, lane4, sc, nz, tcol + 8);
                    } else {
                        piece<BF16, 1>(row, (4 * hw) ^ xq, (4 * hw + 1) ^ xq, lut, lane4, sc, nz, tcol);
                        piece<BF16, 1>(row, (4 * hw + 2) ^ xq, (4 * hw + 3) ^ xq, lut, lane4, sc, nz, tcol + 8);
                    }
                    PPROF_ADD(dw_piece, dt);
                    tc_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&ctl->a_full[aslot]));
                    PPROF_ADD(dw_st, dt);
                }
                if (++aslot == AS) aslot = 0;
                if (++stage == S) { stage = 0; fphase ^= 1u; }
            }
            if (cur_blk >= 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl->sc_empty[sc_idx]));
                if (++sc_idx == kScSlots) { sc_idx = 0; sc_par ^= 1u; }
            }

            // ---------------- epilogue: field ef of the pair, rows [emh*64, emh*64+64), lane quarter q ----------------
            wait(smem_u32(&ctl->acc_full), (uint32_t)(seg & 1), p, PSITE_ACCFULL);
            tc_fence_after();
            PPROF_ADD(dw_epiw, dt);
            const bool full_k = (kb == 0) && (ke == p.k_iters);
            const int m0 = tc.mt * kMb + emh * 64;
            const int n = tc.nt * TN + (ef ? nl1 : nl0);
            const int rows_valid = min(64, p.M - m0);     // may be <= 0
            const uint32_t dcol = tmem + lane_sel + kDCol0 + ef * kMb + emh * 64;
            if (full_k) {
#pragma unroll 1
                for (int mc = 0; mc < rows_valid; mc += 16) {
                    uint32_t r[16];
                    tmem_ld_32x32b_x16(dcol + mc, r);
                    tc_wait_ld();
                    if (n < p.N) {
#pragma unroll
                        for (int x = 0; x < 16; ++x)
                            if (mc + x < rows_valid) p.D[(size_t)(m0 + mc + x) * p.N + n] = f32_to_t<BF16>(__uint_as_float(r[x]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl->acc_empty));
            } else {
                // Partial K range.  This CTA's partial tile goes to one of its two scratch slots (slot 0: the
                // segment starts at the CTA's first stage; slot 1: the CTA's last segment); the CTA that arrives last on
                // the tile's counter sums every contributor's slot.  Slot layout: float4 number (c4 * 512 + thread) holds
                // accumulator rows 4*c4 .. 4*c4+3 of that thread, so a warp's store or load is one 512-byte run.
                const int tile_it0 = tile * p.k_iters;
                const int first_cta = cta_of(p, tile_it0, grid);
                const int last_cta = cta_of(p, tile_it0 + p.k_iters - 1, grid);
                float4* scratch = reinterpret_cast<float4*>(p.workspace + p.scratch_offset);
                const size_t part_f4 = kPartBytes / 16;
                const int tid = warp * 32 + lane;
                {
                    float4* dst = scratch + ((size_t)blockIdx.x * 2 + (it == rg.it0 ? 0 : 1)) * part_f4 + tid;
#pragma unroll 1
                    for (int mc = 0; mc < 64; mc += 16) {
                        uint32_t r[16];
                        tmem_ld_32x32b_x16(dcol + mc, r);
                        tc_wait_ld();
#pragma unroll
                        for (int x = 0; x < 16; x += 4)
                            __stcg(reinterpret_cast<uint4*>(dst + (size_t)((mc + x) >> 2) * (kDqWarps * 32)),
                                   make_uint4(r[x], r[x + 1], r[x + 2], r[x + 3]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl->acc_empty));
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"n"(kDqWarps * 32) : "memory");
                if (threadIdx.x == 0) {
                    const int old = atom_add_acq_rel(reinterpret_cast<int*>(p.workspace) + tile, 1);
                    const int last = (old == last_cta - first_cta) ? 1 : 0;
                    if (last) reinterpret_cast<int*>(p.workspace)[tile] = 0;
                    ctl->is_last = last;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kDqWarps * 32) : "memory");
                if (ctl->is_last) {
                    __threadfence();
                    // contributor c's slot for this tile: 0 if its Stream-K share starts inside the tile, else 1
                    auto slot_of = [&](int c) -> const float4* {
                        const Range rc = cta_part(p, c, grid, 1);
                        return scratch + ((size_t)c * 2 + (rc.it0 >= tile_it0 ? 0 : 1)) * part_f4 + tid;
                    };
#pragma unroll 1
                    for (int mc = 0; mc < rows_valid; mc += 16) {
                        float acc[16];
#pragma unroll
                        for (int x = 0; x < 16; ++x) acc[x] = 0.f;
                        // two contributors' loads in flight at a time (a tile usually has two or three)
#pragma unroll 1
                        for (int c = first_cta; c <= last_cta; c += 2) {
                            const float4* s0 = slot_of(c) + (size_t)(mc >> 2) * (kDqWarps * 32);
                            const bool two = c + 1 <= last_cta;
                            const float4* s1 = two ? slot_of(c + 1) + (size_t)(mc >> 2) * (kDqWarps * 32) : s0;
                            float4 v0[4], v1[4];
#pragma unroll
                            for (int x = 0; x < 4; ++x) v0[x] = __ldcg(s0 + (size_t)x * (kDqWarps * 32));
#pragma unroll
                            for (int x = 0; x < 4; ++x) v1[x] = two ? __ldcg(s1 + (size_t)x * (kDqWarps * 32)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int x = 0; x < 4; ++x) {
                                acc[4 * x] += v0[x].x; acc[4 * x + 1] += v0[x].y; acc[4 * x + 2] += v0[x].z; acc[4 * x + 3] += v0[x].w;
                                acc[4 * x] += v1[x].x; acc[4 * x + 1] += v1[x].y; acc[4 * x + 2] += v1[x].z; acc[4 * x + 3] += v1[x].w;
                            }
                        }
                        if (n < p.N) {
#pragma unroll
                            for (int x = 0; x < 16; ++x)
                                if (mc + x < rows_valid) p.D[(size_t)(m0 + mc + x) * p.N + n] = f32_to_t<BF16>(acc[x]);
                        }
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kDqWarps * 32) : "memory");   // is_last is reused by the next segment
            }
            PPROF_ADD(dw_epi, dt);
            it += ke - kb;
            ++seg;
        }
        }
#ifdef FB_PROFILE
        if (lane == 0 && (warp == 0 || warp == 9)) {
            const int o = (warp == 0) ? 24 : 32;
            PPROF_OUT(o + 0, dw_full); PPROF_OUT(o + 1, dw_aslot); PPROF_OUT(o + 2, dw_piece); PPROF_OUT(o + 3, dw_st); PPROF_OUT(o + 4, dw_sc);
            PPROF_OUT(o + 5, dw_epiw); PPROF_OUT(o + 6, dw_epi);
        }
#endif
    }

    // ---- teardown ------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (p.trace != nullptr && threadIdx.x == 0) p.trace[blockIdx.x * 48 + 7] = globaltimer_ns();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <bool BF16>
static int launch_t(const QgemmArgs& a, cudaStream_t stream) {
    Params p{};
    p.S = static_cast<const uint16_t*>(a.S);
    p.table2 = static_cast<const uint32_t*>(a.table2);
    p.D = static_cast<uint16_t*>(a.D);
    p.workspace = static_cast<uint8_t*>(a.workspace);
    p.diag = a.diag;
    p.trace = a.trace;
    p.timeout_ns = a.timeout_ns;
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.G = a.K / a.group_size;
    p.tile_p = a.tile_p;
    p.gshift = (a.group_size == 64) ? 0 : (a.group_size == 128) ? 1 : 2;
    p.n_tiles = (a.N + TN - 1) / TN;
    p.m_tiles = (a.M + kMb - 1) / kMb;
    p.k_iters = a.K / 64;
    p.tma_scales = ((p.G % 8) == 0 && (reinterpret_cast<uintptr_t>(a.S) & 15) == 0) ? 1 : 0;   // 16-byte scale rows: cp.async

    const uint32_t fixed = kScSlots * kScBytes + 256 * kLutStride + sizeof(Ctl) + 1024;
    const uint32_t smem_budget = 232448u;
    int stages = (int)((smem_budget - fixed) / kStageBytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (a.force_stages > 0 && a.force_stages < stages) stages = a.force_stages;
    if (stages < 2) return FB_ERR_INTERNAL;
    p.stages = stages;
    const uint32_t smem_bytes = stages * kStageBytes + fixed;

    const long long tiles = (long long)p.n_tiles * 2 * p.m_tiles;
    const long long total = tiles * p.k_iters;
    if (total > 0x3fffffffLL || tiles >= (1 << 19)) return FB_ERR_SHAPE;
    int grid = a.num_sms;
    if (a.force_grid > 0) grid = a.force_grid;
    // Whole tiles per CTA, or Stream-K?  In stage units: whole tiles cost ceil(tiles / grid) * k_iters per CTA; Stream-K
    // costs the even share plus the partial-tile hand-over (slot write, counter, the last arriver's reduction), which
    // measured ~16 + k_iters / 8 stages (gpurun r02p2, M = 128 .. 4096 on the four Llama-3-8B shapes: e.g. 4096x4096
    // M = 1024, 128 tiles: whole 38 us vs split 46; M = 512, 64 tiles: whole 38 us vs split 34).
    constexpr size_t kCounterBytes = 65536;
    {
        const long long waves = (tiles + grid - 1) / grid;
        const long long whole_cost = waves * p.k_iters;
        const long long split_cost = (total + grid - 1) / grid + 16 + p.k_iters / 8;
        p.streamk = (split_cost < whole_cost) ? 1 : 0;
    }
    if (a.force_streamk >= 0) p.streamk = a.force_streamk;
    // Stream-K needs one arrival counter per tile in the 64 KB counter region; beyond that (very large M x N, where
    // the last-wave loss is negligible anyway) every CTA takes whole tiles and no counter is touched.
    if ((size_t)tiles * 4 > kCounterBytes) p.streamk = 0;
    // data-parallel waves, then ONE Stream-K wave over the remainder plus a full wave's worth of tiles: every CTA's
    // split share stays >= one tile of work and a tile has 2-3 contributors, not dozens
    p.dp_tiles = (p.streamk && tiles / grid >= 2) ? (int)((tiles / grid - 1) * grid) : 0;
    if (p.streamk) { if (grid > total) grid = (int)total; }
    else           { if (grid > tiles) grid = (int)tiles; }

    // workspace: [tile counters: fixed 64 KB, zero between launches][... zero-invariant fp32 accumulators of the
    // other kernels ...][2 partial-tile slots per CTA at the very end, contents undefined between launches]
    const size_t scratch = prefill_scratch_bytes(a.num_sms);
    if (a.workspace_bytes < kCounterBytes + scratch || grid > a.num_sms) return FB_ERR_WORKSPACE;
    p.scratch_offset = (a.workspace_bytes - scratch) & ~(size_t)255;

    CUtensorMap tm_w, tm_a;
    const uint64_t P = (uint64_t)a.N / 16 * 4;
    int rc = make_tmap_2d(&tm_w, CU_TENSOR_MAP_DATA_TYPE_UINT16, a.Q, (uint64_t)a.K, P, (uint64_t)a.K * 2, 64, 128,
                          CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FB_OK) return rc;
    rc = make_tmap_2d(&tm_a, BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, a.A, (uint64_t)a.K,
                      (uint64_t)a.M, (uint64_t)a.K * 2, 64, kMb, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != FB_OK) return rc;
    auto kern = qgemm_prefill_kernel<BF16>;
    static PerDeviceOnce attr_set;
    if (!attr_set.done(a.device)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_budget) != cudaSuccess) {
            cudaGetLastError();
            return FB_ERR_LAUNCH;
        }
        attr_set.mark(a.device);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
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

}  // namespace pre

bool qgemm_prefill_supported(const QgemmArgs& a) { return a.num_bits == 4 && a.M > 16; }

int qgemm_prefill_launch(const QgemmArgs& a, cudaStream_t stream) {
    return a.bf16 ? pre::launch_t<true>(a, stream) : pre::launch_t<false>(a, stream);
}

}  // namespace fb
