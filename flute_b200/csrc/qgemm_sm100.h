// Internal host <-> kernel interface of the sm_100a LUT-qGEMM (not part of the C ABI).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <stddef.h>
#include <stdint.h>

#include "../../include/flute_b200.h"

namespace fb {

struct Diag;

enum : int { FB_FLAG_PDL = 1, FB_FLAG_STATIC_WEIGHTS = 2 };

// Arguments as the C ABI receives them, plus test-only overrides (0 / -1 = engine's choice).
struct QgemmArgs {
    const void* A;
    const void* Q;
    void* D;
    const void* S;
    const void* table2;
    void* workspace;
    size_t workspace_bytes;
    int M, N, K;
    int num_bits, group_size, tile_p;
    int bf16;
    int flags;
    int device;
    int num_sms;
    Diag* diag;
    uint32_t* dbg;
    unsigned long long* trace;
    uint64_t timeout_ns;
    int force_mb;
    int force_stages;
    int force_grid;
    int force_streamk;
    int ablate;    // perf ablation bits (tests only): 1 skip MMA issue, 2 skip dequant math
    int variant;   // -1 auto; general kernel: 0 LARGE (1 CTA/SM), 1 SMALL (2 CTAs/SM); decode kernel: 2 also 5 <= M <= 16
    int l2_prefetch;   // decode kernel: L2 prefetch distance in stages, -1 = engine's choice
    const flute_b200_tp* tp;   // tensor-parallel fused exchange (decode kernel only), or nullptr
};

// Kernel parameters (passed by value).
struct QgemmParams {
    const uint16_t* S;        // [N, G] T
    const uint32_t* table2;   // [4^bits] packed T pairs
    uint16_t* D;              // [M, N] T
    uint8_t* workspace;       // [tile counters | fp32 partials]
    Diag* diag;
    uint32_t* dbg;
    unsigned long long* trace;   // optional [grid][8] globaltimer stamps
    uint64_t timeout_ns;
    int M, N, K, G;
    int group_size, tile_p;
    int group_shift;          // log2(group_size / 64)
    int mb;                   // activation rows per tile == MMA N
    int n_tiles, m_tiles, k_iters;
    int stages, nchunk, streamk;
    uint32_t partial_offset;
    uint32_t stage_bytes, w_bytes, b_bytes;
    int ablate;
    int static_weights;       // weights may be prefetched before griddepcontrol.wait
    uint32_t neg_zero2;       // packed (-0, -0): addend that keeps the scale multiply an FMA-pipe instruction
    uint32_t plane1_row0;     // 3-bit: first row of planes 1/2 (N/16)
};

// "Function attribute already set on this device" bits, safe against concurrent first calls from several host
// threads (setting an attribute twice is harmless; a torn bool array is not).
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask{0};
    bool done(int device) const { return device >= 0 && device < 64 && ((mask.load(std::memory_order_acquire) >> device) & 1ull); }
    void mark(int device) { if (device >= 0 && device < 64) mask.fetch_or(1ull << device, std::memory_order_release); }
};

int qgemm_launch(const QgemmArgs& a, cudaStream_t stream);
// decode-shaped kernel (M <= 16, 2/4-bit): qgemm_decode_sm100.cu
bool qgemm_decode_supported(const QgemmArgs& a);
int qgemm_decode_launch(const QgemmArgs& a, cudaStream_t stream);
// prefill-shaped kernel (M > 16, 4-bit): qgemm_prefill_sm100.cu
// The tail of the workspace holds its per-CTA partial-tile slots (2 x 128 KB per SM, contents undefined between
// launches); the zero-invariant fp32 accumulators of the other kernels must stay below it.
inline size_t prefill_scratch_bytes(int num_sms) { return (size_t)num_sms * 2 * 131072 + 256; }
bool qgemm_prefill_supported(const QgemmArgs& a);
int qgemm_prefill_launch(const QgemmArgs& a, cudaStream_t stream);
// CTAs per launch.  Default: every SM the launch may use, each with a contiguous share of the (tile, k) stages.  When
// a CTA's share is only a few stages the launch is latency-bound, and shares that do not straddle tiles are worth a
// few idle SMs: every CTA then has ONE partial segment (one hand-over) and a tile has exactly k_iters / share
// contributors.  Swept per shape on B200 (gpurun r02g2, profiles/r02_experiments.md; us per launch, default -> aligned):
// 4096x4096 7.05 -> 6.63 (128 CTAs x 4 stages), 3584x4096 7.16 -> 6.36 (112 x 4), 1024x4096 5.24 -> 4.86 (64 x 2);
// large shapes lose (28672x4096 on 112 CTAs x 32 stages: 17.3 -> 19.5), hence the bounds on share and grid.  The general
// kernel's decode-shaped launches (3-bit, M <= 16; gpurun r02g3) use the same rule: 28672x4096 W3 48.5 -> 41.0 (112 x 8),
// 4096x14336 W3 35.8 -> 29.2 (112 x 4).
inline int decode_grid_for(long long total, int k_iters, int max_grid, bool halve_single_stage_shares) {
    if (total <= max_grid) {
        if (!halve_single_stage_shares) return (int)total;
        // at most one stage per CTA: two per CTA halve the contributors per tile, as long as ~64 CTAs remain
        if ((k_iters % 2) == 0 && total / 2 >= 64) return (int)(total / 2);
        return (int)total;
    }
    for (int share = 2; share <= 8; ++share) {          // smallest share = largest grid first
        if (k_iters % share != 0) continue;
        const long long g = total / share;
        if (g > max_grid) continue;
        if (g * 4 >= (long long)max_grid * 3 || (share <= 4 && g * 5 >= (long long)max_grid * 3)) return (int)g;
        break;                                          // larger shares only give smaller grids
    }
    return max_grid;
}

int qgemm_max_mb(int bits);
int tp_advance_launch(unsigned* epoch, cudaStream_t stream);
int tp_publish_launch(unsigned* const* flags, int tp, cudaStream_t stream);
int tp_wait_launch(const unsigned* flag, unsigned per_step, unsigned offset, const unsigned* epoch, uint64_t timeout_ns, Diag* diag,
                   cudaStream_t stream);
const char* qgemm_dispatch_name(int M, int num_bits, bool bf16);
int make_tmap_2d(CUtensorMap* tm, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t outer,
                 uint64_t row_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle);

int make_tmap_3d(CUtensorMap* tm, CUtensorMapDataType dt, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                 CUtensorMapSwizzle swizzle);

}  // namespace fb
