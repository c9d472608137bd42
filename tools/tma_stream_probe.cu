// Standalone probe: how fast can a persistent TMA pipeline stream a row-major int16 [P, K] matrix
// from HBM on B200 as a function of the box shape (bytes contiguous per row visit)?
// No compute: one producer thread issues TMA loads into a ring, one consumer thread frees slots.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/tma_stream_probe.cu -o /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(par) : "memory");
    }
}
__device__ __forceinline__ void tma2d(uint32_t dst, const void* tm, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma3d(uint32_t dst, const void* tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"((uint64_t)src), "r"(bytes), "r"(bar) : "memory");
}

struct Params {
    int mode;          // 0: 2D box {64, rows}; 1: 3D box {64, rows, chunks}; 2: 1D bulk contiguous
    int rows;          // rows per load
    int chunks;        // 64-element k-chunks per load (mode 1)
    int stages;
    int load_bytes;
    int K, P;
    int loads_per_cta; // contiguous range of loads per CTA
    int total_loads;
    int batch;         // issue this many loads back to back (waits for all their slots first)
    int extra_a;       // also load a {128B x 16 rows} activation box per load (only 1 row in bounds)
    int extra_s;       // also load two {16B x 256 rows} scale boxes every 8 loads
    const uint8_t* base;
};

__global__ void __launch_bounds__(64) probe(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tms, Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[16], empty[16];
    const uint32_t ring = (smem_u32(smem) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int l0 = blockIdx.x * p.loads_per_cta;
    const int l1 = min(p.total_loads, l0 + p.loads_per_cta);
    const int kchunks = p.K / 64;                       // 64-element chunks per row
    const int per_row_block = (p.mode == 1) ? kchunks / p.chunks : kchunks;   // loads per row-block
    if (threadIdx.x == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int l = l0; l < l1; l += p.batch) {
            const int nb = min(p.batch, l1 - l);
            int st = stage; uint32_t ph = phase;
            for (int b = 0; b < nb; ++b) { mbar_wait(smem_u32(&empty[st]), ph ^ 1u); if (++st == p.stages) { st = 0; ph ^= 1u; } }
            for (int b = 0; b < nb; ++b) {
                const int ll = l + b;
                const uint32_t dst = ring + stage * p.load_bytes;
                const uint32_t bar = smem_u32(&full[stage]);
                mbar_expect(bar, p.load_bytes + (p.extra_a ? 2048 : 0) + ((p.extra_s && (ll & 7) == 0) ? 8192 : 0));
                if (p.extra_a) tma2d(ring + p.stages * p.load_bytes, &tma, bar, (ll % kchunks) * 64, 0);
                if (p.extra_s && (ll & 7) == 0) {
                    tma2d(ring + p.stages * p.load_bytes + 2048, &tms, bar, ((ll / 8) % 8) * 8, (ll / kchunks) * 512);
                    tma2d(ring + p.stages * p.load_bytes + 2048 + 4096, &tms, bar, ((ll / 8) % 8) * 8, (ll / kchunks) * 512 + 256);
                }
                if (p.mode == 2) {
                    bulk1d(dst, p.base + (size_t)ll * p.load_bytes, p.load_bytes, bar);
                } else {
                    const int rb = ll / per_row_block, kc = ll - rb * per_row_block;   // k fastest
                    if (p.mode == 0) tma2d(dst, &tm, bar, kc * 64, rb * p.rows);
                    else tma3d(dst, &tm, bar, 0, rb * p.rows, kc * p.chunks);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0;
        for (int l = l0; l < l1; ++l) {
            mbar_wait(smem_u32(&full[stage]), phase);
            mbar_arrive(smem_u32(&empty[stage]));
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
    }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int K = 4096;
    const long P = 7168 * 4;           // 28672 rows x 8 KB = 235 MB  (> L2)
    const size_t bytes = (size_t)P * K * 2;
    uint8_t* d;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMemset(d, 1, bytes));
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    PFN_enc enc = (PFN_enc)fp;
    int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

    struct Case { const char* name; int mode, rows, chunks, stages, batch, ctas_per_sm, extra_a, extra_s; };
    // activation-like [16, K] and scale-like [P*4, 64] side matrices
    uint8_t *da, *ds;
    CK(cudaMalloc(&da, 16 * K * 2)); CK(cudaMemset(da, 1, 16 * K * 2));
    CK(cudaMalloc(&ds, (size_t)P * 4 * 64 * 2)); CK(cudaMemset(ds, 1, (size_t)P * 4 * 64 * 2));
    CUtensorMap tma_, tms_;
    {
        cuuint64_t dims[2] = {(cuuint64_t)K, 1}; cuuint64_t str[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, 16}; cuuint32_t es[2] = {1, 1};
        if (enc(&tma_, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, da, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) printf("enc a fail\n");
        cuuint64_t dims2[2] = {64, (cuuint64_t)P * 4}; cuuint64_t str2[1] = {128};
        cuuint32_t box2[2] = {8, 256};
        if (enc(&tms_, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, ds, dims2, str2, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) printf("enc s fail\n");
    }
    std::vector<Case> cases = {
        {"2D {128B x 128 rows}  16KB x8 + A box  ", 0, 128, 1, 8, 1, 1, 1, 0},
        {"2D {128B x 128 rows}  16KB x8 + A + S  ", 0, 128, 1, 8, 1, 1, 1, 1},
        {"2D {128B x 128 rows}  16KB x8 + S      ", 0, 128, 1, 8, 1, 1, 0, 1},
        {"2D {128B x 32 rows}    4KB x32         ", 0, 32, 1, 32, 1, 1, 0, 0},
        {"2D {128B x 64 rows}    8KB x8          ", 0, 64, 1, 8, 1, 1, 0, 0},
        {"2D {128B x 128 rows}  16KB x8 (current)", 0, 128, 1, 8, 1, 1},
        {"2D {128B x 128 rows}  16KB x8 batch4   ", 0, 128, 1, 8, 4, 1},
        {"2D {128B x 128 rows}  16KB x12         ", 0, 128, 1, 12, 1, 1},
        {"2D {128B x 256 rows}  32KB x6          ", 0, 256, 1, 6, 1, 1},
        {"2D {128B x 64 rows}    8KB x16         ", 0, 64, 1, 16, 1, 1},
        {"3D {128B x 64r x 2c}  16KB x8  (256B)  ", 1, 64, 2, 8, 1, 1},
        {"3D {128B x 32r x 4c}  16KB x8  (512B)  ", 1, 32, 4, 8, 1, 1},
        {"3D {128B x 16r x 8c}  16KB x8  (1KB)   ", 1, 16, 8, 8, 1, 1},
        {"3D {128B x 8r x 16c}  16KB x8  (2KB)   ", 1, 8, 16, 8, 1, 1},
        {"3D {128B x 32r x 8c}  32KB x6  (1KB)   ", 1, 32, 8, 6, 1, 1},
        {"3D {128B x 128r x 2c} 32KB x6  (256B)  ", 1, 128, 2, 6, 1, 1},
        {"1D bulk contiguous    16KB x8          ", 2, 0, 0, 8, 1, 1},
        {"1D bulk contiguous    16KB x12         ", 2, 0, 0, 12, 1, 1},
        {"2D {128B x 128 rows}  16KB x4, 2 CTA/SM", 0, 128, 1, 4, 1, 2},
        {"3D {128B x 32r x 4c}  16KB x4, 2 CTA/SM", 1, 32, 4, 4, 1, 2},
        {"3D {128B x 32r x 4c}  16KB x3, 2 CTA/SM", 1, 32, 4, 3, 1, 2},
    };
    for (auto& c : cases) {
        CUtensorMap tm;
        Params p{};
        p.mode = c.mode; p.rows = c.rows; p.chunks = c.chunks; p.stages = c.stages; p.K = K; p.P = (int)P; p.batch = c.batch; p.extra_a = c.extra_a; p.extra_s = c.extra_s;
        p.base = d;
        if (c.mode == 0) {
            cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)P}; cuuint64_t str[1] = {(cuuint64_t)K * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)c.rows}; cuuint32_t es[2] = {1, 1};
            if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc fail\n"); continue; }
            p.load_bytes = 128 * c.rows;
            p.total_loads = (int)(P / c.rows) * (K / 64);
        } else if (c.mode == 1) {
            // dims {64 elems, P rows, K/64 chunks}; strides {K*2 bytes, 128 bytes}
            cuuint64_t dims[3] = {64, (cuuint64_t)P, (cuuint64_t)(K / 64)}; cuuint64_t str[2] = {(cuuint64_t)K * 2, 128};
            cuuint32_t box[3] = {64, (cuuint32_t)c.rows, (cuuint32_t)c.chunks}; cuuint32_t es[3] = {1, 1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r) { printf("%s: enc fail %d\n", c.name, (int)r); continue; }
            p.load_bytes = 128 * c.rows * c.chunks;
            p.total_loads = (int)(P / c.rows) * (K / 64 / c.chunks);
        } else {
            memset(&tm, 0, sizeof(tm));
            p.load_bytes = 16384;
            p.total_loads = (int)(bytes / 16384);
        }
        const int grid = nsm * c.ctas_per_sm;
        p.loads_per_cta = (p.total_loads + grid - 1) / grid;
        const size_t smem = (size_t)c.stages * p.load_bytes + 1024 + 2048 + 8192;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int w = 0; w < 2; ++w) probe<<<grid, 64, smem>>>(tm, tma_, tms_, p);
        CK(cudaDeviceSynchronize());
        float best = 1e9;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0);
            probe<<<grid, 64, smem>>>(tm, tma_, tms_, p);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("%s : %7.1f us  %7.1f GB/s\n", c.name, best * 1e3, bytes / (best * 1e-3) / 1e9);
        fflush(stdout);
    }
    return 0;
}
