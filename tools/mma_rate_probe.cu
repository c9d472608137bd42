// Standalone probe: issue rate of tcgen05.mma (kind::f16, cta_group::1) on B200 as a function of N,
// with the A operand in tensor memory (.ts form, what the qgemm kernels use) or in shared memory (.ss).
// One CTA per SM, one issuing thread, R back-to-back MMAs then one commit; prints cycles per MMA.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iflute_b200/csrc tools/mma_rate_probe.cu -o tools/mma_probe.bin
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

using namespace fb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ bool elect1() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

struct P {
    int mode;     // 0 = A in TMEM, 1 = A in smem
    int m, n;     // MMA shape (K = 16)
    int reps;
    int nacc;     // distinct accumulators cycled through
    long long* out;
};

template <int MODE>
__global__ void __launch_bounds__(128) probe(P p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    // zero B (256 rows x 128 B) and A (128 rows x 128 B) tiles
    for (int i = threadIdx.x; i < (32768 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + (base - smem_u32(smem)))[i] = 0;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tbase), 512); tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tbase;
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc_f16(true, p.m, p.n);
        const uint64_t bdesc = make_smem_desc_sw128(base);
        const uint64_t adesc = make_smem_desc_sw128(base + 32768);
        // warm-up
        for (int i = 0; i < 8; ++i) {
            if (elect1()) {
                if (MODE == 0) tc_mma_ts(tmem + 256, tmem + (i & 3) * 8, bdesc, idesc, 0);
                else tc_mma_ss(tmem + 256, adesc, bdesc, idesc, 0);
            }
            __syncwarp();
        }
        if (elect1()) tc_commit(smem_u32(&bar));
        __syncwarp();
        while (!mbar_try_wait(smem_u32(&bar), 0)) {}
        const long long t0 = clock64();
        const uint32_t dstep = (p.nacc > 1) ? (uint32_t)p.n : 0u;
        for (int i = 0; i < p.reps; i += 16) {
            if (elect1()) {
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const uint32_t d = tmem + 256 + (u & 3) * dstep;
                    if (MODE == 0) tc_mma_ts(d, tmem + (u & 3) * 8 + (u >> 2) * 32, bdesc + (u & 3) * 2, idesc, 1);
                    else tc_mma_ss(d, adesc + (u & 3) * 2, bdesc + (u & 3) * 2, idesc, 1);
                }
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        if (elect1()) tc_commit(smem_u32(&bar));
        __syncwarp();
        while (!mbar_try_wait(smem_u32(&bar), 1)) {}
        const long long t2 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) { p.out[0] = t1 - t0; p.out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    long long* out;
    CK(cudaMalloc(&out, 16));
    CK(cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    int nsm = 0;
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    for (int reps : {16, 64, 512})
    for (int mode = 0; mode < 2; ++mode) {
        for (int m : {128}) {
            for (int n : {16, 32, 64, 128, 256}) {
                for (int nacc : {1, 4}) {
                    if (n * nacc > 256) continue;
                    if (m == 64 && mode == 0) continue;   // keep to the M=128 lane layout for TMEM A
                    P p{mode, m, n, reps, nacc, out};
                    if (mode == 0) probe<0><<<nsm, 128, 50 * 1024>>>(p); else probe<1><<<nsm, 128, 50 * 1024>>>(p);
                    CK(cudaDeviceSynchronize());
                    long long h[2];
                    CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
                    printf("reps %3d: %s M=%3d N=%3d K=16 nacc=%d : issue total %6lld, complete total %6lld | issue %6.1f cyc/MMA, complete %6.1f cyc/MMA  (%.0f MAC/clk/SM, %.1f A-rows*K elems/clk)\n",
                           reps, mode == 0 ? "A=TMEM" : "A=SMEM", m, n, nacc, h[0], h[1], (double)h[0] / reps, (double)h[1] / reps,
                           (double)m * n * 16 * reps / h[1], (double)m * 16 * reps / h[1]);
                }
            }
        }
    }
    return 0;
}
