// Auxiliary sm_100a kernels of the LUT-qGEMM path:
//   * dequantize_kernel : packed weights -> dense W_hat[K, N] (replaces the identity-GEMM
//     reconstruct/unpack of flute/utils.py:347-407)
//   * hadamard_kernel   : blockwise orthonormal Walsh-Hadamard pre-transform (replaces
//     flute/csrc/hadamard_transform_cuda.cu:92-748 as used by qgemm.cpp:201-244)
// Both are HBM-bound elementwise / small-transform kernels: coalesced 16-bit traffic, smem staging.
#include "aux_kernels.h"
#include "ptx.cuh"

namespace fb {

// ----------------------------------------------------------------------------------------
// dequantize: one thread per (packed row unit u, k-pair k2); NJ columns each
// ----------------------------------------------------------------------------------------
template <int BITS, bool BF16>
__global__ void __launch_bounds__(256) dequantize_kernel(const uint32_t* __restrict__ Q32, const uint16_t* __restrict__ S,
                                                         const uint32_t* __restrict__ table2, uint16_t* __restrict__ What,
                                                         int N, int K, int group_size, int tile_p) {
    constexpr int NJ = (BITS == 4) ? 4 : (BITS == 2) ? 8 : 16;
    constexpr int LUTN = 1 << (2 * BITS);
    __shared__ uint32_t lut[LUTN];
    for (int i = threadIdx.y * 32 + threadIdx.x; i < LUTN; i += 256) lut[i] = table2[i];
    __syncthreads();

    const int K2 = K / 2;
    const int G = K / group_size;
    const int units = (BITS == 3) ? N / 16 : N / 16 * BITS;   // rows that start a word triple / a word
    const int u = blockIdx.x * 32 + threadIdx.x;
    const int k2 = blockIdx.y * 8 + threadIdx.y;
    if (u >= units || k2 >= K2) return;

    uint32_t codes[NJ];
    int n0, nstride;
    if (BITS == 3) {
        const int nb = u >> 5, t = u & 31;
        const uint32_t w0 = Q32[(size_t)u * K2 + k2];
        const uint32_t w1 = Q32[(size_t)(N / 16 + nb * 64 + t) * K2 + k2];
        const uint32_t w2 = Q32[(size_t)(N / 16 + nb * 64 + 32 + t) * K2 + k2];
#pragma unroll
        for (int j = 0; j < 15; ++j) {
            const uint32_t w = (j % 3 == 0) ? w0 : (j % 3 == 1) ? w1 : w2;
            codes[j] = (w >> (6 * (j / 3))) & 0x3fu;
        }
        codes[NJ - 1] = (w0 >> 30) | ((w1 >> 30) << 2) | ((w2 >> 30) << 4);
        n0 = nb * 512 + t;
        nstride = 32;
    } else {
        const uint32_t w = Q32[(size_t)u * K2 + k2];
#pragma unroll
        for (int j = 0; j < NJ; ++j) codes[j] = (w >> (2 * BITS * j)) & (LUTN - 1);
        n0 = (u / tile_p) * (NJ * tile_p) + (u % tile_p);
        nstride = tile_p;
    }
    const int g = (2 * k2) / group_size;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int n = n0 + j * nstride;
        uint32_t s = S[(size_t)n * G + g];
        uint32_t v = mul2<BF16>(lut[codes[j]], s | (s << 16));
        What[(size_t)(2 * k2) * N + n] = (uint16_t)(v & 0xffffu);
        What[(size_t)(2 * k2 + 1) * N + n] = (uint16_t)(v >> 16);
    }
}

int dequantize_launch(const void* Q, const void* S, const void* table2, void* What, int N, int K, int bits, int group,
                      int tile_p, int bf16, cudaStream_t stream) {
    const int units = (bits == 3) ? N / 16 : N / 16 * bits;
    dim3 grid((units + 31) / 32, (K / 2 + 7) / 8), block(32, 8);
    const uint32_t* q = static_cast<const uint32_t*>(Q);
    const uint16_t* s = static_cast<const uint16_t*>(S);
    const uint32_t* t2 = static_cast<const uint32_t*>(table2);
    uint16_t* w = static_cast<uint16_t*>(What);
#define FB_DQ(B, T) dequantize_kernel<B, T><<<grid, block, 0, stream>>>(q, s, t2, w, N, K, group, tile_p)
    switch (bits * 2 + (bf16 ? 1 : 0)) {
        case 8: FB_DQ(4, false); break;
        case 9: FB_DQ(4, true); break;
        case 4: FB_DQ(2, false); break;
        case 5: FB_DQ(2, true); break;
        case 6: FB_DQ(3, false); break;
        case 7: FB_DQ(3, true); break;
        default: return FB_ERR_BITS;
    }
#undef FB_DQ
    if (cudaGetLastError() != cudaSuccess) return FB_ERR_LAUNCH;
    return FB_OK;
}

// ----------------------------------------------------------------------------------------
// Hadamard: each block transforms `rows_per_block` rows of length h held in smem as fp32.
// Butterflies in Sylvester order, one rounding to T at the end (the reference rounds between
// its tensor-core passes; it has no test pinning that).
// ----------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(256) hadamard_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                                                       long rows, int h, int log_h, int rows_per_block, float scale) {
    extern __shared__ float buf[];
    // Launched with programmatic stream serialisation: it may start while the producer of `in` is still running, so it
    // waits for that first -- and releases its own dependents at once: a qgemm that follows (flute.qgemm_hadamard) is
    // then resident, with its weights streaming and its TMEM slots filling, while these few CTAs transform the row.
    pdl_wait_prior_grids();
    pdl_launch_dependents();
    const long row0 = (long)blockIdx.x * rows_per_block;
    const int nrows = (int)min((long)rows_per_block, rows - row0);
    if (nrows <= 0) return;
    const int total = nrows * h;
    const uint16_t* src = in + row0 * h;
    uint16_t* dst = out + row0 * h;
    for (int i = threadIdx.x; i < total; i += blockDim.x) buf[i] = t_to_f32<BF16>(src[i]);
    __syncthreads();
    const int half = total / 2;
    for (int s = 0; s < log_h; ++s) {
        const int stride = 1 << s;
        for (int b = threadIdx.x; b < half; b += blockDim.x) {
            // butterfly b -> element index with bit s cleared
            const int lo = ((b >> s) << (s + 1)) | (b & (stride - 1));
            const float x = buf[lo], y = buf[lo + stride];
            buf[lo] = x + y;
            buf[lo + stride] = x - y;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < total; i += blockDim.x) dst[i] = f32_to_t<BF16>(buf[i] * scale);
}

int hadamard_launch(const void* in, void* out, long rows, int h, int bf16, cudaStream_t stream) {
    if (h <= 0 || (h & (h - 1)) || h > 32768) return FB_ERR_HADAMARD;
    if (rows <= 0) return FB_OK;
    int log_h = 0;
    while ((1 << log_h) < h) ++log_h;
    int rpb = 2048 / h;   // small transforms: several rows per block
    if (rpb < 1) rpb = 1;
    const size_t smem = (size_t)rpb * h * sizeof(float);
    const long blocks = (rows + rpb - 1) / rpb;
    const float scale = 1.0f / sqrtf((float)h);
    const uint16_t* i16 = static_cast<const uint16_t*>(in);
    uint16_t* o16 = static_cast<uint16_t*>(out);
    auto kern = bf16 ? hadamard_kernel<true> : hadamard_kernel<false>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, i16, o16, rows, h, log_h, rpb, scale) != cudaSuccess) {
        cudaGetLastError();
        return FB_ERR_LAUNCH;
    }
    if (cudaGetLastError() != cudaSuccess) return FB_ERR_LAUNCH;
    return FB_OK;
}

}  // namespace fb
