/*
 * CPU oracle for the FLUTE LUT-quantized GEMM hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of what the reference computes on the `flute.qgemm` path, for
 * cases too large for the numpy oracle (oracle/flute_oracle.py) to finish in seconds,
 * and as the threaded CPU baseline timed by bench.py.  Nothing under flute_b200/ links
 * or loads this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.  Pinned against tests/golden/wire_format.npz (vectors
 * produced by the reference's own packers) by tests/test_oracle_golden.py.
 *
 * Citations are into /root/reference.
 *   wire format ......... flute/utils.py:59-253, flute/packbits_utils.py:84-140,191-220
 *   in-register unpack .. flute/csrc/packbits_utils.hpp:82-142 (2/4-bit), :322-363 (3-bit)
 *   dequant numerics .... flute/csrc/packbits_utils.hpp:105,139,343-361 (one __hmul2 in T)
 *   ground-truth GEMM ... tests/kernel.py:68-71, flute/tune.py:332-335
 *   Hadamard ............ flute/csrc/qgemm.cpp:201-211 (PARITY UNPINNED: no reference test)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

/* ---- minimal pthread parallel-for (the image has no libgomp) ---------------------- */
typedef void (*range_fn)(long begin, long end, void* ctx);
typedef struct { range_fn fn; void* ctx; long begin, end; } task_t;
static void* task_main(void* p) { task_t* t = (task_t*)p; t->fn(t->begin, t->end, t->ctx); return NULL; }

int oracle_num_threads(void) {
    const char* e = getenv("ORACLE_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}

static void parallel_for(long n, range_fn fn, void* ctx) {
    int nt = oracle_num_threads();
    if (nt > n) nt = (int)(n > 0 ? n : 1);
    if (nt <= 1) { fn(0, n, ctx); return; }
    pthread_t th[256];
    task_t tk[256];
    long chunk = (n + nt - 1) / nt;
    int started = 0;
    for (int i = 0; i < nt; ++i) {
        long b = i * chunk, e = b + chunk > n ? n : b + chunk;
        if (b >= e) break;
        tk[i].fn = fn; tk[i].ctx = ctx; tk[i].begin = b; tk[i].end = e;
        if (pthread_create(&th[i], NULL, task_main, &tk[i]) != 0) { fn(b, e, ctx); th[i] = 0; tk[i].fn = NULL; }
        started = i + 1;
    }
    for (int i = 0; i < started; ++i) if (tk[i].fn) pthread_join(th[i], NULL);
}

/* ---- T <-> float ---------------------------------------------------------------- */
static inline float bf16_to_f32(uint16_t b) {
    uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline uint16_t f32_to_bf16(float f) { /* round-to-nearest-even */
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float f16_to_f32(uint16_t h) {
    _Float16 x;
    memcpy(&x, &h, 2);
    return (float)x;
}
static inline uint16_t f32_to_f16(float f) {
    _Float16 x = (_Float16)f; /* IEEE RN */
    uint16_t h;
    memcpy(&h, &x, 2);
    return h;
}
static inline float t_to_f32(uint16_t v, int is_bf16) { return is_bf16 ? bf16_to_f32(v) : f16_to_f32(v); }
static inline uint16_t f32_to_t(float v, int is_bf16) { return is_bf16 ? f32_to_bf16(v) : f32_to_f16(v); }

/* ---- wire format ---------------------------------------------------------------- */
/* Pair code presented to table2 for (k2, n): (q[2k2,n] << bits) | q[2k2+1,n]. */
static inline uint32_t pair_code(const uint32_t* Q32, int N, int K2, int bits, int tile_p, int k2, int n) {
    if (bits == 3) {
        int nb = n / 512, j = (n % 512) / 32, t = n % 32;
        const uint32_t* r0 = Q32 + (size_t)(nb * 32 + t) * K2;
        const uint32_t* r1 = Q32 + (size_t)(N / 16 + nb * 64 + t) * K2;
        const uint32_t* r2 = Q32 + (size_t)(N / 16 + nb * 64 + 32 + t) * K2;
        uint32_t w0 = r0[k2], w1 = r1[k2], w2 = r2[k2];
        if (j < 15) {
            uint32_t w = (j % 3 == 0) ? w0 : (j % 3 == 1) ? w1 : w2;
            return (w >> (6 * (j / 3))) & 0x3fu;
        }
        return ((w0 >> 30) & 3u) | (((w1 >> 30) & 3u) << 2) | (((w2 >> 30) & 3u) << 4);
    }
    int fields = 32 / (2 * bits);
    int blk = fields * tile_p;
    int b = n / blk, j = (n % blk) / tile_p, t = n % tile_p;
    uint32_t w = Q32[(size_t)(b * tile_p + t) * K2 + k2];
    return (w >> (2 * bits * j)) & ((1u << (2 * bits)) - 1u);
}

/* int16 [P, K] -> uint8 [K, N] */
typedef struct { const uint32_t* Q32; int N, K2, bits, tile_p; uint8_t* W; } unpack_ctx;
static void unpack_range(long b, long e, void* p) {
    unpack_ctx* c = (unpack_ctx*)p;
    uint32_t mask = (1u << c->bits) - 1u;
    for (long k2 = b; k2 < e; ++k2)
        for (int n = 0; n < c->N; ++n) {
            uint32_t code = pair_code(c->Q32, c->N, c->K2, c->bits, c->tile_p, (int)k2, n);
            c->W[(size_t)(2 * k2) * c->N + n] = (uint8_t)((code >> c->bits) & mask);
            c->W[(size_t)(2 * k2 + 1) * c->N + n] = (uint8_t)(code & mask);
        }
}
int oracle_unpack(const int16_t* Q, int N, int K, int bits, int tile_p, uint8_t* W) {
    if (bits < 2 || bits > 4 || K % 2) return -1;
    unpack_ctx c = {(const uint32_t*)Q, N, K / 2, bits, tile_p, W};
    parallel_for(K / 2, unpack_range, &c);
    return 0;
}

/* uint8 [K, N] -> int16 [P, K] (zero-initialised by this function) */
int oracle_pack(const uint8_t* W, int N, int K, int bits, int tile_p, int16_t* Q) {
    if (bits < 2 || bits > 4 || K % 2) return -1;
    if (bits == 3 && tile_p != 32) return -2;
    int blkc = (bits == 3) ? 512 : (32 / (2 * bits)) * tile_p;
    if (N % blkc) return -3;
    uint32_t* Q32 = (uint32_t*)Q;
    int K2 = K / 2;
    size_t P = (size_t)N / 16 * bits;
    memset(Q32, 0, P * K2 * sizeof(uint32_t));
    for (int k2 = 0; k2 < K2; ++k2)
        for (int n = 0; n < N; ++n) {
            uint32_t c = ((uint32_t)W[(size_t)(2 * k2) * N + n] << bits) | W[(size_t)(2 * k2 + 1) * N + n];
            if (bits == 3) {
                int nb = n / 512, j = (n % 512) / 32, t = n % 32;
                size_t r[3] = {(size_t)nb * 32 + t, (size_t)N / 16 + nb * 64 + t, (size_t)N / 16 + nb * 64 + 32 + t};
                if (j < 15) {
                    Q32[r[j % 3] * K2 + k2] |= c << (6 * (j / 3));
                } else {
                    for (int w = 0; w < 3; ++w) Q32[r[w] * K2 + k2] |= ((c >> (2 * w)) & 3u) << 30;
                }
            } else {
                int fields = 32 / (2 * bits);
                int blk = fields * tile_p;
                int b = n / blk, j = (n % blk) / tile_p, t = n % tile_p;
                Q32[((size_t)b * tile_p + t) * K2 + k2] |= c << (2 * bits * j);
            }
        }
    return 0;
}

/* ---- dequantisation: W_hat[k, n] = round_T(table2[code].{lo,hi} * S[n, k / group]) ---- */
typedef struct {
    const uint32_t* Q32; const uint16_t* S; const uint32_t* table2;
    int N, K2, G, bits, group, tile_p, is_bf16; uint16_t* What;
} deq_ctx;
static void deq_range(long b, long e, void* p) {
    deq_ctx* c = (deq_ctx*)p;
    int bf = c->is_bf16;
    for (long k2 = b; k2 < e; ++k2)
        for (int n = 0; n < c->N; ++n) {
            uint32_t code = pair_code(c->Q32, c->N, c->K2, c->bits, c->tile_p, (int)k2, n);
            uint32_t ent = c->table2[code];
            float s = t_to_f32(c->S[(size_t)n * c->G + (2 * k2) / c->group], bf);
            /* the fp32 product of two T values is exact; one rounding to T == __hmul2 */
            c->What[(size_t)(2 * k2) * c->N + n] = f32_to_t(t_to_f32((uint16_t)(ent & 0xffffu), bf) * s, bf);
            c->What[(size_t)(2 * k2 + 1) * c->N + n] = f32_to_t(t_to_f32((uint16_t)(ent >> 16), bf) * s, bf);
        }
}
int oracle_dequantize(const int16_t* Q, const uint16_t* S, const uint32_t* table2, int N, int K, int bits,
                      int group, int tile_p, int is_bf16, uint16_t* What /* [K, N] */) {
    if (bits < 2 || bits > 4 || K % 2 || K % group) return -1;
    deq_ctx c = {(const uint32_t*)Q, S, table2, N, K / 2, K / group, bits, group, tile_p, is_bf16, What};
    parallel_for(K / 2, deq_range, &c);
    return 0;
}

/* ---- D = round_T(A @ W_hat), fp64 accumulation ------------------------------------ */
typedef struct {
    const float* Af; const uint32_t* Q32; const uint16_t* S; const uint32_t* table2;
    int M, N, K, K2, G, bits, group, tile_p, is_bf16; uint16_t* D;
} gemm_ctx;
static void gemm_range(long b, long e, void* p) {
    gemm_ctx* c = (gemm_ctx*)p;
    int bf = c->is_bf16;
    double* acc = (double*)malloc((size_t)c->M * sizeof(double));
    for (long n = b; n < e; ++n) {
        for (int m = 0; m < c->M; ++m) acc[m] = 0.0;
        for (int k2 = 0; k2 < c->K2; ++k2) {
            uint32_t code = pair_code(c->Q32, c->N, c->K2, c->bits, c->tile_p, k2, (int)n);
            uint32_t ent = c->table2[code];
            float s = t_to_f32(c->S[(size_t)n * c->G + (2 * k2) / c->group], bf);
            float w0 = t_to_f32(f32_to_t(t_to_f32((uint16_t)(ent & 0xffffu), bf) * s, bf), bf);
            float w1 = t_to_f32(f32_to_t(t_to_f32((uint16_t)(ent >> 16), bf) * s, bf), bf);
            for (int m = 0; m < c->M; ++m) {
                const float* a = c->Af + (size_t)m * c->K + 2 * k2;
                acc[m] += (double)a[0] * w0 + (double)a[1] * w1;
            }
        }
        for (int m = 0; m < c->M; ++m) c->D[(size_t)m * c->N + n] = f32_to_t((float)acc[m], bf);
    }
    free(acc);
}
int oracle_qgemm(const uint16_t* A, const int16_t* Q, const uint16_t* S, const uint32_t* table2, int M, int N,
                 int K, int bits, int group, int tile_p, int is_bf16, uint16_t* D /* [M, N] */) {
    if (bits < 2 || bits > 4 || K % 2 || K % group) return -1;
    float* Af = (float*)malloc((size_t)M * K * sizeof(float));
    if (!Af) return -4;
    for (size_t i = 0; i < (size_t)M * K; ++i) Af[i] = t_to_f32(A[i], is_bf16);
    gemm_ctx c = {Af, (const uint32_t*)Q, S, table2, M, N, K, K / 2, K / group, bits, group, tile_p, is_bf16, D};
    parallel_for(N, gemm_range, &c);
    free(Af);
    return 0;
}

/* ---- the reference's CPU path as its tests write it (tests/kernel.py:68-71): materialise
 *      W_hat = table[W] * S in T, then a dense T x T GEMM with fp32 accumulation (what
 *      torch.mm does for half/bf16 on CPU).  Timed by bench.py as the "port" CPU baseline. */
typedef struct {
    const uint16_t* A; const uint8_t* W; const uint16_t* S; const uint16_t* table;
    int M, N, K, G, group, is_bf16; uint16_t* What; uint16_t* D; float* Wf;
} dtm_ctx;
static void dtm_deq_range(long b, long e, void* p) {
    dtm_ctx* c = (dtm_ctx*)p;
    int bf = c->is_bf16;
    for (long k = b; k < e; ++k)
        for (int n = 0; n < c->N; ++n) {
            uint16_t w = f32_to_t(t_to_f32(c->table[c->W[(size_t)k * c->N + n]], bf) *
                                  t_to_f32(c->S[(size_t)n * c->G + k / c->group], bf), bf);
            c->What[(size_t)k * c->N + n] = w;
        }
}
static void dtm_mm_range(long b, long e, void* p) {
    dtm_ctx* c = (dtm_ctx*)p;
    int bf = c->is_bf16;
    float* acc = (float*)malloc((size_t)(e - b) * sizeof(float));
    for (int m = 0; m < c->M; ++m) {
        for (long n = b; n < e; ++n) acc[n - b] = 0.f;
        for (int k = 0; k < c->K; ++k) {
            float a = t_to_f32(c->A[(size_t)m * c->K + k], bf);
            const uint16_t* wrow = c->What + (size_t)k * c->N;
            for (long n = b; n < e; ++n) acc[n - b] += a * t_to_f32(wrow[n], bf);
        }
        for (long n = b; n < e; ++n) c->D[(size_t)m * c->N + n] = f32_to_t(acc[n - b], bf);
    }
    free(acc);
}
int oracle_dequant_then_matmul(const uint16_t* A, const uint8_t* W, const uint16_t* S, const uint16_t* table,
                               int M, int N, int K, int group, int is_bf16, uint16_t* What /* scratch [K,N] */,
                               uint16_t* D) {
    if (K % group) return -1;
    dtm_ctx c = {A, W, S, table, M, N, K, K / group, group, is_bf16, What, D, NULL};
    parallel_for(K, dtm_deq_range, &c);
    parallel_for(N, dtm_mm_range, &c);
    return 0;
}

/* ---- Hadamard: rows of length h, x @ H_h / sqrt(h), Sylvester order, fp64 inside ---- */
typedef struct { const uint16_t* X; uint16_t* Y; int h, is_bf16; } had_ctx;
static void had_range(long b, long e, void* p) {
    had_ctx* c = (had_ctx*)p;
    int h = c->h, bf = c->is_bf16;
    double inv = 1.0 / sqrt((double)h);
    double* v = (double*)malloc((size_t)h * sizeof(double));
    for (long r = b; r < e; ++r) {
        for (int i = 0; i < h; ++i) v[i] = t_to_f32(c->X[(size_t)r * h + i], bf);
        for (int s = 1; s < h; s <<= 1)
            for (int blk = 0; blk < h; blk += 2 * s)
                for (int i = blk; i < blk + s; ++i) {
                    double a = v[i], d = v[i + s];
                    v[i] = a + d;
                    v[i + s] = a - d;
                }
        for (int i = 0; i < h; ++i) c->Y[(size_t)r * h + i] = f32_to_t((float)(v[i] * inv), bf);
    }
    free(v);
}
int oracle_hadamard(const uint16_t* X, long rows, int h, int is_bf16, uint16_t* Y) {
    if (h <= 0 || (h & (h - 1)) || h > (1 << 15)) return -1;
    had_ctx c = {X, Y, h, is_bf16};
    parallel_for(rows, had_range, &c);
    return 0;
}
