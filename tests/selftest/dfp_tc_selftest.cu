// TEST INFRASTRUCTURE (built as tests/selftest/libdfp_selftest.so, not part of libdfprior_b200.so or its public header).
// Self-test of the tcgen05 building blocks (dfp_tc.cuh): D[128 x N] = A[128 x K] * W[N x K]^T with the A operand
// written into TMEM by its owner threads, W in the canonical shared-memory layout, the 3xTF32 split, and D read
// back from TMEM.  tests/test_gpu_tc.py compares it with an f64 matmul.
#include "dfp_device.cuh"
#include "dfp_tc.cuh"
#include <cuda_bf16.h>
#include <stdarg.h>
#include <stdio.h>

namespace dfp {

// the plumbing the shared internal header expects from its host library (the product's lives in dfp_api.cu)
static thread_local char g_selftest_err[512] = "";
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_selftest_err, sizeof(g_selftest_err), fmt, ap);
    va_end(ap);
    return code;
}
void timing_begin(const char*, cudaStream_t) {}
void timing_end(cudaStream_t) {}

__global__ void __launch_bounds__(128) k_tc_selftest(const float* __restrict__ A, const float* __restrict__ W,
                                                     float* __restrict__ D, int K, int N, int split) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Bhi = reinterpret_cast<float*>(smem_raw);
    float* Blo = Bhi + N * K;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int idx = tid; idx < N * K; idx += 128) {
        const int n = idx / K, k = idx % K;
        const float w = W[idx];
        const float hi = tc::tf32_rna(w);
        Bhi[tc::b_index(n, k, N)] = hi;
        Blo[tc::b_index(n, k, N)] = tc::tf32_rna(w - hi);
    }
    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();                  // weights written through the generic proxy, read by the tensor core
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t col_hi = 0, col_lo = (uint32_t)K, col_d = 2u * (uint32_t)K;

    for (int c0 = 0; c0 < K; c0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            tc::split_tf32(A[tid * K + c0 + i], hi[i], lo[i]);
            if (!split) lo[i] = 0u;
        }
        tc::tmem_st8(lane_base + col_hi + c0, hi);
        tc::tmem_st8(lane_base + col_lo + c0, lo);
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after_sync();
        if (split) {
            tc::mma_3xtf32(tbase + col_d, tbase + col_hi, tbase + col_lo, tc::smem_u32(Bhi), tc::smem_u32(Blo), K, N, false);
        } else {
            const uint32_t idesc = tc::idesc_tf32_m128(N);
            const uint32_t lbo = (uint32_t)N * 16u;
            for (int ks = 0; ks < K / 8; ++ks)
                tc::mma_tf32_ts(tbase + col_d, tbase + col_hi + ks * 8, tc::smem_desc_kmajor(tc::smem_u32(Bhi) + ks * 2 * lbo, lbo, 128u),
                                idesc, ks > 0);
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tc::tmem_ld8(lane_base + col_d + c0, v);
        tc::wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) D[tid * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// mode 1: D[128 x KIN] = A[128 x J] * Wf[J x KIN] using the transposed view of the canonical Wf buffer (dH pattern)
__global__ void __launch_bounds__(128) k_tc_selftest_bT(const float* __restrict__ A, const float* __restrict__ Wf,
                                                        float* __restrict__ D, int J, int KIN) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Bhi = reinterpret_cast<float*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
    float* Blo = Bhi + J * ((KIN + 31) / 32) * 32;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int idx = tid; idx < J * KIN; idx += 128) {          // natural [j][kin] rows in 32-column swizzled tiles
        const int j = idx / KIN, k = idx % KIN;
        const float w = Wf[idx];
        const float hi = tc::tf32_rna(w);
        const int o = (k >> 5) * (J * 32) + tc::mn_tile_index(j, k & 31);
        Bhi[o] = hi;
        Blo[o] = tc::tf32_rna(w - hi);
    }
    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t col_hi = 0, col_lo = (uint32_t)J, col_d = 2u * (uint32_t)J;
    for (int c0 = 0; c0 < J; c0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) tc::split_tf32(A[tid * J + c0 + i], hi[i], lo[i]);
        tc::tmem_st8(lane_base + col_hi + c0, hi);
        tc::tmem_st8(lane_base + col_lo + c0, lo);
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after_sync();
        tc::mma_3xtf32_bT(tbase + col_d, tbase + col_hi, tbase + col_lo, tc::smem_u32(Bhi), tc::smem_u32(Blo), J, KIN, (uint32_t)J * 128u, false);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < KIN; c0 += 8) {
        uint32_t v[8];
        tc::tmem_ld8(lane_base + col_d + c0, v);
        tc::wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) D[tid * KIN + c0 + i] = __uint_as_float(v[i]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// mode 2: D[128 x N] (rows < F valid) = H[128 x F]^T * Dl[128 x N], both operands sample-major tiles in smem (dW pattern)
__global__ void __launch_bounds__(128) k_tc_selftest_outer(const float* __restrict__ H, const float* __restrict__ Dl,
                                                           float* __restrict__ D, int F, int N) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sm = reinterpret_cast<float*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
    // [128][32] row-per-sample swizzled tiles (16 KB each): H hi, H lo, then delta hi (one tile per 32 columns)
    float* Hhi = sm;
    float* Hlo = sm + 4096;
    float* Dhi = sm + 2 * 4096;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int f = 0; f < 32; ++f) {
        const float v = f < F ? H[tid * F + f] : 0.f;
        const float hi = tc::tf32_rna(v);
        Hhi[tc::mn_tile_index(tid, f)] = hi;
        Hlo[tc::mn_tile_index(tid, f)] = tc::tf32_rna(v - hi);
    }
    for (int j = 0; j < N; ++j) Dhi[(j >> 5) * 4096 + tc::mn_tile_index(tid, j & 31)] = tc::tf32_rna(Dl[tid * N + j]);
    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    if (tid == 0) {
        tc::mma_outer(tbase, tc::smem_u32(Hhi), tc::smem_u32(Hlo), 16384u, tc::smem_u32(Dhi), 16384u, N, false);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tc::tmem_ld8(lane_base + c0, v);
        tc::wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) D[tid * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// Layout probe: shared memory word i holds float(i); A (TMEM, 8 K columns) is one-hot at column k_hot for every row,
// so after ONE tcgen05.mma (M=128, N, K=8) D[s][n] is the index of the shared-memory word the tensor core used as
// B(n, k_hot) under the descriptor being probed.
__global__ void __launch_bounds__(128) k_tc_probe(float* __restrict__ D, int N, int b_mn, int lbo, int sbo, int k_hot,
                                                  int layout_type, int start_off) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // 1024-byte aligned window inside the dynamic allocation: swizzle patterns are functions of absolute address bits
    float* sm = reinterpret_cast<float*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16384; i += 128) sm[i] = (float)i;
    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (i == k_hot) ? __float_as_uint(1.0f) : 0u;
    tc::tmem_st8(lane_base, a);
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after_sync();
        const uint32_t idesc = tc::idesc_tf32_m128(N, false, b_mn != 0);
        uint64_t desc = tc::smem_desc_kmajor(tc::smem_u32(sm) + start_off, lbo, sbo);
        desc |= ((uint64_t)(layout_type & 7)) << 61;
        tc::mma_tf32_ts(tbase + 64, tbase, desc, idesc, 0u);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tc::tmem_ld8(lane_base + 64 + c0, v);
        tc::wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) D[tid * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// Probe of the accumulator layout for M = 64: SS-mode MMA with A[m][0] = m + 1 (K-major canonical), B[n][0] = 1, so that
// D[m][n] = m + 1; all 128 lanes x 16 columns are dumped to see which TMEM lane holds row m.
__global__ void __launch_bounds__(128) k_tc_probe_m64(float* __restrict__ D, int M) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sm = reinterpret_cast<float*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    float* A = sm;            // [M][8] canonical K-major: chunk ck at ck*(M*4 floats) + m*4
    float* B = sm + 2048;     // [16][8]
    for (int i = tid; i < 4096; i += 128) sm[i] = 0.f;
    __syncthreads();
    if (tid < M) A[tc::b_index(tid, 0, M)] = (float)(tid + 1);
    if (tid < 16) B[tc::b_index(tid, 0, 16)] = 1.0f;
    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 32); tc::tmem_relinquish(); }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    {   // clear the accumulator columns first so untouched lanes read back as 0
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        tc::tmem_st8(lane_base, z);
        tc::tmem_st8(lane_base + 8, z);
        tc::wait_st();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after_sync();
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | (((uint32_t)M >> 4) << 24);
        const uint64_t da = tc::smem_desc_kmajor(tc::smem_u32(A), (uint32_t)M * 16u, 128u);
        const uint64_t db = tc::smem_desc_kmajor(tc::smem_u32(B), 16u * 16u, 128u);
        tc::mma_tf32_ss(tbase, da, db, idesc, 0u);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < 16; c0 += 8) {
        uint32_t v[8];
        tc::tmem_ld8(lane_base + c0, v);
        tc::wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) D[tid * 16 + c0 + i] = __uint_as_float(v[i]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 32);
}

// bf16 MN-major operand probe: D[64 x N] = A^T * B over 128 samples with both operands row-per-sample (MN contiguous) in
// shared memory under a candidate layout `fmt` (0: no-swizzle interleaved core matrices, 1: 64-byte rows + SWIZZLE_64B,
// 2: 128-byte rows + SWIZZLE_128B); descriptor fields are passed in so one build can try several readings of LBO/SBO.
__device__ __forceinline__ uint32_t bf16_off(int fmt, int k, int mn, int MN) {
    if (fmt == 0) return (uint32_t)((k >> 3) * ((MN >> 3) * 128) + (mn >> 3) * 128 + (k & 7) * 16 + (mn & 7) * 2);
    if (fmt == 1) return (uint32_t)((mn >> 5) * 8192 + k * 64 + ((((mn & 31) >> 3) ^ ((k >> 1) & 3)) << 4) + (mn & 7) * 2);
    return (uint32_t)((mn >> 6) * 16384 + k * 128 + ((((mn & 63) >> 3) ^ (k & 7)) << 4) + (mn & 7) * 2);
}
__global__ void __launch_bounds__(128) k_tc_probe_bf16(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                       int N, int fmt, int layout_type, int lbo_a, int lbo_b, int sbo, int kstep_a, int kstep_b) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sa = sm;                 // A: 128 samples x 64 features bf16 (16 KB)
    unsigned char* sb = sm + 32768;         // B: 128 x N (<= 128) bf16
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    __syncthreads();
    for (int f = 0; f < 64; ++f) *reinterpret_cast<__nv_bfloat16*>(sa + bf16_off(fmt, tid, f, 64)) = __float2bfloat16_rn(A[tid * 64 + f]);
    for (int j = 0; j < N; ++j) *reinterpret_cast<__nv_bfloat16*>(sb + bf16_off(fmt, tid, j, N)) = __float2bfloat16_rn(B[tid * N + j]);
    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 128); tc::tmem_relinquish(); }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 128; c0 += 8) { uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; tc::tmem_st8(lane_base + c0, z); }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after_sync();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((64u >> 4) << 24);
        for (int ks = 0; ks < 8; ++ks) {
            auto desc = [&](uint32_t addr, uint32_t lbo) {
                return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(((uint32_t)sbo >> 4) & 0x3FFF) << 32) |
                       (1ull << 46) | ((uint64_t)(layout_type & 7) << 61);
            };
            const uint64_t da = desc(tc::smem_u32(sa) + ks * kstep_a, (uint32_t)lbo_a), db = desc(tc::smem_u32(sb) + ks * kstep_b, (uint32_t)lbo_b);
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                ::"r"(tbase), "l"(da), "l"(db), "r"(idesc), "r"(ks > 0 ? 1u : 0u), "r"(0), "r"(0), "r"(0), "r"(0)
                : "memory");
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tc::tmem_ld8(lane_base + c0, v);
        tc::wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) D[tid * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 128);
}
// bf16 K-major probe for the planned attention chain: D[128 x N] = A[128 x K] * W[N x K]^T with A stored in the PLANE
// layout ([8-sample group][8-feature chunk][sample][feature], dfp_planes.cuh) read as a K-major operand whose MN
// dimension is the sample (candidate: LBO = 128 between feature chunks, SBO = 16 K between sample groups) and W in the
// canonical K-major core-matrix layout ((k/8) * N*16 + n*16 + (k%8)*2 bytes; candidate LBO = N*16, SBO = 128).
// BMN = true: the transposed use of the same weight image (backward chain): D[128 x N] = A[128 x K] * Wt[K x N] where
// Wt[j][k] = weight of output j, input k, stored as the FORWARD's canonical image ((k/8) * K*16 + j*16 + (k%8)*2 bytes,
// K = number of outputs j = contraction here) and read as an MN-major B operand (candidate: LBO = 128 between 8-j
// groups, SBO = K*16 between 8-k chunks).
template <bool BMN>
__global__ void __launch_bounds__(128) k_tc_probe_bf16_k(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D,
                                                         int K, int N, int lbo_a, int sbo_a, int lbo_b, int sbo_b) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sa = sm;                 // A planes (hi only): 128 x K bf16 <= 32 KB
    unsigned char* sb = sm + 32768;         // W: N x K bf16 <= 32 KB
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    __syncthreads();
    for (int f = 0; f < K; ++f)
        *reinterpret_cast<__nv_bfloat16*>(sa + (tid >> 3) * 16 * K + (f >> 3) * 128 + (tid & 7) * 16 + (f & 7) * 2) = __float2bfloat16_rn(A[tid * K + f]);
    for (int i = tid; i < N * K; i += 128) {
        if (!BMN) {
            const int n = i / K, k = i % K;
            *reinterpret_cast<__nv_bfloat16*>(sb + (k >> 3) * N * 16 + n * 16 + (k & 7) * 2) = __float2bfloat16_rn(W[n * K + k]);
        } else {
            const int j = i / N, k = i % N;             // W given as Wt[K = J][N = KIN]
            *reinterpret_cast<__nv_bfloat16*>(sb + (k >> 3) * K * 16 + j * 16 + (k & 7) * 2) = __float2bfloat16_rn(W[j * N + k]);
        }
    }
    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 128); tc::tmem_relinquish(); }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((BMN ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        auto desc = [&](uint32_t addr, uint32_t lbo, uint32_t sbo) {
            return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
        };
        for (int ks = 0; ks < K / 16; ++ks) {                 // 16 features per MMA = two 8-feature chunks
            const uint64_t da = desc(tc::smem_u32(sa) + ks * 2 * 128, (uint32_t)lbo_a, (uint32_t)sbo_a);
            const uint64_t db = desc(tc::smem_u32(sb) + (BMN ? ks * 2 * 128 : ks * 2 * N * 16), (uint32_t)lbo_b, (uint32_t)sbo_b);
            tc::mma_bf16_ss(tbase, da, db, idesc, ks > 0 ? 1u : 0u);
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tc::tmem_ld8(lane_base + c0, v);
        tc::wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) D[tid * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 128);
}
int launch_tc_probe_bf16_k(const float* A, const float* W, float* D, int K, int N, int lbo_a, int sbo_a, int lbo_b, int sbo_b,
                           cudaStream_t st, bool bmn = false) {
    if (K < 16 || K > 128 || K % 16 || N < 16 || N > 128 || N % 16) return fail(DFP_ERR_INVALID, "bad probe shape");
    const size_t smem = 65536 + 1024;
    int rc = bmn ? set_smem(k_tc_probe_bf16_k<true>, smem) : set_smem(k_tc_probe_bf16_k<false>, smem);
    if (rc) return rc;
    if (bmn) k_tc_probe_bf16_k<true><<<1, 128, smem, st>>>(A, W, D, K, N, lbo_a, sbo_a, lbo_b, sbo_b);
    else k_tc_probe_bf16_k<false><<<1, 128, smem, st>>>(A, W, D, K, N, lbo_a, sbo_a, lbo_b, sbo_b);
    DFP_LAUNCH_CHECK("k_tc_probe_bf16_k");
    return DFP_OK;
}

int launch_tc_probe_bf16(const float* A, const float* B, float* D, int N, int fmt, int layout_type, int lbo_a, int lbo_b, int sbo,
                         int kstep_a, int kstep_b, cudaStream_t st) {
    if (N < 16 || N > 128 || N % 16) return fail(DFP_ERR_INVALID, "bad probe shape");
    const size_t smem = 65536 + 1024;
    int rc = set_smem(k_tc_probe_bf16, smem);
    if (rc) return rc;
    DFP_TIMED("k_tc_probe_bf16", st, k_tc_probe_bf16<<<1, 128, smem, st>>>(A, B, D, N, fmt, layout_type, lbo_a, lbo_b, sbo, kstep_a, kstep_b));
    DFP_LAUNCH_CHECK("k_tc_probe_bf16");
    return DFP_OK;
}

int launch_tc_probe(float* D, int N, int b_mn, int lbo, int sbo, int k_hot, int layout_type, int start_off, cudaStream_t st) {
    if (layout_type >= 100) {          // M = 64 / 128 accumulator-layout probe: N is reused as M
        const size_t smem2 = 4096 * sizeof(float) + 1024;
        int rc2 = set_smem(k_tc_probe_m64, smem2);
        if (rc2) return rc2;
        DFP_TIMED("k_tc_probe_m64", st, k_tc_probe_m64<<<1, 128, smem2, st>>>(D, N));
        DFP_LAUNCH_CHECK("k_tc_probe_m64");
        return DFP_OK;
    }
    const size_t smem = 16384 * sizeof(float) + 1024;
    int rc = set_smem(k_tc_probe, smem);
    if (rc) return rc;
    DFP_TIMED("k_tc_probe", st, k_tc_probe<<<1, 128, smem, st>>>(D, N, b_mn, lbo, sbo, k_hot, layout_type, start_off));
    DFP_LAUNCH_CHECK("k_tc_probe");
    return DFP_OK;
}

int launch_tc_selftest_mode(const float* A, const float* W, float* D, int K, int N, int mode, cudaStream_t st) {
    if (mode == 1) {            // K = J, N = KIN
        if (K < 8 || K % 8 || N < 16 || N % 16 || N > 256 || 2 * K + N > 512) return fail(DFP_ERR_INVALID, "bad selftest shape");
        const size_t smem = 2 * (size_t)K * ((N + 31) / 32) * 32 * sizeof(float) + 1024;
        int rc = set_smem(k_tc_selftest_bT, smem);
        if (rc) return rc;
        DFP_TIMED("k_tc_selftest_bT", st, k_tc_selftest_bT<<<1, 128, smem, st>>>(A, W, D, K, N));
        DFP_LAUNCH_CHECK("k_tc_selftest_bT");
        return DFP_OK;
    }
    if (mode == 2) {            // K = F (features of H, <= 32), N = delta columns (<= 64)
        if (K < 4 || K > 32 || K % 4 || N < 16 || N % 16 || N > 64) return fail(DFP_ERR_INVALID, "bad selftest shape");
        const size_t smem = 160 * 1024;     // H tiles + delta tiles + room for the M=128 over-read
        int rc = set_smem(k_tc_selftest_outer, smem);
        if (rc) return rc;
        DFP_TIMED("k_tc_selftest_outer", st, k_tc_selftest_outer<<<1, 128, smem, st>>>(A, W, D, K, N));
        DFP_LAUNCH_CHECK("k_tc_selftest_outer");
        return DFP_OK;
    }
    return fail(DFP_ERR_INVALID, "bad selftest mode");
}

int launch_tc_selftest(const float* A, const float* W, float* D, int K, int N, int split, cudaStream_t st) {
    if (split >= 10) return launch_tc_selftest_mode(A, W, D, K, N, split - 10, st);
    if (K < 8 || K % 8 || N < 16 || N % 16 || N > 256 || 2 * K + N > 512)
        return fail(DFP_ERR_INVALID, "selftest needs K%%8==0, N%%16==0, N<=256, 2K+N<=512");
    const size_t smem = 2 * (size_t)N * K * sizeof(float);
    int rc = set_smem(k_tc_selftest, smem);
    if (rc) return rc;
    DFP_TIMED("k_tc_selftest", st, k_tc_selftest<<<1, 128, smem, st>>>(A, W, D, K, N, split));
    DFP_LAUNCH_CHECK("k_tc_selftest");
    return DFP_OK;
}

}  // namespace dfp

// C entry points of the self-test library (bound by tests/selftest/binding.py)
using namespace dfp;
extern "C" {
const char* dfpt_last_error(void) { return g_selftest_err; }
// D[128,N] = A[128,K] * W[N,K]^T through the tcgen05/TMEM path (split=1: 3xTF32; 11 / 12: transposed-weight / outer-product patterns)
int dfpt_selftest_tc(const float* A, const float* W, float* D, int32_t K, int32_t N, int32_t split, void* stream) {
    if (!A || !W || !D) return fail(DFP_ERR_INVALID, "null argument");
    return launch_tc_selftest(A, W, D, K, N, split, (cudaStream_t)stream);
}
// probes of tcgen05 shared-memory descriptor semantics
int dfpt_probe_tc(float* D, int32_t N, int32_t b_mn, int32_t lbo, int32_t sbo, int32_t k_hot, int32_t layout_type,
                  int32_t start_off, void* stream) {
    if (!D) return fail(DFP_ERR_INVALID, "null argument");
    return launch_tc_probe(D, N, b_mn, lbo, sbo, k_hot, layout_type, start_off, (cudaStream_t)stream);
}
int dfpt_probe_bf16_k(const float* A, const float* W, float* D, int32_t K, int32_t N, int32_t lbo_a, int32_t sbo_a, int32_t lbo_b,
                      int32_t sbo_b, void* stream) {
    if (!A || !W || !D) return fail(DFP_ERR_INVALID, "null argument");
    return launch_tc_probe_bf16_k(A, W, D, K, N, lbo_a, sbo_a, lbo_b, sbo_b, (cudaStream_t)stream);
}
int dfpt_probe_bf16_kt(const float* A, const float* Wt, float* D, int32_t K, int32_t N, int32_t lbo_a, int32_t sbo_a, int32_t lbo_b,
                       int32_t sbo_b, void* stream) {
    if (!A || !Wt || !D) return fail(DFP_ERR_INVALID, "null argument");
    return launch_tc_probe_bf16_k(A, Wt, D, K, N, lbo_a, sbo_a, lbo_b, sbo_b, (cudaStream_t)stream, true);
}
int dfpt_probe_bf16(const float* A, const float* B, float* D, int32_t N, int32_t fmt, int32_t layout_type, int32_t lbo_a, int32_t lbo_b,
                    int32_t sbo, int32_t kstep_a, int32_t kstep_b, void* stream) {
    if (!A || !B || !D) return fail(DFP_ERR_INVALID, "null argument");
    return launch_tc_probe_bf16(A, B, D, N, fmt, layout_type, lbo_a, lbo_b, sbo, kstep_a, kstep_b, (cudaStream_t)stream);
}
}  // extern "C"
