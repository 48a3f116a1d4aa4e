// Self-test of the tcgen05 building blocks (dfp_tc.cuh): D[128 x N] = A[128 x K] * W[N x K]^T with the A operand
// written into TMEM by its owner threads, W in the canonical shared-memory layout, the 3xTF32 split, and D read
// back from TMEM.  tests/test_gpu_tc.py compares it with an f64 matmul.
#include "dfp_device.cuh"
#include "dfp_tc.cuh"

namespace dfp {

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

int launch_tc_selftest(const float* A, const float* W, float* D, int K, int N, int split, cudaStream_t st) {
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
