// tcgen05 / TMEM building blocks (sm_100a inline PTX) for the tensor-core MLP kernels.
//
// Scheme used by every tensor-core kernel in this library ("TS" form of tcgen05.mma):
//   * one thread owns one sample = one TMEM lane (row of the M=128 tile);
//   * the A operand (activations) is written by its owner thread straight into TMEM with
//     tcgen05.st (32x32b shape: lane = row, consecutive columns = consecutive K elements);
//   * the B operand (weights, [N][K]) sits in shared memory in the canonical K-major, no-swizzle
//     layout: 16-byte chunk ck = k/4 of row n at byte  ck*(N*16) + n*16   (core matrix = 8 rows x 16 B,
//     SBO = 128 B between 8-row groups, LBO = N*16 B between K chunks);
//   * the accumulator D [128 x N] f32 lives in TMEM and is read back by the owner threads with
//     tcgen05.ld for the epilogue (bias, ReLU, feature term, split for the next layer).
// fp32 accuracy is kept with the 3xTF32 split: x = hi + lo (both exactly representable in TF32),
// D = A_hi*B_hi + A_lo*B_hi + A_hi*B_lo, accumulated in f32 by the tensor core.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dfp {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM allocation (one full warp executes alloc / dealloc) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- ordering ----
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// producer side of a named barrier: count this thread's arrival and go on (the consumer warp uses named_bar_sync)
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// one lane of a converged warp (elect.sync): the compiler then knows that a single thread runs the guarded code and
// keeps descriptors in uniform registers instead of looping over the lanes' values for every tcgen05.mma
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// software prefetch of a line that a later tile of this CTA will load (no register is held across the tile)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// 16 consecutive floats of a 16-byte aligned shared-memory row as four 128-bit loads (bias / output-row slices of the
// epilogues: scalar loads of the same values were a measurable share of the shared-memory pipe)
__device__ __forceinline__ void lds16(const float* p, float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// all previously issued tcgen05.mma of this thread arrive on `bar` when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- bulk asynchronous copies (TMA, 1-D): global -> shared, completion counted in bytes on an mbarrier ----
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- bf16 operands, both MN-major (row-per-sample), no swizzle: see dfp_planes.cuh for the layout ----
__host__ __device__ constexpr uint32_t idesc_bf16_mn(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ uint64_t smem_desc_plane(uint32_t saddr, uint32_t group_bytes) {      // LBO = 8-sample group stride, SBO = 128
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((group_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// D (+)= A^T B over 16 samples with the two-plane split: hi*hi' + hi*mid' + mid*hi' (small terms first)
__device__ __forceinline__ void mma_planes(uint32_t d_tmem, uint32_t a_hi, uint32_t a_mid, uint32_t a_group_bytes, uint32_t b_hi,
                                           uint32_t b_mid, uint32_t b_group_bytes, uint32_t idesc, bool accumulate) {
    const uint64_t ah = smem_desc_plane(a_hi, a_group_bytes), am = smem_desc_plane(a_mid, a_group_bytes);
    const uint64_t bh = smem_desc_plane(b_hi, b_group_bytes), bm = smem_desc_plane(b_mid, b_group_bytes);
    mma_bf16_ss(d_tmem, am, bh, idesc, accumulate ? 1u : 0u);
    mma_bf16_ss(d_tmem, ah, bm, idesc, 1u);
    mma_bf16_ss(d_tmem, ah, bh, idesc, 1u);
}

// ---- descriptors ----
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (see header comment for the layout)
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: kind::tf32, f32 accumulate, M = 128 (or 64); a_mn / b_mn select the MN-major (transposed) view
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n, bool a_mn = false, bool b_mn = false) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_tf32_m128(int n, bool a_mn = false, bool b_mn = false) {
    return idesc_tf32(128, n, a_mn, b_mn);
}
// ---- MN-major (transposed) TF32 operands ----------------------------------------------------------------
// Probed on B200 (dfp_probe_tc): for kind::tf32 the transposed view only exists with layout type 1
// (SWIZZLE_128B_BASE32B).  An operand X[K rows][MN cols] is then stored as row-major tiles of 32 columns:
//     word(k, mn) = (mn/32)*LBO/4 + (k/4)*SBO/4 + (k%4)*32 + (((mn%32)/8) ^ (k%4))*8 + mn%8
// i.e. each K row holds 32 consecutive MN elements (128 B), 4-row groups are SBO apart (512 B when rows are
// simply consecutive), 32-column blocks are LBO apart, and the four 32-byte groups of a row are XOR-swizzled
// with address bits 7..8 (the tile base must be 512-byte aligned).
__host__ __device__ __forceinline__ int mn_tile_index(int row, int col) {              // [rows][32] tile, in floats
    return row * 32 + ((((col >> 3) ^ row) & 3) << 3) + (col & 7);
}
__device__ __forceinline__ uint64_t smem_desc_mnmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes = 512u) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (1ull << 61);
}
// D[tmem] (+)= A[smem] * B[smem], one K=8 step
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T, one K=8 step
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}

// ---- TF32 split ----
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// x = hi + lo with both parts TF32 numbers.  cvt.rna.tf32.f32 is not a native instruction on sm_100a (it expands to
// add / compare-with-infinity / select / mask, ~4 instructions), and this split runs for every activation and delta of
// every layer; for the finite values that reach it the rounding is an integer add of half a TF32 ulp and a mask.
// The tensor core reads only the upper 19 bits of a TF32 operand, so the residual needs the add but not the mask.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}

// ---- TMEM <-> registers, 32x32b shape: thread t of warp w touches lane 32*(w%4)+t ----
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// 16-column variants for kernels that split one sample's 32 columns between two threads
__device__ __forceinline__ void tmem_store_split16(uint32_t taddr_hi, uint32_t taddr_lo, const float (&x)[16]) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) split_tf32(x[i], hi[i], lo[i]);
    tmem_st16(taddr_hi, hi);
    tmem_st16(taddr_lo, lo);
}
// split 32 activations and store them as the (hi, lo) A-operand column blocks
__device__ __forceinline__ void tmem_store_split32(uint32_t taddr_hi, uint32_t taddr_lo, const float (&x)[32]) {
    uint32_t hi[32], lo[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) split_tf32(x[i], hi[i], lo[i]);
    tmem_st32(taddr_hi, hi);
    tmem_st32(taddr_lo, lo);
}

// canonical B-operand position (in floats) of element (n, k) of an [N][K] matrix
__host__ __device__ __forceinline__ int b_index(int n, int k, int N) { return (k >> 2) * (N * 4) + n * 4 + (k & 3); }

// D[128 x N] (+)= A[128 x K] * B[N x K]^T with the 3xTF32 split.  a_hi / a_lo: TMEM column addresses of the split A
// (lane field 0); b_hi / b_lo: shared addresses of the split canonical B; one thread calls this.
__device__ __forceinline__ void mma_3xtf32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int K,
                                           int N, bool accumulate) {
    const uint32_t idesc = idesc_tf32_m128(N);
    const uint32_t lbo = (uint32_t)N * 16u, sbo = 128u;
    uint32_t acc = accumulate ? 1u : 0u;
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t dh = smem_desc_kmajor(b_hi + ks * 2 * lbo, lbo, sbo);
        const uint64_t dl = smem_desc_kmajor(b_lo + ks * 2 * lbo, lbo, sbo);
        mma_tf32_ts(d_tmem, a_lo + ks * 8, dh, idesc, acc);      // small terms first
        mma_tf32_ts(d_tmem, a_hi + ks * 8, dl, idesc, 1u);
        mma_tf32_ts(d_tmem, a_hi + ks * 8, dh, idesc, 1u);
        acc = 1u;
    }
}

// D[128 x N] (+)= A[128 x J] * W[J x N]: A (hi/lo) in TMEM, W as natural row-major [J][N] swizzled tiles (mn_tile_index,
// one tile per 32 columns, `colblock_bytes` apart), i.e. the MN-major B operand.              (dH = delta * W)
__device__ __forceinline__ void mma_3xtf32_bT(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, uint32_t w_lo, int J,
                                              int N, uint32_t colblock_bytes, bool accumulate) {
    const uint32_t idesc = idesc_tf32_m128(N, false, true);
    uint32_t acc = accumulate ? 1u : 0u;
    for (int ks = 0; ks < J / 8; ++ks) {                      // 8 rows of W per step = two 4-row groups = 1024 B
        const uint64_t dh = smem_desc_mnmajor(w_hi + ks * 1024u, colblock_bytes);
        const uint64_t dl = smem_desc_mnmajor(w_lo + ks * 1024u, colblock_bytes);
        mma_tf32_ts(d_tmem, a_lo + ks * 8, dh, idesc, acc);
        mma_tf32_ts(d_tmem, a_hi + ks * 8, dl, idesc, 1u);
        mma_tf32_ts(d_tmem, a_hi + ks * 8, dh, idesc, 1u);
        acc = 1u;
    }
}

// D[128 (first F rows valid) x N] (+)= H^T * Dl over the 128 samples of a tile.  H: [128][F] row-per-sample swizzled
// tiles split hi/lo (32-feature blocks `h_block_bytes` apart); Dl: [128][N] likewise, hi part only (the dropped
// H_hi*Dl_lo term is < 2^-12 of the product).  M is always 128: feature blocks beyond F read whatever lies at the
// following block addresses and only produce junk in accumulator lanes >= F.                 (dW = H^T * delta)
__device__ __forceinline__ void mma_outer(uint32_t d_tmem, uint32_t h_hi, uint32_t h_lo, uint32_t h_block_bytes, uint32_t d_hi,
                                          uint32_t d_block_bytes, int N, bool accumulate) {
    const uint32_t idesc = idesc_tf32_m128(N, true, true);
    uint32_t acc = accumulate ? 1u : 0u;
    for (int ks = 0; ks < 16; ++ks) {                         // 128 samples, 8 per step (1024 B of each tile)
        const uint64_t ah = smem_desc_mnmajor(h_hi + ks * 1024u, h_block_bytes), al = smem_desc_mnmajor(h_lo + ks * 1024u, h_block_bytes);
        const uint64_t bh = smem_desc_mnmajor(d_hi + ks * 1024u, d_block_bytes);
        mma_tf32_ss(d_tmem, al, bh, idesc, acc);
        mma_tf32_ss(d_tmem, ah, bh, idesc, 1u);
        acc = 1u;
    }
}

// M = 64 variant of mma_outer: D[64 x N] (+)= [H | X]^T * Dl, where the A operand's second 32-feature block X lives
// `h*_lbo_bytes` after the first (any shared block: the callers point the hi part at a constant-one block, giving the
// column sums of Dl in accumulator row 32, and the lo part at a zero block).  Accumulator row m sits in TMEM lane
// (m/16)*32 + m%16 (probed: dfp_probe_tc, M = 64).  Half the shared-memory operand traffic of the M = 128 form.
__device__ __forceinline__ void mma_outer64(uint32_t d_tmem, uint32_t h_hi, uint32_t h_hi_lbo_bytes, uint32_t h_lo,
                                            uint32_t h_lo_lbo_bytes, uint32_t d_hi, int N, bool accumulate) {
    const uint32_t idesc = idesc_tf32(64, N, true, true);
    uint32_t acc = accumulate ? 1u : 0u;
    for (int ks = 0; ks < 16; ++ks) {
        const uint64_t bh = smem_desc_mnmajor(d_hi + ks * 1024u, 16384u);
        mma_tf32_ss(d_tmem, smem_desc_mnmajor(h_lo + ks * 1024u, h_lo_lbo_bytes), bh, idesc, acc);
        mma_tf32_ss(d_tmem, smem_desc_mnmajor(h_hi + ks * 1024u, h_hi_lbo_bytes), bh, idesc, 1u);
        acc = 1u;
    }
}

}  // namespace tc
}  // namespace dfp
