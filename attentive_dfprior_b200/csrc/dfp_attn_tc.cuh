// Shared pieces of the tensor-core attentive-fusion kernels (dfp_attn_tc.cu): the fusion MLP 2-64-128-128-64-2 of
// src/conv_onet/models/decoder.py:206-258 evaluated for tiles of 128 TSDF-band samples.
//
// All operands of the three wide layers are bf16 hi + mid planes in shared memory and every product is three SS-mode
// tcgen05.mma (hi*hi' + hi*mid' + mid*hi', fp32 accumulation in TMEM, ~1e-5 relative):
//   * weights: the image the pack kernel appends to the packed parameters (dfp_layout.h AttImg), one bulk copy;
//   * activations / deltas: written by their owner threads (one sample = one row) in the plane layout of
//     dfp_planes.cuh, [8-sample group][8-feature chunk][sample][feature], which the tensor core reads K-major with the
//     sample as the M dimension (LBO = 128 between feature chunks, SBO = 16 F between sample groups).
// The same bytes are, for the weight-gradient kernel, MN-major operands with the sample as K -- so what a layer leaves
// in shared memory for the next layer's MMA is stored to global memory with ONE bulk copy and read back by bulk copies.
#pragma once
#include "dfp_device.cuh"
#include "dfp_planes.cuh"
#include "dfp_tc.cuh"

namespace dfp {
namespace att {

constexpr int THREADS = 256;                    // two threads per sample: thread (row r, half hf) owns half of every layer's columns
constexpr uint32_t ACT_PLANE = 32768;           // bytes reserved per plane in shared memory (128 samples x 128 features x bf16)
// dynamic shared memory (bytes, from a 1024-aligned base)
constexpr uint32_t SM_IMG = 0;                              // AttImg::BYTES weight image
constexpr uint32_t SM_ACT = AttImg::BYTES;                  // activation / delta planes: hi at +0, mid at +ACT_PLANE
constexpr uint32_t SM_SMALL = SM_ACT + 2 * ACT_PLANE;       // fp32: W1 [2][64] | B1 [64] | B2 [128] | B3 [128] | B4 [64] | W5 [2][64] | B5 [4]
constexpr int S_W1 = 0, S_B1 = 128, S_B2 = 192, S_B3 = 320, S_B4 = 448, S_W5 = 512, S_B5 = 640, SMALL_FLOATS = 644;
constexpr uint32_t SM_BYTES = SM_SMALL + SMALL_FLOATS * 4;
// backward chain only: [1 x0 x1 0*5] (and [dl0 dl1 0*6]) planes of the tile, 2 KB per plane, + slack for the M = 64 over-read
constexpr uint32_t SM_X8 = (SM_BYTES + 127) / 128 * 128;     // hi at +0, mid at +2048
constexpr uint32_t SM_DL8 = SM_X8 + 4096;                    // hi at +0, mid at +2048
constexpr uint32_t SM_BYTES_BWD = SM_DL8 + 4096 + 2048;

// saved per tile for the backward (global memory): the four activation arrays as planes, hi then mid each
constexpr uint32_t G_H1 = 0, G_H2 = G_H1 + 2 * 128 * 64 * 2, G_H3 = G_H2 + 2 * 128 * 128 * 2, G_H4 = G_H3 + 2 * 128 * 128 * 2;
constexpr uint32_t G_TILE = G_H4 + 2 * 128 * 64 * 2;        // 196608 bytes per 128-sample tile (1536 B per sample)
// the chain kernel's deltas for the weight-gradient kernel: d2 (128), d3 (128), d4 (64), hi then mid each
constexpr uint32_t G_D2 = 0, G_D3 = G_D2 + 2 * 128 * 128 * 2, G_D4 = G_D3 + 2 * 128 * 128 * 2;
constexpr uint32_t G_DL = G_D4 + 2 * 128 * 64 * 2;          // [dl0 dl1 0*6]: gradient of the two logits
constexpr uint32_t G_DTILE = G_DL + 2 * 128 * 8 * 2;        // 167936 bytes per tile (1312 B per sample)
constexpr int MASK_WORDS = 12;                  // ReLU masks per sample: h1 2 words, h2 4, h3 4, h4 2; word w of a tile at [w][128 samples]

__device__ __forceinline__ void load_small(float* sm, const float* __restrict__ P /* packed fusion parameters */) {
    for (int i = threadIdx.x; i < SMALL_FLOATS; i += blockDim.x) {
        float v;
        if (i < S_B1) v = P[AttOff::W1 + i];
        else if (i < S_B2) v = P[AttOff::B1 + i - S_B1];
        else if (i < S_B3) v = P[AttOff::B2 + i - S_B2];
        else if (i < S_B4) v = P[AttOff::B3 + i - S_B3];
        else if (i < S_W5) v = P[AttOff::B4 + i - S_B4];
        else if (i < S_B5) v = P[AttOff::W5 + i - S_W5];
        else v = P[AttOff::B5 + i - S_B5];
        sm[i] = v;
    }
}

// 8 consecutive features [f0, f0+8) of row r into the shared-memory planes of an F-feature array (F16: fp16 planes
// for the forward's activations, else bf16 planes for the backward's deltas)
template <bool F16>
__device__ __forceinline__ void put8(unsigned char* act, int F, int r, int f0, const float* v) {
    uint4 hi, mid;
    pl::split8<F16>(v, hi, mid);
    const uint32_t off = (uint32_t)(r >> 3) * (16u * F) + (uint32_t)(f0 >> 3) * 128u + (uint32_t)(r & 7) * 16u;
    *reinterpret_cast<uint4*>(act + off) = hi;
    *reinterpret_cast<uint4*>(act + ACT_PLANE + off) = mid;
}

// D[128 x N] = A[128 x K] * W^T: A = planes in shared memory (K features), W = forward image (N outputs, K inputs)
__device__ __forceinline__ void mma_layer_fwd(uint32_t d_tmem, uint32_t act, int K, uint32_t w_hi, uint32_t w_mid, int N) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);   // f16 x f16 -> f32, both K-major
    const uint32_t sbo_a = 16u * K, lbo_b = 16u * N;
    for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t ah = tc::smem_desc_kmajor(act + ks * 256u, 128u, sbo_a), am = tc::smem_desc_kmajor(act + ACT_PLANE + ks * 256u, 128u, sbo_a);
        const uint64_t bh = tc::smem_desc_kmajor(w_hi + ks * 2u * lbo_b, lbo_b, 128u), bm = tc::smem_desc_kmajor(w_mid + ks * 2u * lbo_b, lbo_b, 128u);
        tc::mma_bf16_ss(d_tmem, am, bh, idesc, ks > 0 ? 1u : 0u);     // small terms first
        tc::mma_bf16_ss(d_tmem, ah, bm, idesc, 1u);
        tc::mma_bf16_ss(d_tmem, ah, bh, idesc, 1u);
    }
}
// D[128 x NIN] = Delta[128 x J] * W: Delta = planes (J features), W = the forward image of a layer with J outputs and
// NIN inputs read transposed (MN-major B: LBO = 128 between 8-output groups, SBO = 16 J between 8-input chunks)
__device__ __forceinline__ void mma_layer_bwd(uint32_t d_tmem, uint32_t act, int J, uint32_t w_hi, uint32_t w_mid, int NIN) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(NIN >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sbo_a = 16u * J, sbo_b = 16u * J;
    for (int ks = 0; ks < J / 16; ++ks) {
        const uint64_t ah = tc::smem_desc_kmajor(act + ks * 256u, 128u, sbo_a), am = tc::smem_desc_kmajor(act + ACT_PLANE + ks * 256u, 128u, sbo_a);
        const uint64_t bh = tc::smem_desc_kmajor(w_hi + ks * 256u, 128u, sbo_b), bm = tc::smem_desc_kmajor(w_mid + ks * 256u, 128u, sbo_b);
        tc::mma_bf16_ss(d_tmem, am, bh, idesc, ks > 0 ? 1u : 0u);
        tc::mma_bf16_ss(d_tmem, ah, bm, idesc, 1u);
        tc::mma_bf16_ss(d_tmem, ah, bh, idesc, 1u);
    }
}

// bulk copy shared -> global (one thread); completion tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_store_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace att
}  // namespace dfp
