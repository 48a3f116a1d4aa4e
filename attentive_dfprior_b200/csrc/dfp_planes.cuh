// "Planes": the per-sample operands of the weight-gradient products, stored by their producers (the tensor-core
// forward and the delta-chain kernel) in exactly the form the tcgen05 tensor cores read them, so that the
// weight-gradient kernel moves them HBM -> shared memory with bulk copies and touches no value in registers.
//
// A value x is split into two bf16 numbers hi = bf16(x), mid = bf16(x - hi) (|x - hi - mid| <= 2^-18 |x|); a product
// sum is accumulated in fp32 as  hi*hi' + hi*mid' + mid*hi'  (relative error ~1e-5, better than the single TF32 rounding
// of the deltas it replaces).  One array holds F features of every sample; for each 128-sample tile it stores the hi
// plane and then the mid plane, each in the no-swizzle MN-major core-matrix layout (probed on B200 with
// tools/probe_bf16.py: layout type 0, LBO = byte distance between 8-sample groups, SBO = 128):
//
//     byte(tile, plane, k, f) = (tile*2 + plane) * 256*F + (k/8) * 16*F + (f/8) * 128 + (k%8) * 16 + (f%8) * 2
//
// i.e. [8-sample group][8-feature chunk][sample in group][feature in chunk].  Any feature range [f0, f0+M) (multiples of
// 8) of any run of 8-sample groups is an MMA operand: start = byte(.., k, f0), LBO = 16*F, SBO = 128.  A thread that
// owns one sample writes 16-byte chunks (8 features of one plane); 8 consecutive samples fill a 128-byte line.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace dfp {
namespace pl {

// ---- the arrays of one decoder (features per sample) --------------------------------------------------------
// H   : [h3 | h1 | h0 | h2 | h4 | ones8]     inputs of pts_linears.4/.2/.1/.3 and of output_linear, then a constant 1
// CX  : [c (CD) | ones8]                     gathered grid features, then a constant 1
//       (the constant chunks [1 0 0 0 0 0 0 0] exist only in the weight-gradient kernel's shared-memory stages, written
//        once per CTA: global memory holds F_H_ST = 160 and CD features per sample, copied group by group)
// (the gradients dh_i of the blocks' outputs, which feed fc_c.i, are NOT stored: dh_{i-1} = dacc_i W_i is linear in the
//  masked delta, so  c^T dh_{i-1} = (c^T dacc_i) W_i  is formed from the accumulated CX^T DL0 product afterwards)
// DLX : [dout8 | dacc4 | dacc2 | dacc1 | dacc3 | dacc0 | darg96]   ONE array of which the weight-gradient kernel reads two
//       overlapping feature ranges, so that dacc3 (needed by both) is stored once:
//   DL0 = DLX[0:136)   : upstream gradient, then ReLU-masked deltas in H's block order
//   DL1 = DLX[104:264) : [dacc3 | dacc0 | darg96], deltas that meet the embedding, then d(embedding argument)
// EX  : [e96 (column 93 == 1) | p0 p1 p2 0*5] sin(embedding argument) and the point (for embedder._B)
constexpr int F_H = 168, F_DL0 = 136, F_DL1 = 160, F_EX = 104;
constexpr int F_DLX = 264, DL1_AT = 104;
constexpr int F_H_ST = 160;                       // features of H kept in global memory (H without its constant chunk)          // features of the stored array; first feature of the DL1 range
__host__ __device__ constexpr int f_cx(int CD) { return CD + 8; }
// position (32-feature block) of h_i inside H, and of dacc_i inside DL0 (after the 8 dout features)
__host__ __device__ constexpr int h_pos(int i) { return i == 0 ? 2 : i == 1 ? 1 : i == 2 ? 3 : i == 3 ? 0 : 4; }
__host__ __device__ constexpr int dl0_pos(int i) { return i == 4 ? 0 : i == 2 ? 1 : i == 1 ? 2 : 3; }      // i in 1..4
__host__ __device__ constexpr int64_t array_bytes(int64_t n_padded, int F) { return n_padded * F * 4; }      // two bf16 planes

__device__ __forceinline__ unsigned char* chunk_ptr(unsigned char* base, int F, int64_t s, int plane, int f0) {
    const int64_t tile = s >> 7;
    const int k = (int)(s & 127);
    return base + (tile * 2 + plane) * (int64_t)(256 * F) + (k >> 3) * (16 * F) + (f0 >> 3) * 128 + (k & 7) * 16;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {        // lo -> bits 0..15
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// fp16 variant of the split (11-bit mantissas: hi + mid carries 22 bits, for operands of bounded range)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {         // lo -> bits 0..15
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void unpack_f16x2(uint32_t v, float& lo, float& hi) {
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(v));
}
template <bool F16>
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& mid) {
    uint32_t h[4], m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float f0, f1;
        if (F16) { h[i] = pack_f16x2(v[2 * i], v[2 * i + 1]); unpack_f16x2(h[i], f0, f1); }
        else { h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]); f0 = __uint_as_float(h[i] << 16); f1 = __uint_as_float(h[i] & 0xffff0000u); }
        m[i] = F16 ? pack_f16x2(v[2 * i] - f0, v[2 * i + 1] - f1) : pack_bf16x2(v[2 * i] - f0, v[2 * i + 1] - f1);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    mid = make_uint4(m[0], m[1], m[2], m[3]);
}

// 8 consecutive features [f0, f0+8) of sample s (list position): hi and mid planes, one 16-byte store each
__device__ __forceinline__ void store8(unsigned char* base, int F, int64_t s, int f0, const float* v) {
    uint32_t h[4], m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        const float r0 = v[2 * i] - __uint_as_float(h[i] << 16), r1 = v[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u);
        m[i] = pack_bf16x2(r0, r1);
    }
    *reinterpret_cast<uint4*>(chunk_ptr(base, F, s, 0, f0)) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(chunk_ptr(base, F, s, 1, f0)) = make_uint4(m[0], m[1], m[2], m[3]);
}
__device__ __forceinline__ void store16(unsigned char* base, int F, int64_t s, int f0, const float (&v)[16]) {
    store8(base, F, s, f0, v);
    store8(base, F, s, f0 + 8, v + 8);
}
// the constant-one chunk [1, 0, 0, 0, 0, 0, 0, 0] (hi plane; the mid plane is zero)
__device__ __forceinline__ void store_ones8(unsigned char* base, int F, int64_t s, int f0) {
    *reinterpret_cast<uint4*>(chunk_ptr(base, F, s, 0, f0)) = make_uint4(0x00003f80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(chunk_ptr(base, F, s, 1, f0)) = make_uint4(0u, 0u, 0u, 0u);
}

// A producer thread's row of one array: the two plane addresses of its sample are computed once per tile, every store is
// then `row + chunk * 128` (the 64-bit address arithmetic of chunk_ptr otherwise repeats for each of ~40 stores).
struct Row {
    unsigned char* hi;          // chunk 0 of the hi plane
    uint32_t plane;             // distance to the mid plane
};
__device__ __forceinline__ Row row_of(void* base, int F, int64_t s) {
    Row r;
    r.hi = chunk_ptr(reinterpret_cast<unsigned char*>(base), F, s, 0, 0);
    r.plane = 256u * (uint32_t)F;
    return r;
}
__device__ __forceinline__ void store8(const Row& r, int f0, const float* v) {
    uint4 hi, mid;
    split8<false>(v, hi, mid);
    unsigned char* p = r.hi + (f0 >> 3) * 128;
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + r.plane) = mid;
}
__device__ __forceinline__ void store16(const Row& r, int f0, const float (&v)[16]) {
    store8(r, f0, v);
    store8(r, f0 + 8, v + 8);
}
__device__ __forceinline__ void store_ones8(const Row& r, int f0) {
    unsigned char* p = r.hi + (f0 >> 3) * 128;
    *reinterpret_cast<uint4*>(p) = make_uint4(0x00003f80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(p + r.plane) = make_uint4(0u, 0u, 0u, 0u);
}

}  // namespace pl
}  // namespace dfp
