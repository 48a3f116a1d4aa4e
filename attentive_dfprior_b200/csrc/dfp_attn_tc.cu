// Attentive fusion on the tcgen05 tensor cores: forward (k_attn_fwd_tc), backward delta chain (k_attn_dh_tc) and
// weight gradients (k_attn_dw_tc).  See dfp_attn_tc.cuh for the operand scheme.
#include "dfp_attn_tc.cuh"

namespace dfp {

using namespace att;

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
// Epilogue of one wide layer: this thread's NH = N/2 accumulator columns -> + bias, ReLU -> planes of the next layer's
// input (in place of the previous activations, which the finished MMAs no longer need).  Returns the ReLU bits.
// 8 features of row r straight into a saved bf16 plane pair in global memory (hi plane, then mid plane, 256 F bytes each)
__device__ __forceinline__ void gput8(unsigned char* garr, int F, int r, int f0, const float* v) {
    uint4 hi, mid;
    pl::split8<false>(v, hi, mid);
    const uint32_t off = (uint32_t)(r >> 3) * (16u * F) + (uint32_t)(f0 >> 3) * 128u + (uint32_t)(r & 7) * 16u;
    *reinterpret_cast<uint4*>(garr + off) = hi;
    *reinterpret_cast<uint4*>(garr + 256u * F + off) = mid;
}

// fp16 planes carry 11 + 11 mantissa bits but only 2^-14 .. 2^15 of range, and the fusion MLP's activations scale with
// its inputs (logits near zero at the surface: rows of 1e-4 and below, where an fp16 "mid" part underflows).  Every row
// of a layer's input is therefore multiplied by a power of two that puts its largest element in [256, 512) -- exact,
// and undone (together with the weight image's constant 2^6) when the accumulator is read back.
__device__ __forceinline__ float row_scale(float rowmax, float& inv) {
    const int e = (int)((__float_as_uint(rowmax) >> 23) & 0xffu);
    if (e == 0) { inv = 1.f; return 1.f; }                             // an all-zero (or denormal) row
    int es = 262 - e;                                                  // 2^(8 - floor(log2(rowmax)))
    es = es < 1 ? 1 : (es > 253 ? 253 : es);
    inv = __uint_as_float((uint32_t)(254 - es) << 23);
    return __uint_as_float((uint32_t)es << 23);
}
// the two threads of a row agree on the row's largest activation through shared memory
__device__ __forceinline__ float row_max_exchange(float mine, float (*rowmax_s)[128], int r, int hf) {
    rowmax_s[hf][r] = mine;
    tc::named_bar_sync(1, THREADS);
    return fmaxf(mine, rowmax_s[hf ^ 1][r]);
}

// Epilogue of one wide layer: this thread's NH = N/2 accumulator columns -> undo the operand scales, + bias, ReLU ->
// fp16 planes of the next layer's input (in place of the previous activations, which the finished MMAs no longer need),
// scaled by the row's new power of two (returned through `inv`).  Two passes over the accumulator: row maximum, then values.
template <int NH, bool LAST>
__device__ __forceinline__ void fwd_epilogue(uint32_t tl, const float* __restrict__ bias, unsigned char* act, int r, int hf, int col0,
                                             float kin, float& inv, float (*rowmax_s)[128],
                                             uint32_t* mask /* NH/32 words */, float* keep /* LAST: the NH activations */,
                                             unsigned char* gsave /* bf16 planes kept for the backward, or null */) {
    constexpr int F = 2 * NH;
    float scale = 1.f;
    if (!LAST) {
        float mx = 0.f;
#pragma unroll
        for (int b = 0; b < NH / 32; ++b) {
            float v[32];
            tc::tmem_ld32(tl + col0 + 32 * b, v);
            tc::wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaf(v[j], kin, bias[col0 + 32 * b + j]));
        }
        scale = row_scale(row_max_exchange(mx, rowmax_s, r, hf), inv);
    }
#pragma unroll
    for (int b = 0; b < NH / 32; ++b) {
        float v[32];
        tc::tmem_ld32(tl + col0 + 32 * b, v);
        tc::wait_ld();
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float z = fmaf(v[j], kin, bias[col0 + 32 * b + j]);
            m |= (z > 0.f ? 1u : 0u) << j;
            v[j] = fmaxf(z, 0.f);
            if (LAST) keep[32 * b + j] = v[j];
        }
        mask[b] = m;
        if (gsave) {
#pragma unroll
            for (int c = 0; c < 4; ++c) gput8(gsave, F, r, col0 + 32 * b + 8 * c, v + 8 * c);
        }
        if (!LAST) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= scale;
#pragma unroll
            for (int c = 0; c < 4; ++c) put8<true>(act, F, r, col0 + 32 * b + 8 * c, v + 8 * c);
        }
    }
}

__global__ void __launch_bounds__(THREADS, 1) k_attn_fwd_tc(const AttArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bar_w, bar_mma;
    __shared__ uint32_t tmem_base_s;
    __shared__ float part[2][128];
    __shared__ float rowmax_s[2][128];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & 127, hf = tid >> 7;
    const int count = *a.band_count;
    const int tiles = (count + 127) >> 7;
    if ((int)blockIdx.x >= tiles) return;                             // nothing to do for this CTA (short band list)
    float* small = reinterpret_cast<float*>(sm + SM_SMALL);
    unsigned char* act = sm + SM_ACT;
    const uint32_t act_s = tc::smem_u32(act), img_s = tc::smem_u32(sm + SM_IMG);

    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 128); tc::tmem_relinquish(); }
    if (tid == 0) {
        tc::mbar_init(&bar_w, 1); tc::mbar_init(&bar_mma, 1); tc::fence_mbar_init();
        tc::mbar_arrive_expect_tx(&bar_w, AttImg::BYTES);            // the (fp16) weight image: one bulk copy, 128 KB
        tc::bulk_copy_g2s(img_s, a.params + OFF_ATT_IMG_F16, AttImg::BYTES, &bar_w);
    }
    load_small(small, a.params + OFF_ATT);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tb = tmem_base_s;
    const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16);
    const bool save = a.tc_planes != nullptr;
    uint32_t parity = 0;
    bool weights_ready = false;

    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int pos = tile * 128 + r;
        const bool valid = pos < count;
        const int s = a.band_list[valid ? pos : count - 1];
        const float x0 = a.high[s] + a.low[s];                         // decoder.py:326,343
        const float x1 = tsdf_logit(a.tsdf[s]).inv;
        uint32_t masks[6];
        float inv = 1.f;                                               // 1 / (this row's current operand scale)
        {   // pts_linears.0 (2 -> 64) on the CUDA cores: this thread's 32 units
            float h[32], mx = 0.f;
            uint32_t m = 0;
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                const int j = 32 * hf + u;
                const float z = fmaf(small[S_W1 + 64 + j], x1, fmaf(small[S_W1 + j], x0, small[S_B1 + j]));
                m |= (z > 0.f ? 1u : 0u) << u;
                h[u] = fmaxf(z, 0.f);
                mx = fmaxf(mx, h[u]);
            }
            masks[0] = m;
            if (save) {
#pragma unroll
                for (int c = 0; c < 4; ++c) gput8(a.tc_planes + (size_t)tile * G_TILE + G_H1, 64, r, 32 * hf + 8 * c, h + 8 * c);
            }
            const float scale = row_scale(row_max_exchange(mx, rowmax_s, r, hf), inv);
#pragma unroll
            for (int u = 0; u < 32; ++u) h[u] *= scale;
#pragma unroll
            for (int c = 0; c < 4; ++c) put8<true>(act, 64, r, 32 * hf + 8 * c, h + 8 * c);
        }
        unsigned char* gsave = save ? a.tc_planes + (size_t)tile * G_TILE : nullptr;
        float h4[32];
#pragma unroll 1
        for (int layer = 0; layer < 3; ++layer) {                      // pts_linears.1 (64->128), .2 (128->128), .3 (128->64)
            tc::fence_proxy_async();                                   // plane writes -> visible to the tensor core
            tc::fence_before_sync();
            tc::named_bar_sync(1, THREADS);
            if (warp == 0 && !weights_ready) { tc::mbar_wait(&bar_w, 0); weights_ready = true; }
            if (warp == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                const int K = layer == 0 ? 64 : 128, N = layer == 2 ? 64 : 128;
                const uint32_t wh = img_s + (layer == 0 ? AttImg::W2H : layer == 1 ? AttImg::W3H : AttImg::W4H);
                const uint32_t wm = img_s + (layer == 0 ? AttImg::W2M : layer == 1 ? AttImg::W3M : AttImg::W4M);
                mma_layer_fwd(tb, act_s, K, wh, wm, N);
                tc::mma_commit(&bar_mma);
            }
            tc::mbar_wait(&bar_mma, parity); parity ^= 1;
            tc::fence_after_sync();
            // the activations go to shared memory as fp16 planes (next layer's operand) and, when a backward follows,
            // to global memory as bf16 planes (operands of the weight-gradient kernel, which pairs them with bf16 deltas)
            const float kin = inv * (1.f / ATT_IMG_F16_SCALE);
            if (layer == 0) fwd_epilogue<64, false>(tl, small + S_B2, act, r, hf, 64 * hf, kin, inv, rowmax_s, masks + 1, nullptr, save ? gsave + G_H2 : nullptr);
            else if (layer == 1) fwd_epilogue<64, false>(tl, small + S_B3, act, r, hf, 64 * hf, kin, inv, rowmax_s, masks + 3, nullptr, save ? gsave + G_H3 : nullptr);
            else fwd_epilogue<32, true>(tl, small + S_B4, act, r, hf, 32 * hf, kin, inv, rowmax_s, masks + 5, h4, save ? gsave + G_H4 : nullptr);
        }
        if (save) {                                                    // ReLU masks for the backward
            uint32_t* gm = a.tc_masks + (size_t)tile * MASK_WORDS * 128;
            gm[(0 + hf) * 128 + r] = masks[0];
            gm[(2 + 2 * hf) * 128 + r] = masks[1]; gm[(3 + 2 * hf) * 128 + r] = masks[2];
            gm[(6 + 2 * hf) * 128 + r] = masks[3]; gm[(7 + 2 * hf) * 128 + r] = masks[4];
            gm[(10 + hf) * 128 + r] = masks[5];
        }
        // output_linear (64 -> 2) and the softmax: each half sums its 32 inputs, the lower half finishes
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            l0 = fmaf(small[S_W5 + 32 * hf + j], h4[j], l0);
            l1 = fmaf(small[S_W5 + 64 + 32 * hf + j], h4[j], l1);
        }
        if (hf == 1) { part[0][r] = l0; part[1][r] = l1; }
        tc::fence_before_sync();                                       // accumulator reads are done before the next tile's MMAs
        tc::named_bar_sync(1, THREADS);
        if (hf == 0 && valid) {
            l0 += part[0][r] + small[S_B5];
            l1 += part[1][r] + small[S_B5 + 1];
            const float mx = fmaxf(l0, l1);
            const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
            const float inv = 1.f / (e0 + e1);
            const float a0 = e0 * inv, a1 = e1 * inv;
            a.fused[s] = a0 * x0 + a1 * x1;                            // decoder.py:256-257
            a.w_attn[s] = a1;
            if (save) { a.tc_soft[(size_t)tile * 256 + r] = a0; a.tc_soft[(size_t)tile * 256 + 128 + r] = a1; }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, 128);
}

// ---------------------------------------------------------------------------------------------------------------
// backward, kernel 1 of 2: the delta chain (decoder.py:252-257 reversed)
// ---------------------------------------------------------------------------------------------------------------
// d4 = (W5^T dl) * m4,  d3 = (d4 W4) * m3,  d2 = (d3 W3) * m2,  d1 = (d2 W2) * m1,  dx = d1 W1 -- the three wide products
// on the tensor cores with the forward's weight image read transposed.  With WP (parameter gradients wanted) the same
// delta planes also feed  [1 x0 x1]^T d_i  (bias gradients in accumulator row 0, pts_linears.0's weight gradient in
// rows 1-2 of the d1 product), accumulated in TMEM over all of the CTA's tiles and flushed once; and d2, d3, d4, dl go
// out as planes for the weight-gradient kernel.
constexpr uint32_t T_D = 0, T_X1 = 128, T_X2 = 192, T_X3 = 320, T_X4 = 448;      // TMEM columns (512 with WP, 128 without)

template <int NH>
__device__ __forceinline__ void bwd_epilogue(uint32_t tl, unsigned char* act, int r, int col0, const uint32_t* mask, float* keep) {
    constexpr int F = 2 * NH;
#pragma unroll
    for (int b = 0; b < NH / 32; ++b) {
        float v[32];
        tc::tmem_ld32(tl + col0 + 32 * b, v);
        tc::wait_ld();
        const uint32_t m = mask[b];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            v[j] = ((m >> j) & 1u) ? v[j] : 0.f;
            if (keep) keep[32 * b + j] = v[j];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) put8<false>(act, F, r, col0 + 32 * b + 8 * c, v + 8 * c);
    }
}

// [v0 v1 v2 0 0 0 0 0] of row r into an 8-feature plane pair (hi at base, mid at base + 2048)
__device__ __forceinline__ void put_small8(unsigned char* base, int r, float v0, float v1, float v2) {
    const uint32_t h0 = pl::pack_bf16x2(v0, v1), h1 = pl::pack_bf16x2(v2, 0.f);
    const uint32_t m0 = pl::pack_bf16x2(v0 - __uint_as_float(h0 << 16), v1 - __uint_as_float(h0 & 0xffff0000u));
    const uint32_t m1 = pl::pack_bf16x2(v2 - __uint_as_float(h1 << 16), 0.f);
    *reinterpret_cast<uint4*>(base + r * 16) = make_uint4(h0, h1, 0u, 0u);
    *reinterpret_cast<uint4*>(base + 2048 + r * 16) = make_uint4(m0, m1, 0u, 0u);
}

template <int MODE, bool WP>
__global__ void __launch_bounds__(THREADS, 1) k_attn_dh_tc(const AttArgs a) {
    using PT = typename PtT<MODE>::T;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bar_w, bar_mma;
    __shared__ uint32_t tmem_base_s;
    __shared__ float part[2][128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = tid & 127, hf = tid >> 7;
    const int count = *a.band_count;
    const int tiles = (count + 127) >> 7;
    if ((int)blockIdx.x >= tiles) return;
    float* small = reinterpret_cast<float*>(sm + SM_SMALL);
    unsigned char* act = sm + SM_ACT;
    unsigned char* x8 = sm + SM_X8;
    unsigned char* dl8 = sm + SM_DL8;
    const uint32_t act_s = tc::smem_u32(act), img_s = tc::smem_u32(sm + SM_IMG), x8_s = tc::smem_u32(x8), dl8_s = tc::smem_u32(dl8);
    constexpr uint32_t TCOLS = WP ? 512 : 128;

    for (uint32_t i = tid; i < (SM_BYTES_BWD - SM_X8) / 16; i += THREADS) reinterpret_cast<uint4*>(x8)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, TCOLS); tc::tmem_relinquish(); }
    if (tid == 0) {
        tc::mbar_init(&bar_w, 1); tc::mbar_init(&bar_mma, 1); tc::fence_mbar_init();
        tc::mbar_arrive_expect_tx(&bar_w, AttImg::BYTES);
        tc::bulk_copy_g2s(img_s, a.params + OFF_ATT_IMG, AttImg::BYTES, &bar_w);
    }
    load_small(small, a.params + OFF_ATT);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tb = tmem_base_s;
    const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16);
    const bool issuer = tid == 0;
    const BoxT<PT> tbox = make_box<PT>(a.tsdf_bnds.b);
    float* g = WP ? a.gparams + OFF_ATT : nullptr;
    uint32_t parity = 0;
    bool weights_ready = false;
    uint32_t xacc = 0;                                                  // 0 until the first tile has initialised the X accumulators

    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int pos = tile * 128 + r;
        const bool valid = pos < count;
        const int s = a.band_list[valid ? pos : count - 1];
        const float g_out = valid ? a.d_fused[s] : 0.f;
        const float g_w = (valid && a.g_w) ? a.g_w[s] : 0.f;
        const bool any = (g_out != 0.f) || (g_w != 0.f);
        const bool tile_any = __syncthreads_or(any) != 0;               // THREADS == blockDim: the whole CTA
        if (tid == 0 && a.tc_tileflag) a.tc_tileflag[tile] = tile_any ? 1 : 0;
        if (!tile_any) continue;                                        // d_low / d_high keep their seeds (zero on band samples)
        const float x0 = a.high[s] + a.low[s];
        const AttLogit lg = tsdf_logit(a.tsdf[s]);
        const float x1 = lg.inv;
        const float a0 = a.tc_soft[(size_t)tile * 256 + r], a1 = a.tc_soft[(size_t)tile * 256 + 128 + r];
        const float da0 = g_out * x0, da1 = g_out * x1 + g_w;
        const float dot = a0 * da0 + a1 * da1;
        const float dl0 = a0 * (da0 - dot), dl1 = a1 * (da1 - dot);      // softmax backward
        float dx0 = g_out * a0, dx1 = g_out * a1;
        const uint32_t* gm = a.tc_masks + (size_t)tile * MASK_WORDS * 128;
        const uint32_t m1 = gm[(0 + hf) * 128 + r];
        const uint32_t m2[2] = {gm[(2 + 2 * hf) * 128 + r], gm[(3 + 2 * hf) * 128 + r]};
        const uint32_t m3[2] = {gm[(6 + 2 * hf) * 128 + r], gm[(7 + 2 * hf) * 128 + r]};
        const uint32_t m4 = gm[(10 + hf) * 128 + r];
        {   // output_linear backward (2 -> 64) through pts_linears.3's ReLU: this thread's 32 units
            float d[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                const int j = 32 * hf + u;
                const float gv = small[S_W5 + j] * dl0 + small[S_W5 + 64 + j] * dl1;
                d[u] = ((m4 >> u) & 1u) ? gv : 0.f;
            }
            if (WP) {                                                  // the previous tile's stores have read the planes
                if (issuer) bulk_wait_read0();
                tc::named_bar_sync(1, THREADS);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) put8<false>(act, 64, r, 32 * hf + 8 * c, d + 8 * c);
            if (WP && hf == 0) {
                put_small8(x8, r, 1.f, x0, x1);
                put_small8(dl8, r, dl0, dl1, 0.f);
                float s0 = dl0, s1 = dl1;                               // output_linear.bias gradient
                for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
                if (lane == 0) { atomicAdd(g + AttOff::B5, s0); atomicAdd(g + AttOff::B5 + 1, s1); }
            }
        }
        unsigned char* gsave = WP ? a.tc_dplanes + (size_t)tile * G_DTILE : nullptr;
        float d1[32];
#pragma unroll 1
        for (int step = 0; step < 3; ++step) {                          // through pts_linears.3 (W4), .2 (W3), .1 (W2)
            tc::fence_proxy_async();
            tc::fence_before_sync();
            tc::named_bar_sync(1, THREADS);
            const int J = step == 0 ? 64 : 128, NIN = step == 2 ? 64 : 128;
            if (WP && issuer) {                                        // bulk groups are per-thread state: always thread 0
                const uint32_t goff = step == 0 ? G_D4 : step == 1 ? G_D3 : G_D2, pbytes = 256u * J;
                bulk_store_s2g(gsave + goff, act_s, pbytes);
                bulk_store_s2g(gsave + goff + pbytes, act_s + ACT_PLANE, pbytes);
                if (step == 0) { bulk_store_s2g(gsave + G_DL, dl8_s, 2048u); bulk_store_s2g(gsave + G_DL + 2048u, dl8_s + 2048u, 2048u); }
                bulk_commit();
            }
            if (warp == 0 && !weights_ready) { tc::mbar_wait(&bar_w, 0); weights_ready = true; }
            if (warp == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                const uint32_t wh = img_s + (step == 0 ? AttImg::W4H : step == 1 ? AttImg::W3H : AttImg::W2H);
                const uint32_t wm = img_s + (step == 0 ? AttImg::W4M : step == 1 ? AttImg::W3M : AttImg::W2M);
                mma_layer_bwd(tb + T_D, act_s, J, wh, wm, NIN);
                if (WP) {                                              // [1 x0 x1]^T d over the tile's 128 samples
                    const uint32_t xcol = step == 0 ? T_X4 : step == 1 ? T_X3 : T_X2;
                    const uint32_t idesc = tc::idesc_bf16_mn(64, J);
                    for (uint32_t ks = 0; ks < 8; ++ks)
                        tc::mma_planes(tb + xcol, x8_s + ks * 256u, x8_s + 2048u + ks * 256u, 128u, act_s + ks * 32u * J,
                                       act_s + ACT_PLANE + ks * 32u * J, 16u * J, idesc, xacc != 0 || ks > 0);
                }
                tc::mma_commit(&bar_mma);
            }
            tc::mbar_wait(&bar_mma, parity); parity ^= 1;
            tc::fence_after_sync();
            if (WP) {
                if (issuer) bulk_wait_read0();
                tc::named_bar_sync(1, THREADS);
            }
            if (step == 0) bwd_epilogue<64>(tl + T_D, act, r, 64 * hf, m3, nullptr);
            else if (step == 1) bwd_epilogue<64>(tl + T_D, act, r, 64 * hf, m2, nullptr);
            else bwd_epilogue<32>(tl + T_D, act, r, 32 * hf, &m1, d1);
        }
        if (WP) {                                                      // pts_linears.0: weight + bias gradient from the d1 planes
            tc::fence_proxy_async();
            tc::fence_before_sync();
            tc::named_bar_sync(1, THREADS);
            if (warp == 0 && tc::elect_one()) {
                tc::fence_after_sync();
                const uint32_t idesc = tc::idesc_bf16_mn(64, 64);
                for (uint32_t ks = 0; ks < 8; ++ks)
                    tc::mma_planes(tb + T_X1, x8_s + ks * 256u, x8_s + 2048u + ks * 256u, 128u, act_s + ks * 32u * 64u,
                                   act_s + ACT_PLANE + ks * 32u * 64u, 16u * 64u, idesc, xacc != 0 || ks > 0);
                tc::mma_commit(&bar_mma);
            }
            tc::mbar_wait(&bar_mma, parity); parity ^= 1;
            tc::fence_after_sync();
            xacc = 1;
        }
        // input gradient (pts_linears.0, 64 -> 2) on the CUDA cores
        float gx0 = 0.f, gx1 = 0.f;
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            gx0 = fmaf(small[S_W1 + 32 * hf + u], d1[u], gx0);
            gx1 = fmaf(small[S_W1 + 64 + 32 * hf + u], d1[u], gx1);
        }
        if (hf == 1) { part[0][r] = gx0; part[1][r] = gx1; }
        tc::fence_before_sync();
        tc::named_bar_sync(1, THREADS);
        if (hf == 0 && valid) {
            dx0 += gx0 + part[0][r];
            dx1 += gx1 + part[1][r];
            a.d_low[s] = dx0;                                          // low_occ enters only through high + low on band points
            a.d_high[s] = dx0;
            if (a.dp) {
                const float dt = dx1 * lg.dinv_dtsdf;                  // gradient into the TSDF value gathered at p (decoder.py:241)
                if (dt != 0.f) {
                    PT p[3];
                    load_point<MODE>(a.ps, s, p);
                    float xn[3];
                    normalize(p, tbox, xn);
                    const Tri t = tri_setup(xn, a.vol.W, a.vol.H, a.vol.D);
                    float gn[3];
                    tsdf_interp_grad(a.vol.data, a.vol.sD, a.vol.sH, a.vol.sW, a.vol.W, a.vol.H, a.vol.D, t, gn);
                    for (int k = 0; k < 3; ++k)
                        a.dp[3 * (int64_t)s + k] += dt * gn[k] * (float)(2.0 / (a.tsdf_bnds.b[k][1] - a.tsdf_bnds.b[k][0]));
                }
            }
        }
    }
    if (WP) {
        if (issuer) bulk_wait0();
        // flush: accumulator rows 0..2 of the M = 64 products sit in TMEM lanes 0..2 (warp 0's quarter)
        if (warp == 0 && xacc) {
            tc::fence_after_sync();
            float v[32];
            auto flush = [&](uint32_t col, int n, int boff, bool w1) {
                for (int b = 0; b < n / 32; ++b) {
                    tc::tmem_ld32(tb + col + 32 * b, v);               // all lanes take part; lanes 0..2 hold rows 0..2
                    tc::wait_ld();
                    if (lane == 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(g + boff + 32 * b + j, v[j]);
                    } else if (w1 && lane <= 2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(g + AttOff::W1 + (lane - 1) * 64 + 32 * b + j, v[j]);
                    }
                }
            };
            flush(T_X1, 64, AttOff::B1, true);
            flush(T_X2, 128, AttOff::B2, false);
            flush(T_X3, 128, AttOff::B3, false);
            flush(T_X4, 64, AttOff::B4, false);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, TCOLS);
}

// ---------------------------------------------------------------------------------------------------------------
// backward, kernel 2 of 2: weight gradients of the wide layers and of output_linear as outer products over samples
// ---------------------------------------------------------------------------------------------------------------
// Same machinery as k_decoder_dw_tc: one thread streams 32-sample slices of the saved planes into a shared-memory ring
// with bulk copies, one thread issues SS-mode MMAs whose accumulators stay in TMEM for the CTA's lifetime, one flush.
//   pts_linears.1: h1^T d2 (M 64,  N 128)     pts_linears.2: h2^T d3 (M 128, N 128)
//   pts_linears.3: h3^T d4 (M 128, N 64)      output_linear: h4^T dl (M 64,  N 8)
namespace adw {
constexpr int NST = 2, THREADS = 192, MAX_TILES = 1024;
__host__ __device__ constexpr uint32_t piece(int F) { return 64u * (uint32_t)F; }        // one plane, 32 samples
constexpr uint32_t O_H1 = 0, O_H2 = O_H1 + 2 * piece(64), O_H3 = O_H2 + 2 * piece(128), O_H4 = O_H3 + 2 * piece(128);
constexpr uint32_t O_D2 = O_H4 + 2 * piece(64), O_D3 = O_D2 + 2 * piece(128), O_D4 = O_D3 + 2 * piece(128), O_DL = O_D4 + 2 * piece(64);
constexpr uint32_t STAGE = O_DL + 2 * piece(8);           // 91136 bytes
constexpr uint32_t T_W2 = 0, T_W3 = 128, T_W4 = 256, T_W5 = 320;
}  // namespace adw

__global__ void __launch_bounds__(adw::THREADS, 1) k_attn_dw_tc(const AttArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full_bar[adw::NST], empty_bar[adw::NST], done_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int tile_list[adw::MAX_TILES];
    __shared__ int n_active_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int count = *a.band_count;
    const int tiles = (count + 127) >> 7;
    if (tid == 0) n_active_s = 0;
    __syncthreads();
    if (warp == 0) {                                                   // this CTA's tiles that carry gradient, in order
        int cnt = 0;
        for (int t0 = blockIdx.x; t0 < tiles && cnt < adw::MAX_TILES; t0 += gridDim.x * 32) {
            const int t = t0 + lane * gridDim.x;
            const bool on = t < tiles && a.tc_tileflag[t] != 0;
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) { const int at = cnt + __popc(m & ((1u << lane) - 1)); if (at < adw::MAX_TILES) tile_list[at] = t; }
            cnt += __popc(m);
        }
        if (lane == 0) n_active_s = cnt < adw::MAX_TILES ? cnt : adw::MAX_TILES;
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
    if (tid == 0) {
        for (int i = 0; i < adw::NST; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
        tc::mbar_init(&done_bar, 1);
        tc::fence_mbar_init();
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tb = tmem_base_s;
    const int n_active = n_active_s;
    const uint32_t sm_base = tc::smem_u32(sm);
    const uint32_t total = 4u * (uint32_t)n_active;                      // stages: (tile, quarter)

    if (n_active > 0) {
        if (warp == 0) {
            if (lane == 0) {                                           // producer
                for (uint32_t q = 0; q < total; ++q) {
                    const int st = q % adw::NST;
                    if (q >= adw::NST) tc::mbar_wait(&empty_bar[st], ((q / adw::NST) - 1) & 1);
                    const size_t tile = (size_t)tile_list[q >> 2];
                    const uint32_t quarter = q & 3;
                    const uint32_t dst = sm_base + st * adw::STAGE;
                    const unsigned char* gh = a.tc_planes + tile * G_TILE;
                    const unsigned char* gd = a.tc_dplanes + tile * G_DTILE;
                    tc::mbar_arrive_expect_tx(&full_bar[st], adw::STAGE);
                    auto copy = [&](const unsigned char* garr, int F, uint32_t off) {
                        for (uint32_t plane = 0; plane < 2; ++plane)
                            tc::bulk_copy_g2s(dst + off + plane * adw::piece(F), garr + plane * (256u * F) + quarter * adw::piece(F), adw::piece(F), &full_bar[st]);
                    };
                    copy(gh + G_H1, 64, adw::O_H1); copy(gh + G_H2, 128, adw::O_H2); copy(gh + G_H3, 128, adw::O_H3); copy(gh + G_H4, 64, adw::O_H4);
                    copy(gd + G_D2, 128, adw::O_D2); copy(gd + G_D3, 128, adw::O_D3); copy(gd + G_D4, 64, adw::O_D4); copy(gd + G_DL, 8, adw::O_DL);
                }
            }
        } else if (warp == 1) {
            if (tc::elect_one()) {                                     // MMA issuer
                for (uint32_t q = 0; q < total; ++q) {
                    const int st = q % adw::NST;
                    tc::mbar_wait(&full_bar[st], (q / adw::NST) & 1);
                    tc::fence_after_sync();
                    const uint32_t s0 = sm_base + st * adw::STAGE;
                    auto prod = [&](uint32_t d_col, int M, int N, uint32_t a_off, int FA, uint32_t b_off, int FB) {
                        const uint32_t idesc = tc::idesc_bf16_mn(M, N);
                        const uint32_t ga = 16u * FA, gb = 16u * FB;
                        for (uint32_t ks = 0; ks < 2; ++ks) {
                            const uint32_t ah = s0 + a_off + ks * 2 * ga, bh = s0 + b_off + ks * 2 * gb;
                            tc::mma_planes(tb + d_col, ah, ah + adw::piece(FA), ga, bh, bh + adw::piece(FB), gb, idesc, q > 0 || ks > 0);
                        }
                    };
                    prod(adw::T_W2, 64, 128, adw::O_H1, 64, adw::O_D2, 128);
                    prod(adw::T_W3, 128, 128, adw::O_H2, 128, adw::O_D3, 128);
                    prod(adw::T_W4, 128, 64, adw::O_H3, 128, adw::O_D4, 64);
                    prod(adw::T_W5, 64, 8, adw::O_H4, 64, adw::O_DL, 8);
                    tc::mma_commit(&empty_bar[st]);
                }
                tc::mma_commit(&done_bar);
            }
        } else {                                                       // flush (warps 2..5 = TMEM lane quarters 2, 3, 0, 1)
            tc::mbar_wait(&done_bar, 0);
            tc::fence_after_sync();
            float* g = a.gparams + OFF_ATT;
            const int qtr = warp & 3;
            const uint32_t tl = tb + ((uint32_t)(qtr * 32) << 16);
            const int row128 = qtr * 32 + lane;
            const int row64 = lane < 16 ? qtr * 16 + lane : -1;          // M = 64: row m in lane (m/16)*32 + m%16
            float d[32];
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {                               // pts_linears.1: [k 64][j 128]
                tc::tmem_ld32(tl + adw::T_W2 + 32 * b, d);
                tc::wait_ld();
                if (row64 >= 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(g + AttOff::W2 + row64 * 128 + 32 * b + j, d[j]);
                }
            }
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {                               // pts_linears.2: [k 128][j 128]
                tc::tmem_ld32(tl + adw::T_W3 + 32 * b, d);
                tc::wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(g + AttOff::W3 + row128 * 128 + 32 * b + j, d[j]);
            }
#pragma unroll 1
            for (int b = 0; b < 2; ++b) {                               // pts_linears.3: [k 128][j 64]
                tc::tmem_ld32(tl + adw::T_W4 + 32 * b, d);
                tc::wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(g + AttOff::W4 + row128 * 64 + 32 * b + j, d[j]);
            }
            {                                                           // output_linear: rows = h4 feature j, columns = logit o
                uint32_t v[8];
                tc::tmem_ld8(tl + adw::T_W5, v);
                tc::wait_ld();
                if (row64 >= 0) {
                    atomicAdd(g + AttOff::W5 + row64, __uint_as_float(v[0]));
                    atomicAdd(g + AttOff::W5 + 64 + row64, __uint_as_float(v[1]));
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tb, 512);
}

int launch_attn_fwd_tc(const AttArgs& aa, int64_t n, cudaStream_t st) {
    const size_t smem = SM_BYTES + 1024;
    int rc = set_smem(k_attn_fwd_tc, smem);
    if (rc) return rc;
    const int blocks = grid_for((n + 127) / 128, 1, num_sms());
    DFP_TIMED("k_attn_fwd_tc", st, k_attn_fwd_tc<<<blocks, THREADS, smem, st>>>(aa));
    DFP_LAUNCH_CHECK("k_attn_fwd_tc");
    return DFP_OK;
}

template <int MODE>
int launch_attn_bwd_tc(const AttArgs& aa, int64_t n, cudaStream_t st) {
    const size_t smem = SM_BYTES_BWD + 1024;
    const int blocks = grid_for((n + 127) / 128, 1, num_sms());
    int rc;
    if (aa.gparams) {
        if ((rc = set_smem(k_attn_dh_tc<MODE, true>, smem))) return rc;
        DFP_TIMED("k_attn_dh_tc", st, k_attn_dh_tc<MODE, true><<<blocks, THREADS, smem, st>>>(aa));
        DFP_LAUNCH_CHECK("k_attn_dh_tc");
        const size_t smem2 = adw::NST * adw::STAGE + 1024;
        if ((rc = set_smem(k_attn_dw_tc, smem2))) return rc;
        int blocks2 = blocks;
        const int64_t min_blocks = ((n + 127) / 128 + adw::MAX_TILES - 1) / adw::MAX_TILES;
        if (blocks2 < min_blocks) blocks2 = (int)min_blocks;
        DFP_TIMED("k_attn_dw_tc", st, k_attn_dw_tc<<<blocks2, adw::THREADS, smem2, st>>>(aa));
        DFP_LAUNCH_CHECK("k_attn_dw_tc");
    } else {
        if ((rc = set_smem(k_attn_dh_tc<MODE, false>, smem))) return rc;
        DFP_TIMED("k_attn_dh_tc", st, k_attn_dh_tc<MODE, false><<<blocks, THREADS, smem, st>>>(aa));
        DFP_LAUNCH_CHECK("k_attn_dh_tc");
    }
    return DFP_OK;
}
template int launch_attn_bwd_tc<0>(const AttArgs&, int64_t, cudaStream_t);
template int launch_attn_bwd_tc<1>(const AttArgs&, int64_t, cudaStream_t);
template int launch_attn_bwd_tc<2>(const AttArgs&, int64_t, cudaStream_t);

}  // namespace dfp
