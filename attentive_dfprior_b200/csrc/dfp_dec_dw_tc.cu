// Tensor-core decoder backward, kernel 2 of 2: every parameter gradient of one decoder as outer products over
// samples,  dW = A^T * Delta,  on the tcgen05 tensor cores, organised as a streaming GEMM pipeline.
//
// The decoder's parameter tensors are cut into "products" of one 32-feature A block times one 32-column delta block
// (pts_linears.1..4, output_linear, fc_c.i per 32 features, the three 32-wide chunks of the embedding against the
// deltas of blocks 0 and 3, embedder._B).  The kernel runs in two passes (pass 0: pts_linears.1..4, output_linear,
// fc_c; pass 1: the embedding-side products) so that the accumulators of ALL products of a pass fit the 512 TMEM
// columns and stay resident for the CTA's lifetime; a CTA then walks  tile (outer) x product (inner): every operand
// block comes from HBM once per tile and its re-stagings for other products hit L2.  Sixteen
// producer warps load the saved activations / deltas (feature-major arrays written by the forward and by kernel 1),
// split them to TF32 hi/lo and store them as row-per-sample swizzled tiles (the MN-major operand layout of
// dfp_tc.cuh) into a 3-stage shared-memory ring; a ninth warp issues the SS-mode tcgen05.mma chains (K = 128 samples
// per stage) and hands the stage back through tcgen05.commit.  (An earlier product-outer variant re-read every operand
// from HBM for each product: 13.9 GB per decoder at 65 536 rays, ncu profiles/r01.)  The MMAs are M = 64: rows 0..31 are the product, row 32 (from a
// constant-one feature block addressed as the A operand's second MN block) the bias sums.  The accumulator (32 TMEM
// columns) lives in one of two TMEM sets, so flushing product p (one atomic pass per CTA
// and product) overlaps the MMAs of product p+1.  No per-sample or per-tile atomics.
#include "dfp_device.cuh"
#include "dfp_tc.cuh"

namespace dfp {

namespace dw {
constexpr int BLK = 4096;                     // one [128][32] f32 tile (16 KB)
constexpr uint32_t BLK_BYTES = BLK * 4;
constexpr int NSTAGE = 2;
constexpr int STAGE = 5 * BLK;                // A_hi | A_lo | B0 | B1 | B2  (an item = one A block against up to 3 delta blocks)
constexpr int ONES = NSTAGE * STAGE;          // constant-one A block (column 0 == 1): second MN block of every hi operand
constexpr int ZERO = ONES + BLK;              // zero block: second MN block of every lo operand
constexpr int FLOATS = ZERO + BLK;
constexpr int MAX_TILES = 1024;               // active tiles of one CTA
constexpr int PRODUCERS = 256;               // threads staging one item (128 A rows + 128 B rows)
constexpr int GROUPS = 2;                     // producer groups working on alternating items
}  // namespace dw

struct DwProduct {
    int kind;      // 0: pts_linears.i  1: output_linear  2: fc_c.i (feature half hf)  3: embedding chunk cb x delta of block `which`  4: embedder._B chunk cb
    int i, hf, cb, which;
};

template <int CD>
__device__ __forceinline__ DwProduct dw_product(int p) {
    DwProduct d = {0, 0, 0, 0, 0};
    constexpr int NC = 5 * (CD / 32);
    if (p < 4) { d.kind = 0; d.i = p + 1; }
    else if (p == 4) { d.kind = 1; }
    else if (p < 5 + NC) { d.kind = 2; d.i = (p - 5) % 5; d.hf = (p - 5) / 5; }
    else if (p < 5 + NC + 6) { d.kind = 3; d.cb = (p - 5 - NC) % 3; d.which = ((p - 5 - NC) / 3) ? 3 : 0; }
    else { d.kind = 4; d.cb = p - 5 - NC - 6; }
    return d;
}
template <int CD> __host__ __device__ constexpr int dw_num_products() { return 5 + 5 * (CD / 32) + 6 + 3; }

// An item = up to three products that share their A block; its MMAs are N = 32 * nprod wide (the transposed SS-mode
// TF32 MMA has a high fixed cost, ~27 ns measured, so fewer and wider MMAs win).  slot = first 32-column TMEM slot.
struct DwItem { int nprod; int prod[3]; int slot; };
template <int CD, int PASS> __host__ __device__ constexpr int dw_num_items() { return PASS == 0 ? 5 + 2 * (CD / 32) : 4; }
template <int CD, int PASS>
__device__ __forceinline__ DwItem dw_item(int it) {
    DwItem d = {1, {0, 0, 0}, 0};
    if (PASS == 0) {
        if (it < 5) { d.prod[0] = it; d.slot = it; return d; }                 // pts_linears.1..4, output_linear
        const int hf = (it - 5) >> 1, second = (it - 5) & 1;                   // fc_c: (c half, [dh_0 dh_1 dh_2]) and (c half, [dh_3 dh_4])
        const int base = 5 + hf * 5 + (second ? 3 : 0);
        d.nprod = second ? 2 : 3;
        d.prod[0] = base; d.prod[1] = base + 1; d.prod[2] = base + 2;
        d.slot = base;
        return d;
    }
    constexpr int B0 = 5 + 5 * (CD / 32);
    if (it < 3) { d.nprod = 2; d.prod[0] = B0 + it; d.prod[1] = B0 + 3 + it; d.slot = 2 * it; return d; }   // sin chunk x [dacc_0 | dacc_3]
    d.nprod = 3; d.prod[0] = B0 + 6; d.prod[1] = B0 + 7; d.prod[2] = B0 + 8; d.slot = 6;                        // point x [darg chunks]
    return d;
}

// write 32 values of sample-row `row` into a swizzled [128][32] tile, as eight 16-byte stores
__device__ __forceinline__ void tile_store_row(float* tile, int row, const float (&v)[32]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float* dst = tile + row * 32 + ((((q >> 1) ^ row) & 3) << 3) + ((q & 1) << 2);
        *reinterpret_cast<float4*>(dst) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}
// 32 consecutive features of a tile-major F-feature array (tm_index) for sample s (zeros when !valid)
__device__ __forceinline__ void load_features(const float* __restrict__ base, int F, int64_t s, int f0, bool valid, float (&v)[32]) {
    const float* p = base + tm_index(s, f0, F);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = valid ? __ldg(p + j * 128) : 0.f;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

template <int DEC, int PASS>
__global__ void __launch_bounds__(dw::GROUPS * dw::PRODUCERS + 32, 1) k_decoder_dw_tc(const DecArgs a) {
    constexpr int CD = (DEC == 1) ? 64 : 32;
    constexpr int NOUT = (DEC == 2) ? 4 : 1;
    constexpr int NI = dw_num_items<CD, PASS>();                     // items per tile: 7 / 9 (pass 0), 4 (pass 1)
    using L = DecOff<CD>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sm = reinterpret_cast<float*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
    __shared__ __align__(8) uint64_t full_bar[dw::NSTAGE], empty_bar[dw::NSTAGE], done_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int tile_list[dw::MAX_TILES];
    __shared__ int n_active_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n = dec_items(a);
    const int64_t tiles = (n + 127) / 128;

    for (int i = tid; i < dw::FLOATS; i += blockDim.x) sm[i] = 0.f;
    if (tid == 0) n_active_s = 0;
    __syncthreads();
    if (tid < 128) sm[dw::ONES + tc::mn_tile_index(tid, 0)] = 1.0f;          // ZERO block stays zero
    // the CTA's tiles that carry any upstream gradient (kernel 1 wrote scratch only for those), in order
    if (warp == 0) {
        int cnt = 0;
        for (int64_t tile = blockIdx.x; tile < tiles && cnt < dw::MAX_TILES; tile += gridDim.x) {
            bool any = false;
            for (int r = lane; r < 128; r += 32) {
                const int64_t ii = tile * 128 + r;
                if (ii < n) {
                    const int64_t s = dec_sample(a, ii);
                    if (NOUT == 1) any = any || a.dout[s] != 0.f;
                    else { const float4 v = *reinterpret_cast<const float4*>(a.dout + 4 * s); any = any || v.x != 0.f || v.y != 0.f || v.z != 0.f; }
                }
            }
            if (__any_sync(0xffffffffu, any)) { if (lane == 0) tile_list[cnt] = (int)tile; ++cnt; }
        }
        if (lane == 0) n_active_s = cnt;
    }
    if (warp == 3) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
    if (tid == 0) {
        for (int i = 0; i < dw::NSTAGE; ++i) { tc::mbar_init(&full_bar[i], dw::PRODUCERS); tc::mbar_init(&empty_bar[i], 1); }
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

    if (n_active > 0) {
        if (warp == dw::GROUPS * dw::PRODUCERS / 32) {
            // ===================== MMA issuer (one lane) =====================
            if (lane == 0) {
                const uint32_t total = (uint32_t)NI * (uint32_t)n_active;
                for (uint32_t q = 0; q < total; ++q) {                   // q = tile * NI + item (tile-outer)
                    const int st = q % dw::NSTAGE;
                    const DwItem it = dw_item<CD, PASS>((int)(q % NI));
                    tc::mbar_wait(&full_bar[st], (q / dw::NSTAGE) & 1);
                    tc::fence_after_sync();
                    const uint32_t sa = sm_base + st * dw::STAGE * 4;
                    const uint32_t a_hi = sa, a_lo = sa + dw::BLK_BYTES, b = sa + 2 * dw::BLK_BYTES;
                    // M = 64: rows 0..31 = the products, row 32 = column sums of the delta tiles (constant-one block)
                    tc::mma_outer64(tb + it.slot * 32, a_hi, sm_base + dw::ONES * 4 - a_hi, a_lo, sm_base + dw::ZERO * 4 - a_lo, b,
                                    32 * it.nprod, q >= (uint32_t)NI);
                    tc::mma_commit(&empty_bar[st]);                      // stage reusable once these MMAs have read it
                }
                tc::mma_commit(&done_bar);                               // everything accumulated
            }
        } else {
            // ===================== producers (256 threads) + flush (warp 0) =====================
            const int pgrp = tid / dw::PRODUCERS;                  // this thread's producer group
            const int half = (tid % dw::PRODUCERS) >> 7;           // 0: stages A (and delta block 2), 1: stages delta blocks 0 and 1
            const int row = tid & 127;
            float* g = a.gparams + (DEC == 0 ? OFF_LOW : (DEC == 1 ? OFF_HIGH : OFF_COLOR));
            // the A block (32 features) of product p for sample s
            auto load_A = [&](const DwProduct& pr, int64_t s, int64_t sid, bool valid, float (&v)[32]) {
                if (pr.kind == 0) load_features(a.act_h, 160, s, (pr.i - 1) * 32, valid, v);
                else if (pr.kind == 1) load_features(a.act_h, 160, s, 4 * 32, valid, v);
                else if (pr.kind == 2) load_features(a.act_c, CD, s, pr.hf * 32, valid, v);
                else if (pr.kind == 3) load_features(a.sc_e, 96, s, pr.cb * 32, valid, v);
                else {
                    float pw[3] = {0.f, 0.f, 0.f};
                    if (valid) {
                        if (a.ps.pts == nullptr) {
                            const int64_t r = sid / a.ps.S;
                            double pq[3];
                            const float o[3] = {a.ps.rays_o[3 * r], a.ps.rays_o[3 * r + 1], a.ps.rays_o[3 * r + 2]};
                            const float d[3] = {a.ps.rays_d[3 * r], a.ps.rays_d[3 * r + 1], a.ps.rays_d[3 * r + 2]};
                            ray_point(o, d, a.ps.z[sid], pq);
                            pw[0] = (float)pq[0]; pw[1] = (float)pq[1]; pw[2] = (float)pq[2];
                        } else if (a.ps.pts_f64) {
                            const double* pq = reinterpret_cast<const double*>(a.ps.pts) + 3 * sid;
                            pw[0] = (float)pq[0]; pw[1] = (float)pq[1]; pw[2] = (float)pq[2];
                        } else {
                            const float* pq = reinterpret_cast<const float*>(a.ps.pts) + 3 * sid;
                            pw[0] = pq[0]; pw[1] = pq[1]; pw[2] = pq[2];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = j < 3 ? pw[j] : 0.f;
                }
            };
            // the delta block (32 columns) of product p for sample s
            auto load_B = [&](const DwProduct& pr, int64_t s, int64_t sid, bool valid, float (&v)[32]) {
                if (pr.kind == 0 || pr.kind == 3) {
                    const int blk = pr.kind == 0 ? pr.i : pr.which;
                    load_features(a.sc_dh, 160, s, blk * 32, valid, v);
                    const uint32_t m = a.act_m[tm_index(s, blk, 5)];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = ((m >> j) & 1u) ? v[j] : 0.f;
                } else if (pr.kind == 1) {
                    float d4[4] = {0.f, 0.f, 0.f, 0.f};
                    if (valid) {
                        if (NOUT == 1) d4[0] = a.dout[sid];
                        else { const float4 x = *reinterpret_cast<const float4*>(a.dout + 4 * sid); d4[0] = x.x; d4[1] = x.y; d4[2] = x.z; }
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = j < 4 ? d4[j] : 0.f;     // colour: the 4th output carries no gradient
                } else if (pr.kind == 2) {
                    load_features(a.sc_dh, 160, s, pr.i * 32, valid, v);
                } else {
                    load_features(a.sc_darg, 96, s, pr.cb * 32, valid, v);
                }
            };
            // items are numbered q = tile * NI + item (tile-outer); this group stages the items with q % GROUPS == pgrp
            const uint32_t total = (uint32_t)NI * (uint32_t)n_active;
            for (uint32_t q = (uint32_t)pgrp; q < total; q += dw::GROUPS) {
                const int st = q % dw::NSTAGE;
                const DwItem it = dw_item<CD, PASS>((int)(q % NI));
                const int64_t i_raw = (int64_t)tile_list[q / NI] * 128 + row;
                const bool valid = i_raw < n;
                const int64_t s = valid ? i_raw : n - 1;             // list position (saved activations / scratch index)
                const int64_t sid = dec_sample(a, s);                // sample id (points, dout)
                float v0[32], v1[32];
                // half 0: the A block and (if present) delta block 2;  half 1: delta blocks 0 and (if present) 1
                const bool second = half == 0 ? it.nprod > 2 : it.nprod > 1;
                if (half == 0) {
                    load_A(dw_product<CD>(it.prod[0]), s, sid, valid, v0);
                    if (second) load_B(dw_product<CD>(it.prod[2]), s, sid, valid, v1);
                } else {
                    load_B(dw_product<CD>(it.prod[0]), s, sid, valid, v0);
                    if (second) load_B(dw_product<CD>(it.prod[1]), s, sid, valid, v1);
                }
                if (q >= dw::NSTAGE) tc::mbar_wait(&empty_bar[st], ((q / dw::NSTAGE) - 1) & 1);      // MMAs of the previous use are done
                float* stage = sm + st * dw::STAGE;
                if (half == 0) {
                    float hi[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) { hi[j] = tc::tf32_rna(v0[j]); v0[j] = tc::tf32_rna(v0[j] - hi[j]); }
                    tile_store_row(stage, row, hi);
                    tile_store_row(stage + dw::BLK, row, v0);
                    if (second) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v1[j] = tc::tf32_rna(v1[j]);
                        tile_store_row(stage + 4 * dw::BLK, row, v1);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v0[j] = tc::tf32_rna(v0[j]);
                    tile_store_row(stage + 2 * dw::BLK, row, v0);
                    if (second) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v1[j] = tc::tf32_rna(v1[j]);
                        tile_store_row(stage + 3 * dw::BLK, row, v1);
                    }
                }
                tc::fence_proxy_async();
                mbar_arrive(&full_bar[st]);
            }
            // ---- flush (warps 0..2), once per CTA: accumulator row m is TMEM lane (m/16)*32 + m%16; row 32 holds the bias sums
            if (warp < 3) {
                tc::mbar_wait(&done_bar, 0);
                tc::fence_after_sync();
#pragma unroll 1
                for (int pp = 0; pp < NI * 3; ++pp) {
                    const DwItem it = dw_item<CD, PASS>(pp / 3);
                    const int blk = pp % 3;
                    if (blk >= it.nprod) continue;
                    const DwProduct pr = dw_product<CD>(it.prod[blk]);
                    float d[32];
                    tc::tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + (it.slot + blk) * 32, d);
                    tc::wait_ld();
                    if (lane < 16) {
                        const int k = warp * 16 + lane;                 // accumulator row
                        if (k == 32) {                                   // bias sums
                            if (pr.kind == 0) { for (int j = 0; j < 32; ++j) atomicAdd(g + L::BP + pr.i * 32 + j, d[j]); }
                            else if (pr.kind == 1) { for (int j = 0; j < NOUT; ++j) atomicAdd(g + L::BO + j, d[j]); }
                            else if (pr.kind == 2 && pr.hf == 0) { for (int j = 0; j < 32; ++j) atomicAdd(g + L::BC + pr.i * 32 + j, d[j]); }
                            else if (pr.kind == 3 && pr.cb == 0 && pr.which == 0) { for (int j = 0; j < 32; ++j) atomicAdd(g + L::BP + j, d[j]); }
                        } else if (k < 32) {
                            if (pr.kind == 0) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) atomicAdd(g + L::WH + (pr.i - 1) * 1024 + k * 32 + j, d[j]);
                            } else if (pr.kind == 1) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) if (j < NOUT) atomicAdd(g + L::WO + j * 32 + k, d[j]);
                            } else if (pr.kind == 2) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) atomicAdd(g + L::WC + pr.i * CD * 32 + (pr.hf * 32 + k) * 32 + j, d[j]);
                            } else if (pr.kind == 3) {
                                const int kk = pr.cb * 32 + k;
                                if (kk < EMB) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) atomicAdd(g + L::W03 + kk * 64 + (pr.which ? 32 : 0) + j, d[j]);
                                }
                            } else if (k < 3) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) if (pr.cb * 32 + j < EMB) atomicAdd(g + L::B + k * EMB_PAD + pr.cb * 32 + j, d[j]);
                            }
                        }
                    }
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 3) tc::tmem_dealloc(tb, 512);
}

template <int DEC>
int launch_decoder_dw_tc(const DecArgs& a, cudaStream_t st) {
    const size_t smem = dw::FLOATS * sizeof(float) + 1024;
    int rc = set_smem(k_decoder_dw_tc<DEC, 0>, smem);
    if (rc) return rc;
    if ((rc = set_smem(k_decoder_dw_tc<DEC, 1>, smem))) return rc;
    const int64_t tiles = (a.n + 127) / 128;
    int blocks = grid_for(tiles, 1, num_sms());
    const int64_t min_blocks = (tiles + dw::MAX_TILES - 1) / dw::MAX_TILES;      // a CTA lists at most MAX_TILES tiles
    if (blocks < min_blocks) blocks = (int)min_blocks;
    DFP_TIMED("k_decoder_dw_tc", st, k_decoder_dw_tc<DEC, 0><<<blocks, dw::GROUPS * dw::PRODUCERS + 32, smem, st>>>(a));
    DFP_LAUNCH_CHECK("k_decoder_dw_tc");
    DFP_TIMED("k_decoder_dw_tc", st, k_decoder_dw_tc<DEC, 1><<<blocks, dw::GROUPS * dw::PRODUCERS + 32, smem, st>>>(a));
    DFP_LAUNCH_CHECK("k_decoder_dw_tc");
    return DFP_OK;
}
template int launch_decoder_dw_tc<0>(const DecArgs&, cudaStream_t);
template int launch_decoder_dw_tc<1>(const DecArgs&, cudaStream_t);
template int launch_decoder_dw_tc<2>(const DecArgs&, cudaStream_t);

}  // namespace dfp
