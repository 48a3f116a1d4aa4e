// Tensor-core decoder backward, kernel 2 of 2: every parameter gradient of one decoder as outer products over
// samples,  dW = A^T * Delta,  on the tcgen05 tensor cores.
//
// Per 128-sample tile the CTA stages the saved activations (A side, hi/lo TF32 split) and the deltas produced by
// kernel 1 (B side, TF32) as row-per-sample swizzled tiles in shared memory -- the MN-major operand layout probed in
// dfp_tc.cuh -- and issues SS-mode tcgen05.mma with K = the 128 samples.  The accumulators of ALL parameter tensors of
// the decoder (496 TMEM columns) stay resident in TMEM for the CTA's whole lifetime and are flushed to the packed
// gradient with one atomic pass per CTA: no per-sample or per-tile atomics.
//   DW_H  [  0,128): pts_linears.1..4       rows k<32 = weight row k, row 32 = bias (from the constant-one feature block)
//   DW_C  [128,288): fc_c.0..4              rows k<CD, row CD = bias
//   DW_03 [288,384): pts_linears.0 | .3[:, :93], transposed: rows j<64, columns k<96 (column 93 = bias of block 0)
//   DW_B  [384,480): embedder._B            rows r<3, columns k<96
//   DW_O  [480,496): output_linear          rows k<32, columns o<4, row 32 = bias
#include "dfp_device.cuh"
#include "dfp_tc.cuh"

namespace dfp {

namespace dw {
constexpr int BLK = 4096;                 // one [128][32] f32 tile (16 KB)
constexpr int A_HI = 0;                   // 3 blocks; block 2 is the constant-one feature (column 0 == 1)
constexpr int A_LO = 3 * BLK;             // 3 blocks; block 2 stays zero
constexpr int B_HI = 6 * BLK;             // 3 blocks
constexpr int FLOATS = 9 * BLK;           // 144 KB
constexpr uint32_t BLK_BYTES = BLK * 4;
constexpr int T_H = 0, T_C = 128, T_03 = 288, T_B = 384, T_O = 480;
}  // namespace dw

// write 32 values of sample-row `row` into a swizzled [128][32] tile, as eight 16-byte stores
__device__ __forceinline__ void tile_store_row(float* tile, int row, const float (&v)[32]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float* dst = tile + row * 32 + ((((q >> 1) ^ row) & 3) << 3) + ((q & 1) << 2);
        *reinterpret_cast<float4*>(dst) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}
__device__ __forceinline__ void tile_store_row_split(float* tile_hi, float* tile_lo, int row, const float (&v)[32]) {
    float hi[32], lo[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { hi[j] = tc::tf32_rna(v[j]); lo[j] = tc::tf32_rna(v[j] - hi[j]); }
    tile_store_row(tile_hi, row, hi);
    tile_store_row(tile_lo, row, lo);
}
__device__ __forceinline__ void tile_store_row_tf32(float* tile, int row, const float (&v)[32]) {
    float hi[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) hi[j] = tc::tf32_rna(v[j]);
    tile_store_row(tile, row, hi);
}
// 32 consecutive features of a feature-major [f][n] array for sample s (zeros when !valid)
__device__ __forceinline__ void load_features(const float* __restrict__ base, int64_t n, int64_t s, int f0, bool valid, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = valid ? __ldg(base + (int64_t)(f0 + j) * n + s) : 0.f;
}

template <int DEC>
__global__ void __launch_bounds__(256, 1) k_decoder_dw_tc(const DecArgs a) {
    constexpr int CD = (DEC == 1) ? 64 : 32;
    constexpr int NOUT = (DEC == 2) ? 4 : 1;
    using L = DecOff<CD>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sm = reinterpret_cast<float*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool roleA = tid < 128;                       // threads 0..127 stage the A side, 128..255 the B side
    const int row = tid & 127;

    for (int i = tid; i < dw::FLOATS; i += 256) sm[i] = 0.f;
    __syncthreads();
    if (tid < 128) sm[dw::A_HI + 2 * dw::BLK + tc::mn_tile_index(tid, 0)] = 1.0f;        // constant-one feature
    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tb = tmem_base_s;
    uint32_t parity = 0;
    const uint32_t a_hi = tc::smem_u32(sm + dw::A_HI), a_lo = tc::smem_u32(sm + dw::A_LO), b_hi = tc::smem_u32(sm + dw::B_HI);
    float* const Ahi = sm + dw::A_HI;
    float* const Alo = sm + dw::A_LO;
    float* const Bhi = sm + dw::B_HI;
    const int64_t n = a.n;
    const int64_t tiles = (n + 127) / 128;
    bool first = true;                                   // first processed tile overwrites the accumulators

    // one round: everything staged -> make it visible to the tensor core -> one thread issues -> wait for completion
    auto sync_issue = [&](auto issue) {
        tc::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            issue();
            tc::mma_commit(&bar);
        }
        tc::mbar_wait(&bar, parity);
        parity ^= 1;
        tc::fence_after_sync();
    };

    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t s_raw = tile * 128 + row;
        const bool valid = s_raw < n;
        const int64_t s = valid ? s_raw : n - 1;
        // same skip rule as kernel 1: a tile whose upstream gradients are all zero contributes nothing
        bool any = false;
        float dout[4] = {0.f, 0.f, 0.f, 0.f};
        if (valid) {
            if (NOUT == 1) dout[0] = a.dout[s];
            else { const float4 v = *reinterpret_cast<const float4*>(a.dout + 4 * s); dout[0] = v.x; dout[1] = v.y; dout[2] = v.z; }
            any = dout[0] != 0.f || dout[1] != 0.f || dout[2] != 0.f;
        }
        if (!__syncthreads_or(any)) continue;
        const uint32_t acc = first ? 0u : 1u;
        float v[32];

        // ---- pts_linears.1..4: A = input of block i (block 1, followed by the constant-one block), B = mask_i(dh_i)
#pragma unroll 1
        for (int i = 1; i <= 4; ++i) {
            if (roleA) {
                load_features(a.act_h, n, s, (i - 1) * 32, valid, v);
                tile_store_row_split(Ahi + dw::BLK, Alo + dw::BLK, row, v);
            } else {
                load_features(a.sc_dh, n, s, i * 32, valid, v);
                const uint32_t m = a.act_m[(int64_t)i * n + s];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ((m >> j) & 1u) ? v[j] : 0.f;
                tile_store_row_tf32(Bhi, row, v);
            }
            sync_issue([&] {
                tc::mma_outer(tb + dw::T_H + (i - 1) * 32, a_hi + dw::BLK_BYTES, a_lo + dw::BLK_BYTES, dw::BLK_BYTES, b_hi, dw::BLK_BYTES, 32,
                              acc != 0);
            });
        }
        // ---- output_linear: A = input of output_linear, B = upstream gradient (columns 0..NOUT-1)
        {
            if (roleA) {
                load_features(a.act_h, n, s, 4 * 32, valid, v);
                tile_store_row_split(Ahi + dw::BLK, Alo + dw::BLK, row, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (j < 4) ? dout[j] : 0.f;      // colour: the 4th output carries no gradient (dout[3] == 0)
                tile_store_row_tf32(Bhi, row, v);
            }
            sync_issue([&] {
                tc::mma_outer(tb + dw::T_O, a_hi + dw::BLK_BYTES, a_lo + dw::BLK_BYTES, dw::BLK_BYTES, b_hi, dw::BLK_BYTES, 16, acc != 0);
            });
        }
        // ---- fc_c.0..4: A = gathered features (staged once; CD = 64 fills blocks 0 and 1), B = dh_i (before the mask)
        constexpr int CBLK0 = (CD == 64) ? 0 : 1;
        if (roleA) {
            load_features(a.act_c, n, s, 0, valid, v);
            tile_store_row_split(Ahi + CBLK0 * dw::BLK, Alo + CBLK0 * dw::BLK, row, v);
            if (CD == 64) {
                load_features(a.act_c, n, s, 32, valid, v);
                tile_store_row_split(Ahi + dw::BLK, Alo + dw::BLK, row, v);
            }
        }
#pragma unroll 1
        for (int i = 0; i <= 4; ++i) {
            if (!roleA) {
                load_features(a.sc_dh, n, s, i * 32, valid, v);
                tile_store_row_tf32(Bhi, row, v);
            }
            sync_issue([&] {
                tc::mma_outer(tb + dw::T_C + i * 32, a_hi + CBLK0 * dw::BLK_BYTES, a_lo + CBLK0 * dw::BLK_BYTES, dw::BLK_BYTES, b_hi,
                              dw::BLK_BYTES, 32, acc != 0);
            });
        }
        // ---- pts_linears.0 and the embedding half of pts_linears.3, transposed: A = [dacc_0 | dacc_3], B = sin (3 blocks)
        {
            if (roleA) {
                load_features(a.sc_dh, n, s, 0, valid, v);
                const uint32_t m0 = a.act_m[s];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ((m0 >> j) & 1u) ? v[j] : 0.f;
                tile_store_row_split(Ahi, Alo, row, v);
                load_features(a.sc_dh, n, s, 3 * 32, valid, v);
                const uint32_t m3 = a.act_m[(int64_t)3 * n + s];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ((m3 >> j) & 1u) ? v[j] : 0.f;
                tile_store_row_split(Ahi + dw::BLK, Alo + dw::BLK, row, v);
            } else {
#pragma unroll 1
                for (int cb = 0; cb < 3; ++cb) {
                    load_features(a.sc_e, n, s, cb * 32, valid, v);
                    tile_store_row_tf32(Bhi + cb * dw::BLK, row, v);
                }
            }
            sync_issue([&] { tc::mma_outer(tb + dw::T_03, a_hi, a_lo, dw::BLK_BYTES, b_hi, dw::BLK_BYTES, 96, acc != 0); });
        }
        // ---- embedder._B: A = world point (columns 0..2 of block 1), B = d(arg) (3 blocks)
        {
            if (roleA) {
                float pw[3] = {0.f, 0.f, 0.f};
                if (valid) {
                    if (a.ps.pts == nullptr) {
                        const int64_t r = s / a.ps.S;
                        double q[3];
                        const float o[3] = {a.ps.rays_o[3 * r], a.ps.rays_o[3 * r + 1], a.ps.rays_o[3 * r + 2]};
                        const float d[3] = {a.ps.rays_d[3 * r], a.ps.rays_d[3 * r + 1], a.ps.rays_d[3 * r + 2]};
                        ray_point(o, d, a.ps.z[s], q);
                        pw[0] = (float)q[0]; pw[1] = (float)q[1]; pw[2] = (float)q[2];
                    } else if (a.ps.pts_f64) {
                        const double* q = reinterpret_cast<const double*>(a.ps.pts) + 3 * s;
                        pw[0] = (float)q[0]; pw[1] = (float)q[1]; pw[2] = (float)q[2];
                    } else {
                        const float* q = reinterpret_cast<const float*>(a.ps.pts) + 3 * s;
                        pw[0] = q[0]; pw[1] = q[1]; pw[2] = q[2];
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = j < 3 ? pw[j] : 0.f;
                tile_store_row_split(Ahi + dw::BLK, Alo + dw::BLK, row, v);
            } else {
#pragma unroll 1
                for (int cb = 0; cb < 3; ++cb) {
                    load_features(a.sc_darg, n, s, cb * 32, valid, v);
                    tile_store_row_tf32(Bhi + cb * dw::BLK, row, v);
                }
            }
            sync_issue([&] {
                tc::mma_outer(tb + dw::T_B, a_hi + dw::BLK_BYTES, a_lo + dw::BLK_BYTES, dw::BLK_BYTES, b_hi, dw::BLK_BYTES, 96, acc != 0);
            });
        }
        first = false;
    }

    // ---- flush: TMEM lane = accumulator row; one pass of atomics per CTA
    if (!first && tid < 128) {
        const uint32_t tl = tb + ((uint32_t)(warp * 32) << 16);
        float* g = a.gparams + (DEC == 0 ? OFF_LOW : (DEC == 1 ? OFF_HIGH : OFF_COLOR));
        const int k = tid;
        float d[32];
        if (warp < 2) {                                   // rows 0..63 carry something in at least one region
#pragma unroll 1
            for (int i = 1; i <= 4; ++i) {
                tc::tmem_ld32(tl + dw::T_H + (i - 1) * 32, d);
                tc::wait_ld();
                if (k < 32) { for (int j = 0; j < 32; ++j) atomicAdd(g + L::WH + (i - 1) * 1024 + k * 32 + j, d[j]); }
                else if (k == 32) { for (int j = 0; j < 32; ++j) atomicAdd(g + L::BP + i * 32 + j, d[j]); }
            }
#pragma unroll 1
            for (int i = 0; i <= 4; ++i) {
                tc::tmem_ld32(tl + dw::T_C + i * 32, d);
                tc::wait_ld();
                if (k < CD) { for (int j = 0; j < 32; ++j) atomicAdd(g + L::WC + i * CD * 32 + k * 32 + j, d[j]); }
                else if (k == CD) { for (int j = 0; j < 32; ++j) atomicAdd(g + L::BC + i * 32 + j, d[j]); }
            }
            // the constant-one block sits two blocks after block 0 when CD == 64 starts at block 0: row 64 then holds the bias
#pragma unroll 1
            for (int cb = 0; cb < 3; ++cb) {
                tc::tmem_ld32(tl + dw::T_03 + cb * 32, d);
                tc::wait_ld();
                for (int j = 0; j < 32; ++j) {
                    const int kk = cb * 32 + j;
                    if (kk < EMB) atomicAdd(g + L::W03 + kk * 64 + k, d[j]);                 // k = row j of [W0 ; W3e]
                    else if (kk == EMB && k < 32) atomicAdd(g + L::BP + k, d[j]);             // bias of block 0
                }
            }
#pragma unroll 1
            for (int cb = 0; cb < 3; ++cb) {
                tc::tmem_ld32(tl + dw::T_B + cb * 32, d);
                tc::wait_ld();
                if (k < 3) for (int j = 0; j < 32; ++j) { const int kk = cb * 32 + j; if (kk < EMB) atomicAdd(g + L::B + k * EMB_PAD + kk, d[j]); }
            }
            {
                uint32_t r[8];
                tc::tmem_ld8(tl + dw::T_O, r);
                tc::wait_ld();
                if (k < 32) { for (int o = 0; o < NOUT; ++o) atomicAdd(g + L::WO + o * 32 + k, __uint_as_float(r[o])); }
                else if (k == 32) { for (int o = 0; o < NOUT; ++o) atomicAdd(g + L::BO + o, __uint_as_float(r[o])); }
            }
        } else if (CD == 64 && warp == 2) {               // row 64: fc_c biases of the 64-feature decoder
#pragma unroll 1
            for (int i = 0; i <= 4; ++i) {
                tc::tmem_ld32(tl + dw::T_C + i * 32, d);
                tc::wait_ld();
                if (k == CD) { for (int j = 0; j < 32; ++j) atomicAdd(g + L::BC + i * 32 + j, d[j]); }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, 512);
}

template <int DEC>
int launch_decoder_dw_tc(const DecArgs& a, cudaStream_t st) {
    const size_t smem = dw::FLOATS * sizeof(float) + 1024;
    int rc = set_smem(k_decoder_dw_tc<DEC>, smem);
    if (rc) return rc;
    const int64_t tiles = (a.n + 127) / 128;
    const int blocks = grid_for(tiles, 1, num_sms());
    DFP_TIMED("k_decoder_dw_tc", st, k_decoder_dw_tc<DEC><<<blocks, 256, smem, st>>>(a));
    DFP_LAUNCH_CHECK("k_decoder_dw_tc");
    return DFP_OK;
}
template int launch_decoder_dw_tc<0>(const DecArgs&, cudaStream_t);
template int launch_decoder_dw_tc<1>(const DecArgs&, cudaStream_t);
template int launch_decoder_dw_tc<2>(const DecArgs&, cudaStream_t);

}  // namespace dfp
