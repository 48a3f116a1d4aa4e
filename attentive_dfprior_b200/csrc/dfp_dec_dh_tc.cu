// Tensor-core decoder backward, kernel 1 of 2: the delta chain.
//
// For tiles of 128 samples (one sample = one TMEM lane, served by two threads that own 16 columns each) this back-propagates through the five blocks
// of one decoder (src/conv_onet/models/decoder.py:192-203 reversed) using the activations the tensor-core forward
// kept (ReLU masks): per block ONE product  [d_c | dh_{i-1}] (+)= dacc_i * [M_i | W_i]  with dacc_i = mask_i(dh_i) written
// into TMEM by its owner threads (the feature gradient d_c = sum_i dh_i Wc_i is linear in the masked deltas, so
// M_i = W_i Wc_{i-1} comes precomputed from the pack kernel, dfp_layout.h DecM, and dh_i itself is never an operand; the
// dh_4 = dout W_o term is added on the CUDA cores) and the weights resident in shared memory in their natural [out][in]
// layout (the MN-major / transposed B operand, dfp_tc.cuh).  The tensor pipe is shared by the CTA's two tile groups and
// runs their MMAs in order, so halving the MMA count shortens both groups' chains.  The embedding gradient
// d_e = [dacc_0 | dacc_3] * [W_0 ; W_3e] is one K=64, N=96 product.  Outputs: the feature gradient scattered into
// the dense grid gradient (vector reductions), d(point) for the Tracker, and -- for kernel 2 (dfp_dec_dw_tc.cu) --
// the per-block deltas and d(arg) (one array, DLX) and sin(arg) (EX) as bf16 hi/mid planes (dfp_planes.cuh), plus one flag
// per tile (any upstream gradient?) that is kernel 2's tile list.
#include "dfp_device.cuh"
#include "dfp_planes.cuh"
#include "dfp_tc.cuh"

namespace dfp {

#ifdef DFP_EXPERIMENTS
__device__ unsigned long long g_prof_dh[32];
extern "C" int dfp_exp_read_prof_dh(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_prof_dh, sizeof(g_prof_dh)); }
extern "C" int dfp_exp_clear_prof_dh() { unsigned long long z[32] = {}; return (int)cudaMemcpyToSymbol(g_prof_dh, z, sizeof(z)); }
#define PROF_DECL long long pt_ = clock64(); const bool prof_ = (blockIdx.x == 0 && tid == 32 && DEC == 2);
#define PROF_TILE if (prof_) atomicAdd(&g_prof_dh[31], 1ull);
#define PROF(i) { const long long n_ = clock64(); if (prof_) atomicAdd(&g_prof_dh[i], (unsigned long long)(n_ - pt_)); pt_ = n_; }
#else
#define PROF_DECL
#define PROF_TILE
#define PROF(i)
#endif

struct DhLayout {                                       // shared memory, in floats
    static constexpr int PAIR = 0;                      // 4 x {M_i tile [32 j][32 q], W_i tile [32 j][32 kin]} for blocks i = 1..4 (W: pts_linears.1, .2, .3[:, 93:], .4)
    static constexpr int W03 = PAIR + 4 * 2048;         // 3 column blocks of [64 j][32 kin]: rows 0..31 pts_linears.0, 32..63 pts_linears.3[:, :93]
    static constexpr int BSIZE = W03 + 3 * 2048;        // one split half (hi or lo)
    static constexpr int S_WO = 2 * BSIZE;              // [4][32] output rows
    static constexpr int S_B = S_WO + 128;              // [3][96] embedding matrix
    static constexpr int S_V = S_B + 288;               // [4][32] V = W_o Wc_4
    static constexpr int FLOATS = S_V + 128;
    // TMEM columns of one 128-sample group
    static constexpr int T_A = 0;                       // 128 columns of A operands (see kernel)
    static constexpr int T_DC = 128;                    // d_c accumulator ...
    static constexpr int T_DX = 160;                    // ... directly followed by dh_{i-1} (32) / later d_e (96)
    static constexpr int T_COLS = 256;
};

template <int CD>
__device__ __forceinline__ void stage_weights_dh(float* sm, const float* __restrict__ P, const float* __restrict__ PM /* DecM block */) {
    using L = DecOff<CD>;
    float* hi = sm;
    float* lo = sm + DhLayout::BSIZE;
    const int nt = blockDim.x;
    for (int i = threadIdx.x; i < 4 * 1024; i += nt) {                 // packed WH: [slot][kin][j]
        const int slot = i >> 10, kin = (i >> 5) & 31, j = i & 31;
        const float w = P[L::WH + i];
        const float h = tc::tf32_rna(w);
        const int o = DhLayout::PAIR + slot * 2048 + 1024 + tc::mn_tile_index(j, kin);
        hi[o] = h; lo[o] = tc::tf32_rna(w - h);
    }
    for (int i = threadIdx.x; i < 4 * 1024; i += nt) {                 // DecM::M: [slot][j][q]
        const int slot = i >> 10, j = (i >> 5) & 31, q = i & 31;
        const float w = PM[DecM::M + i];
        const float h = tc::tf32_rna(w);
        const int o = DhLayout::PAIR + slot * 2048 + tc::mn_tile_index(j, q);
        hi[o] = h; lo[o] = tc::tf32_rna(w - h);
    }
    for (int i = threadIdx.x; i < 96 * 64; i += nt) {                  // packed W03: [kin][64 j]
        const int kin = i >> 6, j = i & 63;
        const float w = (kin < EMB) ? P[L::W03 + i] : 0.f;
        const float h = tc::tf32_rna(w);
        const int o = DhLayout::W03 + (kin >> 5) * 2048 + tc::mn_tile_index(j, kin & 31);
        hi[o] = h; lo[o] = tc::tf32_rna(w - h);
    }
    for (int i = threadIdx.x; i < 128; i += nt) { sm[DhLayout::S_WO + i] = P[L::WO + i]; sm[DhLayout::S_V + i] = PM[DecM::V + i]; }
    for (int i = threadIdx.x; i < 288; i += nt) sm[DhLayout::S_B + i] = P[L::B + i];
}

// OR-reduction of a predicate over the 256 threads of one group (named barrier `id`)
__device__ __forceinline__ bool group_any(bool v, int id) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 q, %1, 0;\n\t"
        "barrier.cta.red.or.pred p, %2, 256, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(r)
        : "r"((uint32_t)v), "r"(id)
        : "memory");
    return r != 0;
}

// Two threads per sample (as in the forward, dfp_dec_fwd_tc.cu): thread (row r, half hf) owns columns 16*hf..16*hf+15
// of every 32-wide delta block of sample r.
// `cta` of `nctas` CTAs work on this decoder (the whole grid, or a share of it in the merged kernel).
template <int DEC, int MODE>
__device__ __forceinline__ void decoder_dh_tc_body(const DecArgs& a, const int cta, const int nctas) {
    constexpr int CD = (DEC == 1) ? 64 : 32;
    constexpr int NOUT = (DEC == 2) ? 4 : 1;
    using PT = typename PtT<MODE>::T;
    using T = DhLayout;
    if ((int64_t)cta * 256 >= dec_items(a)) return;                      // no tile for this CTA: skip the staging too
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sm = reinterpret_cast<float*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float part[2][6][128];                                  // d(point) partial sums of the upper halves
    // feature gradients of a tile (rows padded against bank conflicts) and, per (sample, corner), the voxel offset (in
    // float4 units, -1 = nothing to add) and trilinear weight, for the cooperative scatter: carved from the dynamic
    // allocation behind the weights
    float (*s_dc)[128][36] = reinterpret_cast<float (*)[128][36]>(sm + T::FLOATS);
    int2 (*s_cw)[1024] = reinterpret_cast<int2 (*)[1024]>(sm + T::FLOATS + 2 * 128 * 36);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int grp = tid >> 8, gt = tid & 255;
    const int r = gt & 127, hf = gt >> 7;
    stage_weights_dh<CD>(sm, a.params + (DEC == 0 ? OFF_LOW : (DEC == 1 ? OFF_HIGH : OFF_COLOR)), a.params + OFF_DEC_M + DEC * DecM::SIZE);
    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
    if (tid == 0) { tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    const uint32_t tb = tmem_base_s + (uint32_t)(grp * T::T_COLS);
    const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16) + 16u * hf;      // this warp's lanes, this half's columns
    uint64_t* bar = &bars[grp];
    uint32_t parity = 0;
    const uint32_t whi = tc::smem_u32(sm), wlo = tc::smem_u32(sm + T::BSIZE);
    const BoxT<PT> box = make_box<PT>(a.bound.b);
    const int64_t n = dec_items(a);
    const int64_t tiles = (n + 127) / 128;
    const bool issuer_warp = __shfl_sync(0xffffffffu, gt >> 5, 0) == 0;   // warp-uniform by construction
    const bool planes = a.gparams != nullptr && !DFP_EXP(a, 1);
    const float bmax = embed_bmax(sm + T::S_B);
    const int c0 = 16 * hf;

    for (int64_t tile = (int64_t)cta * 2 + grp; tile < tiles; tile += (int64_t)nctas * 2) {
        PROF_DECL
        const int64_t i_raw = tile * 128 + r;
        const bool valid = i_raw < n;
        const int64_t ia = valid ? i_raw : n - 1;                    // list position: index of saved activations and scratch
        const int64_t s = dec_sample(a, ia);                         // sample id: index of points, dout, dp
        const pl::Row r_dx = pl::row_of(a.sc_darg, pl::F_DLX, i_raw), r_ex = pl::row_of(a.sc_e, pl::F_EX, i_raw);
        constexpr int DL1 = pl::DL1_AT;                              // [dacc3 | dacc0 | darg96] starts at dacc3's place in DLX
        float dout[NOUT];
        if (NOUT == 1) dout[0] = valid ? a.dout[s] : 0.f;
        else {
            const float4 v = *reinterpret_cast<const float4*>(a.dout + 4 * s);
            dout[0] = valid ? v.x : 0.f; dout[NOUT > 1 ? 1 : 0] = valid ? v.y : 0.f; dout[NOUT > 2 ? 2 : 0] = valid ? v.z : 0.f;
            dout[NOUT > 3 ? 3 : 0] = 0.f;                    // the colour decoder's 4th output is overwritten (decoder.py:352)
        }
        bool any = false;
#pragma unroll
        for (int o = 0; o < NOUT; ++o) any = any || (dout[o] != 0.f);
        // kernel 2 reads the planes of every sample of a tile it processes; a tile whose upstream gradients are all zero is
        // skipped by both kernels: the flag written here is kernel 2's tile list
        PROF(0)
        const bool tile_any = group_any(any, 1 + grp);
        if (gt == 0 && planes) a.tileflag[tile] = tile_any ? 1 : 0;
        if (!tile_any) continue;
        PROF(1)

        float dh[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) dh[j] = 0.f;
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
            float wo[16];
            tc::lds16(sm + T::S_WO + o * 32 + c0, wo);
#pragma unroll
            for (int j = 0; j < 16; ++j) dh[j] = fmaf(wo[j], dout[o], dh[j]);
        }
        float dacc3[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) dacc3[j] = 0.f;

        uint32_t m_next = a.act_m[tm_index(ia, 4, 5)];             // ReLU masks are fetched one block ahead of their use
#pragma unroll 1
        for (int i = 4; i >= 1; --i) {
            const uint32_t m = m_next >> c0;
            m_next = a.act_m[tm_index(ia, i - 1, 5)];
            float dacc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) dacc[j] = ((m >> j) & 1u) ? dh[j] : 0.f;
            tc::tmem_store_split16(tl + T::T_A, tl + T::T_A + 32, dacc);          // the block's only operand
            if (i != 4) {                                                         // the product accumulates (d_c adds up): dh starts from zero
                const uint32_t zero[16] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                tc::tmem_st16(tl + T::T_DX, zero);
            }
            tc::wait_st();
            tc::fence_before_sync();
            PROF(2)
            tc::named_bar_sync(3 + grp, 256);
            PROF(3)
            if (issuer_warp && tc::elect_one()) {
                tc::fence_after_sync();
                // [d_c | dh_{i-1}] (+)= dacc_i * [M_i | W_i]
                tc::mma_3xtf32_bT(tb + T::T_DC, tb + T::T_A, tb + T::T_A + 32, whi + (T::PAIR + (i - 1) * 2048) * 4,
                                  wlo + (T::PAIR + (i - 1) * 2048) * 4, 32, 64, 4096u, i != 4);
                tc::mma_commit(bar);
            }
            // while the product runs (off the chain store -> barrier -> issue -> wait): operands of kernel 2 (rows past the
            // end carry zero deltas); dacc3 lands at DL1 + 0, stored once for both passes
            if (planes) pl::store16(r_dx, 8 + pl::dl0_pos(i) * 32 + c0, dacc);
            if (i == 3) {
#pragma unroll
                for (int j = 0; j < 16; ++j) dacc3[j] = dacc[j];
            }
            PROF(4)
            tc::mbar_wait(bar, parity); parity ^= 1;
            tc::fence_after_sync();
            PROF(5)
            tc::tmem_ld16(tl + T::T_DX, dh);
            tc::wait_ld();
            PROF(6)
        }
        // block 0: the embedding
        {
            const uint32_t m0 = m_next >> c0;
#pragma unroll
            for (int j = 0; j < 16; ++j) dh[j] = ((m0 >> j) & 1u) ? dh[j] : 0.f;      // dacc_0
            // A = [dacc_0 | dacc_3]: hi in columns 0..63, lo in 64..127
            tc::tmem_store_split16(tl + T::T_A, tl + T::T_A + 64, dh);
            tc::tmem_store_split16(tl + T::T_A + 32, tl + T::T_A + 96, dacc3);
            tc::wait_st();
            tc::fence_before_sync();
            tc::named_bar_sync(3 + grp, 256);
            if (issuer_warp && tc::elect_one()) {
                tc::fence_after_sync();
                tc::mma_3xtf32_bT(tb + T::T_DX, tb + T::T_A, tb + T::T_A + 64, whi + T::W03 * 4, wlo + T::W03 * 4, 64, 96, 8192u, false);
                tc::mma_commit(bar);
            }
            if (planes) {                                     // while the embedding product runs
                pl::store16(r_dx, DL1 + 32 + c0, dh);
                if (hf == 0) {                                // upstream gradient, padded to 8 columns
                    float d8[8];
#pragma unroll
                    for (int o = 0; o < 8; ++o) d8[o] = 0.f;
#pragma unroll
                    for (int o = 0; o < NOUT; ++o) d8[o] = dout[o];
                    pl::store8(r_dx, 0, d8);
                }
            }
        }
        PROF(9)
        // while the embedding product runs: point, trilinear set-up
        float xn[3], pf[3];
        point_inputs<MODE>(a, box, s, xn, pf);
        const bool fast = embed_fast_ok(pf, bmax);
        const Tri tri = tri_setup(xn, a.grid.W, a.grid.H, a.grid.D);
        PROF(10)
        tc::mbar_wait(bar, parity); parity ^= 1;
        tc::fence_after_sync();
        PROF(11)
        float gp0 = 0.f, gp1 = 0.f, gp2 = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < (DFP_EXP(a, 4) ? 0 : 3); ++ch) {
            float de[16];
            tc::tmem_ld16(tl + T::T_DX + ch * 32, de);
            tc::wait_ld();
            float sn[16], da[16];
            if (fast) {                                   // branch hoisted out of the unrolled loop (code size / I-cache)
                float arg[16];
                embed_arg16(sm + T::S_B, pf, ch * 32 + c0, arg);
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) {
                    const int k = ch * 32 + c0 + kk;
                    float s1 = (k == EMB) ? 1.f : 0.f, c1 = 0.f;    // column 93 of the sin array: constant 1 (bias of block 0 in kernel 2)
                    if (k < EMB) sincos_sfu(arg[kk], s1, c1);
                    sn[kk] = s1;
                    da[kk] = de[kk] * c1;
                }
            } else {
#pragma unroll 1
                for (int kk = 0; kk < 16; ++kk) {
                    const int k = ch * 32 + c0 + kk;
                    float s1 = (k == EMB) ? 1.f : 0.f, c1 = 0.f;
                    if (k < EMB) sincosf(embed_arg(sm + T::S_B, pf, k), &s1, &c1);
#pragma unroll
                    for (int u = 0; u < 16; ++u) if (u == kk) { sn[u] = s1; da[u] = de[u] * c1; }
                }
            }
            if (a.dp != nullptr) {                            // d(point) through the embedding: Tracker only
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                const int k = ch * 32 + c0 + kk;
                if (k < EMB) {
                    gp0 = fmaf(sm[T::S_B + k], da[kk], gp0);
                    gp1 = fmaf(sm[T::S_B + EMB_PAD + k], da[kk], gp1);
                    gp2 = fmaf(sm[T::S_B + 2 * EMB_PAD + k], da[kk], gp2);
                }
            }
            }
            if (planes) {
                pl::store16(r_ex, ch * 32 + c0, sn);
                pl::store16(r_dx, DL1 + 64 + ch * 32 + c0, da);
            }
        }
        if (planes && hf == 0) {                              // the point itself (A operand of embedder._B's gradient)
            const float p8[8] = {pf[0], pf[1], pf[2], 0.f, 0.f, 0.f, 0.f, 0.f};
            pl::store8(r_ex, 96, p8);
        }
        PROF(12)
        float d_c[16];
        tc::tmem_ld16(tl + T::T_DC, d_c);
        tc::wait_ld();
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {                                 // + dh_4 Wc_4 = dout (W_o Wc_4)
            float vo[16];
            tc::lds16(sm + T::S_V + o * 32 + c0, vo);
#pragma unroll
            for (int j = 0; j < 16; ++j) d_c[j] = fmaf(vo[j], dout[o], d_c[j]);
        }
        const bool want_coord = a.dp != nullptr;
        float gn[3] = {0.f, 0.f, 0.f};
        if (a.ggrid && !want_coord && !DFP_EXP(a, 2)) {
            // Cooperative scatter.  A thread's own reduction would touch 32 different lines per warp instruction (one
            // 16-byte piece per lane, ~2 cycles of the load/store unit each: measured, this was half of the kernel).
            // Instead the tile's feature gradients go through shared memory and eight lanes add one corner's whole
            // 128-byte line (32 channels): four lines per warp instruction.
            {
                float* row = &s_dc[grp][r][16 * hf];
#pragma unroll
                for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(row + 4 * q) = make_float4(d_c[4 * q], d_c[4 * q + 1], d_c[4 * q + 2], d_c[4 * q + 3]);
                // each owner thread fills four of its sample's eight (offset, weight) pairs once
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    int x, y, z;
                    const bool ok = tri_corner(tri, 4 * hf + cc, a.grid.W, a.grid.H, a.grid.D, x, y, z) && valid && any;
                    s_cw[grp][r * 8 + 4 * hf + cc] = make_int2(ok ? ((z * a.grid.H + y) * a.grid.W + x) * (CDIM / 4) : -1,
                                                               __float_as_int(tri_weight(tri, 4 * hf + cc)));
                }
            }
            PROF(13)
            tc::named_bar_sync(3 + grp, 256);
            PROF(14)
            // Eight lanes own one corner index (0..7) of 32 consecutive samples of the tile.  Consecutive samples of a ray
            // mostly sit in the same voxel cell (the near-surface samples are centimetres apart, a voxel is 16 / 32 cm), so
            // the contributions to one corner line are summed in registers for as long as the line stays the same and
            // only one reduction per run goes to L2 (roughly half the reductions on the fine grids, a quarter on the coarse one).
            const int sub = gt >> 3, q = gt & 7;
            const int corner = sub & 7, smp0 = (sub >> 3) * 32;
            float4* gg4 = reinterpret_cast<float4*>(a.ggrid);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int cur = -1;
#pragma unroll 4
            for (int it = 0; it < 32; ++it) {
                const int smp = smp0 + it;
                const int2 ent = s_cw[grp][smp * 8 + corner];
                if (ent.x != cur) {
                    if (cur >= 0) red_add_v4(reinterpret_cast<float*>(gg4 + cur + q), acc.x, acc.y, acc.z, acc.w);
                    acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    cur = ent.x;
                }
                if (ent.x >= 0) {
                    const float wgt = __int_as_float(ent.y);
                    const float4 v = *reinterpret_cast<const float4*>(&s_dc[grp][smp][4 * q]);
                    acc.x = fmaf(v.x, wgt, acc.x); acc.y = fmaf(v.y, wgt, acc.y); acc.z = fmaf(v.z, wgt, acc.z); acc.w = fmaf(v.w, wgt, acc.w);
                }
            }
            if (cur >= 0) red_add_v4(reinterpret_cast<float*>(gg4 + cur + q), acc.x, acc.y, acc.z, acc.w);
        } else if (valid && any && (a.ggrid || want_coord) && !DFP_EXP(a, 2)) {
            scatter_features16(a.grid, a.ggrid, tri, hf, d_c, want_coord, gn);
        }
        PROF(15)
        PROF_TILE
        if (want_coord) {                                            // uniform: the two halves of a sample add up in the lower one
            if (hf == 1) {
                part[grp][0][r] = gn[0]; part[grp][1][r] = gn[1]; part[grp][2][r] = gn[2];
                part[grp][3][r] = gp0; part[grp][4][r] = gp1; part[grp][5][r] = gp2;
            }
            tc::named_bar_sync(3 + grp, 256);
            if (hf == 0 && valid && any) {
                gn[0] += part[grp][0][r]; gn[1] += part[grp][1][r]; gn[2] += part[grp][2][r];
                gp0 += part[grp][3][r]; gp1 += part[grp][4][r]; gp2 += part[grp][5][r];
                // reductions: in the merged launch the decoders' chains add to the same d(point) concurrently
                atomicAdd(a.dp + 3 * s + 0, gn[0] * (float)(2.0 / (a.bound.b[0][1] - a.bound.b[0][0])) + gp0);
                atomicAdd(a.dp + 3 * s + 1, gn[1] * (float)(2.0 / (a.bound.b[1][1] - a.bound.b[1][0])) + gp1);
                atomicAdd(a.dp + 3 * s + 2, gn[2] * (float)(2.0 / (a.bound.b[2][1] - a.bound.b[2][0])) + gp2);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}

template <int DEC, int MODE>
__global__ void __launch_bounds__(512, 1) k_decoder_dh_tc(const DecArgs a) {
    decoder_dh_tc_body<DEC, MODE>(a, blockIdx.x, gridDim.x);
}

__device__ __forceinline__ int64_t dh_iterations(const DecArgs& a, int on) { return on ? (dec_items(a) + 255) / 256 : 0; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_decoder_dh_tc_all(const DecArgs3 A) {
    const int64_t it[3] = {dh_iterations(A.d[0], A.on[0]), dh_iterations(A.d[1], A.on[1]), dh_iterations(A.d[2], A.on[2])};
    int first, count;
    split_ctas(it, gridDim.x, 0, first, count);
    if ((int)blockIdx.x < first + count) { decoder_dh_tc_body<1, MODE>(A.d[0], blockIdx.x - first, count); return; }
    split_ctas(it, gridDim.x, 1, first, count);
    if ((int)blockIdx.x < first + count) { decoder_dh_tc_body<2, MODE>(A.d[1], blockIdx.x - first, count); return; }
    split_ctas(it, gridDim.x, 2, first, count);
    if ((int)blockIdx.x < first + count) decoder_dh_tc_body<0, MODE>(A.d[2], blockIdx.x - first, count);
}

template <int MODE>
int launch_decoder_dh_tc_all(const DecArgs3& a, cudaStream_t st) {
    const size_t smem = (DhLayout::FLOATS + 2 * 128 * 36 + 2 * 1024 * 2) * sizeof(float) + 1024;
    int rc = set_smem(k_decoder_dh_tc_all<MODE>, smem);
    if (rc) return rc;
    int64_t its = 0;                                         // upper bound (the high decoder's list may be shorter)
    for (int d = 0; d < 3; ++d) if (a.on[d]) its += (a.d[d].n + 255) / 256;
    const int blocks = (int)(its < (int64_t)num_sms() ? (its < 1 ? 1 : its) : num_sms());
    DFP_TIMED("k_decoder_dh_tc_all", st, k_decoder_dh_tc_all<MODE><<<blocks, 512, smem, st>>>(a));
    DFP_LAUNCH_CHECK("k_decoder_dh_tc_all");
    return DFP_OK;
}
template int launch_decoder_dh_tc_all<0>(const DecArgs3&, cudaStream_t);
template int launch_decoder_dh_tc_all<1>(const DecArgs3&, cudaStream_t);
template int launch_decoder_dh_tc_all<2>(const DecArgs3&, cudaStream_t);

template <int DEC, int MODE>
int launch_decoder_dh_tc(const DecArgs& a, cudaStream_t st) {
    const size_t smem = (DhLayout::FLOATS + 2 * 128 * 36 + 2 * 1024 * 2) * sizeof(float) + 1024;
    int rc = set_smem(k_decoder_dh_tc<DEC, MODE>, smem);
    if (rc) return rc;
    const int64_t tiles = (a.n + 127) / 128;
    const int blocks = grid_for(tiles, 2, num_sms());
    DFP_TIMED("k_decoder_dh_tc", st, k_decoder_dh_tc<DEC, MODE><<<blocks, 512, smem, st>>>(a));
    DFP_LAUNCH_CHECK("k_decoder_dh_tc");
    return DFP_OK;
}
template int launch_decoder_dh_tc<0, 0>(const DecArgs&, cudaStream_t);
template int launch_decoder_dh_tc<1, 0>(const DecArgs&, cudaStream_t);
template int launch_decoder_dh_tc<2, 0>(const DecArgs&, cudaStream_t);
template int launch_decoder_dh_tc<0, 1>(const DecArgs&, cudaStream_t);
template int launch_decoder_dh_tc<1, 1>(const DecArgs&, cudaStream_t);
template int launch_decoder_dh_tc<2, 1>(const DecArgs&, cudaStream_t);
template int launch_decoder_dh_tc<0, 2>(const DecArgs&, cudaStream_t);
template int launch_decoder_dh_tc<1, 2>(const DecArgs&, cudaStream_t);
template int launch_decoder_dh_tc<2, 2>(const DecArgs&, cudaStream_t);

}  // namespace dfp
