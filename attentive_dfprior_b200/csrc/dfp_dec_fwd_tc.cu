// Tensor-core decoder forward (tcgen05 + TMEM, 3xTF32): the grid-conditioned MLP of
// src/conv_onet/models/decoder.py:177-203 for tiles of 128 samples.
//
// One thread owns one sample = one TMEM lane.  It gathers the trilinear features and produces the
// Fourier embedding on the CUDA cores / special-function unit (range reduction of f32 arguments of several hundred
// radians stays there), writes those activations straight into TMEM as the A operand, and the five 32-wide blocks run as
// tcgen05.mma with the weights resident in shared memory (canonical K-major layout, hi/lo TF32 split).
// Epilogues (bias, ReLU, feature term, split for the next layer) read the accumulators back with tcgen05.ld.
// A CTA runs two independent 128-sample groups (one for the 64-feature high decoder) so that one group's
// epilogue overlaps the other group's MMAs.
#include "dfp_device.cuh"
#include "dfp_planes.cuh"
#include "dfp_tc.cuh"

namespace dfp {

#ifdef DFP_EXPERIMENTS
// phase timeline of one non-issuer thread (group 0, warp 1, lane 0) of CTA 0: cycles summed per phase over its tiles
__device__ unsigned long long g_prof_fwd[32];
extern "C" int dfp_exp_read_prof_fwd(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, g_prof_fwd, sizeof(g_prof_fwd)); }
extern "C" int dfp_exp_clear_prof_fwd() { unsigned long long z[32] = {}; return (int)cudaMemcpyToSymbol(g_prof_fwd, z, sizeof(z)); }
#define PROF_DECL long long pt_ = clock64(); const bool prof_ = (blockIdx.x == 0 && tid == 32 && DEC == 2);
#define PROF_TILE if (prof_) atomicAdd(&g_prof_fwd[31], 1ull);
#define PROF(i) { const long long n_ = clock64(); if (prof_) atomicAdd(&g_prof_fwd[i], (unsigned long long)(n_ - pt_)); pt_ = n_; }
#else
#define PROF_DECL
#define PROF_TILE
#define PROF(i)
#endif

template <int CD>
struct TcLayout {                                   // shared memory, in floats
    static constexpr int W03 = 0;                   // B [64][96]  rows 0..31 pts_linears.0, 32..63 pts_linears.3[:, :93]
    static constexpr int WH = W03 + 64 * 96;        // B 4 x [32][32]
    static constexpr int WC = WH + 4 * 32 * 32;     // B 5 x [32][CD]
    static constexpr int BSIZE = WC + 5 * 32 * CD;  // one split half
    static constexpr int SMALL = 2 * BSIZE;         // biases etc. (plain f32)
    static constexpr int S_BP = SMALL;              // [5][32]
    static constexpr int S_BC = S_BP + 160;         // [5][32]
    static constexpr int S_B = S_BC + 160;          // [3][96] embedding matrix
    static constexpr int S_WO = S_B + 288;          // [4][32]
    static constexpr int S_BO = S_WO + 128;         // [4]
    static constexpr int FLOATS = S_BO + 4;
    static constexpr int GROUPS = (CD == 32) ? 2 : 1;
    static constexpr int GATHER = (FLOATS + 3) / 4 * 4;                                   // staging of the cooperative gather
    static constexpr int NGRID = (CD == 64) ? 2 : 1;                                      // grids a sample gathers from
    static constexpr int TOTAL = GATHER + GROUPS * 128 * (CD + 4) + GROUPS * NGRID * 128 * 8 * 2;   // + corner table (int2 per sample and corner)
    // TMEM columns of one group
    static constexpr int T_C_HI = 0, T_C_LO = CD, T_X_HI = 2 * CD, T_X_LO = 2 * CD + 32;
    static constexpr int T_D03 = 2 * CD + 64, T_DH = 2 * CD + 128, T_DC = 2 * CD + 160;
    static constexpr int T_COLS = 2 * CD + 192;     // 256 (CD=32) / 320 (CD=64)
};

template <int CD>
__device__ __forceinline__ void stage_weights_tc(float* sm, const float* __restrict__ P /* packed decoder params */) {
    using L = DecOff<CD>;
    using T = TcLayout<CD>;
    float* hi = sm;
    float* lo = sm + T::BSIZE;
    const int nt = blockDim.x;
    for (int i = threadIdx.x; i < 64 * 96; i += nt) {               // W03: packed [k][64]
        const int n = i & 63, k = i >> 6;
        const float w = (k < EMB) ? P[L::W03 + k * 64 + n] : 0.f;
        const float h = tc::tf32_rna(w);
        hi[T::W03 + tc::b_index(n, k, 64)] = h;
        lo[T::W03 + tc::b_index(n, k, 64)] = tc::tf32_rna(w - h);
    }
    for (int i = threadIdx.x; i < 4 * 1024; i += nt) {              // WH: packed [slot][k][32]
        const int slot = i >> 10, k = (i >> 5) & 31, n = i & 31;
        const float w = P[L::WH + i];
        const float h = tc::tf32_rna(w);
        hi[T::WH + slot * 1024 + tc::b_index(n, k, 32)] = h;
        lo[T::WH + slot * 1024 + tc::b_index(n, k, 32)] = tc::tf32_rna(w - h);
    }
    for (int i = threadIdx.x; i < 5 * CD * 32; i += nt) {           // WC: packed [layer][k][32]
        const int layer = i / (CD * 32), r = i % (CD * 32), k = r >> 5, n = r & 31;
        const float w = P[L::WC + i];
        const float h = tc::tf32_rna(w);
        hi[T::WC + layer * CD * 32 + tc::b_index(n, k, 32)] = h;
        lo[T::WC + layer * CD * 32 + tc::b_index(n, k, 32)] = tc::tf32_rna(w - h);
    }
    for (int i = threadIdx.x; i < 160; i += nt) { sm[T::S_BP + i] = P[L::BP + i]; sm[T::S_BC + i] = P[L::BC + i]; }
    for (int i = threadIdx.x; i < 288; i += nt) sm[T::S_B + i] = P[L::B + i];
    for (int i = threadIdx.x; i < 128; i += nt) sm[T::S_WO + i] = P[L::WO + i];
    if (threadIdx.x < 4) sm[T::S_BO + threadIdx.x] = P[L::BO + threadIdx.x];
}

// Two threads per sample: thread (row r, half hf) owns columns 16*hf..16*hf+15 of every 32-wide block of sample r
// (both halves sit in warps with the same warp%4, i.e. the same TMEM lane quarter).  That doubles the warps per SM at
// the same TMEM budget; the single-thread-per-sample version of this kernel ran at ~30 % issue utilisation with 8 warps.
// `cta` of `nctas` CTAs work on this decoder (the whole grid for the one-decoder kernel, a share of it for the merged one;
// the CTA may have more threads than this decoder's GROUPS * 256: the rest only takes part in the CTA-wide barriers).
template <int DEC, int MODE>
__device__ __forceinline__ void decoder_fwd_tc_body(const DecArgs& a, const int cta, const int nctas) {
    constexpr int CD = (DEC == 1) ? 64 : 32;
    constexpr int NOUT = (DEC == 2) ? 4 : 1;
    using PT = typename PtT<MODE>::T;
    using T = TcLayout<CD>;
    if ((int64_t)cta * T::GROUPS * 128 >= dec_items(a)) return;          // no tile for this CTA (short band list, small batch): skip the staging too
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sm = reinterpret_cast<float*>(smem_raw);
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float part[2][4][128];                                  // output-layer partial sums of the upper halves
    // cooperative gather staging behind the weights: interpolated features [group][128][CD + 4] (padded rows) and the
    // corner table [group][grid][128 samples][8 corners] of (voxel offset in float4 units or -1, trilinear weight)
    constexpr int C_W = CD + 4, NGRID = T::NGRID;
    float* s_c = sm + T::GATHER;
    int2* s_cw = reinterpret_cast<int2*>(s_c + T::GROUPS * 128 * C_W);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int grp = tid >> 8, gt = tid & 255;                         // group, thread within the group
    const int r = gt & 127, hf = gt >> 7;                             // sample row of the tile, column half
    stage_weights_tc<CD>(sm, a.params + (DEC == 0 ? OFF_LOW : (DEC == 1 ? OFF_HIGH : OFF_COLOR)));
    if (warp == 0) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
    if (tid == 0) { tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    const uint32_t tb = tmem_base_s + (uint32_t)(grp * T::T_COLS);                 // this group's columns, lane 0
    const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16) + 16u * hf;       // ... this warp's lanes, this half's columns
    uint64_t* bar = &bars[grp];
    uint32_t parity = 0;
    const uint32_t bhi = tc::smem_u32(sm), blo = tc::smem_u32(sm + T::BSIZE);
    const BoxT<PT> box = make_box<PT>(a.bound.b);
    const int64_t nitems = dec_items(a);
    const int64_t tiles = (nitems + 127) / 128;
    const bool issuer_warp = __shfl_sync(0xffffffffu, gt >> 5, 0) == 0;   // warp-uniform by construction
    const float bmax = embed_bmax(sm + T::S_B);

    if (grp < T::GROUPS)
    for (int64_t tile = (int64_t)cta * T::GROUPS + grp; tile < tiles; tile += (int64_t)nctas * T::GROUPS) {
        PROF_DECL
        const int64_t i_raw = tile * 128 + r;
        const bool valid = i_raw < nitems;
        const int64_t ia = valid ? i_raw : nitems - 1;               // position in the (optional) list: index of saved activations
        const int64_t s = dec_sample(a, ia);                         // sample id: index of points and outputs
        float xn[3], pf[3];
        point_inputs<MODE>(a, box, s, xn, pf);
        const bool fast = embed_fast_ok(pf, bmax);
        pl::Row r_c, r_h;
        const bool save_pl = a.act_c && !DFP_EXP(a, 16);
        if (a.act_c) { r_c = pl::row_of(a.act_c, CD, i_raw); r_h = pl::row_of(a.act_h, pl::F_H_ST, i_raw); }
        {   // trilinear features -> A_C.  Cooperative gather: a thread fetching its own 16 channels touches 32 different
            // lines per warp instruction (one 16-byte piece per lane); instead eight lanes fetch one corner's whole
            // 128-byte line and the interpolated features go to their owners through shared memory.
            // Every (sample, corner) pair's voxel offset and weight are computed ONCE by the sample's two owner threads
            // (four corners each); the eight lanes that then fetch a corner's 128-byte line only read the pair.
#pragma unroll
            for (int gsel = 0; gsel < NGRID; ++gsel) {
                const DfpGrid& g = gsel == 0 ? a.grid : a.grid_low;
                const Tri t0 = tri_setup(xn, g.W, g.H, g.D);
                int2* e = &s_cw[((grp * NGRID + gsel) * 128 + r) * 8 + 4 * hf];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    int x, y, z;
                    const bool ok = tri_corner(t0, 4 * hf + cc, g.W, g.H, g.D, x, y, z);
                    e[cc] = make_int2(ok ? ((z * g.H + y) * g.W + x) * (CDIM / 4) : -1, __float_as_int(tri_weight(t0, 4 * hf + cc)));
                }
            }
            PROF(0)
            tc::named_bar_sync(3 + grp, 256);
            PROF(1)
            {
                const int sub = gt >> 3, q = gt & 7;
#pragma unroll 1
                for (int j = 0; j < 4; ++j) {
                    const int smp = j * 32 + sub;
#pragma unroll
                    for (int gsel = 0; gsel < NGRID; ++gsel) {
                        const float4* gdata = reinterpret_cast<const float4*>((gsel == 0 ? a.grid : a.grid_low).data);
                        const int4* e4 = reinterpret_cast<const int4*>(&s_cw[((grp * NGRID + gsel) * 128 + smp) * 8]);
                        float4 v[8];
                        float wgt[8];
#pragma unroll
                        for (int c2 = 0; c2 < 4; ++c2) {
                            const int4 ent = e4[c2];                       // two (offset, weight) pairs
                            wgt[2 * c2] = __int_as_float(ent.y); wgt[2 * c2 + 1] = __int_as_float(ent.w);
                            v[2 * c2] = (ent.x >= 0 && !DFP_EXP(a, 32)) ? __ldg(gdata + ent.x + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                            v[2 * c2 + 1] = (ent.z >= 0 && !DFP_EXP(a, 32)) ? __ldg(gdata + ent.z + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int cn = 0; cn < 8; ++cn) {
                            acc.x = fmaf(v[cn].x, wgt[cn], acc.x); acc.y = fmaf(v[cn].y, wgt[cn], acc.y);
                            acc.z = fmaf(v[cn].z, wgt[cn], acc.z); acc.w = fmaf(v[cn].w, wgt[cn], acc.w);
                        }
                        *reinterpret_cast<float4*>(&s_c[(grp * 128 + smp) * C_W + 32 * gsel + 4 * q]) = acc;
                    }
                }
            }
            PROF(2)
            tc::named_bar_sync(3 + grp, 256);
            PROF(3)
            float c[16];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const float4 v = *reinterpret_cast<const float4*>(&s_c[(grp * 128 + r) * C_W + 16 * hf + 4 * q4]);
                c[4 * q4] = v.x; c[4 * q4 + 1] = v.y; c[4 * q4 + 2] = v.z; c[4 * q4 + 3] = v.w;
            }
            tc::tmem_store_split16(tl + T::T_C_HI, tl + T::T_C_LO, c);
            // saved operands go out as bf16 planes at the row's own position (every row of a tile is written: rows past
            // the end repeat the last sample, their deltas are zero in the backward)
            if (save_pl) {
                pl::store16(r_c, 16 * hf, c);
            }
            if (DEC == 1) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 v = *reinterpret_cast<const float4*>(&s_c[(grp * 128 + r) * C_W + 32 + 16 * hf + 4 * q4]);
                    c[4 * q4] = v.x; c[4 * q4 + 1] = v.y; c[4 * q4 + 2] = v.z; c[4 * q4 + 3] = v.w;
                }
                tc::tmem_store_split16(tl + T::T_C_HI + 32, tl + T::T_C_LO + 32, c);
                if (save_pl) pl::store16(r_c, 32 + 16 * hf, c);
            }
        }
        PROF(4)
        // Fourier embedding in three 32-wide K chunks (16 per thread), each accumulated into D03 = [z0 | skip part of z3]
#pragma unroll 1
        for (int ch = 0; ch < 3; ++ch) {
            float e[16];
            if (fast) {                                   // branch hoisted: one copy of each path keeps the kernel inside the I-cache
                float arg[16];
                embed_arg16(sm + T::S_B, pf, ch * 32 + 16 * hf, arg);           // columns 93..95 of the staged matrix are zero padding
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) {
                    const int k = ch * 32 + 16 * hf + kk;
                    e[kk] = (k < EMB && !DFP_EXP(a, 64)) ? __sinf(reduce_2pi(arg[kk])) : 0.f;
                }
            } else {
#pragma unroll 1
                for (int kk = 0; kk < 16; ++kk) {
                    const int k = ch * 32 + 16 * hf + kk;
                    const float v = (k < EMB) ? sinf(embed_arg(sm + T::S_B, pf, k)) : 0.f;
#pragma unroll
                    for (int u = 0; u < 16; ++u) if (u == kk) e[u] = v;
                }
            }
            PROF(5)
            if (ch > 0) { tc::mbar_wait(bar, parity); parity ^= 1; tc::fence_after_sync(); }   // previous chunk consumed A_X
            PROF(6)
            tc::tmem_store_split16(tl + T::T_X_HI, tl + T::T_X_LO, e);
            tc::wait_st();
            tc::fence_before_sync();
            PROF(7)
            tc::named_bar_sync(1 + grp, 256);
            PROF(8)
            if (issuer_warp && tc::elect_one()) {
                tc::fence_after_sync();
                const uint32_t koff = (uint32_t)ch * 8u * (64u * 16u);             // 8 K-chunks of 16 B per 32 k, LBO = N*16
                tc::mma_3xtf32(tb + T::T_D03, tb + T::T_X_HI, tb + T::T_X_LO, bhi + T::W03 * 4 + koff, blo + T::W03 * 4 + koff,
                               32, 64, ch > 0);
                if (ch == 2)                                                         // fc_c[0](c)
                    tc::mma_3xtf32(tb + T::T_DC, tb + T::T_C_HI, tb + T::T_C_LO, bhi + T::WC * 4, blo + T::WC * 4, CD, 32, false);
                tc::mma_commit(bar);
            }
        }
        PROF(9)
        tc::mbar_wait(bar, parity); parity ^= 1;
        tc::fence_after_sync();
        PROF(10)
        float h[16];
        uint32_t m_prev;                                              // ReLU bits of the block whose output `h` holds
        {
            float z[16], f[16];
            tc::tmem_ld16(tl + T::T_D03, z);
            tc::tmem_ld16(tl + T::T_DC, f);
            tc::wait_ld();
            uint32_t m = 0;
            float bp[16], bc[16];
            tc::lds16(sm + T::S_BP + 16 * hf, bp);
            tc::lds16(sm + T::S_BC + 16 * hf, bc);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float pre = z[j] + bp[j];
                m |= (pre > 0.f ? 1u : 0u) << j;
                h[j] = fmaxf(pre, 0.f) + (f[j] + bc[j]);
            }
            m_prev = m;
        }
        PROF(11)
#pragma unroll 1
        for (int i = 1; i < 5; ++i) {
            tc::tmem_store_split16(tl + T::T_X_HI, tl + T::T_X_LO, h);
            tc::wait_st();
            tc::fence_before_sync();
            PROF(12)
            tc::named_bar_sync(1 + grp, 256);
            PROF(13)
            if (issuer_warp && tc::elect_one()) {
                tc::fence_after_sync();
                tc::mma_3xtf32(tb + T::T_DH, tb + T::T_X_HI, tb + T::T_X_LO, bhi + (T::WH + (i - 1) * 1024) * 4,
                               blo + (T::WH + (i - 1) * 1024) * 4, 32, 32, false);
                tc::mma_3xtf32(tb + T::T_DC, tb + T::T_C_HI, tb + T::T_C_LO, bhi + (T::WC + i * CD * 32) * 4,
                               blo + (T::WC + i * CD * 32) * 4, CD, 32, false);
                tc::mma_commit(bar);
            }
            // while the block's products run: what the backward needs of the PREVIOUS block's output (ReLU bits, H plane) --
            // kept off the chain store -> barrier -> issue -> wait, which the tile's latency is made of
            if (a.act_m) {
                if (valid) reinterpret_cast<uint16_t*>(a.act_m + tm_index(ia, i - 1, 5))[hf] = (uint16_t)m_prev;
                if (save_pl) pl::store16(r_h, pl::h_pos(i - 1) * 32 + 16 * hf, h);
            }
            PROF(14)
            tc::mbar_wait(bar, parity); parity ^= 1;
            tc::fence_after_sync();
            PROF(15)
            float z[16], f[16];
            tc::tmem_ld16(tl + T::T_DH, z);
            tc::tmem_ld16(tl + T::T_DC, f);
            if (i == 3) {
                float z3[16];
                tc::tmem_ld16(tl + T::T_D03 + 32, z3);
                tc::wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) z[j] += z3[j];
            } else {
                tc::wait_ld();
            }
            PROF(16)
            uint32_t m = 0;
            float bp[16], bc[16];
            tc::lds16(sm + T::S_BP + i * 32 + 16 * hf, bp);
            tc::lds16(sm + T::S_BC + i * 32 + 16 * hf, bc);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float pre = z[j] + bp[j];
                m |= (pre > 0.f ? 1u : 0u) << j;
                h[j] = fmaxf(pre, 0.f) + (f[j] + bc[j]);
            }
            m_prev = m;
        }
        if (a.act_m) {                                                // the last block's output
            if (valid) reinterpret_cast<uint16_t*>(a.act_m + tm_index(ia, 4, 5))[hf] = (uint16_t)m_prev;
            if (save_pl) pl::store16(r_h, pl::h_pos(4) * 32 + 16 * hf, h);
        }
        PROF(17)
        // output layer: each half sums its 16 inputs, the lower half adds the upper half's partial and the bias
        float out[NOUT];
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
            float acc = 0.f, wo[16];
            tc::lds16(sm + T::S_WO + o * 32 + 16 * hf, wo);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc = fmaf(wo[j], h[j], acc);
            out[o] = acc;
        }
        if (hf == 1) {
#pragma unroll
            for (int o = 0; o < NOUT; ++o) part[grp][o][r] = out[o];
        }
        tc::named_bar_sync(1 + grp, 256);
        if (hf == 0 && valid) {
#pragma unroll
            for (int o = 0; o < NOUT; ++o) out[o] += part[grp][o][r] + sm[T::S_BO + o];
            if (NOUT == 1) a.out[s] = out[0];
            else *reinterpret_cast<float4*>(a.out + 4 * s) = make_float4(out[0], out[NOUT > 1 ? 1 : 0], out[NOUT > 2 ? 2 : 0], out[NOUT > 3 ? 3 : 0]);
        }
        PROF(18)
        PROF_TILE
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}

template <int DEC, int MODE>
__global__ void __launch_bounds__(TcLayout<(DEC == 1) ? 64 : 32>::GROUPS * 256, 1) k_decoder_fwd_tc(const DecArgs a) {
    decoder_fwd_tc_body<DEC, MODE>(a, blockIdx.x, gridDim.x);
}

// CTA-iterations a decoder needs: tiles / groups per CTA
template <int DEC>
__device__ __forceinline__ int64_t fwd_iterations(const DecArgs& a, int on) {
    if (!on) return 0;
    constexpr int G = TcLayout<(DEC == 1) ? 64 : 32>::GROUPS;
    const int64_t tiles = (dec_items(a) + 127) / 128;
    return (tiles + G - 1) / G;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_decoder_fwd_tc_all(const DecArgs3 A) {
    const int64_t it[3] = {fwd_iterations<1>(A.d[0], A.on[0]), fwd_iterations<2>(A.d[1], A.on[1]), fwd_iterations<0>(A.d[2], A.on[2])};
    int first, count;
    split_ctas(it, gridDim.x, 0, first, count);
    if ((int)blockIdx.x < first + count) { decoder_fwd_tc_body<1, MODE>(A.d[0], blockIdx.x - first, count); return; }
    split_ctas(it, gridDim.x, 1, first, count);
    if ((int)blockIdx.x < first + count) { decoder_fwd_tc_body<2, MODE>(A.d[1], blockIdx.x - first, count); return; }
    split_ctas(it, gridDim.x, 2, first, count);
    if ((int)blockIdx.x < first + count) decoder_fwd_tc_body<0, MODE>(A.d[2], blockIdx.x - first, count);
}

template <int MODE>
int launch_decoder_fwd_tc_all(const DecArgs3& a, cudaStream_t st) {
    const size_t smem = (TcLayout<64>::TOTAL > TcLayout<32>::TOTAL ? TcLayout<64>::TOTAL : TcLayout<32>::TOTAL) * sizeof(float);
    int rc = set_smem(k_decoder_fwd_tc_all<MODE>, smem);
    if (rc) return rc;
    int64_t its = 0;                                         // upper bound (the high decoder's list may be shorter)
    for (int d = 0; d < 3; ++d) if (a.on[d]) { const int64_t tiles = (a.d[d].n + 127) / 128; its += d == 0 ? tiles : (tiles + 1) / 2; }
    const int blocks = (int)(its < (int64_t)num_sms() ? (its < 1 ? 1 : its) : num_sms());
    DFP_TIMED("k_decoder_fwd_tc_all", st, k_decoder_fwd_tc_all<MODE><<<blocks, 512, smem, st>>>(a));
    DFP_LAUNCH_CHECK("k_decoder_fwd_tc_all");
    return DFP_OK;
}
template int launch_decoder_fwd_tc_all<0>(const DecArgs3&, cudaStream_t);
template int launch_decoder_fwd_tc_all<1>(const DecArgs3&, cudaStream_t);
template int launch_decoder_fwd_tc_all<2>(const DecArgs3&, cudaStream_t);

template <int DEC, int MODE>
int launch_decoder_fwd_tc(const DecArgs& a, cudaStream_t st) {
    constexpr int CD = (DEC == 1) ? 64 : 32;
    using T = TcLayout<CD>;
    const size_t smem = T::TOTAL * sizeof(float);
    int rc = set_smem(k_decoder_fwd_tc<DEC, MODE>, smem);
    if (rc) return rc;
    const int64_t tiles = (a.n + 127) / 128;
    const int blocks = grid_for(tiles, T::GROUPS, num_sms());
    DFP_TIMED("k_decoder_fwd_tc", st, k_decoder_fwd_tc<DEC, MODE><<<blocks, T::GROUPS * 256, smem, st>>>(a));
    DFP_LAUNCH_CHECK("k_decoder_fwd_tc");
    return DFP_OK;
}
template int launch_decoder_fwd_tc<0, 0>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd_tc<1, 0>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd_tc<2, 0>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd_tc<0, 1>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd_tc<1, 1>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd_tc<2, 1>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd_tc<0, 2>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd_tc<1, 2>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd_tc<2, 2>(const DecArgs&, cudaStream_t);

}  // namespace dfp
