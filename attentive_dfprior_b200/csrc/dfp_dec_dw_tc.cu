// Tensor-core decoder backward, kernel 2 of 2: every parameter gradient of one decoder as outer products over
// samples,  dW = A^T * Delta,  on the tcgen05 tensor cores, fed by bulk copies.
//
// The forward and kernel 1 leave their per-sample operands in global memory as bf16 hi/mid "planes" already in the
// tensor cores' MN-major operand layout (dfp_planes.cuh).  This kernel therefore never touches a value in registers:
// one thread streams 32-sample slices of the planes into a shared-memory ring (as many stages as fit 216 KB: 4 to 6) with cp.async.bulk (completion
// counted in bytes on an mbarrier), one thread issues SS-mode tcgen05.mma chains whose accumulators stay resident in
// TMEM for the CTA's lifetime, and tcgen05.commit hands the stage back.  (The previous version staged fp32
// feature-major scratch through registers -- 32 scalar loads + tf32 split + 16 shared stores per thread and block --
// and was bound by the load/store unit, not by the MMAs: profiles/r01, "K2 without MMAs" experiment.)
//
// Products per 16 samples (each as hi*hi' + hi*mid' + mid*hi'), by pass:
//   pass 0   P : H[0:128)   x DL0[8:136)   M=128 N=128  diagonal 32x32 blocks = pts_linears.4/.2/.1/.3 (hidden part)
//            O : H[128:192) x DL0[0:104)   M=64  N=104  rows 0..31 x dout = output_linear; the constant-1 row gives
//                                                       output_linear.bias and pts_linears.4/.2/.1 biases
//            C : CX[0:..)   x DL0[0:136)   M=64/128 N=136/144  -> scratch; k_fcc_finish forms fc_c.0..4 weights and biases from it
//   pass 1   E : EX[0:128)  x DL1[0:160)   M=128 N=160  rows 0..92 x dacc3|dacc0 = pts_linears.3 (embedding part) / .0,
//                                                       row 93 (constant 1) = their biases, rows 96..98 x darg = embedder._B
// Feature ranges that run past an array (H[168:192), CX beyond CD+8, EX[104:128)) read whatever finite plane data
// follows in the stage and only produce accumulator rows nobody reads.  One atomic flush per CTA and pass.
// DL0 and DL1 are two overlapping feature ranges of ONE stored array (DLX, dfp_planes.cuh: dacc3 is stored once), copied
// group by group; the constant-one chunks of H and CX are not stored at all: they are written into the stages once per CTA
// and the copies go around them.  The tile list comes from the per-tile flags kernel 1 leaves (a scan of the upstream
// gradients here would delay the first copy by ~0.7 us per tile).
#include "dfp_device.cuh"
#include "dfp_planes.cuh"
#include "dfp_tc.cuh"

namespace dfp {

namespace dw {
constexpr int RING_BYTES = 216 * 1024;        // shared memory given to the ring: as many stages as fit (bytes in flight per SM bound the HBM rate)
template <int CD, int PASS> __host__ __device__ constexpr int nst();
constexpr int QS = 32;                        // samples per stage (4 groups of 8)
constexpr int MAX_TILES = 1024;               // active tiles of one CTA
constexpr int THREADS = 192;                  // warp 0: tile list + producer, warp 1: MMA issuer, warps 2..5: flush
constexpr int FCC_N = 144;                    // row length of the [c | 1]^T [dout8 | dacc4 | dacc2 | dacc1 | dacc3] scratch (136 used)
__host__ __device__ constexpr uint32_t piece(int F) { return 64u * (uint32_t)F; }          // one plane, 32 samples
template <int CD, int PASS> __host__ __device__ constexpr uint32_t stage_bytes() {          // shared-memory size of a stage
    return PASS == 0 ? 2u * (piece(pl::F_H) + piece(pl::f_cx(CD)) + piece(pl::F_DL0)) : 2u * (piece(pl::F_EX) + piece(pl::F_DL1));
}
template <int CD, int PASS> __host__ __device__ constexpr uint32_t copied_bytes() {         // bytes the bulk copies bring per stage (no constant chunks)
    return PASS == 0 ? 2u * (piece(pl::F_H_ST) + piece(CD) + piece(pl::F_DL0)) : 2u * (piece(pl::F_EX) + piece(pl::F_DL1));
}
template <int CD, int PASS> __host__ __device__ constexpr int nst() { return RING_BYTES / (int)stage_bytes<CD, PASS>(); }
}  // namespace dw

// state shared by the two passes' code (one CTA runs exactly one pass)
struct DwShared {
    uint64_t full_bar[8], empty_bar[8], done_bar;
    uint32_t tmem_base;
    int n_active;
    int tile_list[dw::MAX_TILES];
};

// One pass for CTA `cta` of `nctas` CTAs working on it (the kernel gives the first n0 CTAs of the grid to pass 0 and the
// rest to pass 1, so that both passes of a decoder run in ONE launch: at the reference's batch sizes a launch is mostly
// fixed cost, and on large batches the two HBM streams share the machine in proportion to their bytes).
template <int DEC, int PASS>
__device__ __forceinline__ void dw_body(const DecArgs& a, const int cta, const int nctas, DwShared& S) {
    constexpr int CD = (DEC == 1) ? 64 : 32;
    constexpr int NOUT = (DEC == 2) ? 4 : 1;
    constexpr int FCX = pl::f_cx(CD);
    constexpr int MC = (CD == 32) ? 64 : 128;                        // M of the fc_c product
    using L = DecOff<CD>;
    constexpr uint32_t STAGE = dw::stage_bytes<CD, PASS>();
    constexpr int NST = dw::nst<CD, PASS>();
    // byte offsets of the pieces inside a stage
    constexpr uint32_t O_H = 0, O_CX = O_H + 2 * dw::piece(pl::F_H), O_DL0 = O_CX + 2 * dw::piece(FCX);
    constexpr int NC = (MC == 64) ? dw::FCC_N - 8 : dw::FCC_N;       // columns of the CX^T DL0 product (M = 128 needs N % 16 == 0)
    constexpr uint32_t O_EX = 0, O_DL1 = O_EX + 2 * dw::piece(pl::F_EX);
    // TMEM columns
    constexpr uint32_t T_P = 0, T_O = 128, T_C = 232, T_E = 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = S.full_bar; uint64_t* empty_bar = S.empty_bar; uint64_t& done_bar = S.done_bar;
    uint32_t& tmem_base_s = S.tmem_base;
    int* tile_list = S.tile_list;
    int& n_active_s = S.n_active;
    static_assert(NST <= 8, "ring deeper than the barrier arrays");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n = dec_items(a);
    const int64_t tiles = (n + 127) / 128;
    if (cta >= tiles) return;                                        // nothing for this CTA (short band list, small batch)

    // the stage ring must never hold NaN bit patterns where the over-long feature ranges read: clear it once
    for (uint32_t i = tid; i < NST * STAGE / 16; i += dw::THREADS) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (PASS == 0) {
        // the constant-one chunks of H and CX ([1 0 0 0 0 0 0 0] per sample in the hi plane, zero in the mid plane) live
        // only here: the copies below never touch them
        __syncthreads();
        for (uint32_t i = tid; i < (uint32_t)NST * 32u; i += dw::THREADS) {
            const uint32_t st = i >> 5, g8 = (i >> 3) & 3, k = i & 7;
            unsigned char* s0 = sm + st * STAGE;
            *reinterpret_cast<uint4*>(s0 + O_H + g8 * 16u * pl::F_H + (pl::F_H_ST / 8) * 128u + k * 16u) = make_uint4(0x00003f80u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(s0 + O_CX + g8 * 16u * FCX + (CD / 8) * 128u + k * 16u) = make_uint4(0x00003f80u, 0u, 0u, 0u);
        }
    }
    if (tid == 0) n_active_s = 0;
    __syncthreads();
    // the CTA's tiles that carry any upstream gradient (kernel 1 wrote its planes only for those, and left one flag per
    // tile), in order.  (Scanning the tiles' upstream gradients here instead cost ~0.7 us per tile in ONE warp before the
    // first bulk copy could be issued: 0.2 ms of every launch at 65 536 rays.)
    if (warp == 0) {
        int cnt = 0;
        for (int64_t t0 = cta; t0 < tiles && cnt < dw::MAX_TILES; t0 += (int64_t)nctas * 32) {
            const int64_t tile = t0 + (int64_t)lane * nctas;
            const bool on = tile < tiles && a.tileflag[tile] != 0;
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) { const int at = cnt + __popc(m & ((1u << lane) - 1)); if (at < dw::MAX_TILES) tile_list[at] = (int)tile; }
            cnt += __popc(m);
        }
        if (lane == 0) n_active_s = cnt < dw::MAX_TILES ? cnt : dw::MAX_TILES;
    }
    if (warp == 1) { tc::tmem_alloc(&tmem_base_s, 512); tc::tmem_relinquish(); }
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
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
    const uint32_t total = 4u * (uint32_t)n_active;                  // stages: (tile, quarter)

    if (n_active > 0) {
        if (warp == 0) {
            // ===================== producer: bulk copies of plane slices =====================
            if (tc::elect_one()) {
                const unsigned char* gH = reinterpret_cast<const unsigned char*>(a.act_h);
                const unsigned char* gCX = reinterpret_cast<const unsigned char*>(a.act_c);
                const unsigned char* gDL0 = reinterpret_cast<const unsigned char*>(a.sc_darg);
                const unsigned char* gEX = reinterpret_cast<const unsigned char*>(a.sc_e);
                for (uint32_t q = 0; q < total; ++q) {
                    const int st = q % NST;
                    if (q >= NST) tc::mbar_wait(&empty_bar[st], ((q / NST) - 1) & 1);     // MMAs of the previous use are done
                    const int64_t tile = tile_list[q >> 2];
                    const uint32_t quarter = q & 3;
                    const uint32_t dst = sm_base + st * STAGE;
                    tc::mbar_arrive_expect_tx(&full_bar[st], dw::copied_bytes<CD, PASS>());
                    auto copy = [&](const unsigned char* g, int F, uint32_t off) {
                        for (int plane = 0; plane < 2; ++plane)
                            tc::bulk_copy_g2s(dst + off + plane * dw::piece(F), g + (tile * 2 + plane) * (int64_t)(256 * F) + quarter * dw::piece(F),
                                              dw::piece(F), &full_bar[st]);
                    };
                    // feature range [f0, f0 + F) of an array stored with FS features per sample: per 8-sample group the range
                    // is contiguous (16 F bytes), the groups are 16 FS apart; in shared memory the groups are FSM features apart
                    auto copy_range = [&](const unsigned char* g, int FS, int f0, int F, uint32_t off, int FSM) {
                        for (int plane = 0; plane < 2; ++plane)
                            for (int grp8 = 0; grp8 < 4; ++grp8)
                                tc::bulk_copy_g2s(dst + off + plane * dw::piece(FSM) + grp8 * 16u * FSM,
                                                  g + (tile * 2 + plane) * (int64_t)(256 * FS) + (quarter * 4 + grp8) * (16 * FS) + f0 * 16,
                                                  16u * F, &full_bar[st]);
                    };
                    if (PASS == 0) {
                        copy_range(gH, pl::F_H_ST, 0, pl::F_H_ST, O_H, pl::F_H);              // the stage keeps its constant chunk
                        copy_range(gCX, CD, 0, CD, O_CX, FCX);
                        copy_range(gDL0, pl::F_DLX, 0, pl::F_DL0, O_DL0, pl::F_DL0);
                    } else {
                        copy(gEX, pl::F_EX, O_EX);
                        copy_range(gDL0, pl::F_DLX, pl::DL1_AT, pl::F_DL1, O_DL1, pl::F_DL1);
                    }
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            if (tc::elect_one()) {
                for (uint32_t q = 0; q < total; ++q) {
                    const int st = q % NST;
                    tc::mbar_wait(&full_bar[st], (q / NST) & 1);
                    tc::fence_after_sync();
                    const uint32_t s0 = sm_base + st * STAGE;
                    auto prod = [&](uint32_t d_col, int M, int N, uint32_t a_off, int FA, int fa0, uint32_t b_off, int FB, int fb0) {
                        const uint32_t idesc = tc::idesc_bf16_mn(M, N);
                        const uint32_t ga = 16u * FA, gb = 16u * FB;                       // bytes per 8-sample group
                        for (uint32_t ks = 0; ks < 2; ++ks) {                              // 16 samples = two groups per MMA
                            const uint32_t ah = s0 + a_off + (fa0 >> 3) * 128 + ks * 2 * ga, bh = s0 + b_off + (fb0 >> 3) * 128 + ks * 2 * gb;
                            tc::mma_planes(tb + d_col, ah, ah + dw::piece(FA), ga, bh, bh + dw::piece(FB), gb, idesc, q > 0 || ks > 0);
                        }
                    };
                    if (PASS == 0) {
                        if (!DFP_EXP(a, 256)) prod(T_P, 128, 128, O_H, pl::F_H, 0, O_DL0, pl::F_DL0, 8);
                        if (!DFP_EXP(a, 128)) prod(T_O, 64, 104, O_H, pl::F_H, 128, O_DL0, pl::F_DL0, 0);
                        if (!DFP_EXP(a, 128)) prod(T_C, MC, NC, O_CX, FCX, 0, O_DL0, pl::F_DL0, 0);
                    } else {
                        if (!DFP_EXP(a, 256)) prod(T_E, 128, 160, O_EX, pl::F_EX, 0, O_DL1, pl::F_DL1, 0);
                    }
                    tc::mma_commit(&empty_bar[st]);                      // stage reusable once these MMAs have read it
                }
                tc::mma_commit(&done_bar);                               // everything accumulated
            }
        } else {
            // ===================== flush (warps 2..5 = TMEM lane quarters 2, 3, 0, 1), once per CTA =====================
            tc::mbar_wait(&done_bar, 0);
            tc::fence_after_sync();
            float* g = a.gparams + (DEC == 0 ? OFF_LOW : (DEC == 1 ? OFF_HIGH : OFF_COLOR));
            const int qtr = warp & 3;
            const uint32_t tl = tb + ((uint32_t)(qtr * 32) << 16);
            const int row128 = qtr * 32 + lane;                          // accumulator row of an M = 128 product
            const int row64 = lane < 16 ? qtr * 16 + lane : -1;          // ... of an M = 64 product (lane (m/16)*32 + m%16)
            float d[32];
            if (PASS == 0) {
                {   // P: this warp's rows are H block `qtr`; only the matching delta block is a parameter
                    tc::tmem_ld32(tl + T_P + qtr * 32, d);
                    tc::wait_ld();
                    const int li = qtr == 0 ? 4 : qtr == 1 ? 2 : qtr == 2 ? 1 : 3;
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(g + L::WH + (li - 1) * 1024 + lane * 32 + j, d[j]);
                }
                {   // O: columns [dout8 | dacc4 | dacc2 | dacc1]
                    float d8[8];
                    {
                        uint32_t v[8];
                        tc::tmem_ld8(tl + T_O, v);
                        tc::wait_ld();
#pragma unroll
                        for (int j = 0; j < 8; ++j) d8[j] = __uint_as_float(v[j]);
                    }
                    if (row64 >= 0 && row64 < 32) {
#pragma unroll
                        for (int o = 0; o < NOUT; ++o) atomicAdd(g + L::WO + o * 32 + row64, d8[o]);
                    } else if (row64 == 32) {
#pragma unroll
                        for (int o = 0; o < NOUT; ++o) atomicAdd(g + L::BO + o, d8[o]);
                    }
#pragma unroll 1
                    for (int b = 0; b < 3; ++b) {
                        tc::tmem_ld32(tl + T_O + 8 + b * 32, d);
                        tc::wait_ld();
                        const int li = b == 0 ? 4 : b == 1 ? 2 : 1;
                        if (row64 == 32) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) atomicAdd(g + L::BP + li * 32 + j, d[j]);
                        }
                    }
                }
                // C: [c | 1]^T [dout8 | dacc4 | dacc2 | dacc1 | dacc3] goes to the scratch; k_fcc_finish turns it into the
                // fc_c.i gradients (c^T dh_i = (c^T dacc_{i+1}) W_{i+1}: the dh planes are never materialised)
                {
                    const int row = (MC == 64) ? row64 : row128;
#pragma unroll 1
                    for (int b = 0; b < 5; ++b) {
                        float d8[8];
                        if (b < 4) { tc::tmem_ld32(tl + T_C + b * 32, d); }
                        else { uint32_t v[8]; tc::tmem_ld8(tl + T_C + 128, v);
#pragma unroll
                               for (int j = 0; j < 8; ++j) d8[j] = __uint_as_float(v[j]); }
                        tc::wait_ld();
                        if (row >= 0 && row <= CD) {
                            if (b < 4) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) atomicAdd(a.fcc + row * dw::FCC_N + b * 32 + j, d[j]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) atomicAdd(a.fcc + row * dw::FCC_N + 128 + j, d8[j]);
                            }
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int b = 0; b < 5; ++b) {   // E: columns [dacc3 | dacc0 | darg 0..95]
                    tc::tmem_ld32(tl + T_E + b * 32, d);
                    tc::wait_ld();
                    const int m = row128;
                    if (b < 2) {
                        const int coff = b == 0 ? 32 : 0, li = b == 0 ? 3 : 0;
                        if (m < EMB) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) atomicAdd(g + L::W03 + m * 64 + coff + j, d[j]);
                        } else if (m == EMB) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) atomicAdd(g + L::BP + li * 32 + j, d[j]);
                        }
                    } else if (m >= 96 && m < 99) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if ((b - 2) * 32 + j < EMB) atomicAdd(g + L::B + (m - 96) * EMB_PAD + (b - 2) * 32 + j, d[j]);
                    }
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tb, 512);
}

// fc_c.i gradients from the accumulated G = [c | 1]^T [dout8 | dacc4 | dacc2 | dacc1 | dacc3]  (rows: feature c, last row the
// constant 1):  dh_4 = dout W_o and dh_{i-1} = dacc_i W_i (hidden part of W_3 for i = 3), hence
//   d fc_c.4[c][k] = sum_o G[c][o] W_o[o][k],   d fc_c.(i-1)[c][k] = sum_j G[c][col_i + j] W_i[j][k],   biases from the last row.
// grid (5 layers, CD + 1 rows), 32 threads = output units k.
template <int CD, int NOUT>
__global__ void k_fcc_finish(const float* __restrict__ fcc, const float* __restrict__ P /* packed decoder params */, float* __restrict__ g) {
    using L = DecOff<CD>;
    const int layer = blockIdx.x, row = blockIdx.y, k = threadIdx.x;
    const float* G = fcc + row * dw::FCC_N;
    float acc = 0.f;
    if (layer == 4) {
#pragma unroll
        for (int o = 0; o < NOUT; ++o) acc = fmaf(G[o], P[L::WO + o * 32 + k], acc);
    } else {
        const int i = layer + 1;                                  // the block whose masked delta produced dh_layer
        const int col0 = i == 4 ? 8 : i == 2 ? 40 : i == 1 ? 72 : 104;
        const float* W = P + L::WH + (i - 1) * 1024 + k * 32;     // packed k-major: WH[slot][kin = k][j]
#pragma unroll 8
        for (int j = 0; j < 32; ++j) acc = fmaf(G[col0 + j], W[j], acc);
    }
    if (row < CD) g[L::WC + layer * CD * 32 + row * 32 + k] += acc;
    else g[L::BC + layer * 32 + k] += acc;
}

template <int DEC>
__global__ void __launch_bounds__(dw::THREADS, 1) k_decoder_dw_tc(const DecArgs a, const int n0) {
    __shared__ __align__(8) DwShared S;
    if ((int)blockIdx.x < n0) dw_body<DEC, 0>(a, blockIdx.x, n0, S);
    else dw_body<DEC, 1>(a, blockIdx.x - n0, gridDim.x - n0, S);
}

// fc_c finish of up to three decoders in one launch (grid: 5 layers x 65 rows x 3 decoders, order high / colour / low)
struct FccArgs3 { const float* fcc[3]; const float* params; float* gparams[3]; int on[3]; };
template <int CD, int NOUT>
__device__ __forceinline__ void fcc_finish_body(const float* __restrict__ fcc, const float* __restrict__ P, float* __restrict__ g, int layer, int row, int k) {
    using L = DecOff<CD>;
    if (row > CD) return;
    const float* G = fcc + row * dw::FCC_N;
    float acc = 0.f;
    if (layer == 4) {
#pragma unroll
        for (int o = 0; o < NOUT; ++o) acc = fmaf(G[o], P[L::WO + o * 32 + k], acc);
    } else {
        const int i = layer + 1;
        const int col0 = i == 4 ? 8 : i == 2 ? 40 : i == 1 ? 72 : 104;
        const float* W = P + L::WH + (i - 1) * 1024 + k * 32;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) acc = fmaf(G[col0 + j], W[j], acc);
    }
    if (row < CD) g[L::WC + layer * CD * 32 + row * 32 + k] += acc;
    else g[L::BC + layer * 32 + k] += acc;
}
__global__ void k_fcc_finish_all(const FccArgs3 a) {
    const int d = blockIdx.z;
    if (!a.on[d]) return;
    if (d == 0) fcc_finish_body<64, 1>(a.fcc[0], a.params + OFF_HIGH, a.gparams[0] + OFF_HIGH, blockIdx.x, blockIdx.y, threadIdx.x);
    else if (d == 1) fcc_finish_body<32, 4>(a.fcc[1], a.params + OFF_COLOR, a.gparams[1] + OFF_COLOR, blockIdx.x, blockIdx.y, threadIdx.x);
    else fcc_finish_body<32, 1>(a.fcc[2], a.params + OFF_LOW, a.gparams[2] + OFF_LOW, blockIdx.x, blockIdx.y, threadIdx.x);
}

// Both passes of up to three decoders in one launch: bounds[i] = first CTA of (decoder i/2, pass i%2), order high / colour / low
struct DwSplit { int first[7]; };
__global__ void __launch_bounds__(dw::THREADS, 1) k_decoder_dw_tc_all(const DecArgs3 A, const DwSplit sp) {
    __shared__ __align__(8) DwShared S;
    const int b = blockIdx.x;
    if (b < sp.first[1]) dw_body<1, 0>(A.d[0], b - sp.first[0], sp.first[1] - sp.first[0], S);
    else if (b < sp.first[2]) dw_body<1, 1>(A.d[0], b - sp.first[1], sp.first[2] - sp.first[1], S);
    else if (b < sp.first[3]) dw_body<2, 0>(A.d[1], b - sp.first[2], sp.first[3] - sp.first[2], S);
    else if (b < sp.first[4]) dw_body<2, 1>(A.d[1], b - sp.first[3], sp.first[4] - sp.first[3], S);
    else if (b < sp.first[5]) dw_body<0, 0>(A.d[2], b - sp.first[4], sp.first[5] - sp.first[4], S);
    else if (b < sp.first[6]) dw_body<0, 1>(A.d[2], b - sp.first[5], sp.first[6] - sp.first[5], S);
}

int launch_decoder_dw_tc_all(const DecArgs3& a, cudaStream_t st) {
    auto smem_of = [](auto cd_tag) {
        constexpr int CD = decltype(cd_tag)::value;
        constexpr size_t s0 = dw::nst<CD, 0>() * dw::stage_bytes<CD, 0>() + 1024 + 256, s1 = dw::nst<CD, 1>() * dw::stage_bytes<CD, 1>() + 1024 + 256;
        return s0 > s1 ? s0 : s1;
    };
    const size_t sm32 = smem_of(std::integral_constant<int, 32>{}), sm64 = smem_of(std::integral_constant<int, 64>{});
    const size_t smem = sm32 > sm64 ? sm32 : sm64;
    int rc = set_smem(k_decoder_dw_tc_all, smem);
    if (rc) return rc;
    // CTAs in proportion to (tiles x bytes per tile) of each (decoder, pass); at least one where there is work
    double wgt[6]; double total = 0;
    int64_t tiles[3];
    for (int d = 0; d < 3; ++d) {
        tiles[d] = a.on[d] ? (a.d[d].n + 127) / 128 : 0;
        const double b0 = d == 0 ? (double)dw::copied_bytes<64, 0>() : (double)dw::copied_bytes<32, 0>();
        const double b1 = d == 0 ? (double)dw::copied_bytes<64, 1>() : (double)dw::copied_bytes<32, 1>();
        wgt[2 * d] = tiles[d] * b0; wgt[2 * d + 1] = tiles[d] * b1;
        total += wgt[2 * d] + wgt[2 * d + 1];
    }
    if (total <= 0) return DFP_OK;
    const int G = num_sms();
    DwSplit sp;
    int at = 0;
    for (int i = 0; i < 6; ++i) {
        sp.first[i] = at;
        if (wgt[i] <= 0) continue;
        int64_t c = (int64_t)(G * wgt[i] / total + 0.5);
        if (c < 1) c = 1;
        if (c > tiles[i / 2]) c = tiles[i / 2];
        const int64_t min_blocks = (tiles[i / 2] + dw::MAX_TILES - 1) / dw::MAX_TILES;
        if (c < min_blocks) c = min_blocks;
        at += (int)c;
    }
    sp.first[6] = at;
    DFP_TIMED("k_decoder_dw_tc_all", st, k_decoder_dw_tc_all<<<at, dw::THREADS, smem, st>>>(a, sp));
    DFP_LAUNCH_CHECK("k_decoder_dw_tc_all");
    FccArgs3 f;
    memset(&f, 0, sizeof(f));
    for (int d = 0; d < 3; ++d) { f.fcc[d] = a.d[d].fcc; f.gparams[d] = a.d[d].gparams; f.on[d] = a.on[d]; f.params = a.on[d] ? a.d[d].params : f.params; }
    DFP_TIMED("k_fcc_finish_all", st, k_fcc_finish_all<<<dim3(5, 65, 3), 32, 0, st>>>(f));
    DFP_LAUNCH_CHECK("k_fcc_finish_all");
    return DFP_OK;
}

template <int DEC>
int launch_decoder_dw_tc(const DecArgs& a, cudaStream_t st) {
    constexpr int CD = (DEC == 1) ? 64 : 32;
    constexpr size_t smem0 = dw::nst<CD, 0>() * dw::stage_bytes<CD, 0>() + 1024 + 256, smem1 = dw::nst<CD, 1>() * dw::stage_bytes<CD, 1>() + 1024 + 256;   // + alignment + over-read slack
    constexpr size_t smem = smem0 > smem1 ? smem0 : smem1;
    int rc = set_smem(k_decoder_dw_tc<DEC>, smem);
    if (rc) return rc;
    const int64_t tiles = (a.n + 127) / 128;
    // CTAs per pass in proportion to the bytes a tile costs in it (pass 0: H + CX + DL0, pass 1: EX + DL1)
    constexpr double share0 = (double)dw::stage_bytes<CD, 0>() / (double)(dw::stage_bytes<CD, 0>() + dw::stage_bytes<CD, 1>());
    int blocks = grid_for(2 * tiles, 1, num_sms());
    if (blocks < 2) blocks = 2;
    int n0 = (int)(blocks * share0 + 0.5);
    if (n0 < 1) n0 = 1;
    if (n0 > blocks - 1) n0 = blocks - 1;
    int n1 = blocks - n0;
    const int min_blocks = (int)((tiles + dw::MAX_TILES - 1) / dw::MAX_TILES);   // a CTA lists at most MAX_TILES tiles
    if (n0 < min_blocks) n0 = min_blocks;
    if (n1 < min_blocks) n1 = min_blocks;
    DFP_TIMED("k_decoder_dw_tc", st, k_decoder_dw_tc<DEC><<<n0 + n1, dw::THREADS, smem, st>>>(a, n0));
    DFP_LAUNCH_CHECK("k_decoder_dw_tc");
    constexpr int NOUT = (DEC == 2) ? 4 : 1;
    const int off = DEC == 0 ? OFF_LOW : (DEC == 1 ? OFF_HIGH : OFF_COLOR);
    DFP_TIMED("k_fcc_finish", st, k_fcc_finish<CD, NOUT><<<dim3(5, CD + 1), 32, 0, st>>>(a.fcc, a.params + off, a.gparams + off));
    DFP_LAUNCH_CHECK("k_fcc_finish");
    return DFP_OK;
}
template int launch_decoder_dw_tc<0>(const DecArgs&, cudaStream_t);
template int launch_decoder_dw_tc<1>(const DecArgs&, cudaStream_t);
template int launch_decoder_dw_tc<2>(const DecArgs&, cudaStream_t);

}  // namespace dfp
