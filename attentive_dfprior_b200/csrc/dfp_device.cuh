// Internal header shared by the translation units of libdfprior_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/dfprior_b200.h"
#include "dfp_math.h"

namespace dfp {

// ---- error plumbing (defined in dfp_api.cu)
int fail(int code, const char* fmt, ...);
#define DFP_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(DFP_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
#define DFP_LAUNCH_CHECK(name)                                                                              \
    do {                                                                                                    \
        cudaError_t e_ = cudaGetLastError();                                                                \
        if (e_ != cudaSuccess) return fail(DFP_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    } while (0)

int num_sms();

// launch accounting / optional per-kernel CUDA-event timing (dfp_launch_count, dfp_timing_*)
void timing_begin(const char* name, cudaStream_t st);
void timing_end(cudaStream_t st);
#define DFP_TIMED(name, st, ...)     \
    do {                             \
        ::dfp::timing_begin(name, st); \
        __VA_ARGS__;                 \
        ::dfp::timing_end(st);       \
    } while (0)

constexpr int NT = 256;            // threads per CTA of the forward / attention per-sample kernels
constexpr int WARPS = NT / 32;
constexpr int NT_BWD = 256;        // decoder backward: 8 warps, 1 CTA/SM (512 threads at 128 registers measured 50% slower)
constexpr int WARPS_BWD = NT_BWD / 32;
constexpr int SINK_FLOATS_PER_WARP = 32 * 32 + 32 * 33;

// ------------------------------------------------------------------------------------------
// device-side argument blocks
// ------------------------------------------------------------------------------------------
struct PointSrc {                  // where sample s comes from
    const float* rays_o;           // MODE 0: p = o + d * z[s], ray = s / S
    const float* rays_d;
    const double* z;
    int S;
    const void* pts;               // MODE 1: f64 [n,3]; MODE 2: f32 [n,3]
    int pts_f64;                   // dtype tag of pts for kernels that are not templated on MODE
};
template <int MODE> struct PtT { using T = double; };
template <> struct PtT<2> { using T = float; };

template <int MODE>
__device__ __forceinline__ void load_point(const PointSrc& ps, int64_t s, typename PtT<MODE>::T p[3]) {
    if (MODE == 0) {
        const int64_t r = s / ps.S;
        const float o[3] = {ps.rays_o[3 * r], ps.rays_o[3 * r + 1], ps.rays_o[3 * r + 2]};
        const float d[3] = {ps.rays_d[3 * r], ps.rays_d[3 * r + 1], ps.rays_d[3 * r + 2]};
        double q[3];
        ray_point(o, d, ps.z[s], q);
        p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
    } else if (MODE == 1) {
        const double* q = reinterpret_cast<const double*>(ps.pts) + 3 * s;
        p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
    } else {
        const float* q = reinterpret_cast<const float*>(ps.pts) + 3 * s;
        p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
    }
}

struct Bounds { double b[3][2]; };

// Saved activations and backward scratch are tile-major: element (feature f, sample s) of an F-feature array lives at
// ((s/128)*F + f)*128 + s%128, so the 32-feature operand block of one 128-sample tile is 16 KB contiguous in HBM
// (feature-major [f][n] put consecutive features ~12 MB apart and thrashed DRAM pages).
__host__ __device__ __forceinline__ int64_t tm_index(int64_t s, int f, int F) { return ((s >> 7) * F + f) * 128 + (s & 127); }

struct Ws {                        // per-sample workspace arrays (NS samples)
    double* z;                     // [NS] sorted sample depths (render only)
    float* depthmax;               // [nchunks][2]: max(gt_depth*1.2), max(gt_depth)
    float* low; float* high; float* tsdf; float* fused; float* w_attn;   // [NS]
    float* rgb;                    // [NS][4]  colour decoder output (rgb + its unused 4th channel)
    float4* xnpf;                  // [NS][2]: (point normalised to the scene bound, 0) and ((float)point, 0), written by k_prep
    int* flags;                    // [NS] bit0 in-bound, bit1 in TSDF band
    int* band_list;                // [NS] compacted band sample ids
    int* band_count;               // [4]
    // backward scratch
    float* d_low; float* d_high; float* d_fused;    // [NS]
    float* d_rgb;                  // [NS][4]
    float* dp;                     // [NS][3] gradient w.r.t. the sample point
    // DFP_FLAG_SAVE_ACT: activations of the three decoders kept by the tensor-core forward, tile-major (tm_index)
    // bf16 plane arrays (dfp_planes.cuh; the float* type is only a handle)
    float* act_h[3];               // H:  inputs of pts_linears.1..4 and of output_linear (160 features; the constant chunk is synthesised in K2)
    uint32_t* act_m[3];            // [5][NS] tile-major: ReLU masks of pts_linears.0..4
    float* act_c[3];               // CX: gathered grid features (CD features; likewise)
    // scratch shared by the decoders during the tensor-core backward
    float* sc_e[3];                // EX:  sin(embedding argument), point.  Per decoder; the three entries alias ONE buffer unless the
                                   //      call is small enough for the merged launches (then the decoders' K1 run concurrently)
    float* fcc[3];                 // per decoder: accumulated [c | 1]^T [dout8 | dacc..] ((CD + 1) x 144 floats, zeroed before the backward)
    float* sc_darg[3];             // DLX: upstream gradient, masked deltas, gradient of the embedding arguments (same aliasing rule)
    int* dec_tileflag[3];          // [tiles]: 1 if the 128-sample tile carries any upstream gradient (written by K1, read by K2 of the same decoder)
    float* att_act;                // SIMT path: fusion MLP activations per band-list position (AttArgs::act)
    unsigned char* att_planes; uint32_t* att_masks; float* att_soft; unsigned char* att_dplanes; int* att_tileflag;   // tensor-core path (AttArgs::tc_*)
};


// ------------------------------------------------------------------------------------------
// trilinear feature gather / scatter on a channels-last grid
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void gather_features(const DfpGrid& g, const Tri& t, float* c) {
    float acc[CDIM];
#pragma unroll
    for (int j = 0; j < CDIM; ++j) acc[j] = 0.f;
    // two passes of four corners: the four 128-byte corner reads of a pass are independent, so their
    // loads are all in flight before the first FMA (the serial 8-corner loop was latency-bound)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const float4* src[4];
        float wgt[4];
        bool ok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int x, y, z;
            const int cn = half * 4 + i;
            ok[i] = tri_corner(t, cn, g.W, g.H, g.D, x, y, z);      // index == size: skipped (GridSampler.cuh:225)
            wgt[i] = tri_weight(t, cn);
            src[i] = reinterpret_cast<const float4*>(g.data + (ok[i] ? (((int64_t)z * g.H + y) * g.W + x) * CDIM : 0));
        }
#pragma unroll
        for (int q = 0; q < CDIM / 4; ++q) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = ok[i] ? __ldg(src[i] + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[4 * q + 0] = fmaf(v[i].x, wgt[i], acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(v[i].y, wgt[i], acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(v[i].z, wgt[i], acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(v[i].w, wgt[i], acc[4 * q + 3]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CDIM; ++j) c[j] = acc[j];
}

// 16-channel half (channels 16*hf .. 16*hf+15) of gather_features, for kernels with two threads per sample
__device__ __forceinline__ void gather_features16(const DfpGrid& g, const Tri& t, int hf, float (&c)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const float4* src[4];
        float wgt[4];
        bool ok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int x, y, z;
            const int cn = half * 4 + i;
            ok[i] = tri_corner(t, cn, g.W, g.H, g.D, x, y, z);
            wgt[i] = tri_weight(t, cn);
            src[i] = reinterpret_cast<const float4*>(g.data + (ok[i] ? (((int64_t)z * g.H + y) * g.W + x) * CDIM : 0)) + 4 * hf;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = ok[i] ? __ldg(src[i] + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                c[4 * q + 0] = fmaf(v[i].x, wgt[i], c[4 * q + 0]);
                c[4 * q + 1] = fmaf(v[i].y, wgt[i], c[4 * q + 1]);
                c[4 * q + 2] = fmaf(v[i].z, wgt[i], c[4 * q + 2]);
                c[4 * q + 3] = fmaf(v[i].w, wgt[i], c[4 * q + 3]);
            }
        }
    }
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// grid-gradient scatter (dense zero-filled grid, vector reductions) and, optionally, the gradient
// w.r.t. the normalised coordinates                                (ATen grid_sampler_3d backward)
__device__ __forceinline__ void scatter_features(const DfpGrid& g, float* ggrid, const Tri& t, const float* d_c,
                                                 bool want_coord, float gn[3]) {
    float dc[CDIM];
#pragma unroll
    for (int j = 0; j < CDIM; ++j) dc[j] = d_c[j];
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll 1
    for (int cn = 0; cn < 8; ++cn) {
        int x, y, z;
        if (!tri_corner(t, cn, g.W, g.H, g.D, x, y, z)) continue;
        const int64_t off = (((int64_t)z * g.H + y) * g.W + x) * CDIM;
        if (ggrid) {
            const float wgt = tri_weight(t, cn);
#pragma unroll
            for (int q = 0; q < CDIM / 4; ++q)
                red_add_v4(ggrid + off + 4 * q, dc[4 * q] * wgt, dc[4 * q + 1] * wgt, dc[4 * q + 2] * wgt, dc[4 * q + 3] * wgt);
        }
        if (want_coord) {
            const float4* src = reinterpret_cast<const float4*>(g.data + off);
            float dot = 0.f;
#pragma unroll
            for (int q = 0; q < CDIM / 4; ++q) {
                const float4 v = __ldg(src + q);
                dot = fmaf(v.x, dc[4 * q], dot); dot = fmaf(v.y, dc[4 * q + 1], dot);
                dot = fmaf(v.z, dc[4 * q + 2], dot); dot = fmaf(v.w, dc[4 * q + 3], dot);
            }
            float dx, dy, dz;
            tri_weight_grad(t, cn, dx, dy, dz);
            gx = fmaf(dot, dx, gx); gy = fmaf(dot, dy, gy); gz = fmaf(dot, dz, gz);
        }
    }
    gn[0] = gx * t.gx; gn[1] = gy * t.gy; gn[2] = gz * t.gz;
}

// sin / cos of an embedding argument for the tensor-core kernels: two-constant Cody-Waite reduction to [-pi, pi] (exact
// to ~1e-7 for |x| < 4.8e4, the range embed_fast_ok guarantees) and then the special-function unit, whose absolute
// error on that interval is 2^-21 (5e-7).  ~6 instructions instead of ~20 for the polynomial sin_cw.  Used by the delta
// chain and, since round 2, by the forward: its error is below the 3xTF32 products' own (2e-6), the parity margins of
// every fixture are unchanged (profiles/r02_parity_margins.json).  (Round 1 kept the polynomial in the forward because
// the 64-ray smoke case failed with it; that failure was the stale scene-bounds cache, fixed since.)
__device__ __forceinline__ float reduce_2pi(float x) {
    const float k = rintf(x * 0.15915494309189535f);
    return fmaf(-k, -1.7484555e-7f, fmaf(-k, 6.2831854820251465f, x));
}
__device__ __forceinline__ void sincos_sfu(float x, float& s, float& c) {
    const float r = reduce_2pi(x);
    s = __sinf(r);
    c = __cosf(r);
}

// 16-channel half of scatter_features: gn receives this half's partial coordinate gradient
__device__ __forceinline__ void scatter_features16(const DfpGrid& g, float* ggrid, const Tri& t, int hf, const float (&dc)[16],
                                                   bool want_coord, float gn[3]) {
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll 1
    for (int cn = 0; cn < 8; ++cn) {
        int x, y, z;
        if (!tri_corner(t, cn, g.W, g.H, g.D, x, y, z)) continue;
        const int64_t off = (((int64_t)z * g.H + y) * g.W + x) * CDIM + 16 * hf;
        if (ggrid) {
            const float wgt = tri_weight(t, cn);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                red_add_v4(ggrid + off + 4 * q, dc[4 * q] * wgt, dc[4 * q + 1] * wgt, dc[4 * q + 2] * wgt, dc[4 * q + 3] * wgt);
        }
        if (want_coord) {
            const float4* src = reinterpret_cast<const float4*>(g.data + off);
            float dot = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = __ldg(src + q);
                dot = fmaf(v.x, dc[4 * q], dot); dot = fmaf(v.y, dc[4 * q + 1], dot);
                dot = fmaf(v.z, dc[4 * q + 2], dot); dot = fmaf(v.w, dc[4 * q + 3], dot);
            }
            float dx, dy, dz;
            tri_weight_grad(t, cn, dx, dy, dz);
            gx = fmaf(dot, dx, gx); gy = fmaf(dot, dy, gy); gz = fmaf(dot, dz, gz);
        }
    }
    gn[0] = gx * t.gx; gn[1] = gy * t.gy; gn[2] = gz * t.gz;
}

__device__ __forceinline__ void load_weights_to_smem(float* dst, const float* __restrict__ src, int nfloats) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < nfloats / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
}


// ------------------------------------------------------------------------------------------
// decoder forward
// ------------------------------------------------------------------------------------------
struct DecArgs {
    PointSrc ps; int64_t n; Bounds bound;
    DfpGrid grid;            // this decoder's grid
    DfpGrid grid_low;        // high decoder only: the low grid it concatenates (decoder.py:182-187)
    const float* params;     // packed parameters (global)
    float* out;              // DEC 0/1: [n]; DEC 2: [n][4]
    // backward only
    const float* dout;       // DEC 0/1: [n]; DEC 2: [n][4] (4th ignored)
    float* ggrid;            // gradient of `grid` (NDHWC) or null
    float* gparams;          // packed parameter gradient or null
    float* dp;               // [n][3] accumulated, or null
    // tensor-core training path (DFP_FLAG_SAVE_ACT): saved activations / backward scratch, feature-major with stride n
    float* act_h; uint32_t* act_m; float* act_c; float* sc_e; float* sc_darg; float* fcc;   // see Ws
    // optional compacted sample list (device): the kernel then processes samples list[0..*count) only and indexes the
    // saved activations / scratch by list position.  Used for the high decoder, whose output and gradient only matter
    // on TSDF-band samples (decoder.py:326-333: f_add_m_occ is read under the band mask only).
    const int* list; const int* count;
    int* tileflag;           // tensor-core backward: per 128-sample tile, 1 if any upstream gradient is non-zero (K1 writes, K2 reads)
    // per-sample (normalised point | float point) pairs written by k_prep (same dtype chain as load_point + normalize,
    // so the values are bit-identical): the decoder kernels then skip the f64 ray-point / division chain.  May be null.
    const float4* xnpf;
#ifdef DFP_EXPERIMENTS
    int dbg;                 // timing experiments only (tools/exp_build.sh): bits switch parts of a kernel off; never in the product build
#endif
};
#ifdef DFP_EXPERIMENTS
#define DFP_EXP(a, bit) (((a).dbg & (bit)) != 0)
#else
#define DFP_EXP(a, bit) false
#endif
// the two per-sample inputs of a decoder: the point normalised to the scene bound (trilinear coordinates) and the point
// itself in f32 (Fourier embedding)
template <int MODE>
__device__ __forceinline__ void point_inputs(const DecArgs& a, const BoxT<typename PtT<MODE>::T>& box, int64_t s, float xn[3], float pf[3]) {
    if (a.xnpf) {
        const float4 u = __ldg(a.xnpf + 2 * s), v = __ldg(a.xnpf + 2 * s + 1);
        xn[0] = u.x; xn[1] = u.y; xn[2] = u.z;
        pf[0] = v.x; pf[1] = v.y; pf[2] = v.z;
    } else {
        typename PtT<MODE>::T p[3];
        load_point<MODE>(a.ps, s, p);
        normalize(p, box, xn);
        pf[0] = (float)p[0]; pf[1] = (float)p[1]; pf[2] = (float)p[2];
    }
}
// number of work items and the sample id of item idx
__device__ __forceinline__ int64_t dec_items(const DecArgs& a) { return a.list ? (int64_t)*a.count : a.n; }
__device__ __forceinline__ int64_t dec_sample(const DecArgs& a, int64_t idx) { return a.list ? (int64_t)a.list[idx] : idx; }


// ------------------------------------------------------------------------------------------
// attentive fusion over the compacted band list
// ------------------------------------------------------------------------------------------
struct AttArgs {
    const float* params; const int* band_list; const int* band_count;
    const float* low; const float* high; const float* tsdf;
    float* fused; float* w_attn;
    // backward
    PointSrc ps; Bounds tsdf_bnds; DfpTsdf vol;
    const float* d_fused; const float* g_w; float* d_low; float* d_high; float* dp; float* gparams;
    float* act;                    // SIMT path: saved activations of the fusion MLP (ATT_ACT floats per band-list position), or null
    // tensor-core path (dfp_attn_tc.cu), indexed by 128-sample tile of the band list; null = nothing kept
    unsigned char* tc_planes;      // activations h1..h4 as bf16 planes (att::G_TILE bytes per tile)
    uint32_t* tc_masks;            // ReLU masks (att::MASK_WORDS x 128 words per tile)
    float* tc_soft;                // softmax outputs a0, a1 ([2][128] per tile)
    unsigned char* tc_dplanes;     // backward: deltas d2, d3, d4, dl as planes (att::G_DTILE bytes per tile)
    int* tc_tileflag;              // backward: 1 if the tile carries any upstream gradient
};
// h1 (64) | h2 (128) | h3 (128) | h4 (64) | a0 a1 | pad: index ((pos/32)*ATT_ACT + f)*32 + pos%32, so a warp's 32 positions coalesce
constexpr int ATT_ACT = 388;
__device__ __forceinline__ int64_t att_act_index(int pos, int f) { return ((int64_t)(pos >> 5) * ATT_ACT + f) * 32 + (pos & 31); }


// ------------------------------------------------------------------------------------------
// warp-cooperative parameter-gradient sink: every lane holds one sample's (a, d) pair; the warp
// forms sum_samples a[k]*d[j] through a shared-memory transpose and issues one coalesced
// reduction per (k, 32 j) row.
// ------------------------------------------------------------------------------------------
struct WarpSink {
    float* g;        // packed parameter gradient (global)
    float* As;       // [32][32]  As[kk*32 + sample]
    float* Dt;       // [32][33]  Dt[sample*33 + j]
    int lane;

    __device__ __forceinline__ void stage_d(const float (&d)[32]) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) Dt[lane * 33 + j] = d[j];
    }
    template <int KB>
    __device__ __forceinline__ void block(int woff, int ld, const float* a, int k0, int K) {
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) As[kk * 32 + lane] = (k0 + kk < K) ? a[k0 + kk] : 0.f;
        __syncwarp();
        float acc[KB];
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) acc[kk] = 0.f;
#pragma unroll 2
        for (int m = 0; m < 32; m += 4) {
            const float d0 = Dt[(m + 0) * 33 + lane], d1 = Dt[(m + 1) * 33 + lane];
            const float d2 = Dt[(m + 2) * 33 + lane], d3 = Dt[(m + 3) * 33 + lane];
#pragma unroll
            for (int kk = 0; kk < KB; ++kk) {
                const float4 a4 = *reinterpret_cast<const float4*>(As + kk * 32 + m);
                acc[kk] = fmaf(a4.x, d0, fmaf(a4.y, d1, fmaf(a4.z, d2, fmaf(a4.w, d3, acc[kk]))));
            }
        }
#pragma unroll
        for (int kk = 0; kk < KB; ++kk)
            if (k0 + kk < K) atomicAdd(g + woff + (k0 + kk) * ld + lane, acc[kk]);
        __syncwarp();
    }
    __device__ __forceinline__ void bias_from_staged(int boff) {
        float s = 0.f;
#pragma unroll 8
        for (int m = 0; m < 32; ++m) s += Dt[m * 33 + lane];
        atomicAdd(g + boff + lane, s);
    }
    __device__ __forceinline__ void outer32(int woff, int ld, const float* a, int K, const float (&d)[32], int boff) {
        stage_d(d);
        for (int k0 = 0; k0 < K; k0 += 32) block<32>(woff, ld, a, k0, K);
        if (boff >= 0) bias_from_staged(boff);
    }
    __device__ __forceinline__ void outer4(int woff, int ld, const float* a, int K, const float (&d)[32]) {
        stage_d(d);
        block<4>(woff, ld, a, 0, K);
    }
    __device__ __forceinline__ void bias32(int boff, const float (&d)[32]) {
        stage_d(d);
        __syncwarp();
        bias_from_staged(boff);
    }
    __device__ __forceinline__ void scalar(int off, float v) {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(g + off, v);
    }
};


template <typename K>
inline int set_smem(K kernel, size_t bytes) {
    DFP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DFP_OK;
}

inline int grid_for(int64_t n, int per_block, int max_blocks) {
    int64_t b = (n + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

inline Bounds to_bounds(const double b[3][2]) {
    Bounds r;
    memcpy(r.b, b, sizeof(r.b));
    return r;
}


// launchers defined in the per-kernel translation units
template <int DEC, int MODE> int launch_decoder_fwd(const DecArgs& a, cudaStream_t st);
template <int DEC, int MODE> int launch_decoder_bwd(const DecArgs& a, cudaStream_t st);
template <int DEC, int MODE> int launch_decoder_fwd_tc(const DecArgs& a, cudaStream_t st);   // tcgen05 path
// Small batches: the decoders of a call in ONE launch.  d[0] = high, d[1] = colour, d[2] = low; n_dec[i] = 0 leaves decoder i
// out.  Every CTA works out the same split of the grid from the decoders' tile counts (the band list's length is only
// known on the device) and runs one decoder's body.  At the reference's batch sizes (1000 / 200 rays) a decoder kernel is
// one or two tile rounds plus weight staging: three launches leave most SMs idle in every tail.
struct DecArgs3 { DecArgs d[3]; int on[3]; };
template <int MODE> int launch_decoder_fwd_tc_all(const DecArgs3& a, cudaStream_t st);
template <int MODE> int launch_decoder_dh_tc_all(const DecArgs3& a, cudaStream_t st);
int launch_decoder_dw_tc_all(const DecArgs3& a, cudaStream_t st);      // K2 of the decoders that are `on` + the fc_c finish, two launches
// split_ctas (the grid split every CTA of a merged launch derives for itself) lives in dfp_layout.h: it is host-testable
template <int DEC, int MODE> int launch_decoder_dh_tc(const DecArgs& a, cudaStream_t st);      // backward K1: delta chain
template <int DEC> int launch_decoder_dw_tc(const DecArgs& a, cudaStream_t st);                // backward K2: weight gradients
int launch_attn_fwd(const AttArgs& a, int64_t n, cudaStream_t st);
int launch_attn_fwd_tc(const AttArgs& a, int64_t n, cudaStream_t st);                            // tcgen05 path
template <int MODE> int launch_attn_bwd_tc(const AttArgs& a, int64_t n, cudaStream_t st);        // delta chain + weight gradients
template <int MODE> int launch_attn_bwd(const AttArgs& a, int64_t n, cudaStream_t st);

}  // namespace dfp
