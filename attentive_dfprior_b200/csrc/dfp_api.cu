// sm_100a kernels + C ABI of libdfprior_b200.so (see include/dfprior_b200.h).
//
// Pipeline of one Renderer.render_batch_ray call (src/utils/Renderer.py:110-255), all on the
// caller's stream, no host synchronisation, no device allocation:
//
//   k_depth_max      per-chunk max(gt_depth)                       (Renderer.py:159,195)
//   k_sample_z       48 candidate depths per ray, rank-sorted      (Renderer.py:134-221)
//   k_prep           in-bound flag, TSDF gather, band flag, band compaction
//   k_decoder_fwd    x3: trilinear feature gather + Fourier embedding + 5-block MLP, one sample per
//                    thread, decoder weights broadcast from shared memory  (decoder.py:168-203)
//   k_attn_fwd       attentive fusion MLP over the compacted band samples (decoder.py:240-258)
//   k_composite_fwd  stage select, out-of-bound override, alpha compositing (common.py:206-251)
//
// and of its backward: k_composite_bwd -> k_attn_bwd -> k_decoder_bwd x3 (recompute + backprop,
// feature-gradient scatter with vector reductions, warp-cooperative weight gradients) -> k_ray_grad.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "dfp_device.cuh"
#include "dfp_planes.cuh"
#include "dfp_attn_tc.cuh"

namespace dfp {

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
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


// ---- launch counter and opt-in per-kernel timing with CUDA events on the launching stream
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_timing{0};
struct TimedRec { const char* name; cudaEvent_t a, b; };
static std::vector<TimedRec> g_recs;
static std::mutex g_rec_mu;

void timing_begin(const char* name, cudaStream_t st) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_timing.load(std::memory_order_relaxed)) return;
    TimedRec r;
    r.name = name;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    std::lock_guard<std::mutex> lk(g_rec_mu);
    g_recs.push_back(r);
}
void timing_end(cudaStream_t st) {
    if (!g_timing.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_rec_mu);
    if (!g_recs.empty()) cudaEventRecord(g_recs.back().b, st);
}

// Which MLP kernels a call uses is decided by the call's own flags (DfpCfg.flags / DfpPoints.flags), never by process
// state: default = tcgen05 tensor-core kernels; DFP_FLAG_SIMT selects the fp32 SIMT kernels (in-tree cross-check).
// The saved-activation training path (DFP_FLAG_SAVE_ACT) exists only on the tensor-core kernels.
static inline bool flags_tc(int flags) { return (flags & DFP_FLAG_SIMT) == 0; }
static int train_flags(int flags) {
    return flags_tc(flags) ? (flags & (DFP_FLAG_SAVE_ACT | DFP_FLAG_FROZEN_LOW | DFP_FLAG_FROZEN_HIGH | DFP_FLAG_FROZEN_COLOR)) : DFP_FLAG_SIMT;
}
// decoder d (0 low, 1 high, 2 colour) takes part in the weight-gradient path of this call
static inline bool dec_frozen(int flags, int d) { return (flags & (d == 0 ? DFP_FLAG_FROZEN_LOW : d == 1 ? DFP_FLAG_FROZEN_HIGH : DFP_FLAG_FROZEN_COLOR)) != 0; }

static int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            g_num_sms = n;
        else
            g_num_sms = 148;
    }
    return g_num_sms;
}


static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
// calls with at most this many samples run their decoders in one launch per direction (148 SMs x 2 tiles x 10 rounds = 7 893 rays of
// 48 samples: covers the reference's mapping batches, 1000 rays on Replica and 5000 on ScanNet -- measured at 5000 rays: decoder
// kernels 834 -> 712 us; equal at 16 384 rays; the per-decoder scratch planes it needs cost 3 KB per sample more)
constexpr int64_t MERGED_MAX_SAMPLES = 148 * 2 * 128 * 10;

static size_t carve_ws(Ws& w, char* base, int64_t ns, int64_t nchunks, bool with_z, int stage = DFP_STAGE_COLOR, int flags = 0) {
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
    w.z = (double*)take(with_z ? ns * sizeof(double) : 0);
    w.depthmax = (float*)take((nchunks > 0 ? nchunks : 1) * 2 * sizeof(float));
    w.low = (float*)take(ns * 4); w.high = (float*)take(ns * 4); w.tsdf = (float*)take(ns * 4);
    w.fused = (float*)take(ns * 4); w.w_attn = (float*)take(ns * 4);
    w.rgb = (float*)take(ns * 16);
    w.xnpf = (float4*)take(ns * 32);
    w.flags = (int*)take(ns * 4); w.band_list = (int*)take(ns * 4); w.band_count = (int*)take(16);
    w.d_low = (float*)take(ns * 4); w.d_high = (float*)take(ns * 4); w.d_fused = (float*)take(ns * 4);
    w.d_rgb = (float*)take(ns * 16);
    w.dp = (float*)take(ns * 12);
    const bool act = (flags & DFP_FLAG_SAVE_ACT) != 0;
    const int64_t nsp = (ns + 127) / 128 * 128;          // tile-major arrays cover whole 128-sample tiles
    for (int d = 0; d < 3; ++d) {
        const bool used = act && (d == 0 || (d == 1 && stage >= DFP_STAGE_HIGH) || (d == 2 && stage == DFP_STAGE_COLOR));
        const bool wg = used && !dec_frozen(flags, d);      // weight-gradient operands only for decoders that are trained
        w.act_h[d] = (float*)take(wg ? pl::array_bytes(nsp, pl::F_H_ST) : 0);
        w.act_m[d] = (uint32_t*)take(used ? nsp * 5 * 4 : 0);
        w.act_c[d] = (float*)take(wg ? pl::array_bytes(nsp, d == 1 ? 64 : 32) : 0);
        if (!wg) { w.act_h[d] = nullptr; w.act_c[d] = nullptr; }
        if (!used) w.act_m[d] = nullptr;
    }
    const int nscr = (act && ns <= MERGED_MAX_SAMPLES) ? 3 : 1;       // merged launches: the decoders' delta chains run side by side
    for (int d = 0; d < 3; ++d) w.sc_e[d] = d < nscr ? (float*)take(act ? pl::array_bytes(nsp, pl::F_EX) : 0) : w.sc_e[0];
    for (int d = 0; d < 3; ++d) w.fcc[d] = (float*)take(act ? (size_t)(65 * 144) * 4 : 0);      // contiguous: one memset
    for (int d = 0; d < 3; ++d) w.sc_darg[d] = d < nscr ? (float*)take(act ? pl::array_bytes(nsp, pl::F_DLX) + 4096 : 0) : w.sc_darg[0];   // DLX (+ slack for the last group's over-read)
    for (int d = 0; d < 3; ++d) w.dec_tileflag[d] = d < nscr ? (int*)take(act ? (size_t)(nsp / 128) * 4 : 0) : w.dec_tileflag[0];
    // fusion MLP (tensor-core path): activations, ReLU masks, softmax outputs kept by the forward and the delta planes of
    // the backward, per 128-sample tile of the band list (which can be as long as the sample list)
    const bool att = act && stage >= DFP_STAGE_HIGH;
    const size_t ntile = (size_t)(nsp / 128);
    w.att_act = nullptr;
    w.att_planes = (unsigned char*)take(att ? ntile * att::G_TILE : 0);
    w.att_masks = (uint32_t*)take(att ? ntile * att::MASK_WORDS * 128 * 4 : 0);
    w.att_soft = (float*)take(att ? ntile * 256 * 4 : 0);
    w.att_dplanes = (unsigned char*)take(att ? ntile * att::G_DTILE : 0);
    w.att_tileflag = (int*)take(att ? ntile * 4 : 0);
    if (!att) { w.att_planes = nullptr; w.att_masks = nullptr; w.att_soft = nullptr; w.att_dplanes = nullptr; w.att_tileflag = nullptr; }
    return off;
}


// ------------------------------------------------------------------------------------------
// sampling kernels
// ------------------------------------------------------------------------------------------
__global__ void k_depth_max(const float* __restrict__ gt, int64_t n, int64_t chunk, float* __restrict__ out) {
    __shared__ float red[32];
    const int64_t c = blockIdx.x;
    const int64_t lo = c * chunk, hi = min(n, lo + chunk);
    float m = -INFINITY;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) m = fmaxf(m, gt[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) {
            out[2 * c] = __fmul_rn(m, 1.2f);     // max(gt*1.2) == max(gt)*1.2 (monotone rounding)   Renderer.py:159
            out[2 * c + 1] = m;                  // Renderer.py:195
        }
    }
}

__global__ void k_depth_max_given(const float* __restrict__ mx, int64_t nchunks, float* __restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nchunks) { out[2 * c] = __fmul_rn(mx[0], 1.2f); out[2 * c + 1] = mx[0]; }
}

struct SampleArgs {
    const float* rays_o; const float* rays_d; const float* gt; int64_t n_rays; int64_t chunk;
    const float* depthmax; double* z; Bounds bound; SampleTables T;
};

// one thread per (ray, candidate); the ray's candidates are ranked in shared memory
__global__ void __launch_bounds__(256) k_sample_z(const SampleArgs a) {
    __shared__ double cand[256];
    const bool has_depth = a.gt != nullptr;
    const int S = a.T.n_samples + (has_depth ? a.T.n_surface : 0);
    const int rpb = 256 / S;
    const int rl = threadIdx.x / S, i = threadIdx.x - rl * S;
    const int64_t r = (int64_t)blockIdx.x * rpb + rl;
    const bool active = rl < rpb && r < a.n_rays;
    double v = 0.0;
    if (active) {
        const float o[3] = {a.rays_o[3 * r], a.rays_o[3 * r + 1], a.rays_o[3 * r + 2]};
        const float d[3] = {a.rays_d[3 * r], a.rays_d[3 * r + 1], a.rays_d[3 * r + 2]};
        float max12 = 0.f, maxd = 0.f, gt = 0.f;
        if (has_depth) {
            const int64_t c = r / a.chunk;
            max12 = a.depthmax[2 * c]; maxd = a.depthmax[2 * c + 1];
            gt = a.gt[r];
        }
        const double far = ray_far(o, d, a.bound.b, has_depth, max12);
        v = z_candidate(i, has_depth, gt, far, maxd, a.T);
        cand[threadIdx.x] = v;
    }
    __syncthreads();
    if (!active) return;
    int rank = i;
    if (has_depth && a.T.n_surface > 0) {       // torch.sort of the concatenation (values only)   Renderer.py:220-221
        rank = 0;
        const double* row = cand + rl * S;
        for (int j = 0; j < S; ++j) {
            const double u = row[j];
            rank += (u < v) || (u == v && j < i);
        }
    }
    a.z[r * S + rank] = v;
}

// ------------------------------------------------------------------------------------------
// prep: in-bound flag, TSDF value, band flag + compaction
// ------------------------------------------------------------------------------------------
struct PrepArgs {
    PointSrc ps; int64_t n; Bounds bound; Bounds tsdf_bnds; DfpTsdf tsdf; int need_tsdf;
    float* tsdf_out; int* flags; float* w_attn; int* band_list; int* band_count; int apply_mask;
    float4* xnpf;
};

template <int MODE>
__global__ void __launch_bounds__(256) k_prep(const PrepArgs a) {
    using PT = typename PtT<MODE>::T;
    const BoxT<PT> box = make_box<PT>(a.bound.b);
    const BoxT<PT> tbox = make_box<PT>(a.tsdf_bnds.b);
    const int lane = threadIdx.x & 31;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < a.n; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = base + threadIdx.x;
        bool band = false;
        if (s < a.n) {
            PT p[3];
            load_point<MODE>(a.ps, s, p);
            int f = (!a.apply_mask || in_bound(p, box)) ? 1 : 0;
            {   // the decoders' per-sample inputs, once for all of them (DecArgs::xnpf)
                float xb[3];
                normalize(p, box, xb);
                a.xnpf[2 * s] = make_float4(xb[0], xb[1], xb[2], 0.f);
                a.xnpf[2 * s + 1] = make_float4((float)p[0], (float)p[1], (float)p[2], 0.f);
            }
            if (a.need_tsdf) {
                float xn[3];
                normalize(p, tbox, xn);
                const Tri t = tri_setup(xn, a.tsdf.W, a.tsdf.H, a.tsdf.D);
                const float tv = tsdf_interp(a.tsdf.data, a.tsdf.sD, a.tsdf.sH, a.tsdf.sW, a.tsdf.W, a.tsdf.H, a.tsdf.D, t);
                a.tsdf_out[s] = tv;
                band = in_band(tv);
                if (band) f |= 2;
            }
            a.flags[s] = f;
            a.w_attn[s] = 1.f;                                 // decoder.py:317,332,349
        }
        if (a.need_tsdf) {                                     // warp-aggregated append to the band list
            const unsigned m = __ballot_sync(0xffffffffu, band);
            if (m) {
                int pos = 0;
                if (lane == __ffs(m) - 1) pos = atomicAdd(a.band_count, __popc(m));
                pos = __shfl_sync(0xffffffffu, pos, __ffs(m) - 1);
                if (band) a.band_list[pos + __popc(m & ((1u << lane) - 1))] = (int)s;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------
// stage select + compositing
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float final_occ(int stage, int flags, float low, float fused) {
    float occ = (stage != DFP_STAGE_LOW && (flags & 2)) ? fused : low;   // decoder.py:333-336, 350-352
    if (!(flags & 1)) occ = 100.f;                                       // Renderer.py:64
    return occ;
}

struct CompArgs {
    int64_t n_rays; int S; int stage;
    const double* z; const float* low; const float* fused; const float* rgb4; const int* flags; const float* w_attn;
    double* depth; double* unc; float* color; float* weight; float* raw;
    // backward
    const double* g_depth; const double* g_unc; const float* g_color; const float* g_weight;
    float* d_low; float* d_high; float* d_fused; float* d_rgb4; float* dp;
};

// Compositing along the ray with warp scans: 16 lanes per ray (two rays per warp), each lane owns PER = ceil(S/16)
// consecutive samples.  Transmittance T_s = prod_{k<s} (1 - alpha_k + 1e-10) (common.py:244-245) is a lane-local running
// product seeded by a multiplicative exclusive scan over the 16 lanes; the ray sums (depth and variance in f64, colour
// in f32) are butterfly reductions; the backward's suffix sums are the mirrored scan.
constexpr int COMP_LANES = 16;
constexpr int COMP_MAXPER = DFP_MAX_RAY_SAMPLES / COMP_LANES;          // 4
__device__ __forceinline__ double red16(double v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, COMP_LANES);
    return v;
}
__device__ __forceinline__ float red16(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, COMP_LANES);
    return v;
}
// product of the factors of all lanes before this one (1 for lane 0)
__device__ __forceinline__ float excl_prod16(float p, int lane16) {
    float incl = p;
#pragma unroll
    for (int d = 1; d < COMP_LANES; d <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, d, COMP_LANES);
        if (lane16 >= d) incl *= t;
    }
    const float e = __shfl_up_sync(0xffffffffu, incl, 1, COMP_LANES);
    return lane16 == 0 ? 1.f : e;
}
// sum of the values of all lanes after this one (0 for lane 15)
__device__ __forceinline__ double excl_suffix16(double v, int lane16) {
    double incl = v;
#pragma unroll
    for (int d = 1; d < COMP_LANES; d <<= 1) {
        const double t = __shfl_down_sync(0xffffffffu, incl, d, COMP_LANES);
        if (lane16 + d < COMP_LANES) incl += t;
    }
    const double e = __shfl_down_sync(0xffffffffu, incl, 1, COMP_LANES);
    return lane16 == COMP_LANES - 1 ? 0.0 : e;
}

__global__ void __launch_bounds__(256) k_composite_fwd(const CompArgs a) {
    const int lane16 = threadIdx.x & (COMP_LANES - 1);
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / COMP_LANES;
    const bool live = r < a.n_rays;
    const int S = a.S, per = (S + COMP_LANES - 1) / COMP_LANES;
    const bool col = a.stage == DFP_STAGE_COLOR;
    float alpha[COMP_MAXPER], rgb[COMP_MAXPER][3];
    double z[COMP_MAXPER];
    float prod = 1.f;
#pragma unroll
    for (int k = 0; k < COMP_MAXPER; ++k) {
        const int s = lane16 * per + k;
        alpha[k] = 0.f; z[k] = 0.0; rgb[k][0] = rgb[k][1] = rgb[k][2] = 0.f;
        if (k < per && s < S && live) {
            const int64_t i = r * S + s;
            const float occ = final_occ(a.stage, a.flags[i], a.low[i], a.fused[i]);
            if (col) { const float4 v = *reinterpret_cast<const float4*>(a.rgb4 + 4 * i); rgb[k][0] = v.x; rgb[k][1] = v.y; rgb[k][2] = v.z; }
            a.weight[i] = a.w_attn[i];
            if (a.raw) *reinterpret_cast<float4*>(a.raw + 4 * i) = make_float4(rgb[k][0], rgb[k][1], rgb[k][2], occ);
            alpha[k] = sigmoid10(occ);                                   // common.py:236
            z[k] = a.z[i];
            prod *= (1.f - alpha[k]) + 1e-10f;
        }
    }
    float T = excl_prod16(prod, lane16);
    float w[COMP_MAXPER];
    double dsum = 0.0;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < COMP_MAXPER; ++k) {
        const bool on = k < per && lane16 * per + k < S && live;
        w[k] = on ? alpha[k] * T : 0.f;                                  // common.py:244-245
        if (on) T = T * ((1.f - alpha[k]) + 1e-10f);
        dsum += (double)w[k] * z[k];                                     // common.py:248 (f64)
        c0 += w[k] * rgb[k][0]; c1 += w[k] * rgb[k][1]; c2 += w[k] * rgb[k][2];
    }
    const double depth = red16(dsum);
    c0 = red16(c0); c1 = red16(c1); c2 = red16(c2);
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < COMP_MAXPER; ++k) { const double t = z[k] - depth; v += ((double)w[k] * t) * t; }      // common.py:249-250
    v = red16(v);
    if (live && lane16 == 0) {
        a.depth[r] = depth; a.unc[r] = v;
        a.color[3 * r] = c0; a.color[3 * r + 1] = c1; a.color[3 * r + 2] = c2;
    }
}

__global__ void __launch_bounds__(256) k_composite_bwd(const CompArgs a) {
    const int lane16 = threadIdx.x & (COMP_LANES - 1);
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / COMP_LANES;
    const bool live = r < a.n_rays;
    const int S = a.S, per = (S + COMP_LANES - 1) / COMP_LANES;
    const bool col = a.stage == DFP_STAGE_COLOR;
    float alpha[COMP_MAXPER], rgb[COMP_MAXPER][3], Ts[COMP_MAXPER], w[COMP_MAXPER];
    double z[COMP_MAXPER];
    int fl[COMP_MAXPER];
    float prod = 1.f;
#pragma unroll
    for (int k = 0; k < COMP_MAXPER; ++k) {
        const int s = lane16 * per + k;
        alpha[k] = 0.f; z[k] = 0.0; fl[k] = 0; rgb[k][0] = rgb[k][1] = rgb[k][2] = 0.f;
        if (k < per && s < S && live) {
            const int64_t i = r * S + s;
            fl[k] = a.flags[i];
            alpha[k] = sigmoid10(final_occ(a.stage, fl[k], a.low[i], a.fused[i]));
            if (col) { const float4 v = *reinterpret_cast<const float4*>(a.rgb4 + 4 * i); rgb[k][0] = v.x; rgb[k][1] = v.y; rgb[k][2] = v.z; }
            z[k] = a.z[i];
            prod *= (1.f - alpha[k]) + 1e-10f;
        }
    }
    float T = excl_prod16(prod, lane16);
    double dsum = 0.0;
#pragma unroll
    for (int k = 0; k < COMP_MAXPER; ++k) {
        const bool on = k < per && lane16 * per + k < S && live;
        Ts[k] = T;
        w[k] = on ? alpha[k] * T : 0.f;
        if (on) T = T * ((1.f - alpha[k]) + 1e-10f);
        dsum += (double)w[k] * z[k];
    }
    const double depth = red16(dsum);
    double dv = 0.0;                                 // var = sum w (z - D)^2  ->  d var/d D = -2 sum w (z - D)
#pragma unroll
    for (int k = 0; k < COMP_MAXPER; ++k) dv += -2.0 * (double)w[k] * (z[k] - depth);
    dv = red16(dv);
    const double g_depth = (live && a.g_depth) ? a.g_depth[r] : 0.0, g_var = (live && a.g_unc) ? a.g_unc[r] : 0.0;
    float gc[3] = {0.f, 0.f, 0.f};
    if (live && a.g_color) { gc[0] = a.g_color[3 * r]; gc[1] = a.g_color[3 * r + 1]; gc[2] = a.g_color[3 * r + 2]; }
    const double gd = g_depth + g_var * dv;
    double G[COMP_MAXPER], local = 0.0;
#pragma unroll
    for (int k = 0; k < COMP_MAXPER; ++k) {
        const double dz = z[k] - depth;
        G[k] = gd * z[k] + g_var * dz * dz;
        if (col) G[k] += (double)gc[0] * rgb[k][0] + (double)gc[1] * rgb[k][1] + (double)gc[2] * rgb[k][2];
        local += G[k] * (double)w[k];
    }
    double suffix = excl_suffix16(local, lane16);    // sum_{k > s} G_k w_k over the lanes after this one
#pragma unroll
    for (int k = COMP_MAXPER - 1; k >= 0; --k) {
        const int s = lane16 * per + k;
        if (!(k < per && s < S && live)) continue;
        const int64_t i = r * S + s;
        const double one_minus = (double)((1.f - alpha[k]) + 1e-10f);
        const double dalpha = G[k] * (double)Ts[k] - suffix / one_minus;
        suffix += G[k] * (double)w[k];
        const float d_occ = (float)(dalpha * 10.0 * (double)alpha[k] * (1.0 - (double)alpha[k]));
        const float g = (fl[k] & 1) ? d_occ : 0.f;                       // occ overwritten by 100 outside the bound
        const bool band = a.stage != DFP_STAGE_LOW && (fl[k] & 2);
        a.d_low[i] = band ? 0.f : g;
        a.d_high[i] = 0.f;
        a.d_fused[i] = band ? g : 0.f;
        if (col) *reinterpret_cast<float4*>(a.d_rgb4 + 4 * i) = make_float4(w[k] * gc[0], w[k] * gc[1], w[k] * gc[2], 0.f);
        if (a.dp) { a.dp[3 * i] = 0.f; a.dp[3 * i + 1] = 0.f; a.dp[3 * i + 2] = 0.f; }
    }
}

// rays_o / rays_d gradients: pts = o + d*z  (Renderer.py:223-224)
__global__ void k_ray_grad(int64_t n_rays, int S, const float* __restrict__ dp, const double* __restrict__ z,
                           float* __restrict__ d_o, float* __restrict__ d_d) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    double so[3] = {0, 0, 0}, sd[3] = {0, 0, 0};
    for (int s = 0; s < S; ++s) {
        const int64_t i = r * S + s;
        for (int k = 0; k < 3; ++k) { const double g = dp[3 * i + k]; so[k] += g; sd[k] += g * z[i]; }
    }
    for (int k = 0; k < 3; ++k) { if (d_o) d_o[3 * r + k] = (float)so[k]; if (d_d) d_d[3 * r + k] = (float)sd[k]; }
}

// ---- point-query epilogues (DF.forward / eval_points) ----
struct PtsArgs {
    int64_t n; int stage;
    const float* low; const float* fused; const float* rgb4; const int* flags; const float* w_attn;
    float* raw; float* w;
    const float* g_raw; const float* g_w; float* d_low; float* d_high; float* d_fused; float* d_rgb4; float* dp;
};
__global__ void k_points_finalize(const PtsArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const bool col = a.stage == DFP_STAGE_COLOR;
    const float occ = final_occ(a.stage, a.flags[i], a.low[i], a.fused[i]);
    float4 v = make_float4(0.f, 0.f, 0.f, occ);
    if (col) { const float4 c = *reinterpret_cast<const float4*>(a.rgb4 + 4 * i); v.x = c.x; v.y = c.y; v.z = c.z; }
    *reinterpret_cast<float4*>(a.raw + 4 * i) = v;
    if (a.w) a.w[i] = a.w_attn[i];
}
__global__ void k_points_seed_bwd(const PtsArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int f = a.flags[i];
    const float4 gr = a.g_raw ? *reinterpret_cast<const float4*>(a.g_raw + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float g = (f & 1) ? gr.w : 0.f;
    const bool band = a.stage != DFP_STAGE_LOW && (f & 2);
    a.d_low[i] = band ? 0.f : g;
    a.d_high[i] = 0.f;
    a.d_fused[i] = band ? g : 0.f;
    if (a.stage == DFP_STAGE_COLOR) *reinterpret_cast<float4*>(a.d_rgb4 + 4 * i) = make_float4(gr.x, gr.y, gr.z, 0.f);
    if (a.dp) { a.dp[3 * i] = 0.f; a.dp[3 * i + 1] = 0.f; a.dp[3 * i + 2] = 0.f; }
}
template <typename PT>
__global__ void k_points_grad_out(int64_t n, const float* __restrict__ dp, PT* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * n) out[i] = (PT)dp[i];
}

template <int MODE>
__global__ void k_tsdf_sample(PointSrc ps, int64_t n, Bounds tb, DfpTsdf vol, float* out) {
    using PT = typename PtT<MODE>::T;
    const BoxT<PT> tbox = make_box<PT>(tb.b);
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    PT p[3];
    load_point<MODE>(ps, s, p);
    float xn[3];
    normalize(p, tbox, xn);
    out[s] = tsdf_interp(vol.data, vol.sD, vol.sH, vol.sW, vol.W, vol.H, vol.D, tri_setup(xn, vol.W, vol.H, vol.D));
}

// ------------------------------------------------------------------------------------------
// parameter pack / unpack and grid layout transposes
// ------------------------------------------------------------------------------------------
struct PtrTable { const float* p[N_TENSORS]; };

constexpr int ATT_IMG_BLOCKS = 16;      // 4096 (row, 8-input chunk) tasks of the three wide fusion layers, 256 per block

constexpr int PACK_SPLIT = 8;           // blocks per tensor (the largest tensor has 16 384 elements: one block per tensor is latency-bound)

constexpr int DEC_M_BLOCKS = 3 * 17;     // per decoder: 4096 entries of M_1..M_4 (16 blocks of 256) + one block for V

__global__ void k_pack_params(const PtrTable tab, float* __restrict__ packed) {
    if (blockIdx.x >= N_TENSORS + ATT_IMG_BLOCKS) {
        // delta-chain products of the decoders (dfp_layout.h, DecM), straight from the native [out][in] tensors
        if (blockIdx.y != 0) return;
        const int b = blockIdx.x - N_TENSORS - ATT_IMG_BLOCKS, dec = b / 17, blk = b % 17;
        const int cd = dec_cd(dec), nout = dec_nout(dec);
        float* dst = packed + OFF_DEC_M + dec * DecM::SIZE;
        const int t0 = 23 * dec;                                         // fc_c.i.weight = t0 + 2 i, pts_linears.i.weight = t0 + 11 + 2 i
        if (blk < 16) {
            const int e = blk * 256 + threadIdx.x, slot = e >> 10, j = (e >> 5) & 31, q = e & 31, i = slot + 1;
            const float* W = tab.p[t0 + 11 + 2 * i];                     // [32][32], or [32][93 + 32] for i = 3 (hidden part last)
            const int ld = (i == 3) ? EMB + HID : HID, k0 = (i == 3) ? EMB : 0;
            const float* Wc = tab.p[t0 + 2 * (i - 1)];                   // fc_c.(i-1).weight [32][cd]
            float acc = 0.f;
            for (int k = 0; k < HID; ++k) acc = fmaf(W[j * ld + k0 + k], Wc[k * cd + q], acc);
            dst[DecM::M + e] = acc;
        } else if (threadIdx.x < 4 * HID) {
            const int o = threadIdx.x >> 5, q = threadIdx.x & 31;
            float acc = 0.f;
            if (o < nout) {
                const float* Wo = tab.p[t0 + 21];                        // output_linear.weight [nout][32]
                const float* Wc = tab.p[t0 + 8];                         // fc_c.4.weight [32][cd]
                for (int k = 0; k < HID; ++k) acc = fmaf(Wo[o * HID + k], Wc[k * cd + q], acc);
            }
            dst[DecM::V + threadIdx.x] = acc;
        }
        return;
    }
    if (blockIdx.x >= N_TENSORS) {
        if (blockIdx.y != 0) return;
        // tensor-core image of the fusion MLP (dfp_layout.h, AttImg): 8 consecutive inputs of one output row are 32
        // contiguous source bytes (torch keeps [out][in]) and one 16-byte core-matrix row of each plane
        int task = (blockIdx.x - N_TENSORS) * blockDim.x + threadIdx.x;
        unsigned char* img = reinterpret_cast<unsigned char*>(packed + OFF_ATT_IMG);
        unsigned char* img16 = reinterpret_cast<unsigned char*>(packed + OFF_ATT_IMG_F16);
        const float* src; int nin, nout, hi_off, mid_off;
        if (task < 1024) { src = tab.p[71]; nin = 64; nout = 128; hi_off = AttImg::W2H; mid_off = AttImg::W2M; }
        else if (task < 3072) { task -= 1024; src = tab.p[73]; nin = 128; nout = 128; hi_off = AttImg::W3H; mid_off = AttImg::W3M; }
        else { task -= 3072; src = tab.p[75]; nin = 128; nout = 64; hi_off = AttImg::W4H; mid_off = AttImg::W4M; }
        const int nchunk = nin / 8, j = task / nchunk, c = task % nchunk;
        if (j < nout) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = src[j * nin + c * 8 + i];
            uint4 hi, mid;
            pl::split8<false>(v, hi, mid);
            *reinterpret_cast<uint4*>(img + hi_off + c * (nout * 16) + j * 16) = hi;
            *reinterpret_cast<uint4*>(img + mid_off + c * (nout * 16) + j * 16) = mid;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= ATT_IMG_F16_SCALE;
            pl::split8<true>(v, hi, mid);
            *reinterpret_cast<uint4*>(img16 + hi_off + c * (nout * 16) + j * 16) = hi;
            *reinterpret_cast<uint4*>(img16 + mid_off + c * (nout * 16) + j * 16) = mid;
        }
        return;
    }
    const int t = blockIdx.x;
    int rows, cols;
    tensor_shape(t, rows, cols);
    const float* src = tab.p[t];
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < rows * cols; i += PACK_SPLIT * blockDim.x) packed[packed_index(t, i / cols, i % cols)] = src[i];
}
struct OffTable { int off[N_TENSORS]; };
__global__ void k_unpack_grads(const float* __restrict__ packed, float* __restrict__ flat, const OffTable offs) {
    const int t = blockIdx.x;
    int rows, cols;
    tensor_shape(t, rows, cols);
    float* dst = flat + offs.off[t];
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < rows * cols; i += PACK_SPLIT * blockDim.x) dst[i] = packed[packed_index(t, i / cols, i % cols)];
}

// [C][V] <-> [V][C] with V = D*H*W, C = 32: a 32x32 shared-memory tile per block
__global__ void k_transpose_c_v(const float* __restrict__ src, float* __restrict__ dst, int C, int64_t V, int to_last) {
    __shared__ float tile[32][33];
    const int64_t v0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    for (int c0 = 0; c0 < C; c0 += 32) {
        if (to_last) {            // src [C][V] -> dst [V][C]
            for (int j = ty; j < 32; j += 8) { const int64_t v = v0 + tx; tile[j][tx] = (v < V && c0 + j < C) ? src[(int64_t)(c0 + j) * V + v] : 0.f; }
            __syncthreads();
            for (int j = ty; j < 32; j += 8) { const int64_t v = v0 + j; if (v < V && c0 + tx < C) dst[v * C + c0 + tx] = tile[tx][j]; }
        } else {                  // src [V][C] -> dst [C][V]
            for (int j = ty; j < 32; j += 8) { const int64_t v = v0 + j; tile[j][tx] = (v < V && c0 + tx < C) ? src[v * C + c0 + tx] : 0.f; }
            __syncthreads();
            for (int j = ty; j < 32; j += 8) { const int64_t v = v0 + tx; if (v < V && c0 + j < C) dst[(int64_t)(c0 + j) * V + v] = tile[tx][j]; }
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------
// host-side launch helpers
// ------------------------------------------------------------------------------------------
static int check_scene(const DfpScene* sc, int stage) {
    if (!sc) return fail(DFP_ERR_INVALID, "scene is null");
    if (stage < DFP_STAGE_LOW || stage > DFP_STAGE_COLOR) return fail(DFP_ERR_INVALID, "bad stage %d", stage);
    const DfpGrid* gs[3] = {&sc->low, &sc->high, &sc->color};
    const int need[3] = {1, stage >= DFP_STAGE_HIGH, stage == DFP_STAGE_COLOR};
    for (int i = 0; i < 3; ++i) {
        if (!need[i]) continue;
        if (!gs[i]->data) return fail(DFP_ERR_INVALID, "grid %d is null", i);
        if (gs[i]->C != DFP_C_DIM) return fail(DFP_ERR_INVALID, "grid %d has C=%d, only %d supported", i, gs[i]->C, DFP_C_DIM);
        if (gs[i]->D < 1 || gs[i]->H < 1 || gs[i]->W < 1) return fail(DFP_ERR_INVALID, "grid %d has an empty dimension", i);
        if (((uintptr_t)gs[i]->data) & 15) return fail(DFP_ERR_INVALID, "grid %d is not 16-byte aligned", i);
    }
    if (stage >= DFP_STAGE_HIGH) {
        if (!sc->tsdf.data) return fail(DFP_ERR_INVALID, "tsdf volume is null");
        if (sc->tsdf.D < 1 || sc->tsdf.H < 1 || sc->tsdf.W < 1) return fail(DFP_ERR_INVALID, "tsdf volume has an empty dimension");
    }
    for (int k = 0; k < 3; ++k)
        if (!(sc->bound[k][1] > sc->bound[k][0])) return fail(DFP_ERR_INVALID, "bound axis %d is empty", k);
    return DFP_OK;
}

// forward of everything between "points are known" and "low/high/rgb/fused/w/flags are in the workspace"
template <int MODE>
static int eval_samples_fwd(const DfpScene* sc, const float* params, const PointSrc& ps, int64_t n, int stage,
                            int apply_mask, const Ws& w, int flags, cudaStream_t st) {
    const bool hi = stage >= DFP_STAGE_HIGH;
    if (hi) DFP_CUDA(cudaMemsetAsync(w.band_count, 0, 16, st));
    PrepArgs pa;
    pa.ps = ps; pa.n = n; pa.bound = to_bounds(sc->bound); pa.tsdf_bnds = to_bounds(sc->tsdf_bnds); pa.tsdf = sc->tsdf;
    pa.need_tsdf = hi; pa.tsdf_out = w.tsdf; pa.flags = w.flags; pa.w_attn = w.w_attn;
    pa.band_list = w.band_list; pa.band_count = w.band_count; pa.apply_mask = apply_mask; pa.xnpf = w.xnpf;
    DFP_TIMED("k_prep", st, k_prep<MODE><<<grid_for(n, 256, 8 * num_sms()), 256, 0, st>>>(pa));
    DFP_LAUNCH_CHECK("k_prep");

    DecArgs da;
    memset(&da, 0, sizeof(da));
    da.ps = ps; da.n = n; da.bound = to_bounds(sc->bound); da.params = params; da.xnpf = w.xnpf;
#ifdef DFP_EXPERIMENTS
    { const char* e = getenv("DFP_EXP"); da.dbg = e ? atoi(e) : 0; }
#endif
    int rc;
    const bool tcp = flags_tc(flags);
    const bool save = tcp && (flags & DFP_FLAG_SAVE_ACT);
    auto set_act = [&](int d) {
        da.act_h = save ? w.act_h[d] : nullptr; da.act_m = save ? w.act_m[d] : nullptr; da.act_c = save ? w.act_c[d] : nullptr;
    };
    // small batches: one launch for all decoders of the stage (dfp_device.cuh, DecArgs3)
    if (tcp && n <= MERGED_MAX_SAMPLES) {
        DecArgs3 a3;
        memset(&a3, 0, sizeof(a3));
        if (hi) {
            set_act(1);
            da.grid = sc->high; da.grid_low = sc->low; da.out = w.high; da.list = w.band_list; da.count = w.band_count;
            a3.d[0] = da; a3.on[0] = 1;
            da.list = nullptr; da.count = nullptr;
        }
        if (stage == DFP_STAGE_COLOR) { set_act(2); da.grid = sc->color; da.out = w.rgb; a3.d[1] = da; a3.on[1] = 1; }
        set_act(0);
        da.grid = sc->low; da.out = w.low;
        a3.d[2] = da; a3.on[2] = 1;
        if ((rc = launch_decoder_fwd_tc_all<MODE>(a3, st))) return rc;
    } else {
    if (hi) {
        set_act(1);
        da.grid = sc->high; da.grid_low = sc->low; da.out = w.high;
        if (tcp) { da.list = w.band_list; da.count = w.band_count; }      // high occupancy is only read on band samples
        if ((rc = tcp ? launch_decoder_fwd_tc<1, MODE>(da, st) : launch_decoder_fwd<1, MODE>(da, st))) return rc;
        da.list = nullptr; da.count = nullptr;
    }
    if (stage == DFP_STAGE_COLOR) {
        set_act(2);
        da.grid = sc->color; da.out = w.rgb;
        if ((rc = tcp ? launch_decoder_fwd_tc<2, MODE>(da, st) : launch_decoder_fwd<2, MODE>(da, st))) return rc;
    }
    set_act(0);
    da.grid = sc->low; da.out = w.low;
    if ((rc = tcp ? launch_decoder_fwd_tc<0, MODE>(da, st) : launch_decoder_fwd<0, MODE>(da, st))) return rc;
    }
    if (hi) {
        AttArgs aa;
        memset(&aa, 0, sizeof(aa));
        aa.params = params; aa.band_list = w.band_list; aa.band_count = w.band_count;
        aa.low = w.low; aa.high = w.high; aa.tsdf = w.tsdf; aa.fused = w.fused; aa.w_attn = w.w_attn;
        if (tcp) {
            aa.tc_planes = save ? w.att_planes : nullptr; aa.tc_masks = save ? w.att_masks : nullptr; aa.tc_soft = save ? w.att_soft : nullptr;
            if ((rc = launch_attn_fwd_tc(aa, n, st))) return rc;
        } else if ((rc = launch_attn_fwd(aa, n, st))) return rc;
    }
    return DFP_OK;
}

// backward counterpart: d_low / d_high / d_fused / d_rgb / dp seeded in the workspace
template <int MODE>
static int eval_samples_bwd(const DfpScene* sc, const float* params, const PointSrc& ps, int64_t n, int stage,
                            const float* g_w, const DfpGrads* g, bool want_dp, const Ws& w, int flags, cudaStream_t st) {
    const bool hi = stage >= DFP_STAGE_HIGH;
    int rc;
    if (hi) {
        AttArgs aa;
        memset(&aa, 0, sizeof(aa));
        aa.params = params; aa.band_list = w.band_list; aa.band_count = w.band_count;
        aa.low = w.low; aa.high = w.high; aa.tsdf = w.tsdf;
        aa.ps = ps; aa.tsdf_bnds = to_bounds(sc->tsdf_bnds); aa.vol = sc->tsdf;
        aa.d_fused = w.d_fused; aa.g_w = g_w; aa.d_low = w.d_low; aa.d_high = w.d_high;
        aa.dp = want_dp ? w.dp : nullptr; aa.gparams = g->params_packed;
        if (flags_tc(flags) && (flags & DFP_FLAG_SAVE_ACT) && w.att_planes != nullptr) {
            aa.tc_planes = w.att_planes; aa.tc_masks = w.att_masks; aa.tc_soft = w.att_soft; aa.tc_dplanes = w.att_dplanes;
            aa.tc_tileflag = w.att_tileflag;
            if ((rc = launch_attn_bwd_tc<MODE>(aa, n, st))) return rc;
        } else if ((rc = launch_attn_bwd<MODE>(aa, n, st))) return rc;
    }
    DecArgs da;
    memset(&da, 0, sizeof(da));
    da.ps = ps; da.n = n; da.bound = to_bounds(sc->bound); da.params = params; da.xnpf = w.xnpf;
#ifdef DFP_EXPERIMENTS
    { const char* e = getenv("DFP_EXP"); da.dbg = e ? atoi(e) : 0; }
#endif
    da.gparams = g->params_packed; da.dp = want_dp ? w.dp : nullptr;
    // tensor-core backward (needs the activations saved by the tensor-core forward): K1 = delta chain, K2 = weight grads
    const bool tcb = flags_tc(flags) && (flags & DFP_FLAG_SAVE_ACT) && w.act_m[0] != nullptr;
    if (tcb && da.gparams) DFP_CUDA(cudaMemsetAsync(w.fcc[0], 0, (size_t)((char*)w.fcc[2] - (char*)w.fcc[0]) + 65 * 144 * 4, st));
    auto set_dec = [&](int dec) {
        da.act_h = w.act_h[dec]; da.act_m = w.act_m[dec]; da.act_c = w.act_c[dec]; da.fcc = w.fcc[dec];
        da.sc_e = w.sc_e[dec]; da.sc_darg = w.sc_darg[dec]; da.tileflag = w.dec_tileflag[dec];
        da.list = (dec == 1) ? w.band_list : nullptr; da.count = (dec == 1) ? w.band_count : nullptr;   // as in the forward
        if (dec == 1) { da.grid = sc->high; da.grid_low = sc->low; da.dout = w.d_high; da.ggrid = g->grid_high; }
        else if (dec == 2) { da.grid = sc->color; da.dout = w.d_rgb; da.ggrid = g->grid_color; }
        else { da.grid = sc->low; da.dout = w.d_low; da.ggrid = g->grid_low; }
    };
    // small batches (no all-reduce events to record in between): the decoders' delta chains in one launch, then their
    // weight-gradient kernels in one launch
    if (tcb && n <= MERGED_MAX_SAMPLES && !g->ev_high && !g->ev_color) {
        DecArgs3 a3;
        memset(&a3, 0, sizeof(a3));
        float* const gp = da.gparams;
        const int decs[3] = {1, 2, 0};
        bool any_w = false;
        for (int i = 0; i < 3; ++i) {
            const int dec = decs[i];
            if ((dec == 1 && !hi) || (dec == 2 && stage != DFP_STAGE_COLOR)) continue;
            set_dec(dec);
            da.gparams = dec_frozen(flags, dec) ? nullptr : gp;
            a3.d[i] = da; a3.on[i] = 1;
            any_w = any_w || da.gparams != nullptr;
        }
        da.gparams = gp;
        if ((rc = launch_decoder_dh_tc_all<MODE>(a3, st))) return rc;
        if (any_w) {
            for (int i = 0; i < 3; ++i) a3.on[i] = a3.on[i] && a3.d[i].gparams != nullptr;
            if ((rc = launch_decoder_dw_tc_all(a3, st))) return rc;
        }
        return DFP_OK;
    }
    auto run = [&](auto dec_tag) -> int {
        constexpr int DEC = decltype(dec_tag)::value;
        if (!tcb) {
            if (DEC == 1) { da.grid = sc->high; da.grid_low = sc->low; da.dout = w.d_high; da.ggrid = g->grid_high; }
            else if (DEC == 2) { da.grid = sc->color; da.dout = w.d_rgb; da.ggrid = g->grid_color; }
            else { da.grid = sc->low; da.dout = w.d_low; da.ggrid = g->grid_low; }
            return launch_decoder_bwd<DEC, MODE>(da, st);
        }
        set_dec(DEC);
        float* const gp = da.gparams;
        if (dec_frozen(flags, DEC)) da.gparams = nullptr;        // K1 then writes no operands for K2, and K2 does not run
        int r = launch_decoder_dh_tc<DEC, MODE>(da, st);
        if (!r && da.gparams) r = launch_decoder_dw_tc<DEC>(da, st);
        da.gparams = gp;
        return r;
    };
    if (hi && (rc = run(std::integral_constant<int, 1>{}))) return rc;
    if (g->ev_high) DFP_CUDA(cudaEventRecord((cudaEvent_t)g->ev_high, st));        // grid_high's gradient is final
    if (stage == DFP_STAGE_COLOR && (rc = run(std::integral_constant<int, 2>{}))) return rc;
    if (g->ev_color) DFP_CUDA(cudaEventRecord((cudaEvent_t)g->ev_color, st));      // grid_color's gradient is final
    if ((rc = run(std::integral_constant<int, 0>{}))) return rc;
    return DFP_OK;
}

static int fill_tables(const DfpCfg* cfg, SampleTables& T) {
    if (cfg->n_importance != 0) return fail(DFP_ERR_UNSUPPORTED, "N_importance must be 0 (the reference branch is dead, Renderer.py:235-252)");
    if (cfg->n_samples < 1 || cfg->n_surface < 0 || cfg->n_samples + cfg->n_surface > DFP_MAX_RAY_SAMPLES)
        return fail(DFP_ERR_INVALID, "N_samples + N_surface must be in [1, %d]", DFP_MAX_RAY_SAMPLES);
    T.n_samples = cfg->n_samples; T.n_surface = cfg->n_surface;
    memcpy(T.tv, cfg->t_vals, sizeof(T.tv)); memcpy(T.omt, cfg->one_minus_t, sizeof(T.omt));
    memcpy(T.ts, cfg->t_surf, sizeof(T.ts)); memcpy(T.omts, cfg->one_minus_t_surf, sizeof(T.omts));
    return DFP_OK;
}


}  // namespace dfp

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace dfp;

extern "C" {

int dfp_abi_version(void) { return DFP_ABI_VERSION; }
const char* dfp_last_error(void) { return g_err; }
int64_t dfp_packed_param_floats(void) { return PACKED_TOTAL_FLOATS; }
int64_t dfp_packed_grad_floats(void) { return PACKED_FLOATS; }

int64_t dfp_launch_count(void) { return (int64_t)g_launches.load(); }

int dfp_timing_enable(int on) {
    g_timing.store(on ? 1 : 0);
    return DFP_OK;
}

// Synchronises the recorded events and writes one line per kernel name: "<name> <launches> <total_ms>\n".
int dfp_timing_collect(char* buf, size_t buf_bytes) {
    if (!buf || buf_bytes == 0) return fail(DFP_ERR_INVALID, "null buffer");
    std::vector<TimedRec> recs;
    {
        std::lock_guard<std::mutex> lk(g_rec_mu);
        recs.swap(g_recs);
    }
    std::map<std::string, std::pair<long long, double>> agg;
    for (auto& r : recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto& e = agg[r.name];
            e.first += 1;
            e.second += ms;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    std::string out;
    for (auto& kv : agg) {
        char line[256];
        snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (out.size() + 1 > buf_bytes) return fail(DFP_ERR_INVALID, "timing buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return DFP_OK;
}

int dfp_pack_params(const float* const* tensors79, float* packed, void* stream) {
    if (!tensors79 || !packed) return fail(DFP_ERR_INVALID, "null argument");
    PtrTable tab;
    for (int t = 0; t < N_TENSORS; ++t) {
        if (!tensors79[t]) return fail(DFP_ERR_INVALID, "parameter tensor %d is null", t);
        tab.p[t] = tensors79[t];
    }
    DFP_TIMED("k_pack_params", (cudaStream_t)stream, k_pack_params<<<dim3(N_TENSORS + ATT_IMG_BLOCKS + DEC_M_BLOCKS, PACK_SPLIT), 256, 0, (cudaStream_t)stream>>>(tab, packed));
    DFP_LAUNCH_CHECK("k_pack_params");
    return DFP_OK;
}

int dfp_unpack_param_grads(const float* packed_grad, float* native_flat, void* stream) {
    if (!packed_grad || !native_flat) return fail(DFP_ERR_INVALID, "null argument");
    OffTable offs;
    int off = 0;
    for (int t = 0; t < N_TENSORS; ++t) {
        int r, c;
        tensor_shape(t, r, c);
        offs.off[t] = off;
        off += r * c;
    }
    DFP_TIMED("k_unpack_grads", (cudaStream_t)stream, k_unpack_grads<<<dim3(N_TENSORS, PACK_SPLIT), 256, 0, (cudaStream_t)stream>>>(packed_grad, native_flat, offs));
    DFP_LAUNCH_CHECK("k_unpack_grads");
    return DFP_OK;
}

static int transpose_grid(const float* src, float* dst, int32_t C, int32_t D, int32_t H, int32_t W, int to_last, void* stream) {
    if (!src || !dst) return fail(DFP_ERR_INVALID, "null argument");
    if (C < 1 || D < 1 || H < 1 || W < 1) return fail(DFP_ERR_INVALID, "empty grid");
    const int64_t V = (int64_t)D * H * W;
    DFP_TIMED("k_transpose_c_v", (cudaStream_t)stream, k_transpose_c_v<<<(unsigned)((V + 31) / 32), 256, 0, (cudaStream_t)stream>>>(src, dst, C, V, to_last));
    DFP_LAUNCH_CHECK("k_transpose_c_v");
    return DFP_OK;
}
int dfp_grid_to_ndhwc(const float* s, float* d, int32_t C, int32_t D, int32_t H, int32_t W, void* st) { return transpose_grid(s, d, C, D, H, W, 1, st); }
int dfp_grid_to_cdhw(const float* s, float* d, int32_t C, int32_t D, int32_t H, int32_t W, void* st) { return transpose_grid(s, d, C, D, H, W, 0, st); }

size_t dfp_render_workspace_bytes(int64_t n_rays, int32_t S, int64_t chunk, int32_t stage, int32_t flags) {
    if (n_rays < 0 || S < 1) return 0;
    if (chunk <= 0) chunk = n_rays > 0 ? n_rays : 1;
    Ws w;
    return carve_ws(w, nullptr, n_rays * S, (n_rays + chunk - 1) / chunk, true, stage, train_flags(flags));
}
size_t dfp_points_workspace_bytes(int64_t n, int32_t stage, int32_t flags) {
    if (n < 0) return 0;
    Ws w;
    return carve_ws(w, nullptr, n, 1, false, stage, train_flags(flags));
}

int dfp_render_fwd(const DfpScene* sc, const float* params, const DfpRays* rays, const DfpCfg* cfg,
                   const DfpRenderOut* out, void* workspace, size_t ws_bytes, void* stream) {
    if (!rays || !cfg || !out || !params || !workspace) return fail(DFP_ERR_INVALID, "null argument");
    int rc = check_scene(sc, cfg->stage);
    if (rc) return rc;
    if (rays->n_rays < 1) return fail(DFP_ERR_INVALID, "n_rays must be >= 1 (torch.max of an empty depth batch raises upstream too)");
    if (!rays->rays_o || !rays->rays_d) return fail(DFP_ERR_INVALID, "rays are null");
    if (!out->depth || !out->uncertainty || !out->color || !out->weight) return fail(DFP_ERR_INVALID, "output pointer is null");
    SampleTables T;
    if ((rc = fill_tables(cfg, T))) return rc;
    const bool has_depth = rays->gt_depth != nullptr;
    const int S = T.n_samples + (has_depth ? T.n_surface : 0);
    const int64_t N = rays->n_rays, NS = N * S;
    if (NS > 0x7fffffffLL) return fail(DFP_ERR_INVALID, "too many samples in one call (%lld)", (long long)NS);
    const int64_t chunk = rays->chunk > 0 ? rays->chunk : N;
    const int64_t nchunks = (N + chunk - 1) / chunk;
    Ws w;
    const int flags = train_flags(cfg->flags);
    const size_t need = carve_ws(w, (char*)workspace, NS, nchunks, true, cfg->stage, flags);
    if (need > ws_bytes) return fail(DFP_ERR_WORKSPACE, "workspace too small: need %zu bytes, have %zu", need, ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;

    if (has_depth && rays->depth_max) {
        DFP_TIMED("k_depth_max_given", st, k_depth_max_given<<<(unsigned)((nchunks + 127) / 128), 128, 0, st>>>(rays->depth_max, nchunks, w.depthmax));
        DFP_LAUNCH_CHECK("k_depth_max_given");
    } else if (has_depth) {
        DFP_TIMED("k_depth_max", st, k_depth_max<<<(unsigned)nchunks, 1024, 0, st>>>(rays->gt_depth, N, chunk, w.depthmax));
        DFP_LAUNCH_CHECK("k_depth_max");
    }
    SampleArgs sa;
    sa.rays_o = rays->rays_o; sa.rays_d = rays->rays_d; sa.gt = rays->gt_depth; sa.n_rays = N; sa.chunk = chunk;
    sa.depthmax = w.depthmax; sa.z = w.z; sa.bound = to_bounds(sc->bound); sa.T = T;
    const int rpb = 256 / S;
    DFP_TIMED("k_sample_z", st, k_sample_z<<<(unsigned)((N + rpb - 1) / rpb), 256, 0, st>>>(sa));
    DFP_LAUNCH_CHECK("k_sample_z");

    PointSrc ps;
    ps.rays_o = rays->rays_o; ps.rays_d = rays->rays_d; ps.z = w.z; ps.S = S; ps.pts = nullptr; ps.pts_f64 = 0;
    if ((rc = eval_samples_fwd<0>(sc, params, ps, NS, cfg->stage, 1, w, flags, st))) return rc;

    CompArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.n_rays = N; ca.S = S; ca.stage = cfg->stage; ca.z = w.z; ca.low = w.low; ca.fused = w.fused; ca.rgb4 = w.rgb;
    ca.flags = w.flags; ca.w_attn = w.w_attn;
    ca.depth = out->depth; ca.unc = out->uncertainty; ca.color = out->color; ca.weight = out->weight; ca.raw = out->raw;
    DFP_TIMED("k_composite_fwd", st, k_composite_fwd<<<(unsigned)((N * COMP_LANES + 255) / 256), 256, 0, st>>>(ca));
    DFP_LAUNCH_CHECK("k_composite_fwd");
    if (out->z_vals) DFP_CUDA(cudaMemcpyAsync(out->z_vals, w.z, NS * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return DFP_OK;
}

int dfp_render_bwd(const DfpScene* sc, const float* params, const DfpRays* rays, const DfpCfg* cfg,
                   const DfpRenderOutGrad* og, const DfpGrads* g, void* workspace, size_t ws_bytes, void* stream) {
    if (!rays || !cfg || !og || !g || !params || !workspace) return fail(DFP_ERR_INVALID, "null argument");
    int rc = check_scene(sc, cfg->stage);
    if (rc) return rc;
    SampleTables T;
    if ((rc = fill_tables(cfg, T))) return rc;
    const bool has_depth = rays->gt_depth != nullptr;
    const int S = T.n_samples + (has_depth ? T.n_surface : 0);
    const int64_t N = rays->n_rays, NS = N * S;
    if (N < 1) return fail(DFP_ERR_INVALID, "n_rays must be >= 1");
    const int64_t chunk = rays->chunk > 0 ? rays->chunk : N;
    Ws w;
    const int flags = train_flags(cfg->flags);
    const size_t need = carve_ws(w, (char*)workspace, NS, (N + chunk - 1) / chunk, true, cfg->stage, flags);
    if (need > ws_bytes) return fail(DFP_ERR_WORKSPACE, "workspace too small: need %zu bytes, have %zu", need, ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    const bool want_dp = g->rays_o || g->rays_d;

    CompArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.n_rays = N; ca.S = S; ca.stage = cfg->stage; ca.z = w.z; ca.low = w.low; ca.fused = w.fused; ca.rgb4 = w.rgb;
    ca.flags = w.flags; ca.w_attn = w.w_attn;
    ca.g_depth = og->depth; ca.g_unc = og->uncertainty; ca.g_color = og->color; ca.g_weight = og->weight;
    ca.d_low = w.d_low; ca.d_high = w.d_high; ca.d_fused = w.d_fused; ca.d_rgb4 = w.d_rgb; ca.dp = want_dp ? w.dp : nullptr;
    DFP_TIMED("k_composite_bwd", st, k_composite_bwd<<<(unsigned)((N * COMP_LANES + 255) / 256), 256, 0, st>>>(ca));
    DFP_LAUNCH_CHECK("k_composite_bwd");

    PointSrc ps;
    ps.rays_o = rays->rays_o; ps.rays_d = rays->rays_d; ps.z = w.z; ps.S = S; ps.pts = nullptr; ps.pts_f64 = 0;
    if ((rc = eval_samples_bwd<0>(sc, params, ps, NS, cfg->stage, og->weight, g, want_dp, w, flags, st))) return rc;
    if (want_dp) {
        DFP_TIMED("k_ray_grad", st, k_ray_grad<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(N, S, w.dp, w.z, g->rays_o, g->rays_d));
        DFP_LAUNCH_CHECK("k_ray_grad");
    }
    return DFP_OK;
}

int dfp_points_fwd(const DfpScene* sc, const float* params, const DfpPoints* pts, int32_t stage, float* raw, float* wout,
                   void* workspace, size_t ws_bytes, void* stream) {
    if (!pts || !params || !workspace || !raw) return fail(DFP_ERR_INVALID, "null argument");
    int rc = check_scene(sc, stage);
    if (rc) return rc;
    if (pts->n < 1 || !pts->p) return fail(DFP_ERR_INVALID, "no points");
    if (pts->n > 0x7fffffffLL) return fail(DFP_ERR_INVALID, "too many points in one call");
    Ws w;
    const int flags = train_flags(pts->flags);
    const size_t need = carve_ws(w, (char*)workspace, pts->n, 1, false, stage, flags);
    if (need > ws_bytes) return fail(DFP_ERR_WORKSPACE, "workspace too small: need %zu bytes, have %zu", need, ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    PointSrc ps;
    memset(&ps, 0, sizeof(ps));
    ps.pts = pts->p; ps.S = 1; ps.pts_f64 = pts->is_f64;
    if (pts->is_f64) rc = eval_samples_fwd<1>(sc, params, ps, pts->n, stage, pts->apply_bound_mask, w, flags, st);
    else rc = eval_samples_fwd<2>(sc, params, ps, pts->n, stage, pts->apply_bound_mask, w, flags, st);
    if (rc) return rc;
    PtsArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.n = pts->n; pa.stage = stage; pa.low = w.low; pa.fused = w.fused; pa.rgb4 = w.rgb; pa.flags = w.flags; pa.w_attn = w.w_attn;
    pa.raw = raw; pa.w = wout;
    DFP_TIMED("k_points_finalize", st, k_points_finalize<<<(unsigned)((pts->n + 255) / 256), 256, 0, st>>>(pa));
    DFP_LAUNCH_CHECK("k_points_finalize");
    return DFP_OK;
}

int dfp_points_bwd(const DfpScene* sc, const float* params, const DfpPoints* pts, int32_t stage, const float* g_raw,
                   const float* g_w, const DfpGrads* g, void* workspace, size_t ws_bytes, void* stream) {
    if (!pts || !params || !workspace || !g) return fail(DFP_ERR_INVALID, "null argument");
    int rc = check_scene(sc, stage);
    if (rc) return rc;
    if (pts->n < 1 || !pts->p) return fail(DFP_ERR_INVALID, "no points");
    Ws w;
    const int flags = train_flags(pts->flags);
    const size_t need = carve_ws(w, (char*)workspace, pts->n, 1, false, stage, flags);
    if (need > ws_bytes) return fail(DFP_ERR_WORKSPACE, "workspace too small: need %zu bytes, have %zu", need, ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    const bool want_dp = g->points != nullptr;
    PtsArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.n = pts->n; pa.stage = stage; pa.flags = w.flags; pa.g_raw = g_raw; pa.g_w = g_w;
    pa.d_low = w.d_low; pa.d_high = w.d_high; pa.d_fused = w.d_fused; pa.d_rgb4 = w.d_rgb; pa.dp = want_dp ? w.dp : nullptr;
    DFP_TIMED("k_points_seed_bwd", st, k_points_seed_bwd<<<(unsigned)((pts->n + 255) / 256), 256, 0, st>>>(pa));
    DFP_LAUNCH_CHECK("k_points_seed_bwd");
    PointSrc ps;
    memset(&ps, 0, sizeof(ps));
    ps.pts = pts->p; ps.S = 1; ps.pts_f64 = pts->is_f64;
    if (pts->is_f64) rc = eval_samples_bwd<1>(sc, params, ps, pts->n, stage, g_w, g, want_dp, w, flags, st);
    else rc = eval_samples_bwd<2>(sc, params, ps, pts->n, stage, g_w, g, want_dp, w, flags, st);
    if (rc) return rc;
    if (want_dp) {
        const unsigned blocks = (unsigned)((3 * pts->n + 255) / 256);
        if (pts->is_f64) DFP_TIMED("k_points_grad_out", st, k_points_grad_out<double><<<blocks, 256, 0, st>>>(pts->n, w.dp, (double*)g->points));
        else DFP_TIMED("k_points_grad_out", st, k_points_grad_out<float><<<blocks, 256, 0, st>>>(pts->n, w.dp, (float*)g->points));
        DFP_LAUNCH_CHECK("k_points_grad_out");
    }
    return DFP_OK;
}

int dfp_tsdf_sample(const DfpScene* sc, const DfpPoints* pts, float* out, void* stream) {
    if (!sc || !pts || !out || !pts->p) return fail(DFP_ERR_INVALID, "null argument");
    if (!sc->tsdf.data) return fail(DFP_ERR_INVALID, "tsdf volume is null");
    if (pts->n < 1) return fail(DFP_ERR_INVALID, "no points");
    PointSrc ps;
    memset(&ps, 0, sizeof(ps));
    ps.pts = pts->p; ps.S = 1; ps.pts_f64 = pts->is_f64;
    const unsigned blocks = (unsigned)((pts->n + 255) / 256);
    if (pts->is_f64) DFP_TIMED("k_tsdf_sample", (cudaStream_t)stream, k_tsdf_sample<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(ps, pts->n, to_bounds(sc->tsdf_bnds), sc->tsdf, out));
    else DFP_TIMED("k_tsdf_sample", (cudaStream_t)stream, k_tsdf_sample<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(ps, pts->n, to_bounds(sc->tsdf_bnds), sc->tsdf, out));
    DFP_LAUNCH_CHECK("k_tsdf_sample");
    return DFP_OK;
}

}  // extern "C"
