// Mapper-side kernels around the render hot path (the "next" rows):
//   (f1) masked-grid Adam step        src/Mapper.py:351-361, 374-388, 472-482
//   (f3) frustum feature selection    src/Mapper.py:90-158 (numpy + cv2.remap upstream)
//   (f2) ray batch assembly and loss  src/Mapper.py:414-469, src/common.py:76-136
// All of them are streaming byte/float kernels bounded by HBM; none touches the tensor cores.
#include "dfp_device.cuh"

namespace dfp {

// ------------------------------------------------------------------------------------------
// (f1) masked Adam.  One thread = 4 channels of one voxel; threads of a voxel share its mask byte.
// Arithmetic follows torch.optim.Adam (single-tensor path): lerp for the first moment, mul+addcmul for the second,
// denom = sqrt(v)/sqrt(bias_correction2) + eps, param += -(lr/bias_correction1) * m / denom.
// ------------------------------------------------------------------------------------------
struct AdamArgs {
    float4* param; const float4* grad; float4* m; float4* v; const uint8_t* mask;
    int64_t n4;            // n_vox * C / 4
    int c4;                // C / 4
    float w1, b2, w2, bc2_sqrt, eps, neg_step_size;
};

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamArgs& a) {
    m = __fadd_rn(m, __fmul_rn(a.w1, __fsub_rn(g, m)));                           // exp_avg.lerp_(grad, 1 - beta1)
    v = __fadd_rn(__fmul_rn(v, a.b2), __fmul_rn(__fmul_rn(a.w2, g), g));          // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), a.bc2_sqrt), a.eps);
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(a.neg_step_size, m), denom));            // addcdiv_(exp_avg, denom, -step_size)
}

__global__ void __launch_bounds__(256) k_masked_adam(const AdamArgs a) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += stride) {
        if (a.mask && !a.mask[i / a.c4]) continue;
        const float4 g = __ldg(a.grad + i);
        float4 p = a.param[i], m = a.m[i], v = a.v[i];
        adam1(p.x, g.x, m.x, v.x, a);
        adam1(p.y, g.y, m.y, v.y, a);
        adam1(p.z, g.z, m.z, v.z, a);
        adam1(p.w, g.w, m.w, v.w, a);
        a.param[i] = p; a.m[i] = m; a.v[i] = v;
    }
}

// ------------------------------------------------------------------------------------------
// (f3) frustum mask.  Vertex order upstream is x-major (meshgrid 'ij' over X, Y, Z; Mapper.py:106-110); we visit
// the vertices in the grid's own voxel order [D,H,W] and write the mask there.
// ------------------------------------------------------------------------------------------
struct FrustumArgs {
    const float* depth; int H, W;
    double fx, fy, cx, cy;
    float w2c[16]; float cam_o[3];
    const float* xs; const float* ys; const float* zs;
    int Dg, Hg, Wg;
    float* sdepth; double* negz; uint8_t* inimg; unsigned int* maxbits;
    uint8_t* mask;
};

// cv2.remap(..., INTER_LINEAR) for a CV_32FC1 image and float maps, BORDER_CONSTANT 0
// (modules/imgproc/src/imgwarp.cpp: remap -> remapBilinear): the coordinates are rounded to 1/32 pixel
// (cvRound(x * INTER_TAB_SIZE), INTER_BITS = 5), the integer part saturates to int16, the four weights come from a
// float table (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy*fx, and the sum runs left to right.
__device__ __forceinline__ int cv_round_tab(float x) {
    const float v = x * 32.f;
    if (!(fabsf(v) < 2147483520.f)) return INT_MIN;          // cvtss2si's "integer indefinite" for overflow / NaN
    return __float2int_rn(v);
}
__device__ __forceinline__ int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }
__device__ __forceinline__ float cv_remap_linear(const float* img, int H, int W, float u, float v) {
    const int su = cv_round_tab(u), sv = cv_round_tab(v);
    const int x0 = sat_short(su >> 5), y0 = sat_short(sv >> 5);
    const float fx = (float)(su & 31) * (1.f / 32.f), fy = (float)(sv & 31) * (1.f / 32.f);
    const float w00 = __fmul_rn(1.f - fy, 1.f - fx), w01 = __fmul_rn(1.f - fy, fx);
    const float w10 = __fmul_rn(fy, 1.f - fx), w11 = __fmul_rn(fy, fx);
    auto px = [&](int y, int x) -> float { return ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) ? __ldg(img + (int64_t)y * W + x) : 0.f; };
    float acc = __fmul_rn(px(y0, x0), w00);
    acc = __fadd_rn(acc, __fmul_rn(px(y0, x0 + 1), w01));
    acc = __fadd_rn(acc, __fmul_rn(px(y0 + 1, x0), w10));
    acc = __fadd_rn(acc, __fmul_rn(px(y0 + 1, x0 + 1), w11));
    return acc;
}

__global__ void __launch_bounds__(256) k_frustum_project(const FrustumArgs a) {
    const int64_t n = (int64_t)a.Dg * a.Hg * a.Wg;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float best = 0.f;
    if (i < n) {
        const int ix = (int)(i % a.Wg), iy = (int)((i / a.Wg) % a.Hg), iz = (int)(i / ((int64_t)a.Wg * a.Hg));
        const float p[3] = {a.xs[ix], a.ys[iy], a.zs[iz]};
        // cam = w2c @ [p, 1] in f32 (numpy matmul of two f32 arrays), sequential sum
        float cam[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float acc = __fmul_rn(a.w2c[4 * r], p[0]);
            acc = __fadd_rn(acc, __fmul_rn(a.w2c[4 * r + 1], p[1]));
            acc = __fadd_rn(acc, __fmul_rn(a.w2c[4 * r + 2], p[2]));
            acc = __fadd_rn(acc, a.w2c[4 * r + 3]);                       // * 1.0f
            cam[r] = acc;
        }
        cam[0] = -cam[0];                                                  // cam_cord[:, 0] *= -1
        // uv = K @ cam in f64 (K is f64, Mapper.py:119-121)
        const double c0 = (double)cam[0], c1 = (double)cam[1], c2 = (double)cam[2];
        const double u3 = __dadd_rn(__dadd_rn(__dmul_rn(a.fx, c0), __dmul_rn(0.0, c1)), __dmul_rn(a.cx, c2));
        const double v3 = __dadd_rn(__dadd_rn(__dmul_rn(0.0, c0), __dmul_rn(a.fy, c1)), __dmul_rn(a.cy, c2));
        const double w3 = __dadd_rn(__dadd_rn(__dmul_rn(0.0, c0), __dmul_rn(0.0, c1)), c2);
        const double z = w3 + 1e-5;
        const float u = (float)(u3 / z), v = (float)(v3 / z);
        const float d = cv_remap_linear(a.depth, a.H, a.W, u, v);
        a.sdepth[i] = d;
        a.negz[i] = -z;
        a.inimg[i] = (u < (float)a.W && u > 0.f && v < (float)a.H && v > 0.f) ? 1 : 0;
        best = d;
    }
    // np.max(depths): non-negative in practice; NaN/negative depths are clamped away by the uint trick's domain
    best = fmaxf(best, 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) atomicMax(a.maxbits, __float_as_uint(best));
}

__global__ void __launch_bounds__(256) k_frustum_test(const FrustumArgs a) {
    const int64_t n = (int64_t)a.Dg * a.Hg * a.Wg;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ix = (int)(i % a.Wg), iy = (int)((i / a.Wg) % a.Hg), iz = (int)(i / ((int64_t)a.Wg * a.Hg));
    float d = a.sdepth[i];
    if (d == 0.f) d = __uint_as_float(*a.maxbits);                        // depths[zero_mask] = np.max(depths)
    const double nz = a.negz[i];
    const bool vis = a.inimg[i] && (0.0 <= nz) && (nz <= (double)__fadd_rn(d, 0.5f));
    // feature grid near the camera centre: |p - o|^2 < 0.25 in f32 (Mapper.py:147-153)
    const float dx = __fsub_rn(a.xs[ix], a.cam_o[0]), dy = __fsub_rn(a.ys[iy], a.cam_o[1]), dz = __fsub_rn(a.zs[iz], a.cam_o[2]);
    const float dist = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    a.mask[i] = (vis || dist < 0.25f) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// (f2) ray batch assembly: gather + ray + box filter, then an in-order compaction (three small kernels: flags,
// single-CTA exclusive scan of per-block counts, scatter).
// ------------------------------------------------------------------------------------------
constexpr int MAX_FRAMES = 16;
struct AssembleArgs {
    DfpFrame frames[MAX_FRAMES];
    int n_frames; const int64_t* indices; int n_per_frame;
    int H0, H1, W0, W1, W;
    float fx, fy, cx, cy;
    double bound[3][2];
    int filter;
    float* rays_o; float* rays_d; float* gt_depth; float* gt_color; int32_t* n_kept;
    float* tmp;            // [n, 10] staged rays before compaction
    int32_t* block_count;  // [blocks] kept rays per block, then exclusive offsets
    int blocks;
};

__device__ __forceinline__ bool assemble_one(const AssembleArgs& a, int64_t i, float (&r)[10]) {
    const int f = (int)(i / a.n_per_frame);
    const DfpFrame& fr = a.frames[f];
    const int ww = a.W1 - a.W0;
    const int64_t idx = a.indices[i];
    const int px = a.W0 + (int)(idx % ww), py = a.H0 + (int)(idx / ww);
    // i, j come from torch.linspace(W0, W1-1, W1-W0): integers, exact in f32 (common.py:115-118)
    const float fi = (float)px, fj = (float)py;
    const float d0 = __fdiv_rn(__fsub_rn(fi, a.cx), a.fx), d1 = -__fdiv_rn(__fsub_rn(fj, a.cy), a.fy), d2 = -1.f;
    // rays_d = sum(dirs * c2w[:3,:3], -1) (common.py:89): a 3-term f32 sum per row, sequential
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float acc = __fmul_rn(d0, fr.c2w[4 * k]);
        acc = __fadd_rn(acc, __fmul_rn(d1, fr.c2w[4 * k + 1]));
        acc = __fadd_rn(acc, __fmul_rn(d2, fr.c2w[4 * k + 2]));
        r[3 + k] = acc;
        r[k] = fr.c2w[4 * k + 3];
    }
    const int64_t pix = (int64_t)py * a.W + px;
    r[6] = __ldg(fr.depth + pix);
    r[7] = __ldg(fr.color + 3 * pix); r[8] = __ldg(fr.color + 3 * pix + 1); r[9] = __ldg(fr.color + 3 * pix + 2);
    if (!a.filter) return true;
    // t = min over axes of max((bound - o) / d) >= depth  (Mapper.py:439-445); the bound is an f64 tensor upstream, so
    // the whole expression is evaluated in f64
    double t = INFINITY;
    bool nan = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double t0 = (a.bound[k][0] - (double)r[k]) / (double)r[3 + k];
        const double t1 = (a.bound[k][1] - (double)r[k]) / (double)r[3 + k];
        const double m = (t0 != t0 || t1 != t1) ? (double)NAN : fmax(t0, t1);      // torch.max / min propagate NaN
        if (m != m) nan = true;
        t = fmin(t, m);
    }
    if (nan) return false;                                                  // NaN >= depth is False
    return t >= (double)r[6];
}

__global__ void __launch_bounds__(256) k_assemble_flags(const AssembleArgs a) {
    const int64_t n = (int64_t)a.n_frames * a.n_per_frame;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    bool keep = false;
    if (i < n) {
        float r[10];
        keep = assemble_one(a, i, r);
#pragma unroll
        for (int k = 0; k < 10; ++k) a.tmp[i * 10 + k] = r[k];
        a.tmp[n * 10 + i] = keep ? 1.f : 0.f;
    }
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) a.block_count[blockIdx.x] = c;
}
__global__ void __launch_bounds__(1024) k_assemble_scan(const AssembleArgs a) {
    __shared__ int warp_sum[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < a.blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < a.blocks ? a.block_count[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_sum[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
            warp_sum[threadIdx.x] = w;
        }
        __syncthreads();
        const int incl = x + ((threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0) + carry_s;
        if (i < a.blocks) a.block_count[i] = incl - v;                      // exclusive offset
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *a.n_kept = carry_s;
}
__global__ void __launch_bounds__(256) k_assemble_scatter(const AssembleArgs a) {
    const int64_t n = (int64_t)a.n_frames * a.n_per_frame;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool keep = i < n && a.tmp[n * 10 + i] != 0.f;
    // in-block exclusive position
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    __shared__ int wcnt[8];
    if ((threadIdx.x & 31) == 0) wcnt[threadIdx.x >> 5] = __popc(bal);
    __syncthreads();
    int off = a.block_count[blockIdx.x];
    for (int w = 0; w < (threadIdx.x >> 5); ++w) off += wcnt[w];
    off += __popc(bal & ((1u << (threadIdx.x & 31)) - 1u));
    if (!keep) return;
    const float* r = a.tmp + i * 10;
#pragma unroll
    for (int k = 0; k < 3; ++k) { a.rays_o[3 * off + k] = r[k]; a.rays_d[3 * off + k] = r[3 + k]; a.gt_color[3 * off + k] = r[7 + k]; }
    a.gt_depth[off] = r[6];
}

// Mapper loss + gradient (Mapper.py:457-469)
struct LossArgs {
    const double* depth; const float* color; const float* weight; const float* gt_depth; const float* gt_color;
    int64_t n; int S; double w_color; int use_weight;
    double* g_depth; float* g_color; float* g_weight; double* loss;
};
__device__ __forceinline__ float sgnf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
__global__ void __launch_bounds__(256) k_mapper_loss(const LossArgs a) {
    const int64_t total = a.use_weight ? a.n * a.S : a.n;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        if (i < a.n) {
            const float gt = a.gt_depth[i];
            // batch_gt_depth (f32) - depth (f64) promotes to f64
            const double diff = (double)gt - a.depth[i];
            const bool on = gt > 0.f;
            if (on) acc += fabs(diff);
            if (a.g_depth) a.g_depth[i] = on ? (diff > 0.0 ? -1.0 : (diff < 0.0 ? 1.0 : 0.0)) : 0.0;
            if (a.color) {
                float ca = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float dc = __fsub_rn(a.gt_color[3 * i + k], a.color[3 * i + k]);
                    ca += fabsf(dc);
                    if (a.g_color) a.g_color[3 * i + k] = (float)(-a.w_color) * sgnf(dc);
                }
                acc += a.w_color * (double)ca;
            }
        }
        if (a.use_weight) {
            const float dw = __fsub_rn(a.weight[i], 1.f);
            acc += (double)fabsf(dw);
            if (a.g_weight) a.g_weight[i] = sgnf(dw);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += ws[w];
        atomicAdd(a.loss, s);
    }
}

}  // namespace dfp

using namespace dfp;

// ------------------------------------------------------------------------------------------
// ray generation without gradient: get_rays (src/common.py:254-272) and get_rays_from_uv (:76-91)
// ------------------------------------------------------------------------------------------
struct RayGenArgs {
    int64_t n; int W;                      // full frame: pixel r = (r % W, r / W)
    const float* pix_i; const float* pix_j; // or explicit pixel coordinates (null = full frame)
    float fx, fy, cx, cy;
    const float* c2w; int row_stride;      // device pose, rows 0..2 x columns 0..3
    float* rays_o; float* rays_d;
};
__global__ void k_gen_rays(const RayGenArgs a) {
    __shared__ float m[12];
    if (threadIdx.x < 12) m[threadIdx.x] = a.c2w[(threadIdx.x >> 2) * a.row_stride + (threadIdx.x & 3)];
    __syncthreads();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.n; r += (int64_t)gridDim.x * blockDim.x) {
        // i, j of the full frame come from torch.linspace(0, W-1, W): integers, exact in f32
        const float fi = a.pix_i ? a.pix_i[r] : (float)(r % a.W), fj = a.pix_j ? a.pix_j[r] : (float)(r / a.W);
        const float d0 = __fdiv_rn(__fsub_rn(fi, a.cx), a.fx), d1 = -__fdiv_rn(__fsub_rn(fj, a.cy), a.fy), d2 = -1.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {      // rays_d = sum(dirs * c2w[:3,:3], -1): a sequential 3-term f32 sum per row
            float acc = __fmul_rn(d0, m[4 * k]);
            acc = __fadd_rn(acc, __fmul_rn(d1, m[4 * k + 1]));
            acc = __fadd_rn(acc, __fmul_rn(d2, m[4 * k + 2]));
            a.rays_d[3 * r + k] = acc;
            a.rays_o[3 * r + k] = m[4 * k + 3];
        }
    }
}

extern "C" {

int dfp_gen_rays(int64_t n, int32_t W, const float* pix_i, const float* pix_j, double fx, double fy, double cx, double cy,
                 const float* c2w_dev, int32_t row_stride, float* rays_o, float* rays_d, void* stream) {
    if (!c2w_dev || !rays_o || !rays_d) return fail(DFP_ERR_INVALID, "null argument");
    if (n < 1 || (!pix_i && W < 1) || ((pix_i == nullptr) != (pix_j == nullptr))) return fail(DFP_ERR_INVALID, "bad ray generation request");
    RayGenArgs a;
    a.n = n; a.W = W; a.pix_i = pix_i; a.pix_j = pix_j;
    a.fx = (float)fx; a.fy = (float)fy; a.cx = (float)cx; a.cy = (float)cy;
    a.c2w = c2w_dev; a.row_stride = row_stride; a.rays_o = rays_o; a.rays_d = rays_d;
    const int64_t want = (n + 255) / 256;
    const int blocks = (int)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
    DFP_TIMED("k_gen_rays", (cudaStream_t)stream, k_gen_rays<<<blocks, 256, 0, (cudaStream_t)stream>>>(a));
    DFP_LAUNCH_CHECK("k_gen_rays");
    return DFP_OK;
}

int dfp_masked_adam_step(const DfpAdam* a, void* stream) {
    if (!a || !a->param || !a->grad || !a->exp_avg || !a->exp_avg_sq) return fail(DFP_ERR_INVALID, "null argument");
    if (a->n_vox < 1 || a->C < 4 || (a->C & 3)) return fail(DFP_ERR_INVALID, "n_vox must be >= 1 and C a multiple of 4");
    if (a->step < 1) return fail(DFP_ERR_INVALID, "Adam step is 1-based");
    if ((((uintptr_t)a->param) | ((uintptr_t)a->grad) | ((uintptr_t)a->exp_avg) | ((uintptr_t)a->exp_avg_sq)) & 15)
        return fail(DFP_ERR_INVALID, "buffers must be 16-byte aligned");
    AdamArgs k;
    k.param = reinterpret_cast<float4*>(a->param); k.grad = reinterpret_cast<const float4*>(a->grad);
    k.m = reinterpret_cast<float4*>(a->exp_avg); k.v = reinterpret_cast<float4*>(a->exp_avg_sq);
    k.mask = a->voxel_mask; k.c4 = a->C / 4; k.n4 = a->n_vox * k.c4;
    // the scalars are formed in double exactly as torch's Python code does, then narrowed to the tensor dtype
    const double bc1 = 1.0 - pow(a->beta1, (double)a->step), bc2 = 1.0 - pow(a->beta2, (double)a->step);
    k.w1 = (float)(1.0 - a->beta1); k.b2 = (float)a->beta2; k.w2 = (float)(1.0 - a->beta2);
    k.bc2_sqrt = (float)sqrt(bc2); k.eps = (float)a->eps; k.neg_step_size = (float)(-(a->lr / bc1));
    const int64_t want = (k.n4 + 255) / 256;
    const int blocks = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
    DFP_TIMED("k_masked_adam", (cudaStream_t)stream, k_masked_adam<<<blocks, 256, 0, (cudaStream_t)stream>>>(k));
    DFP_LAUNCH_CHECK("k_masked_adam");
    return DFP_OK;
}

size_t dfp_frustum_scratch_bytes(int32_t Dg, int32_t Hg, int32_t Wg) {
    const size_t n = (size_t)Dg * Hg * Wg;
    return ((n * 8 + 255) & ~(size_t)255) + ((n * 4 + 255) & ~(size_t)255) + ((n + 255) & ~(size_t)255) + 256;
}

int dfp_frustum_mask(const DfpFrustum* f, uint8_t* mask, void* scratch, size_t scratch_bytes, void* stream) {
    if (!f || !mask || !scratch || !f->depth || !f->xs || !f->ys || !f->zs) return fail(DFP_ERR_INVALID, "null argument");
    if (f->Dg < 1 || f->Hg < 1 || f->Wg < 1 || f->H < 1 || f->W < 1) return fail(DFP_ERR_INVALID, "empty grid or image");
    if (scratch_bytes < dfp_frustum_scratch_bytes(f->Dg, f->Hg, f->Wg)) return fail(DFP_ERR_INVALID, "scratch too small");
    if (((uintptr_t)scratch) & 255) return fail(DFP_ERR_INVALID, "scratch must be 256-byte aligned");
    const size_t n = (size_t)f->Dg * f->Hg * f->Wg;
    FrustumArgs a;
    a.depth = f->depth; a.H = f->H; a.W = f->W; a.fx = f->fx; a.fy = f->fy; a.cx = f->cx; a.cy = f->cy;
    memcpy(a.w2c, f->w2c, sizeof(a.w2c)); memcpy(a.cam_o, f->cam_o, sizeof(a.cam_o));
    a.xs = f->xs; a.ys = f->ys; a.zs = f->zs; a.Dg = f->Dg; a.Hg = f->Hg; a.Wg = f->Wg;
    unsigned char* p = reinterpret_cast<unsigned char*>(scratch);
    a.negz = reinterpret_cast<double*>(p); p += (n * 8 + 255) & ~(size_t)255;
    a.sdepth = reinterpret_cast<float*>(p); p += (n * 4 + 255) & ~(size_t)255;
    a.inimg = p; p += (n + 255) & ~(size_t)255;
    a.maxbits = reinterpret_cast<unsigned int*>(p);
    a.mask = mask;
    cudaStream_t st = (cudaStream_t)stream;
    DFP_CUDA(cudaMemsetAsync(a.maxbits, 0, 4, st));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    DFP_TIMED("k_frustum_project", st, k_frustum_project<<<blocks, 256, 0, st>>>(a));
    DFP_LAUNCH_CHECK("k_frustum_project");
    DFP_TIMED("k_frustum_test", st, k_frustum_test<<<blocks, 256, 0, st>>>(a));
    DFP_LAUNCH_CHECK("k_frustum_test");
    return DFP_OK;
}

size_t dfp_assemble_scratch_bytes(int64_t n_total) {
    if (n_total < 1) return 0;
    const size_t blocks = (size_t)((n_total + 255) / 256);
    return (((size_t)n_total * 11 * 4 + 255) & ~(size_t)255) + ((blocks * 4 + 255) & ~(size_t)255);
}

int dfp_assemble_rays(const DfpRayBatch* b, const DfpRayBatchOut* out, void* scratch, size_t scratch_bytes, void* stream) {
    if (!b || !out || !scratch || !b->frames || !b->indices) return fail(DFP_ERR_INVALID, "null argument");
    if (!out->rays_o || !out->rays_d || !out->gt_depth || !out->gt_color || !out->n_kept) return fail(DFP_ERR_INVALID, "null output");
    if (b->n_frames < 1 || b->n_frames > MAX_FRAMES) return fail(DFP_ERR_INVALID, "n_frames must be in [1, %d]", MAX_FRAMES);
    if (b->n_per_frame < 1) return fail(DFP_ERR_INVALID, "n_per_frame must be >= 1");
    if (b->H1 <= b->H0 || b->W1 <= b->W0 || b->W < b->W1) return fail(DFP_ERR_INVALID, "bad sampling window");
    const int64_t n = (int64_t)b->n_frames * b->n_per_frame;
    if (scratch_bytes < dfp_assemble_scratch_bytes(n)) return fail(DFP_ERR_INVALID, "scratch too small");
    AssembleArgs a;
    for (int f = 0; f < b->n_frames; ++f) {
        if (!b->frames[f].depth || !b->frames[f].color) return fail(DFP_ERR_INVALID, "frame %d has a null image", f);
        a.frames[f] = b->frames[f];
    }
    a.n_frames = b->n_frames; a.indices = b->indices; a.n_per_frame = b->n_per_frame;
    a.H0 = b->H0; a.H1 = b->H1; a.W0 = b->W0; a.W1 = b->W1; a.W = b->W;
    // the intrinsics enter f32 tensor arithmetic as Python scalars upstream
    a.fx = (float)b->fx; a.fy = (float)b->fy; a.cx = (float)b->cx; a.cy = (float)b->cy;
    for (int k = 0; k < 3; ++k) { a.bound[k][0] = b->bound[k][0]; a.bound[k][1] = b->bound[k][1]; }
    a.filter = b->filter;
    a.rays_o = out->rays_o; a.rays_d = out->rays_d; a.gt_depth = out->gt_depth; a.gt_color = out->gt_color; a.n_kept = out->n_kept;
    a.blocks = (int)((n + 255) / 256);
    a.tmp = reinterpret_cast<float*>(scratch);
    a.block_count = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(scratch) + (((size_t)n * 11 * 4 + 255) & ~(size_t)255));
    cudaStream_t st = (cudaStream_t)stream;
    DFP_TIMED("k_assemble_flags", st, k_assemble_flags<<<a.blocks, 256, 0, st>>>(a));
    DFP_LAUNCH_CHECK("k_assemble_flags");
    DFP_TIMED("k_assemble_scan", st, k_assemble_scan<<<1, 1024, 0, st>>>(a));
    DFP_LAUNCH_CHECK("k_assemble_scan");
    DFP_TIMED("k_assemble_scatter", st, k_assemble_scatter<<<a.blocks, 256, 0, st>>>(a));
    DFP_LAUNCH_CHECK("k_assemble_scatter");
    return DFP_OK;
}

int dfp_mapper_loss(const DfpMapLoss* l, void* stream) {
    if (!l || !l->depth || !l->gt_depth || !l->loss) return fail(DFP_ERR_INVALID, "null argument");
    if (l->n < 1) return fail(DFP_ERR_INVALID, "no rays");
    if (l->use_weight && (!l->weight || l->S < 1)) return fail(DFP_ERR_INVALID, "weight term requested without weights");
    if (l->color && !l->gt_color) return fail(DFP_ERR_INVALID, "colour term requested without gt colour");
    LossArgs a;
    a.depth = l->depth; a.color = l->w_color != 0.0 ? l->color : nullptr; a.weight = l->weight; a.gt_depth = l->gt_depth; a.gt_color = l->gt_color;
    a.n = l->n; a.S = l->S; a.w_color = l->w_color; a.use_weight = l->use_weight;
    a.g_depth = l->g_depth; a.g_color = l->g_color; a.g_weight = l->g_weight; a.loss = l->loss;
    const int64_t total = a.use_weight ? a.n * a.S : a.n;
    const int64_t want = (total + 255) / 256;
    const int blocks = (int)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
    DFP_TIMED("k_mapper_loss", (cudaStream_t)stream, k_mapper_loss<<<blocks, 256, 0, (cudaStream_t)stream>>>(a));
    DFP_LAUNCH_CHECK("k_mapper_loss");
    return DFP_OK;
}

}  // extern "C"
