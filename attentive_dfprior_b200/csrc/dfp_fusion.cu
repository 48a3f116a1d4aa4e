// Offline TSDF fusion of RGB-D frames into the prior volume: the `integrate` kernel of
// src/fusion.py:69-142 for sm_100a, and the Mesher's seen / forecast / unseen point classification
// (src/utils/Mesher.py:58-204) as one kernel over points x keyframes.
#include "dfp_device.cuh"

namespace dfp {

struct FuseArgs {
    float* tsdf; float* weight; float* color;
    int dx, dy, dz; float ox, oy, oz; float voxel, trunc;
    const float* depth; const float* color_im; int H, W;
    float K[9]; float pose[16]; float obs_weight;
};

// One thread per voxel, grid-stride; the arithmetic of the upstream kernel operation by operation in f32 (including its
// float voxel-index decomposition), without fused multiply-adds so that a numpy restatement reproduces it bit for bit.
__global__ void __launch_bounds__(256) k_tsdf_integrate(const FuseArgs a) {
    const int64_t total = (int64_t)a.dx * a.dy * a.dz;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        const int idx = (int)v;
        const float vx = floorf(__fdiv_rn((float)idx, (float)(a.dy * a.dz)));                       // fusion.py:89
        const float vy = floorf(__fdiv_rn((float)(idx - ((int)vx) * a.dy * a.dz), (float)a.dz));   // :90
        const float vz = (float)(idx - ((int)vx) * a.dy * a.dz - ((int)vy) * a.dz);                 // :91
        const float px = __fadd_rn(a.ox, __fmul_rn(vx, a.voxel)), py = __fadd_rn(a.oy, __fmul_rn(vy, a.voxel)),
                    pz = __fadd_rn(a.oz, __fmul_rn(vz, a.voxel));                                 // :94-96
        const float tx = __fsub_rn(px, a.pose[3]), ty = __fsub_rn(py, a.pose[7]), tz = __fsub_rn(pz, a.pose[11]);   // :98-100
        const float cxp = __fadd_rn(__fadd_rn(__fmul_rn(a.pose[0], tx), __fmul_rn(a.pose[4], ty)), __fmul_rn(a.pose[8], tz));   // :101-103
        const float cyp = __fadd_rn(__fadd_rn(__fmul_rn(a.pose[1], tx), __fmul_rn(a.pose[5], ty)), __fmul_rn(a.pose[9], tz));
        const float czp = __fadd_rn(__fadd_rn(__fmul_rn(a.pose[2], tx), __fmul_rn(a.pose[6], ty)), __fmul_rn(a.pose[10], tz));
        const int pixx = (int)roundf(__fadd_rn(__fmul_rn(a.K[0], __fdiv_rn(cxp, czp)), a.K[2]));     // :105-106
        const int pixy = (int)roundf(__fadd_rn(__fmul_rn(a.K[4], __fdiv_rn(cyp, czp)), a.K[5]));
        if (pixx < 0 || pixx >= a.W || pixy < 0 || pixy >= a.H || czp < 0.f) continue;               // :110-111
        const float dv = a.depth[pixy * a.W + pixx];
        if (dv == 0.f) continue;                                                                     // :114-115
        const float diff = __fsub_rn(dv, czp);
        if (diff < -a.trunc) continue;                                                               // :119-120
        const float dist = fminf(1.0f, __fdiv_rn(diff, a.trunc));
        const float w_old = a.weight[v];
        const float w_new = __fadd_rn(w_old, a.obs_weight);
        a.weight[v] = w_new;
        a.tsdf[v] = __fdiv_rn(__fadd_rn(__fmul_rn(a.tsdf[v], w_old), __fmul_rn(a.obs_weight, dist)), w_new);   // :125
        // colour: running average per channel of the b*65536 + g*256 + r packing (:127-138)
        const float oc = a.color[v];
        const float ob = floorf(__fdiv_rn(oc, 65536.f));
        const float og = floorf(__fdiv_rn(__fsub_rn(oc, __fmul_rn(ob, 65536.f)), 256.f));
        const float orr = __fsub_rn(__fsub_rn(oc, __fmul_rn(ob, 65536.f)), __fmul_rn(og, 256.f));
        const float nc = a.color_im[pixy * a.W + pixx];
        float nb = floorf(__fdiv_rn(nc, 65536.f));
        float ng = floorf(__fdiv_rn(__fsub_rn(nc, __fmul_rn(nb, 65536.f)), 256.f));
        float nr = __fsub_rn(__fsub_rn(nc, __fmul_rn(nb, 65536.f)), __fmul_rn(ng, 256.f));
        nb = fminf(roundf(__fdiv_rn(__fadd_rn(__fmul_rn(ob, w_old), __fmul_rn(a.obs_weight, nb)), w_new)), 255.0f);
        ng = fminf(roundf(__fdiv_rn(__fadd_rn(__fmul_rn(og, w_old), __fmul_rn(a.obs_weight, ng)), w_new)), 255.0f);
        nr = fminf(roundf(__fdiv_rn(__fadd_rn(__fmul_rn(orr, w_old), __fmul_rn(a.obs_weight, nr)), w_new)), 255.0f);
        a.color[v] = __fadd_rn(__fadd_rn(__fmul_rn(nb, 65536.f), __fmul_rn(ng, 256.f)), nr);
    }
}

// ---- Mesher.point_masks: seen / forecast per point over the keyframes (frustum test, optional depth bound) ----
constexpr int PM_MAX_FRAMES = 64;
struct PointMaskArgs {
    const float* points; int64_t n;
    int n_frames; const float* w2c;          // [n_frames][16] f32 (np.linalg.inv of the pose, computed by the caller as upstream)
    const float* max_depth;                  // [n_frames]: torch.max(keyframe depth) * 1.1 (depth_test = False, Mesher.py:186)
    float fx, fy, cx, cy; int H, W;
    uint8_t* seen; uint8_t* forecast;
};
__global__ void __launch_bounds__(256) k_point_masks(const PointMaskArgs a) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = a.points[3 * i], y = a.points[3 * i + 1], z = a.points[3 * i + 2];
        bool seen = false, fore = false;
        for (int f = 0; f < a.n_frames; ++f) {
            const float* m = a.w2c + 16 * f;
            // cam = w2c @ [x y z 1]  (4-term f32 dot products, left to right), then x *= -1 (Mesher.py:140-147)
            float c0 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z)), m[3]);
            const float c1 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], x), __fmul_rn(m[5], y)), __fmul_rn(m[6], z)), m[7]);
            const float c2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[8], x), __fmul_rn(m[9], y)), __fmul_rn(m[10], z)), m[11]);
            c0 = -c0;
            // uv = K @ cam; z = uv[2] + 1e-8; uv = uv[:2] / z  (:148-150)
            const float u0 = __fadd_rn(__fmul_rn(a.fx, c0), __fmul_rn(a.cx, c2));
            const float u1 = __fadd_rn(__fmul_rn(a.fy, c1), __fmul_rn(a.cy, c2));
            const float zz = __fadd_rn(c2, 1e-8f);
            const float u = __fdiv_rn(u0, zz), v = __fdiv_rn(u1, zz);
            const bool front = zz < 0.f;
            bool s = (u < (float)a.W) && (u > 0.f) && (v < (float)a.H) && (v > 0.f) && front;              // edge = 0
            bool fo = (u < (float)(a.W + 1000)) && (u > -1000.f) && (v < (float)(a.H + 1000)) && (v > -1000.f) && front;   // edge = -1000
            const float proj = -c2;
            const float md = a.max_depth[f];
            s = s && (proj < md);                                                                            // :188-201
            fo = fo && (proj < md);
            seen = seen || s; fore = fore || fo;
        }
        a.seen[i] = seen ? 1 : 0;
        a.forecast[i] = (fore && !seen) ? 1 : 0;                                                             // :206
    }
}

}  // namespace dfp

using namespace dfp;
extern "C" {

int dfp_tsdf_integrate(const DfpTsdfFuse* f, void* stream) {
    if (!f || !f->tsdf_vol || !f->weight_vol || !f->color_vol || !f->depth_im || !f->color_im) return fail(DFP_ERR_INVALID, "null argument");
    if (f->dim[0] < 1 || f->dim[1] < 1 || f->dim[2] < 1 || f->H < 1 || f->W < 1) return fail(DFP_ERR_INVALID, "empty volume or image");
    if ((int64_t)f->dim[0] * f->dim[1] * f->dim[2] > 0x7fffffffLL) return fail(DFP_ERR_INVALID, "volume too large (int voxel index upstream)");
    FuseArgs a;
    a.tsdf = f->tsdf_vol; a.weight = f->weight_vol; a.color = f->color_vol;
    a.dx = f->dim[0]; a.dy = f->dim[1]; a.dz = f->dim[2];
    a.ox = f->origin[0]; a.oy = f->origin[1]; a.oz = f->origin[2]; a.voxel = f->voxel_size; a.trunc = f->trunc_margin;
    a.depth = f->depth_im; a.color_im = f->color_im; a.H = f->H; a.W = f->W;
    memcpy(a.K, f->intr, sizeof(a.K)); memcpy(a.pose, f->pose, sizeof(a.pose)); a.obs_weight = f->obs_weight;
    const int64_t total = (int64_t)a.dx * a.dy * a.dz;
    const int64_t want = (total + 255) / 256;
    const int blocks = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
    DFP_TIMED("k_tsdf_integrate", (cudaStream_t)stream, k_tsdf_integrate<<<blocks, 256, 0, (cudaStream_t)stream>>>(a));
    DFP_LAUNCH_CHECK("k_tsdf_integrate");
    return DFP_OK;
}

int dfp_point_masks(const float* points, int64_t n, const float* w2c, const float* max_depth, int32_t n_frames, double fx, double fy,
                    double cx, double cy, int32_t H, int32_t W, uint8_t* seen, uint8_t* forecast, void* stream) {
    if (!points || !w2c || !max_depth || !seen || !forecast) return fail(DFP_ERR_INVALID, "null argument");
    if (n < 1 || n_frames < 1) return fail(DFP_ERR_INVALID, "no points or no frames");
    PointMaskArgs a;
    a.points = points; a.n = n; a.n_frames = n_frames; a.w2c = w2c; a.max_depth = max_depth;
    a.fx = (float)fx; a.fy = (float)fy; a.cx = (float)cx; a.cy = (float)cy; a.H = H; a.W = W; a.seen = seen; a.forecast = forecast;
    const int64_t want = (n + 255) / 256;
    const int blocks = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
    DFP_TIMED("k_point_masks", (cudaStream_t)stream, k_point_masks<<<blocks, 256, 0, (cudaStream_t)stream>>>(a));
    DFP_LAUNCH_CHECK("k_point_masks");
    return DFP_OK;
}

}  // extern "C"
