/*
 * dfprior_b200.h -- C ABI of libdfprior_b200.so
 *
 * B200-native (sm_100a) implementation of DF-Prior's per-ray volume-rendering
 * hot path.  The reference (MachinePerceptionLab/Attentive_DFPrior) is pure
 * Python/PyTorch and has no FFI layer of its own; its seam for this path is
 * the set of Python call signatures cited beside each entry point below
 * (paths relative to the reference checkout).  A binding only needs ctypes
 * (see INTEGRATION.md); no torch types cross this boundary.
 *
 * Conventions
 *  - every function returns 0 (DFP_OK) or a negative DfpStatus; nothing throws
 *    or aborts across the boundary; dfp_last_error() gives a thread-local text.
 *  - all buffers are caller-owned DEVICE pointers unless stated otherwise; the
 *    library never allocates device memory and never synchronises the stream.
 *  - kernels are enqueued on the caller's stream (cudaStream_t passed as void*).
 *  - gradient buffers are ACCUMULATED INTO (caller zeroes them), so a caller
 *    may chunk a batch over several calls.
 *  - stateless and re-entrant: forward and backward may come from different
 *    host threads (autograd engine thread); no environment variable or setter
 *    changes what a call computes (the only process-wide state is the launch
 *    counter and the opt-in event timing of the diagnostics section).
 */
#ifndef DFPRIOR_B200_H
#define DFPRIOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFP_ABI_VERSION 9

#define DFP_C_DIM 32            /* configs/df_prior.yaml:102 (model.c_dim) */
#define DFP_HIDDEN 32           /* src/conv_onet/models/decoder.py:278 */
#define DFP_EMBED 93            /* decoder.py:129 */
#define DFP_N_PARAM_TENSORS 79  /* DF.state_dict() */
#define DFP_N_PARAMS 86029
#define DFP_MAX_RAY_SAMPLES 64  /* N_samples + N_surface (32 + 16 in every shipped config) */

typedef enum DfpStatus {
    DFP_OK = 0,
    DFP_ERR_INVALID = -1,      /* bad argument (null pointer, unsupported shape, ...) */
    DFP_ERR_CUDA = -2,         /* a CUDA runtime call failed */
    DFP_ERR_WORKSPACE = -3,    /* workspace too small */
    DFP_ERR_UNSUPPORTED = -4   /* e.g. N_importance > 0 (dead + broken upstream, Renderer.py:235-252) */
} DfpStatus;

typedef enum DfpStage { DFP_STAGE_LOW = 0, DFP_STAGE_HIGH = 1, DFP_STAGE_COLOR = 2 } DfpStage;

/* forward-call flags */
#define DFP_FLAG_SAVE_ACT 1    /* a backward call will follow: the tensor-core forward keeps the decoders' layer
                                  activations in the (larger) workspace so that the backward need not recompute them */
#define DFP_FLAG_SIMT 2        /* evaluate the MLPs with the fp32 SIMT kernels instead of the tcgen05 tensor-core
                                  kernels (in-tree cross-check of the product path; several times slower).  Like
                                  every flag it belongs to the call: pass the same value to the workspace-size
                                  function, the forward and the backward.  The library keeps no switches. */
#define DFP_FLAG_FROZEN_LOW 4   /* with DFP_FLAG_SAVE_ACT: no PARAMETER gradient is wanted for the low / high / colour */
#define DFP_FLAG_FROZEN_HIGH 8  /* decoder (none of its tensors requires grad -- the reference's Mapper never optimises the */
#define DFP_FLAG_FROZEN_COLOR 16 /* low decoder and, with fix_high / fix_color, leaves those out too: src/Mapper.py:364-371).
                                  The forward then keeps no weight-gradient operands for that decoder and the backward
                                  skips its weight-gradient kernel; its rows of DfpGrads.params_packed stay untouched.
                                  Grid gradients, the fusion MLP's gradients and d(rays) are unaffected. */

/* One feature grid.  Logical shape is the reference's [1, C, D, H, W]
 * (src/DF_Prior.py:243-266) but memory is channels-last: element (z, y, x, c)
 * lives at ((z*H + y)*W + x)*C + c, i.e. torch.channels_last_3d strides.
 * dfp_grid_to_ndhwc() converts a channels-first grid. */
typedef struct DfpGrid {
    const float* data;
    int32_t D, H, W, C;        /* C must be DFP_C_DIM */
} DfpGrid;

/* The fused TSDF prior volume, logical [1,1,D,H,W] = [.,.,Z,Y,X] with explicit
 * element strides (the reference keeps a non-contiguous permuted view,
 * get_tsdf.py:95-97). */
typedef struct DfpTsdf {
    const float* data;
    int32_t D, H, W;
    int64_t sD, sH, sW;
} DfpTsdf;

typedef struct DfpScene {
    DfpGrid low, high, color;  /* c['grid_low'|'grid_high'|'grid_color'] */
    DfpTsdf tsdf;
    double bound[3][2];        /* slam.bound after rounding, src/DF_Prior.py:185-190 */
    double tsdf_bnds[3][2];    /* slam.vol_bnds, src/DF_Prior.py:87 */
} DfpScene;

/* A batch of rays (Renderer.render_batch_ray arguments, src/utils/Renderer.py:110). */
typedef struct DfpRays {
    const float* rays_o;       /* [n_rays,3] f32 */
    const float* rays_d;       /* [n_rays,3] f32 */
    const float* gt_depth;     /* [n_rays] f32 or NULL (=> N_surface = 0, near = 0.01; Renderer.py:140-142) */
    int64_t n_rays;
    int64_t chunk;             /* rays per batch-global max(gt_depth) group: n_rays for one
                                  render_batch_ray call, ray_batch_size (100000) for render_img
                                  (Renderer.py:159,195,294-306); 0 means n_rays */
    const float* depth_max;    /* optional device scalar: max(gt_depth) of the WHOLE batch when this call
                                  holds only a shard of it (ray-sharded multi-GPU: all-reduce(max) first);
                                  NULL = computed from gt_depth per chunk */
} DfpRays;

/* cfg['rendering'] (configs/df_prior.yaml:93-98).  The linspace tables are passed in from
 * torch so that both sides use the very same constants (Renderer.py:183-184, 203). */
typedef struct DfpCfg {
    int32_t stage;             /* DfpStage */
    int32_t n_samples;         /* N_samples (32) */
    int32_t n_surface;         /* N_surface (16); ignored when gt_depth == NULL */
    int32_t n_importance;      /* must be 0 */
    int32_t flags;             /* DFP_FLAG_* (pass the same value to the forward, the backward and the workspace size) */
    int32_t reserved;
    float t_vals[DFP_MAX_RAY_SAMPLES];              /* linspace(0,1,N_samples) f32 */
    float one_minus_t[DFP_MAX_RAY_SAMPLES];         /* 1 - t_vals, f32 */
    double t_surf[DFP_MAX_RAY_SAMPLES];             /* linspace(0,1,N_surface) f32 -> f64 */
    double one_minus_t_surf[DFP_MAX_RAY_SAMPLES];   /* 1 - t_surf, f64 */
} DfpCfg;

/* render_batch_ray return tuple (Renderer.py:255): depth, uncertainty, color, weight. */
typedef struct DfpRenderOut {
    double* depth;             /* [n_rays] f64 */
    double* uncertainty;       /* [n_rays] f64 */
    float* color;              /* [n_rays,3] f32 */
    float* weight;             /* [n_rays,S] f32 attention weight (NOT the compositing weight) */
    double* z_vals;            /* optional [n_rays,S] f64 sorted sample depths (parity checks); may be NULL */
    float* raw;                /* optional [n_rays,S,4] f32 pre-sigmoid raw (parity checks); may be NULL */
} DfpRenderOut;

/* upstream gradients of the four outputs; any pointer may be NULL (= zero) */
typedef struct DfpRenderOutGrad {
    const double* depth;       /* [n_rays] */
    const double* uncertainty; /* [n_rays] */
    const float* color;        /* [n_rays,3] */
    const float* weight;       /* [n_rays,S] */
} DfpRenderOutGrad;

/* gradient sinks; NULL = not needed (the kernels skip that work) */
typedef struct DfpGrads {
    float* grid_low;           /* NDHWC, same shape as the grid, accumulated */
    float* grid_high;
    float* grid_color;
    float* params_packed;      /* packed layout (dfp_packed_grad_floats() floats), accumulated */
    float* rays_o;             /* [n_rays,3] f32, overwritten */
    float* rays_d;             /* [n_rays,3] f32, overwritten */
    void* points;              /* dfp_points_bwd only: [n,3] in the points' dtype, overwritten */
    /* optional cudaEvent_t handles (NULL = none), recorded on the call's stream by dfp_render_bwd / dfp_points_bwd:
     * ev_high once grid_high's gradient is final (the fusion MLP's and the high decoder's backward are enqueued),
     * ev_color once grid_color's is.  A multi-GPU caller starts the all-reduce of those buckets on a side stream
     * while the remaining decoders' backward still runs (grid_low and params_packed are final when the call's work is). */
    void* ev_high;
    void* ev_color;
} DfpGrads;

/* A batch of query points (Renderer.eval_points / Mesher.eval_points / DF.forward). */
typedef struct DfpPoints {
    const void* p;             /* [n,3] */
    int32_t is_f64;            /* 1: f64 (ray path), 0: f32 (Mesher grid points) */
    int32_t apply_bound_mask;  /* 1: eval_points semantics (occ=100 outside bound, Renderer.py:51-64);
                                  0: bare DF.forward (decoder.py:307-353) */
    int64_t n;
    int32_t flags;             /* DFP_FLAG_* */
    int32_t reserved;
} DfpPoints;

int dfp_abi_version(void);
const char* dfp_last_error(void);

/* ---- decoder parameters ------------------------------------------------------------
 * The 79 tensors of DF.state_dict() in state_dict order (host array of device pointers)
 * are gathered into one kernel-friendly packed buffer (k-major weights, padded rows);
 * dfp_unpack_param_grads scatters a packed gradient back into a flat buffer holding
 * the 79 gradients back to back in native [out,in] layout (DFP_N_PARAMS floats). */
int64_t dfp_packed_param_floats(void);   /* size of the packed PARAMETER buffer (fp32 parameters + tensor-core operand images) */
int64_t dfp_packed_grad_floats(void);    /* size of a packed GRADIENT buffer (DfpGrads.params_packed): the fp32 part only */
int dfp_pack_params(const float* const* tensors79, float* packed, void* stream);
int dfp_unpack_param_grads(const float* packed_grad, float* native_flat, void* stream);

/* channels-first [C,D,H,W] <-> channels-last [D,H,W,C] feature grid transposes */
int dfp_grid_to_ndhwc(const float* src_cdhw, float* dst_dhwc, int32_t C, int32_t D, int32_t H, int32_t W, void* stream);
int dfp_grid_to_cdhw(const float* src_dhwc, float* dst_cdhw, int32_t C, int32_t D, int32_t H, int32_t W, void* stream);

/* ---- Renderer.render_batch_ray / render_img  (src/utils/Renderer.py:110-255, 258-327) ----
 * The workspace written by the forward call is what the backward call reads; keep it
 * alive (and unmodified) between the two. */
size_t dfp_render_workspace_bytes(int64_t n_rays, int32_t samples_per_ray, int64_t chunk, int32_t stage, int32_t flags);
int dfp_render_fwd(const DfpScene* scene, const float* params_packed, const DfpRays* rays, const DfpCfg* cfg,
                   const DfpRenderOut* out, void* workspace, size_t workspace_bytes, void* stream);
int dfp_render_bwd(const DfpScene* scene, const float* params_packed, const DfpRays* rays, const DfpCfg* cfg,
                   const DfpRenderOutGrad* out_grad, const DfpGrads* grads,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- Renderer.eval_points / Mesher.eval_points / DF.forward --------------------------
 * (src/utils/Renderer.py:27-71, src/utils/Mesher.py:286-326, decoder.py:307-353)
 * raw [n,4] f32, w [n] f32. */
size_t dfp_points_workspace_bytes(int64_t n_points, int32_t stage, int32_t flags);
int dfp_points_fwd(const DfpScene* scene, const float* params_packed, const DfpPoints* pts, int32_t stage,
                   float* raw, float* w, void* workspace, size_t workspace_bytes, void* stream);
int dfp_points_bwd(const DfpScene* scene, const float* params_packed, const DfpPoints* pts, int32_t stage,
                   const float* grad_raw, const float* grad_w, const DfpGrads* grads,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- Renderer.eval_points_tsdf / sample_grid_tsdf (Renderer.py:73-107) ---------------- */
int dfp_tsdf_sample(const DfpScene* scene, const DfpPoints* pts, float* tsdf_out, void* stream);

/* ---- ray generation without gradient: get_rays (src/common.py:254-272, every pixel of an H x W frame in row-major
 * order; pix_i = pix_j = NULL) and get_rays_from_uv (src/common.py:76-91, explicit pixel coordinates as f32 device
 * arrays).  The pose is read from device memory (rows 0..2 of a row-major c2w with `row_stride` floats per row), so a
 * render_img call makes no host round trip.  The Tracker's differentiable use of get_rays_from_uv stays with autograd. */
int dfp_gen_rays(int64_t n_rays, int32_t W, const float* pix_i, const float* pix_j, double fx, double fy, double cx, double cy,
                 const float* c2w_dev, int32_t row_stride, float* rays_o, float* rays_d, void* stream);

/* ---- Mapper-side rows --------------------------------------------
 *
 * (f1) Masked-grid optimiser step, src/Mapper.py:351-361 (val_grad = val[mask]), :374-378 (one Adam,
 * per-group lr set per stage :392-398), :383-388 (val[mask] = val_grad), :472-482 (step, zero_grad,
 * write-back).  One kernel updates the dense channels-last grid in place for the voxels of the mask:
 * torch.optim.Adam's arithmetic (betas, eps, bias correction by `step`, no weight decay / amsgrad) on
 * dense moment buffers; voxels outside the mask, their moments and their gradients are not touched.
 * lr = 0 still advances the moments, as torch does.  voxel_mask NULL = every voxel (the
 * frustum_feature_selection = False branch, Mapper.py:333-342). */
typedef struct {
    float* param;              /* [n_vox, C] grid values (NDHWC), updated in place */
    const float* grad;         /* [n_vox, C] dense gradient from dfp_render_bwd */
    float* exp_avg;            /* [n_vox, C] first moment  (zero-initialised by the caller) */
    float* exp_avg_sq;         /* [n_vox, C] second moment */
    const uint8_t* voxel_mask; /* [n_vox] 0/1, or NULL */
    int64_t n_vox;
    int32_t C;                 /* multiple of 4 */
    int32_t step;              /* 1-based Adam step of this parameter */
    double lr, beta1, beta2, eps;
} DfpAdam;
int dfp_masked_adam_step(const DfpAdam* a, void* stream);

/* (f3) Frustum feature selection, Mapper.get_mask_from_c2w (src/Mapper.py:90-158): project every grid
 * vertex into the camera (w2c in f32 as numpy computes it, K in f64), bilinear depth look-up with
 * cv2.remap's INTER_LINEAR arithmetic (1/32-pixel fixed-point coordinates, constant-0 border), fill
 * zero depths with the maximum, depth test (-z <= depth + 0.5), or within 0.5 m of the camera centre.
 * mask is written in the grid's voxel order [D,H,W] (what Mapper.py:347-348 permutes to). */
typedef struct {
    const float* depth;        /* [H, W] f32 depth image (device) */
    int32_t H, W;
    double fx, fy, cx, cy;
    float w2c[16];             /* row-major inverse pose, f32 (np.linalg.inv of the f32 c2w, computed by the caller) */
    float cam_o[3];            /* c2w[:3, 3] */
    const float* xs;           /* [Wg] torch.linspace(bound[0][0], bound[0][1], Wg) (device, f32) */
    const float* ys;           /* [Hg] */
    const float* zs;           /* [Dg] */
    int32_t Dg, Hg, Wg;        /* grid shape val.shape[2:] = (D, H, W) */
} DfpFrustum;
/* scratch: n_vox f32 (sampled depths) + n_vox f64 (-z) + n_vox u8 (in-image flag) + 4 bytes (max): use
 * dfp_frustum_scratch_bytes */
size_t dfp_frustum_scratch_bytes(int32_t Dg, int32_t Hg, int32_t Wg);
int dfp_frustum_mask(const DfpFrustum* f, uint8_t* mask, void* scratch, size_t scratch_bytes, void* stream);

/* (f2) Ray batch assembly + loss, src/Mapper.py:414-449 (get_samples per frame, bbox pre-filter) and
 * :457-469 (loss); src/common.py:76-136.  The random pixel pick stays torch.randint on the caller's side
 * (same generator stream as the reference); everything downstream of the indices is fused:
 * dfp_assemble_rays gathers depth/colour, builds rays (get_rays_from_uv, common.py:76-91), applies the
 * bounding-box filter and compacts the survivors in order (what the boolean indexing of Mapper.py:446-449
 * yields); n_kept is written to device memory. */
typedef struct {
    const float* depth;        /* [H, W] */
    const float* color;        /* [H, W, 3] */
    float c2w[12];             /* rows 0..2 of the pose, f32 */
} DfpFrame;
typedef struct {
    const DfpFrame* frames;    /* HOST array [n_frames] (copied into the launch arguments; at most 16 per call) */
    int32_t n_frames;
    const int64_t* indices;    /* [n_frames, n_per_frame] flat pixel indices into the H0..H1 x W0..W1 window (device) */
    int32_t n_per_frame;
    int32_t H0, H1, W0, W1;    /* sampling window (Mapper: whole image; Tracker: edge-cropped) */
    int32_t W;                 /* image width (row stride of depth / colour) */
    double fx, fy, cx, cy;
    double bound[3][2];
    int32_t filter;            /* 1: drop rays whose depth lies beyond the scene box (Mapper.py:439-449) */
} DfpRayBatch;
typedef struct {
    float* rays_o;             /* [n, 3] capacity n_frames * n_per_frame */
    float* rays_d;             /* [n, 3] */
    float* gt_depth;           /* [n] */
    float* gt_color;           /* [n, 3] */
    int32_t* n_kept;           /* [1] device */
} DfpRayBatchOut;
size_t dfp_assemble_scratch_bytes(int64_t n_total);
int dfp_assemble_rays(const DfpRayBatch* b, const DfpRayBatchOut* out, void* scratch, size_t scratch_bytes, void* stream);

/* Mapper loss (Mapper.py:457-469) and its gradient in one pass: loss = sum_{gt>0} |gt - depth|
 * (+ sum |weight - 1| when w_weight != 0) (+ w_color * sum |gt_color - color| when color != NULL);
 * writes d(loss)/d(depth) f64 [n], d/d(color) f32 [n,3], d/d(weight) f32 [n*S] (each optional) and
 * accumulates the loss (f64) into *loss, which the caller zeroes. */
typedef struct {
    const double* depth; const float* color; const float* weight;      /* renderer outputs */
    const float* gt_depth; const float* gt_color;
    int64_t n; int32_t S;
    double w_color;            /* 0 = no colour term */
    int32_t use_weight;        /* 1 = add sum |weight - 1| */
    double* g_depth; float* g_color; float* g_weight;
    double* loss;
} DfpMapLoss;
int dfp_mapper_loss(const DfpMapLoss* l, void* stream);

/* (f4) Offline TSDF fusion, the `integrate` kernel of src/fusion.py:69-142 (pycuda upstream): one RGB-D frame into the
 * [X][Y][Z] f32 volumes (tsdf initialised to -1, weight and colour to 0 by the caller, fusion.py:52-56), truncation
 * 5 voxels, running averages weighted by obs_weight, colour packed b*65536 + g*256 + r.  Same f32 arithmetic as the
 * upstream kernel (including its float voxel-index decomposition), without fused multiply-adds. */
typedef struct {
    float* tsdf_vol; float* weight_vol; float* color_vol;
    int32_t dim[3];            /* X, Y, Z */
    float origin[3];           /* vol_bnds[:, 0] as f32 (fusion.py:44) */
    float voxel_size, trunc_margin;
    const float* depth_im;     /* [H, W] f32 (device) */
    const float* color_im;     /* [H, W] f32, folded floor(b*65536 + g*256 + r) (fusion.py:207-208) */
    int32_t H, W;
    float intr[9];             /* row-major 3x3 */
    float pose[16];            /* row-major 4x4 camera-to-world */
    float obs_weight;
} DfpTsdfFuse;
int dfp_tsdf_integrate(const DfpTsdfFuse* f, void* stream);

/* (f4) Mesher.point_masks without depth test (src/utils/Mesher.py:58-204, cfg meshing.depth_test = False): points [n,3]
 * f32 are classified against n_frames keyframes given as world-to-camera matrices (np.linalg.inv of the estimated pose,
 * f32, row-major [n_frames][16]) and max_depth[f] = max(depth_f) * 1.1; seen = inside some frame's image with -z <
 * max_depth, forecast = inside some frame's image extended by 1000 px and not seen.  unseen = neither. */
int dfp_point_masks(const float* points, int64_t n, const float* w2c, const float* max_depth, int32_t n_frames, double fx, double fy,
                    double cx, double cy, int32_t H, int32_t W, uint8_t* seen, uint8_t* forecast, void* stream);

/* ---- diagnostics (used by bench.py; not part of the reference seam) ---------------------
 * dfp_launch_count: kernels launched by this library since load.  dfp_timing_enable(1) makes
 * every launch be bracketed by CUDA events on its own stream; dfp_timing_collect waits for them
 * and writes "<kernel> <launches> <total_ms>\n" per kernel name into buf. */
int64_t dfp_launch_count(void);
int dfp_timing_enable(int on);
int dfp_timing_collect(char* buf, size_t buf_bytes);

#ifdef __cplusplus
}
#endif
#endif /* DFPRIOR_B200_H */
