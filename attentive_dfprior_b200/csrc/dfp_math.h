// Per-thread arithmetic of the hot path: one sample (or one ray) per thread.
//
// Everything here is plain sequential code on thread-private data, so the same
// functions run inside the sm_100a kernels (dfp_kernels.cu) and inside the
// host emulation that the CPU unit tests build with g++ (tests/hostemu).  The
// cooperative parts (weight-gradient reduction, band compaction, scatter) are
// supplied by the caller through the Sink template parameter.
//
// Reference lines restated by each function are cited beside it (paths relative
// to the reference checkout).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "dfp_layout.h"

namespace dfp {

#if !defined(__CUDACC__)
struct float4 { float x, y, z, w; };
#endif

// ---- correctly rounded, never-fused scalar ops (the f64/f32 sampling chain must be bit-exact,
//      SURVEY.md appendix A.2; host builds use -ffp-contract=off)
#if defined(__CUDA_ARCH__)
DFP_HD float op_mul(float a, float b) { return __fmul_rn(a, b); }
DFP_HD float op_add(float a, float b) { return __fadd_rn(a, b); }
DFP_HD float op_sub(float a, float b) { return __fsub_rn(a, b); }
DFP_HD float op_div(float a, float b) { return __fdiv_rn(a, b); }
DFP_HD double op_mul(double a, double b) { return __dmul_rn(a, b); }
DFP_HD double op_add(double a, double b) { return __dadd_rn(a, b); }
DFP_HD double op_sub(double a, double b) { return __dsub_rn(a, b); }
DFP_HD double op_div(double a, double b) { return __ddiv_rn(a, b); }
#else
DFP_HD float op_mul(float a, float b) { return a * b; }
DFP_HD float op_add(float a, float b) { return a + b; }
DFP_HD float op_sub(float a, float b) { return a - b; }
DFP_HD float op_div(float a, float b) { return a / b; }
DFP_HD double op_mul(double a, double b) { return a * b; }
DFP_HD double op_add(double a, double b) { return a + b; }
DFP_HD double op_sub(double a, double b) { return a - b; }
DFP_HD double op_div(double a, double b) { return a / b; }
#endif

// =====================================================================================
// ray sampling  (src/utils/Renderer.py:134-225)
// =====================================================================================
struct SampleTables {
    int n_samples, n_surface;
    float tv[64];        // linspace(0,1,N_samples), f32                         Renderer.py:203
    float omt[64];       // 1 - tv, f32
    double ts[64];       // linspace(0,1,N_surface) f32 -> f64                   Renderer.py:183-184
    double omts[64];     // 1 - ts, f64
};

// far bound of one ray: min over axes of the larger slab distance, +0.01, clamped to
// [0, max(gt_depth*1.2)] when depth is given                                   Renderer.py:148-162
DFP_HD double ray_far(const float o[3], const float d[3], const double bound[3][2], bool has_depth, float max12) {
    double far = 0.0;
    for (int k = 0; k < 3; ++k) {
        const double t0 = op_div(op_sub(bound[k][0], (double)o[k]), (double)d[k]);
        const double t1 = op_div(op_sub(bound[k][1], (double)o[k]), (double)d[k]);
        const double m = t0 > t1 ? t0 : t1;
        far = (k == 0 || m < far) ? m : far;
    }
    far = op_add(far, 0.01);
    if (has_depth) {
        far = far < 0.0 ? 0.0 : far;
        const double hi = (double)max12;
        far = far > hi ? hi : far;
    }
    return far;
}

// candidate depth i of a ray, before sorting.  i < n_samples: uniform sample between
// near = 0.01*depth (f32) and far (f64), Renderer.py:146,206; otherwise near-surface sample
// (depth > 0, :186-188) or camera-to-max-depth sample (depth == 0, :194-201).
DFP_HD double z_candidate(int i, bool has_depth, float gt, double far, float max_d, const SampleTables& T) {
    if (i < T.n_samples) {
        const float near = has_depth ? op_mul(gt, 0.01f) : 0.01f;
        const float a = op_mul(near, T.omt[i]);
        return op_add((double)a, op_mul(far, (double)T.tv[i]));
    }
    const int k = i - T.n_samples;
    if (gt > 0.f) {
        const float lo = op_mul(0.95f, gt), hi = op_mul(1.05f, gt);
        return op_add(op_mul((double)lo, T.omts[k]), op_mul((double)hi, T.ts[k]));
    }
    return op_add(op_mul(0.001, T.omts[k]), op_mul((double)max_d, T.ts[k]));
}

// pts = rays_o + rays_d * z in f64, separate multiply and add            Renderer.py:223-224
DFP_HD void ray_point(const float o[3], const float d[3], double z, double p[3]) {
    for (int k = 0; k < 3; ++k) p[k] = op_add((double)o[k], op_mul((double)d[k], z));
}

// =====================================================================================
// coordinate normalisation + trilinear set-up
// =====================================================================================
template <typename PT>
struct BoxT {                // bound scalars rounded into the points' dtype
    PT lo[3], hi[3], len[3];
};
template <typename PT>
DFP_HD BoxT<PT> make_box(const double b[3][2]) {
    BoxT<PT> r;
    for (int k = 0; k < 3; ++k) {
        r.lo[k] = (PT)b[k][0];
        r.hi[k] = (PT)b[k][1];
        r.len[k] = (PT)(b[k][1] - b[k][0]);
    }
    return r;
}

// strict in-bound test in the points' dtype                               Renderer.py:51-54
template <typename PT>
DFP_HD bool in_bound(const PT p[3], const BoxT<PT>& b) {
    bool m = true;
    for (int k = 0; k < 3; ++k) m = m && (p[k] < b.hi[k]) && (p[k] > b.lo[k]);
    return m;
}

// ((p-b0)/(b1-b0))*2-1 in the points' dtype, then .float()                common.py:287-289, decoder.py:171
template <typename PT>
DFP_HD void normalize(const PT p[3], const BoxT<PT>& b, float xn[3]) {
    for (int k = 0; k < 3; ++k)
        xn[k] = (float)op_sub(op_mul(op_div(op_sub(p[k], b.lo[k]), b.len[k]), (PT)2), (PT)1);
}

// grid_sample(align_corners=True, padding_mode='border', mode='bilinear') set-up for one point
// (ATen GridSampler.cuh:23-31 unnormalise, :55-80 clip and its gradient)
struct Tri {
    int x0, y0, z0;            // floor corner
    float ix, iy, iz;          // clipped continuous index
    float gx, gy, gz;          // d(i*)/d(normalised coord): (size-1)/2, or 0 where clipped
};
DFP_HD void tri_axis(float xn, int size, float& ic, int& i0, float& g) {
    float v = ((xn + 1.f) / 2) * (float)(size - 1);
    const float mx = (float)(size - 1);
    g = mx / 2;
    if (v <= 0.f) { v = 0.f; g = 0.f; }
    else if (v >= mx) { v = mx; g = 0.f; }
    ic = v;
    i0 = (int)floorf(v);
}
DFP_HD Tri tri_setup(const float xn[3], int W, int H, int D) {
    Tri t;
    tri_axis(xn[0], W, t.ix, t.x0, t.gx);
    tri_axis(xn[1], H, t.iy, t.y0, t.gy);
    tri_axis(xn[2], D, t.iz, t.z0, t.gz);
    return t;
}
// weight of corner c (bit0: +x, bit1: +y, bit2: +z) in ATen's order tnw,tne,tsw,tse,bnw,bne,bsw,bse
DFP_HD float tri_weight(const Tri& t, int c) {
    const float x1 = (float)t.x0 + 1.f, y1 = (float)t.y0 + 1.f, z1 = (float)t.z0 + 1.f;
    const float wx = (c & 1) ? (t.ix - (float)t.x0) : (x1 - t.ix);
    const float wy = (c & 2) ? (t.iy - (float)t.y0) : (y1 - t.iy);
    const float wz = (c & 4) ? (t.iz - (float)t.z0) : (z1 - t.iz);
    return wx * wy * wz;
}
DFP_HD bool tri_corner(const Tri& t, int c, int W, int H, int D, int& x, int& y, int& z) {
    x = t.x0 + (c & 1);
    y = t.y0 + ((c >> 1) & 1);
    z = t.z0 + ((c >> 2) & 1);
    return x < W && y < H && z < D;      // x0,y0,z0 >= 0 after the clip
}
// d(weight of corner c)/d(ix, iy, iz)
DFP_HD void tri_weight_grad(const Tri& t, int c, float& dwx, float& dwy, float& dwz) {
    const float x1 = (float)t.x0 + 1.f, y1 = (float)t.y0 + 1.f, z1 = (float)t.z0 + 1.f;
    const float wx = (c & 1) ? (t.ix - (float)t.x0) : (x1 - t.ix);
    const float wy = (c & 2) ? (t.iy - (float)t.y0) : (y1 - t.iy);
    const float wz = (c & 4) ? (t.iz - (float)t.z0) : (z1 - t.iz);
    const float sx = (c & 1) ? 1.f : -1.f, sy = (c & 2) ? 1.f : -1.f, sz = (c & 4) ? 1.f : -1.f;
    dwx = sx * wy * wz;
    dwy = wx * sy * wz;
    dwz = wx * wy * sz;
}

// scalar (C = 1) trilinear value with arbitrary strides: the TSDF prior    decoder.py:295-303
DFP_HD float tsdf_interp(const float* vol, int64_t sD, int64_t sH, int64_t sW, int W, int H, int D, const Tri& t) {
    float acc = 0.f;
    for (int c = 0; c < 8; ++c) {
        int x, y, z;
        if (tri_corner(t, c, W, H, D, x, y, z)) acc += vol[z * sD + y * sH + x * sW] * tri_weight(t, c);
    }
    return acc;
}
// d(value)/d(normalised coords) of the same                               (ATen grid_sampler_3d backward)
DFP_HD void tsdf_interp_grad(const float* vol, int64_t sD, int64_t sH, int64_t sW, int W, int H, int D, const Tri& t,
                             float g[3]) {
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int c = 0; c < 8; ++c) {
        int x, y, z;
        if (tri_corner(t, c, W, H, D, x, y, z)) {
            float dx, dy, dz;
            tri_weight_grad(t, c, dx, dy, dz);
            const float v = vol[z * sD + y * sH + x * sW];
            gx += v * dx; gy += v * dy; gz += v * dz;
        }
    }
    g[0] = gx * t.gx; g[1] = gy * t.gy; g[2] = gz * t.gz;
}

// TSDF truncation band                                                    decoder.py:329, :346
DFP_HD bool in_band(float tsdf) { return tsdf > (float)(-1.0 + 1e-4) && tsdf < (float)(1.0 - 1e-4); }

// =====================================================================================
// sin / cos of the Fourier-embedding arguments                              decoder.py:30
// =====================================================================================
// Branch-free sin/cos: round(x*2/pi) by the magic-number trick, three-term Cody-Waite reduction by pi/2,
// minimax polynomials on [-pi/4, pi/4].  Max abs error 7.1e-8 for |x| <= 3000 (measured against f64 sin/cos,
// i.e. 1 ulp-class like libm's sinf) and good to |x| < 4.8e4; the callers fall back to libm's sinf above
// that bound (never reached with _B ~ N(0, 25^2) on metre-scale coordinates, where |arg| stays below ~1000).
// libm's own sinf costs a divergent slow-path check per call, which dominated the tensor-core kernels.
#if defined(__CUDA_ARCH__)
DFP_HD int f32_bits(float f) { return __float_as_int(f); }
#else
DFP_HD int f32_bits(float f) { int u; memcpy(&u, &f, 4); return u; }
#endif
constexpr float SINCOS_CW_MAX_ARG = 4.8e4f;
DFP_HD void sincos_cw(float x, float& sn, float& cs) {
    const float t = fmaf(x, 0.636619747f, 12582912.f);
    const int q = f32_bits(t);
    const float j = op_sub(t, 12582912.f);
    float r = fmaf(j, -1.57079601e+00f, x);
    r = fmaf(j, -3.13916473e-07f, r);
    r = fmaf(j, -5.39030253e-15f, r);
    const float s = r * r;
    float ps = fmaf(2.86567956e-6f, s, -1.98559923e-4f);
    ps = fmaf(ps, s, 8.33338592e-3f);
    ps = fmaf(ps, s, -1.66666672e-1f);
    const float sr = fmaf(r * s, ps, r);
    float pc = fmaf(2.44677067e-5f, s, -1.38877297e-3f);
    pc = fmaf(pc, s, 4.16666567e-2f);
    pc = fmaf(pc, s, -0.5f);
    const float cr = fmaf(pc, s, 1.0f);
    const float a = (q & 1) ? cr : sr;
    const float b = (q & 1) ? sr : cr;
    sn = (q & 2) ? -a : a;
    cs = ((q + 1) & 2) ? -b : b;
}
DFP_HD float sin_cw(float x) {
    float sn, cs;
    sincos_cw(x, sn, cs);
    return sn;
}
// largest |_B| entry of one decoder's embedding matrix ([3][EMB_PAD] rows, EMB used)
DFP_HD float embed_bmax(const float* Bm) {
    float m = 0.f;
    for (int r = 0; r < 3; ++r)
        for (int k = 0; k < EMB; ++k) m = fmaxf(m, fabsf(Bm[r * EMB_PAD + k]));
    return m;
}
// true when every embedding argument of this point is inside sincos_cw's range
DFP_HD bool embed_fast_ok(const float pf[3], float bmax) {
    return (fabsf(pf[0]) + fabsf(pf[1]) + fabsf(pf[2])) * bmax < SINCOS_CW_MAX_ARG;
}

// =====================================================================================
// grid-conditioned decoder, one sample per thread                          decoder.py:177-203
// =====================================================================================
struct DecState {                 // recomputed-forward state the backward needs
    float hist[5 * HID];          // hist[(i-1)*32..] = input of pts_linears.i (i=1..4); hist[128..] = input of output_linear
    unsigned relu[5];             // bit j: pre-activation of pts_linears.i neuron j is > 0
};

template <int CD>
DFP_HD void add_fc_c(const float* Wc, const float* bc, const float* c, float (&h)[HID]) {
#pragma unroll
    for (int j = 0; j < HID; ++j) h[j] += bc[j];
#pragma unroll 4
    for (int k = 0; k < CD; ++k) {
        const float ck = c[k];
        const float4* w4 = reinterpret_cast<const float4*>(Wc + k * HID);
#pragma unroll
        for (int q = 0; q < HID / 4; ++q) {
            const float4 w = w4[q];
            h[4 * q + 0] = fmaf(w.x, ck, h[4 * q + 0]);
            h[4 * q + 1] = fmaf(w.y, ck, h[4 * q + 1]);
            h[4 * q + 2] = fmaf(w.z, ck, h[4 * q + 2]);
            h[4 * q + 3] = fmaf(w.w, ck, h[4 * q + 3]);
        }
    }
}

DFP_HD float embed_arg(const float* Bm, const float pf[3], int k) {          // (p @ _B)[k]      decoder.py:29
    return fmaf(pf[2], Bm[2 * EMB_PAD + k], fmaf(pf[1], Bm[EMB_PAD + k], pf[0] * Bm[k]));
}

// 16 consecutive arguments k0 .. k0+15 (k0 a multiple of 4, Bm 16-byte aligned): the three rows of _B come in as
// 128-bit loads (a quarter of the shared-memory instructions of 16 embed_arg calls); same operation order per element
DFP_HD void embed_arg16(const float* Bm, const float pf[3], int k0, float arg[16]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b0 = *reinterpret_cast<const float4*>(Bm + k0 + 4 * q);
        const float4 b1 = *reinterpret_cast<const float4*>(Bm + EMB_PAD + k0 + 4 * q);
        const float4 b2 = *reinterpret_cast<const float4*>(Bm + 2 * EMB_PAD + k0 + 4 * q);
        arg[4 * q + 0] = fmaf(pf[2], b2.x, fmaf(pf[1], b1.x, pf[0] * b0.x));
        arg[4 * q + 1] = fmaf(pf[2], b2.y, fmaf(pf[1], b1.y, pf[0] * b0.y));
        arg[4 * q + 2] = fmaf(pf[2], b2.z, fmaf(pf[1], b1.z, pf[0] * b0.z));
        arg[4 * q + 3] = fmaf(pf[2], b2.w, fmaf(pf[1], b1.w, pf[0] * b0.w));
    }
#else
    for (int i = 0; i < 16; ++i) arg[i] = embed_arg(Bm, pf, k0 + i);
#endif
}

// W: this decoder's packed parameters (shared memory on the device), pf: world point (f32),
// c: gathered features [CD].  out: 1 (occupancy) or 4 (rgb, occ) values.
template <int CD, int NOUT, bool SAVE>
DFP_HD void decoder_fwd_thread(const float* W, const float pf[3], const float* c, float (&out)[NOUT], DecState* st,
                               float bmax) {
    using L = DecOff<CD>;
    const bool fast = embed_fast_ok(pf, bmax);
    float a0[HID], a3[HID];
#pragma unroll
    for (int j = 0; j < HID; ++j) { a0[j] = W[L::BP + j]; a3[j] = W[L::BP + 3 * HID + j]; }
    // sin(p @ B) feeds pts_linears.0 and, through the skip connection, the first 93 inputs of
    // pts_linears.3 (decoder.py:198-199): both are accumulated while the embedding is produced
#pragma unroll 3
    for (int k = 0; k < EMB; ++k) {
        const float arg = embed_arg(W + L::B, pf, k);
        const float e = fast ? sin_cw(arg) : sinf(arg);
        const float4* w4 = reinterpret_cast<const float4*>(W + L::W03 + k * 64);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 w = w4[q];
            a0[4 * q + 0] = fmaf(w.x, e, a0[4 * q + 0]);
            a0[4 * q + 1] = fmaf(w.y, e, a0[4 * q + 1]);
            a0[4 * q + 2] = fmaf(w.z, e, a0[4 * q + 2]);
            a0[4 * q + 3] = fmaf(w.w, e, a0[4 * q + 3]);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 w = w4[8 + q];
            a3[4 * q + 0] = fmaf(w.x, e, a3[4 * q + 0]);
            a3[4 * q + 1] = fmaf(w.y, e, a3[4 * q + 1]);
            a3[4 * q + 2] = fmaf(w.z, e, a3[4 * q + 2]);
            a3[4 * q + 3] = fmaf(w.w, e, a3[4 * q + 3]);
        }
    }
    float h[HID];
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < HID; ++j) { if (a0[j] > 0.f) m |= 1u << j; h[j] = fmaxf(a0[j], 0.f); }
    add_fc_c<CD>(W + L::WC, W + L::BC, c, h);                               // h = relu(.) + fc_c[0](c)   decoder.py:194-197
    if (SAVE) {
        st->relu[0] = m;
#pragma unroll
        for (int j = 0; j < HID; ++j) st->hist[j] = h[j];
    }
#pragma unroll 1
    for (int i = 1; i < 5; ++i) {
        const float* Wi = W + L::WH + (i - 1) * HID * HID;
        float acc[HID];
#pragma unroll
        for (int j = 0; j < HID; ++j) acc[j] = (i == 3) ? a3[j] : W[L::BP + i * HID + j];
#pragma unroll
        for (int k = 0; k < HID; ++k) {
            const float hk = h[k];
            const float4* w4 = reinterpret_cast<const float4*>(Wi + k * HID);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 w = w4[q];
                acc[4 * q + 0] = fmaf(w.x, hk, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(w.y, hk, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(w.z, hk, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(w.w, hk, acc[4 * q + 3]);
            }
        }
        m = 0;
#pragma unroll
        for (int j = 0; j < HID; ++j) { if (acc[j] > 0.f) m |= 1u << j; h[j] = fmaxf(acc[j], 0.f); }
        add_fc_c<CD>(W + L::WC + i * CD * HID, W + L::BC + i * HID, c, h);
        if (SAVE) {
            st->relu[i] = m;
#pragma unroll
            for (int j = 0; j < HID; ++j) st->hist[i * HID + j] = h[j];
        }
    }
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
        float s = W[L::BO + o];
#pragma unroll
        for (int j = 0; j < HID; ++j) s = fmaf(W[L::WO + o * HID + j], h[j], s);
        out[o] = s;
    }
}

// Backward of one decoder for one sample.  dout: gradient of the decoder outputs.
// Produces d_c[CD] (gradient of the gathered features) and adds the embedding path's
// gradient w.r.t. the world point into dp[3].  Parameter gradients go to the Sink:
//   sink.outer32(woff, ld, a, K, d, boff): g[woff + k*ld + j] += a[k]*d[j] (k<K, j<32), g[boff + j] += d[j]
//   sink.outer4 (woff, ld, a, K, d):       the same for K <= 4
//   sink.scalar(off, v):                   g[off] += v
// all summed over the samples the sink spans.  WANT_PARAMS = false skips them (Tracker).
template <int CD, int NOUT, bool WANT_PARAMS, class Sink>
DFP_HD void decoder_bwd_thread(const float* W, int gbase, const float pf[3], const float* c, const DecState& st,
                               const float (&dout)[NOUT], float* d_c, float dp[3], Sink& sink, float bmax) {
    using L = DecOff<CD>;
    const bool fast = embed_fast_ok(pf, bmax);
    float dh[HID];
    // output_linear
#pragma unroll
    for (int j = 0; j < HID; ++j) {
        float s = 0.f;
#pragma unroll
        for (int o = 0; o < NOUT; ++o) s = fmaf(W[L::WO + o * HID + j], dout[o], s);
        dh[j] = s;
    }
    if (WANT_PARAMS) {
        float h5[HID], dd[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < HID; ++j) h5[j] = st.hist[4 * HID + j];
#pragma unroll
        for (int o = 0; o < NOUT; ++o) dd[o] = dout[o];
        sink.outer4(gbase + L::WO, HID, dd, NOUT, h5);
#pragma unroll
        for (int o = 0; o < NOUT; ++o) sink.scalar(gbase + L::BO + o, dout[o]);
    }
    for (int k = 0; k < CD; ++k) d_c[k] = 0.f;
    float dacc3[HID];
#pragma unroll
    for (int j = 0; j < HID; ++j) dacc3[j] = 0.f;

#pragma unroll 1
    for (int i = 4; i >= 1; --i) {
        const float* Wc = W + L::WC + i * CD * HID;
        const float* Wi = W + L::WH + (i - 1) * HID * HID;
        if (WANT_PARAMS) sink.outer32(gbase + L::WC + i * CD * HID, HID, c, CD, dh, gbase + L::BC + i * HID);
        // d_c += Wc_i . dh
#pragma unroll 2
        for (int k = 0; k < CD; ++k) {
            const float4* w4 = reinterpret_cast<const float4*>(Wc + k * HID);
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 w = w4[q];
                s0 = fmaf(w.x, dh[4 * q + 0], s0); s1 = fmaf(w.y, dh[4 * q + 1], s1);
                s2 = fmaf(w.z, dh[4 * q + 2], s2); s3 = fmaf(w.w, dh[4 * q + 3], s3);
            }
            d_c[k] += (s0 + s1) + (s2 + s3);
        }
        // through the ReLU of pts_linears.i
        const unsigned m = st.relu[i];
#pragma unroll
        for (int j = 0; j < HID; ++j) dh[j] = ((m >> j) & 1u) ? dh[j] : 0.f;
        if (WANT_PARAMS) sink.outer32(gbase + L::WH + (i - 1) * HID * HID, HID, st.hist + (i - 1) * HID, HID, dh,
                                      gbase + L::BP + i * HID);
        if (i == 3) {
#pragma unroll
            for (int j = 0; j < HID; ++j) dacc3[j] = dh[j];
        }
        // dh_prev[k] = sum_j W_i[k][j] * dacc[j]
        float dprev[HID];
#pragma unroll
        for (int k = 0; k < HID; ++k) {
            const float4* w4 = reinterpret_cast<const float4*>(Wi + k * HID);
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 w = w4[q];
                s0 = fmaf(w.x, dh[4 * q + 0], s0); s1 = fmaf(w.y, dh[4 * q + 1], s1);
                s2 = fmaf(w.z, dh[4 * q + 2], s2); s3 = fmaf(w.w, dh[4 * q + 3], s3);
            }
            dprev[k] = (s0 + s1) + (s2 + s3);
        }
#pragma unroll
        for (int j = 0; j < HID; ++j) dh[j] = dprev[j];
    }
    // layer 0: fc_c[0] and the ReLU of pts_linears.0
    if (WANT_PARAMS) sink.outer32(gbase + L::WC, HID, c, CD, dh, gbase + L::BC);
#pragma unroll 2
    for (int k = 0; k < CD; ++k) {
        const float4* w4 = reinterpret_cast<const float4*>(W + L::WC + k * HID);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 w = w4[q];
            s0 = fmaf(w.x, dh[4 * q + 0], s0); s1 = fmaf(w.y, dh[4 * q + 1], s1);
            s2 = fmaf(w.z, dh[4 * q + 2], s2); s3 = fmaf(w.w, dh[4 * q + 3], s3);
        }
        d_c[k] += (s0 + s1) + (s2 + s3);
    }
    {
        const unsigned m = st.relu[0];
#pragma unroll
        for (int j = 0; j < HID; ++j) dh[j] = ((m >> j) & 1u) ? dh[j] : 0.f;
    }
    // embedding: e_k = sin(arg_k) feeds pts_linears.0 (grad dh) and pts_linears.3[:, :93] (grad dacc3)
    float gp0 = 0.f, gp1 = 0.f, gp2 = 0.f;
#pragma unroll 1
    for (int k0 = 0; k0 < EMB; k0 += 32) {
        float eb[32], darg[32];
#pragma unroll 1
        for (int kk = 0; kk < 32; ++kk) {
            const int k = k0 + kk;
            float e = 0.f, da = 0.f;
            if (k < EMB) {
                const float arg = embed_arg(W + L::B, pf, k);
                float sn, cs;
                if (fast) sincos_cw(arg, sn, cs);
                else sincosf(arg, &sn, &cs);
                const float4* w4 = reinterpret_cast<const float4*>(W + L::W03 + k * 64);
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 w = w4[q];
                    s0 = fmaf(w.x, dh[4 * q + 0], s0); s1 = fmaf(w.y, dh[4 * q + 1], s1);
                    s2 = fmaf(w.z, dh[4 * q + 2], s2); s3 = fmaf(w.w, dh[4 * q + 3], s3);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 w = w4[8 + q];
                    s0 = fmaf(w.x, dacc3[4 * q + 0], s0); s1 = fmaf(w.y, dacc3[4 * q + 1], s1);
                    s2 = fmaf(w.z, dacc3[4 * q + 2], s2); s3 = fmaf(w.w, dacc3[4 * q + 3], s3);
                }
                e = sn;
                da = ((s0 + s1) + (s2 + s3)) * cs;
                gp0 = fmaf(W[L::B + k], da, gp0);
                gp1 = fmaf(W[L::B + EMB_PAD + k], da, gp1);
                gp2 = fmaf(W[L::B + 2 * EMB_PAD + k], da, gp2);
            }
            eb[kk] = e;
            darg[kk] = da;
        }
        if (WANT_PARAMS) {
            const int kb = (EMB - k0) < 32 ? (EMB - k0) : 32;
            sink.outer32(gbase + L::W03 + k0 * 64, 64, eb, kb, dh, k0 == 0 ? gbase + L::BP : -1);
            sink.outer32(gbase + L::W03 + k0 * 64 + 32, 64, eb, kb, dacc3, -1);
            float pp[4] = {pf[0], pf[1], pf[2], 0.f};
            sink.outer4(gbase + L::B + k0, EMB_PAD, pp, 3, darg);                // d _B[r][k0+j] += p[r]*darg[j]
        }
    }
    dp[0] += gp0; dp[1] += gp1; dp[2] += gp2;
}

// =====================================================================================
// attentive fusion of learned occupancy and the TSDF prior                 decoder.py:240-258
// =====================================================================================
struct AttLogit {            // tsdf -> occupancy logit with what its derivative needs
    float inv;               // clamp(-0.1*log(1/(t+1e-8) - 1 + 1e-7), -100, 100)
    float dinv_dtsdf;        // d inv / d tsdf (0 where a clamp is active)
};
DFP_HD AttLogit tsdf_logit(float tsdf) {                                        // decoder.py:244-248
    AttLogit r;
    const float t_raw = op_sub(1.f, op_div(op_add(tsdf, 1.f), 2.f));
    const float t = fminf(fmaxf(t_raw, 0.f), 1.f);
    const float u = op_add(t, 1e-8f);
    const float q = op_add(op_sub(op_div(1.f, u), 1.f), 1e-7f);
    const float raw = op_mul(-0.1f, logf(q));
    r.inv = fminf(fmaxf(raw, -100.f), 100.f);
    const bool pass_inv = (raw >= -100.f) && (raw <= 100.f);
    const bool pass_t = (t_raw >= 0.f) && (t_raw <= 1.f);
    // d raw/dq = -0.1/q ; dq/du = -1/u^2 ; dt_raw/dtsdf = -1/2
    r.dinv_dtsdf = (pass_inv && pass_t) ? (0.1f / (q * u * u)) * (-0.5f) : 0.f;
    return r;
}

struct AttState {            // activations of the attention MLP (post-ReLU), kept for the backward
    float h1[64], h2[128], h3[128], h4[64];
    float a0, a1;
};

// acc[32] = bias[jc*32..] + sum_k W[k][jc*32 + j] * in[k]   (k-major W with row length NJ)
template <int K>
DFP_HD void att_dense_chunk(const float* Wm, const float* bias, int NJ, int jc, const float* in, float (&acc)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = bias[jc * 32 + j];
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float v = in[k];
        const float4* w4 = reinterpret_cast<const float4*>(Wm + k * NJ + jc * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 w = w4[q];
            acc[4 * q + 0] = fmaf(w.x, v, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(w.y, v, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(w.z, v, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(w.w, v, acc[4 * q + 3]);
        }
    }
}

// x0 = learned occupancy (high + low), x1 = TSDF logit.  Returns fused occupancy; w = weight on the TSDF term.
DFP_HD float attention_fwd_thread(const float* A, float x0, float x1, float& w_out, AttState& st) {
    const float xin[2] = {x0, x1};
    float acc[32];
    for (int jc = 0; jc < 2; ++jc) {
        att_dense_chunk<2>(A + AttOff::W1, A + AttOff::B1, 64, jc, xin, acc);
#pragma unroll
        for (int j = 0; j < 32; ++j) st.h1[jc * 32 + j] = fmaxf(acc[j], 0.f);
    }
    for (int jc = 0; jc < 4; ++jc) {
        att_dense_chunk<64>(A + AttOff::W2, A + AttOff::B2, 128, jc, st.h1, acc);
#pragma unroll
        for (int j = 0; j < 32; ++j) st.h2[jc * 32 + j] = fmaxf(acc[j], 0.f);
    }
    for (int jc = 0; jc < 4; ++jc) {
        att_dense_chunk<128>(A + AttOff::W3, A + AttOff::B3, 128, jc, st.h2, acc);
#pragma unroll
        for (int j = 0; j < 32; ++j) st.h3[jc * 32 + j] = fmaxf(acc[j], 0.f);
    }
    for (int jc = 0; jc < 2; ++jc) {
        att_dense_chunk<128>(A + AttOff::W4, A + AttOff::B4, 64, jc, st.h3, acc);
#pragma unroll
        for (int j = 0; j < 32; ++j) st.h4[jc * 32 + j] = fmaxf(acc[j], 0.f);
    }
    float l0 = A[AttOff::B5], l1 = A[AttOff::B5 + 1];
#pragma unroll 8
    for (int j = 0; j < 64; ++j) {
        l0 = fmaf(A[AttOff::W5 + j], st.h4[j], l0);
        l1 = fmaf(A[AttOff::W5 + 64 + j], st.h4[j], l1);
    }
    const float mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
    const float inv = 1.f / (e0 + e1);
    st.a0 = e0 * inv;
    st.a1 = e1 * inv;
    w_out = st.a1;
    return st.a0 * x0 + st.a1 * x1;
}

// dprev[k] (+)= sum_j W[k][jc*32 + j] * d[j] for k < K
template <int K>
DFP_HD void att_back_chunk(const float* Wm, int NJ, int jc, const float (&d)[32], float* dprev, bool first) {
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        const float4* w4 = reinterpret_cast<const float4*>(Wm + k * NJ + jc * 32);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 w = w4[q];
            s0 = fmaf(w.x, d[4 * q + 0], s0); s1 = fmaf(w.y, d[4 * q + 1], s1);
            s2 = fmaf(w.z, d[4 * q + 2], s2); s3 = fmaf(w.w, d[4 * q + 3], s3);
        }
        const float s = (s0 + s1) + (s2 + s3);
        dprev[k] = first ? s : dprev[k] + s;
    }
}

// Backward of the fusion for one sample.  g_out: grad of fused occupancy, g_w: grad of the returned
// weight a1.  Returns d x0, d x1.
template <bool WANT_PARAMS, class Sink>
DFP_HD void attention_bwd_thread(const float* A, int gbase, float x0, float x1, const AttState& st, float g_out,
                                 float g_w, float& dx0, float& dx1, Sink& sink) {
    const float da0 = g_out * x0, da1 = g_out * x1 + g_w;
    const float dot = st.a0 * da0 + st.a1 * da1;
    const float dl[2] = {st.a0 * (da0 - dot), st.a1 * (da1 - dot)};             // softmax backward
    dx0 = g_out * st.a0;
    dx1 = g_out * st.a1;
    float bufA[128], bufB[128];
    float d[32];
    // output_linear 64 -> 2
    for (int jc = 0; jc < 2; ++jc) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int jj = jc * 32 + j;
            const float g = A[AttOff::W5 + jj] * dl[0] + A[AttOff::W5 + 64 + jj] * dl[1];
            bufA[jj] = st.h4[jj] > 0.f ? g : 0.f;                              // through ReLU of pts_linears.3
        }
        if (WANT_PARAMS) {
            float hh[32], dd[4] = {dl[0], dl[1], 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 32; ++j) hh[j] = st.h4[jc * 32 + j];
            sink.outer4(gbase + AttOff::W5 + jc * 32, 64, dd, 2, hh);
        }
    }
    if (WANT_PARAMS) { sink.scalar(gbase + AttOff::B5, dl[0]); sink.scalar(gbase + AttOff::B5 + 1, dl[1]); }
    // pts_linears.3: 128 -> 64   (delta in bufA[64], produces bufB[128] masked by h3)
    for (int jc = 0; jc < 2; ++jc) {
#pragma unroll
        for (int j = 0; j < 32; ++j) d[j] = bufA[jc * 32 + j];
        if (WANT_PARAMS) sink.outer32(gbase + AttOff::W4 + jc * 32, 64, st.h3, 128, d, gbase + AttOff::B4 + jc * 32);
        att_back_chunk<128>(A + AttOff::W4, 64, jc, d, bufB, jc == 0);
    }
    for (int k = 0; k < 128; ++k) bufB[k] = st.h3[k] > 0.f ? bufB[k] : 0.f;
    // pts_linears.2: 128 -> 128  (delta bufB[128] -> bufA[128] masked by h2)
    for (int jc = 0; jc < 4; ++jc) {
#pragma unroll
        for (int j = 0; j < 32; ++j) d[j] = bufB[jc * 32 + j];
        if (WANT_PARAMS) sink.outer32(gbase + AttOff::W3 + jc * 32, 128, st.h2, 128, d, gbase + AttOff::B3 + jc * 32);
        att_back_chunk<128>(A + AttOff::W3, 128, jc, d, bufA, jc == 0);
    }
    for (int k = 0; k < 128; ++k) bufA[k] = st.h2[k] > 0.f ? bufA[k] : 0.f;
    // pts_linears.1: 64 -> 128   (delta bufA[128] -> bufB[64] masked by h1)
    for (int jc = 0; jc < 4; ++jc) {
#pragma unroll
        for (int j = 0; j < 32; ++j) d[j] = bufA[jc * 32 + j];
        if (WANT_PARAMS) sink.outer32(gbase + AttOff::W2 + jc * 32, 128, st.h1, 64, d, gbase + AttOff::B2 + jc * 32);
        att_back_chunk<64>(A + AttOff::W2, 128, jc, d, bufB, jc == 0);
    }
    for (int k = 0; k < 64; ++k) bufB[k] = st.h1[k] > 0.f ? bufB[k] : 0.f;
    // pts_linears.0: 2 -> 64
    float gx[2] = {0.f, 0.f};
    const float xin[4] = {x0, x1, 0.f, 0.f};
    for (int jc = 0; jc < 2; ++jc) {
#pragma unroll
        for (int j = 0; j < 32; ++j) d[j] = bufB[jc * 32 + j];
        if (WANT_PARAMS) {
            sink.outer4(gbase + AttOff::W1 + jc * 32, 64, xin, 2, d);
            sink.bias32(gbase + AttOff::B1 + jc * 32, d);
        }
        att_back_chunk<2>(A + AttOff::W1, 64, jc, d, gx, jc == 0);
    }
    dx0 += gx[0];
    dx1 += gx[1];
}

// =====================================================================================
// compositing                                                              src/common.py:206-251
// =====================================================================================
DFP_HD float sigmoid10(float occ) { return 1.f / (1.f + expf(-(10.f * occ))); }   // common.py:236

// One ray, S samples.  occ[s]: final raw occupancy (100 where out of bound), rgb: [S][3] or nullptr.
// wts_out (optional, [S]) receives the compositing weights.
DFP_HD void composite_ray(int S, const float* occ, const float* rgb, const double* z, double& depth, double& var,
                          float col[3], float* wts_out) {
    float T = 1.f;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
    double dsum = 0.0;
    float wloc[64];
    for (int s = 0; s < S; ++s) {
        const float alpha = sigmoid10(occ[s]);
        const float w = alpha * T;                                            // common.py:244-245
        wloc[s] = w;
        T = T * ((1.f - alpha) + 1e-10f);
        if (rgb) { c0 += w * rgb[3 * s]; c1 += w * rgb[3 * s + 1]; c2 += w * rgb[3 * s + 2]; }
        dsum += (double)w * z[s];                                             // common.py:248 (f64)
    }
    double v = 0.0;
    for (int s = 0; s < S; ++s) {
        const double t = z[s] - dsum;
        v += ((double)wloc[s] * t) * t;                                       // common.py:249-250
        if (wts_out) wts_out[s] = wloc[s];
    }
    depth = dsum; var = v; col[0] = c0; col[1] = c1; col[2] = c2;
}

// Backward of composite_ray.  g_*: upstream grads.  d_occ[S], d_rgb[S][3] (optional) are written.
DFP_HD void composite_ray_bwd(int S, const float* occ, const float* rgb, const double* z, double g_depth, double g_var,
                              const float g_col[3], float* d_occ, float* d_rgb) {
    float alpha[64], T[64], w[64];
    float t = 1.f;
    double depth = 0.0;
    for (int s = 0; s < S; ++s) {
        alpha[s] = sigmoid10(occ[s]);
        T[s] = t;
        w[s] = alpha[s] * t;
        t = t * ((1.f - alpha[s]) + 1e-10f);
        depth += (double)w[s] * z[s];
    }
    double dvar_ddepth = 0.0;                       // var = sum w (z - D)^2  ->  d var/d D = -2 sum w (z - D)
    for (int s = 0; s < S; ++s) dvar_ddepth += -2.0 * (double)w[s] * (z[s] - depth);
    const double gd = g_depth + g_var * dvar_ddepth;
    double suffix = 0.0;                            // sum_{k>s} G_k w_k
    for (int s = S - 1; s >= 0; --s) {
        const double dz = z[s] - depth;
        double G = gd * z[s] + g_var * dz * dz;
        if (rgb) G += (double)g_col[0] * rgb[3 * s] + (double)g_col[1] * rgb[3 * s + 1] + (double)g_col[2] * rgb[3 * s + 2];
        const double one_minus = (double)((1.f - alpha[s]) + 1e-10f);
        const double dalpha = G * (double)T[s] - suffix / one_minus;
        suffix += G * (double)w[s];
        d_occ[s] = (float)(dalpha * 10.0 * (double)alpha[s] * (1.0 - (double)alpha[s]));
        if (d_rgb) { d_rgb[3 * s] = w[s] * g_col[0]; d_rgb[3 * s + 1] = w[s] * g_col[1]; d_rgb[3 * s + 2] = w[s] * g_col[2]; }
    }
}

}  // namespace dfp
