// TEST INFRASTRUCTURE ONLY -- never linked into libdfprior_b200.so and never imported by the package.
//
// Serial host emulation of the kernel pipeline: it calls the very same per-thread
// functions (attentive_dfprior_b200/csrc/dfp_math.h, dfp_layout.h) that the sm_100a kernels
// call, one "thread" at a time, so that the arithmetic (sampling chain, trilinear set-up,
// decoder / attention forward and backward formulas, compositing and its backward, packed
// parameter layout) can be checked against the oracle on the GPU-less authoring box.
// The cooperative pieces of the kernels (shared-memory weight staging, warp sink, band
// compaction, vector reductions) have no host counterpart and are only exercised by the
// `-m gpu` tests.
//
// Build: g++ -O2 -ffp-contract=off -shared -fPIC (tests/test_hostemu.py does it).
#include <string.h>

#include <vector>

#include "../../attentive_dfprior_b200/csrc/dfp_math.h"
#include "../../include/dfprior_b200.h"

using namespace dfp;

namespace {

struct HostSink {
    float* g;
    void outer32(int woff, int ld, const float* a, int K, const float (&d)[32], int boff) {
        for (int k = 0; k < K; ++k)
            for (int j = 0; j < 32; ++j) g[woff + k * ld + j] += a[k] * d[j];
        if (boff >= 0)
            for (int j = 0; j < 32; ++j) g[boff + j] += d[j];
    }
    void outer4(int woff, int ld, const float* a, int K, const float (&d)[32]) { outer32(woff, ld, a, K, d, -1); }
    void bias32(int boff, const float (&d)[32]) { for (int j = 0; j < 32; ++j) g[boff + j] += d[j]; }
    void scalar(int off, float v) { g[off] += v; }
};

void gather(const DfpGrid& g, const Tri& t, float* c) {
    for (int j = 0; j < 32; ++j) c[j] = 0.f;
    for (int cn = 0; cn < 8; ++cn) {
        int x, y, z;
        if (!tri_corner(t, cn, g.W, g.H, g.D, x, y, z)) continue;
        const float w = tri_weight(t, cn);
        const float* src = g.data + (((int64_t)z * g.H + y) * g.W + x) * 32;
        for (int j = 0; j < 32; ++j) c[j] = fmaf(src[j], w, c[j]);
    }
}

void scatter(const DfpGrid& g, float* gg, const Tri& t, const float* dc, bool want_coord, float gn[3]) {
    float gx = 0, gy = 0, gz = 0;
    for (int cn = 0; cn < 8; ++cn) {
        int x, y, z;
        if (!tri_corner(t, cn, g.W, g.H, g.D, x, y, z)) continue;
        const int64_t off = (((int64_t)z * g.H + y) * g.W + x) * 32;
        const float w = tri_weight(t, cn);
        if (gg) for (int j = 0; j < 32; ++j) gg[off + j] += dc[j] * w;
        if (want_coord) {
            float dot = 0;
            for (int j = 0; j < 32; ++j) dot = fmaf(g.data[off + j], dc[j], dot);
            float dx, dy, dz;
            tri_weight_grad(t, cn, dx, dy, dz);
            gx += dot * dx; gy += dot * dy; gz += dot * dz;
        }
    }
    gn[0] = gx * t.gx; gn[1] = gy * t.gy; gn[2] = gz * t.gz;
}

template <int DEC>
void decoder_fwd(const DfpScene* sc, const float* packed, const double p[3], float* out) {
    constexpr int CD = DEC == 1 ? 64 : 32;
    constexpr int NOUT = DEC == 2 ? 4 : 1;
    const DfpGrid& g = DEC == 0 ? sc->low : (DEC == 1 ? sc->high : sc->color);
    const BoxT<double> box = make_box<double>(sc->bound);
    float xn[3];
    normalize(p, box, xn);
    float c[CD];
    gather(g, tri_setup(xn, g.W, g.H, g.D), c);
    if (DEC == 1) gather(sc->low, tri_setup(xn, sc->low.W, sc->low.H, sc->low.D), c + 32);
    const float pf[3] = {(float)p[0], (float)p[1], (float)p[2]};
    float o[NOUT];
    decoder_fwd_thread<CD, NOUT, false>(packed + dec_base(DEC), pf, c, o, nullptr, embed_bmax(packed + dec_base(DEC) + DecOff<CD>::B));
    for (int k = 0; k < NOUT; ++k) out[k] = o[k];
}

template <int DEC>
void decoder_bwd(const DfpScene* sc, const float* packed, const double p[3], const float* dout_in, float* ggrid,
                 float* gparams, float* dp) {
    constexpr int CD = DEC == 1 ? 64 : 32;
    constexpr int NOUT = DEC == 2 ? 4 : 1;
    const DfpGrid& g = DEC == 0 ? sc->low : (DEC == 1 ? sc->high : sc->color);
    const BoxT<double> box = make_box<double>(sc->bound);
    float xn[3];
    normalize(p, box, xn);
    const Tri tri = tri_setup(xn, g.W, g.H, g.D);
    float c[CD];
    gather(g, tri, c);
    if (DEC == 1) gather(sc->low, tri_setup(xn, sc->low.W, sc->low.H, sc->low.D), c + 32);
    const float pf[3] = {(float)p[0], (float)p[1], (float)p[2]};
    float o[NOUT], dout[NOUT];
    for (int k = 0; k < NOUT; ++k) dout[k] = (DEC == 2 && k == 3) ? 0.f : dout_in[k];
    DecState st;
    const float bmax = embed_bmax(packed + dec_base(DEC) + DecOff<CD>::B);
    decoder_fwd_thread<CD, NOUT, true>(packed + dec_base(DEC), pf, c, o, &st, bmax);
    float d_c[CD], dpe[3] = {0, 0, 0};
    HostSink sink{gparams};
    if (gparams) decoder_bwd_thread<CD, NOUT, true>(packed + dec_base(DEC), dec_base(DEC), pf, c, st, dout, d_c, dpe, sink, bmax);
    else decoder_bwd_thread<CD, NOUT, false>(packed + dec_base(DEC), dec_base(DEC), pf, c, st, dout, d_c, dpe, sink, bmax);
    float gn[3] = {0, 0, 0};
    scatter(g, ggrid, tri, d_c, dp != nullptr, gn);
    if (dp)
        for (int k = 0; k < 3; ++k) dp[k] += gn[k] * (float)(2.0 / (sc->bound[k][1] - sc->bound[k][0])) + dpe[k];
}

}  // namespace

extern "C" {

int emu_packed_floats() { return PACKED_FLOATS; }

void emu_pack(const float* const* tensors, float* packed) {
    for (int t = 0; t < N_TENSORS; ++t) {
        int r, c;
        tensor_shape(t, r, c);
        for (int i = 0; i < r * c; ++i) packed[packed_index(t, i / c, i % c)] = tensors[t][i];
    }
}
void emu_unpack(const float* packed, float* flat) {
    int off = 0;
    for (int t = 0; t < N_TENSORS; ++t) {
        int r, c;
        tensor_shape(t, r, c);
        for (int i = 0; i < r * c; ++i) flat[off + i] = packed[packed_index(t, i / c, i % c)];
        off += r * c;
    }
}

// forward (+ backward when og != NULL) of one render_batch_ray call, all pointers are HOST pointers
int emu_render(const DfpScene* sc, const float* packed, const DfpRays* rays, const DfpCfg* cfg, const DfpRenderOut* out,
               const DfpRenderOutGrad* og, const DfpGrads* g) {
    SampleTables T;
    T.n_samples = cfg->n_samples; T.n_surface = cfg->n_surface;
    memcpy(T.tv, cfg->t_vals, sizeof(T.tv)); memcpy(T.omt, cfg->one_minus_t, sizeof(T.omt));
    memcpy(T.ts, cfg->t_surf, sizeof(T.ts)); memcpy(T.omts, cfg->one_minus_t_surf, sizeof(T.omts));
    const bool has_depth = rays->gt_depth != nullptr;
    const int S = T.n_samples + (has_depth ? T.n_surface : 0);
    const int64_t N = rays->n_rays, NS = N * S;
    const int64_t chunk = rays->chunk > 0 ? rays->chunk : N;
    const int stage = cfg->stage;
    std::vector<double> z(NS);
    std::vector<float> low(NS), high(NS, 0.f), tsdf(NS, 0.f), fused(NS, 0.f), wat(NS, 1.f), rgb(NS * 4, 0.f);
    std::vector<int> flags(NS);
    const BoxT<double> box = make_box<double>(sc->bound), tbox = make_box<double>(sc->tsdf_bnds);
    const DfpTsdf& V = sc->tsdf;

    for (int64_t r = 0; r < N; ++r) {
        float mx = -INFINITY;
        if (has_depth) {
            const int64_t c0 = (r / chunk) * chunk, c1 = c0 + chunk < N ? c0 + chunk : N;
            for (int64_t i = c0; i < c1; ++i) mx = fmaxf(mx, rays->gt_depth[i]);
        }
        const float max12 = mx * 1.2f;
        const float* o = rays->rays_o + 3 * r;
        const float* d = rays->rays_d + 3 * r;
        const float gt = has_depth ? rays->gt_depth[r] : 0.f;
        const double far = ray_far(o, d, sc->bound, has_depth, max12);
        double cand[64];
        for (int i = 0; i < S; ++i) cand[i] = z_candidate(i, has_depth, gt, far, mx, T);
        for (int i = 0; i < S; ++i) {
            int rank = i;
            if (has_depth && T.n_surface > 0) {
                rank = 0;
                for (int j = 0; j < S; ++j) rank += (cand[j] < cand[i]) || (cand[j] == cand[i] && j < i);
            }
            z[r * S + rank] = cand[i];
        }
        for (int s = 0; s < S; ++s) {
            const int64_t i = r * S + s;
            double p[3];
            ray_point(o, d, z[i], p);
            int f = in_bound(p, box) ? 1 : 0;
            if (stage >= DFP_STAGE_HIGH) {
                float xn[3];
                normalize(p, tbox, xn);
                tsdf[i] = tsdf_interp(V.data, V.sD, V.sH, V.sW, V.W, V.H, V.D, tri_setup(xn, V.W, V.H, V.D));
                if (in_band(tsdf[i])) f |= 2;
                decoder_fwd<1>(sc, packed, p, &high[i]);
            }
            flags[i] = f;
            if (stage == DFP_STAGE_COLOR) decoder_fwd<2>(sc, packed, p, &rgb[4 * i]);
            decoder_fwd<0>(sc, packed, p, &low[i]);
            if (f & 2) {
                AttState st;
                float w;
                fused[i] = attention_fwd_thread(packed + OFF_ATT, high[i] + low[i], tsdf_logit(tsdf[i]).inv, w, st);
                wat[i] = w;
            }
        }
        float occ[64], c3[3], rg[192];
        for (int s = 0; s < S; ++s) {
            const int64_t i = r * S + s;
            occ[s] = (stage != DFP_STAGE_LOW && (flags[i] & 2)) ? fused[i] : low[i];
            if (!(flags[i] & 1)) occ[s] = 100.f;
            rg[3 * s] = rgb[4 * i]; rg[3 * s + 1] = rgb[4 * i + 1]; rg[3 * s + 2] = rgb[4 * i + 2];
            out->weight[i] = wat[i];
            if (out->raw) { out->raw[4 * i] = rg[3 * s]; out->raw[4 * i + 1] = rg[3 * s + 1]; out->raw[4 * i + 2] = rg[3 * s + 2]; out->raw[4 * i + 3] = occ[s]; }
            if (out->z_vals) out->z_vals[i] = z[i];
        }
        double depth, var;
        composite_ray(S, occ, stage == DFP_STAGE_COLOR ? rg : nullptr, &z[r * S], depth, var, c3, nullptr);
        out->depth[r] = depth; out->uncertainty[r] = var;
        for (int k = 0; k < 3; ++k) out->color[3 * r + k] = c3[k];

        if (!og) continue;
        // ---------------- backward of this ray ----------------
        float d_occ[64], d_rgb[192];
        const float gc[3] = {og->color ? og->color[3 * r] : 0.f, og->color ? og->color[3 * r + 1] : 0.f, og->color ? og->color[3 * r + 2] : 0.f};
        composite_ray_bwd(S, occ, stage == DFP_STAGE_COLOR ? rg : nullptr, &z[r * S], og->depth ? og->depth[r] : 0.0,
                          og->uncertainty ? og->uncertainty[r] : 0.0, gc, d_occ, d_rgb);
        const bool want_dp = g->rays_o || g->rays_d;
        double so[3] = {0, 0, 0}, sd[3] = {0, 0, 0};
        for (int s = 0; s < S; ++s) {
            const int64_t i = r * S + s;
            const int f = flags[i];
            const float gocc = (f & 1) ? d_occ[s] : 0.f;
            const bool band = stage != DFP_STAGE_LOW && (f & 2);
            float d_low = band ? 0.f : gocc, d_high = 0.f;
            float dp[3] = {0, 0, 0};
            double p[3];
            ray_point(o, d, z[i], p);
            if (band) {
                const float x0 = high[i] + low[i];
                const AttLogit lg = tsdf_logit(tsdf[i]);
                AttState st;
                float w, dx0, dx1;
                attention_fwd_thread(packed + OFF_ATT, x0, lg.inv, w, st);
                HostSink sink{g->params_packed};
                const float gw = og->weight ? og->weight[i] : 0.f;
                if (g->params_packed) attention_bwd_thread<true>(packed + OFF_ATT, OFF_ATT, x0, lg.inv, st, gocc, gw, dx0, dx1, sink);
                else attention_bwd_thread<false>(packed + OFF_ATT, OFF_ATT, x0, lg.inv, st, gocc, gw, dx0, dx1, sink);
                d_low = dx0; d_high = dx0;
                if (want_dp) {
                    float xn[3], gn[3];
                    normalize(p, tbox, xn);
                    tsdf_interp_grad(V.data, V.sD, V.sH, V.sW, V.W, V.H, V.D, tri_setup(xn, V.W, V.H, V.D), gn);
                    const float dt = dx1 * lg.dinv_dtsdf;
                    for (int k = 0; k < 3; ++k) dp[k] += dt * gn[k] * (float)(2.0 / (sc->tsdf_bnds[k][1] - sc->tsdf_bnds[k][0]));
                }
            }
            if (stage >= DFP_STAGE_HIGH) decoder_bwd<1>(sc, packed, p, &d_high, g->grid_high, g->params_packed, want_dp ? dp : nullptr);
            if (stage == DFP_STAGE_COLOR) {
                const float dr[4] = {d_rgb[3 * s], d_rgb[3 * s + 1], d_rgb[3 * s + 2], 0.f};
                decoder_bwd<2>(sc, packed, p, dr, g->grid_color, g->params_packed, want_dp ? dp : nullptr);
            }
            decoder_bwd<0>(sc, packed, p, &d_low, g->grid_low, g->params_packed, want_dp ? dp : nullptr);
            for (int k = 0; k < 3; ++k) { so[k] += dp[k]; sd[k] += (double)dp[k] * z[i]; }
        }
        for (int k = 0; k < 3; ++k) {
            if (g->rays_o) g->rays_o[3 * r + k] = (float)so[k];
            if (g->rays_d) g->rays_d[3 * r + k] = (float)sd[k];
        }
    }
    return 0;
}

// grid split of the merged decoder launches (dfp_layout.h): first CTA and CTA count of each of the three decoders
void emu_split_ctas(const int64_t* it, int G, int* first3, int* count3) {
    for (int d = 0; d < 3; ++d) dfp::split_ctas(it, G, d, first3[d], count3[d]);
}

}  // extern "C"
