// attentive fusion kernels over the compacted band list
#include "dfp_device.cuh"

namespace dfp {

__global__ void __launch_bounds__(NT) k_attn_fwd(const AttArgs a) {
    extern __shared__ float4 smem4[];
    float* A = reinterpret_cast<float*>(smem4);
    load_weights_to_smem(A, a.params + OFF_ATT, AttOff::SIZE);
    __syncthreads();
    const int count = *a.band_count;
    // 32-sample chunks are dealt round-robin over the CTAs first and their warps second, so that a short band list
    // (the Mapper's 1000-ray batches) runs one warp on every SM instead of eight warps on a few
    for (int chunk = (threadIdx.x >> 5) * gridDim.x + blockIdx.x; chunk * 32 < count; chunk += WARPS * gridDim.x) {
        const int i = chunk * 32 + (threadIdx.x & 31);
        if (i >= count) continue;
        const int s = a.band_list[i];
        const float x0 = a.high[s] + a.low[s];                         // decoder.py:326,343
        const float x1 = tsdf_logit(a.tsdf[s]).inv;
        AttState st;
        float w;
        const float f = attention_fwd_thread(A, x0, x1, w, st);
        a.fused[s] = f;
        a.w_attn[s] = w;
        if (a.act) {                                           // kept for the backward (DFP_FLAG_SAVE_ACT): no recompute there
            float* dst = a.act + att_act_index(i, 0);
#pragma unroll 8
            for (int k = 0; k < 64; ++k) dst[k * 32] = st.h1[k];
#pragma unroll 8
            for (int k = 0; k < 128; ++k) dst[(64 + k) * 32] = st.h2[k];
#pragma unroll 8
            for (int k = 0; k < 128; ++k) dst[(192 + k) * 32] = st.h3[k];
#pragma unroll 8
            for (int k = 0; k < 64; ++k) dst[(320 + k) * 32] = st.h4[k];
            dst[384 * 32] = st.a0;
            dst[385 * 32] = st.a1;
        }
    }
}

template <int MODE, bool WANT_PARAMS>
__global__ void __launch_bounds__(NT) k_attn_bwd(const AttArgs a) {
    using PT = typename PtT<MODE>::T;
    extern __shared__ float4 smem4[];
    float* A = reinterpret_cast<float*>(smem4);
    load_weights_to_smem(A, a.params + OFF_ATT, AttOff::SIZE);
    WarpSink sink;
    sink.lane = threadIdx.x & 31;
    sink.g = a.gparams;
    sink.As = A + AttOff::SIZE + (threadIdx.x >> 5) * SINK_FLOATS_PER_WARP;
    sink.Dt = sink.As + 32 * 32;
    __syncthreads();
    const int count = *a.band_count;
    const BoxT<PT> tbox = make_box<PT>(a.tsdf_bnds.b);
    for (int chunk = (threadIdx.x >> 5) * gridDim.x + blockIdx.x; chunk * 32 < count; chunk += WARPS * gridDim.x) {     // as in k_attn_fwd
        const int i = chunk * 32 + (threadIdx.x & 31);
        const bool valid = i < count;
        const int s = a.band_list[valid ? i : count - 1];
        const float g_out = valid ? a.d_fused[s] : 0.f;
        const float g_w = (valid && a.g_w) ? a.g_w[s] : 0.f;
        const bool any = (g_out != 0.f) || (g_w != 0.f);
        if (!__any_sync(0xffffffffu, any)) continue;
        const float x0 = a.high[s] + a.low[s];
        const AttLogit lg = tsdf_logit(a.tsdf[s]);
        AttState st;
        if (a.act) {                                           // activations saved by k_attn_fwd
            const float* src = a.act + att_act_index(valid ? i : count - 1, 0);
#pragma unroll 8
            for (int k = 0; k < 64; ++k) st.h1[k] = src[k * 32];
#pragma unroll 8
            for (int k = 0; k < 128; ++k) st.h2[k] = src[(64 + k) * 32];
#pragma unroll 8
            for (int k = 0; k < 128; ++k) st.h3[k] = src[(192 + k) * 32];
#pragma unroll 8
            for (int k = 0; k < 64; ++k) st.h4[k] = src[(320 + k) * 32];
            st.a0 = src[384 * 32];
            st.a1 = src[385 * 32];
        } else {
            float w;
            attention_fwd_thread(A, x0, lg.inv, w, st);
        }
        float dx0, dx1;
        attention_bwd_thread<WANT_PARAMS>(A, OFF_ATT, x0, lg.inv, st, g_out, g_w, dx0, dx1, sink);
        if (valid) {
            a.d_low[s] = dx0;                                          // low_occ enters only through high+low on band points
            a.d_high[s] = dx0;
            if (a.dp) {
                const float dt = dx1 * lg.dinv_dtsdf;                  // gradient into the TSDF value gathered at p (decoder.py:241)
                if (dt != 0.f) {
                    PT p[3];
                    load_point<MODE>(a.ps, s, p);
                    float xn[3];
                    normalize(p, tbox, xn);
                    const Tri t = tri_setup(xn, a.vol.W, a.vol.H, a.vol.D);
                    float gn[3];
                    tsdf_interp_grad(a.vol.data, a.vol.sD, a.vol.sH, a.vol.sW, a.vol.W, a.vol.H, a.vol.D, t, gn);
                    for (int k = 0; k < 3; ++k)
                        a.dp[3 * (int64_t)s + k] += dt * gn[k] * (float)(2.0 / (a.tsdf_bnds.b[k][1] - a.tsdf_bnds.b[k][0]));
                }
            }
        }
    }
}


int launch_attn_fwd(const AttArgs& aa, int64_t n, cudaStream_t st) {
    const size_t smem = AttOff::SIZE * sizeof(float);
    int rc = set_smem(k_attn_fwd, smem);
    if (rc) return rc;
    DFP_TIMED("k_attn_fwd", st, k_attn_fwd<<<grid_for(n, NT, num_sms()), NT, smem, st>>>(aa));
    DFP_LAUNCH_CHECK("k_attn_fwd");
    return DFP_OK;
}

template <int MODE>
int launch_attn_bwd(const AttArgs& aa, int64_t n, cudaStream_t st) {
    const size_t smem = (AttOff::SIZE + WARPS * SINK_FLOATS_PER_WARP) * sizeof(float);
    const int blocks = grid_for(n, NT, num_sms());
    int rc;
    if (aa.gparams) {
        if ((rc = set_smem(k_attn_bwd<MODE, true>, smem))) return rc;
        DFP_TIMED("k_attn_bwd", st, k_attn_bwd<MODE, true><<<blocks, NT, smem, st>>>(aa));
    } else {
        if ((rc = set_smem(k_attn_bwd<MODE, false>, smem))) return rc;
        DFP_TIMED("k_attn_bwd", st, k_attn_bwd<MODE, false><<<blocks, NT, smem, st>>>(aa));
    }
    DFP_LAUNCH_CHECK("k_attn_bwd");
    return DFP_OK;
}
template int launch_attn_bwd<0>(const AttArgs&, int64_t, cudaStream_t);
template int launch_attn_bwd<1>(const AttArgs&, int64_t, cudaStream_t);
template int launch_attn_bwd<2>(const AttArgs&, int64_t, cudaStream_t);

}  // namespace dfp
