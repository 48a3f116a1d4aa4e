// decoder backward kernel + launcher; included once per point mode (DFP_BWD_MODE) so the units build in parallel
#include "dfp_device.cuh"

namespace dfp {

// ------------------------------------------------------------------------------------------
// decoder backward: recompute the forward with saved activations, back-propagate
// ------------------------------------------------------------------------------------------
template <int DEC, int MODE, bool WANT_PARAMS>
__global__ void __launch_bounds__(NT_BWD, 1) k_decoder_bwd(const DecArgs a) {
    constexpr int CD = (DEC == 1) ? 64 : 32;
    constexpr int NOUT = (DEC == 2) ? 4 : 1;
    using PT = typename PtT<MODE>::T;
    using L = DecOff<CD>;
    extern __shared__ float4 smem4[];
    float* W = reinterpret_cast<float*>(smem4);
    constexpr int gbase = (DEC == 0 ? OFF_LOW : (DEC == 1 ? OFF_HIGH : OFF_COLOR));
    load_weights_to_smem(W, a.params + gbase, L::SIZE);
    WarpSink sink;
    sink.lane = threadIdx.x & 31;
    sink.g = a.gparams;
    sink.As = W + L::SIZE + (threadIdx.x >> 5) * SINK_FLOATS_PER_WARP;
    sink.Dt = sink.As + 32 * 32;
    __syncthreads();
    const BoxT<PT> box = make_box<PT>(a.bound.b);
    const float bmax = embed_bmax(W + DecOff<CD>::B);
    for (int64_t base = (int64_t)blockIdx.x * NT_BWD; base < a.n; base += (int64_t)gridDim.x * NT_BWD) {
        const int64_t s_raw = base + threadIdx.x;
        const bool valid = s_raw < a.n;
        const int64_t s = valid ? s_raw : a.n - 1;          // keep the whole warp in the cooperative sink
        float dout[NOUT];
        if (NOUT == 1) dout[0] = valid ? a.dout[s] : 0.f;
        else {
            const float4 v = *reinterpret_cast<const float4*>(a.dout + 4 * s);
            dout[0] = valid ? v.x : 0.f; dout[NOUT > 1 ? 1 : 0] = valid ? v.y : 0.f; dout[NOUT > 2 ? 2 : 0] = valid ? v.z : 0.f;
            dout[NOUT > 3 ? 3 : 0] = 0.f;                   // colour decoder's 4th output is overwritten (decoder.py:352)
        }
        bool any = false;
#pragma unroll
        for (int o = 0; o < NOUT; ++o) any = any || (dout[o] != 0.f);
        if (!__any_sync(0xffffffffu, any)) continue;         // nothing flows into this warp's samples
        PT p[3];
        load_point<MODE>(a.ps, s, p);
        float xn[3];
        normalize(p, box, xn);
        const Tri tri = tri_setup(xn, a.grid.W, a.grid.H, a.grid.D);
        float c[CD];
        gather_features(a.grid, tri, c);
        if (DEC == 1) gather_features(a.grid_low, tri_setup(xn, a.grid_low.W, a.grid_low.H, a.grid_low.D), c + CDIM);
        const float pf[3] = {(float)p[0], (float)p[1], (float)p[2]};
        DecState st;
        float out[NOUT];
        decoder_fwd_thread<CD, NOUT, true>(W, pf, c, out, &st, bmax);
        float d_c[CD];
        float dpe[3] = {0.f, 0.f, 0.f};
        decoder_bwd_thread<CD, NOUT, WANT_PARAMS>(W, gbase, pf, c, st, dout, d_c, dpe, sink, bmax);
        if (valid && any) {
            const bool want_coord = a.dp != nullptr;
            float gn[3] = {0.f, 0.f, 0.f};
            if (a.ggrid || want_coord) scatter_features(a.grid, a.ggrid, tri, d_c, want_coord, gn);
            if (want_coord) {
                // d p = d xn * 2/(b1-b0)  (common.py:287-289)  +  embedding path (decoder.py:189-191)
                a.dp[3 * s + 0] += gn[0] * (float)(2.0 / (a.bound.b[0][1] - a.bound.b[0][0])) + dpe[0];
                a.dp[3 * s + 1] += gn[1] * (float)(2.0 / (a.bound.b[1][1] - a.bound.b[1][0])) + dpe[1];
                a.dp[3 * s + 2] += gn[2] * (float)(2.0 / (a.bound.b[2][1] - a.bound.b[2][0])) + dpe[2];
            }
        }
    }
}


template <int DEC, int MODE>
int launch_decoder_bwd(const DecArgs& a, cudaStream_t st) {
    constexpr int CD = (DEC == 1) ? 64 : 32;
    const size_t smem = (DecOff<CD>::SIZE + WARPS_BWD * SINK_FLOATS_PER_WARP) * sizeof(float);
    const int blocks = grid_for(a.n, NT_BWD, num_sms());
    if (a.gparams) {
        int rc = set_smem(k_decoder_bwd<DEC, MODE, true>, smem);
        if (rc) return rc;
        DFP_TIMED("k_decoder_bwd", st, k_decoder_bwd<DEC, MODE, true><<<blocks, NT_BWD, smem, st>>>(a));
    } else {
        int rc = set_smem(k_decoder_bwd<DEC, MODE, false>, smem);
        if (rc) return rc;
        DFP_TIMED("k_decoder_bwd", st, k_decoder_bwd<DEC, MODE, false><<<blocks, NT_BWD, smem, st>>>(a));
    }
    DFP_LAUNCH_CHECK("k_decoder_bwd");
    return DFP_OK;
}


template int launch_decoder_bwd<0, DFP_BWD_MODE>(const DecArgs&, cudaStream_t);
template int launch_decoder_bwd<1, DFP_BWD_MODE>(const DecArgs&, cudaStream_t);
template int launch_decoder_bwd<2, DFP_BWD_MODE>(const DecArgs&, cudaStream_t);

}  // namespace dfp
