// decoder forward kernels (one sample per thread, weights broadcast from shared memory)
#include "dfp_device.cuh"

namespace dfp {

template <int DEC, int MODE>
__global__ void __launch_bounds__(NT, 2) k_decoder_fwd(const DecArgs a) {
    constexpr int CD = (DEC == 1) ? 64 : 32;
    constexpr int NOUT = (DEC == 2) ? 4 : 1;
    using PT = typename PtT<MODE>::T;
    extern __shared__ float4 smem4[];
    float* W = reinterpret_cast<float*>(smem4);
    load_weights_to_smem(W, a.params + (DEC == 0 ? OFF_LOW : (DEC == 1 ? OFF_HIGH : OFF_COLOR)), DecOff<CD>::SIZE);
    __syncthreads();
    const BoxT<PT> box = make_box<PT>(a.bound.b);
    const float bmax = embed_bmax(W + DecOff<CD>::B);
    for (int64_t base = (int64_t)blockIdx.x * NT; base < a.n; base += (int64_t)gridDim.x * NT) {
        const int64_t s = base + threadIdx.x;
        if (s >= a.n) continue;
        PT p[3];
        load_point<MODE>(a.ps, s, p);
        float xn[3];
        normalize(p, box, xn);
        float c[CD];
        gather_features(a.grid, tri_setup(xn, a.grid.W, a.grid.H, a.grid.D), c);
        if (DEC == 1) gather_features(a.grid_low, tri_setup(xn, a.grid_low.W, a.grid_low.H, a.grid_low.D), c + CDIM);
        const float pf[3] = {(float)p[0], (float)p[1], (float)p[2]};       // p.float()   decoder.py:189
        float out[NOUT];
        decoder_fwd_thread<CD, NOUT, false>(W, pf, c, out, nullptr, bmax);
        if (NOUT == 1) a.out[s] = out[0];
        else *reinterpret_cast<float4*>(a.out + 4 * s) = make_float4(out[0], out[NOUT > 1 ? 1 : 0], out[NOUT > 2 ? 2 : 0], out[NOUT > 3 ? 3 : 0]);
    }
}


template <int DEC, int MODE>
int launch_decoder_fwd(const DecArgs& a, cudaStream_t st) {
    constexpr int CD = (DEC == 1) ? 64 : 32;
    const size_t smem = DecOff<CD>::SIZE * sizeof(float);
    int rc = set_smem(k_decoder_fwd<DEC, MODE>, smem);
    if (rc) return rc;
    const int blocks = grid_for(a.n, NT, 2 * num_sms());
    DFP_TIMED("k_decoder_fwd", st, k_decoder_fwd<DEC, MODE><<<blocks, NT, smem, st>>>(a));
    DFP_LAUNCH_CHECK("k_decoder_fwd");
    return DFP_OK;
}

template int launch_decoder_fwd<0, 0>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd<0, 1>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd<0, 2>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd<1, 0>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd<1, 1>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd<1, 2>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd<2, 0>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd<2, 1>(const DecArgs&, cudaStream_t);
template int launch_decoder_fwd<2, 2>(const DecArgs&, cudaStream_t);

}  // namespace dfp
