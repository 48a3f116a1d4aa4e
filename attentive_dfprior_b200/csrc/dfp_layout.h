// Packed decoder-parameter layout shared by the kernels, the pack/unpack kernels and
// the host-side emulation used in unit tests.
//
// The reference keeps 79 nn.Parameter tensors in [out,in] layout
// (src/conv_onet/models/decoder.py:124-159, 218-225).  The kernels evaluate the MLPs
// one sample per thread with weights broadcast from shared memory, which wants each
// weight matrix k-major (W[k][j], j contiguous) so that one 128-bit shared load gives
// four output neurons of the same input k.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DFP_HD __host__ __device__ __forceinline__
#else
#define DFP_HD inline
#endif

namespace dfp {

constexpr int HID = 32;
constexpr int EMB = 93;
constexpr int EMB_PAD = 96;
constexpr int CDIM = 32;
constexpr int N_TENSORS = 79;
constexpr int N_PARAMS = 86029;

// ---- one grid-conditioned decoder (decoder.py:91-203), feature width CD (32, or 64 for "high")
template <int CD>
struct DecOff {
    static constexpr int B = 0;                        // [3][96]     embedder._B rows (93 used)
    static constexpr int W03 = B + 3 * EMB_PAD;        // [93][64]    cols 0..31 pts_linears.0, cols 32..63 pts_linears.3[:, :93]
    static constexpr int WH = W03 + EMB * 64;          // [4][32][32] pts_linears.1, .2, .3[:, 93:], .4  (k-major)
    static constexpr int WC = WH + 4 * HID * HID;      // [5][CD][32] fc_c.i (k-major)
    static constexpr int BP = WC + 5 * CD * HID;       // [5][32]     pts_linears.i.bias
    static constexpr int BC = BP + 5 * HID;            // [5][32]     fc_c.i.bias
    static constexpr int WO = BC + 5 * HID;            // [4][32]     output_linear.weight rows (1 or 4 used)
    static constexpr int BO = WO + 4 * HID;            // [4]         output_linear.bias
    static constexpr int SIZE = BO + 4;
};

// ---- attention fusion MLP 2->64->128->128->64->2 (decoder.py:206-258)
struct AttOff {
    static constexpr int W1 = 0;                       // [2][64]    k-major
    static constexpr int B1 = W1 + 2 * 64;             // [64]
    static constexpr int W2 = B1 + 64;                 // [64][128]
    static constexpr int B2 = W2 + 64 * 128;           // [128]
    static constexpr int W3 = B2 + 128;                // [128][128]
    static constexpr int B3 = W3 + 128 * 128;          // [128]
    static constexpr int W4 = B3 + 128;                // [128][64]
    static constexpr int B4 = W4 + 128 * 64;           // [64]
    static constexpr int W5 = B4 + 64;                 // [2][64]    output rows
    static constexpr int B5 = W5 + 2 * 64;             // [2] (+2 pad)
    static constexpr int SIZE = B5 + 4;
};

constexpr int OFF_LOW = 0;
constexpr int OFF_HIGH = OFF_LOW + DecOff<32>::SIZE;
constexpr int OFF_COLOR = OFF_HIGH + DecOff<64>::SIZE;
constexpr int OFF_ATT = OFF_COLOR + DecOff<32>::SIZE;
constexpr int PACKED_FLOATS = OFF_ATT + AttOff::SIZE;      // fp32 parameters (and the size of a packed gradient)

// ---- tensor-core operand images appended to the packed PARAMETER buffer by the pack kernel (not part of gradients) ----
// Fusion MLP: the three wide layers as bf16 hi + mid planes (w ~= hi + mid, |w - hi - mid| <= 2^-18 |w|) in the
// canonical K-major core-matrix order of a B operand with N = outputs:  byte(n = out j, k = in) =
// (k/8) * (N*16) + n*16 + (k%8)*2.  The forward reads it K-major (LBO = N*16, SBO = 128); the backward chain reads the
// very same bytes transposed as an MN-major operand (LBO = 128, SBO = N*16) -- both readings probed on B200
// (tools/probe_bf16_kmajor.py).  One bulk copy brings the 128 KB into shared memory.
struct AttImg {
    static constexpr int W2H = 0;                          // pts_linears.1: in 64,  out 128   (bytes)
    static constexpr int W2M = W2H + 64 * 128 * 2;
    static constexpr int W3H = W2M + 64 * 128 * 2;         // pts_linears.2: in 128, out 128
    static constexpr int W3M = W3H + 128 * 128 * 2;
    static constexpr int W4H = W3M + 128 * 128 * 2;        // pts_linears.3: in 128, out 64
    static constexpr int W4M = W4H + 128 * 64 * 2;
    static constexpr int BYTES = W4M + 128 * 64 * 2;       // 131072
};
// Two copies: bf16 planes (fp32's exponent range: the backward chain multiplies them with gradients of arbitrary
// scale) and fp16 planes for the forward (hi + mid = 22 mantissa bits, the accuracy of the decoders' 3xTF32 scheme:
// the forward decides ReLU masks, and with bf16's 2^-17 about one band sample in 300 lands on the other side of an
// edge than the fp32 reference does, which small gradient parity cases notice).
constexpr float ATT_IMG_F16_SCALE = 64.f;          // the fp16 image holds 64 w: a weight of 0.1 keeps its "mid" part a normal fp16 number
constexpr int OFF_ATT_IMG = (PACKED_FLOATS + 31) / 32 * 32;             // 128-byte aligned (bulk-copy source)
constexpr int OFF_ATT_IMG_F16 = OFF_ATT_IMG + AttImg::BYTES / 4;
// Decoders, backward: the feature gradient d_c = sum_i dh_i Wc_i is linear in the ReLU-masked deltas (dh_{i-1} = dacc_i W_i),
// so the delta-chain kernel multiplies every masked delta by [M_i | W_i] in ONE product, M_i = W_i Wc_{i-1} (restricted to
// the decoder's own 32 grid features), and adds the dh_4 = dout W_o term as dout V, V = W_o Wc_4.  Per decoder:
struct DecM {
    static constexpr int M = 0;                            // [4][32 j][32 q]  M_1 .. M_4 (slot i - 1), row = delta index j
    static constexpr int V = M + 4 * HID * HID;            // [4][32 q]        rows of V (1 or 4 used)
    static constexpr int SIZE = V + 4 * HID;
};
constexpr int OFF_DEC_M = OFF_ATT_IMG_F16 + AttImg::BYTES / 4;          // low, high, colour
constexpr int PACKED_TOTAL_FLOATS = OFF_DEC_M + 3 * DecM::SIZE;

static_assert(DecOff<32>::SIZE % 4 == 0 && DecOff<64>::SIZE % 4 == 0 && AttOff::SIZE % 4 == 0, "float4 copies");

// ---- the 79 tensors in state_dict order --------------------------------------------
// per decoder (23): fc_c.{0..4}.{weight,bias}, embedder._B, pts_linears.{0..4}.{weight,bias},
//                   output_linear.{weight,bias};   then mlp (10): pts_linears.{0..3}.{weight,bias},
//                   output_linear.{weight,bias}.
struct TensorDesc {
    int rows, cols;      // native [rows][cols] (bias: rows = 1)
    int native_off;      // offset in the flat native buffer (tensors back to back)
};

DFP_HD int dec_cd(int dec) { return dec == 1 ? 64 : 32; }
DFP_HD int dec_nout(int dec) { return dec == 2 ? 4 : 1; }
DFP_HD int dec_base(int dec) { return dec == 0 ? OFF_LOW : (dec == 1 ? OFF_HIGH : OFF_COLOR); }

// shape of tensor t
DFP_HD void tensor_shape(int t, int& rows, int& cols) {
    if (t < 69) {
        const int dec = t / 23, u = t % 23;
        const int cd = dec_cd(dec), nout = dec_nout(dec);
        if (u < 10) { if (u & 1) { rows = 1; cols = HID; } else { rows = HID; cols = cd; } }
        else if (u == 10) { rows = 3; cols = EMB; }
        else if (u < 21) {
            const int i = (u - 11) >> 1;
            if ((u - 11) & 1) { rows = 1; cols = HID; }
            else { rows = HID; cols = (i == 0) ? EMB : (i == 3 ? HID + EMB : HID); }
        } else if (u == 21) { rows = nout; cols = HID; }
        else { rows = 1; cols = nout; }
    } else {
        const int u = t - 69;
        const int din[5] = {2, 64, 128, 128, 64};
        const int dout[5] = {64, 128, 128, 64, 2};
        const int i = u >> 1;
        if (u & 1) { rows = 1; cols = dout[i]; } else { rows = dout[i]; cols = din[i]; }
    }
}

// packed index of native element (r, c) of tensor t
DFP_HD int packed_index(int t, int r, int c) {
    if (t < 69) {
        const int dec = t / 23, u = t % 23;
        const int cd = dec_cd(dec);
        const int base = dec_base(dec);
        // the member offsets that depend on CD
        const int WC = DecOff<32>::WC;                       // same for both (WC precedes the CD-sized block)
        const int BP = WC + 5 * cd * HID, BC = BP + 5 * HID, WO = BC + 5 * HID, BO = WO + 4 * HID;
        if (u < 10) {
            const int i = u >> 1;
            if (u & 1) return base + BC + i * HID + c;                         // fc_c.i.bias[j=c]
            return base + WC + i * cd * HID + c * HID + r;                      // fc_c.i.weight[j=r][k=c]
        }
        if (u == 10) return base + DecOff<32>::B + r * EMB_PAD + c;             // _B[r][k]
        if (u < 21) {
            const int i = (u - 11) >> 1;
            if ((u - 11) & 1) return base + BP + i * HID + c;                   // pts_linears.i.bias
            if (i == 0) return base + DecOff<32>::W03 + c * 64 + r;
            if (i == 3) return (c < EMB) ? base + DecOff<32>::W03 + c * 64 + 32 + r
                                         : base + DecOff<32>::WH + 2 * HID * HID + (c - EMB) * HID + r;
            const int slot = (i == 4) ? 3 : i - 1;
            return base + DecOff<32>::WH + slot * HID * HID + c * HID + r;
        }
        if (u == 21) return base + WO + r * HID + c;                            // output_linear.weight[o][j]
        return base + BO + c;                                                   // output_linear.bias[o]
    }
    const int u = t - 69;
    const int i = u >> 1;
    const int woff[5] = {AttOff::W1, AttOff::W2, AttOff::W3, AttOff::W4, AttOff::W5};
    const int boff[5] = {AttOff::B1, AttOff::B2, AttOff::B3, AttOff::B4, AttOff::B5};
    const int dout[5] = {64, 128, 128, 64, 2};
    if (u & 1) return OFF_ATT + boff[i] + c;
    if (i == 4) return OFF_ATT + AttOff::W5 + r * 64 + c;                       // output rows [o][j]
    return OFF_ATT + woff[i] + c * dout[i] + r;                                 // k-major: [k=c][j=r]
}


// ---- merged launches: one kernel for up to three decoders (dfp_device.cuh, DecArgs3) ----
// (first CTA, number of CTAs) of decoder `which` when `G` CTAs are split over decoders needing it[0..2] CTA-iterations
DFP_HD void split_ctas(const int64_t it[3], int G, int which, int& first, int& count) {
    const int64_t total = it[0] + it[1] + it[2];
    int n[3] = {0, 0, 0};
    if (total > 0) {
        int used = 0, big = 0;
        for (int d = 0; d < 3; ++d) {
            if (it[d] > 0) { n[d] = (int)((it[d] * G) / total); if (n[d] < 1) n[d] = 1; if ((int64_t)n[d] > it[d]) n[d] = (int)it[d]; }
            used += n[d];
            if (it[d] > it[big]) big = d;
        }
        // rounding: hand spare CTAs to (or take missing ones from) the decoder with the most work
        if (used < G) { const int64_t room = it[big] - n[big]; n[big] += (int)((G - used) < room ? (G - used) : room); }
        while (used > G) { int m = 0; for (int d = 1; d < 3; ++d) if (n[d] > n[m]) m = d; --n[m]; --used; }
    }
    first = 0;
    for (int d = 0; d < which; ++d) first += n[d];
    count = n[which];
}

}  // namespace dfp
