"""GPU: tcgen05 / TMEM building blocks (dfp_tc.cuh, through the test-only library tests/selftest) against an f64
matmul, and the tensor-core product path against the in-tree fp32 SIMT kernels (forward and backward)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "selftest"))
import binding as ST  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(32, 32), (32, 64), (64, 160), (96, 64), (8, 16)])
def test_tc_selftest_3xtf32(K, N):
    lib = ST.lib()
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    ref = (A.double() @ W.double().t())
    for split, tol in ((1, 2e-6), (0, 5e-3)):
        D = torch.full((128, N), float("nan"), device="cuda")
        ST.check(lib.dfpt_selftest_tc(A.data_ptr(), W.data_ptr(), D.data_ptr(), K, N, split,
                                    torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        err = float((D.double() - ref).abs().max()) / float(ref.abs().max())
        assert err <= tol, (K, N, split, err)
        if split == 0:
            assert err > 1e-6, "single-pass TF32 should be visibly less accurate than the 3x split"


@pytest.mark.parametrize("stage", ["low", "high", "color"])
def test_tensor_core_and_simt_paths_agree(stage):
    """The tcgen05 decoder path and the fp32 SIMT path (DFP_FLAG_SIMT, a per-call flag) give the same point query within
    1e-5 forward and 1e-4 in every gradient."""
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200 import functional as Fn
    from attentive_dfprior_b200 import synth
    torch.manual_seed(1)
    scene = synth.make_scene("small", seed=5, device="cuda", high_std=1e-2, channels_last=True)
    model = A.DF().cuda()
    with torch.no_grad():
        for dec in (model.low_decoder, model.high_decoder):
            dec.output_linear.weight.mul_(0.05)
    g = torch.Generator().manual_seed(3)
    lo, hi = scene.bound[:, 0], scene.bound[:, 1]
    p0 = (torch.rand(5000, 3, generator=g, dtype=torch.float64) * (hi - lo) + lo).cuda()
    gr = torch.randn(5000, 4, generator=g).cuda()
    gw = torch.randn(5000, generator=g).cuda()
    res = {}
    for simt in (False, True):
        grids = {k: v.detach().clone().requires_grad_(True) for k, v in scene.grids.items()}
        for prm in model.parameters():
            prm.grad = None
        p = p0.clone().requires_grad_(True)
        raw, w = Fn.eval_points(p, model, grids, scene.tsdf_volume, scene.tsdf_bnds, scene.bound, stage, simt=simt)
        ((raw * gr).sum() + (w * gw).sum()).backward()
        res[simt] = (raw.detach(), w.detach(), p.grad.clone(), {k: v.grad.clone() for k, v in grids.items() if v.grad is not None},
                     [q.grad.clone() if q.grad is not None else None for q in model.parameters()])
    raw_tc, w_tc, dp_tc, gg_tc, gp_tc = res[False]
    raw_sm, w_sm, dp_sm, gg_sm, gp_sm = res[True]
    scale = float(raw_sm[:, 3][raw_sm[:, 3] != 100].abs().max())
    assert float((raw_tc - raw_sm).abs().max()) <= 1e-5 * max(scale, 1.0)
    assert float((w_tc - w_sm).abs().max()) <= 1e-5

    def rel(a, b):
        return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)
    assert rel(dp_tc, dp_sm) <= 1e-4
    for k in gg_sm:
        assert rel(gg_tc[k], gg_sm[k]) <= 1e-4, k
    for a, b in zip(gp_tc, gp_sm):
        if b is not None and float(b.abs().max()) > 0:
            assert rel(a, b) <= 1e-4


@pytest.mark.parametrize("J,KIN", [(32, 32), (64, 96), (32, 64)])
def test_tc_transposed_weight_view(J, KIN):
    """dH pattern: D = A[128,J] @ Wf[J,KIN] through the MN-major view of the canonical (forward) weight buffer."""
    lib = ST.lib()
    g = torch.Generator().manual_seed(J * 100 + KIN)
    A = torch.randn(128, J, generator=g).cuda()
    Wf = torch.randn(J, KIN, generator=g).cuda()
    D = torch.full((128, KIN), float("nan"), device="cuda")
    ST.check(lib.dfpt_selftest_tc(A.data_ptr(), Wf.data_ptr(), D.data_ptr(), J, KIN, 11, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = A.double() @ Wf.double()
    assert float((D.double() - ref).abs().max()) / float(ref.abs().max()) <= 2e-6


@pytest.mark.parametrize("F,N", [(32, 32), (32, 64), (4, 16), (32, 16)])
def test_tc_outer_product_from_smem(F, N):
    """dW pattern: D[:F] = H[128,F]^T @ Dl[128,N], both operands sample-major tiles in shared memory, M=128 over-read."""
    lib = ST.lib()
    g = torch.Generator().manual_seed(F * 100 + N)
    H = torch.randn(128, F, generator=g).cuda()
    Dl = torch.randn(128, N, generator=g).cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    ST.check(lib.dfpt_selftest_tc(H.data_ptr(), Dl.data_ptr(), D.data_ptr(), F, N, 12, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = H.double().t() @ Dl.double()
    # the delta operand is used in TF32 (hi part only): ~2^-12 relative per product
    assert float((D[:F].double() - ref).abs().max()) / float(ref.abs().max()) <= 3e-4
