"""GPU: tcgen05 / TMEM building blocks (dfp_tc.cuh) against an f64 matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(32, 32), (32, 64), (64, 160), (96, 64), (8, 16)])
def test_tc_selftest_3xtf32(K, N):
    from attentive_dfprior_b200 import _lib as L
    lib = L.lib()
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    ref = (A.double() @ W.double().t())
    for split, tol in ((1, 2e-6), (0, 5e-3)):
        D = torch.full((128, N), float("nan"), device="cuda")
        L.check(lib.dfp_selftest_tc(A.data_ptr(), W.data_ptr(), D.data_ptr(), K, N, split,
                                    torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        err = float((D.double() - ref).abs().max()) / float(ref.abs().max())
        assert err <= tol, (K, N, split, err)
        if split == 0:
            assert err > 1e-6, "single-pass TF32 should be visibly less accurate than the 3x split"


@pytest.mark.parametrize("stage", ["low", "high", "color"])
def test_tensor_core_and_simt_paths_agree(stage):
    """The tcgen05 (3xTF32) decoder path and the fp32 SIMT path give the same point query within 1e-5."""
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200 import _lib as L
    from attentive_dfprior_b200 import functional as Fn
    from attentive_dfprior_b200 import synth
    torch.manual_seed(1)
    scene = synth.make_scene("small", seed=5, device="cuda", high_std=1e-2, channels_last=True)
    model = A.DF().cuda()
    model.bound = scene.bound
    g = torch.Generator().manual_seed(3)
    lo, hi = scene.bound[:, 0], scene.bound[:, 1]
    p = (torch.rand(5000, 3, generator=g, dtype=torch.float64) * (hi - lo) + lo).cuda()
    lib = L.lib()
    try:
        with torch.no_grad():
            lib.dfp_set_tensor_cores(1)
            raw_tc, w_tc = Fn.eval_points(p, model, scene.grids, scene.tsdf_volume, scene.tsdf_bnds, scene.bound, stage)
            lib.dfp_set_tensor_cores(0)
            raw_sm, w_sm = Fn.eval_points(p, model, scene.grids, scene.tsdf_volume, scene.tsdf_bnds, scene.bound, stage)
    finally:
        lib.dfp_set_tensor_cores(1)
    scale = float(raw_sm[:, 3][raw_sm[:, 3] != 100].abs().max())
    assert float((raw_tc - raw_sm).abs().max()) <= 1e-5 * max(scale, 1.0)
    assert float((w_tc - w_sm).abs().max()) <= 1e-5


@pytest.mark.parametrize("J,KIN", [(32, 32), (64, 96), (32, 64)])
def test_tc_transposed_weight_view(J, KIN):
    """dH pattern: D = A[128,J] @ Wf[J,KIN] through the MN-major view of the canonical (forward) weight buffer."""
    from attentive_dfprior_b200 import _lib as L
    lib = L.lib()
    g = torch.Generator().manual_seed(J * 100 + KIN)
    A = torch.randn(128, J, generator=g).cuda()
    Wf = torch.randn(J, KIN, generator=g).cuda()
    D = torch.full((128, KIN), float("nan"), device="cuda")
    L.check(lib.dfp_selftest_tc(A.data_ptr(), Wf.data_ptr(), D.data_ptr(), J, KIN, 11, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = A.double() @ Wf.double()
    assert float((D.double() - ref).abs().max()) / float(ref.abs().max()) <= 2e-6


@pytest.mark.parametrize("F,N", [(32, 32), (32, 64), (4, 16), (32, 16)])
def test_tc_outer_product_from_smem(F, N):
    """dW pattern: D[:F] = H[128,F]^T @ Dl[128,N], both operands sample-major tiles in shared memory, M=128 over-read."""
    from attentive_dfprior_b200 import _lib as L
    lib = L.lib()
    g = torch.Generator().manual_seed(F * 100 + N)
    H = torch.randn(128, F, generator=g).cuda()
    Dl = torch.randn(128, N, generator=g).cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    L.check(lib.dfp_selftest_tc(H.data_ptr(), Dl.data_ptr(), D.data_ptr(), F, N, 12, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = H.double().t() @ Dl.double()
    # the delta operand is used in TF32 (hi part only): ~2^-12 relative per product
    assert float((D[:F].double() - ref).abs().max()) / float(ref.abs().max()) <= 3e-4
