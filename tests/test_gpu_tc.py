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
