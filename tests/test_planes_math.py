"""CPU: the arithmetic and the address map of the bf16 hi/mid "planes" (attentive_dfprior_b200/csrc/dfp_planes.cuh),
restated in torch/numpy.  The CUDA side is covered by the parity tests; this pins the two claims DESIGN.md makes:
the three-product scheme is accurate to ~1e-5, and the layout is a bijection whose 8-sample / 8-feature chunks are
what the descriptors (LBO = 16 F, SBO = 128) address."""
import numpy as np
import torch


def split(x):
    hi = x.to(torch.bfloat16).to(torch.float32)
    mid = (x - hi).to(torch.bfloat16).to(torch.float32)
    return hi, mid


def test_three_product_scheme_accuracy():
    g = torch.Generator().manual_seed(0)
    A = torch.randn(4096, 64, generator=g)                 # samples x features (activations)
    B = torch.randn(4096, 96, generator=g) * 1e-3          # samples x delta columns
    ah, am = split(A)
    bh, bm = split(B)
    assert float((A - ah - am).abs().max() / A.abs().max()) < 2.0 ** -16
    ref = A.double().t() @ B.double()
    got = (am.double().t() @ bh.double()) + (ah.double().t() @ bm.double()) + (ah.double().t() @ bh.double())
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err
    one = ah.double().t() @ bh.double()                    # a single bf16 product for comparison: ~1e-3
    assert float((one - ref).abs().max() / ref.abs().max()) > 20 * err


def plane_byte(tile, plane, k, f, F):
    return (tile * 2 + plane) * 256 * F + (k // 8) * 16 * F + (f // 8) * 128 + (k % 8) * 16 + (f % 8) * 2


def test_plane_layout_is_a_bijection_and_chunked():
    for F in (40, 72, 104, 136, 160, 168):
        seen = np.zeros(2 * 2 * 128 * F, dtype=np.int32)
        for tile in range(2):
            for plane in range(2):
                k, f = np.meshgrid(np.arange(128), np.arange(F), indexing="ij")
                off = plane_byte(tile, plane, k, f, F)
                assert (off % 2 == 0).all()
                np.add.at(seen, off // 2, 1)
        assert (seen == 1).all()                           # every bf16 slot of the array is used exactly once
        # 8 consecutive features of one sample are one 16-byte chunk; 8 consecutive samples fill a 128-byte line
        assert plane_byte(0, 0, 5, 16, F) + 14 == plane_byte(0, 0, 5, 23, F)
        assert plane_byte(0, 0, 8 * 3 + 7, 8, F) - plane_byte(0, 0, 8 * 3, 8, F) == 7 * 16
        # operand addressing used by the weight-gradient kernel: feature chunks 128 B apart, sample groups 16 F apart
        assert plane_byte(0, 0, 0, 8, F) - plane_byte(0, 0, 0, 0, F) == 128
        assert plane_byte(0, 0, 8, 0, F) - plane_byte(0, 0, 0, 0, F) == 16 * F
        # a 32-sample slice of one plane (what one bulk copy moves) is contiguous: 64 F bytes
        lo, hi = plane_byte(1, 1, 32, 0, F), plane_byte(1, 1, 63, F - 1, F) + 2
        assert hi - lo == 64 * F
