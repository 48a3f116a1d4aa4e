"""CPU: checkpoints / TSDF prior files in the reference's formats (src/utils/Logger.py:24-34, src/DF_Prior.py:75-91,
205-218).  When /root/reference is mounted the checkpoint is produced from the reference's own DF module."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _grids():
    g = torch.Generator().manual_seed(0)
    return {"grid_low": torch.randn(1, 32, 3, 4, 5, generator=g), "grid_high": torch.randn(1, 32, 6, 7, 9, generator=g),
            "grid_color": torch.randn(1, 32, 6, 7, 9, generator=g)}


def test_checkpoint_round_trip(tmp_path):
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200 import io as IO
    torch.manual_seed(1)
    src, dst = A.DF(), A.DF()
    c = {k: v.contiguous(memory_format=torch.channels_last_3d) for k, v in _grids().items()}
    tsdf = torch.rand(5, 4, 3).reshape(1, 1, 5, 4, 3).permute(0, 1, 4, 3, 2)           # non-contiguous, as get_tsdf.py leaves it
    path = IO.save_checkpoint(str(tmp_path / "00010.tar"), c, src, torch.eye(4)[None], torch.eye(4)[None], [0, 5],
                              [{"idx": 0}], 10, tsdf)
    raw = torch.load(path, weights_only=False)
    assert tuple(raw.keys()) == IO.CKPT_KEYS                                            # Logger.log's keys, in its order
    ck = IO.load_checkpoint(path, dst)
    for (k1, v1), (k2, v2) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    for k, v in _grids().items():
        assert torch.equal(ck["c"][k], v) and ck["c"][k].is_contiguous(memory_format=torch.channels_last_3d)
    assert ck["idx"] == 10 and ck["keyframe_list"] == [0, 5]
    assert torch.equal(ck["tsdf_volume"], tsdf)


def test_reads_a_checkpoint_of_the_reference_module(tmp_path):
    ref_root = "/root/reference"
    if not os.path.isdir(ref_root):
        pytest.skip("reference checkout not mounted")
    sys.path.insert(0, ref_root)
    cwd = os.getcwd()
    try:
        os.chdir(ref_root)
        from src.conv_onet.models.decoder import DF as RefDF
        torch.manual_seed(3)
        ref = RefDF(dim=3, c_dim=32, low_grid_len=0.32, high_grid_len=0.16, color_grid_len=0.16, hidden_size=32,
                    pos_embedding_method="fourier")
    finally:
        os.chdir(cwd)
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200 import io as IO
    path = str(tmp_path / "ref.tar")
    torch.save({"c": _grids(), "decoder_state_dict": ref.state_dict(), "gt_c2w_list": None, "estimate_c2w_list": None,
                "keyframe_list": [], "keyframe_dict": [], "selected_keyframes": None, "idx": 0, "tsdf_volume": None},
               path, _use_new_zipfile_serialization=False)                             # exactly what Logger.log writes
    mine = A.DF()
    IO.load_checkpoint(path, mine)
    sd = mine.state_dict()
    assert list(sd.keys()) == list(ref.state_dict().keys()) and len(sd) == 79
    assert all(torch.equal(sd[k], v) for k, v in ref.state_dict().items())


def test_tsdf_prior_and_pretrained_decoders(tmp_path):
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200 import io as IO
    vol = torch.rand(6, 5, 4).reshape(1, 1, 6, 5, 4).permute(0, 1, 4, 3, 2)
    torch.save(vol, str(tmp_path / "room0_tsdf_volume.pt"))
    torch.save(np.array([[-1.0, 2.0], [-1.5, 1.5], [0.0, 3.0]]), str(tmp_path / "room0_bounds.pt"))
    v, b = IO.load_tsdf_prior(str(tmp_path / "room0_tsdf_volume.pt"), str(tmp_path / "room0_bounds.pt"))
    assert torch.equal(v, vol) and b.dtype == torch.float64 and tuple(b.shape) == (3, 2)
    torch.manual_seed(5)
    donor, target = A.DF(), A.DF()
    model = {"decoder_coarse." + k: v for k, v in donor.low_decoder.state_dict().items()}
    model.update({"decoder_fine." + k: v for k, v in donor.high_decoder.state_dict().items()})
    model["encoder.conv.weight"] = torch.zeros(1)
    torch.save({"model": model}, str(tmp_path / "low_high.pt"))
    IO.load_pretrained_decoders(target, str(tmp_path / "low_high.pt"))
    for name in ("low_decoder", "high_decoder"):
        for (k1, v1), (k2, v2) in zip(getattr(donor, name).state_dict().items(), getattr(target, name).state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2)
