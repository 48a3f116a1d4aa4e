"""On-disk formats the reference's run produces and consumes, for this implementation's state (SURVEY.md 8(f) row 4,
formats part).  Pure host code, no kernels.

* checkpoints: `Logger.log` (src/utils/Logger.py:24-34) torch.saves a dict with keys ``c`` (the grid dict),
  ``decoder_state_dict`` (79 tensors, same names here), ``gt_c2w_list``, ``estimate_c2w_list``, ``keyframe_list``,
  ``keyframe_dict``, ``selected_keyframes``, ``idx``, ``tsdf_volume`` -- with the legacy (non-zipfile) serialisation.
* TSDF prior: a ``*_tsdf_volume.pt`` tensor ``[1,1,Z,Y,X]`` (a non-contiguous permuted view, get_tsdf.py:95-97) and a
  ``*_bounds.pt`` array ``[3,2]`` (src/DF_Prior.py:75-91).
* pretrained decoders: ConvONet checkpoint whose ``model`` dict holds ``decoder_coarse.*`` / ``decoder_fine.*``
  entries (DF_Prior.load_pretrain, src/DF_Prior.py:205-218).
"""
from __future__ import annotations

import os

import torch

CKPT_KEYS = ("c", "decoder_state_dict", "gt_c2w_list", "estimate_c2w_list", "keyframe_list", "keyframe_dict",
             "selected_keyframes", "idx", "tsdf_volume")


def save_checkpoint(path, c, decoders, gt_c2w_list, estimate_c2w_list, keyframe_list, keyframe_dict, idx,
                    tsdf_volume, selected_keyframes=None):
    """Write a checkpoint the reference's tools can read (same keys, legacy serialisation).  Channels-last grids are
    stored as they are: the logical shape ``[1,C,D,H,W]`` is what every reader indexes."""
    torch.save({
        "c": c,
        "decoder_state_dict": decoders.state_dict(),
        "gt_c2w_list": gt_c2w_list,
        "estimate_c2w_list": estimate_c2w_list,
        "keyframe_list": keyframe_list,
        "keyframe_dict": keyframe_dict,
        "selected_keyframes": selected_keyframes,
        "idx": idx,
        "tsdf_volume": tsdf_volume,
    }, path, _use_new_zipfile_serialization=False)
    return path


def load_checkpoint(path, decoders=None, device="cpu", channels_last: bool = True):
    """Read a checkpoint written by the reference's Logger or by ``save_checkpoint``.  Loads the decoder weights into
    ``decoders`` (strict: the 79 names must match) and returns the dict; the grids are moved to ``device`` and, by
    default, laid out channels-last so the render kernels use them in place."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    missing = [k for k in ("c", "decoder_state_dict") if k not in ckpt]
    if missing:
        raise KeyError(f"{path}: not a DF-Prior checkpoint (missing {missing})")
    if decoders is not None:
        decoders.load_state_dict(ckpt["decoder_state_dict"], strict=True)
    grids = {}
    for key, val in ckpt["c"].items():
        val = val.to(device)
        if channels_last and val.dim() == 5:
            val = val.contiguous(memory_format=torch.channels_last_3d)
        grids[key] = val
    ckpt["c"] = grids
    return ckpt


def load_tsdf_prior(volume_path, bounds_path, device="cpu"):
    """The fused TSDF prior and its bounds as src/DF_Prior.py:75-91 loads them: the volume keeps whatever strides it was
    saved with (the kernels take explicit strides), the bounds become an f64 ``[3,2]`` tensor."""
    vol = torch.load(volume_path, map_location="cpu", weights_only=False)
    bnds = torch.load(bounds_path, map_location="cpu", weights_only=False)
    vol = vol.to(device)
    bnds = torch.as_tensor(bnds, dtype=torch.float64).to(device)
    if vol.dim() != 5 or tuple(bnds.shape) != (3, 2):
        raise ValueError(f"unexpected TSDF prior shapes: volume {tuple(vol.shape)}, bounds {tuple(bnds.shape)}")
    return vol, bnds


def load_pretrained_decoders(decoders, path, device="cpu"):
    """DF_Prior.load_pretrain (src/DF_Prior.py:205-218): take ``decoder_coarse.*`` / ``decoder_fine.*`` entries of a
    ConvONet checkpoint's ``model`` dict into the low / high decoders."""
    ckpt = torch.load(path, map_location=device, weights_only=False)
    low, high = {}, {}
    for key, val in ckpt["model"].items():
        if "decoder" in key and "encoder" not in key:
            if "coarse" in key:
                low[key[len("decoder_coarse."):]] = val
            elif "fine" in key:
                high[key[len("decoder_fine."):]] = val
    decoders.low_decoder.load_state_dict(low)
    decoders.high_decoder.load_state_dict(high)
    return decoders
