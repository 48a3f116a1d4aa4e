"""Decoder modules with the reference's names, constructor arguments and ``state_dict`` keys
(src/conv_onet/models/decoder.py), whose forward passes run in the sm_100a kernels.

Only what the shipped configuration uses is provided (``pos_embedding_method='fourier'``,
hidden size 32, 5 blocks, skip after block 2: configs/df_prior.yaml:100-103 and
decoder.py:276-292); other embedding choices raise instead of silently differing.

The 79 parameters stay ordinary ``nn.Parameter``s, so ``.to()``, ``.share_memory()``,
``copy.deepcopy``, ``parameters()`` grouping by submodule (src/Mapper.py:364-371) and
``state_dict()/load_state_dict()`` (src/utils/Logger.py:26, src/DF_Prior.py:205-218) behave
as upstream.  Each call gathers them into the kernels' packed layout with one launch.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib as L
from . import functional as Fn


class GaussianFourierFeatureTransform(nn.Module):
    """sin(x @ _B), _B ~ N(0, scale^2) learnable -- decoder.py:7-30."""

    def __init__(self, num_input_channels, mapping_size=93, scale=25, learnable=True):
        super().__init__()
        B = torch.randn((num_input_channels, mapping_size)) * scale
        if learnable:
            self._B = nn.Parameter(B)
        else:
            raise NotImplementedError("only the learnable Fourier embedding is used by the shipped configs")

    def forward(self, x):
        x = x.squeeze(0)
        assert x.dim() == 2, "Expected 2D input (got {}D input)".format(x.dim())
        return torch.sin(x @ self._B.to(x.device))


class DenseLayer(nn.Linear):
    """Linear layer with xavier-uniform weights for the given activation gain -- decoder.py:70-79."""

    def __init__(self, in_dim: int, out_dim: int, activation: str = "relu", *args, **kwargs) -> None:
        self.activation = activation
        super().__init__(in_dim, out_dim, *args, **kwargs)

    def reset_parameters(self) -> None:
        nn.init.xavier_uniform_(self.weight, gain=nn.init.calculate_gain(self.activation))
        if self.bias is not None:
            nn.init.zeros_(self.bias)


class MLP(nn.Module):
    """One grid-conditioned decoder -- decoder.py:91-203.  ``forward(p, c_grid)`` evaluates it alone."""

    def __init__(self, name="", dim=3, c_dim=128, hidden_size=256, n_blocks=5, leaky=False,
                 sample_mode="bilinear", color=False, skips=[2], grid_len=0.16,
                 pos_embedding_method="fourier", concat_feature=False):
        super().__init__()
        if pos_embedding_method != "fourier" or leaky or sample_mode != "bilinear":
            raise NotImplementedError("kernels implement the shipped configuration only (fourier, relu, bilinear)")
        if hidden_size != L_HIDDEN or n_blocks != 5 or list(skips) != [2] or dim != 3:
            raise NotImplementedError("kernels implement hidden_size=32, n_blocks=5, skips=[2], dim=3")
        self.name = name
        self.color = color
        self.no_grad_feature = False
        self.c_dim = c_dim
        self.grid_len = grid_len
        self.concat_feature = concat_feature
        self.n_blocks = n_blocks
        self.skips = skips
        self.sample_mode = sample_mode
        embedding_size = 93
        self.fc_c = nn.ModuleList([nn.Linear(c_dim, hidden_size) for _ in range(n_blocks)])
        self.embedder = GaussianFourierFeatureTransform(dim, mapping_size=embedding_size, scale=25)
        self.pts_linears = nn.ModuleList(
            [DenseLayer(embedding_size, hidden_size, activation="relu")] +
            [DenseLayer(hidden_size + (embedding_size if i in skips else 0), hidden_size, activation="relu")
             for i in range(n_blocks - 1)])
        self.output_linear = DenseLayer(hidden_size, 4 if color else 1, activation="linear")

    def forward(self, p, c_grid=None):
        raise NotImplementedError(
            "a single decoder is evaluated through DF.forward / Renderer; call DF.forward(stage=...) instead")


L_HIDDEN = 32


class mlp_tsdf(nn.Module):
    """Attention MLP 2-64-128-128-64-2 blending learned occupancy with the TSDF prior -- decoder.py:206-258."""

    def __init__(self):
        super().__init__()
        self.no_grad_feature = False
        self.sample_mode = "bilinear"
        dims = [2, 64, 128, 128, 64]
        self.pts_linears = nn.ModuleList([DenseLayer(dims[i], dims[i + 1], activation="relu") for i in range(4)])
        self.output_linear = DenseLayer(64, 2, activation="linear")
        self.softmax = nn.Softmax(dim=1)
        self.sigmoid = nn.Sigmoid()

    def forward(self, p, occ, tsdf_volume, tsdf_bnds, **kwargs):
        raise NotImplementedError("the fusion MLP is evaluated inside DF.forward / Renderer (fused kernels)")


class ParamPack:
    """Gathers a decoder module's 79 parameters into the kernels' packed layout.

    Every call packs afresh into its OWN buffer (one small kernel, 344 KB from the caching allocator).  Round 1 cached
    the packed copy keyed on ``(data_ptr, _version)``: writes through ``param.data`` / raw pointers (``p.data.copy_``,
    an external fused optimiser) do not bump ``_version`` and kept stale weights in use, and one shared buffer meant a
    second forward after an optimiser step overwrote the weights an earlier, not-yet-run backward still pointed at.
    A private copy per forward has neither problem: its backward sees exactly the weights its forward used."""

    def get(self, tensors):
        lib = L.lib()
        dev = tensors[0].device
        packed = torch.zeros(int(lib.dfp_packed_param_floats()), device=dev, dtype=torch.float32)   # padding stays zero
        keep = [t.detach() if (t.dtype == torch.float32 and t.is_contiguous()) else t.detach().float().contiguous()
                for t in tensors]
        arr = (C.c_void_p * L.N_PARAM_TENSORS)(*[t.data_ptr() for t in keep])
        with torch.cuda.device(dev):
            L.check(lib.dfp_pack_params(C.cast(arr, C.POINTER(C.c_void_p)), packed.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream))
        return packed


_PACK = ParamPack()


def decoder_param_list(decoders):
    """The 79 tensors in state_dict order; works for this package's DF and for the reference's own DF."""
    ps = list(decoders.parameters())
    if len(ps) != L.N_PARAM_TENSORS or (not decoders.__dict__.get("_dfp_checked") and sum(p.numel() for p in ps) != L.N_PARAMS):
        raise ValueError(f"expected {L.N_PARAM_TENSORS} decoder tensors / {L.N_PARAMS} parameters, "
                         f"got {len(ps)} / {sum(p.numel() for p in ps)}")
    decoders.__dict__["_dfp_checked"] = True          # shapes of a module's parameters do not change: count them once
    return ps


def param_pack_of(decoders) -> ParamPack:
    return _PACK


class DF(nn.Module):
    """The three decoders + the fusion MLP -- decoder.py:262-353.

    ``forward(p, c_grid, tsdf_volume, tsdf_bnds, stage)`` returns ``(raw [P,4], w [P])`` exactly as
    upstream; ``.bound`` (set by the caller, src/DF_Prior.py:191-194) supplies the scene bound."""

    def __init__(self, dim=3, c_dim=32, low_grid_len=0.16, high_grid_len=0.16, color_grid_len=0.16,
                 hidden_size=32, pos_embedding_method="fourier"):
        super().__init__()
        if c_dim != 32:
            raise NotImplementedError("kernels implement c_dim=32 (configs/df_prior.yaml:102)")
        self.low_decoder = MLP(name="low", dim=dim, c_dim=c_dim, color=False, skips=[2], n_blocks=5,
                               hidden_size=hidden_size, grid_len=low_grid_len,
                               pos_embedding_method=pos_embedding_method)
        self.high_decoder = MLP(name="high", dim=dim, c_dim=c_dim * 2, color=False, skips=[2], n_blocks=5,
                                hidden_size=hidden_size, grid_len=high_grid_len, concat_feature=True,
                                pos_embedding_method=pos_embedding_method)
        self.color_decoder = MLP(name="color", dim=dim, c_dim=c_dim, color=True, skips=[2], n_blocks=5,
                                 hidden_size=hidden_size, grid_len=color_grid_len,
                                 pos_embedding_method=pos_embedding_method)
        self.mlp = mlp_tsdf()

    def __deepcopy__(self, memo):                 # Tracker deep-copies the shared decoders (src/Tracker.py:144)
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        import copy
        for k, v in self.__dict__.items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def forward(self, p, c_grid, tsdf_volume, tsdf_bnds, stage="low", **kwargs):
        bound = getattr(self, "bound", None)
        if bound is None:
            raise AttributeError("DF.bound must be set before calling forward (src/DF_Prior.py:191-194)")
        pts = p.reshape(-1, 3)
        raw, w = Fn.eval_points(pts, self, c_grid, tsdf_volume, tsdf_bnds, bound, stage, apply_bound_mask=False)
        return raw, w


def get_model(cfg):
    """Factory with the reference's contract -- src/conv_onet/config.py:4-27 (hidden size left at 32)."""
    return DF(dim=cfg["data"]["dim"], c_dim=cfg["model"]["c_dim"],
              low_grid_len=cfg["grid_len"]["low"], high_grid_len=cfg["grid_len"]["high"],
              color_grid_len=cfg["grid_len"]["color"],
              pos_embedding_method=cfg["model"]["pos_embedding_method"])
