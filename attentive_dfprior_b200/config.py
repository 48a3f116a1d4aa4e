"""The hot path's configuration knobs with the reference's defaults (configs/df_prior.yaml)."""
from __future__ import annotations

import copy
from types import SimpleNamespace

DEFAULT_CFG = {
    "scale": 1,                                   # df_prior.yaml:2
    "occupancy": True,                            # :4
    "grid_len": {"low": 0.32, "high": 0.16, "color": 0.16, "bound_divisible": 0.32},   # :6-10
    "meshing": {"resolution": 256},               # :15
    "rendering": {"N_samples": 32, "N_surface": 16, "N_importance": 0, "lindisp": False, "perturb": 0.0},  # :93-98
    "data": {"dim": 3},                           # :99-100
    "model": {"c_dim": 32, "pos_embedding_method": "fourier"},   # :101-103
    "cam": {"H": 680, "W": 1200, "fx": 600.0, "fy": 600.0, "cx": 599.5, "cy": 339.5},   # :84-90
    "mapping": {"pixels": 1000, "w_color_loss": 0.2}, "tracking": {"pixels": 200, "w_color_loss": 0.5},
}


def default_cfg():
    return copy.deepcopy(DEFAULT_CFG)


def slam_namespace(bound, tsdf_bnds, cam=None):
    """The attributes Renderer/Mesher read from the SLAM object (src/utils/Renderer.py:19-23)."""
    cam = cam or DEFAULT_CFG["cam"]
    return SimpleNamespace(bound=bound, vol_bnds=tsdf_bnds, tsdf_bnds=tsdf_bnds, H=cam["H"], W=cam["W"], fx=cam["fx"], fy=cam["fy"],
                           cx=cam["cx"], cy=cam["cy"])
