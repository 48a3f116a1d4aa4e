"""Offline TSDF fusion with the reference's ``TSDFVolume`` interface (src/fusion.py:22-332), the per-frame ``integrate``
running as one sm_100a kernel (dfp_tsdf_integrate) on device-resident volumes.

Covered: the constructor's bound / dimension rules (:33-56), ``integrate`` (:189-290; upstream's own kernel is the pycuda
source at :69-142) and ``get_volume``; plus ``tsdf_prior()``, the [1,1,Z,Y,X] view + bounds that get_tsdf.py:95-97 saves
and src/DF_Prior.py:75-91 loads.  Mesh / point-cloud extraction (:295-332) is skimage marching cubes upstream and is
not part of this package.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


class TSDFVolume:
    def __init__(self, vol_bnds, voxel_size, use_gpu=True, device="cuda:0"):
        vol_bnds = np.array(vol_bnds, dtype=np.float64)
        assert vol_bnds.shape == (3, 2), "[!] `vol_bnds` should be of shape (3, 2)."
        self._vol_bnds = vol_bnds
        self._voxel_size = float(voxel_size)
        self._trunc_margin = 5 * self._voxel_size                       # fusion.py:38
        self._color_const = 256 * 256
        self._vol_dim = np.ceil((vol_bnds[:, 1] - vol_bnds[:, 0]) / self._voxel_size).copy(order="C").astype(int)   # :42
        self._vol_bnds[:, 1] = self._vol_bnds[:, 0] + self._vol_dim * self._voxel_size                              # :43
        self._vol_origin = self._vol_bnds[:, 0].copy(order="C").astype(np.float32)                                  # :44
        if not use_gpu:
            raise L.DfpError("attentive_dfprior_b200.fusion runs on the GPU only (there is no CPU fallback)")
        self.device = torch.device(device)
        dims = tuple(int(d) for d in self._vol_dim)
        self._tsdf_vol = -torch.ones(dims, dtype=torch.float32, device=self.device)        # :52
        self._weight_vol = torch.zeros(dims, dtype=torch.float32, device=self.device)
        self._color_vol = torch.zeros(dims, dtype=torch.float32, device=self.device)

    def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight=1.):
        """Integrate one RGB-D frame: color_im (H,W,3) in 0..255, depth_im (H,W), cam_intr (3,3), cam_pose (4,4) camera-to-world."""
        dev = self.device
        color = torch.as_tensor(np.asarray(color_im) if not torch.is_tensor(color_im) else color_im).to(dev).to(torch.float32)
        depth = torch.as_tensor(np.asarray(depth_im) if not torch.is_tensor(depth_im) else depth_im).to(dev).to(torch.float32).contiguous()
        H, W = depth.shape
        folded = torch.floor(color[..., 2] * self._color_const + color[..., 1] * 256 + color[..., 0]).contiguous()   # :207-208 (f32)
        f = L.DfpTsdfFuse()
        f.tsdf_vol, f.weight_vol, f.color_vol = self._tsdf_vol.data_ptr(), self._weight_vol.data_ptr(), self._color_vol.data_ptr()
        for k in range(3):
            f.dim[k] = int(self._vol_dim[k])
            f.origin[k] = float(self._vol_origin[k])
        f.voxel_size = float(np.float32(self._voxel_size))
        f.trunc_margin = float(np.float32(self._trunc_margin))
        f.depth_im, f.color_im, f.H, f.W = depth.data_ptr(), folded.data_ptr(), H, W
        K = np.asarray(cam_intr.cpu() if torch.is_tensor(cam_intr) else cam_intr, dtype=np.float64).reshape(-1).astype(np.float32)
        P = np.asarray(cam_pose.cpu() if torch.is_tensor(cam_pose) else cam_pose, dtype=np.float64).reshape(-1).astype(np.float32)
        f.intr[:] = [float(v) for v in K]
        f.pose[:] = [float(v) for v in P]
        f.obs_weight = float(np.float32(obs_weight))
        with torch.cuda.device(dev):
            L.check(L.lib().dfp_tsdf_integrate(C.byref(f), torch.cuda.current_stream().cuda_stream))

    def get_volume(self):
        return self._tsdf_vol.cpu().numpy(), self._color_vol.cpu().numpy(), self._vol_bnds

    def tsdf_prior(self):
        """(tsdf_volume [1,1,Z,Y,X] non-contiguous view, bounds [3,2] f64 tensor on the device): what get_tsdf.py:95-97 stores."""
        X, Y, Z = self._tsdf_vol.shape
        vol = self._tsdf_vol.reshape(1, 1, X, Y, Z).permute(0, 1, 4, 3, 2)
        return vol, torch.from_numpy(self._vol_bnds.copy()).to(self.device)
