"""Ray generation and pose utilities around the hot path (src/common.py).

Where the Tracker's pose gradient enters (camera 7-vector -> c2w -> rays_o/rays_d) they are PyTorch ops, so autograd
sees them: the kernels return d rays_o / d rays_d and autograd carries those into the quaternion and translation.
Without gradient (render_img's full frames, a Mapper's fixed poses) and with the pose on the GPU, ``get_rays`` and
``get_rays_from_uv`` run as one kernel (dfp_gen_rays) that reproduces the torch ops' f32 arithmetic bit for bit.
"""
from __future__ import annotations

import numpy as np
import torch


def _kernel_rays(n, W, pix_i, pix_j, fx, fy, cx, cy, c2w):
    import ctypes as C  # noqa: F401
    from . import _lib as L
    pose = c2w.detach()
    if pose.dtype != torch.float32 or pose.stride(-1) != 1:
        pose = pose.float().contiguous()
    rays_o = torch.empty((n, 3), dtype=torch.float32, device=pose.device)
    rays_d = torch.empty((n, 3), dtype=torch.float32, device=pose.device)
    with torch.cuda.device(pose.device):
        L.check(L.lib().dfp_gen_rays(n, W, None if pix_i is None else pix_i.data_ptr(), None if pix_j is None else pix_j.data_ptr(),
                                     float(fx), float(fy), float(cx), float(cy), pose.data_ptr(), pose.stride(0),
                                     rays_o.data_ptr(), rays_d.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return rays_o, rays_d


def _kernel_ok(c2w, *more):
    return (torch.is_tensor(c2w) and c2w.is_cuda and not (torch.is_grad_enabled() and c2w.requires_grad)
            and c2w.dim() == 2 and c2w.shape[0] >= 3 and c2w.shape[1] == 4
            and all(torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and not t.requires_grad for t in more))


def get_rays_from_uv(i, j, c2w, H, W, fx, fy, cx, cy, device):
    """Rays of chosen pixels -- src/common.py:76-91."""
    if isinstance(c2w, np.ndarray):
        c2w = torch.from_numpy(c2w).to(device)
    if _kernel_ok(c2w, i, j):
        n = i.numel()
        rays_o, rays_d = _kernel_rays(n, W, i.reshape(-1).contiguous(), j.reshape(-1).contiguous(), fx, fy, cx, cy, c2w)
        return rays_o.reshape(*i.shape, 3), rays_d.reshape(*i.shape, 3)
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1).to(device).reshape(-1, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def select_uv(i, j, n, depth, color, device="cuda:0"):
    """n pixels with replacement -- src/common.py:94-109."""
    i, j = i.reshape(-1), j.reshape(-1)
    idx = torch.randint(i.shape[0], (n,), device=device).clamp(0, i.shape[0])
    return i[idx], j[idx], depth.reshape(-1)[idx], color.reshape(-1, 3)[idx]


def get_sample_uv(H0, H1, W0, W1, n, depth, color, device="cuda:0"):
    """src/common.py:112-124."""
    depth = depth[H0:H1, W0:W1]
    color = color[H0:H1, W0:W1]
    i, j = torch.meshgrid(torch.linspace(W0, W1 - 1, W1 - W0).to(device),
                          torch.linspace(H0, H1 - 1, H1 - H0).to(device), indexing="ij")
    return select_uv(i.t(), j.t(), n, depth, color, device=device)


def get_samples(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, depth, color, device):
    """n rays from an image window with their depth/colour -- src/common.py:127-136."""
    i, j, sample_depth, sample_color = get_sample_uv(H0, H1, W0, W1, n, depth, color, device=device)
    rays_o, rays_d = get_rays_from_uv(i, j, c2w, H, W, fx, fy, cx, cy, device)
    return rays_o, rays_d, sample_depth, sample_color


def get_rays(H, W, fx, fy, cx, cy, c2w, device):
    """All rays of a frame -- src/common.py:254-272."""
    if isinstance(c2w, np.ndarray):
        c2w = torch.from_numpy(c2w)
    if _kernel_ok(c2w):
        rays_o, rays_d = _kernel_rays(H * W, W, None, None, fx, fy, cx, cy, c2w)
        return rays_o.reshape(H, W, 3), rays_d.reshape(H, W, 3)
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1).to(device).reshape(H, W, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3].to(device), -1)
    rays_o = c2w[:3, -1].to(device).expand(rays_d.shape)
    return rays_o, rays_d


def quad2rotation(quad):
    """Quaternion (r,i,j,k) batch -> rotation matrices, differentiable -- src/common.py:139-162."""
    qr, qi, qj, qk = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    two_s = 2.0 / (quad * quad).sum(-1)
    r0 = torch.stack([1 - two_s * (qj ** 2 + qk ** 2), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr)], -1)
    r1 = torch.stack([two_s * (qi * qj + qk * qr), 1 - two_s * (qi ** 2 + qk ** 2), two_s * (qj * qk - qi * qr)], -1)
    r2 = torch.stack([two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi ** 2 + qj ** 2)], -1)
    return torch.stack([r0, r1, r2], 1)


def get_camera_from_tensor(inputs):
    """(quaternion, translation) 7-vector -> [3,4] -- src/common.py:165-178."""
    single = inputs.dim() == 1
    x = inputs.unsqueeze(0) if single else inputs
    RT = torch.cat([quad2rotation(x[:, :4]), x[:, 4:, None]], 2)
    return RT[0] if single else RT


def normalize_3d_coordinate(p, bound):
    """((p-b0)/(b1-b0))*2-1 -- src/common.py:275-290 (host-side helper; the kernels do this themselves)."""
    p = p.reshape(-1, 3)
    for k in range(3):
        p[:, k] = ((p[:, k] - bound[k, 0]) / (bound[k, 1] - bound[k, 0])) * 2 - 1.0
    return p
