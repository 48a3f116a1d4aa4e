"""CPU oracle for the DF-Prior per-ray volume-rendering hot path.

TEST INFRASTRUCTURE ONLY.  This file is a functional restatement, in plain
PyTorch ops, of what the reference computes on the path
``Renderer.render_batch_ray`` -> ``eval_points`` -> ``DF.forward`` ->
``raw2outputs_nerf_color``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package ``attentive_dfprior_b200`` never does.

Parity pinning: the reference ships no golden vectors or tests for this path
("parity unpinned by upstream tests").  The oracle is
therefore pinned against the *reference itself*, imported from
``/root/reference`` in the authoring container by ``oracle/make_golden.py``;
the resulting input/output vectors are committed under ``tests/golden/`` and
``tests/test_oracle_golden.py`` checks this file against them.

Every function cites the reference lines it restates (paths relative to the
reference checkout).  Decoder parameters are passed as a flat ``dict`` using
the reference's ``state_dict()`` names (79 tensors).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

N_BLOCKS = 5          # src/conv_onet/models/decoder.py:283 (n_blocks=5)
SKIP_AFTER = 2        # decoder.py:283 (skips=[2])
HIDDEN = 32           # decoder.py:278 (hidden_size default; the factory does not pass one)
EMBED = 93            # decoder.py:129
STAGES = ("low", "high", "color")


# --------------------------------------------------------------------------
# scene set-up rules (host glue whose results define the kernel's inputs)
# --------------------------------------------------------------------------
def round_bound(bound, divisible: float = 0.32, scale: float = 1.0) -> torch.Tensor:
    """Upper-edge rounding of the scene bound -- src/DF_Prior.py:185-190."""
    b = torch.from_numpy(np.array(bound, dtype=np.float64) * scale)
    b[:, 1] = (((b[:, 1] - b[:, 0]) / divisible).int() + 1) * divisible + b[:, 0]
    return b


def grid_shape(bound: torch.Tensor, grid_len: float):
    """[D, H, W] = [z, y, x] vertex counts of one feature grid -- src/DF_Prior.py:241-246."""
    xyz_len = bound[:, 1] - bound[:, 0]
    s = list(map(int, (xyz_len / grid_len).tolist()))
    s[0], s[2] = s[2], s[0]
    return s


def init_decoder_params(seed: int = 0, c_dim: int = 32, dtype=torch.float32) -> Params:
    """Random-init decoder parameters with the reference's names, shapes and
    distributions (xavier-uniform DenseLayer with relu/linear gain and zero
    bias, decoder.py:75-79; default nn.Linear init for fc_c, decoder.py:124-126;
    _B = randn*25, decoder.py:21-22).  The RNG stream is not the reference's."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}

    def xavier(out_d, in_d, gain):
        a = gain * math.sqrt(6.0 / (in_d + out_d))
        return (torch.rand(out_d, in_d, generator=g, dtype=torch.float64) * 2 - 1) * a

    def default_linear(out_d, in_d):
        k = 1.0 / math.sqrt(in_d)
        w = (torch.rand(out_d, in_d, generator=g, dtype=torch.float64) * 2 - 1) * k
        b = (torch.rand(out_d, generator=g, dtype=torch.float64) * 2 - 1) * k
        return w, b

    relu_gain = math.sqrt(2.0)
    for name, cd, n_out in (("low", c_dim, 1), ("high", 2 * c_dim, 1), ("color", c_dim, 4)):
        pre = f"{name}_decoder."
        for i in range(N_BLOCKS):
            w, b = default_linear(HIDDEN, cd)
            p[pre + f"fc_c.{i}.weight"], p[pre + f"fc_c.{i}.bias"] = w, b
        p[pre + "embedder._B"] = torch.randn(3, EMBED, generator=g, dtype=torch.float64) * 25
        dims = [EMBED, HIDDEN, HIDDEN, HIDDEN + EMBED, HIDDEN]
        for i in range(N_BLOCKS):
            p[pre + f"pts_linears.{i}.weight"] = xavier(HIDDEN, dims[i], relu_gain)
            p[pre + f"pts_linears.{i}.bias"] = torch.zeros(HIDDEN, dtype=torch.float64)
        p[pre + "output_linear.weight"] = xavier(n_out, HIDDEN, 1.0)
        p[pre + "output_linear.bias"] = torch.zeros(n_out, dtype=torch.float64)
    dims = [2, 64, 128, 128, 64]
    for i in range(4):
        p[f"mlp.pts_linears.{i}.weight"] = xavier(dims[i + 1], dims[i], relu_gain)
        p[f"mlp.pts_linears.{i}.bias"] = torch.zeros(dims[i + 1], dtype=torch.float64)
    p["mlp.output_linear.weight"] = xavier(2, 64, 1.0)
    p["mlp.output_linear.bias"] = torch.zeros(2, dtype=torch.float64)
    return {k: v.to(dtype) for k, v in p.items()}


# --------------------------------------------------------------------------
# point evaluation
# --------------------------------------------------------------------------
def normalize_coords(p: torch.Tensor, bound: torch.Tensor) -> torch.Tensor:
    """((p-b0)/(b1-b0))*2-1 per axis, evaluated in p's dtype -- src/common.py:275-290.
    ``bound`` entries act as 0-dim scalars exactly as in the reference."""
    q = p.reshape(-1, 3).clone()
    for k in range(3):
        q[:, k] = ((q[:, k] - bound[k, 0]) / (bound[k, 1] - bound[k, 0])) * 2 - 1.0
    return q


def trilinear_gather(volume: torch.Tensor, p: torch.Tensor, bound: torch.Tensor) -> torch.Tensor:
    """Trilinear feature gather -- decoder.py:168-175 (feature grids) and
    :230-238 / :295-303 (TSDF).  volume [1,C,D,H,W]; returns [1,C,P]."""
    vgrid = normalize_coords(p, bound).unsqueeze(0)[:, :, None, None].float()
    if volume.dtype != torch.float32:           # f64 "truth" mode
        vgrid = vgrid.to(volume.dtype)
    return F.grid_sample(volume, vgrid.to(volume.device), padding_mode="border",
                         align_corners=True, mode="bilinear").squeeze(-1).squeeze(-1)


def fourier_embed(p32: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """sin(p @ _B) on world coordinates -- decoder.py:26-30."""
    return torch.sin(p32 @ B)


def decoder_mlp(params: Params, name: str, p: torch.Tensor, grids: Dict[str, torch.Tensor],
                bound: torch.Tensor) -> torch.Tensor:
    """One grid-conditioned decoder -- decoder.py:177-203.  p is [1,P,3] or [P,3].
    Returns [P] (low/high) or [P,4] (color)."""
    pre = f"{name}_decoder."
    wdt = params[pre + "embedder._B"].dtype
    c = trilinear_gather(grids["grid_" + name], p, bound).transpose(1, 2).squeeze(0)
    if name == "high":                                           # decoder.py:182-187
        with torch.no_grad():
            c_low = trilinear_gather(grids["grid_low"], p, bound).transpose(1, 2).squeeze(0)
        c = torch.cat([c, c_low], dim=1)
    pf = p.reshape(-1, 3).to(wdt)                                # decoder.py:189 (.float())
    emb = fourier_embed(pf, params[pre + "embedder._B"])
    h = emb
    for i in range(N_BLOCKS):
        h = F.linear(h, params[pre + f"pts_linears.{i}.weight"], params[pre + f"pts_linears.{i}.bias"])
        h = F.relu(h)
        h = h + F.linear(c, params[pre + f"fc_c.{i}.weight"], params[pre + f"fc_c.{i}.bias"])
        if i == SKIP_AFTER:
            h = torch.cat([emb, h], -1)
    out = F.linear(h, params[pre + "output_linear.weight"], params[pre + "output_linear.bias"])
    return out if name == "color" else out.squeeze(-1)


def relu_margin(params: Params, p: torch.Tensor, grids, tsdf_volume, tsdf_bnds, bound, stage: str = "color") -> torch.Tensor:
    """Conditioning of a point for GRADIENT comparisons: the smallest RELATIVE pre-activation |z| / (|W| |h| + |b|) over
    every ReLU the stage evaluates at p (decoders and, inside the TSDF band, the fusion MLP), computed in f64 from f32
    parameters.  The denominator is the magnitude of the terms the unit sums, i.e. the scale of any fp32 evaluation's
    rounding error.  The path's gradient is discontinuous where a pre-activation crosses zero, so two correct fp32
    implementations (relative error ~1e-6 of that scale) can differ by a whole unit's gradient at a point whose margin
    is that small; parity tests that put equal upstream weight on every point drop such points (test infrastructure,
    not a reference line)."""
    p = p.reshape(-1, 3).double()
    P64 = {k: v.detach().double() for k, v in params.items()}
    G64 = {k: v.detach().double() for k, v in grids.items()}
    margin = torch.full((p.shape[0],), float("inf"), dtype=torch.float64)
    outs = {}
    for name in {"low": ("low",), "high": ("high", "low"), "color": ("high", "color", "low")}[stage]:
        pre = f"{name}_decoder."
        c = trilinear_gather(G64["grid_" + name], p, bound).transpose(1, 2).squeeze(0)
        if name == "high":
            c = torch.cat([c, trilinear_gather(G64["grid_low"], p, bound).transpose(1, 2).squeeze(0)], dim=1)
        emb = fourier_embed(p.float().double(), P64[pre + "embedder._B"])
        h = emb
        for i in range(N_BLOCKS):
            Wm, bm = P64[pre + f"pts_linears.{i}.weight"], P64[pre + f"pts_linears.{i}.bias"]
            z = F.linear(h, Wm, bm)
            margin = torch.minimum(margin, (z.abs() / (F.linear(h.abs(), Wm.abs(), bm.abs()) + 1e-300)).min(dim=1)[0])
            h = F.relu(z) + F.linear(c, P64[pre + f"fc_c.{i}.weight"], P64[pre + f"fc_c.{i}.bias"])
            if i == SKIP_AFTER:
                h = torch.cat([emb, h], -1)
        outs[name] = F.linear(h, P64[pre + "output_linear.weight"], P64[pre + "output_linear.bias"])
    if stage != "low":
        tv = trilinear_gather(tsdf_volume.double(), p, tsdf_bnds).reshape(-1)
        m = band_mask(tv)
        if bool(m.any()):
            x = torch.stack([(outs["high"].squeeze(-1) + outs["low"].squeeze(-1))[m], tsdf_to_logit(tv[m])], dim=1)
            h = x
            mm = torch.full((int(m.sum()),), float("inf"), dtype=torch.float64)
            for i in range(4):
                Wm, bm = P64[f"mlp.pts_linears.{i}.weight"], P64[f"mlp.pts_linears.{i}.bias"]
                z = F.linear(h, Wm, bm)
                mm = torch.minimum(mm, (z.abs() / (F.linear(h.abs(), Wm.abs(), bm.abs()) + 1e-300)).min(dim=1)[0])
                h = F.relu(z)
            margin = margin.index_put((m,), torch.minimum(margin[m], mm))
    return margin


def tsdf_to_logit(tsdf_val: torch.Tensor) -> torch.Tensor:
    """TSDF value -> occupancy logit -- decoder.py:244-248."""
    t = 1. - (tsdf_val + 1.) / 2.
    t = torch.clamp(t, 0.0, 1.0)
    inv = -0.1 * torch.log((1 / (t + 1e-8)) - 1 + 1e-7)
    return torch.clamp(inv, -100.0, 100.0)


def attentive_fusion(params: Params, p: torch.Tensor, occ: torch.Tensor, tsdf_volume: torch.Tensor,
                     tsdf_bnds: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Attention MLP blending learned occupancy with the TSDF prior -- decoder.py:240-258.
    p [M,3], occ [M]; returns (fused [M], weight-on-TSDF-term [M])."""
    tsdf_val = trilinear_gather(tsdf_volume, p, tsdf_bnds).squeeze(0)          # [1,M]
    inv = tsdf_to_logit(tsdf_val)
    x = torch.cat([occ.reshape(tsdf_val.shape), inv], dim=0).t()               # [M,2]
    h = x
    for i in range(4):
        h = F.relu(F.linear(h, params[f"mlp.pts_linears.{i}.weight"], params[f"mlp.pts_linears.{i}.bias"]))
    a = torch.softmax(F.linear(h, params["mlp.output_linear.weight"], params["mlp.output_linear.bias"]), dim=1)
    return (a * x).sum(dim=1), a[:, 1]


def band_mask(tsdf_val: torch.Tensor) -> torch.Tensor:
    """Points inside the truncation band -- decoder.py:329, :346."""
    return (tsdf_val > -1.0 + 1e-4) & (tsdf_val < 1.0 - 1e-4)


def df_forward(params: Params, p: torch.Tensor, grids: Dict[str, torch.Tensor], tsdf_volume: torch.Tensor,
               tsdf_bnds: torch.Tensor, bound: torch.Tensor, stage: str = "low"
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stage dispatch -- decoder.py:307-353.  p [P,3] (or [1,P,3]).  Returns raw [P,4], w [P]."""
    p = p.reshape(-1, 3)
    P = p.shape[0]
    dev = p.device
    wdt = params["low_decoder.embedder._B"].dtype
    if stage == "low":
        low = decoder_mlp(params, "low", p, grids, bound)
        raw = torch.zeros(P, 4, device=dev, dtype=wdt)
        raw[:, 3] = low
        return raw, torch.ones(P, device=dev, dtype=wdt)
    if stage not in ("high", "color"):
        raise ValueError(stage)
    high = decoder_mlp(params, "high", p, grids, bound)
    raw = decoder_mlp(params, "color", p, grids, bound) if stage == "color" \
        else torch.zeros(P, 4, device=dev, dtype=wdt)
    low = decoder_mlp(params, "low", p, grids, bound)
    both = high + low                                                # decoder.py:326, :343
    mask = band_mask(trilinear_gather(tsdf_volume, p, tsdf_bnds)).reshape(-1)
    w = torch.ones(P, device=dev, dtype=wdt)
    occ = low.clone()
    if bool(mask.any()):
        fused, a1 = attentive_fusion(params, p[mask], both[mask], tsdf_volume, tsdf_bnds)
        occ = occ.index_put((mask,), fused)                           # decoder.py:333, :350
        w = w.index_put((mask,), a1)
    raw = torch.cat([raw[:, :3], occ[:, None]], dim=1)                # decoder.py:336, :352
    return raw, w


def in_bound_mask(p: torch.Tensor, bound: torch.Tensor) -> torch.Tensor:
    """Strict in-bound test in p's dtype -- src/utils/Renderer.py:51-54."""
    m = torch.ones(p.shape[0], dtype=torch.bool, device=p.device)
    for k in range(3):
        m = m & (p[:, k] < bound[k][1]) & (p[:, k] > bound[k][0])
    return m


def eval_points(params: Params, p: torch.Tensor, grids, tsdf_volume, tsdf_bnds, bound, stage="color",
                points_batch_size: int = 500000) -> Tuple[torch.Tensor, torch.Tensor]:
    """Batched point query with the out-of-bound occupancy override --
    Renderer.py:27-71 (and its twin src/utils/Mesher.py:286-326)."""
    rets, ws = [], []
    for pi in torch.split(p, points_batch_size):
        m = in_bound_mask(pi, bound)
        raw, w = df_forward(params, pi, grids, tsdf_volume, tsdf_bnds, bound, stage)
        occ = torch.where(m, raw[:, 3], torch.full_like(raw[:, 3], 100.0))   # Renderer.py:64
        rets.append(torch.cat([raw[:, :3], occ[:, None]], dim=1))
        ws.append(w)
    return torch.cat(rets, 0), torch.cat(ws, 0)


# --------------------------------------------------------------------------
# ray sampling and compositing
# --------------------------------------------------------------------------
def sample_depths(rays_o: torch.Tensor, rays_d: torch.Tensor, gt_depth: Optional[torch.Tensor],
                  bound: torch.Tensor, n_samples: int = 32, n_surface: int = 16) -> torch.Tensor:
    """Sorted per-ray sample depths z [N, n_samples+n_surface] (f64) --
    Renderer.py:134-221 with lindisp=False, perturb=0 (configs/df_prior.yaml:93-98)."""
    dev = rays_o.device
    with torch.no_grad():
        o = rays_o.detach().unsqueeze(-1)
        d = rays_d.detach().unsqueeze(-1)
        t = (bound.unsqueeze(0).to(dev) - o) / d                               # :151-152 (f64)
        far_bb = torch.min(torch.max(t, dim=2)[0], dim=1)[0].unsqueeze(-1)     # :153-154
        far_bb = far_bb + 0.01                                                 # :155
    if gt_depth is None:                                                       # :140-142, :161-162
        n_surface = 0
        near = 0.01
        far = far_bb
    else:
        gt_depth = gt_depth.reshape(-1, 1)
        near = gt_depth.repeat(1, n_samples) * 0.01                            # :145-146 (f32)
        far = torch.clamp(far_bb, 0, torch.max(gt_depth * 1.2))                # :159 batch-global
    t_vals = torch.linspace(0., 1., steps=n_samples, device=dev)               # :203
    z = near * (1. - t_vals) + far * t_vals                                    # :206
    if n_surface > 0:
        nz = (gt_depth > 0).squeeze(-1)                                        # :179
        ts = torch.linspace(0., 1., steps=n_surface).double().to(dev)          # :183-184
        zs = torch.zeros(gt_depth.shape[0], n_surface, device=dev).double()    # :189-190
        d_nz = gt_depth[nz].reshape(-1, 1).repeat(1, n_surface)
        zs[nz] = 0.95 * d_nz * (1. - ts) + 1.05 * d_nz * ts                    # :186-193
        zs[~nz] = 0.001 * (1. - ts) + torch.max(gt_depth) * ts                 # :194-201
        z, _ = torch.sort(torch.cat([z, zs.double()], -1), -1)                 # :220-221
    return z


def composite(raw: torch.Tensor, z: torch.Tensor):
    """Occupancy compositing -- src/common.py:206-251 (occupancy branch :234-239).
    raw [N,S,4], z [N,S] (f64).  Returns depth, var (z's dtype), rgb [N,3], weights [N,S]."""
    alpha = torch.sigmoid(10 * raw[..., 3])                                    # :236
    ones = torch.ones((alpha.shape[0], 1), device=alpha.device, dtype=alpha.dtype)
    trans = torch.cumprod(torch.cat([ones, (1. - alpha + 1e-10)], -1), -1)[:, :-1]   # :244-245
    weights = alpha * trans
    rgb = torch.sum(weights[..., None] * raw[..., :3], -2)                     # :247
    depth = torch.sum(weights * z, -1)                                         # :248
    tmp = z - depth.unsqueeze(-1)
    var = torch.sum(weights * tmp * tmp, dim=1)                                # :249-250
    return depth, var, rgb, weights


def render_batch_ray(params: Params, grids, rays_d, rays_o, tsdf_volume, tsdf_bnds, bound, stage,
                     gt_depth=None, n_samples=32, n_surface=16, points_batch_size=500000,
                     return_aux: bool = False):
    """Renderer.render_batch_ray -- Renderer.py:110-255 (N_importance == 0).
    Returns depth [N] f64, uncertainty [N] f64, color [N,3], attention weight [N,S,1]."""
    z = sample_depths(rays_o, rays_d, gt_depth, bound, n_samples, n_surface)
    S = z.shape[1]
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]        # :223-224 (f64)
    raw, w = eval_points(params, pts.reshape(-1, 3), grids, tsdf_volume, tsdf_bnds, bound, stage,
                         points_batch_size)
    N = rays_o.shape[0]
    raw = raw.reshape(N, S, 4)
    depth, var, rgb, cw = composite(raw, z)
    out = (depth, var, rgb, w.reshape(N, S, 1))
    if return_aux:
        return out, {"z": z, "pts": pts, "raw": raw, "composite_weights": cw}
    return out


def get_rays(H, W, fx, fy, cx, cy, c2w: torch.Tensor):
    """Full-image rays -- src/common.py:254-272."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1).to(c2w.device)
    rays_d = torch.sum(dirs.reshape(H, W, 1, 3) * c2w[:3, :3], -1)
    return c2w[:3, -1].expand(rays_d.shape), rays_d


def get_rays_from_uv(i, j, c2w, fx, fy, cx, cy):
    """Rays for chosen pixels -- src/common.py:76-91.  Pose gradients enter here."""
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1).reshape(-1, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)
    return c2w[:3, -1].expand(rays_d.shape), rays_d


def quad2rotation(quad: torch.Tensor) -> torch.Tensor:
    """Quaternion [B,4] (r,i,j,k) -> rotation [B,3,3] -- src/common.py:139-162."""
    qr, qi, qj, qk = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    two_s = 2.0 / (quad * quad).sum(-1)
    rows = [
        torch.stack([1 - two_s * (qj ** 2 + qk ** 2), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr)], -1),
        torch.stack([two_s * (qi * qj + qk * qr), 1 - two_s * (qi ** 2 + qk ** 2), two_s * (qj * qk - qi * qr)], -1),
        torch.stack([two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi ** 2 + qj ** 2)], -1),
    ]
    return torch.stack(rows, 1)


def get_camera_from_tensor(inputs: torch.Tensor) -> torch.Tensor:
    """7-vector (quaternion, translation) -> [3,4] -- src/common.py:165-178."""
    one = inputs.dim() == 1
    x = inputs.unsqueeze(0) if one else inputs
    RT = torch.cat([quad2rotation(x[:, :4]), x[:, 4:, None]], 2)
    return RT[0] if one else RT


def render_img(params, grids, c2w, tsdf_volume, tsdf_bnds, bound, stage, gt_depth, H, W, fx, fy, cx, cy,
               ray_batch_size=100000, **kw):
    """Chunked full-frame render -- Renderer.py:258-327 (per-chunk depth maxima)."""
    with torch.no_grad():
        rays_o, rays_d = get_rays(H, W, fx, fy, cx, cy, c2w)
        rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
        gt = gt_depth.reshape(-1)
        ds, us, cs = [], [], []
        for i in range(0, rays_d.shape[0], ray_batch_size):
            d, u, c, _ = render_batch_ray(params, grids, rays_d[i:i + ray_batch_size], rays_o[i:i + ray_batch_size],
                                          tsdf_volume, tsdf_bnds, bound, stage, gt[i:i + ray_batch_size], **kw)
            ds.append(d.double()); us.append(u.double()); cs.append(c)
        return torch.cat(ds).reshape(H, W), torch.cat(us).reshape(H, W), torch.cat(cs).reshape(H, W, 3)


# --------------------------------------------------------------------------
# callers' losses (used to build gradient test cases and the CPU baseline)
# --------------------------------------------------------------------------
def mapper_loss(depth, color, weight, gt_depth, gt_color, stage, w_color_loss=0.2, attn_term=False):
    """src/Mapper.py:457-469."""
    m = gt_depth > 0
    loss = torch.abs(gt_depth[m] - depth[m]).sum()
    if attn_term:
        loss = loss + torch.abs(weight - torch.ones_like(weight)).sum()
    if stage == "color":
        loss = loss + w_color_loss * torch.abs(gt_color - color).sum()
    return loss


def tracker_loss(depth, uncertainty, color, gt_depth, gt_color, w_color_loss=0.5):
    """src/Tracker.py:116-129 (handle_dynamic=True, use_color_in_tracking=True)."""
    unc = uncertainty.detach()
    tmp = torch.abs(gt_depth - depth) / torch.sqrt(unc + 1e-10)
    m = (tmp < 10 * tmp.median()) & (gt_depth > 0)
    loss = (torch.abs(gt_depth - depth) / torch.sqrt(unc + 1e-10))[m].sum()
    return loss + w_color_loss * torch.abs(gt_color - color)[m].sum()


def get_grid_uniform(mc_bound, resolution: int, padding: float = 0.05) -> torch.Tensor:
    """Mesher query points -- src/utils/Mesher.py:365-393 (np.meshgrid default 'xy' indexing)."""
    ax = [np.linspace(mc_bound[k][0] - padding, mc_bound[k][1] + padding, resolution) for k in range(3)]
    xx, yy, zz = np.meshgrid(*ax)
    return torch.tensor(np.vstack([xx.ravel(), yy.ravel(), zz.ravel()]).T, dtype=torch.float)
