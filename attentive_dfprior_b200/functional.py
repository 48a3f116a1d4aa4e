"""autograd glue between PyTorch tensors and the C ABI (include/dfprior_b200.h).

PyTorch is used for device memory, streams and autograd bookkeeping only; all arithmetic of
the hot path happens in libdfprior_b200.so.  Differentiable inputs: the three feature grids,
the 79 decoder tensors, ``rays_o`` / ``rays_d`` (or the query points).  ``tsdf_volume``, bounds
and ``gt_depth`` never receive gradients.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, Optional

import torch

from . import _lib as L

GRID_KEYS = ("grid_low", "grid_high", "grid_color")


# ----------------------------------------------------------------------------------------
# bounds
# ----------------------------------------------------------------------------------------
class _BoundsCache:
    """[3,2] bound tensors -> python floats.

    CPU tensors (``bound`` as the reference keeps it, src/DF_Prior.py:185-194) are read on every call: ``tolist`` of six
    doubles costs microseconds.  CUDA tensors (``tsdf_bnds``, src/DF_Prior.py:87) need a device->host copy, so their
    values are remembered -- per tensor OBJECT, held through a weak reference and checked for identity and version on
    every hit.  Nothing is keyed on addresses: a freed tensor's address (and ``id``) is handed to the next allocation at
    once, and an address-keyed cache then returns the previous scene's box (round-1 bug: 16-50 % wrong outputs when a
    second scene was rendered in the same process)."""

    def __init__(self):
        self.d = {}

    def get(self, t: torch.Tensor):
        if not t.is_cuda:
            return t.detach().double().reshape(3, 2).tolist()
        key = id(t)
        hit = self.d.get(key)
        if hit is not None:
            ref, version, vals = hit
            if ref() is t and version == t._version:
                return vals
        vals = t.detach().to("cpu").double().reshape(3, 2).tolist()
        d = self.d
        self.d[key] = (weakref.ref(t, lambda _r, k=key: d.pop(k, None)), t._version, vals)
        return vals


_bounds = _BoundsCache()
_tables = {}


_cfg_cache = {}


def _cfg_struct(stage: str, n_samples: int, n_surface: int, n_importance: int = 0, flags: int = 0) -> L.DfpCfg:
    key = (n_samples, n_surface)
    tab = _tables.get(key)
    if tab is None:
        tv = torch.linspace(0., 1., steps=n_samples)                       # Renderer.py:203
        ts = torch.linspace(0., 1., steps=max(n_surface, 1)).double()      # Renderer.py:183-184
        tab = (tv.tolist(), (1. - tv).tolist(), ts.tolist(), (1. - ts).tolist())
        _tables[key] = tab
    ckey = (stage, n_samples, n_surface, n_importance, flags)
    cfg = _cfg_cache.get(ckey)                       # read-only once built: shared between calls
    if cfg is not None:
        return cfg
    cfg = L.DfpCfg()
    cfg.stage = L.STAGES[stage]
    cfg.n_samples, cfg.n_surface, cfg.n_importance = n_samples, n_surface, n_importance
    cfg.flags = flags
    for i in range(n_samples):
        cfg.t_vals[i] = tab[0][i]
        cfg.one_minus_t[i] = tab[1][i]
    for i in range(n_surface):
        cfg.t_surf[i] = tab[2][i]
        cfg.one_minus_t_surf[i] = tab[3][i]
    _cfg_cache[ckey] = cfg
    return cfg


def _is_ndhwc(g: torch.Tensor) -> bool:
    return g.dim() == 5 and g.shape[0] == 1 and g.is_contiguous(memory_format=torch.channels_last_3d)


def _grid_ndhwc(g: torch.Tensor) -> torch.Tensor:
    """Return a tensor whose memory is [D,H,W,C] for the logical [1,C,D,H,W] grid ``g``.

    Grids allocated with ``memory_format=torch.channels_last_3d`` are used in place (the fast path
    INTEGRATION.md recommends); channels-first grids are transposed by dfp_grid_to_ndhwc."""
    if g.dtype != torch.float32 or g.dim() != 5 or g.shape[0] != 1:
        raise ValueError("feature grids must be float32 [1,C,D,H,W]")
    gd = g.detach()
    if _is_ndhwc(gd):
        return gd
    src = gd.contiguous()
    _, Cc, D, H, W = src.shape
    dst = torch.empty((1, Cc, D, H, W), device=g.device, dtype=torch.float32,
                      memory_format=torch.channels_last_3d)
    L.check(L.lib().dfp_grid_to_ndhwc(src.data_ptr(), dst.data_ptr(), Cc, D, H, W,
                                      torch.cuda.current_stream().cuda_stream))
    return dst


def _scene_struct(grids_cl: Dict[str, Optional[torch.Tensor]], tsdf_volume, tsdf_bnds, bound) -> L.DfpScene:
    sc = L.DfpScene()
    for key in GRID_KEYS:
        g = grids_cl.get(key)
        if g is None:
            continue
        s = getattr(sc, key[5:])
        s.data = g.data_ptr()
        s.C, s.D, s.H, s.W = g.shape[1], g.shape[2], g.shape[3], g.shape[4]
    if tsdf_volume is not None:
        if tsdf_volume.dtype != torch.float32 or tsdf_volume.dim() != 5:
            raise ValueError("tsdf_volume must be float32 [1,1,Z,Y,X]")
        sc.tsdf.data = tsdf_volume.data_ptr()
        sc.tsdf.D, sc.tsdf.H, sc.tsdf.W = tsdf_volume.shape[2:]
        sc.tsdf.sD, sc.tsdf.sH, sc.tsdf.sW = tsdf_volume.stride()[2:]
    b = _bounds.get(bound)
    tb = _bounds.get(tsdf_bnds) if tsdf_bnds is not None else b
    for k in range(3):
        sc.bound[k][0], sc.bound[k][1] = b[k][0], b[k][1]
        sc.tsdf_bnds[k][0], sc.tsdf_bnds[k][1] = tb[k][0], tb[k][1]
    return sc


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.DfpError("attentive_dfprior_b200 runs on CUDA tensors only (there is no CPU fallback)")


class _Ctx:
    """Non-tensor arguments of one call."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def _needed_grids(stage):
    return {"low": ("grid_low",), "high": ("grid_low", "grid_high"),
            "color": ("grid_low", "grid_high", "grid_color")}[stage]


def _grads_struct(ctx, needs_grid, need_params, dev):
    """Allocate zeroed gradient sinks and the C struct pointing at them.

    All sinks are views of ONE flat zero-filled buffer (one memset; one all-reduce when the batch is sharded over
    ranks).  Grid gradients are channels-last views ([1,C,D,H,W] logical shape over an NDHWC slice)."""
    g = L.DfpGrads()
    # bucket order of the sharded all-reduce: the big grids first (final early in the backward), then grid_low + parameters
    order = ("grid_high", "grid_color", "grid_low")
    want = [k for k in order if needs_grid.get(k) and ctx.grids_cl.get(k) is not None]
    sizes = [ctx.grids_cl[k].numel() for k in want]
    npk = int(L.lib().dfp_packed_grad_floats()) if need_params else 0
    pad = lambda n: (n + 63) // 64 * 64                              # keep every slice 256-byte aligned
    total = sum(pad(n) for n in sizes) + npk
    flat = torch.zeros(total, device=dev, dtype=torch.float32) if total else None
    ggrids, off, spans = {}, 0, {}
    for k, n in zip(want, sizes):
        _, C_, D_, H_, W_ = ctx.grids_cl[k].shape
        t = flat[off:off + n].view(1, D_, H_, W_, C_).permute(0, 4, 1, 2, 3)
        ggrids[k] = t
        setattr(g, k, t.data_ptr())
        spans[k] = (off, off + pad(n))
        off += pad(n)
    g._spans = spans
    gp = None
    if need_params:
        gp = flat[off:off + npk]
        g.params_packed = gp.data_ptr()
    return g, ggrids, gp, flat


def _param_grads(gp, params, needs):
    """Packed gradient -> 79 views of one flat native-layout buffer."""
    if gp is None:
        return [None] * len(params)
    flat = torch.empty(L.N_PARAMS, device=gp.device, dtype=torch.float32)
    L.check(L.lib().dfp_unpack_param_grads(gp.data_ptr(), flat.data_ptr(), _stream()))
    # one C++ call makes the 79 shaped views (a Python slice/view/to per tensor cost 0.4 ms per step, measured)
    views = torch._C._nn.unflatten_dense_tensors(flat, list(params))
    return [(v if v.dtype == p.dtype else v.to(p.dtype)) if need else None for v, p, need in zip(views, params, needs)]


_comm_streams = {}


def _comm_stream(dev):
    s = _comm_streams.get(dev)
    if s is None:
        s = _comm_streams[dev] = torch.cuda.Stream(device=dev)
    return s


def _reduce_buckets(flat, spans, overlap, group):
    """All-reduce ``flat`` in three buckets on the side stream: grid_high as soon as the library has recorded its event,
    grid_color likewise, then the rest (grid_low + packed parameter gradients) when the whole backward is enqueued.
    Every rank issues the same collectives in the same order; the compute stream waits for the side stream at the end."""
    import torch.distributed as dist
    ev_high, ev_color, main, comm = overlap
    flat.record_stream(comm)
    done = 0
    with torch.cuda.stream(comm):
        for key, ev in (("grid_high", ev_high), ("grid_color", ev_color)):
            if key in spans:
                lo, hi = spans[key]
                if lo != done:                                           # buckets are laid out in this order; be safe
                    continue
                comm.wait_event(ev)
                dist.all_reduce(flat[lo:hi], group=group)
                done = hi
    tail = torch.cuda.Event()
    tail.record(main)
    with torch.cuda.stream(comm):
        comm.wait_event(tail)
        if done < flat.numel():
            dist.all_reduce(flat[done:], group=group)
    fin = torch.cuda.Event()
    fin.record(comm)
    main.wait_event(fin)


# ----------------------------------------------------------------------------------------
# Renderer.render_batch_ray
# ----------------------------------------------------------------------------------------
# A training call keeps per-sample operands for its backward in the workspace.  Above this many bytes the call is run
# in ray chunks instead: the forward keeps nothing, and the backward re-runs each chunk's forward right before that
# chunk's backward (one extra forward, bounded memory).  SURVEY 8d's 1 048 576-ray variant needs this.
WORKSPACE_BUDGET_BYTES = 24 << 30
MERGED_MAX_SAMPLES = 148 * 2 * 128 * 10      # dfp_api.cu: calls up to this size run their decoders in one launch per direction


def _frozen_flags(need_params, train):
    """DFP_FLAG_FROZEN_* for the decoders none of whose 23 tensors (state_dict order: low, high, colour) needs a gradient:
    the library then keeps no weight-gradient operands for them and skips their weight-gradient kernel.  The reference's
    Mapper never optimises the low decoder and, with fix_high / fix_color, leaves those out as well (src/Mapper.py:364-371);
    a caller that also sets ``requires_grad_(False)`` on them gets this for free, like autograd's own pruning."""
    if not train or len(need_params) < 69:
        return 0
    f = 0
    for d in range(3):
        if not any(need_params[23 * d:23 * (d + 1)]):
            f |= L.FLAG_FROZEN[d]
    return f


class _RenderBatchRay(torch.autograd.Function):
    @staticmethod
    def _fwd_call(aux, scene, packed, ro, rd, gt, dmax, lo, hi, flags, outs, want_aux):
        """One dfp_render_fwd over rays [lo, hi); returns the workspace it wrote."""
        lib = L.lib()
        n = hi - lo
        depth, unc, color, weight, z, raw = outs
        S = weight.shape[1]
        cfg = _cfg_struct(aux.stage, aux.n_samples, aux.n_surface, aux.n_importance, flags)
        rays = L.DfpRays(ro[lo:hi].data_ptr(), rd[lo:hi].data_ptr(), _ptr(gt[lo:hi]) if gt is not None else None, n,
                         aux.chunk, _ptr(dmax))
        ws_bytes = int(lib.dfp_render_workspace_bytes(n, S, aux.chunk, L.STAGES[aux.stage], flags))
        ws = torch.empty(ws_bytes, device=ro.device, dtype=torch.uint8)
        out = L.DfpRenderOut(depth[lo:hi].data_ptr(), unc[lo:hi].data_ptr(), color[lo:hi].data_ptr(), weight[lo:hi].data_ptr(),
                             _ptr(z[lo:hi]) if want_aux else None, _ptr(raw[lo:hi]) if want_aux else None)
        L.check(lib.dfp_render_fwd(C.byref(scene), packed.data_ptr(), C.byref(rays), C.byref(cfg), C.byref(out),
                                   ws.data_ptr(), ws_bytes, _stream()))
        return ws, ws_bytes, rays, cfg

    @staticmethod
    def forward(ctx, aux, rays_o, rays_d, grid_low, grid_high, grid_color, *params):
        lib = L.lib()
        dev = rays_o.device
        ro = rays_o.detach().float().contiguous()
        rd = rays_d.detach().float().contiguous()
        gt = aux.gt_depth
        N = ro.shape[0]
        S = aux.n_samples + (aux.n_surface if gt is not None else 0)
        given = {"grid_low": grid_low, "grid_high": grid_high, "grid_color": grid_color}
        grids_cl = {k: (_grid_ndhwc(given[k]) if k in _needed_grids(aux.stage) else None) for k in GRID_KEYS}
        packed = aux.pack.get(list(params))
        scene = _scene_struct(grids_cl, aux.tsdf_volume, aux.tsdf_bnds, aux.bound)
        train = any(ctx.needs_input_grad)
        flags = (L.FLAG_SIMT if aux.simt else 0) | (L.FLAG_SAVE_ACT if train else 0) | _frozen_flags(ctx.needs_input_grad[6:], train)
        depth = torch.empty(N, device=dev, dtype=torch.float64)
        unc = torch.empty(N, device=dev, dtype=torch.float64)
        color = torch.empty(N, 3, device=dev, dtype=torch.float32)
        weight = torch.empty(N, S, 1, device=dev, dtype=torch.float32)
        z = torch.empty(N, S, device=dev, dtype=torch.float64) if aux.want_aux else None
        raw = torch.empty(N, S, 4, device=dev, dtype=torch.float32) if aux.want_aux else None
        outs = (depth, unc, color, weight, z, raw)
        dmax = aux.depth_max
        # ray chunks for calls whose saved operands would not fit the budget (training calls only; render_img's
        # chunk-exact maxima are a forward-only feature and keep the single-call path)
        per_ray = int(lib.dfp_render_workspace_bytes(1 << 14, S, 0, L.STAGES[aux.stage], flags)) >> 14
        step = N
        if train and aux.chunk == 0 and N * per_ray > aux.ws_budget:
            step = max(1, int(aux.ws_budget // max(per_ray, 1)))
            if gt is not None and dmax is None:
                dmax = gt.max().reshape(1)                   # the far clamp uses the max of the WHOLE batch (Renderer.py:159,195)
        keep = _Ctx(ro=ro, rd=rd, grids_cl=grids_cl, packed=packed, N=N, S=S, params=params, flags=flags, scene=scene,
                    dmax=dmax, step=step, ws=None, ws_bytes=0)
        if step >= N:
            keep.ws, keep.ws_bytes, _, _ = _RenderBatchRay._fwd_call(aux, scene, packed, ro, rd, gt, dmax, 0, N, flags, outs,
                                                                     aux.want_aux)
        else:
            for lo in range(0, N, step):                     # nothing kept: the backward recomputes chunk by chunk
                _RenderBatchRay._fwd_call(aux, scene, packed, ro, rd, gt, dmax, lo, min(N, lo + step),
                                          flags & ~L.FLAG_SAVE_ACT, outs, aux.want_aux)
        if aux.want_aux:
            aux.z_vals, aux.raw = z, raw
        ctx.aux = aux
        ctx.keep = keep
        return depth, unc, color, weight

    @staticmethod
    def backward(ctx, g_depth, g_unc, g_color, g_weight):
        lib = L.lib()
        aux, k = ctx.aux, ctx.keep
        dev = k.ro.device
        need = ctx.needs_input_grad           # (aux, rays_o, rays_d, low, high, color, *params)
        needs_grid = {"grid_low": need[3], "grid_high": need[4], "grid_color": need[5]}
        need_params = any(need[6:])
        with torch.cuda.device(dev):
            g, ggrids, gp, flat = _grads_struct(k, needs_grid, need_params, dev)
            d_o = torch.empty_like(k.ro) if (need[1] or need[2]) else None
            d_d = torch.empty_like(k.rd) if (need[1] or need[2]) else None

            def prep(t, dtype):
                return None if t is None else t.detach().to(dtype).contiguous()

            gd, gu, gc, gw = (prep(g_depth, torch.float64), prep(g_unc, torch.float64), prep(g_color, torch.float32),
                              prep(g_weight, torch.float32))
            gt = aux.gt_depth
            sl = lambda t, lo, hi: None if t is None else t[lo:hi].data_ptr()
            # sharded + single-chunk: the library records an event when each big grid's gradient is final, so that its
            # all-reduce runs on a side stream under the remaining decoders' backward
            overlap = None
            # (not for small shards: those run their decoders in merged launches, which beat three overlapped buckets)
            if aux.group is not None and flat is not None and k.step >= k.N and aux.overlap and k.N * k.S > MERGED_MAX_SAMPLES:
                overlap = (torch.cuda.Event(), torch.cuda.Event(), torch.cuda.current_stream(), _comm_stream(dev))
                overlap[0].record(); overlap[1].record()                  # create the CUDA events (lazily created otherwise)
                g.ev_high, g.ev_color = overlap[0].cuda_event, overlap[1].cuda_event
            for lo in range(0, k.N, k.step):
                hi = min(k.N, lo + k.step)
                if k.ws is not None:
                    ws, ws_bytes = k.ws, k.ws_bytes
                    cfg = _cfg_struct(aux.stage, aux.n_samples, aux.n_surface, aux.n_importance, k.flags)
                    rays = L.DfpRays(k.ro.data_ptr(), k.rd.data_ptr(), _ptr(gt), k.N, aux.chunk, _ptr(k.dmax))
                else:                                        # chunked call: this chunk's forward again, keeping its operands
                    scratch = (torch.empty(hi - lo, device=dev, dtype=torch.float64), torch.empty(hi - lo, device=dev, dtype=torch.float64),
                               torch.empty(hi - lo, 3, device=dev), torch.empty(hi - lo, k.S, 1, device=dev), None, None)
                    ws, ws_bytes, rays, cfg = _RenderBatchRay._fwd_call(aux, k.scene, k.packed, k.ro[lo:hi], k.rd[lo:hi],
                                                                        None if gt is None else gt[lo:hi], k.dmax, 0, hi - lo,
                                                                        k.flags, scratch, False)
                g.rays_o, g.rays_d = sl(d_o, lo, hi), sl(d_d, lo, hi)
                og = L.DfpRenderOutGrad(sl(gd, lo, hi), sl(gu, lo, hi), sl(gc, lo, hi), sl(gw, lo, hi))
                L.check(lib.dfp_render_bwd(C.byref(k.scene), k.packed.data_ptr(), C.byref(rays), C.byref(cfg), C.byref(og),
                                           C.byref(g), ws.data_ptr(), ws_bytes, _stream()))
            if aux.group is not None and flat is not None:
                # ray-sharded data parallelism: sum the replicated state's gradients over the ranks (NCCL)
                if overlap is not None:
                    _reduce_buckets(flat, g._spans, overlap, aux.group)
                else:
                    import torch.distributed as dist
                    dist.all_reduce(flat, group=aux.group)           # grids + packed parameters in one collective
            pgrads = _param_grads(gp, k.params, need[6:])
        return (None, d_o if need[1] else None, d_d if need[2] else None,
                ggrids.get("grid_low"), ggrids.get("grid_high"), ggrids.get("grid_color"), *pgrads)


def render_batch_ray(c, decoders, rays_d, rays_o, tsdf_volume, tsdf_bnds, bound, stage, gt_depth=None,
                     n_samples=32, n_surface=16, n_importance=0, chunk=0, want_aux=False, depth_max=None, group=None,
                     simt=False, workspace_budget=None, overlap_allreduce=True):
    """Functional form of Renderer.render_batch_ray (src/utils/Renderer.py:110-255).

    Returns ``(depth [N] f64, uncertainty [N] f64, color [N,3] f32, weight [N,S,1] f32)``; with
    ``want_aux`` also a dict holding the sorted sample depths and the pre-sigmoid raw values.
    ``simt`` evaluates the MLPs with the fp32 SIMT kernels (cross-check of the tensor-core path)."""
    from .decoder import decoder_param_list, param_pack_of
    _require_cuda(rays_o, rays_d, tsdf_volume)
    if stage not in L.STAGES:
        raise ValueError(f"unknown stage {stage!r}")
    if rays_o.shape[0] == 0:
        raise ValueError("render_batch_ray needs at least one ray")
    params = decoder_param_list(decoders)
    gt = None if gt_depth is None else gt_depth.detach().reshape(-1).float().contiguous()
    aux = _Ctx(stage=stage, n_samples=int(n_samples), n_surface=int(n_surface), n_importance=int(n_importance),
               chunk=int(chunk), gt_depth=gt, tsdf_volume=tsdf_volume.detach() if tsdf_volume is not None else None,
               tsdf_bnds=tsdf_bnds, bound=bound, pack=param_pack_of(decoders), want_aux=want_aux,
               depth_max=None if depth_max is None else depth_max.detach().reshape(-1)[:1].float().contiguous(),
               group=group, simt=bool(simt), overlap=bool(overlap_allreduce),
               ws_budget=WORKSPACE_BUDGET_BYTES if workspace_budget is None else int(workspace_budget))
    dummy = rays_o.new_zeros(())
    gl = c.get("grid_low")
    gh = c.get("grid_high") if stage != "low" else None
    gc = c.get("grid_color") if stage == "color" else None
    with torch.cuda.device(rays_o.device):
        out = _RenderBatchRay.apply(aux, rays_o, rays_d, gl, gh if gh is not None else dummy,
                                    gc if gc is not None else dummy, *params)
    if want_aux:
        return out, {"z": aux.z_vals, "raw": aux.raw}
    return out


# ----------------------------------------------------------------------------------------
# Renderer.eval_points / Mesher.eval_points / DF.forward
# ----------------------------------------------------------------------------------------
class _EvalPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, aux, pts, grid_low, grid_high, grid_color, *params):
        lib = L.lib()
        dev = pts.device
        p = pts.detach()
        if p.dtype not in (torch.float32, torch.float64):
            p = p.float()
        p = p.contiguous()
        n = p.shape[0]
        given = {"grid_low": grid_low, "grid_high": grid_high, "grid_color": grid_color}
        grids_cl = {k: (_grid_ndhwc(given[k]) if k in _needed_grids(aux.stage) else None) for k in GRID_KEYS}
        packed = aux.pack.get(list(params))
        scene = _scene_struct(grids_cl, aux.tsdf_volume, aux.tsdf_bnds, aux.bound)
        flags = ((L.FLAG_SIMT if aux.simt else 0) | (L.FLAG_SAVE_ACT if any(ctx.needs_input_grad) else 0)
                 | _frozen_flags(ctx.needs_input_grad[5:], any(ctx.needs_input_grad)))
        P = L.DfpPoints(p.data_ptr(), 1 if p.dtype == torch.float64 else 0, 1 if aux.apply_bound_mask else 0, n, flags, 0)
        raw = torch.empty(n, 4, device=dev, dtype=torch.float32)
        w = torch.empty(n, device=dev, dtype=torch.float32)
        ws_bytes = int(lib.dfp_points_workspace_bytes(n, L.STAGES[aux.stage], flags))
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        L.check(lib.dfp_points_fwd(C.byref(scene), packed.data_ptr(), C.byref(P), L.STAGES[aux.stage],
                                   raw.data_ptr(), w.data_ptr(), ws.data_ptr(), ws_bytes, _stream()))
        ctx.aux = aux
        ctx.keep = _Ctx(p=p, grids_cl=grids_cl, packed=packed, ws=ws, ws_bytes=ws_bytes, n=n, params=params,
                        scene=scene, flags=flags)
        return raw, w

    @staticmethod
    def backward(ctx, g_raw, g_w):
        lib = L.lib()
        aux, k = ctx.aux, ctx.keep
        dev = k.p.device
        need = ctx.needs_input_grad           # (aux, pts, low, high, color, *params)
        needs_grid = {"grid_low": need[2], "grid_high": need[3], "grid_color": need[4]}
        need_params = any(need[5:])
        with torch.cuda.device(dev):
            g, ggrids, gp, flat = _grads_struct(k, needs_grid, need_params, dev)
            d_p = torch.empty_like(k.p) if need[1] else None
            g.points = _ptr(d_p)
            gr = g_raw.detach().float().contiguous() if g_raw is not None else None
            gw = g_w.detach().float().contiguous() if g_w is not None else None
            scene = k.scene                      # the forward's own scene record (grids, TSDF and both boxes as it saw them)
            P = L.DfpPoints(k.p.data_ptr(), 1 if k.p.dtype == torch.float64 else 0, 1 if aux.apply_bound_mask else 0, k.n,
                            k.flags, 0)
            L.check(lib.dfp_points_bwd(C.byref(scene), k.packed.data_ptr(), C.byref(P), L.STAGES[aux.stage],
                                       _ptr(gr), _ptr(gw), C.byref(g), k.ws.data_ptr(), k.ws_bytes, _stream()))
            pgrads = _param_grads(gp, k.params, need[5:])
        return (None, d_p, ggrids.get("grid_low"), ggrids.get("grid_high"), ggrids.get("grid_color"), *pgrads)


def eval_points(p, decoders, c, tsdf_volume, tsdf_bnds, bound, stage="color", apply_bound_mask=True, simt=False):
    """Point query: ``raw [P,4]``, ``w [P]``.  With ``apply_bound_mask`` it is Renderer/Mesher.eval_points
    (occupancy 100 outside the bound, Renderer.py:51-64); without, the bare DF.forward (decoder.py:307-353)."""
    from .decoder import decoder_param_list, param_pack_of
    _require_cuda(p, tsdf_volume)
    if stage not in L.STAGES:
        raise ValueError(f"unknown stage {stage!r}")
    if p.shape[0] == 0:
        raise ValueError("eval_points needs at least one point")
    params = decoder_param_list(decoders)
    aux = _Ctx(stage=stage, tsdf_volume=tsdf_volume.detach() if tsdf_volume is not None else None,
               tsdf_bnds=tsdf_bnds, bound=bound, pack=param_pack_of(decoders), apply_bound_mask=apply_bound_mask,
               simt=bool(simt))
    dummy = p.new_zeros((), dtype=torch.float32)
    gl = c.get("grid_low")
    gh = c.get("grid_high") if stage != "low" else None
    gc = c.get("grid_color") if stage == "color" else None
    with torch.cuda.device(p.device):
        return _EvalPoints.apply(aux, p, gl, gh if gh is not None else dummy, gc if gc is not None else dummy, *params)


def sample_tsdf(p, tsdf_volume, tsdf_bnds):
    """Trilinear TSDF values at ``p`` -- Renderer.sample_grid_tsdf (Renderer.py:73-81).  Returns [P] f32."""
    _require_cuda(p, tsdf_volume)
    pp = p.detach().reshape(-1, 3)
    if pp.dtype not in (torch.float32, torch.float64):
        pp = pp.float()
    pp = pp.contiguous()
    n = pp.shape[0]
    out = torch.empty(n, device=pp.device, dtype=torch.float32)
    if n == 0:
        return out
    scene = _scene_struct({}, tsdf_volume.detach(), tsdf_bnds, tsdf_bnds)
    P = L.DfpPoints(pp.data_ptr(), 1 if pp.dtype == torch.float64 else 0, 0, n, 0, 0)
    with torch.cuda.device(pp.device):
        L.check(L.lib().dfp_tsdf_sample(C.byref(scene), C.byref(P), out.data_ptr(), _stream()))
    return out
