#!/usr/bin/env python
"""Benchmark of the DF-Prior volume-rendering hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One *step* = one Replica-room0-shaped mapping iteration of the hot path: Renderer.render_batch_ray
forward (stage `color`: low + high + colour decoders, attentive TSDF fusion, compositing), the Mapper
loss (src/Mapper.py:457-469) and the backward into the three dense feature grids and all 79 decoder
tensors.  Synthetic data (random-init grids/decoders, analytic box-room TSDF at the reference's 4/256 m
voxel size, analytic depth); N > 1 shards rays across ranks (weak scaling: fixed rays per GPU) with one
NCCL all-reduce of grid + decoder gradients per step.

Prints ONE JSON line (rank 0).  `value` = rays/s with inputs resident in HBM; `e2e` = the same step driven
from pinned host buffers (H2D of rays/depth/colour and D2H of the loss inside the timed region).
`--impl reference` times the reference algorithm (the oracle port of the reference's PyTorch path; the
reference is pure Python and cannot travel to the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "mapping rays/s (fwd+bwd render_batch_ray)"
UNIT = "rays/s"
STAGE = "color"
S = 48
# algorithmic bytes (C=32 f32, no inter-sample reuse, scatter counted write-only)
BYTES_PER_RAY = {"low": 147704, "high": 298232, "color": 445688}
# per-sample algorithmic bytes of the individual kernels of the pipeline
KERNEL_BYTES_PER_SAMPLE = {
    "k_decoder_fwd": {"low": 1024, "high": 2048, "color": 1024},
    "k_decoder_bwd": {"low": 2048, "high": 3072, "color": 2048},
    "k_prep": 32 + 24, "k_attn_fwd": 20, "k_attn_bwd": 28 + 32,
    # tensor-core path: forward gathers as above; K1 (delta chain) scatters the feature gradient (1024 B) and re-gathers
    # nothing; K2 (weight gradients) touches no grid data: its algorithmic bytes are the 32 B/sample of deltas in and
    # the parameter gradient out (negligible) -- its real traffic is the saved-activation stream
    "k_decoder_fwd_tc": {"low": 1024, "high": 2048, "color": 1024},
    "k_decoder_dh_tc": {"low": 1024, "high": 1024, "color": 1024},
    "k_decoder_dw_tc": {"low": 32, "high": 32, "color": 32},
}


def cpu_model():
    """CPU model string of the host the CPU baseline ran on (state core count and CPU model)."""
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def workload_config(n, world, channels_first=False, tsdf_shape=None, extra=None):
    """The `config` object both arms print: it names the workload, not the implementation."""
    cfg = {"workload": "replica_room0_mapping_iteration", "stage": STAGE, "rays_per_gpu_per_step": n,
           "samples_per_ray": S, "grids": "low[21,28,37] high/color[43,56,74] x32",
           "tsdf": tsdf_shape, "loss": "Mapper depth L1 + 0.2 colour L1",
           "parallelism": f"ray-sharded dp{world}" + (" + all-reduce of grid/decoder grads" if world > 1 else ""),
           "l2": "GPU arm: flushed between timed steps (256 MiB write)"}
    if extra:
        cfg.update(extra)
    return cfg


def oracle_step_factory(scene, n_total, n_sample, device="cpu"):
    """One mapping step of the reference algorithm (oracle port) on `n_sample` rays of the `n_total`-ray batch the
    kernels are benchmarked on (same generator, same seed): returns (step_fn, n_sample)."""
    from attentive_dfprior_b200 import synth
    from oracle import dfprior_oracle as O
    params = {k: v.to(device).requires_grad_(True) for k, v in O.init_decoder_params(0).items()}
    grids = {k: v.to(device).requires_grad_(True) for k, v in scene.grids.items()}
    rays_o, rays_d, gt, gt_color = synth.make_rays(scene, n_total, seed=100)
    n_sample = min(n_sample, n_total)
    idx = torch.arange(0, n_total, n_total // n_sample)[:n_sample]              # an even sample of the batch
    rays_o, rays_d, gt, gt_color = (t[idx].to(device) for t in (rays_o, rays_d, gt, gt_color))
    tsdf, tb = scene.tsdf_volume.to(device), scene.tsdf_bnds.to(device)

    def step():
        for t in list(params.values()) + list(grids.values()):
            t.grad = None
        d, u, c, w = O.render_batch_ray(params, grids, rays_d, rays_o, tsdf, tb, scene.bound, STAGE, gt)
        loss = O.mapper_loss(d, c, w, gt, gt_color, STAGE)
        loss.backward()
        return float(loss)

    return step, n_sample


def run_reference(args):
    """The reference arm: the reference's algorithm for the path (its PyTorch ops, restated in oracle/ and pinned
    bit-exactly to the reference's outputs; the reference itself is Python and does not exist on the GPU box) on the
    host cores, or with --device cuda on the GPU as the honest GPU comparator.  Same workload / config as our arm;
    each step is a bounded sample of the batch (stated in cpu_baseline.sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from attentive_dfprior_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = args.device
    scene = synth.make_scene("room0", seed=0, high_std=1e-2)
    step, n = oracle_step_factory(scene, args.rays, args.ref_rays, dev)
    sync = torch.cuda.synchronize if dev != "cpu" else (lambda: None)
    for _ in range(args.warmup):
        step()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    sync()
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    where = f"torch CPU {cores} threads" if dev == "cpu" else f"torch {dev} (fp32, allow_tf32=False)"
    sample = (f"each step = {n} rays sampled evenly from the {args.rays}-ray batch, x {args.steps} steps, stage {STAGE}, "
              f"fwd+bwd, oracle port of the reference's PyTorch path, {where}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.rays, args.gpus, tsdf_shape=list(scene.tsdf_volume.shape[2:])),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores if dev == "cpu" else 0, "kind": "port", "sample": sample,
                         "device": dev, "cpu_model": cpu_model()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(scene_cpu, n_total, budget_s=20.0):
    """Bounded sample of the same workload through the oracle on the host cores (rank 0, N=1 only)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, n = oracle_step_factory(scene_cpu, n_total, 1024, "cpu")
    step()
    t0 = time.perf_counter()
    reps = 0
    while reps < 3 or (time.perf_counter() - t0 < budget_s and reps < 20):
        step()
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": n * reps / dt, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": cpu_model(),
            "sample": f"{reps} steps of {n} rays sampled evenly from the {n_total}-ray batch (stage {STAGE}, fwd+bwd, room0 shapes) "
                      f"through the oracle, torch CPU {cores} threads"}


def torch_gpu_baseline(scene_cpu, n_total, n_sample=8192, reps=5):
    """The reference's PyTorch ops for the same step on the B200 itself (BASELINE.md section 2 / SURVEY 8d: the honest GPU
    comparator): the oracle port on cuda:0, fp32 with TF32 off, CUDA-event timed."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    step, n = oracle_step_factory(scene_cpu, n_total, n_sample, "cuda:0")
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    return {"value": n / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "kind": "port on cuda:0",
            "sample": f"{reps} steps of {n} rays sampled evenly from the {n_total}-ray batch, torch {torch.__version__} eager, "
                      f"fp32 (allow_tf32=False), loss .item() closes each step"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
NCU_TRAFFIC = "profiles/ncu_traffic.json"       # written by tools/ncu_summarize.py from an `ncu --set full` capture of this bench


def kernel_source_hash():
    """sha256 over the CUDA sources + public header: identifies the build a profile was taken from (there is no .git on
    the GPU box)."""
    import hashlib
    h = hashlib.sha256()
    roots = [os.path.join(REPO, "attentive_dfprior_b200", "csrc"), os.path.join(REPO, "include")]
    for root in roots:
        for name in sorted(os.listdir(root)):
            if name.startswith(("dfp_mesh", "dfprior_mesh")):
                continue                      # mesh extraction after the Mesher query: not on the benchmarked path
            if name.endswith((".cu", ".cuh", ".h", ".inl")):
                h.update(name.encode())
                h.update(open(os.path.join(root, name), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(kernels):
    """DRAM bytes (read + write) of the named kernels' launches during ONE step of the default workload, from the
    committed `ncu --set full` capture (profiles/ncu_traffic.json).  Returns (bytes or None, note): None when there is
    no capture or when it was taken from different kernel sources than the ones being benchmarked."""
    path = os.path.join(REPO, NCU_TRAFFIC)
    if not os.path.exists(path):
        return None, "no ncu capture committed"
    d = json.load(open(path))
    mine = kernel_source_hash()
    if d.get("source_hash") != mine:
        return None, f"ncu capture is from sources {d.get('source_hash')}, benchmarked sources are {mine}: not reported"
    total = 0.0
    for name, rec in d.get("kernels", {}).items():
        name = name.replace("void ", "").strip()
        if any(k == name or name.startswith(k + "<") for k in kernels):
            total += rec["dram_bytes_per_step"]
    return (total or None), f"{NCU_TRAFFIC} (ncu --set full, {d.get('command', '?')}; sources {mine})"


def run_ours(args):
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200 import _lib as L
    from attentive_dfprior_b200 import dist as D
    from attentive_dfprior_b200 import synth
    from attentive_dfprior_b200 import mapper_ops as M
    from attentive_dfprior_b200.config import default_cfg, slam_namespace

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
        if os.environ.get("BENCH_NO_COLLECTIVE") == "1":     # experiment: independent replicas (no gradient exchange)
            group = None
    L.lib()

    n = args.rays if args.scaling == "weak" else max(1, args.rays // world)   # rays per GPU per step
    torch.manual_seed(0)
    scene = synth.make_scene("room0", seed=0, device=dev, high_std=1e-2, channels_last=not args.channels_first)
    model = A.DF().to(dev)
    renderer = A.Renderer(default_cfg(), None, slam_namespace(scene.bound, scene.tsdf_bnds))
    grids = {k: v.requires_grad_(True) for k, v in scene.grids.items()}
    params = list(model.parameters())
    # the global batch is n rays from each of `world` synthetic keyframes (one camera pose per frame); it is dealt to
    # the ranks interleaved, so every rank renders the same mix of frames (a contiguous split gives each rank one
    # frame, whose cost differs by up to 15 % with the pose: measured, profiles/r01 scaling notes)
    parts = [synth.make_rays(scene, n, seed=100 + k, device=dev) for k in range(world)]
    mine = D.interleaved_shard(n * world, rank, world).to(dev)
    rays_o, rays_d, gt, gt_color = (torch.cat([p[i] for p in parts])[mine].contiguous() for i in range(4))
    # every rank holds the keyframes, so the batch-global max(gt_depth) (Renderer.py:159,195) is computed locally before
    # the rays are dealt out: no all-reduce(max), the gradient all-reduce is the step's only collective
    gmax = torch.cat([p[2] for p in parts]).max().reshape(1) if world > 1 else None
    del parts
    host = [t.cpu().pin_memory() for t in (rays_o, rays_d, gt, gt_color)]
    h2d = sum(t.numel() * t.element_size() for t in host)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def zero_grads():
        for t in params:
            t.grad = None
        for t in grids.values():
            t.grad = None

    def step(ro, rd, g, gc, dm=None):
        zero_grads()
        d, u, c, w = D.render_batch_ray_sharded(renderer, grids, model, rd, ro, scene.tsdf_volume, scene.tsdf_bnds,
                                                STAGE, g, group=group, depth_max=dm,
                                                overlap_allreduce=not args.no_overlap)
        if args.torch_loss:
            m = g > 0
            loss = torch.abs(g[m] - d[m]).sum() + 0.2 * torch.abs(gc - c).sum()      # src/Mapper.py:457-469 as torch ops
        else:
            loss = M.mapper_loss(d, c, w, g, gc, 0.2, False, True)                   # the same loss as one kernel (+ gradient)
        loss.backward()
        return loss

    def step_dev():
        return step(rays_o, rays_d, gt, gt_color, gmax)

    def step_e2e():
        dv = [t.to(dev, non_blocking=True) for t in host]
        return float(step(*dv, gmax).detach())                                  # D2H of the loss closes the step

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if group is not None:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        l0 = L.launch_count()
        evs = []
        for _ in range(steps):
            flush.fill_(1)                                             # flush L2 between timed steps (outside the events)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        launches = L.launch_count() - l0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if os.environ.get("BENCH_RANK_TIMES") == "1":          # diagnostics: every rank's own time (stderr)
            print(f"[rank {rank}] {ms / steps:.3f} ms/step", file=sys.stderr, flush=True)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if group is not None:
            torch.distributed.barrier()
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.cuda.synchronize()
        return float(t), launches, clocks

    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches, clocks = timed(step_dev, args.steps, args.warmup, sampler)
    rays_total = n * world * args.steps
    value = rays_total / (ms / 1e3)
    ms_e2e, _, _ = timed(step_e2e, args.steps, max(3, args.warmup // 2))
    e2e = rays_total / (ms_e2e / 1e3)

    # live per-kernel durations (CUDA events on the launching stream) over a few more steps; every rank runs them
    # because the steps contain collectives, rank 0 reports
    prof_steps = 3
    kt = L.kernel_times(lambda: [step_dev() for _ in range(prof_steps)])
    # the reference's own operating points as latency figures, device-resident and end to end (pinned host buffers in,
    # loss / pose gradient out): Mapper 1000 pixels per step (configs/Replica/replica.yaml:23), Tracker 200 (:12-13)
    def time_loop(fn, reps=30, warm=8):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    small = tracker = frozen = None
    if not args.no_small and world == 1:
        # The reference's own optimiser set-up (configs/df_prior.yaml: fix_high True; src/Mapper.py:364-371 never hands the
        # low decoder to Adam): with requires_grad off on those two decoders the library keeps no weight-gradient operands
        # for them (DFP_FLAG_FROZEN_*).  Same batch and loss as the headline; the headline itself differentiates all 79
        # tensors, as the reference's autograd does.
        for prm in list(model.low_decoder.parameters()) + list(model.high_decoder.parameters()):
            prm.requires_grad_(False)
        ms_fr = time_loop(lambda: (flush.fill_(1), step_dev()), reps=10, warm=3)
        ms_fl = time_loop(lambda: flush.fill_(1), reps=10, warm=1)
        for prm in list(model.low_decoder.parameters()) + list(model.high_decoder.parameters()):
            prm.requires_grad_(True)
        frozen = {"ms_per_step": ms_fr - ms_fl, "rays_per_s": n / ((ms_fr - ms_fl) / 1e3),
                  "note": "low + high decoders frozen (requires_grad False, as the reference's Mapper optimises neither with df_prior.yaml); "
                          "grids, colour decoder and fusion MLP still receive gradients; L2 flush time subtracted"}
    if not args.no_small:
        n_small = max(1, 1000 // world)
        sm_dev = synth.make_rays(scene, n_small, seed=7 + rank, device=dev)
        sm_host = [t.cpu().pin_memory() for t in sm_dev]
        ms_dev = time_loop(lambda: step(*sm_dev))
        ms_e2e_small = time_loop(lambda: float(step(*[t.to(dev, non_blocking=True) for t in sm_host])))
        small = {"rays_per_step": n_small * world, "ms_per_step": ms_dev, "rays_per_s": n_small * world / (ms_dev / 1e3),
                 "e2e": {"ms_per_step": ms_e2e_small, "value": n_small * world / (ms_e2e_small / 1e3), "unit": UNIT,
                         "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in sm_host), "d2h_bytes_per_step": 8}}
        if world == 1:
            # the same iteration captured once into a CUDA graph (attentive_dfprior_b200.graphs) and replayed
            try:
                from attentive_dfprior_b200.graphs import GraphedStep
                zero_grads()
                gs = GraphedStep(lambda ro, rd, g, gc: step(ro, rd, g, gc), static_inputs=[t.clone() for t in sm_dev])
                ms_g = time_loop(lambda: gs(*sm_dev))
                ms_g_e2e = time_loop(lambda: float(gs(*[t.to(dev, non_blocking=True) for t in sm_host])))
                small["cuda_graph"] = {"ms_per_step": ms_g, "rays_per_s": n_small / (ms_g / 1e3),
                                       "e2e": {"ms_per_step": ms_g_e2e, "value": n_small / (ms_g_e2e / 1e3), "unit": UNIT}}
                del gs
            except Exception as e:
                small["cuda_graph"] = {"unavailable": repr(e)[:300]}
        if world == 1:
            # Tracker pose step (src/Tracker.py:83-133): 200 pixels, frozen decoders and grids, gradient to the camera 7-vector
            from attentive_dfprior_b200 import common as CM
            cam = scene.cam
            tr_model = A.DF().to(dev)
            for prm in tr_model.parameters():
                prm.requires_grad_(False)
            tgrids = {k: v.detach() for k, v in scene.grids.items()}
            g = torch.Generator().manual_seed(11)
            pi = torch.randint(100, cam["W"] - 100, (200,), generator=g).float()
            pj = torch.randint(100, cam["H"] - 100, (200,), generator=g).float()
            c2w0 = synth.make_pose(scene, 3)
            with torch.no_grad():
                ro0, rd0 = CM.get_rays_from_uv(pi, pj, c2w0, cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], "cpu")
                tgt = synth.depth_along(scene, ro0, rd0)
            tcol = torch.rand(200, 3, generator=g)
            quat = torch.tensor([1.0, 0.0, 0.0, 0.0])
            cam7 = torch.cat([quat, c2w0[:3, 3]]).to(dev).requires_grad_(True)
            R0 = c2w0[:3, :3].to(dev)
            tr_host = [t.pin_memory() for t in (pi, pj, tgt, tcol)]

            def tracker_step(pi_, pj_, tg_, tc_):
                cam7.grad = None
                c2w = CM.get_camera_from_tensor(cam7)
                c2w = torch.cat([R0 @ c2w[:3, :3], c2w[:3, 3:]], 1)
                ro, rd = CM.get_rays_from_uv(pi_, pj_, c2w, cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], dev)
                d, u, c, _ = renderer.render_batch_ray(tgrids, tr_model, rd, ro, dev, scene.tsdf_volume, scene.tsdf_bnds, STAGE, gt_depth=tg_)
                u = u.detach()
                tmp = torch.abs(tg_ - d) / torch.sqrt(u + 1e-10)
                m = ((tmp < 10 * tmp.median()) & (tg_ > 0)).to(tmp.dtype)
                # src/Tracker.py:116-129; the sums over the selected rays written as masked sums (same values, no host sync)
                loss = (tmp * m).sum() + 0.5 * (torch.abs(tc_ - c) * m[:, None]).sum()
                loss.backward()
                return cam7.grad

            tr_dev = [t.to(dev) for t in tr_host]
            ms_tr = time_loop(lambda: tracker_step(*tr_dev))
            ms_tr_e2e = time_loop(lambda: tracker_step(*[t.to(dev, non_blocking=True) for t in tr_host]).cpu())
            tracker = {"rays_per_step": 200, "ms_per_step": ms_tr, "rays_per_s": 200 / (ms_tr / 1e3),
                       "e2e": {"ms_per_step": ms_tr_e2e, "value": 200 / (ms_tr_e2e / 1e3), "unit": UNIT,
                               "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in tr_host), "d2h_bytes_per_step": 28},
                       "note": "pose chain and median-masked loss are torch ops (autograd carries d rays -> 7-vector); render fwd+bwd are the kernels"}
            try:
                from attentive_dfprior_b200.graphs import GraphedStep
                gs = GraphedStep(tracker_step, static_inputs=[t.clone() for t in tr_dev])
                ms_g = time_loop(lambda: gs(*tr_dev))
                ms_g_e2e = time_loop(lambda: gs(*[t.to(dev, non_blocking=True) for t in tr_host]).cpu())
                tracker["cuda_graph"] = {"ms_per_step": ms_g, "rays_per_s": 200 / (ms_g / 1e3),
                                         "e2e": {"ms_per_step": ms_g_e2e, "value": 200 / (ms_g_e2e / 1e3), "unit": UNIT}}
                del gs
            except Exception as e:
                tracker["cuda_graph"] = {"unavailable": repr(e)[:300]}

    out = None
    if rank == 0:
        peak, peak_src = load_peaks()
        ns = n * S
        total_kernel_ms = sum(v[1] for v in kt.values())
        # The roofline object describes the dominant OPERATION of the step.  The decoder backward is one logical op that
        # this library runs as two kernels per decoder on the tensor-core path (k_decoder_dh_tc: delta chain + grid
        # scatter, k_decoder_dw_tc: weight gradients) or as one (k_decoder_bwd) on the SIMT path; its algorithmic bytes
        # are SURVEY 8d's backward figure (re-gather + gradient scatter).
        ops = {
            "decoder_bwd": (["k_decoder_bwd", "k_decoder_dh_tc", "k_decoder_dw_tc", "k_decoder_dh_tc_all", "k_decoder_dw_tc_all"], {"low": 2048, "high": 3072, "color": 2048}),
            "decoder_fwd": (["k_decoder_fwd", "k_decoder_fwd_tc", "k_decoder_fwd_tc_all"], {"low": 1024, "high": 2048, "color": 1024}),
            "attention": (["k_attn_fwd", "k_attn_bwd", "k_attn_fwd_tc", "k_attn_dh_tc", "k_attn_dw_tc"], {"low": 0, "high": 0, "color": 48 + 60}),
        }
        op_ms = {o: sum(kt[k][1] for k in ks if k in kt) / prof_steps for o, (ks, _) in ops.items()}
        top_op = max(op_ms, key=op_ms.get)
        kernels, kb = ops[top_op]
        bytes_per_step = sum(kb.values()) * ns                       # the three decoders of stage `color`
        achieved = bytes_per_step / (op_ms[top_op] / 1e3) / 1e9 if op_ms[top_op] > 0 else 0.0
        step_gbs = value / world * BYTES_PER_RAY[STAGE] / 1e9
        traffic, traffic_note = ncu_traffic(kernels) if (n == 65536 and world == 1) else (None, "only captured for the default 1-GPU workload")
        # arithmetic side: >= 333 kFLOP per sample fwd+bwd (MLP MACs x 2, minimum-needed count)
        flop_per_sample = 333e3
        tflops = value / world * S * flop_per_sample / 1e12
        roofline = {"bound": "hbm", "kernel": top_op + " (" + " + ".join(k for k in kernels if k in kt) + ")",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_note,
                    "peak_source": peak_src,
                    "kernel_share_of_step": op_ms[top_op] * prof_steps / total_kernel_ms if total_kernel_ms else None,
                    "algorithmic_bytes_per_step": bytes_per_step, "op_ms_per_step": op_ms[top_op],
                    "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak,
                                   "bytes_per_ray": BYTES_PER_RAY[STAGE]},
                    "flops": {"algorithmic_flop_per_sample": flop_per_sample, "achieved_tflops": tflops,
                              "frac_of_fp32_simt_peak": tflops / 80.0, "fp32_simt_peak_tflops": 80.0,
                              "frac_of_tf32_tensor_peak": tflops / 1100.0, "tf32_tensor_peak_tflops": 1100.0,
                              "note": "SURVEY 8d minimum-needed MLP flops; peaks are nominal B200 figures (no measured fp32/tf32 peak in MEASURED_PEAKS.json); the kernels spend 3 tensor-core MMAs per algorithmic one (3xTF32 / bf16 hi+mid split)"},
                    "kernels_ms_per_step": {k: round(v[1] / prof_steps, 4) for k, v in sorted(kt.items())}}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(n, world, tsdf_shape=list(scene.tsdf_volume.shape[2:])),
            "impl_details": {
                "grid_layout": "channels_last_3d" if not args.channels_first else "channels_first (transposed per call)",
                "collective": ("NCCL all-reduce (sum) of the grid + decoder gradients in 3 buckets on a side stream, each issued when the "
                               "backward has finished it" if not args.no_overlap else "one blocking NCCL all-reduce after the backward") if world > 1 else None,
                "mlp": "tcgen05 tensor-core kernels", "kernel_source_hash": kernel_source_hash(),
                "loss_impl": "torch ops" if args.torch_loss else "mapper_ops.mapper_loss kernel"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        }
        if frozen is not None:
            out["mapping_frozen_low_high"] = frozen
        if small is not None:
            out["mapping_1000_rays"] = small
        if tracker is not None:
            out["tracker_200_rays"] = tracker
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        del flush
        scene_cpu = synth.make_scene("room0", seed=0, high_std=1e-2)
        if not args.no_gpu_baseline:
            try:
                out["torch_gpu_baseline"] = torch_gpu_baseline(scene_cpu, n)
            except Exception as e:                      # e.g. out of memory on a smaller GPU: report, do not fail the bench
                out["torch_gpu_baseline"] = {"unavailable": repr(e)[:200]}
            torch.cuda.empty_cache()
        out["cpu_baseline"] = cpu_baseline(scene_cpu, n)
    if rank == 0:
        print(json.dumps(out))
    if group is not None:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=65536, help="rays per GPU per step")
    ap.add_argument("--ref-rays", type=int, default=2048, help="reference arm: rays per step sampled from the --rays batch")
    ap.add_argument("--device", default="cpu", help="reference arm: cpu (default) or cuda:0 (the reference's ops on the GPU)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the torch-on-GPU comparator of our arm")
    ap.add_argument("--channels-first", action="store_true", help="keep the grids channels-first (transposed per call)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-small", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: one blocking gradient all-reduce instead of the bucketed overlap")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = --rays per GPU (default, the driver's contract); strong = --rays in total, split over the ranks")
    ap.add_argument("--torch-loss", action="store_true", help="Mapper loss as torch ops (boolean indexing, host sync) instead of the loss kernel")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
