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
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from attentive_dfprior_b200 import synth
    from oracle import dfprior_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = args.ref_rays
    scene = synth.make_scene("room0", seed=0, high_std=1e-2)
    params = {k: v.requires_grad_(True) for k, v in O.init_decoder_params(0).items()}
    grids = {k: v.requires_grad_(True) for k, v in scene.grids.items()}
    rays_o, rays_d, gt, gt_color = synth.make_rays(scene, n, seed=1)

    def step():
        for t in list(params.values()) + list(grids.values()):
            t.grad = None
        d, u, c, w = O.render_batch_ray(params, grids, rays_d, rays_o, scene.tsdf_volume, scene.tsdf_bnds, scene.bound,
                                        STAGE, gt)
        loss = O.mapper_loss(d, c, w, gt, gt_color, STAGE)
        loss.backward()
        return float(loss)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = f"{n} rays/step x {args.steps} steps, room0 shapes, stage {STAGE}, fwd+bwd, torch CPU {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "replica_room0_mapping_iteration", "stage": STAGE, "rays_per_step": n,
                   "samples_per_ray": S, "note": "oracle port of the reference PyTorch path on host cores"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(scene_cpu_factory, budget_s=20.0):
    """Bounded sample of the same workload through the oracle on the host cores (rank 0, N=1 only)."""
    from attentive_dfprior_b200 import synth
    from oracle import dfprior_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    scene = scene_cpu_factory()
    n = 1000
    params = {k: v.requires_grad_(True) for k, v in O.init_decoder_params(0).items()}
    grids = {k: v.requires_grad_(True) for k, v in scene.grids.items()}
    rays_o, rays_d, gt, gt_color = synth.make_rays(scene, n, seed=1)

    def step():
        for t in list(params.values()) + list(grids.values()):
            t.grad = None
        d, u, c, w = O.render_batch_ray(params, grids, rays_d, rays_o, scene.tsdf_volume, scene.tsdf_bnds, scene.bound,
                                        STAGE, gt)
        O.mapper_loss(d, c, w, gt, gt_color, STAGE).backward()

    step()
    t0 = time.perf_counter()
    reps = 0
    while reps < 3 or (time.perf_counter() - t0 < budget_s and reps < 20):
        step()
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": n * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{reps} x {n}-ray mapping steps (stage {STAGE}, fwd+bwd, room0 shapes) through the oracle, "
                      f"torch CPU {cores} threads"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
NCU_SUMMARY = "profiles/r01_ncu_full_summary_planes.csv"


def ncu_traffic(kernels):
    """DRAM bytes (read + write) of the named kernels' launches during ONE step of the default workload, from the
    committed `ncu --set full` capture of `bench.py --steps 1` (profiles/); None when the file is missing."""
    path = os.path.join(REPO, NCU_SUMMARY)
    if not os.path.exists(path):
        return None
    import csv
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    total = 0.0
    for r in rows[2:]:
        if any(k in r[ik] for k in kernels):
            total += float(r[ir]) * scale.get(units[ir], 1.0) + float(r[iw]) * scale.get(units[iw], 1.0)
    return total or None


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

    n = args.rays                                   # rays per GPU per step (weak scaling)
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
    del parts
    host = [t.cpu().pin_memory() for t in (rays_o, rays_d, gt, gt_color)]
    h2d = sum(t.numel() * t.element_size() for t in host)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def zero_grads():
        for t in params:
            t.grad = None
        for t in grids.values():
            t.grad = None

    def step(ro, rd, g, gc):
        zero_grads()
        A.decoder.param_pack_of(model).key = None                      # decoders change every mapping iteration: re-pack
        d, u, c, w = D.render_batch_ray_sharded(renderer, grids, model, rd, ro, scene.tsdf_volume, scene.tsdf_bnds,
                                                STAGE, g, group=group)
        if args.torch_loss:
            m = g > 0
            loss = torch.abs(g[m] - d[m]).sum() + 0.2 * torch.abs(gc - c).sum()      # src/Mapper.py:457-469 as torch ops
        else:
            loss = M.mapper_loss(d, c, w, g, gc, 0.2, False, True)                   # the same loss as one kernel (+ gradient)
        loss.backward()
        return loss

    def step_dev():
        return step(rays_o, rays_d, gt, gt_color)

    def step_e2e():
        dv = [t.to(dev, non_blocking=True) for t in host]
        return float(step(*dv))                                        # D2H of the loss closes the step

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
    # the reference's own batch size (1000 mapping pixels per step over all ranks) as a latency figure
    small = None
    if not args.no_small:
        n_small = max(1, 1000 // world)
        ro, rd, g, gc = synth.make_rays(scene, n_small, seed=7 + rank, device=dev)
        small_step = lambda: step(ro, rd, g, gc)
        for _ in range(5):
            small_step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            small_step()
        b.record()
        torch.cuda.synchronize()
        small = {"rays_per_step": n_small * world, "ms_per_step": a.elapsed_time(b) / 20,
                 "rays_per_s": n_small * world * 20 / (a.elapsed_time(b) / 1e3)}

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
            "decoder_bwd": (["k_decoder_bwd", "k_decoder_dh_tc", "k_decoder_dw_tc"], {"low": 2048, "high": 3072, "color": 2048}),
            "decoder_fwd": (["k_decoder_fwd", "k_decoder_fwd_tc"], {"low": 1024, "high": 2048, "color": 1024}),
            "attention": (["k_attn_fwd", "k_attn_bwd"], {"low": 0, "high": 0, "color": 48 + 60}),
        }
        op_ms = {o: sum(kt[k][1] for k in ks if k in kt) / prof_steps for o, (ks, _) in ops.items()}
        top_op = max(op_ms, key=op_ms.get)
        kernels, kb = ops[top_op]
        bytes_per_step = sum(kb.values()) * ns                       # the three decoders of stage `color`
        achieved = bytes_per_step / (op_ms[top_op] / 1e3) / 1e9 if op_ms[top_op] > 0 else 0.0
        step_gbs = value / world * BYTES_PER_RAY[STAGE] / 1e9
        roofline = {"bound": "hbm", "kernel": top_op + " (" + " + ".join(k for k in kernels if k in kt) + ")",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(kernels) if (n == 65536 and world == 1) else None,
                    "traffic_source": NCU_SUMMARY + " (dram read+write of these kernels' launches in one step, ncu --set full)",
                    "peak_source": peak_src,
                    "kernel_share_of_step": op_ms[top_op] * prof_steps / total_kernel_ms if total_kernel_ms else None,
                    "algorithmic_bytes_per_step": bytes_per_step, "op_ms_per_step": op_ms[top_op],
                    "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak,
                                   "bytes_per_ray": BYTES_PER_RAY[STAGE]},
                    "kernels_ms_per_step": {k: round(v[1] / prof_steps, 4) for k, v in sorted(kt.items())}}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "replica_room0_mapping_iteration", "stage": STAGE, "rays_per_gpu_per_step": n,
                       "samples_per_ray": S, "grids": "low[21,28,37] high/color[43,56,74] x32, " +
                       ("channels_last_3d" if not args.channels_first else "channels_first (transposed per call)"),
                       "tsdf": list(scene.tsdf_volume.shape[2:]), "l2": "flushed between timed steps (256 MiB write)",
                       "parallelism": f"ray-sharded dp{world}" + (" + NCCL all-reduce of grid/decoder grads" if world > 1 else ""),
                       "mlp": "tcgen05 3xTF32 tensor-core path" if os.environ.get("DFP_TENSOR_CORES", "1") != "0" else "fp32 SIMT",
                       "loss": "Mapper depth L1 + 0.2 colour L1" + (" (torch ops)" if args.torch_loss else " (mapper_ops.mapper_loss kernel)")},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        }
        if small is not None:
            out["mapping_1000_rays"] = small
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        del flush
        out["cpu_baseline"] = cpu_baseline(lambda: synth.make_scene("room0", seed=0, high_std=1e-2))
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
    ap.add_argument("--ref-rays", type=int, default=1000, help="rays per step of the CPU reference arm")
    ap.add_argument("--channels-first", action="store_true", help="keep the grids channels-first (transposed per call)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-small", action="store_true")
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
