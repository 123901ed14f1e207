#!/usr/bin/env python
"""bench.py — frame-pairs/s of the photometric loss path, forward + backward.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU loss path (oracle port)

Workload (BASELINE.json configs[1]): flow-mode 4-level pyramid loss (warp + soft occlusion + L1 + SSIM +
second-order smoothness + fwd/bwd consistency), 256x832, batch 8 per GPU (weak scaling), synthetic
KITTI-shaped inputs.  One training sample (l, c, r) = 2 frame pairs.  A step = fused forward + finalize +
backward (d total / d flows) over one batch.

Prints ONE JSON line (rank 0).  `value` = inputs resident in HBM, L2 flushed between steps, timed with
CUDA events on the launching stream, max over ranks.  `e2e` = the same step through the host-buffer API
(H2D of frames + flows from pinned memory, pyramid build, forward, backward, D2H of the losses).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, LEVELS, BATCH = 256, 832, 4, 8
FLOW_PX = 10.0
METRIC = "frame_pairs_per_sec_loss_fwd_bwd"
UNIT = "frame-pairs/s"
WORKLOAD = "flow-mode 4-level pyramid loss fwd+bwd, synthetic 256x832, batch 8 per GPU"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(batch: int, fwd: bool, bwd: bool) -> int:
    """per sample and level fwd reads 13 N floats (3 images x 3 ch + 2 flows x 2 ch);
    bwd re-reads 13 N and writes 4 N gradient floats."""
    n = sum((H >> l) * (W >> l) for l in range(LEVELS))
    floats = (13 * n if fwd else 0) + (17 * n if bwd else 0)
    return floats * 4 * batch


class ClockSampler:
    """SM clock / throttle reasons sampled WHILE the timed region runs: in-process NVML polling on a background thread
    (every ~2 ms; the timed region is only tens of milliseconds, shorter than nvidia-smi's start-up), nvidia-smi as fallback."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, cuda_index: int):
        self.cuda_index, self.samples, self.mask, self.max_mhz = cuda_index, [], 0, None
        self._stop = threading.Event()
        self._thread, self._proc, self._nvml = None, None, None

    def _handle(self):
        import pynvml
        pynvml.nvmlInit()
        self._nvml = pynvml
        try:
            uuid = str(torch.cuda.get_device_properties(self.cuda_index).uuid)
            return pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.cuda_index]) if vis and vis.split(",")[self.cuda_index].isdigit() else self.cuda_index
            return pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _poll(self, h):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                except Exception:
                    pass
            time.sleep(0.002)

    def __enter__(self):
        try:
            h = self._handle()
            self.max_mhz = float(self._nvml.nvmlDeviceGetMaxClockInfo(h, self._nvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._poll, args=(h,), daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None
            try:
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                self._proc = subprocess.Popen(["nvidia-smi", "-i", str(self.cuda_index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                               "-lms", "20"], stdout=subprocess.PIPE, text=True)
                self._thread = threading.Thread(target=self._read_smi, daemon=True)
                self._thread.start()
            except Exception:
                self._proc = None
        return self

    def _read_smi(self):
        for line in self._proc.stdout:
            c = [x.strip() for x in line.split(",")]
            try:
                self.samples.append(float(c[0])); self.max_mhz = float(c[1])
            except Exception:
                continue
            for (name, bit), v in zip(self.REASONS, c[2:6]):
                if v.lower().startswith("active"):
                    self.mask |= bit

    def __exit__(self, *a):
        self._stop.set()
        if self._proc:
            self._proc.terminate()
        if self._thread:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        busy = [c for c in self.samples if c >= 0.5 * max(self.samples)]
        return {"sm_mhz": float(statistics.median(busy)), "sm_max_mhz": self.max_mhz, "samples": len(self.samples),
                "reasons": [name for name, bit in self.REASONS if self.mask & bit]}


def cpu_reference_rate(sample_batch: int, steps: int, warmup: int, threads: int, seed: int = 1234):
    """The reference's CPU loss path (oracle port of model_flow.py:232-254) on the host cores: same synthetic inputs as the GPU
    arm (same generator, seed and flow amplitude), loss forward + backward on materialised image pyramids."""
    from oracle import loss_port as P
    from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
    torch.set_num_threads(threads)
    t = make_triplet(sample_batch, H, W, LEVELS, 1, seed=seed, flow_px=FLOW_PX)
    pyr = tuple(P.box_pyramid(x, LEVELS) for x in (t.img_l, t.img, t.img_r))
    keys = list(P.FLOW_WEIGHTS)
    times = []
    for it in range(warmup + steps):
        for f in t.flows_fwd + t.flows_bwd:
            f.requires_grad_(True); f.grad = None
        t0 = time.perf_counter()
        loss = P.flow_mode_loss(t.img_l, t.img, t.img_r, t.flows_fwd, t.flows_bwd, LEVELS, pyramids=pyr)
        total = sum(P.FLOW_WEIGHTS[k] * loss[k].mean() for k in keys)
        total.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return 2.0 * sample_batch / statistics.mean(times), statistics.mean(times)


def flow_config(batch: int, world: int, launch: str):
    """The `config` object of the contract line: identical keys and values in the GPU arm and the reference arm."""
    return {"workload": WORKLOAD, "height": H, "width": W, "levels": LEVELS, "batch_per_gpu": batch, "global_batch": batch * world,
            "flows": "blurred N(0,1) scaled to %.0f px at level 0 (make_triplet flow_mode='noise', seed 1234 + rank)" % FLOW_PX,
            "launch": launch, "l2": "flushed between steps (256 MiB memset outside the event pair)",
            "timing": "sum of per-step CUDA-event pairs on the launching stream, max over ranks"}


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path (the oracle port: the Python reference cannot travel to
    the GPU box) on all host cores, on the GPU arm's workload: same batch, inputs, warm-up and config."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warm = max(1, min(args.steps, 30)), max(args.warmup, 3)
    rate, sec = cpu_reference_rate(args.batch, steps, warm, cores)
    cfg = flow_config(args.batch, args.gpus, "cpu (torch, %d threads)" % cores)
    cfg["l2"] = "n/a (CPU)"
    cfg["timing"] = "time.perf_counter around forward + backward, mean over the timed steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "the full batch-%d step (one rank's share of the workload), %d timed steps after %d warm-ups "
                                   "(oracle/loss_port.py, torch CPU, %d threads)" % (args.batch, steps, warm, cores)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_torch_cuda(args, rank, local):
    """Extra comparator: the reference's op sequence (oracle port = F.grid_sample / avg_pool2d / softmax ... under autograd)
    executed by stock PyTorch on the same B200, same workload and batch as the driver's line, eager launches."""
    if rank != 0:
        return
    from oracle import loss_port as P
    from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    t = make_triplet(args.batch, H, W, LEVELS, 1, seed=1234, flow_px=FLOW_PX).to(dev)
    flows = [f.requires_grad_(True) for f in t.flows_fwd + t.flows_bwd]

    def step():
        loss = P.flow_mode_loss(t.img_l, t.img, t.img_r, t.flows_fwd, t.flows_bwd, LEVELS)
        total = sum(P.FLOW_WEIGHTS[k] * loss[k].mean() for k in loss)
        return torch.autograd.grad(total, flows)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"impl": "torch-cuda", "metric": METRIC, "value": 2.0 * args.batch / (ms * 1e-3), "unit": UNIT, "n_gpus": 1,
                      "steps": args.steps, "ms_per_step": ms, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": WORKLOAD, "launch": "eager (stock PyTorch ops, autograd)"}}), flush=True)


GEOM_WEIGHTS = {"loss_flow_pixel": 0.15, "loss_flow_ssim": 0.85, "loss_flow_smooth": 10.0, "loss_flow_consis": 0.01, "loss_depth_pixel": 1.0,
                "loss_depth_ssim": 0.85, "loss_depth_smooth": 0.5, "loss_depth_consis": 0.1, "loss_depth_flow_consis": 1.0, "loss_epipolar": 0.1,
                "loss_triangle": 0.001, "loss_pnp": 0.1, "loss_eight_point": 0.1}          # config/kitti_geom.yaml:20-34


def mode_algorithmic_bytes(workload: str, height: int, width: int, scales: int = 3) -> int:
    """SURVEY 8(d) per-sample figures: depth 2 (12 N0 + 18 sum_{s>=1} N_s) + 3 sum N_s floats; geom (also used for the flow+depth
    stress config) 2 (16 N0 + 22 sum_{s>=1} N_s) + 7 sum N_s floats."""
    n = [(height >> l) * (width >> l) for l in range(scales)]
    if workload.startswith("depth"):
        return 4 * (2 * (12 * n[0] + 18 * sum(n[1:])) + 3 * sum(n))
    return 4 * (2 * (16 * n[0] + 22 * sum(n[1:])) + 7 * sum(n))


def mode_step(workload: str, batch: int, height: int, width: int, dev, steps: int, warm: int, flush, rank: int = 0, graph: bool = True,
              world: int = 1, dist=None):
    """One of the mode steps (BASELINE configs[2..4]) through ``losses.*.forward_losses`` (fused kernels) under autograd +
    ``losses.total_loss`` (train.py:211-214): CUDA-graph replay, L2 flushed before every step, per-step event pairs.  The synthetic
    triplet is generated once at batch <= 8 and tiled to ``batch`` (every sample is an independent unit of the same size)."""
    from unsupervised_depth_opticalflow_egomotion_b200 import losses, ops
    from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
    S = 3
    base_b = min(batch, 8)
    t = make_triplet(base_b, height, width, 4, S, seed=1234 + rank, flow_mode="rigid").to(dev)
    if batch != base_b:
        reps = (batch + base_b - 1) // base_b
        tile = lambda x: x.repeat(reps, *([1] * (x.dim() - 1)))[:batch].contiguous()
        for name in ("img_l", "img", "img_r", "pose", "K", "K_inv"):
            setattr(t, name, tile(getattr(t, name)))
        for name in ("flows_fwd", "flows_bwd", "disp", "disp_l", "disp_r"):
            setattr(t, name, [tile(x) for x in getattr(t, name)])
    leaves = [x.requires_grad_(True) for x in t.flows_fwd + t.flows_bwd + t.disp + t.disp_l + t.disp_r + [t.pose]]
    if workload == "geom":
        mod = losses.GeometryLoss(S)
        # training-step form: the weights of the total are declared up front, the flow branch writes its gradients in its forward launches
        fwd = lambda: mod.forward_losses(t.img_l, t.img, t.img_r, t.flows_fwd, t.flows_bwd, t.disp, t.disp_l, t.disp_r, t.pose, t.K, t.K_inv,
                                         step_weights=GEOM_WEIGHTS)[0]
    elif workload in ("depth", "depth-texture", "depth-live"):
        # depth = BASELINE configs[2] (model_depth_texture.py:296-307, reprojection L1 + SSIM + smoothness);
        # depth-texture = that file's whole loss (+ the depth-consistency term, :309-310); depth-live = model_depth.py
        mod = losses.DepthLoss(S, {"depth": "ssim", "depth-texture": "texture", "depth-live": "live"}[workload])
        fwd = lambda: mod.forward_losses(t.img_l, t.img, t.img_r, t.disp, t.disp_l, t.disp_r, t.pose, t.K)[0]
    else:   # flow+depth (BASELINE configs[4]): the flow-mode loss (4 levels) and the live depth-mode loss on the same triplet
        fmod, dmod = losses.FlowLoss(4), losses.DepthLoss(S, "live")
        fwd = lambda: {**fmod.forward_losses(t.img_l, t.img, t.img_r, t.flows_fwd, t.flows_bwd),
                       **dmod.forward_losses(t.img_l, t.img, t.img_r, t.disp, t.disp_l, t.disp_r, t.pose, t.K)[0]}

    one = torch.ones((), device=dev, dtype=torch.float32)     # the seed of backward(): allocated once, no fill launch per step

    def step():
        for x in leaves:
            x.grad = None
        losses.total_loss(fwd(), GEOM_WEIGHTS).backward(one)

    n0 = ops.LAUNCH_COUNTER["n"]
    step()
    launches = ops.LAUNCH_COUNTER["n"] - n0
    run, how = step, "eager"
    if graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for x in leaves:
                x.grad = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                losses.total_loss(fwd(), GEOM_WEIGHTS).backward(one)
            run, how = g.replay, "cuda-graph"
        except Exception as e:
            sys.stderr.write("graph capture failed, eager: %r\n" % (e,))
            torch.cuda.synchronize()
    for _ in range(max(warm, 3)):
        run(); flush.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / steps
    if world > 1:       # no data-path collective: the job's step time is the slowest rank's
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    peak, _ = _peaks()
    bytes_step = mode_algorithmic_bytes(workload, height, width) * batch
    return {"ms_per_step": ms, "frame_pairs_per_sec": 2.0 * batch * world / (ms * 1e-3), "batch_per_gpu": batch, "global_batch": batch * world,
            "height": height, "width": width, "launch": how, "launches_per_step": launches, "steps": steps, "l2": "flushed before every step",
            "algorithmic_bytes_per_step_per_gpu": bytes_step, "achieved_gbs": bytes_step / (ms * 1e-3) / 1e9,
            "hbm_frac": bytes_step / (ms * 1e-3) / 1e9 / peak}


def run_mode_workload(args, rank, world, local):
    """Extra measurement as its own line: ``--workload depth|depth-live|geom|flow+depth`` at ``--batch`` per GPU."""
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    r = mode_step(args.workload, args.batch, args.height, args.width, dev, args.steps, args.warmup, flush, rank, not args.no_graph, world, dist)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"metric": METRIC, "value": r["frame_pairs_per_sec"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                          "ms_per_step": r["ms_per_step"], "dtype": "f32", "data": "synthetic", "gpu_launches": r["launches_per_step"] * args.steps,
                          "config": {"workload": "%s-mode loss body fwd+bwd (fused kernels under autograd + the trainer's weighted total), %dx%d, "
                                                 "batch %d per GPU, S=3" % (args.workload, args.height, args.width, args.batch), **r}}), flush=True)


def run_cost_volume(args, rank, local):
    """Extra measurement: the ten PWC-Net cost volumes of one flow-mode step (5 pyramid levels x 2 directions,
    pwc_tf.py:112-160) forward + backward at batch 8, ours vs the reference's corr_naive op sequence executed by PyTorch on the same GPU."""
    import torch.nn.functional as F
    from unsupervised_depth_opticalflow_egomotion_b200 import ops
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    shapes = [(196, 4, 13), (128, 8, 26), (96, 16, 52), (64, 32, 104), (32, 64, 208)]          # c16 .. c12 at 256x832
    feats = [(torch.randn(B, c, h, w, generator=g).to(dev).requires_grad_(True), torch.randn(B, c, h, w, generator=g).to(dev).requires_grad_(True))
             for c, h, w in shapes for _ in range(2)]
    gos = [torch.randn(B, 81, a.shape[2], a.shape[3], generator=g).to(dev) for a, _ in feats]

    def naive(f1, f2, d=4):                      # the reference's op sequence (pwc_tf.py:97-106), stock PyTorch
        Hh, Ww = f1.shape[2:]
        f2 = F.pad(f2, (d, d, d, d), value=0)
        return torch.cat([(f1 * f2[:, :, i:i + Hh, j:j + Ww]).mean(1).unsqueeze(1) for i in range(2 * d + 1) for j in range(2 * d + 1)], 1)

    def step(fn):
        outs = [fn(a, b) for a, b in feats]
        torch.autograd.grad([o for o in outs], [t for ab in feats for t in ab], grad_outputs=gos)

    res = {}
    for name, fn in (("ours", ops.cost_volume), ("torch_naive", naive)):
        run = lambda: step(fn)
        for _ in range(max(args.warmup, 3)):
            run()
        torch.cuda.synchronize()
        if name == "ours" and not args.no_graph:          # 30 launches of 20-100 us each: replay them as one CUDA graph
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step(fn)
            torch.cuda.current_stream().wait_stream(side)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                step(fn)
            run = gr.replay
            run()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            run()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / args.steps
    macs = 3 * 81 * B * sum(c * h * w for c, h, w in shapes) * 2
    if rank == 0:
        print(json.dumps({"metric": "cost_volume_ms_per_step", "value": res["ours"], "unit": "ms", "higher_is_better": False, "n_gpus": 1,
                          "steps": args.steps, "torch_naive_ms": res["torch_naive"], "speedup_vs_torch_naive": res["torch_naive"] / res["ours"],
                          "gmac_per_step": macs / 1e9, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "PWC-Net cost volumes of one step: 5 levels x 2 directions, fwd+bwd, batch %d" % B,
                                     "launch": "ours: cuda-graph replay of 20 launches; torch_naive: eager (GPU-bound)"}}),
              flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch-cuda"],
                    help="reference = the reference's CPU loss path (oracle port); torch-cuda = the same torch op sequence on the GPU "
                         "(extra comparator: 'stock PyTorch on B200', SURVEY 8(d))")
    ap.add_argument("--batch", type=int, default=BATCH, help="samples per GPU")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the BASELINE configs[2..4] extras of the contract line")
    ap.add_argument("--workload", default="flow", choices=["flow", "depth", "depth-texture", "depth-live", "geom", "flow+depth", "costvolume"],
                    help="flow = BASELINE configs[1] (the driver's line).  Extras: depth (model_depth_texture spec) / depth-live (model_depth) "
                         "= configs[2], geom = configs[3], flow+depth = configs[4] (use --height 384 --width 1280)")
    ap.add_argument("--height", type=int, default=H, help="extras only (the driver's line is always 256x832)")
    ap.add_argument("--width", type=int, default=W)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.impl == "torch-cuda":
        run_torch_cuda(args, rank, local)
        return
    if args.workload == "costvolume":
        run_cost_volume(args, rank, local)
        return
    if args.workload != "flow":
        run_mode_workload(args, rank, world, local)
        return

    import torch.distributed as dist
    from unsupervised_depth_opticalflow_egomotion_b200 import _cabi, ops, parallel, build as ugl_build
    from unsupervised_depth_opticalflow_egomotion_b200.step import FlowLossStep, FLOW_WEIGHTS, weight_matrix
    from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the loss path has no CPU implementation); "
                         "use --impl reference for the CPU baseline")
    if rank == 0:
        ugl_build.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = parallel.bind_to_gpu_numa(local)          # before any pinned allocation: first touch decides the NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    _cabi.lib()
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)

    # ---- inputs resident in HBM -------------------------------------------------------------------------
    host = make_triplet(B, H, W, LEVELS, 1, seed=1234 + rank, flow_px=FLOW_PX)
    t = host.to(dev)
    pl, pc, pr = (ops.image_pyramid(x, LEVELS, "box") for x in (t.img_l, t.img, t.img_r))
    ff = [f.requires_grad_(True) for f in t.flows_fwd]
    fb = [f.requires_grad_(True) for f in t.flows_bwd]
    wmat = weight_matrix(FLOW_WEIGHTS, ops.FLOW_LOSS_KEYS, B, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    state = {"out": None}

    def step():
        # the fused training step with d total / d loss = w_k / B (train.py:211-215): photometry, weight sums, stencil (writes
        # d total / d flow), finalize (losses) -- 4 launches
        state["out"] = ops.flow_loss_step(pl, pc, pr, ff, fb, wmat, LEVELS, out=state["out"])
        return state["out"]

    n0 = ops.LAUNCH_COUNTER["n"]
    res = step()
    launches_per_step = ops.LAUNCH_COUNTER["n"] - n0
    torch.cuda.synchronize()
    # the autograd surface (forward_grad + combine through the basis planes) must give the same numbers as the fused step
    chk = ops.flow_loss(pl, pc, pr, ff, fb, LEVELS, as_matrix=True)
    chk_g = torch.autograd.grad(chk, ff + fb, grad_outputs=wmat)
    rel = lambda a, b: float((a.detach() - b.detach()).abs().max() / b.detach().abs().max().clamp_min(1e-30))
    assert rel(res["loss"], chk) < 1e-6 and all(rel(a, b) < 5e-6 for a, b in zip(res["gf"] + res["gb"], chk_g))

    def capture(fn):
        if args.no_graph:
            return None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            return g
        except Exception as e:   # pragma: no cover - reported in the JSON line
            sys.stderr.write("CUDA graph capture failed, timing eager launches: %r\n" % (e,))
            torch.cuda.synchronize()
            return None

    graph = capture(step)
    run = graph.replay if graph is not None else step

    def timed(fn, n, warm):
        for _ in range(warm):
            fn(); flush.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        for _ in range(n):
            flush.zero_()                         # L2 flush, outside the event pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    kernel_ms = {}
    with ClockSampler(local) as clk:
        ms = timed(run, K, Wm)
        # per-kernel timing (roofline of the dominant kernel): each launch of the step replayed on its own, L2 flushed before it,
        # CUDA events on the launching stream.  The kernels read what the full step left in the workspace.
        for part in ("photo", "norm", "stencil", "finalize"):
            fn = (lambda ph=part: ops.flow_loss_step(pl, pc, pr, ff, fb, wmat, LEVELS, out=state["out"], phase=ph))
            gpart = capture(fn)
            kernel_ms[part] = statistics.mean(timed(gpart.replay if gpart is not None else fn, min(K, 20), 2))
        run()
    clocks = clk.summary()
    total_ms = sum(ms)

    # ---- e2e: host buffers -> losses on the host, through the public step API ------------------------------
    n_e2e = max(5, min(K, 30))

    def fill(stepper, frames):
        for slot in (0, 1):
            hv = stepper.host_views(slot)
            for name, src in zip(("img_l", "img", "img_r"), frames):
                hv[name].copy_(src)
            for l in range(LEVELS):
                hv["ff%d" % l].copy_(host.flows_fwd[l].detach()); hv["fb%d" % l].copy_(host.flows_bwd[l].detach())

    def e2e_run(stepper, n):
        # pipelined: the ONE H2D copy of step k+1 (copy stream) overlaps the kernels of step k; every step's H2D and D2H lie
        # inside the timed region, one event pair around the n steps, results read back step by step.
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        prev = None
        for _ in range(n):
            cur = stepper.submit_staged()
            if prev is not None:
                stepper.result(prev)
            prev = cur
        stepper.result(prev)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def h2d_probe(stepper, n):
        # raw pinned H2D bandwidth of this rank while every rank copies (no kernels): what the box can deliver to the path
        sl = stepper.slots[0]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            sl.d_raw.copy_(sl.h_raw, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return sl.nbytes * n / (e0.elapsed_time(e1) * 1e-3) / 1e9

    stepper = FlowLossStep(B, H, W, LEVELS, device=dev)
    fill(stepper, (host.img_l, host.img, host.img_r))
    e2e_run(stepper, 3)
    if world > 1:
        dist.barrier()
    e2e_total = e2e_run(stepper, n_e2e)
    h2d_raw = h2d_probe(stepper, 10)
    h2d_bytes, d2h_bytes = stepper.h2d_bytes, stepper.d2h_bytes
    del stepper

    # extra (not the contract value): the same step with the three frames shipped as uint8 and `img / 255.0` of the reference's
    # dataset (kitti_prepared.py:89) done on the device -- the frames are bytes at the source; 4x less PCIe traffic for them
    stepper = FlowLossStep(B, H, W, LEVELS, device=dev, frame_dtype=torch.uint8)
    fill(stepper, [(x * 255.0).round().clamp(0, 255).to(torch.uint8) for x in (host.img_l, host.img, host.img_r)])
    e2e_run(stepper, 3)
    if world > 1:
        dist.barrier()
    u8_total = e2e_run(stepper, n_e2e)
    u8_h2d = stepper.h2d_bytes
    del stepper

    # ---- the data-parallel step's collectives inside a timed region (N > 1) ---------------------------------------------------
    coll = None
    if world > 1:
        lar = parallel.LossAllReduce(4, B * world, dev)

        def step_with_allreduce():
            run()                       # the step's graph ...
            lar(state["out"]["loss"])   # ... then sum + NCCL all-reduce + scale on the same stream (launched eagerly: a captured NCCL
                                        # kernel ties the graph's lifetime to the communicator's and can hang the teardown)

        ms_lar = statistics.mean(timed(step_with_allreduce, min(K, 20), 3))
        bucket = parallel.GradBucketAllReduce(dev)

        def bucket_alone():
            bucket.launch(); bucket.join()

        def step_with_bucket():
            bucket.launch()          # the 86 MB parameter-gradient all-reduce of the surrounding DP step, on its own stream ...
            run()                    # ... under the loss step
            bucket.join()

        ms_bucket_alone = statistics.mean(timed(bucket_alone, 10, 3))
        ms_overlap = statistics.mean(timed(step_with_bucket, min(K, 20), 3))
        coll = [ms_lar, ms_bucket_alone, ms_overlap]

    # ---- BASELINE configs[2..4] as extras of the same line --------------------------------------------------------------------
    extras = {}
    if not args.no_extras:
        del graph
        state["out"] = None
        torch.cuda.empty_cache()
        gb32, gb64 = max(1, 32 // world), max(1, 64 // world)
        for name, wl, b_, h_, w_ in (("depth", "depth", B, H, W), ("depth_texture_full", "depth-texture", B, H, W),
                                     ("depth_live", "depth-live", B, H, W), ("geom", "geom", B, H, W),
                                     ("geom_strong_b32", "geom", gb32, H, W), ("highres_b64", "flow+depth", gb64, 384, 1280)):
            try:
                extras[name] = mode_step(wl, b_, h_, w_, dev, min(K, 10), 3, flush, rank, not args.no_graph, world, dist if world > 1 else None)
                extras[name]["scaling"] = "strong" if name in ("geom_strong_b32", "highres_b64") else "weak"
            except Exception as e:   # pragma: no cover
                extras[name] = {"error": repr(e)[:200]}
            torch.cuda.empty_cache()
        try:    # the contract workload on rigid flows (SURVEY 8(d) / appendix B's realistic variant): smooth flow fields gather coherently
            tr = make_triplet(B, H, W, LEVELS, LEVELS, seed=1234 + rank, flow_mode="rigid").to(dev)
            pyr_r = [ops.image_pyramid(x, LEVELS, "box") for x in (tr.img_l, tr.img, tr.img_r)]
            st_r = {"out": None}

            def step_r():
                st_r["out"] = ops.flow_loss_step(pyr_r[0], pyr_r[1], pyr_r[2], tr.flows_fwd, tr.flows_bwd, wmat, LEVELS, out=st_r["out"])
            g_r = capture(step_r)
            ms_r = statistics.mean(timed(g_r.replay if g_r is not None else step_r, min(K, 20), 3))
            extras["flow_rigid_flows"] = {"ms_per_step": ms_r, "frame_pairs_per_sec": 2.0 * B * world / (ms_r * 1e-3),
                                          "flows": "rigid flow of the synthetic disparity + blurred N(0,0.5^2) px"}
        except Exception as e:   # pragma: no cover
            extras["flow_rigid_flows"] = {"error": repr(e)[:200]}

    # ---- max over ranks ------------------------------------------------------------------------------------
    red = torch.tensor([total_ms, e2e_total / n_e2e, u8_total / n_e2e] + [kernel_ms[k] for k in ("photo", "norm", "stencil", "finalize")]
                       + (coll or []), device=dev, dtype=torch.float64)
    mn = torch.tensor([h2d_raw], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    total_ms, e2e_ms_per_step, u8_ms_per_step = float(red[0]), float(red[1]), float(red[2])
    kms = {k: float(red[3 + i]) for i, k in enumerate(("photo", "norm", "stencil", "finalize"))}
    ms_per_step = total_ms / K
    value = 2.0 * B * world / (ms_per_step * 1e-3)
    e2e_value = 2.0 * B * world / (e2e_ms_per_step * 1e-3)

    if rank == 0:
        peak, peak_src = _peaks()
        alg = algorithmic_bytes(B, True, True)
        step_gbs = alg / (ms_per_step * 1e-3) / 1e9
        dom = max(("photo", "stencil"), key=lambda k: kms[k])
        counters = {}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                counters = json.load(f)
        except Exception:
            pass
        kc = counters.get("kernels", {})
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        issue_rate = 148 * 4 * sm_hz                       # warp-instructions per second at one issue per SM sub-partition and clock
        per_kernel = {}
        for k in ("photo", "norm", "stencil", "finalize"):
            e = {"ms": kms[k]}
            c = kc.get({"photo": "flow_photo_kernel", "norm": "flow_photo_norm_kernel", "stencil": "flow_stencil_kernel",
                        "finalize": "flow_loss_finalize_kernel"}[k])
            if c:
                e["dram_bytes"] = c["dram_bytes_read"] + c["dram_bytes_write"]
                e["dram_gbs"] = e["dram_bytes"] / (kms[k] * 1e-3) / 1e9
                e["warp_instructions"] = c.get("warp_instructions")
                if c.get("warp_instructions"):
                    e["issue_frac"] = c["warp_instructions"] / issue_rate / (kms[k] * 1e-3)
                for extra_key in ("shared_wavefronts", "fma_pipe_pct", "issue_active_pct", "warps_active_pct"):
                    if extra_key in c:
                        e[extra_key] = c[extra_key]
                if c.get("shared_wavefronts"):
                    e["shared_pipe_frac"] = c["shared_wavefronts"] / (148 * sm_hz) / (kms[k] * 1e-3)
            per_kernel[k] = e
        traffic = sum(e.get("dram_bytes", 0) for e in per_kernel.values()) or None
        instr = sum((e.get("warp_instructions") or 0) for e in per_kernel.values())
        issue_frac_step = instr / issue_rate / (ms_per_step * 1e-3) if instr else None
        hbm_frac = step_gbs / peak
        bound = "issue" if (issue_frac_step or 0) > max(hbm_frac, (traffic or 0) / (ms_per_step * 1e-3) / 1e9 / peak) else "hbm"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": flow_config(B, world, "cuda-graph" if not args.no_graph else "eager"),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms_per_step,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "h2d_copies_per_step": 1,
                    "h2d_gbs_achieved": h2d_bytes / (e2e_ms_per_step * 1e-3) / 1e9, "h2d_gbs_raw": float(mn[0]),
                    "h2d_raw_how": "the slot's pinned buffer copied 10x with no kernels, all ranks at once; min over ranks",
                    "numa": numa,
                    "returns": "the (4,B) losses (128 B) come back to the host; d total / d flow stays on the device for the "
                               "network backward that consumes it",
                    "how": "step.FlowLossStep (staged): inputs in one pinned buffer -> ONE H2D -> pyramids -> fused fwd+bwd step -> D2H "
                           "losses; 2 staging slots, copy of step k+1 overlaps compute of step k; one event pair around %d steps" % n_e2e,
                    "uint8_frames": {"value": 2.0 * B * world / (u8_ms_per_step * 1e-3), "unit": UNIT, "ms_per_step": u8_ms_per_step,
                                     "h2d_bytes_per_step": u8_h2d,
                                     "how": "extra, not the contract value: FlowLossStep(frame_dtype=uint8) -- frames shipped as bytes, "
                                            "the dataset's img/255.0 evaluated on the device (bit-identical), flows fp32"}},
            "gpu_launches": launches_per_step * K,
            "roofline": {"bound": bound, "kernel": "flow_%s_kernel" % dom,
                         "achieved": alg / (kms[dom] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (kms[dom] * 1e-3) / 1e9 / peak,
                         "how": "contract form: the step's algorithmic bytes (SURVEY 8(d): 120 N bytes per sample and level) over the "
                                "dominant kernel's launch time; `step` below is the same bytes over the whole step (all four launches)",
                         "traffic": traffic, "traffic_over_algorithmic": (traffic / alg) if traffic else None,
                         "traffic_source": "profiles/traffic.json: dram__bytes_read + write of every kernel of the step, one ncu --set full "
                                           "capture of this tree (%s)" % counters.get("capture", "?"),
                         "peak_source": peak_src, "algorithmic_bytes_per_step": alg,
                         "step": {"achieved": step_gbs, "frac": hbm_frac, "ms": ms_per_step},
                         "issue": {"frac": issue_frac_step, "warp_instructions_per_step": instr or None,
                                   "rate": "148 SMs x 4 sub-partitions x %.0f MHz" % (sm_hz / 1e6),
                                   "how": "warp-instructions of the step's kernels (ncu smsp__inst_executed.sum, committed capture) / "
                                          "issue rate / measured step time"},
                         "kernels": per_kernel,
                         "timing": "CUDA events around a graph replay of each launch on the launching stream, L2 flushed before each"},
        }
        if coll:
            lar_ms, bucket_alone_ms, overlap_ms = float(red[7]), float(red[8]), float(red[9])
            line["collective"] = {
                "loss_allreduce": {"ms_per_step_with": lar_ms, "ms_per_step_without": ms_per_step, "exposed_ms": max(0.0, lar_ms - ms_per_step),
                                   "what": "4-float all-reduce (NCCL) of the per-term loss sums on the step's stream, behind the step's CUDA graph"},
                "grad_bucket_allreduce": {"bytes": 4 * 21_570_000, "ms_alone": bucket_alone_ms, "ms_step_overlapped": overlap_ms,
                                          "exposed_ms": max(0.0, overlap_ms - ms_per_step),
                                          "what": "stand-in for the DP step's parameter-gradient all-reduce (21.57 M fp32, 25 MB buckets) on its "
                                                  "own stream under the loss step"}}
        if extras:
            line["extras"] = extras
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            rate, sec = cpu_reference_rate(B, 8, 2, cores)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "the full batch-%d step, 8 timed steps after 2 warm-ups (oracle/loss_port.py, torch CPU, %d threads)" % (B, cores)}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
