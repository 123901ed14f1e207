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


def cpu_reference_rate(sample_batch: int, steps: int, warmup: int, threads: int):
    """The reference's CPU loss path (oracle port of model_flow.py:232-254) on the host cores."""
    from oracle import loss_port as P
    from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
    torch.set_num_threads(threads)
    t = make_triplet(sample_batch, H, W, LEVELS, 1, seed=1234)
    keys = list(P.FLOW_WEIGHTS)
    times = []
    for it in range(warmup + steps):
        for f in t.flows_fwd + t.flows_bwd:
            f.requires_grad_(True); f.grad = None
        t0 = time.perf_counter()
        loss = P.flow_mode_loss(t.img_l, t.img, t.img_r, t.flows_fwd, t.flows_bwd, LEVELS)
        total = sum(P.FLOW_WEIGHTS[k] * loss[k].mean() for k in keys)
        total.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return 2.0 * sample_batch / statistics.median(times), statistics.median(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_b = 1
    steps = max(1, min(args.steps, 20))
    rate, sec = cpu_reference_rate(sample_b, steps, min(args.warmup, 2), cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 2), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "height": H, "width": W, "levels": LEVELS, "global_batch": BATCH * args.gpus},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "batch %d of the batch-%d workload per step (oracle/loss_port.py, torch CPU, %d threads)" % (sample_b, BATCH, cores)},
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
    t = make_triplet(args.batch, H, W, LEVELS, 1, seed=1234, flow_px=10.0).to(dev)
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


def run_mode_workload(args, rank, world, local):
    """Extra measurement (not the driver's line): the depth- / geom-mode loss bodies (BASELINE configs[2] / [3]) through
    ``losses.*.forward_losses`` (fused kernels) under autograd + ``losses.total_loss`` (train.py:211-214), 256x832, batch 8 per GPU,
    S=3, CUDA-graph replay unless --no-graph."""
    from unsupervised_depth_opticalflow_egomotion_b200 import losses, ops
    from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, S = args.batch, 3
    H, W = args.height, args.width
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    t = make_triplet(B, H, W, 4, S, seed=1234 + rank, flow_mode="rigid").to(dev)
    leaves = [x.requires_grad_(True) for x in t.flows_fwd + t.flows_bwd + t.disp + t.disp_l + t.disp_r + [t.pose]]
    weights = {"loss_flow_pixel": 0.15, "loss_flow_ssim": 0.85, "loss_flow_smooth": 10.0, "loss_flow_consis": 0.01, "loss_depth_pixel": 1.0,
               "loss_depth_ssim": 0.85, "loss_depth_smooth": 0.5, "loss_depth_consis": 0.1, "loss_depth_flow_consis": 1.0, "loss_epipolar": 0.1,
               "loss_triangle": 0.001, "loss_pnp": 0.1, "loss_eight_point": 0.1}
    if args.workload == "geom":
        mod = losses.GeometryLoss(S)
        fwd = lambda: mod.forward_losses(t.img_l, t.img, t.img_r, t.flows_fwd, t.flows_bwd, t.disp, t.disp_l, t.disp_r, t.pose, t.K, t.K_inv)[0]
    elif args.workload in ("depth", "depth-live"):
        mod = losses.DepthLoss(S, "texture" if args.workload == "depth" else "live")
        fwd = lambda: mod.forward_losses(t.img_l, t.img, t.img_r, t.disp, t.disp_l, t.disp_r, t.pose, t.K)[0]
    else:   # flow+depth (BASELINE configs[4]): the flow-mode loss (4 levels) and the live depth-mode loss on the same triplet
        fmod, dmod = losses.FlowLoss(4), losses.DepthLoss(S, "live")
        fwd = lambda: {**fmod.forward_losses(t.img_l, t.img, t.img_r, t.flows_fwd, t.flows_bwd),
                       **dmod.forward_losses(t.img_l, t.img, t.img_r, t.disp, t.disp_l, t.disp_r, t.pose, t.K)[0]}

    def step():
        for x in leaves:
            x.grad = None
        loss = fwd()
        losses.total_loss(loss, weights).backward()

    n0 = ops.LAUNCH_COUNTER["n"]
    step()
    launches = ops.LAUNCH_COUNTER["n"] - n0
    run, how = step, "eager"
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for x in leaves:
                x.grad = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss = fwd()
                losses.total_loss(loss, weights).backward()
            run, how = g.replay, "cuda-graph"
        except Exception as e:
            sys.stderr.write("graph capture failed, eager: %r\n" % (e,))
            torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:       # weak scaling, no data-path collective: the job's step time is the slowest rank's
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"metric": METRIC, "value": 2.0 * B * world / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                          "ms_per_step": ms, "dtype": "f32", "data": "synthetic", "gpu_launches": launches * args.steps,
                          "config": {"workload": "%s-mode loss body fwd+bwd (fused kernels under autograd + the trainer's weighted total), %dx%d, "
                                                 "batch %d per GPU, S=3" % (args.workload, H, W, B), "height": H, "width": W, "launch": how,
                                     "timing": "one CUDA-event pair around all steps, max over ranks"}}), flush=True)


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
    ap.add_argument("--workload", default="flow", choices=["flow", "depth", "depth-live", "geom", "flow+depth", "costvolume"],
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
    from unsupervised_depth_opticalflow_egomotion_b200 import _cabi, ops, build as ugl_build
    from unsupervised_depth_opticalflow_egomotion_b200.step import FlowLossStep, FLOW_WEIGHTS, weight_matrix
    from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the loss path has no CPU implementation); "
                         "use --impl reference for the CPU baseline")
    if rank == 0:
        ugl_build.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    _cabi.lib()
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)

    # ---- inputs resident in HBM -------------------------------------------------------------------------
    host = make_triplet(B, H, W, LEVELS, 1, seed=1234 + rank, flow_px=10.0)
    t = host.to(dev)
    pl, pc, pr = (ops.image_pyramid(x, LEVELS, "box") for x in (t.img_l, t.img, t.img_r))
    ff = [f.requires_grad_(True) for f in t.flows_fwd]
    fb = [f.requires_grad_(True) for f in t.flows_bwd]
    wmat = weight_matrix(FLOW_WEIGHTS, ops.FLOW_LOSS_KEYS, B, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    state = {"out": None}

    def step():
        # forward + finalize + backward with d total / d loss = w_k / B (train.py:211-215): 3 launches
        state["out"] = ops.flow_loss_step(pl, pc, pr, ff, fb, wmat, LEVELS, out=state["out"])
        return state["out"]

    n0 = ops.LAUNCH_COUNTER["n"]
    res = step()
    launches_per_step = ops.LAUNCH_COUNTER["n"] - n0
    torch.cuda.synchronize()
    # the autograd surface must give the same numbers as the direct step
    chk = ops.flow_loss(pl, pc, pr, ff, fb, LEVELS, as_matrix=True)
    chk_g = torch.autograd.grad(chk, ff + fb, grad_outputs=wmat)
    assert torch.equal(chk, res["loss"]) and all(torch.equal(a, b) for a, b in zip(chk_g, res["gf"] + res["gb"]))

    graph = None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
        except Exception as e:   # pragma: no cover - reported in the JSON line
            graph = None
            sys.stderr.write("CUDA graph capture failed, timing eager launches: %r\n" % (e,))
    run = (lambda: graph.replay()) if graph is not None else (lambda: step())

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

    with ClockSampler(local) as clk:
        ms = timed(run, K, Wm)
        # per-launch timing for the roofline of the dominant kernel: the forward half (single-pass stencil kernel +
        # finalize) and the backward half (element-wise combine) replayed as separate CUDA graphs, L2 flushed before each
        stats_ms = {"fwd": [], "bwd": []}
        halves = {}
        for name in ("forward", "backward"):
            fn = (lambda ph=name: state.__setitem__("out_sp", ops.flow_loss_step(pl, pc, pr, ff, fb, wmat, LEVELS, out=state.get("out_sp"), phase=ph, mode="single_pass")))
            if graph is not None:
                gh = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    fn()
                torch.cuda.current_stream().wait_stream(side)
                with torch.cuda.graph(gh):
                    fn()
                halves[name] = gh.replay
            else:
                halves[name] = fn
        for _ in range(min(K, 30)):
            for name, key in (("forward", "fwd"), ("backward", "bwd")):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); halves[name](); e1.record()
                torch.cuda.synchronize()
                stats_ms[key].append(e0.elapsed_time(e1))
    clocks = clk.summary()
    total_ms = sum(ms)

    # ---- e2e: host buffers -> losses on the host, through the public step API ------------------------------
    pin = lambda x: x.pin_memory()
    h_imgs = [pin(host.img_l), pin(host.img), pin(host.img_r)]
    h_ff, h_fb = [pin(f.detach()) for f in host.flows_fwd], [pin(f.detach()) for f in host.flows_bwd]
    stepper = FlowLossStep(B, H, W, LEVELS, device=dev)
    n_e2e = max(5, min(K, 30))

    def e2e_run(n):
        # pipelined: the H2D copy of step k+1 (copy stream) overlaps the kernels of step k; every step's H2D and D2H
        # lie inside the timed region, one event pair around the n steps, results read back step by step.
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        prev = None
        for _ in range(n):
            cur = stepper.submit(h_imgs[0], h_imgs[1], h_imgs[2], h_ff, h_fb)
            if prev is not None:
                stepper.result(prev)
            prev = cur
        stepper.result(prev)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    e2e_run(3)
    if world > 1:
        dist.barrier()
    e2e_total = e2e_run(n_e2e)
    e2e_ms = [e2e_total / n_e2e] * n_e2e

    # extra (not the contract value): the same step with the three frames shipped as uint8 and `img / 255.0` of the reference's
    # dataset (kitti_prepared.py:89) done on the device -- the frames are bytes at the source; 4x less PCIe traffic for them
    h_imgs_f32, stepper_f32 = h_imgs, stepper
    h_imgs = [pin((x * 255.0).round().clamp(0, 255).to(torch.uint8)) for x in (host.img_l, host.img, host.img_r)]
    stepper = FlowLossStep(B, H, W, LEVELS, device=dev, frame_dtype=torch.uint8)
    e2e_run(3)
    if world > 1:
        dist.barrier()
    u8_total = e2e_run(n_e2e)
    u8_h2d = stepper.h2d_bytes
    h_imgs, stepper = h_imgs_f32, stepper_f32

    # ---- max over ranks ------------------------------------------------------------------------------------
    red = torch.tensor([total_ms, e2e_total / len(e2e_ms), u8_total / n_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms_per_step, u8_ms_per_step = float(red[0]), float(red[1]), float(red[2])
    ms_per_step = total_ms / K
    value = 2.0 * B * world / (ms_per_step * 1e-3)
    e2e_value = 2.0 * B * world / (e2e_ms_per_step * 1e-3)

    if rank == 0:
        peak, peak_src = _peaks()
        fwd_ms, bwd_ms = statistics.mean(stats_ms["fwd"]), statistics.mean(stats_ms["bwd"])
        # single-pass mode: the forward launch (flow_loss_fwdgrad_kernel + finalize) executes the whole forward+backward
        # algorithm (30 N floats per sample-level); the backward launch is an element-wise combine of its maps.
        dom, dom_ms = "fwdgrad", fwd_ms
        dom_bytes = algorithmic_bytes(B, True, True)
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        step_achieved = algorithmic_bytes(B, True, True) / (ms_per_step * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)["flow_loss_fwdgrad_kernel"]
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "height": H, "width": W, "levels": LEVELS, "batch_per_gpu": B,
                       "global_batch": B * world, "launch": "cuda-graph" if graph is not None else "eager",
                       "l2": "flushed between steps (256 MiB memset outside the event pair)",
                       "timing": "sum of per-step CUDA-event pairs on the launching stream, max over ranks"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms_per_step,
                    "h2d_bytes_per_step": stepper.h2d_bytes, "d2h_bytes_per_step": stepper.d2h_bytes,
                    "how": "step.FlowLossStep: pinned host frames+flows -> H2D -> pyramids -> fused fwd+bwd -> D2H losses; "
                           "2 staging slots, copy of step k+1 overlaps compute of step k; one event pair around %d steps" % n_e2e,
                    "uint8_frames": {"value": 2.0 * B * world / (u8_ms_per_step * 1e-3), "unit": UNIT, "ms_per_step": u8_ms_per_step,
                                     "h2d_bytes_per_step": u8_h2d,
                                     "how": "extra, not the contract value: FlowLossStep(frame_dtype=uint8) -- frames shipped as bytes, "
                                            "the dataset's img/255.0 evaluated on the device (bit-identical), flows fp32"}},
            "gpu_launches": launches_per_step * K,
            "roofline": {"bound": "hbm", "kernel": "flow_loss_%s_kernel" % dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full, bytes per launch)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms,
                         "fwdgrad_plus_finalize_ms": fwd_ms, "combine_ms": bwd_ms,
                         "timing": "CUDA events around a graph replay of that half on the launching stream, L2 flushed before each",
                         "step": {"achieved": step_achieved, "frac": step_achieved / peak,
                                  "algorithmic_bytes_per_step": algorithmic_bytes(B, True, True)}},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            rate, sec = cpu_reference_rate(1, 8, 2, cores)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "batch 1 of the batch-%d workload, 8 timed steps (oracle/loss_port.py, torch CPU)" % B}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
