"""Host-buffer entry points: one training-step worth of the loss path, starting from HOST memory.

``FlowLossStep`` is what a data-parallel trainer (``train.py:171-215`` in the reference: H2D copy of the
batch, loss forward, ``sum_k w_k * mean(loss_k)``, ``backward()``) calls when the frames and the
network outputs live in pinned host buffers: it stages them to the device, builds the image pyramids,
runs the fused forward + backward kernels and brings the per-sample losses back.  It owns all its
buffers (allocated once) and never synchronises except for the final D2H read.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops

FLOW_WEIGHTS = {"loss_flow_pixel": 0.15, "loss_flow_ssim": 0.85, "loss_flow_smooth": 10.0, "loss_flow_consis": 0.01}
"""config/kitti.yaml:18-21"""


def weight_matrix(weights: Dict[str, float], keys: Sequence[str], batch: int, device) -> torch.Tensor:
    """d(sum_k w_k * mean_B(loss_k)) / d loss  as a (K,B) matrix (train.py:211-214)."""
    w = torch.tensor([weights[k] for k in keys], dtype=torch.float32).view(-1, 1).repeat(1, batch) / float(batch)
    return w.to(device).contiguous()


class _Slot:
    """One set of device staging buffers + the events that order its reuse."""

    def __init__(self, batch, height, width, levels, device, frame_dtype=torch.float32):
        f32 = torch.float32
        self.imgs = [torch.empty((batch, 3, height, width), device=device, dtype=f32) for _ in range(3)]
        # uint8 frames: staged as bytes, converted on the device (ops.frames_from_u8) into self.imgs
        self.imgs_u8 = ([torch.empty((batch, 3, height, width), device=device, dtype=torch.uint8) for _ in range(3)]
                        if frame_dtype == torch.uint8 else None)
        self.ff = [torch.empty((batch, 2, height >> l, width >> l), device=device, dtype=f32) for l in range(levels)]
        self.fb = [torch.empty((batch, 2, height >> l, width >> l), device=device, dtype=f32) for l in range(levels)]
        self.h_loss = torch.empty((4, batch), dtype=f32).pin_memory()
        self.h2d_done = torch.cuda.Event()
        self.compute_done = torch.cuda.Event()
        self.grads: List[torch.Tensor] = []
        self.busy = False


class FlowLossStep:
    """Flow-mode loss step (Model_flow.forward loss body, model_flow.py:232-254) from host buffers.

    Two staging slots and a copy stream: ``submit`` enqueues the H2D copy of a step on the copy stream and its
    pyramids + fused forward/backward + D2H of the losses on the compute stream, so the copy of step k+1 overlaps
    the kernels of step k.  ``result(slot)`` waits for that step only.  ``__call__`` = submit + result.

    ``frame_dtype=torch.uint8``: the three frames arrive as bytes (what the image decoder produces) and the dataset's
    ``img / 255.0`` (core/dataset/kitti_prepared.py:89) runs on the device, bit-identical to the host division: the frames then
    cost a quarter of the PCIe traffic (the step from host buffers is PCIe-bound)."""

    def __init__(self, batch: int, height: int, width: int, levels: int = 4, num_scales: Optional[int] = None,
                 weights: Optional[Dict[str, float]] = None, device="cuda:0", frame_dtype=torch.float32):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FlowLossStep needs a CUDA device: the loss path has no CPU implementation")
        self.B, self.H, self.W, self.L = batch, height, width, levels
        self.scales = levels if num_scales is None else num_scales
        with torch.cuda.device(self.device):
            if frame_dtype not in (torch.float32, torch.uint8):
                raise TypeError("frame_dtype must be torch.float32 or torch.uint8")
            self.frame_dtype = frame_dtype
            self.slots = [_Slot(batch, height, width, levels, self.device, frame_dtype) for _ in range(2)]
            self.copy_stream = torch.cuda.Stream()
        self.wmat = weight_matrix(weights or FLOW_WEIGHTS, ops.FLOW_LOSS_KEYS, batch, self.device)
        s0 = self.slots[0]
        frames = s0.imgs_u8 if s0.imgs_u8 is not None else s0.imgs
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in frames + s0.ff + s0.fb)
        self.d2h_bytes = s0.h_loss.numel() * 4
        self._next = 0

    def submit(self, h_img_l: torch.Tensor, h_img: torch.Tensor, h_img_r: torch.Tensor, h_flows_fwd: Sequence[torch.Tensor],
               h_flows_bwd: Sequence[torch.Tensor]) -> int:
        """Enqueue one step (host tensors, pinned for truly asynchronous copies); returns the slot to pass to ``result``."""
        idx = self._next
        self._next ^= 1
        sl = self.slots[idx]
        main = torch.cuda.current_stream(self.device)
        if sl.busy:
            self.copy_stream.wait_event(sl.compute_done)          # the slot's previous step must have consumed its buffers
        with torch.cuda.stream(self.copy_stream):
            for dst, src in zip(sl.imgs_u8 if sl.imgs_u8 is not None else sl.imgs, (h_img_l, h_img, h_img_r)):
                if src.dtype != dst.dtype:
                    raise TypeError("FlowLossStep(frame_dtype=%s) got %s frames" % (dst.dtype, src.dtype))
                dst.copy_(src, non_blocking=True)
            for dst, src in zip(sl.ff + sl.fb, list(h_flows_fwd) + list(h_flows_bwd)):
                dst.copy_(src, non_blocking=True)
            sl.h2d_done.record(self.copy_stream)
        main.wait_event(sl.h2d_done)
        if sl.imgs_u8 is not None:
            ops.frames_from_u8(sl.imgs_u8, out=sl.imgs)            # one launch for the three frames
        pl, pc, pr = (d["box"] for d in ops.image_pyramids(sl.imgs, self.L, ("box", "box", "box")))     # one launch
        ff = [f.detach().requires_grad_(True) for f in sl.ff]       # fresh leaves over the staging buffers
        fb = [f.detach().requires_grad_(True) for f in sl.fb]
        loss = ops.flow_loss(pl, pc, pr, ff, fb, self.scales, as_matrix=True)
        sl.grads = list(torch.autograd.grad(loss, ff[:self.scales] + fb[:self.scales], grad_outputs=self.wmat))
        sl.h_loss.copy_(loss.detach(), non_blocking=True)
        sl.compute_done.record(main)
        sl.busy = True
        return idx

    def result(self, slot: int) -> torch.Tensor:
        """Per-sample losses (4,B) of a submitted step on the host; d(total)/d(flow) stays on the device in
        ``self.slots[slot].grads`` (forward levels, then backward levels)."""
        sl = self.slots[slot]
        sl.compute_done.synchronize()
        return sl.h_loss

    @property
    def grads(self) -> List[torch.Tensor]:
        return self.slots[self._next ^ 1].grads

    def __call__(self, h_img_l, h_img, h_img_r, h_flows_fwd, h_flows_bwd, sync: bool = True) -> torch.Tensor:
        slot = self.submit(h_img_l, h_img, h_img_r, h_flows_fwd, h_flows_bwd)
        return self.result(slot) if sync else self.slots[slot].h_loss
