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


class FlowLossStep:
    """Flow-mode loss step (Model_flow.forward loss body, model_flow.py:232-254) from host buffers."""

    def __init__(self, batch: int, height: int, width: int, levels: int = 4, num_scales: Optional[int] = None,
                 weights: Optional[Dict[str, float]] = None, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FlowLossStep needs a CUDA device: the loss path has no CPU implementation")
        self.B, self.H, self.W, self.L = batch, height, width, levels
        self.scales = levels if num_scales is None else num_scales
        d, f32 = self.device, torch.float32
        self.d_imgs = [torch.empty((batch, 3, height, width), device=d, dtype=f32) for _ in range(3)]
        self.d_ff = [torch.empty((batch, 2, height >> l, width >> l), device=d, dtype=f32) for l in range(levels)]
        self.d_fb = [torch.empty((batch, 2, height >> l, width >> l), device=d, dtype=f32) for l in range(levels)]
        self.h_loss = torch.empty((4, batch), dtype=f32).pin_memory()
        self.wmat = weight_matrix(weights or FLOW_WEIGHTS, ops.FLOW_LOSS_KEYS, batch, d)
        self.h2d_bytes = sum(t.numel() * 4 for t in self.d_imgs + self.d_ff + self.d_fb)
        self.d2h_bytes = self.h_loss.numel() * 4
        self.grads: List[torch.Tensor] = []

    def __call__(self, h_img_l: torch.Tensor, h_img: torch.Tensor, h_img_r: torch.Tensor, h_flows_fwd: Sequence[torch.Tensor],
                 h_flows_bwd: Sequence[torch.Tensor], sync: bool = True) -> torch.Tensor:
        """Host tensors in (pinned for async copies), per-sample losses (4,B) back on the host.
        Flow gradients of d(total)/d(flow) stay on the device in ``self.grads`` (fwd levels, then bwd)."""
        for dst, src in zip(self.d_imgs, (h_img_l, h_img, h_img_r)):
            dst.copy_(src, non_blocking=True)
        for dst, src in zip(self.d_ff + self.d_fb, list(h_flows_fwd) + list(h_flows_bwd)):
            dst.copy_(src, non_blocking=True)
        pl, pc, pr = (ops.image_pyramid(x, self.L, "box") for x in self.d_imgs)
        ff = [f.detach().requires_grad_(True) for f in self.d_ff]   # fresh leaves over the staging buffers
        fb = [f.detach().requires_grad_(True) for f in self.d_fb]
        loss = ops.flow_loss(pl, pc, pr, ff, fb, self.scales, as_matrix=True)
        self.grads = list(torch.autograd.grad(loss, ff[:self.scales] + fb[:self.scales], grad_outputs=self.wmat))
        self.h_loss.copy_(loss.detach(), non_blocking=True)
        if sync:
            torch.cuda.current_stream().synchronize()
        return self.h_loss
