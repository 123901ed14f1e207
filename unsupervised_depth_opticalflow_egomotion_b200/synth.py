"""Seeded synthetic KITTI-shaped inputs for the photometric-loss path.

Shapes and distributions follow SURVEY.md §8(d): blurred-uniform image triplets, a 4-level
forward/backward flow pyramid, S-level disparities for three frames, a small 6-DoF pose pair
and a KITTI-like intrinsic matrix.  Everything is drawn on the CPU from one
``torch.Generator`` (so the same seed gives the same tensors on every box) and then moved to
the requested device.  This module is input generation only: it is used by the tests, by
``bench.py`` and by ``__graft_entry__.smoke()``; it never touches ``oracle/``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import torch
import torch.nn.functional as F


def _blur(t: torch.Tensor, k: int) -> torch.Tensor:
    """k x k box blur with replicate padding (keeps the value range)."""
    if k <= 1:
        return t
    p = k // 2
    tp = F.pad(t, (p, p, p, p), mode="replicate")
    return F.avg_pool2d(tp, k, stride=1)


def kitti_like_intrinsics(batch: int, height: int, width: int) -> torch.Tensor:
    """K = [[0.58 W, 0, 0.5 W], [0, 1.92 H, 0.5 H], [0, 0, 1]]."""
    K = torch.tensor(
        [[0.58 * width, 0.0, 0.5 * width], [0.0, 1.92 * height, 0.5 * height], [0.0, 0.0, 1.0]],
        dtype=torch.float32,
    )
    return K.unsqueeze(0).repeat(batch, 1, 1).contiguous()


def _euler_rot(ang: torch.Tensor) -> torch.Tensor:
    """R = Rx @ Ry @ Rz for angles (B,3); used only to make rigid-consistent synthetic flows."""
    x, y, z = ang[:, 0], ang[:, 1], ang[:, 2]
    o, zero = torch.ones_like(x), torch.zeros_like(x)
    rz = torch.stack([z.cos(), -z.sin(), zero, z.sin(), z.cos(), zero, zero, zero, o], 1).view(-1, 3, 3)
    ry = torch.stack([y.cos(), zero, y.sin(), zero, o, zero, -y.sin(), zero, y.cos()], 1).view(-1, 3, 3)
    rx = torch.stack([o, zero, zero, zero, x.cos(), -x.sin(), zero, x.sin(), x.cos()], 1).view(-1, 3, 3)
    return rx @ ry @ rz


def _synthetic_rigid_flow(depth: torch.Tensor, pose: torch.Tensor, K: torch.Tensor) -> torch.Tensor:
    """Pixel displacement induced by (depth, pose, K); input generation helper only."""
    B, _, h, w = depth.shape
    jj = torch.arange(w, dtype=torch.float32).view(1, 1, w).expand(B, h, w)
    ii = torch.arange(h, dtype=torch.float32).view(1, h, 1).expand(B, h, w)
    pix = torch.stack([jj, ii, torch.ones_like(jj)], 1).reshape(B, 3, -1)
    cam = (torch.linalg.inv(K) @ pix) * depth.reshape(B, 1, -1)
    P = K @ torch.cat([_euler_rot(pose[:, 3:]), pose[:, :3].unsqueeze(-1)], 2)
    q = P[:, :, :3] @ cam + P[:, :, 3:]
    z = q[:, 2].clamp(min=1e-3)
    return torch.stack([q[:, 0] / z - pix[:, 0], q[:, 1] / z - pix[:, 1]], 1).reshape(B, 2, h, w)


@dataclass
class Triplet:
    """One synthetic batch of (left, centre, right) frames plus network-like predictions."""

    img_l: torch.Tensor            # (B,3,H,W)
    img: torch.Tensor              # (B,3,H,W)
    img_r: torch.Tensor            # (B,3,H,W)
    flows_fwd: List[torch.Tensor]  # L x (B,2,H>>s,W>>s)   centre -> right
    flows_bwd: List[torch.Tensor]  # L x (B,2,H>>s,W>>s)   centre -> left
    disp: List[torch.Tensor] = field(default_factory=list)    # S x (B,1,H>>s,W>>s) centre
    disp_l: List[torch.Tensor] = field(default_factory=list)
    disp_r: List[torch.Tensor] = field(default_factory=list)
    pose: torch.Tensor | None = None   # (B,2,6): [:,0] centre->left, [:,1] centre->right
    K: torch.Tensor | None = None      # (B,3,3)
    K_inv: torch.Tensor | None = None  # (B,3,3)

    def to(self, device) -> "Triplet":
        mv = lambda t: None if t is None else t.to(device)
        return Triplet(
            mv(self.img_l), mv(self.img), mv(self.img_r),
            [mv(f) for f in self.flows_fwd], [mv(f) for f in self.flows_bwd],
            [mv(d) for d in self.disp], [mv(d) for d in self.disp_l], [mv(d) for d in self.disp_r],
            mv(self.pose), mv(self.K), mv(self.K_inv),
        )


def make_triplet(
    batch: int,
    height: int,
    width: int,
    flow_levels: int = 4,
    depth_scales: int = 3,
    seed: int = 1234,
    flow_mode: str = "noise",      # "noise": blurred N(0,1)*20/2^s px ; "rigid": rigid flow + N(0,0.5^2)
    flow_px: float = 20.0,
    oob_fraction: float = 0.0,     # >0: push roughly this fraction of pixels out of bounds (adversarial)
    blur: int = 9,
    device="cpu",
) -> Triplet:
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    rnd = lambda *s: torch.rand(*s, generator=g, dtype=torch.float32)
    nrm = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32)

    base = _blur(rnd(batch, 3, height, width), blur)
    # left / right frames: the centre frame plus independent blurred texture, so photometric
    # differences are small but non-zero and SSIM windows are non-degenerate.
    img = base
    img_l = (0.7 * base + 0.3 * _blur(rnd(batch, 3, height, width), blur)).contiguous()
    img_r = (0.7 * base + 0.3 * _blur(rnd(batch, 3, height, width), blur)).contiguous()

    K = kitti_like_intrinsics(batch, height, width)
    K_inv = torch.linalg.inv(K).contiguous()
    pose = 0.01 * nrm(batch, 2, 6)
    pose[:, 0, 2] -= 0.02
    pose[:, 1, 2] += 0.02

    disp, disp_l, disp_r = [], [], []
    for s in range(depth_scales):
        h, w = height >> s, width >> s
        k = max(1, blur >> s) | 1
        disp.append((0.05 + 0.9 * _blur(rnd(batch, 1, h, w), k)).contiguous())
        disp_l.append((0.05 + 0.9 * _blur(rnd(batch, 1, h, w), k)).contiguous())
        disp_r.append((0.05 + 0.9 * _blur(rnd(batch, 1, h, w), k)).contiguous())

    flows_fwd, flows_bwd = [], []
    for s in range(flow_levels):
        h, w = height >> s, width >> s
        k = max(1, blur >> s) | 1
        for out, pidx in ((flows_bwd, 0), (flows_fwd, 1)):
            if flow_mode == "rigid" and s < depth_scales:
                Ks = K.clone()
                Ks[:, 0:2] = Ks[:, 0:2] / float(1 << s)
                f = _synthetic_rigid_flow(disp[s], pose[:, pidx], Ks) + _blur(0.5 * nrm(batch, 2, h, w), k)
            else:
                f = _blur(nrm(batch, 2, h, w), k) * (flow_px * float(k) / float(1 << s))
            if oob_fraction > 0.0:
                push = (rnd(batch, 1, h, w) < oob_fraction).float()
                f = f + push * torch.tensor([2.0 * w, 0.0]).view(1, 2, 1, 1)
            out.append(f.contiguous())

    return Triplet(img_l, img, img_r, flows_fwd, flows_bwd, disp, disp_l, disp_r, pose, K, K_inv).to(device)
