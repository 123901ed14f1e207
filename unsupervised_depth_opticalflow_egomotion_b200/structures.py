"""Drop-in replacements for the free functions of ``core/networks/structures`` and
``core/networks/pytorch_ssim`` that lie on the loss path: same names, positional signatures, return
arity / shape / dtype and error behaviour, backed by the sm_100a kernels.

The per-sample 3x3 / 3x4 matrix algebra (K^-1, R(euler), K [R|t], K^-T [t]x R K^-1) stays in
PyTorch exactly as in the reference (12-element tensors: plumbing, and it keeps the pose gradient in
the reference's own arithmetic); everything per-pixel runs in the CUDA kernels.
"""
from __future__ import annotations

import torch

from . import ops

Tensor = torch.Tensor

warp_flow = ops.warp_flow          # structures/net_utils.py:16-54
SSIM = ops.ssim                    # pytorch_ssim/ssim.py:4-19


def check_sizes(input: Tensor, input_name: str, expected: str) -> None:
    """structures/inverse_warp.py:21-27 — same AssertionError text."""
    ok = [input.ndimension() == len(expected)]
    for i, size in enumerate(expected):
        if size.isdigit():
            ok.append(input.size(i) == int(size))
    assert all(ok), "wrong size for {}, expected {}, got  {}".format(input_name, "x".join(expected), list(input.size()))


def euler2mat(angle: Tensor) -> Tensor:
    """structures/inverse_warp.py:110-145: R = Rx @ Ry @ Rz for (rx, ry, rz) in radians, (B,3) -> (B,3,3)."""
    rx, ry, rz = angle[:, 0], angle[:, 1], angle[:, 2]
    zero = rz.detach() * 0
    one = zero + 1

    def mat(*rows):
        return torch.stack(rows, dim=1).reshape(-1, 3, 3)

    cz, sz = torch.cos(rz), torch.sin(rz)
    cy, sy = torch.cos(ry), torch.sin(ry)
    cx, sx = torch.cos(rx), torch.sin(rx)
    Rz = mat(cz, -sz, zero, sz, cz, zero, zero, zero, one)
    Ry = mat(cy, zero, sy, zero, one, zero, -sy, zero, cy)
    Rx = mat(one, zero, zero, zero, cx, -sx, zero, sx, cx)
    return Rx @ Ry @ Rz


def pose_vec2mat(vec: Tensor, rotation_mode: str = "euler") -> Tensor:
    """structures/inverse_warp.py:172-187: [tx,ty,tz,rx,ry,rz] -> [R|t] (B,3,4)."""
    if rotation_mode != "euler":
        raise NotImplementedError("only the euler parameterisation is on the live loss path")
    return torch.cat([euler2mat(vec[:, 3:]), vec[:, :3].unsqueeze(-1)], dim=2)


def _projection(pose: Tensor, intrinsics: Tensor):
    return intrinsics.inverse().contiguous(), (intrinsics @ pose_vec2mat(pose)).contiguous()


def inverse_warp2(img: Tensor, depth: Tensor, ref_depth: Tensor, pose: Tensor, intrinsics: Tensor, padding_mode: str = "zeros"):
    """structures/inverse_warp.py:263-303 -> (projected_img, valid_mask, projected_depth, computed_depth)."""
    check_sizes(img, "img", "B3HW")
    check_sizes(depth, "depth", "B1HW")
    check_sizes(ref_depth, "ref_depth", "B1HW")
    check_sizes(pose, "pose", "B6")
    check_sizes(intrinsics, "intrinsics", "B33")
    if padding_mode != "zeros":
        raise NotImplementedError("the loss path only uses padding_mode='zeros'")
    Kinv, P = _projection(pose, intrinsics)
    return ops.reproject(img, depth, ref_depth, Kinv, P)


def calculate_rigid_flow(depth: Tensor, pose: Tensor, intrinsics: Tensor) -> Tensor:
    """structures/inverse_warp.py:311-342 -> (B,2,H,W) un-normalised pixel displacement."""
    Kinv, P = _projection(pose, intrinsics)
    return ops.rigid_flow(depth, Kinv, P)


def skewsymmetric(t: Tensor) -> Tensor:
    zero = torch.zeros_like(t[:, 0])
    return torch.stack([zero, -t[:, 2], t[:, 1], t[:, 2], zero, -t[:, 0], -t[:, 1], t[:, 0], zero], dim=1).view(-1, 3, 3)


def compute_essential_matrix(vec: Tensor, rotation_mode: str = "euler") -> Tensor:
    """structures/inverse_warp.py:354-364: E = [t]x @ R."""
    if rotation_mode != "euler":
        raise NotImplementedError("only the euler parameterisation is on the live loss path")
    return skewsymmetric(vec[:, :3]).bmm(euler2mat(vec[:, 3:]))
