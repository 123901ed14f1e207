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


def inv3x3(M: Tensor) -> Tensor:
    """Closed-form inverse of (...,3,3) matrices (adjugate / determinant) in element-wise ops.  Replaces the
    reference's ``intrinsics.inverse()`` (inverse_warp.py:284): ``torch.inverse`` runs a batched LU with a host
    synchronisation and cannot be captured into a CUDA graph; the closed form agrees with it to ~1e-7 relative."""
    a, b, c = M[..., 0, 0], M[..., 0, 1], M[..., 0, 2]
    d, e, f = M[..., 1, 0], M[..., 1, 1], M[..., 1, 2]
    g, h, i = M[..., 2, 0], M[..., 2, 1], M[..., 2, 2]
    A, B, Cc = e * i - f * h, f * g - d * i, d * h - e * g
    det = a * A + b * B + c * Cc
    adj = torch.stack([A, c * h - b * i, b * f - c * e,
                       B, a * i - c * g, c * d - a * f,
                       Cc, b * g - a * h, a * e - b * d], dim=-1)
    return (adj / det.unsqueeze(-1)).reshape(M.shape)


def scaled_intrinsics(intrinsics: Tensor, downscale: float) -> Tensor:
    """model_geometry.py:92-93: rows 0-1 of K divided by the downscale factor, row 2 kept."""
    return torch.cat((intrinsics[:, 0:2] / downscale, intrinsics[:, 2:]), dim=1)


def _projection(pose: Tensor, intrinsics: Tensor):
    return inv3x3(intrinsics).contiguous(), (intrinsics @ pose_vec2mat(pose)).contiguous()


def projection_pyramid(intrinsics: Tensor, poses, downscales):
    """K_s^-1 and K_s [R|t] for every level s and every pose in one batched pass (the per-level glue of
    ``reconstruction`` / ``compute_dynamic_mask`` computed once per step).  Returns
    ``Kinv[s]`` (B,3,3) and ``P[k][s]`` (B,3,4) for pose k."""
    Ks = torch.stack([scaled_intrinsics(intrinsics, ds) for ds in downscales], dim=1)          # (B,S,3,3)
    Kinv = inv3x3(Ks)
    Ps = [Ks @ pose_vec2mat(p).unsqueeze(1) for p in poses]                                      # (B,S,3,4) each
    S = len(downscales)
    return [Kinv[:, s].contiguous() for s in range(S)], [[P[:, s].contiguous() for s in range(S)] for P in Ps]


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


def corr_naive(input1: Tensor, input2: Tensor, d: int = 4) -> Tensor:
    """``PWC_tf.corr_naive`` (structures/pwc_tf.py:97-106) as a free function: bind with ``PWC_tf.corr = staticmethod(corr_naive)``
    or ``self.corr = corr_naive`` (:19)."""
    return ops.cost_volume(input1, input2, d)


def skewsymmetric(t: Tensor) -> Tensor:
    zero = torch.zeros_like(t[:, 0])
    return torch.stack([zero, -t[:, 2], t[:, 1], t[:, 2], zero, -t[:, 0], -t[:, 1], t[:, 0], zero], dim=1).view(-1, 3, 3)


def compute_essential_matrix(vec: Tensor, rotation_mode: str = "euler") -> Tensor:
    """structures/inverse_warp.py:354-364: E = [t]x @ R."""
    if rotation_mode != "euler":
        raise NotImplementedError("only the euler parameterisation is on the live loss path")
    return skewsymmetric(vec[:, :3]).bmm(euler2mat(vec[:, 3:]))
