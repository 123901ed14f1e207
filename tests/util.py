"""Shared test helpers: golden-fixture loading and comparison metrics."""
from __future__ import annotations

import os

import numpy as np
import torch

from unsupervised_depth_opticalflow_egomotion_b200.synth import Triplet

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

LOSS_RTOL = 1e-5      # north_star: 1e-5 relative on losses
GRAD_RTOL = 1e-4      # north_star: 1e-4 relative on gradients (relative to the largest |g| of the tensor)


def load_golden(name: str):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _list(d, prefix):
    out, l = [], 0
    while "%s_%d" % (prefix, l) in d:
        out.append(d["%s_%d" % (prefix, l)].clone())
        l += 1
    return out


def golden_triplet(d, device="cpu") -> Triplet:
    t = Triplet(d["img_l"].clone(), d["img"].clone(), d["img_r"].clone(), _list(d, "flow_fwd"), _list(d, "flow_bwd"),
                _list(d, "disp"), _list(d, "disp_l"), _list(d, "disp_r"), d["pose"].clone(), d["K"].clone(), d["K_inv"].clone())
    return t.to(device)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| relative to max |b| (per-tensor scale)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def loss_rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """element-wise relative error of per-sample losses."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float(((a - b).abs() / b.abs().clamp_min(1e-30)).max())
