#!/usr/bin/env python
"""ORACLE-SIDE (test infrastructure) — pin ``oracle/loss_port.py`` against the unmodified
reference executed in the build container.

    python oracle/validate_against_reference.py [--height 64 --width 208 --batch 2]

For the primitives and for each of the three modes it compares loss values, every mask
(bit-exact) and the autograd gradients w.r.t. flows / disparities / pose.  Exit code 0 = pinned.
The same comparisons run under pytest (tests/test_oracle_vs_reference.py) whenever
/root/reference is mounted.
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import loss_port as P                      # noqa: E402
from oracle import reference_harness as R              # noqa: E402
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet  # noqa: E402


def _leaves(t, names):
    out = []
    for n in names:
        v = getattr(t, n)
        vs = v if isinstance(v, list) else [v]
        for x in vs:
            x.requires_grad_(True)
            out.append(x)
    return out


def _rel(a, b):
    d = (a - b).abs().max().item()
    s = max(b.abs().max().item(), 1e-30)
    return d / s


def compare_mode(name, run_ref, run_port, leaf_names, weights, mk):
    worst = {"loss": 0.0, "grad": 0.0, "mask_flips": 0}
    t1, t2 = mk(), mk()
    l1 = _leaves(t1, leaf_names)
    l2 = _leaves(t2, leaf_names)
    loss_r, aux_r = run_ref(t1)
    loss_p, aux_p = run_port(t2)
    for k in loss_r:
        worst["loss"] = max(worst["loss"], _rel(loss_p[k].detach(), loss_r[k].detach()))
    tot_r = sum(weights[k] * loss_r[k].mean() for k in loss_r)
    tot_p = sum(weights[k] * loss_p[k].mean() for k in loss_r)
    g_r = torch.autograd.grad(tot_r, l1, allow_unused=True)
    g_p = torch.autograd.grad(tot_p, l2, allow_unused=True)
    for a, b in zip(g_p, g_r):
        if b is None:
            assert a is None or a.abs().max() == 0
            continue
        worst["grad"] = max(worst["grad"], _rel(a, b))
    for k, v in aux_r.items():
        if v is None or k not in aux_p:
            continue
        vs_r = v if isinstance(v, list) else [v]
        vs_p = aux_p[k] if isinstance(aux_p[k], list) else [aux_p[k]]
        for a, b in zip(vs_p, vs_r):
            if set(torch.unique(b.detach()).tolist()) <= {0.0, 1.0}:
                worst["mask_flips"] += int((a.detach() != b.detach()).sum())
    print("%-14s loss rel %.2e   grad rel %.2e   mask flips %d" % (name, worst["loss"], worst["grad"], worst["mask_flips"]))
    return worst


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=64)
    ap.add_argument("--width", type=int, default=208)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--seed", type=int, default=7)
    a = ap.parse_args(argv)
    if not R.available():
        print("reference tree not mounted; nothing to validate")
        return 0
    torch.manual_seed(0)
    ok = True
    for flow_mode, oob in (("noise", 0.0), ("rigid", 0.0), ("noise", 0.3)):
        mk = lambda: make_triplet(a.batch, a.height, a.width, 4, 3, seed=a.seed, flow_mode=flow_mode,
                                  flow_px=6.0, oob_fraction=oob)
        print("== inputs: flow_mode=%s oob=%.1f" % (flow_mode, oob))
        res = [
            compare_mode("flow S=4", lambda t: R.reference_flow_mode(t, 4),
                         lambda t: P.flow_mode_loss(t.img_l, t.img, t.img_r, t.flows_fwd, t.flows_bwd, 4, return_aux=True),
                         ["flows_fwd", "flows_bwd"], P.FLOW_WEIGHTS, mk),
            compare_mode("depth live", lambda t: R.reference_depth_mode(t, 3, False),
                         lambda t: _drop(P.depth_mode_loss(t.img_l, t.img, t.img_r, t.disp, t.disp_l, t.disp_r, t.pose, t.K, 3,
                                                           "live", return_aux=True)),
                         ["disp", "disp_l", "disp_r", "pose"], P.GEOM_WEIGHTS, mk),
            compare_mode("depth texture", lambda t: R.reference_depth_mode(t, 3, True),
                         lambda t: P.depth_mode_loss(t.img_l, t.img, t.img_r, t.disp, t.disp_l, t.disp_r, t.pose, t.K, 3,
                                                     "texture", return_aux=True),
                         ["disp", "disp_l", "disp_r", "pose"], P.GEOM_WEIGHTS, mk),
            compare_mode("geom S=3", lambda t: R.reference_geom_mode(t, 3),
                         lambda t: P.geom_mode_loss(t.img_l, t.img, t.img_r, t.flows_fwd, t.flows_bwd, t.disp, t.disp_l,
                                                    t.disp_r, t.pose, t.K, t.K_inv, 3, return_aux=True),
                         ["flows_fwd", "flows_bwd", "disp", "disp_l", "disp_r", "pose"], P.GEOM_WEIGHTS, mk),
        ]
        for r in res:
            ok &= r["loss"] < 1e-6 and r["grad"] < 1e-5 and r["mask_flips"] == 0
    # PWC-Net cost volume: PWC_tf.corr_naive called unbound
    ref = R.load()
    g = torch.Generator().manual_seed(a.seed)
    f1 = torch.randn(a.batch, 24, a.height // 4, a.width // 4, generator=g).requires_grad_(True)
    f2 = torch.randn(a.batch, 24, a.height // 4, a.width // 4, generator=g).requires_grad_(True)
    go = torch.randn(a.batch, 81, a.height // 4, a.width // 4, generator=g)
    o_r, o_p = ref.structures.PWC_tf.corr_naive(None, f1, f2), P.cost_volume(f1, f2)
    g_r, g_p = torch.autograd.grad((o_r * go).sum(), [f1, f2]), torch.autograd.grad((o_p * go).sum(), [f1, f2])
    cv = max(_rel(o_p.detach(), o_r.detach()), _rel(g_p[0], g_r[0]), _rel(g_p[1], g_r[1]))
    print("%-14s worst rel err %.2e" % ("cost volume", cv))
    ok &= cv < 1e-6
    print("PINNED" if ok else "MISMATCH")
    return 0 if ok else 1


def _drop(pair):
    loss, aux = pair
    return {k: v for k, v in loss.items() if k not in ('loss_depth_ssim', 'loss_depth_consis')}, aux


if __name__ == "__main__":
    sys.exit(main())
