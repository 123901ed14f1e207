#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by executing the UNMODIFIED reference
(/root/reference, imported in place through oracle/reference_harness.py) on seeded synthetic inputs.

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, these small .npz files can.  Each fixture stores the
inputs, the reference's outputs (losses, masks) and its autograd gradients.  The oracle
(oracle/loss_port.py) and the CUDA path are both tested against them.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import reference_harness as R                                      # noqa: E402
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet   # noqa: E402

FLOW_W = {"loss_flow_pixel": 0.15, "loss_flow_ssim": 0.85, "loss_flow_smooth": 10.0, "loss_flow_consis": 0.01}
GEOM_W = dict(FLOW_W, loss_depth_pixel=1.0, loss_depth_ssim=0.85, loss_depth_smooth=0.5, loss_depth_consis=0.1,
              loss_depth_flow_consis=1.0, loss_epipolar=0.1)


def _np(t):
    return t.detach().cpu().numpy()


def _store_inputs(d, t):
    d["img_l"], d["img"], d["img_r"] = _np(t.img_l), _np(t.img), _np(t.img_r)
    for l, f in enumerate(t.flows_fwd):
        d["flow_fwd_%d" % l] = _np(f)
    for l, f in enumerate(t.flows_bwd):
        d["flow_bwd_%d" % l] = _np(f)
    for name in ("disp", "disp_l", "disp_r"):
        for l, f in enumerate(getattr(t, name)):
            d["%s_%d" % (name, l)] = _np(f)
    d["pose"], d["K"], d["K_inv"] = _np(t.pose), _np(t.K), _np(t.K_inv)


def _leaves(t, names):
    out = []
    for n in names:
        v = getattr(t, n)
        for x in (v if isinstance(v, list) else [v]):
            out.append((n, x.requires_grad_(True)))
    return out


def _run(kind, t, weights, leaf_names, fn):
    leaves = _leaves(t, leaf_names)
    loss, aux = fn(t)
    d = {}
    _store_inputs(d, t)
    for k, v in loss.items():
        d["out_" + k] = _np(v)
    total = sum(weights[k] * v.mean() for k, v in loss.items())
    grads = torch.autograd.grad(total, [x for _, x in leaves], allow_unused=True)
    counter = {}
    for (n, x), g in zip(leaves, grads):
        i = counter.get(n, 0)
        counter[n] = i + 1
        key = "grad_%s_%d" % (n, i) if isinstance(getattr(t, n), list) else "grad_" + n
        d[key] = _np(g) if g is not None else np.zeros(tuple(x.shape), np.float32)
    for k, v in aux.items():
        if v is None:
            continue
        vs = v if isinstance(v, list) else [v]
        for l, m in enumerate(vs):
            m = m.detach()
            if m.shape[1] == 1:   # masks / weights / distance maps
                d["aux_%s_%d" % (k, l)] = _np(m)
    path = os.path.join(HERE, kind + ".npz")
    np.savez_compressed(path, **d)
    print("wrote %-28s %6.1f KB" % (os.path.basename(path), os.path.getsize(path) / 1024))


def primitives():
    ref = R.load()
    g = torch.Generator().manual_seed(5)
    d = {}
    # warp_flow with and without mask (structures/net_utils.py:16-54), C=5 to exercise C != 3
    x = torch.rand(2, 5, 12, 20, generator=g).requires_grad_(True)
    flow = (4.0 * torch.randn(2, 2, 12, 20, generator=g)).requires_grad_(True)
    go = torch.randn(2, 5, 12, 20, generator=g)
    for use_mask in (False, True):
        out = ref.structures.warp_flow(x, flow, use_mask=use_mask)
        gx, gf = torch.autograd.grad((out * go).sum(), [x, flow])
        tag = "warp_mask%d_" % int(use_mask)
        d[tag + "out"], d[tag + "grad_x"], d[tag + "grad_flow"] = _np(out), _np(gx), _np(gf)
    d["warp_x"], d["warp_flow"], d["warp_go"] = _np(x), _np(flow), _np(go)
    # SSIM map (pytorch_ssim/ssim.py:4-19)
    a = torch.rand(2, 3, 10, 14, generator=g).requires_grad_(True)
    b = (a.detach() + 0.1 * torch.randn(2, 3, 10, 14, generator=g)).requires_grad_(True)
    gs = torch.randn(2, 3, 10, 14, generator=g)
    s = ref.pytorch_ssim.SSIM(a, b)
    ga, gb = torch.autograd.grad((s * gs).sum(), [a, b])
    d.update(ssim_x=_np(a), ssim_y=_np(b), ssim_go=_np(gs), ssim_out=_np(s), ssim_grad_x=_np(ga), ssim_grad_y=_np(gb))
    # pyramids (model_flow.py:58-64 adaptive pooling; model_geometry.py:65-72 bilinear)
    img = torch.rand(1, 3, 16, 32, generator=g)
    mf, mg = R.flow_model(3), R.geom_model(3)
    for l, p in enumerate(mf.generate_img_pyramid(img, 3)):
        d["pyr_box_%d" % l] = _np(p)
    for l, p in enumerate(mg.generate_img_pyramid(img, 3)):
        d["pyr_bilinear_%d" % l] = _np(p)
    d["pyr_img"] = _np(img)
    # inverse_warp2 / calculate_rigid_flow (structures/inverse_warp.py:263-303, 311-342)
    t = make_triplet(2, 16, 24, 1, 1, seed=9)
    depth = t.disp[0].clone().requires_grad_(True)
    refd = t.disp_l[0].clone().requires_grad_(True)
    pose = (5.0 * t.pose[:, 0]).clone().requires_grad_(True)
    src = t.img_l.clone().requires_grad_(True)
    rec, valid, proj, comp = ref.structures.inverse_warp2(src, depth, refd, pose, t.K)
    g1, g2, g3 = torch.randn(rec.shape, generator=g), torch.randn(proj.shape, generator=g), torch.randn(comp.shape, generator=g)
    gd, gr, gp, gi = torch.autograd.grad((rec * g1).sum() + (proj * g2).sum() + (comp * g3).sum(), [depth, refd, pose, src])
    rf = ref.structures.calculate_rigid_flow(depth, pose, t.K)
    g4 = torch.randn(rf.shape, generator=g)
    gd2, gp2 = torch.autograd.grad((rf * g4).sum(), [depth, pose])
    d.update(iw_img=_np(src), iw_depth=_np(depth), iw_ref_depth=_np(refd), iw_pose=_np(pose), iw_K=_np(t.K),
             iw_rec=_np(rec), iw_valid=_np(valid), iw_proj=_np(proj), iw_comp=_np(comp),
             iw_go_rec=_np(g1), iw_go_proj=_np(g2), iw_go_comp=_np(g3),
             iw_grad_depth=_np(gd), iw_grad_ref_depth=_np(gr), iw_grad_pose=_np(gp), iw_grad_img=_np(gi),
             rf_out=_np(rf), rf_go=_np(g4), rf_grad_depth=_np(gd2), rf_grad_pose=_np(gp2))
    path = os.path.join(HERE, "primitives.npz")
    np.savez_compressed(path, **d)
    print("wrote %-28s %6.1f KB" % (os.path.basename(path), os.path.getsize(path) / 1024))


def cost_volume():
    """PWC_tf.corr_naive (structures/pwc_tf.py:97-106) called unbound: output and both input gradients"""
    ref = R.load()
    g = torch.Generator().manual_seed(9)
    d = {}
    for tag, (B, Cc, Hh, Ww) in (("a", (2, 6, 9, 14)), ("b", (1, 33, 5, 40))):
        f1 = torch.randn(B, Cc, Hh, Ww, generator=g).requires_grad_(True)
        f2 = torch.randn(B, Cc, Hh, Ww, generator=g).requires_grad_(True)
        go = torch.randn(B, 81, Hh, Ww, generator=g)
        out = ref.structures.PWC_tf.corr_naive(None, f1, f2)
        g1, g2 = torch.autograd.grad((out * go).sum(), [f1, f2])
        d.update({tag + "_f1": _np(f1), tag + "_f2": _np(f2), tag + "_go": _np(go), tag + "_out": _np(out), tag + "_g1": _np(g1),
                  tag + "_g2": _np(g2)})
    path = os.path.join(HERE, "cost_volume.npz")
    np.savez_compressed(path, **d)
    print("wrote %-28s %6.1f KB" % (os.path.basename(path), os.path.getsize(path) / 1024))


def main():
    if "--only-cost-volume" in sys.argv:
        torch.manual_seed(0)
        torch.set_num_threads(1)
        cost_volume()
        return 0
    if not R.available():
        print("reference tree not mounted; cannot (re)generate golden fixtures")
        return 1
    torch.manual_seed(0)
    torch.set_num_threads(1)   # fixed reduction order inside ATen
    mk = lambda **kw: make_triplet(2, 32, 64, 4, 3, **kw)
    _run("flow_mode_s4", mk(seed=101, flow_px=3.0), FLOW_W, ["flows_fwd", "flows_bwd"], lambda t: R.reference_flow_mode(t, 4))
    _run("flow_mode_s4_oob", mk(seed=102, flow_px=3.0, oob_fraction=0.3), FLOW_W, ["flows_fwd", "flows_bwd"],
         lambda t: R.reference_flow_mode(t, 4))
    _run("flow_mode_s3", mk(seed=103, flow_px=2.0), FLOW_W, ["flows_fwd", "flows_bwd"], lambda t: R.reference_flow_mode(t, 3))
    dl = ["disp", "disp_l", "disp_r", "pose"]
    _run("depth_mode_live", mk(seed=104, flow_mode="rigid"), GEOM_W, dl, lambda t: R.reference_depth_mode(t, 3, False))
    _run("depth_mode_texture", mk(seed=105, flow_mode="rigid"), GEOM_W, dl, lambda t: R.reference_depth_mode(t, 3, True))
    _run("geom_mode_s3", mk(seed=106, flow_mode="rigid"), GEOM_W, ["flows_fwd", "flows_bwd"] + dl,
         lambda t: R.reference_geom_mode(t, 3))
    # multi-tile fixtures (round 2): 112 x 168, batch 1 -- several 32x13 / 32x32 tiles in both directions with ragged
    # remainders, level widths 168 / 84 / 42 / 21 (odd: no vector stores, no TMA staging at the two small levels)
    mt = lambda **kw: make_triplet(1, 112, 168, 4, 3, **kw)
    _run("flow_mode_s4_mt", mt(seed=111, flow_px=4.0, oob_fraction=0.05), FLOW_W, ["flows_fwd", "flows_bwd"], lambda t: R.reference_flow_mode(t, 4))
    _run("depth_mode_live_mt", mt(seed=114, flow_mode="rigid"), GEOM_W, dl, lambda t: R.reference_depth_mode(t, 3, False))
    _run("depth_mode_texture_mt", mt(seed=115, flow_mode="rigid"), GEOM_W, dl, lambda t: R.reference_depth_mode(t, 3, True))
    _run("geom_mode_s3_mt", mt(seed=116, flow_mode="rigid"), GEOM_W, ["flows_fwd", "flows_bwd"] + dl,
         lambda t: R.reference_geom_mode(t, 3))
    primitives()
    cost_volume()
    return 0


if __name__ == "__main__":
    sys.exit(main())
