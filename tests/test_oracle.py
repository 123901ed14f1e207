"""CPU tests of the oracle itself: against the golden fixtures generated from the reference, against
the live reference when it is mounted, and known-answer cases."""
import subprocess
import sys
import os

import pytest
import torch

from oracle import loss_port as P
from oracle import reference_harness as R
from util import load_golden, golden_triplet, rel_err, loss_rel_err, LOSS_RTOL, GRAD_RTOL

FLOW_KEYS = ["loss_flow_pixel", "loss_flow_ssim", "loss_flow_smooth", "loss_flow_consis"]


def _check(loss, d, leaves, weights):
    for k, v in loss.items():
        if "out_" + k in d:
            assert loss_rel_err(v, d["out_" + k]) < LOSS_RTOL, k
    total = sum(weights[k] * v.mean() for k, v in loss.items() if "out_" + k in d)
    names = [n for n, _ in leaves]
    grads = torch.autograd.grad(total, [x for _, x in leaves], allow_unused=True)
    for n, g in zip(names, grads):
        ref = d["grad_" + n]
        g = torch.zeros_like(ref) if g is None else g
        if ref.abs().max() == 0:
            assert g.abs().max() == 0, n
        else:
            assert rel_err(g, ref) < GRAD_RTOL, n


def _leaves(t, fields):
    out = []
    for f in fields:
        v = getattr(t, f)
        if isinstance(v, list):
            out += [("%s_%d" % (f, i), x.requires_grad_(True)) for i, x in enumerate(v)]
        else:
            out.append((f, v.requires_grad_(True)))
    return out


@pytest.mark.parametrize("name,scales", [("flow_mode_s4", 4), ("flow_mode_s4_oob", 4), ("flow_mode_s3", 3)])
def test_flow_mode_vs_golden(name, scales):
    d = load_golden(name)
    t = golden_triplet(d)
    leaves = _leaves(t, ["flows_fwd", "flows_bwd"])
    loss, aux = P.flow_mode_loss(t.img_l, t.img, t.img_r, t.flows_fwd, t.flows_bwd, scales, return_aux=True)
    _check(loss, d, leaves, P.FLOW_WEIGHTS)
    for l in range(scales):
        assert rel_err(aux["w_fwd"][l], d["aux_w_fwd_%d" % l]) < 1e-6
        assert rel_err(aux["w_bwd"][l], d["aux_w_bwd_%d" % l]) < 1e-6


@pytest.mark.parametrize("name,variant", [("depth_mode_live", "live"), ("depth_mode_texture", "texture")])
def test_depth_mode_vs_golden(name, variant):
    d = load_golden(name)
    t = golden_triplet(d)
    leaves = _leaves(t, ["disp", "disp_l", "disp_r", "pose"])
    loss, aux = P.depth_mode_loss(t.img_l, t.img, t.img_r, t.disp, t.disp_l, t.disp_r, t.pose, t.K, 3, variant, return_aux=True)
    _check(loss, d, leaves, P.GEOM_WEIGHTS)
    for l in range(3):
        assert torch.equal(aux["valid_l"][l], d["aux_valid_l_%d" % l])
        assert torch.equal(aux["valid_r"][l], d["aux_valid_r_%d" % l])
        if variant == "live":
            assert torch.equal(aux["tex_b"][l], d["aux_tex_b_%d" % l])


def test_geom_mode_vs_golden():
    d = load_golden("geom_mode_s3")
    t = golden_triplet(d)
    leaves = _leaves(t, ["flows_fwd", "flows_bwd", "disp", "disp_l", "disp_r", "pose"])
    loss, aux = P.geom_mode_loss(t.img_l, t.img, t.img_r, t.flows_fwd, t.flows_bwd, t.disp, t.disp_l, t.disp_r, t.pose,
                                 t.K, t.K_inv, 3, return_aux=True)
    _check(loss, d, leaves, P.GEOM_WEIGHTS)
    for key in ("occ_b", "occ_f", "valid_b", "valid_f", "dyn_b", "dyn_f", "tex_b", "tex_f", "val_l", "val_r"):
        for l in range(3):
            assert torch.equal(aux[key][l], d["aux_%s_%d" % (key, l)]), (key, l)   # masks bit-exact


def test_primitives_vs_golden():
    d = load_golden("primitives")
    x, flow, go = d["warp_x"].requires_grad_(True), d["warp_flow"].requires_grad_(True), d["warp_go"]
    for use_mask in (False, True):
        for restated in (False, True):
            out = P.flow_backwarp(x, flow, use_mask, restated=restated)
            gx, gf = torch.autograd.grad((out * go).sum(), [x, flow])
            tag = "warp_mask%d_" % int(use_mask)
            assert rel_err(out, d[tag + "out"]) < 1e-6
            assert rel_err(gx, d[tag + "grad_x"]) < 1e-5
            assert rel_err(gf, d[tag + "grad_flow"]) < 1e-5
    for restated in (False, True):
        s = P.ssim_map(d["ssim_x"], d["ssim_y"], restated=restated)
        assert rel_err(s, d["ssim_out"]) < 1e-5
    for l in range(3):
        assert torch.equal(P.box_pyramid(d["pyr_img"], 3)[l], d["pyr_box_%d" % l])
        assert torch.equal(P.bilinear_pyramid(d["pyr_img"], 3)[l], d["pyr_bilinear_%d" % l])
    rec, valid, proj, comp = P.reproject(d["iw_img"], d["iw_depth"], d["iw_ref_depth"], d["iw_pose"], d["iw_K"])
    assert rel_err(rec, d["iw_rec"]) < 1e-6 and torch.equal(valid, d["iw_valid"])
    assert rel_err(proj, d["iw_proj"]) < 1e-6 and rel_err(comp, d["iw_comp"]) < 1e-6
    assert rel_err(P.rigid_flow(d["iw_depth"], d["iw_pose"], d["iw_K"]), d["rf_out"]) < 1e-6


def test_cost_volume_vs_golden():
    """oracle cost volume against PWC_tf.corr_naive outputs / gradients recorded from the reference"""
    d = load_golden("cost_volume")
    for tag in ("a", "b"):
        f1, f2 = d[tag + "_f1"].requires_grad_(True), d[tag + "_f2"].requires_grad_(True)
        out = P.cost_volume(f1, f2)
        g1, g2 = torch.autograd.grad((out * d[tag + "_go"]).sum(), [f1, f2])
        assert rel_err(out, d[tag + "_out"]) < 1e-6
        assert rel_err(g1, d[tag + "_g1"]) < 1e-6 and rel_err(g2, d[tag + "_g2"]) < 1e-6


def test_cost_volume_known_answers():
    f = torch.zeros(1, 2, 5, 6)
    f[0, :, 2, 3] = torch.tensor([2.0, 4.0])
    cv = P.cost_volume(f, f)                         # only the zero displacement (channel 40) sees the pixel against itself
    assert float(cv[0, 40, 2, 3]) == 10.0 and float(cv.abs().sum()) == 10.0
    g = torch.zeros(1, 2, 5, 6)
    g[0, :, 0, 1] = 1.0                              # partner displaced by (-2, -2) -> channel (d-2)*9 + (d-2) = 20
    cv = P.cost_volume(f, g)
    assert float(cv[0, 20, 2, 3]) == 3.0 and float(cv.abs().sum()) == 3.0


# ---- known-answer cases -------------------------------------------------------------------------------
def test_zero_flow_resamples_with_the_half_pixel_shift():
    """even zero flow samples at ix = x*W/(W-1) - 0.5 (align_corners mismatch)."""
    W, H = 9, 5
    ramp = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).expand(1, 1, H, W).contiguous()
    out = P.flow_backwarp(ramp, torch.zeros(1, 2, H, W))
    j = torch.arange(W, dtype=torch.float32)
    expect = (j * W / (W - 1) - 0.5)
    # interior columns interpolate the ramp exactly; column 0 and W-1 touch the zero padding
    assert torch.allclose(out[0, 0, 2, 1:-1], expect[1:-1], atol=1e-5)
    assert out[0, 0, 2, 0] == 0.0 * 0.5 + 0.0   # ix=-0.5: half of pixel 0 (value 0) + half padding
    assert torch.allclose(out[0, 0, 2, -1], torch.tensor(0.5 * (W - 1)), atol=1e-5)   # ix=W-0.5: half of last pixel


def test_ssim_of_identical_images_is_one():
    x = torch.rand(1, 3, 8, 8)
    assert torch.allclose(P.ssim_map(x, x), torch.ones(1, 3, 8, 8), atol=1e-5)


def test_fully_masked_loss_is_zero():
    img = [torch.rand(2, 3, 8, 8)]
    assert torch.equal(P.photometric_l1(img, [torch.rand(2, 3, 8, 8)], [torch.zeros(2, 1, 8, 8)], 1), torch.zeros(2))


def test_identity_pose_rigid_flow_is_zero():
    t = torch.rand(1, 1, 6, 10) + 0.1
    K = torch.tensor([[[5.8, 0, 5.0], [0, 11.5, 3.0], [0, 0, 1.0]]])
    assert P.rigid_flow(t, torch.zeros(1, 6), K).abs().max() < 1e-4


@pytest.mark.skipif(not R.available(), reason="reference tree not mounted (build container only)")
def test_oracle_pinned_against_live_reference():
    script = os.path.join(os.path.dirname(os.path.abspath(P.__file__)), "validate_against_reference.py")
    res = subprocess.run([sys.executable, script, "--height", "32", "--width", "64"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
