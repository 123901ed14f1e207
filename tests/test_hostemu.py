"""CPU tests of the CUDA kernels' tile logic through the host emulator (tests/hostemu/): the same
__host__ __device__ phase functions the kernels run, driven tile by tile on the CPU and compared
with the oracle.  The real kernels are tested on the B200 in the -m gpu files."""
import ctypes as C

import pytest
import torch

import hostemu_util as H
from oracle import loss_port as P
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
from util import load_golden, golden_triplet, rel_err, loss_rel_err, LOSS_RTOL, GRAD_RTOL

KEYS = ["loss_flow_pixel", "loss_flow_ssim", "loss_flow_smooth", "loss_flow_consis"]


def _oracle_flow(t, scales, gl):
    for f in t.flows_fwd + t.flows_bwd:
        f.requires_grad_(True)
    loss = P.flow_mode_loss(t.img_l, t.img, t.img_r, t.flows_fwd, t.flows_bwd, scales)
    tot = sum((gl[k] * loss[KEYS[k]]).sum() for k in range(4))
    g = torch.autograd.grad(tot, t.flows_fwd[:scales] + t.flows_bwd[:scales])
    return loss, g[:scales], g[scales:]


@pytest.mark.parametrize("B,Hh,W,scales,px,oob", [(2, 64, 208, 4, 6.0, 0.0), (1, 48, 80, 4, 3.0, 0.3), (2, 40, 72, 3, 1.0, 0.0),
                                                   (1, 34, 50, 2, 2.0, 0.1)])
def test_flow_loss_tiles_vs_oracle(B, Hh, W, scales, px, oob):
    L = 4 if Hh % 8 == 0 else 2
    t = make_triplet(B, Hh, W, L, 1, seed=11, flow_px=px, oob_fraction=oob)
    gl = torch.rand(4, B, generator=torch.Generator().manual_seed(1)) + 0.5
    loss, gf, gb = _oracle_flow(t, scales, gl)
    pl, pc, pr = P.box_pyramid(t.img_l, L), P.box_pyramid(t.img, L), P.box_pyramid(t.img_r, L)
    ff = [f.detach().contiguous() for f in t.flows_fwd]
    fb = [f.detach().contiguous() for f in t.flows_bwd]
    # recompute kernels, fused single-pass tile kernel, split kernels (+ combine), split kernels as the fused training step
    for emu_fn in (H.emu_flow_loss, H.emu_flow_loss_single_pass, H.emu_flow_loss_split, H.emu_flow_loss_step):
        el, egf, egb, _ = emu_fn(pl, pc, pr, ff, fb, scales, gl)
        for k in range(4):
            assert loss_rel_err(el[k], loss[KEYS[k]]) < LOSS_RTOL, KEYS[k]
        for l in range(scales):
            assert rel_err(egf[l], gf[l]) < GRAD_RTOL, ("fwd", l)
            assert rel_err(egb[l], gb[l]) < GRAD_RTOL * 1.5, ("bwd", l)   # SSIM-dominated: fp32 noise of the formula itself


def test_flow_loss_tiles_vs_golden():
    d = load_golden("flow_mode_s4")
    t = golden_triplet(d)
    B = t.img.shape[0]
    w = torch.tensor([P.FLOW_WEIGHTS[k] for k in KEYS]).view(4, 1).repeat(1, B) / B
    pl, pc, pr = P.box_pyramid(t.img_l, 4), P.box_pyramid(t.img, 4), P.box_pyramid(t.img_r, 4)
    el, egf, egb, _ = H.emu_flow_loss(pl, pc, pr, t.flows_fwd, t.flows_bwd, 4, w.contiguous())
    for k in range(4):
        assert loss_rel_err(el[k], d["out_" + KEYS[k]]) < LOSS_RTOL
    for l in range(4):
        assert rel_err(egf[l], d["grad_flows_fwd_%d" % l]) < GRAD_RTOL
        assert rel_err(egb[l], d["grad_flows_bwd_%d" % l]) < GRAD_RTOL * 1.5


def test_pyramid_pixels_bit_exact():
    img = torch.rand(2, 3, 32, 64)
    for mode, ref in ((0, P.box_pyramid(img, 4)), (1, P.bilinear_pyramid(img, 4))):
        outs = [img] + [torch.empty(2, 3, 32 >> l, 64 >> l) for l in range(1, 4)]
        arr = (C.c_void_p * 4)(*[o.data_ptr() for o in outs])
        H.emu().emu_image_pyramid(C.c_void_p(img.data_ptr()), 2, 3, 32, 64, 4, mode, arr)
        for l in range(1, 4):
            if mode == 0:
                assert torch.equal(outs[l], ref[l]), (mode, l)
            else:   # ATen's CPU bilinear path is size/thread dependent in the last ulp (see ugl_primitives.cuh)
                assert (outs[l] - ref[l]).abs().max() <= 1.2e-7, (mode, l)


@pytest.mark.parametrize("use_mask", [0, 1])
def test_warp_pixels_vs_oracle(use_mask):
    d = load_golden("primitives")
    x, flow, go = d["warp_x"].contiguous(), d["warp_flow"].contiguous(), d["warp_go"].contiguous()
    B, Cc, Hh, W = x.shape
    out, mask = torch.empty_like(x), torch.empty(B, 1, Hh, W)
    gflow, gx = torch.empty_like(flow), torch.empty_like(x)
    p = lambda t: C.c_void_p(t.data_ptr())
    H.emu().emu_warp_flow_forward(p(x), p(flow), B, Cc, Hh, W, use_mask, p(out), p(mask))
    H.emu().emu_warp_flow_backward(p(x), p(flow), p(go), B, Cc, Hh, W, use_mask, p(gflow), p(gx))
    tag = "warp_mask%d_" % use_mask
    assert rel_err(out, d[tag + "out"]) < 1e-6
    assert rel_err(gflow, d[tag + "grad_flow"]) < 1e-5
    assert rel_err(gx, d[tag + "grad_x"]) < 1e-5
    # the keep mask is bit-exact against the oracle's
    with torch.no_grad():
        cover = P.grid_sample_restated(torch.ones(B, 1, Hh, W), _grid(flow))
    assert torch.equal(mask, (cover >= 0.9999).float() if use_mask else torch.ones_like(mask))


def _grid(flow):
    B, _, Hh, W = flow.shape
    tgt = P._pixel_grid(B, Hh, W, flow) + flow
    return torch.stack([2.0 * tgt[:, 0] / max(W - 1, 1) - 1.0, 2.0 * tgt[:, 1] / max(Hh - 1, 1) - 1.0], -1)


def _geom_projections(t, S):
    """K_s^-1 and K_s [R|t] per level, as the oracle forms them (scale_intrinsics + pose_to_matrix)"""
    H0 = t.img.shape[2]
    Kinv, Pb, Pf = [], [], []
    for s in range(S):
        Ks = P.scale_intrinsics(t.K, H0 / (H0 >> s))
        Kinv.append(Ks.inverse().contiguous())
        Pb.append((Ks @ P.pose_to_matrix(t.pose[:, 0])).contiguous())
        Pf.append((Ks @ P.pose_to_matrix(t.pose[:, 1])).contiguous())
    return Kinv, Pb, Pf


@pytest.mark.parametrize("split", [False, True, "step"])
@pytest.mark.parametrize("B,Hh,W,S", [(2, 48, 96, 3), (1, 40, 72, 2), (1, 34, 50, 1)])
def test_geom_flow_tiles_vs_oracle(B, Hh, W, S, split):
    """geom-mode variant of the single-pass kernel (split=False: fused tile kernel; True: photometry pixels + stencil tiles; "step": the
    fused training step -- scale factors from the weight sums, gradients written by the stencil tiles): flow-branch losses, flow
    gradients and the packed masks"""
    t = make_triplet(B, Hh, W, flow_levels=S, depth_scales=S, seed=11, flow_mode="rigid", flow_px=1.5)
    keys = ("loss_flow_pixel", "loss_flow_ssim", "loss_flow_smooth", "loss_flow_consis")
    gl = torch.rand(4, B, generator=torch.Generator().manual_seed(2)) + 0.5
    ff = [f.detach().clone().requires_grad_(True) for f in t.flows_fwd]
    fb = [f.detach().clone().requires_grad_(True) for f in t.flows_bwd]
    loss, aux = P.geom_mode_loss(t.img_l, t.img, t.img_r, ff, fb, t.disp, t.disp_l, t.disp_r, t.pose, t.K, t.K_inv, S, return_aux=True)
    sum((loss[k] * gl[i]).sum() for i, k in enumerate(keys)).backward()
    pl, pc, pr = P.bilinear_pyramid(t.img_l, S), P.bilinear_pyramid(t.img, S), P.bilinear_pyramid(t.img_r, S)
    Kinv, Pb, Pf = _geom_projections(t, S)
    el, egf, egb, masks = H.emu_geom_flow(pl, pc, pr, [f.detach().contiguous() for f in ff], [f.detach().contiguous() for f in fb],
                                          [d.detach().contiguous() for d in t.disp[:S]], Kinv, Pb, Pf, 0.01, 0.5, S, gl, split=split)
    for bit, name in ((1, "valid_b"), (2, "valid_f"), (4, "occ_b"), (8, "occ_f"), (16, "dyn_b"), (32, "dyn_f")):
        for s in range(S):
            got = ((masks[s] & bit) != 0).float().unsqueeze(1)
            flips = int((got != aux[name][s]).sum())
            assert flips == 0, (name, s, flips)     # the emulator runs without fma contraction: bit-exact masks
    for i, k in enumerate(keys):
        assert loss_rel_err(el[i], loss[k]) < LOSS_RTOL, k
    for l in range(S):
        assert rel_err(egf[l], ff[l].grad) < GRAD_RTOL, ("fwd", l)
        assert rel_err(egb[l], fb[l].grad) < GRAD_RTOL * 1.5, ("bwd", l)


@pytest.mark.parametrize("B,Hh,W,S", [(2, 48, 96, 3), (1, 40, 72, 2)])
def test_depth_ssim_tiles_vs_oracle(B, Hh, W, S):
    """depth-mode variant of the single-pass kernel (reprojection warps): loss_depth_pixel + loss_depth_ssim of the
    model_depth_texture spec, the masks, and the gradients w.r.t. the centre disparity and K_s [R|t]"""
    t = make_triplet(B, Hh, W, flow_levels=S, depth_scales=S, seed=13, flow_mode="rigid")
    gl = torch.rand(2, B, generator=torch.Generator().manual_seed(4)) + 0.5
    disp = [d.detach().clone().requires_grad_(True) for d in t.disp]
    Kinv, Pb, Pf = _geom_projections(t, S)
    Pl = [[p.clone().requires_grad_(True) for p in Pb], [p.clone().requires_grad_(True) for p in Pf]]
    # oracle terms with P as an explicit leaf: reproject through P directly
    pc, pl, pr = P.bilinear_pyramid(t.img, S), P.bilinear_pyramid(t.img_l, S), P.bilinear_pyramid(t.img_r, S)
    area = [P.box_pyramid(t.img_l, S), P.box_pyramid(t.img_r, S)]
    bil = [pl, pr]
    tot, pix, ssim, valid, tex = 0, 0, 0, [[], []], [[], []]
    for d in range(2):
        rec, val = [], []
        for s in range(S):
            h, w = disp[s].shape[2:]
            jj = torch.arange(w, dtype=torch.float32).view(1, 1, w).expand(1, h, w)
            ii = torch.arange(h, dtype=torch.float32).view(1, h, 1).expand(1, h, w)
            pixg = torch.stack((jj, ii, torch.ones_like(jj)), 1).expand(B, 3, h, w).reshape(B, 3, -1)
            cam = ((Kinv[s] @ pixg).reshape(B, 3, h, w) * disp[s]).reshape(B, 3, -1)
            q = Pl[d][s][:, :, :3] @ cam + Pl[d][s][:, :, 3:]
            Z = q[:, 2].clamp(min=1e-3)
            gx, gy = 2 * (q[:, 0] / Z) / (w - 1) - 1, 2 * (q[:, 1] / Z) / (h - 1) - 1
            gx = torch.where((gx > 1) | (gx < -1), torch.full_like(gx, 2.0), gx)
            gy = torch.where((gy > 1) | (gy < -1), torch.full_like(gy, 2.0), gy)
            grid = torch.stack([gx, gy], 2).reshape(B, h, w, 2)
            rec.append(torch.nn.functional.grid_sample(area[d][s], grid, padding_mode="zeros", align_corners=False))
            val.append((grid.abs().max(-1)[0] <= 1).unsqueeze(1).float())
        tx = P.texture_mask(pc, rec, bil[d], S)
        m = [val[s] * tx[s] for s in range(S)]
        pix = pix + P.photometric_l1(pc, rec, m, S)
        ssim = ssim + P.ssim_loss(pc, rec, val, S)
        valid[d], tex[d] = val, tx
    ((pix * gl[0]).sum() + (ssim * gl[1]).sum()).backward()
    el, egd, egP, ev, et = H.emu_depth_ssim(pc, area, bil, [d.detach().contiguous() for d in disp], Kinv,
                                            [[p.detach().contiguous() for p in Pl[0]], [p.detach().contiguous() for p in Pl[1]]], gl)
    assert loss_rel_err(el[0], pix) < LOSS_RTOL and loss_rel_err(el[1], ssim) < LOSS_RTOL
    assert float(el[2:].abs().max()) == 0.0
    for d in range(2):
        for s in range(S):
            assert torch.equal(ev[d][s], valid[d][s]) and torch.equal(et[d][s], tex[d][s]), (d, s)      # masks bit-exact
            assert rel_err(egP[d][s], Pl[d][s].grad) < GRAD_RTOL, ("P", d, s)
    for s in range(S):
        assert rel_err(egd[s], disp[s].grad) < GRAD_RTOL, ("disp", s)
