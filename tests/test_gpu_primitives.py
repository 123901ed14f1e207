"""GPU parity tests of the primitive ops (warp_flow, image pyramid) through the C-ABI."""
import pytest
import torch

from oracle import loss_port as P
from unsupervised_depth_opticalflow_egomotion_b200 import ops
from util import load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_mask", [False, True])
def test_warp_flow_vs_reference_golden(cuda_device, use_mask):
    d = load_golden("primitives")
    x = d["warp_x"].to(cuda_device).requires_grad_(True)
    flow = d["warp_flow"].to(cuda_device).requires_grad_(True)
    out = ops.warp_flow(x, flow, use_mask)
    gx, gf = torch.autograd.grad((out * d["warp_go"].to(cuda_device)).sum(), [x, flow])
    tag = "warp_mask%d_" % int(use_mask)
    assert rel_err(out, d[tag + "out"]) < 1e-6
    assert rel_err(gf, d[tag + "grad_flow"]) < 1e-5
    assert rel_err(gx, d[tag + "grad_x"]) < 1e-5


@pytest.mark.parametrize("B,C,H,W,px", [(2, 3, 64, 208, 8.0), (1, 32, 32, 104, 3.0), (2, 1, 17, 23, 30.0), (1, 128, 16, 52, 1.0)])
def test_warp_flow_vs_oracle(cuda_device, B, C, H, W, px):
    g = torch.Generator().manual_seed(B * 1000 + C)
    x = torch.rand(B, C, H, W, generator=g)
    flow = px * torch.randn(B, 2, H, W, generator=g)
    go = torch.randn(B, C, H, W, generator=g)
    for use_mask in (False, True):
        xc, fc = x.clone().requires_grad_(True), flow.clone().requires_grad_(True)
        ref = P.flow_backwarp(xc, fc, use_mask)
        rgx, rgf = torch.autograd.grad((ref * go).sum(), [xc, fc])
        xd, fd = x.to(cuda_device).requires_grad_(True), flow.to(cuda_device).requires_grad_(True)
        out = ops.warp_flow(xd, fd, use_mask)
        gx, gf = torch.autograd.grad((out * go.to(cuda_device)).sum(), [xd, fd])
        assert rel_err(out, ref) < 1e-6
        assert torch.equal(out.cpu() == 0, ref == 0)            # the keep mask zeroes exactly the same pixels
        assert rel_err(gf, rgf) < 1e-5 and rel_err(gx, rgx) < 1e-5
        gx2, = torch.autograd.grad((ops.warp_flow(xd, fd, use_mask) * go.to(cuda_device)).sum(), [xd])
        assert torch.equal(gx, gx2)                               # deterministic scatter (no fp atomics)


@pytest.mark.parametrize("B,C,H,W,px", [(2, 3, 64, 208, 2.0), (1, 33, 32, 104, 6.0), (2, 1, 17, 23, 30.0), (1, 130, 16, 52, 1.0),
                                        (1, 6, 45, 70, 12.0)])
def test_scatter_forms_same_bits(cuda_device, monkeypatch, B, C, H, W, px):
    """The tile-local scatter (shared-memory window per CTA, lo/hi 32-bit words with carry) and the global-atomic form sum the same
    64-bit fixed-point terms: grad_x of warp_flow and the forward splat are bit-identical, whatever leaves the window (px > 8)."""
    g = torch.Generator().manual_seed(B * 100 + C)
    x = torch.rand(B, C, H, W, generator=g).to(cuda_device)
    flow = (px * torch.randn(B, 2, H, W, generator=g)).to(cuda_device)
    flow[:, :, : H // 3] *= 0.05                                   # a sub-pixel region: every tap of a tile lands in a few cells (carries)
    go = (1e3 * torch.randn(B, C, H, W, generator=g)).to(cuda_device)
    res = {}
    for form in ("tile_local", "global"):
        monkeypatch.setattr(ops, "SCATTER_FORM", form)
        for use_mask in (False, True):
            xd = x.clone().requires_grad_(True)
            res[form, use_mask], = torch.autograd.grad((ops.warp_flow(xd, flow, use_mask) * go).sum(), [xd])
        res[form, "splat"] = ops.forward_splat(x, flow)
    for key in (False, True, "splat"):
        assert torch.equal(res["tile_local", key], res["global", key]), key
    assert res["tile_local", False].abs().max() > 0


@pytest.mark.parametrize("mode", ["box", "bilinear"])
def test_image_pyramid_bit_exact(cuda_device, mode):
    img = torch.rand(2, 3, 256, 832, generator=torch.Generator().manual_seed(4))
    ref = P.box_pyramid(img, 4) if mode == "box" else P.bilinear_pyramid(img, 4)
    out = ops.image_pyramid(img.to(cuda_device), 4, mode)
    for l in range(4):
        if mode == "box":
            assert torch.equal(out[l].cpu(), ref[l]), l
        else:   # ATen's CPU bilinear path is size/thread dependent in the last ulp (see ugl_primitives.cuh)
            assert (out[l].cpu() - ref[l]).abs().max() <= 1.2e-7, l
    with pytest.raises(ValueError):
        ops.image_pyramid(torch.rand(1, 3, 30, 64, device=cuda_device), 4, mode)


def test_forward_splat_extension(cuda_device):
    """`transformerFwd` is undefined in the reference (dead code): independent CPU oracle + known answers."""
    from unsupervised_depth_opticalflow_egomotion_b200 import losses
    g = torch.Generator().manual_seed(12)
    x = torch.rand(2, 3, 24, 40, generator=g)
    flow = 4.0 * torch.randn(2, 2, 24, 40, generator=g)
    ref = P.forward_splat(x, flow)
    out = ops.forward_splat(x.to(cuda_device), flow.to(cuda_device))
    assert rel_err(out, ref) < 1e-5
    assert torch.equal(out, ops.forward_splat(x.to(cuda_device), flow.to(cuda_device)))          # deterministic
    zero = torch.zeros(2, 2, 24, 40, device=cuda_device)
    assert torch.allclose(ops.forward_splat(x.to(cuda_device), zero), x.to(cuda_device), atol=1e-6)   # zero flow = identity
    shift = zero.clone(); shift[:, 0] = 3.0                                                            # integer shift
    moved = ops.forward_splat(x.to(cuda_device), shift).cpu()
    assert torch.allclose(moved[..., 3:], x[..., :-3], atol=1e-6) and moved[..., :3].abs().max() == 0
    occ = losses.FlowLoss(1).get_occlusion_mask_from_flow((2, 1, 24, 40), flow.to(cuda_device))
    assert occ.shape == (2, 1, 24, 40) and float(occ.min()) >= 0.0 and float(occ.max()) <= 1.0
    assert rel_err(occ, P.forward_splat(torch.ones(2, 1, 24, 40), flow).clamp(0, 1)) < 1e-5


@pytest.mark.parametrize("levels,H,W", [(2, 34, 50), (3, 64, 208), (4, 64, 208), (4, 256, 832), (5, 64, 128)])
def test_image_pyramids_one_launch_bit_equal(cuda_device, levels, H, W):
    """ops.image_pyramids (all levels, both modes, three images in one launch; per-level fallback outside 2..4 levels or for
    unaligned sizes) is bit-identical to the per-level kernel"""
    g = torch.Generator().manual_seed(3)
    imgs = [torch.rand(2, 3, H, W, generator=g).to(cuda_device) for _ in range(3)]
    out = ops.image_pyramids(imgs, levels, ("bilinear", ("bilinear", "area"), "box"))
    assert set(out[0]) == {"bilinear"} and set(out[1]) == {"bilinear", "area"} and set(out[2]) == {"box"}
    for i, d in enumerate(out):
        for mode, pyr in d.items():
            ref = ops.image_pyramid(imgs[i], levels, mode)
            assert len(pyr) == levels and pyr[0] is not None
            for l in range(levels):
                assert torch.equal(pyr[l], ref[l]), (i, mode, l)


def test_cost_volume_vs_reference_golden(cuda_device):
    d = load_golden("cost_volume")
    for tag in ("a", "b"):
        f1, f2 = d[tag + "_f1"].to(cuda_device).requires_grad_(True), d[tag + "_f2"].to(cuda_device).requires_grad_(True)
        out = ops.cost_volume(f1, f2)
        g1, g2 = torch.autograd.grad((out * d[tag + "_go"].to(cuda_device)).sum(), [f1, f2])
        assert rel_err(out.cpu(), d[tag + "_out"]) < 1e-5
        assert rel_err(g1.cpu(), d[tag + "_g1"]) < 1e-5 and rel_err(g2.cpu(), d[tag + "_g2"]) < 1e-5


@pytest.mark.parametrize("B,C,H,W,d", [(2, 32, 64, 208, 4), (1, 196, 4, 13, 4), (2, 7, 9, 33, 2), (1, 3, 2, 3, 1)])
def test_cost_volume_vs_oracle(cuda_device, B, C, H, W, d):
    """PWC pyramid shapes (level 2 and the 4x13 top level), sizes smaller than the search window, other radii"""
    g = torch.Generator().manual_seed(11)
    f1c, f2c = torch.randn(B, C, H, W, generator=g).requires_grad_(True), torch.randn(B, C, H, W, generator=g).requires_grad_(True)
    go = torch.randn(B, (2 * d + 1) ** 2, H, W, generator=g)
    ref = P.cost_volume(f1c, f2c, d)
    r1, r2 = torch.autograd.grad((ref * go).sum(), [f1c, f2c])
    f1, f2 = f1c.detach().to(cuda_device).requires_grad_(True), f2c.detach().to(cuda_device).requires_grad_(True)
    out = ops.cost_volume(f1, f2, d)
    g1, g2 = torch.autograd.grad((out * go.to(cuda_device)).sum(), [f1, f2])
    assert rel_err(out.cpu(), ref.detach()) < 1e-5
    assert rel_err(g1.cpu(), r1) < 1e-5 and rel_err(g2.cpu(), r2) < 1e-5
    out2 = ops.cost_volume(f1, f2, d)
    assert torch.equal(out, out2)
    with pytest.raises(AssertionError):
        ops.cost_volume(f1, f2[:, :, :, :-1].contiguous() if W > 1 else f2[:, :-1], d)
    # only one input needs a gradient
    f3 = f2c.detach().to(cuda_device)
    g1b, = torch.autograd.grad((ops.cost_volume(f1, f3, d) * go.to(cuda_device)).sum(), [f1])
    assert torch.equal(g1b, g1)


def test_packed_pair_ssim_is_bit_identical_to_scalar(cuda_device):
    """The FADD2 / FMUL2 / FFMA2 forms of the SSIM moments, terms and backward coefficients (both warp directions per
    instruction) must reproduce the scalar forms bit for bit: ptxas contracts packed `.rn` products into sums unless the
    code prevents it (ugl_common.cuh: acc2_rn / sub2_rn), which this test would catch as sxx / syy / sxy / d2 / S mismatches."""
    bad = ops.selftest_packed_pairs(cuda_device, blocks=64, windows_per_thread=500)
    assert bad.shape == (2, 14)
    assert int(bad.sum()) == 0, bad.cpu().tolist()


def test_frames_from_u8_matches_the_dataset_division(cuda_device):
    """core/dataset/kitti_prepared.py:89: `img / 255.0` (numpy float64) then `.float()` — all 256 byte values, plus a ragged size."""
    import numpy as np
    k = np.arange(256, dtype=np.uint8)
    ref = torch.from_numpy((k / 255.0).astype(np.float32))
    x = torch.from_numpy(np.tile(k, 37)[: 256 * 37 - 5].copy()).to(cuda_device)          # not a multiple of 16
    y = torch.from_numpy(np.roll(np.tile(k, 37), 3)[: 256 * 37 - 5].copy()).to(cuda_device)
    ox, oy = ops.frames_from_u8([x, y])
    assert torch.equal(ox.cpu(), ref[x.cpu().long()]) and torch.equal(oy.cpu(), ref[y.cpu().long()])
    with pytest.raises(TypeError):
        ops.frames_from_u8([x.float()])


def test_flow_loss_step_uint8_frames_equal_float_frames(cuda_device):
    from unsupervised_depth_opticalflow_egomotion_b200.step import FlowLossStep
    from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
    t = make_triplet(2, 64, 96, 3, 1, seed=11)
    u8 = [(x * 255.0).round().clamp(0, 255).to(torch.uint8).pin_memory() for x in (t.img_l, t.img, t.img_r)]
    f32 = [(x.double() / 255.0).float().pin_memory() for x in u8]
    ff, fb = [f.detach().pin_memory() for f in t.flows_fwd], [f.detach().pin_memory() for f in t.flows_bwd]
    a = FlowLossStep(2, 64, 96, 3, device=cuda_device, frame_dtype=torch.uint8)
    b = FlowLossStep(2, 64, 96, 3, device=cuda_device)
    la = a(u8[0], u8[1], u8[2], ff, fb)
    lb = b(f32[0], f32[1], f32[2], ff, fb)
    assert torch.equal(la, lb)
    assert all(torch.equal(x, y) for x, y in zip(a.grads(0), b.grads(0))) and len(a.grads(0)) == 6
    assert a.h2d_bytes < b.h2d_bytes and a.h2d_copies_per_step == 1
    with pytest.raises(TypeError):
        a(f32[0], f32[1], f32[2], ff, fb)


def test_flow_loss_step_staged_pipeline_and_guards(cuda_device):
    """The staged form (inputs written into the slot's pinned views, one H2D per step), two steps in flight, against the
    resident-input path; a slot may not be re-submitted before its result was read; results are private copies."""
    from unsupervised_depth_opticalflow_egomotion_b200.step import FlowLossStep, FLOW_WEIGHTS, weight_matrix
    from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
    B, H, W, L = 2, 64, 96, 3
    ts = [make_triplet(B, H, W, L, 1, seed=s) for s in (21, 22, 23)]
    st = FlowLossStep(B, H, W, L, device=cuda_device)
    wmat = weight_matrix(FLOW_WEIGHTS, ops.FLOW_LOSS_KEYS, B, cuda_device)

    def fill(slot, t):
        hv = st.host_views(slot)
        hv["img_l"].copy_(t.img_l); hv["img"].copy_(t.img); hv["img_r"].copy_(t.img_r)
        for l in range(L):
            hv["ff%d" % l].copy_(t.flows_fwd[l].detach()); hv["fb%d" % l].copy_(t.flows_bwd[l].detach())

    def resident(t):
        td = t.to(cuda_device)
        pyr = [ops.image_pyramid(x, L, "box") for x in (td.img_l, td.img, td.img_r)]
        out = ops.flow_loss_step(pyr[0], pyr[1], pyr[2], td.flows_fwd, td.flows_bwd, wmat, L)
        return out["loss"].cpu(), [g.cpu() for g in out["gf"] + out["gb"]]

    s0 = st.next_slot(); fill(s0, ts[0]); assert st.submit_staged() == s0
    s1 = st.next_slot(); fill(s1, ts[1]); assert st.submit_staged() == s1 and s1 != s0
    with pytest.raises(RuntimeError, match="re-submitted before result"):
        st.submit_staged()                                  # slot s0 again: its result is still unread
    r0 = st.result(s0)
    g0 = [g.cpu() for g in st.grads(s0)]
    fill(s0, ts[2]); st.submit_staged()                     # now legal; must not disturb r0 (a private copy)
    r1, r2 = st.result(s1), st.result(s0)
    for r, t in ((r0, ts[0]), (r1, ts[1]), (r2, ts[2])):
        ref_l, _ = resident(t)
        assert torch.equal(r, ref_l)
    assert all(torch.equal(a, b) for a, b in zip(g0, resident(ts[0])[1]))
    with pytest.raises(RuntimeError, match="nothing was submitted"):
        FlowLossStep(B, H, W, L, device=cuda_device).result(0)
