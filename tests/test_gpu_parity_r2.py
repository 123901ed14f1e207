"""Round-2 GPU parity tests (VERDICT r1 "next round" item 1 and the advisor's findings):

* depth mode (BASELINE configs[2], both reference variants) against the oracle at multi-tile and full sizes, and the depth half
  of configs[4] at 384x1280 -- until now only compared with the reference at 32x64;
* the multi-tile golden fixtures (112x168, ragged tile remainders, odd level widths) generated from the reference;
* geom-mode pose gradients term by term (depth L1, depth-flow consistency, epipolar), each against its own fp32-vs-fp64 floor,
  and ``ugl_pose_setup_backward`` against the torch chain in fp64;
* ``get_rigid_mask`` (M7) bit-exactly, including values at the thresholds, and the lazily unpacked ``GeomMasks`` entries;
* the internal forms of the single-pass forward (fused / split with plain staging / split with TMA staging) against each other and
  the fused training step (``ugl_flow_loss_step``) against the oracle.
"""
import pytest
import torch

from oracle import loss_port as P
from unsupervised_depth_opticalflow_egomotion_b200 import losses, ops, structures
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
from util import load_golden, golden_triplet, rel_err, loss_rel_err, assert_loss_close, assert_grad_close, LOSS_RTOL, GRAD_RTOL

pytestmark = pytest.mark.gpu
KEYS = list(ops.FLOW_LOSS_KEYS)


def _leaf_list(xs, dev, dtype=torch.float32):
    return [x.detach().to(device=dev, dtype=dtype).requires_grad_(True) for x in xs]


# ---- depth mode against the oracle at multi-tile / full size ----------------------------------------------------------
def _oracle_depth(t, variant, S, dtype):
    cv = lambda x: x.detach().to(dtype)
    disp, disp_l, disp_r = _leaf_list(t.disp, "cpu", dtype), _leaf_list(t.disp_l, "cpu", dtype), _leaf_list(t.disp_r, "cpu", dtype)
    pose = cv(t.pose).requires_grad_(True)
    loss, aux = P.depth_mode_loss(cv(t.img_l), cv(t.img), cv(t.img_r), disp, disp_l, disp_r, pose, cv(t.K), S, variant, return_aux=True)
    keys = [k for k, v in loss.items() if v.requires_grad]
    tot = sum(P.GEOM_WEIGHTS[k] * loss[k].mean() for k in keys)
    g = torch.autograd.grad(tot, disp + disp_l + disp_r + [pose], allow_unused=True)
    return loss, aux, keys, g


@pytest.mark.parametrize("variant,B,H,W", [("live", 2, 128, 416), ("texture", 2, 128, 416), ("ssim", 2, 128, 416), ("live", 1, 256, 832),
                                           ("texture", 1, 256, 832), ("ssim", 1, 256, 832), ("live", 1, 384, 1280)])
def test_depth_mode_vs_oracle(cuda_device, variant, B, H, W):
    """model_depth.py:281-335 ('live') / model_depth_texture.py:296-311 ('texture') through losses.DepthLoss (fused kernels):
    masks bit-exact, losses 1e-5, disparity gradients 1e-4, pose gradient against the fp64 twin."""
    S, dev = 3, cuda_device
    t = make_triplet(B, H, W, 4, S, seed=71, flow_mode="rigid")
    ref, raux, keys, rg = _oracle_depth(t, variant, S, torch.float32)
    ref64, _, _, rg64 = _oracle_depth(t, variant, S, torch.float64)
    disp, disp_l, disp_r = _leaf_list(t.disp, dev), _leaf_list(t.disp_l, dev), _leaf_list(t.disp_r, dev)
    pose = t.pose.detach().to(dev).requires_grad_(True)
    loss, masks = losses.DepthLoss(S, variant).forward_losses(t.img_l.to(dev), t.img.to(dev), t.img_r.to(dev), disp, disp_l, disp_r, pose, t.K.to(dev))
    for k in keys:
        assert_loss_close(k, loss[k], ref[k], ref64[k])
    # reprojection valid masks: bit-exact.  Texture masks [mean|I - rec| < mean|I - src|]: bit-exact except where the ORACLE's
    # own decision is a knife edge -- this path builds K_s^-1 and K_s [R|t] with its pose set-up kernel, the oracle with the
    # CPU BLAS, and the two differ in the last bit of some matrix elements (the reference run on CUDA differs from its CPU run
    # in the same way); test_depth_masks_bit_exact_given_matrices below pins the kernels with identical matrices.
    pc, pl, pr = P.bilinear_pyramid(t.img, S), P.bilinear_pyramid(t.img_l, S), P.bilinear_pyramid(t.img_r, S)
    for l in range(S):
        assert torch.equal(masks["valid_l"][l].cpu(), raux["valid_l"][l]) and torch.equal(masks["valid_r"][l].cpu(), raux["valid_r"][l])
        for mk, rec, src in (("tex_b", raux["rec_l"][l], pl[l]), ("tex_f", raux["rec_r"][l], pr[l])):
            flip = masks[mk][l].cpu() != raux[mk][l]
            if flip.any():
                margin = ((pc[l] - rec.detach()).abs().mean(1, keepdim=True) - (pc[l] - src).abs().mean(1, keepdim=True)).abs()
                assert int(flip.sum()) <= max(1, int(1e-5 * flip.numel())) and float(margin[flip].max()) < 2e-6, (mk, l, int(flip.sum()))
    tot = sum(P.GEOM_WEIGHTS[k] * loss[k].mean() for k in keys)
    og = torch.autograd.grad(tot, disp + disp_l + disp_r + [pose], allow_unused=True)
    n = len(og)
    for i, (a, b, c) in enumerate(zip(og, rg, rg64)):
        if b is None:
            assert a is None or float(a.abs().max()) == 0.0
            continue
        if i == n - 1:
            # pose (B,2,6): a sum over every pixel; its fp32 noise floor is what the fp32 oracle itself shows against fp64
            e32, e_got, e_ref = rel_err(a, b), rel_err(a, c), rel_err(b, c)
            assert e32 < GRAD_RTOL or e_got <= 2.0 * e_ref + 1e-5, ("pose", e32, e_got, e_ref)
        else:
            assert_grad_close("leaf %d" % i, a, b, c)


def test_depth_masks_bit_exact_given_matrices(cuda_device):
    """The reprojection-photometric kernel with the ORACLE's matrices (torch.inverse / bmm on the CPU, as inverse_warp.py:284-289
    evaluates them) at 256x832: valid and texture masks of both directions and all levels bit-exact, loss 1e-5."""
    S, dev = 3, cuda_device
    t = make_triplet(1, 256, 832, 4, S, seed=71, flow_mode="rigid")
    _, raux = P.depth_mode_loss(t.img_l, t.img, t.img_r, t.disp, t.disp_l, t.disp_r, t.pose, t.K, S, "live", return_aux=True)
    ref = P.depth_mode_loss(t.img_l, t.img, t.img_r, t.disp, t.disp_l, t.disp_r, t.pose, t.K, S, "live")
    Ks = [P.scale_intrinsics(t.K, float(1 << s)) for s in range(S)]
    Kinv = [torch.inverse(k).contiguous().to(dev) for k in Ks]
    Pm = [[k.bmm(P.pose_to_matrix(t.pose[:, idx])).contiguous().to(dev) for k in Ks] for idx in (0, 1)]
    td = t.to(dev)
    pyr = ops.image_pyramids((td.img, td.img_l, td.img_r), S, ("bilinear", ("bilinear", "area"), ("bilinear", "area")))
    pc, pl, pr = (d["bilinear"] for d in pyr)
    pix, valid, tex = ops.depth_photo_loss(pc, (pyr[1]["area"], pyr[2]["area"]), (pl, pr), list(td.disp[:S]), Kinv, (Pm[0], Pm[1]))
    assert loss_rel_err(pix, ref["loss_depth_pixel"]) < LOSS_RTOL
    for l in range(S):
        assert torch.equal(valid[0][l].cpu(), raux["valid_l"][l]) and torch.equal(valid[1][l].cpu(), raux["valid_r"][l])
        assert torch.equal(tex[0][l].cpu(), raux["tex_b"][l]) and torch.equal(tex[1][l].cpu(), raux["tex_f"][l])


# ---- multi-tile golden fixtures generated from the reference ----------------------------------------------------------
def test_flow_mode_vs_reference_golden_multitile(cuda_device):
    d = load_golden("flow_mode_s4_mt")
    t = golden_triplet(d).to(cuda_device)
    B = t.img.shape[0]
    w = (torch.tensor([P.FLOW_WEIGHTS[k] for k in KEYS]).view(4, 1).repeat(1, B) / B).contiguous().to(cuda_device)
    pl, pc, pr = (ops.image_pyramid(x, 4, "box") for x in (t.img_l, t.img, t.img_r))
    for variant in ("split", "fused"):
        ops.SINGLE_PASS_VARIANT = variant
        try:
            ff = [f.detach().clone().requires_grad_(True) for f in t.flows_fwd]
            fb = [f.detach().clone().requires_grad_(True) for f in t.flows_bwd]
            loss = ops.flow_loss(pl, pc, pr, ff, fb, 4, as_matrix=True)
            g = torch.autograd.grad(loss, ff + fb, grad_outputs=w)
        finally:
            ops.SINGLE_PASS_VARIANT = "split"
        for k in range(4):
            assert loss_rel_err(loss[k], d["out_" + KEYS[k]]) < LOSS_RTOL, (variant, KEYS[k])
        for l in range(4):
            assert_grad_close("%s fwd%d" % (variant, l), g[l], d["grad_flows_fwd_%d" % l])
            assert_grad_close("%s bwd%d" % (variant, l), g[4 + l], d["grad_flows_bwd_%d" % l])
    st = ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, w, 4)          # fused training step
    for k in range(4):
        assert loss_rel_err(st["loss"][k], d["out_" + KEYS[k]]) < LOSS_RTOL, KEYS[k]
    for l in range(4):
        assert_grad_close("step fwd%d" % l, st["gf"][l], d["grad_flows_fwd_%d" % l])
        assert_grad_close("step bwd%d" % l, st["gb"][l], d["grad_flows_bwd_%d" % l])


@pytest.mark.parametrize("name,variant", [("depth_mode_live_mt", "live"), ("depth_mode_texture_mt", "texture")])
def test_depth_mode_vs_reference_golden_multitile(cuda_device, name, variant):
    d = load_golden(name)
    t = golden_triplet(d)
    dev = cuda_device
    disp, disp_l, disp_r = _leaf_list(t.disp, dev), _leaf_list(t.disp_l, dev), _leaf_list(t.disp_r, dev)
    pose = t.pose.to(dev).requires_grad_(True)
    loss, masks = losses.DepthLoss(3, variant).forward_losses(t.img_l.to(dev), t.img.to(dev), t.img_r.to(dev), disp, disp_l, disp_r, pose, t.K.to(dev))
    live = [k for k in loss if "out_" + k in d and d["out_" + k].numel() == t.img.shape[0] and float(d["out_" + k].abs().max()) > 0]
    for k in live:
        assert loss_rel_err(loss[k], d["out_" + k]) < LOSS_RTOL, k
    for l in range(3):
        assert torch.equal(masks["valid_l"][l].cpu(), d["aux_valid_l_%d" % l])
        assert torch.equal(masks["valid_r"][l].cpu(), d["aux_valid_r_%d" % l])
    tot = sum(P.GEOM_WEIGHTS[k] * loss[k].mean() for k in live)
    names = ["disp_%d" % i for i in range(3)] + ["disp_l_%d" % i for i in range(3)] + ["disp_r_%d" % i for i in range(3)] + ["pose"]
    grads = torch.autograd.grad(tot, disp + disp_l + disp_r + [pose], allow_unused=True)
    # the fixture holds the reference's fp32 gradients; its fp64 twin (the oracle, pinned to the reference at 1e-6) tells how much
    # of a difference is the reference's own rounding noise (SSIM's E[x^2] - mu^2, the depth-consistency ratio)
    _, _, keys64, g64 = _oracle_depth(t, variant, 3, torch.float64)
    assert sorted(keys64) == sorted(live)
    for n, g, c in zip(names, grads, g64):
        ref = d["grad_" + n]
        g = torch.zeros_like(ref) if g is None else g.cpu()
        if float(ref.abs().max()) == 0.0:
            assert float(g.abs().max()) == 0.0, n
        elif n == "pose":
            e32, e_got, e_ref = rel_err(g, ref), rel_err(g, c), rel_err(ref, c)
            assert e32 < GRAD_RTOL or e_got <= 2.0 * e_ref + 1e-5, (n, e32, e_got, e_ref)
        else:
            # texture variant, centre disparity: the SSIM term under the reprojection valid mask.  Next to the mask's boundary the
            # windows hold a few non-zero pixels, sigma ~ 0 and 1 / (d1 d2) ~ 1e7: measured here, the reference's own fp32 gradient
            # is 1.05e-4 (of max|g|) off its fp64 value at such pixels, this path 2.4e-4 at 8 of 18,816 pixels (the composed
            # per-method SSIM kernels give the same 8 pixels) -- same noise source, different summation order.
            assert_grad_close(n, g, ref, c, fp64_factor=2.5 if (variant == "texture" and n == "disp_0") else 1.25)


def _assert_close_up_to_sign_knife_edges(name, got, ref, margin, rtol=1.5 * GRAD_RTOL, edge=3e-5, max_frac=1e-3):
    """d|a - b| = sign(a - b): where the reference's own a - b is within a few ulp of 0 (``margin`` < ``edge``; coordinates are
    O(100) px, one ulp there is 8e-6) the sign -- an O(1) factor of that element's gradient -- is decided by rounding.  Every
    element beyond ``rtol`` must be such a knife edge, there must be few of them, and none may be off by more than the scale."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    assert bool(torch.isfinite(got).all()), name
    scale = max(float(ref.abs().max()), 1e-30)
    bad = ~((got - ref).abs() <= rtol * scale)
    if not bad.any():
        return
    assert int(bad.sum()) <= max(2, int(max_frac * got.numel())), (name, int(bad.sum()))
    assert float(margin.expand_as(bad)[bad].max()) < edge, (name, float(margin.expand_as(bad)[bad].max()))
    assert float((got - ref).abs().max()) <= 4.0 * scale, name


def test_geom_mode_vs_reference_golden_multitile(cuda_device):
    d = load_golden("geom_mode_s3_mt")
    t = golden_triplet(d)
    dev = cuda_device
    ff, fb = _leaf_list(t.flows_fwd, dev), _leaf_list(t.flows_bwd, dev)
    disp, disp_l, disp_r = _leaf_list(t.disp, dev), _leaf_list(t.disp_l, dev), _leaf_list(t.disp_r, dev)
    pose = t.pose.to(dev).requires_grad_(True)
    loss, masks = losses.GeometryLoss(3).forward_losses(t.img_l.to(dev), t.img.to(dev), t.img_r.to(dev), ff, fb, disp, disp_l, disp_r,
                                                        pose, t.K.to(dev), t.K_inv.to(dev))
    live = [k for k in loss if "out_" + k in d and d["out_" + k].numel() == t.img.shape[0] and loss[k].requires_grad]
    for k in live:
        assert loss_rel_err(loss[k], d["out_" + k].clamp_min(1e-30)) < (5e-5 if k == "loss_epipolar" else LOSS_RTOL) or float(d["out_" + k].abs().max()) == 0.0, k
    for key in ("occ_b", "occ_f", "valid_b", "valid_f", "dyn_b", "dyn_f", "tex_b", "tex_f", "val_l", "val_r"):
        for l in range(3):
            assert torch.equal(masks[key][l].cpu(), d["aux_%s_%d" % (key, l)]), (key, l)
    tot = sum(P.GEOM_WEIGHTS[k] * loss[k].mean() for k in live)
    names = (["flows_fwd_%d" % i for i in range(4)] + ["flows_bwd_%d" % i for i in range(4)] + ["disp_%d" % i for i in range(3)]
             + ["disp_l_%d" % i for i in range(3)] + ["disp_r_%d" % i for i in range(3)])
    grads = torch.autograd.grad(tot, ff + fb + disp + disp_l + disp_r, allow_unused=True)
    # level 0 carries the depth-flow consistency term |rigid flow - flow| (model_geometry.py:716-732): its sign knife edges
    with torch.no_grad():
        m_b = (P.rigid_flow(t.disp[0], t.pose[:, 0], t.K) - t.flows_bwd[0]).abs()
        m_f = (P.rigid_flow(t.disp[0], t.pose[:, 1], t.K) - t.flows_fwd[0]).abs()
    margins = {"flows_bwd_0": m_b, "flows_fwd_0": m_f,
               "disp_0": torch.minimum(m_b.min(1, keepdim=True).values, m_f.min(1, keepdim=True).values)}
    for n, g in zip(names, grads):
        ref = d["grad_" + n]
        g = torch.zeros_like(ref) if g is None else g.cpu()
        if float(ref.abs().max()) == 0.0:
            assert float(g.abs().max()) == 0.0, n
        elif n in margins:
            _assert_close_up_to_sign_knife_edges(n, g, ref, margins[n])
        else:
            assert_grad_close(n, g, ref, None, rtol=1.5 * GRAD_RTOL)


# ---- geom-mode pose gradients, term by term ----------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W", [(2, 64, 208), (1, 256, 832)])
@pytest.mark.parametrize("term", ["loss_depth_pixel", "loss_depth_flow_consis", "loss_epipolar"])
def test_geom_pose_gradient_per_term(cuda_device, term, B, H, W):
    """d term / d pose (B,2,6) of the three geom-mode terms that reach the pose, each on its own.  The criterion is the
    north-star 1e-4 against the fp32 oracle, or -- where the oracle's own fp32 arithmetic is noisier than that against the
    same oracle in fp64 (sums over 2e5 pixels of signed, cancelling contributions) -- being at most 2x as far from the fp64
    value as the fp32 oracle is.  A wrong analytic backward (pose set-up, projection chain) fails either."""
    dev = cuda_device
    t = make_triplet(B, H, W, 4, 3, seed=43, flow_mode="rigid")

    def oracle(dtype):
        cv = lambda x: x.detach().to(dtype)
        pose = cv(t.pose).requires_grad_(True)
        loss = P.geom_mode_loss(cv(t.img_l), cv(t.img), cv(t.img_r), [cv(f) for f in t.flows_fwd], [cv(f) for f in t.flows_bwd],
                                [cv(x) for x in t.disp], [cv(x) for x in t.disp_l], [cv(x) for x in t.disp_r], pose, cv(t.K), cv(t.K_inv), 3)
        g, = torch.autograd.grad(loss[term].sum(), [pose])
        return g

    r32, r64 = oracle(torch.float32), oracle(torch.float64)
    pose = t.pose.detach().to(dev).requires_grad_(True)
    mv = lambda xs: [x.detach().to(dev) for x in xs]
    loss, _ = losses.GeometryLoss(3).forward_losses(t.img_l.to(dev), t.img.to(dev), t.img_r.to(dev), mv(t.flows_fwd), mv(t.flows_bwd),
                                                    mv(t.disp), mv(t.disp_l), mv(t.disp_r), pose, t.K.to(dev), t.K_inv.to(dev))
    g, = torch.autograd.grad(loss[term].sum(), [pose])
    assert torch.isfinite(g).all()
    e32, e_got, e_ref = rel_err(g, r32), rel_err(g, r64), rel_err(r32, r64)
    assert e32 < GRAD_RTOL or e_got <= 2.0 * e_ref + 1e-5, (term, e32, e_got, e_ref)
    assert e_got < 5e-2                                              # and never beyond the worst floor ever measured


def test_pose_setup_backward_vs_torch_fp64(cuda_device):
    """ugl_pose_setup_forward / _backward (euler2mat, K_s, K_s^-1, K_s [R|t], F = K^-T [t]x R K^-1) against the composed torch chain
    of structures.projection_pyramid evaluated in fp64 (inverse_warp.py:110-187, model_geometry.py:92-93, 284-289, 354-364)."""
    dev = cuda_device
    g = torch.Generator().manual_seed(7)
    B = 3
    pose = (0.05 * torch.randn(B, 2, 6, generator=g))
    K = torch.tensor([[[0.58 * 832, 0.0, 416.0], [0.0, 1.92 * 256, 128.0], [0.0, 0.0, 1.0]]]).repeat(B, 1, 1)
    K_inv = torch.linalg.inv(K)
    downs = [1.0, 2.0, 4.0]
    pd = pose.to(dev).requires_grad_(True)
    Kinv, (P_b, P_f), Fm = ops.pose_setup(pd, K.to(dev), downs, K_inv.to(dev), fundamental=True)
    outs = list(P_b) + list(P_f) + list(Fm)
    gos = [torch.randn(o.shape, generator=g) for o in outs]
    got, = torch.autograd.grad(sum((o * go.to(dev)).sum() for o, go in zip(outs, gos)), [pd])

    p64 = pose.double().requires_grad_(True)
    K64 = K.double()
    ref_outs = []
    for idx in (0, 1):
        T = P.pose_to_matrix(p64[:, idx])                                                # (B,3,4)
        for s in downs:
            ref_outs.append((idx, P.scale_intrinsics(K64, s).bmm(T)))
    ref_P = [o for i, o in ref_outs if i == 0] + [o for i, o in ref_outs if i == 1]
    ref_F = []
    for idx in (0, 1):
        E = P.essential_matrix(p64[:, idx])
        ref_F.append(K_inv.double().transpose(1, 2).bmm(E.bmm(K_inv.double())))
    refs = ref_P + ref_F
    for o, r in zip(outs, refs):
        assert rel_err(o, r) < 2e-6
    want, = torch.autograd.grad(sum((r * go.double()).sum() for r, go in zip(refs, gos)), [p64])
    assert rel_err(got, want) < 2e-5


# ---- M7: get_rigid_mask -------------------------------------------------------------------------------------------------
def test_rigid_mask_bit_exact_vs_oracle(cuda_device):
    """model_geometry.py:420-425: [dist < 0.5], [dist < 0.1], rigid / (1 + dist) -- including values exactly at and next to the
    thresholds."""
    g = torch.Generator().manual_seed(11)
    dist = (torch.rand(2, 1, 37, 53, generator=g) * 0.8)
    edge = torch.tensor([0.5, 0.1, 0.0])
    special = torch.cat([edge, torch.nextafter(edge, torch.tensor(10.0)), torch.nextafter(edge, torch.tensor(-10.0)).clamp_min(0.0)])
    dist.view(-1)[:special.numel()] = special
    rigid, inlier, score = ops.rigid_mask(dist.to(cuda_device))
    r_ref, i_ref, s_ref = P.rigid_masks(dist)
    assert torch.equal(rigid.cpu(), r_ref) and torch.equal(inlier.cpu(), i_ref)
    assert torch.equal(score.cpu(), s_ref) or rel_err(score, s_ref) < 1e-7
    rigid2, inlier2, _ = ops.rigid_mask(dist.to(cuda_device), 0.3, 0.05)
    r2, i2, _ = P.rigid_masks(dist, 0.3, 0.05)
    assert torch.equal(rigid2.cpu(), r2) and torch.equal(inlier2.cpu(), i2)


def test_geom_masks_lazy_entries_equal_composed(cuda_device):
    """GeomMasks of the fused path: the lazily computed dist_b / dist_f / rigid_f / inlier_f and the unpacked fwd / bwd masks equal
    the composed path's (one kernel per reference method), and the oracle's."""
    dev = cuda_device
    t = make_triplet(2, 64, 208, 4, 3, seed=47, flow_mode="rigid")
    td = t.to(dev)
    mod = losses.GeometryLoss(3)
    args = (td.img_l, td.img, td.img_r, td.flows_fwd, td.flows_bwd, td.disp, td.disp_l, td.disp_r, td.pose, td.K, td.K_inv)
    _, mf = mod.forward_losses(*args, fused=True)
    _, mc = mod.forward_losses(*args, fused=False)
    assert rel_err(mf["dist_f"], mc["dist_f"]) < 1e-6 and rel_err(mf["dist_b"], mc["dist_b"]) < 1e-6
    dist_o = P.epipolar_distance(t.pose[:, 1], t.flows_fwd[0], t.K, t.K_inv)
    r_o, i_o, _ = P.rigid_masks(dist_o)
    for key, ref in (("rigid_f", r_o), ("inlier_f", i_o)):
        a, b = mf[key].cpu(), mc[key].cpu()
        assert int((a != b).sum()) <= 2 and int((a != ref).sum()) <= 4, key      # thresholds on a cancellation-prone fp32 distance
    for key in ("fwd_mask", "bwd_mask"):
        for l in range(3):
            assert torch.equal(mf[key][l], mc[key][l]), (key, l)


# ---- internal forms of the single-pass forward, and the fused training step -------------------------------------------
@pytest.mark.parametrize("B,H,W,L", [(2, 64, 208, 3), (2, 96, 160, 4), (1, 36, 52, 2), (1, 256, 832, 4)])
def test_single_pass_variants_agree(cuda_device, B, H, W, L):
    """fused tile kernel / split kernels with plain staging / split kernels with TMA staging: same losses (1e-6) and gradients
    (2e-6; the split form applies the occlusion weight before instead of after the coefficient box sums)."""
    t = make_triplet(B, H, W, L, 1, seed=81, flow_px=5.0, oob_fraction=0.05).to(cuda_device)
    pl, pc, pr = (ops.image_pyramid(x, L, "box") for x in (t.img_l, t.img, t.img_r))
    gl = (torch.rand(4, B, generator=torch.Generator().manual_seed(2)) + 0.5).to(cuda_device)
    res = {}
    tma_ok = all((W >> l) % 4 == 0 for l in range(L))
    for variant in ("fused", "split_plain", "split", "split_tma"):
        if variant == "split_tma" and not tma_ok:
            ops.SINGLE_PASS_VARIANT = variant
            try:
                with pytest.raises(Exception, match="TMA staging requested"):
                    ops.flow_loss(pl, pc, pr, [f.clone().requires_grad_(True) for f in t.flows_fwd], t.flows_bwd, L, as_matrix=True)
            finally:
                ops.SINGLE_PASS_VARIANT = "split"
            continue
        ops.SINGLE_PASS_VARIANT = variant
        try:
            ff = [f.detach().clone().requires_grad_(True) for f in t.flows_fwd]
            fb = [f.detach().clone().requires_grad_(True) for f in t.flows_bwd]
            loss = ops.flow_loss(pl, pc, pr, ff, fb, L, as_matrix=True)
            res[variant] = (loss, torch.autograd.grad(loss, ff + fb, grad_outputs=gl))
        finally:
            ops.SINGLE_PASS_VARIANT = "split"
    base = res["fused"]
    for variant, (loss, g) in res.items():
        assert loss_rel_err(loss, base[0]) < 1e-6, variant
        for a, b in zip(g, base[1]):
            assert rel_err(a, b) < 2e-6, variant
    if "split_tma" in res:       # the staging form must not change a single bit
        assert torch.equal(res["split_tma"][0], res["split_plain"][0])
        assert all(torch.equal(a, b) for a, b in zip(res["split_tma"][1], res["split_plain"][1]))


def test_geom_single_pass_variants_agree(cuda_device):
    t = make_triplet(2, 64, 208, 4, 3, seed=83, flow_mode="rigid").to(cuda_device)
    mod = losses.GeometryLoss(3)
    res = {}
    for variant in ("fused", "split_plain", "split"):
        ops.SINGLE_PASS_VARIANT = variant
        try:
            ff, fb = _leaf_list(t.flows_fwd, cuda_device), _leaf_list(t.flows_bwd, cuda_device)
            loss, masks = mod.forward_losses(t.img_l, t.img, t.img_r, ff, fb, t.disp, t.disp_l, t.disp_r, t.pose, t.K, t.K_inv)
            keys = ["loss_flow_pixel", "loss_flow_ssim", "loss_flow_smooth", "loss_flow_consis"]
            g = torch.autograd.grad(sum(P.GEOM_WEIGHTS[k] * loss[k].mean() for k in keys), ff[:3] + fb[:3])
            res[variant] = ({k: loss[k] for k in keys}, [m.clone() for m in masks.mask_bytes], g)
        finally:
            ops.SINGLE_PASS_VARIANT = "split"
    for variant in ("split_plain", "split"):
        for k, v in res[variant][0].items():
            assert loss_rel_err(v, res["fused"][0][k]) < 1e-6, (variant, k)
        for a, b in zip(res[variant][1], res["fused"][1]):
            assert torch.equal(a, b), variant                                    # all six packed masks bit-identical
        for a, b in zip(res[variant][2], res["fused"][2]):
            assert rel_err(a, b) < 2e-6, variant


@pytest.mark.parametrize("B,H,W,L,scales", [(2, 64, 208, 4, 4), (1, 112, 168, 4, 3), (3, 39, 57, 1, 1), (1, 256, 832, 4, 4)])
def test_fused_step_vs_oracle_and_autograd(cuda_device, B, H, W, L, scales):
    """ugl_flow_loss_step (losses + flow gradients in one pass, no basis planes) against the oracle and against the autograd path."""
    dev = cuda_device
    even = (H % (1 << (L - 1)) == 0) and (W % (1 << (L - 1)) == 0)
    t = make_triplet(B, H, W, L, 1, seed=85, flow_px=4.0, oob_fraction=0.03)
    gl = torch.rand(4, B, generator=torch.Generator().manual_seed(4)) + 0.5
    td = t.to(dev)
    if even:
        pl, pc, pr = (ops.image_pyramid(x, L, "box") for x in (td.img_l, td.img, td.img_r))
    else:
        pl, pc, pr = [td.img_l], [td.img], [td.img_r]
    st = ops.flow_loss_step(pl, pc, pr, td.flows_fwd, td.flows_bwd, gl.to(dev), scales)
    ff = [f.detach().clone().requires_grad_(True) for f in t.flows_fwd]
    fb = [f.detach().clone().requires_grad_(True) for f in t.flows_bwd]
    ref = P.flow_mode_loss(t.img_l, t.img, t.img_r, ff, fb, scales)
    rg = torch.autograd.grad(sum((gl[k] * ref[KEYS[k]]).sum() for k in range(4)), ff[:scales] + fb[:scales])
    f64 = lambda xs: [x.detach().double().requires_grad_(True) for x in xs]
    ff64, fb64 = f64(t.flows_fwd), f64(t.flows_bwd)
    ref64 = P.flow_mode_loss(t.img_l.double(), t.img.double(), t.img_r.double(), ff64, fb64, scales)
    rg64 = torch.autograd.grad(sum((gl[k].double() * ref64[KEYS[k]]).sum() for k in range(4)), ff64[:scales] + fb64[:scales])
    for k in range(4):
        assert loss_rel_err(st["loss"][k], ref[KEYS[k]]) < LOSS_RTOL, KEYS[k]
    for a, b, c in zip(st["gf"] + st["gb"], rg, rg64):
        assert torch.isfinite(a).all()
        assert rel_err(a, b) < GRAD_RTOL or rel_err(a, c) <= 1.25 * rel_err(b, c)
    a_ff = [f.detach().clone().requires_grad_(True) for f in td.flows_fwd]
    a_fb = [f.detach().clone().requires_grad_(True) for f in td.flows_bwd]
    al = ops.flow_loss(pl, pc, pr, a_ff, a_fb, scales, as_matrix=True)
    ag = torch.autograd.grad(al, a_ff[:scales] + a_fb[:scales], grad_outputs=gl.to(dev))
    assert loss_rel_err(st["loss"], al) < 1e-6
    for a, b in zip(st["gf"] + st["gb"], ag):
        assert rel_err(a, b) < 2e-6


# ---- one autograd node per mode (mode_steps.py) against the same kernels as separate autograd Functions ----------------------
def test_geom_step_node_equals_op_by_op(cuda_device):
    """losses.GeometryLoss(fused=True) runs the fused kernels inside ONE autograd node (gradient accumulation and the weighted total
    as single launches); fused='ops' runs the same kernels as separate Functions under the autograd engine.  Same kernels, same
    inputs: losses and masks bit-identical, gradients equal up to the order of the additions."""
    t = make_triplet(2, 64, 208, 4, 3, seed=91, flow_mode="rigid").to(cuda_device)
    mod = losses.GeometryLoss(3)
    res = {}
    for fused in (True, "ops"):
        ff, fb = _leaf_list(t.flows_fwd, cuda_device), _leaf_list(t.flows_bwd, cuda_device)
        disp, disp_l, disp_r = _leaf_list(t.disp, cuda_device), _leaf_list(t.disp_l, cuda_device), _leaf_list(t.disp_r, cuda_device)
        pose = t.pose.detach().clone().requires_grad_(True)
        loss, masks = mod.forward_losses(t.img_l, t.img, t.img_r, ff, fb, disp, disp_l, disp_r, pose, t.K, t.K_inv, fused=fused)
        total = losses.total_loss(loss, P.GEOM_WEIGHTS)
        g = torch.autograd.grad(total, ff[:3] + fb[:3] + disp + disp_l + disp_r + [pose])
        res[fused] = (loss, masks, total, g)
    a, b = res[True], res["ops"]
    assert set(a[0]) == set(b[0])
    for k in b[0]:
        assert a[0][k].shape == b[0][k].shape and torch.equal(a[0][k].detach(), b[0][k].detach()), k
    for key in ("occ_b", "occ_f", "valid_b", "valid_f", "dyn_b", "dyn_f", "tex_b", "tex_f", "val_l", "val_r", "fwd_mask", "bwd_mask"):
        for l in range(3):
            assert torch.equal(a[1][key][l], b[1][key][l]), (key, l)
    assert rel_err(a[1]["dist_f"], b[1]["dist_f"]) == 0.0
    assert abs(float(a[2].detach()) - float(b[2].detach())) <= 1e-6 * abs(float(b[2].detach()))
    for x, y in zip(a[3], b[3]):
        assert rel_err(x, y) < 1e-6
    assert ff[3].grad is None      # level 3 of the flow pyramids carries no loss in geom mode (model_geometry.py loops over num_scales)


@pytest.mark.parametrize("variant", ["live", "ssim", "texture"])
def test_depth_step_node_equals_op_by_op(cuda_device, variant):
    t = make_triplet(2, 64, 208, 4, 3, seed=93, flow_mode="rigid").to(cuda_device)
    mod = losses.DepthLoss(3, variant)
    res = {}
    for fused in (True, "ops"):
        disp, disp_l, disp_r = _leaf_list(t.disp, cuda_device), _leaf_list(t.disp_l, cuda_device), _leaf_list(t.disp_r, cuda_device)
        pose = t.pose.detach().clone().requires_grad_(True)
        loss, masks = mod.forward_losses(t.img_l, t.img, t.img_r, disp, disp_l, disp_r, pose, t.K, fused=fused)
        total = losses.total_loss(loss, P.GEOM_WEIGHTS)
        g = torch.autograd.grad(total, disp + disp_l + disp_r + [pose], allow_unused=True)
        res[fused] = (loss, masks, total, g)
    a, b = res[True], res["ops"]
    assert set(a[0]) == set(b[0])
    for k in b[0]:
        assert a[0][k].shape == b[0][k].shape and torch.equal(a[0][k].detach(), b[0][k].detach()), k
    for key in ("valid_l", "valid_r", "tex_b", "tex_f"):
        for l in range(3):
            assert torch.equal(a[1][key][l], b[1][key][l]), (key, l)
    assert abs(float(a[2].detach()) - float(b[2].detach())) <= 1e-6 * abs(float(b[2].detach()))
    for x, y in zip(a[3], b[3]):
        assert (x is None) == (y is None)
        if x is not None:
            assert rel_err(x, y) < 1e-6


# ---- programmatic dependent launches / side streams: same bits, no stale data ---------------------------------------------------
def test_fused_step_graph_replay_tracks_inputs(cuda_device):
    """The fused step's kernels are chained by programmatic dependent launches (the stencil kernel starts under the weight-sum kernel,
    the finalize under the stencil kernel's tail).  One captured graph replayed on inputs changed in place must give, every time,
    exactly what a fresh eager call gives on the same inputs: nothing read before its producer finished, nothing stale in a cache."""
    dev = cuda_device
    B, H, W, L = 2, 128, 416, 4
    t = make_triplet(B, H, W, L, 1, seed=97, flow_px=6.0).to(dev)
    pl, pc, pr = (ops.image_pyramid(x, L, "box") for x in (t.img_l, t.img, t.img_r))
    gl = (torch.rand(4, B, generator=torch.Generator().manual_seed(5)) + 0.5).to(dev)
    out = ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L)
    g, s = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(s):
        ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L, out=out)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L, out=out)
    torch.cuda.synchronize()
    for k in range(4):
        for f in t.flows_fwd + t.flows_bwd:
            f.mul_(0.8)
        for x in pc + pl:
            x.mul_(0.97)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        got = [out["loss"].clone()] + [x.clone() for x in out["gf"] + out["gb"]]
        fresh = ops.flow_loss_step(pl, pc, pr, t.flows_fwd, t.flows_bwd, gl, L)
        torch.cuda.synchronize()
        for a, b in zip(got, [fresh["loss"]] + fresh["gf"] + fresh["gb"]):
            assert torch.equal(a, b), k


def test_mode_steps_side_streams_same_bits(cuda_device):
    """mode_steps runs independent branches of a geom / depth step on side streams (fork / join by stream waits).  Same kernels, same
    inputs: every loss and gradient bit-identical with the single-stream order."""
    from unsupervised_depth_opticalflow_egomotion_b200 import mode_steps
    t = make_triplet(2, 64, 208, 4, 3, seed=99, flow_mode="rigid").to(cuda_device)
    res = {}
    saved = mode_steps._Side.enabled
    try:
        for enabled in (True, False):
            mode_steps._Side.enabled = enabled
            ff, fb = _leaf_list(t.flows_fwd, cuda_device), _leaf_list(t.flows_bwd, cuda_device)
            disp, disp_l, disp_r = _leaf_list(t.disp, cuda_device), _leaf_list(t.disp_l, cuda_device), _leaf_list(t.disp_r, cuda_device)
            pose = t.pose.detach().clone().requires_grad_(True)
            loss, _ = losses.GeometryLoss(3).forward_losses(t.img_l, t.img, t.img_r, ff, fb, disp, disp_l, disp_r, pose, t.K, t.K_inv)
            total = losses.total_loss(loss, P.GEOM_WEIGHTS)
            g = torch.autograd.grad(total, ff[:3] + fb[:3] + disp + disp_l + disp_r + [pose])
            d2, d2l, d2r = _leaf_list(t.disp, cuda_device), _leaf_list(t.disp_l, cuda_device), _leaf_list(t.disp_r, cuda_device)
            pose2 = t.pose.detach().clone().requires_grad_(True)
            dl, _ = losses.DepthLoss(3, "texture").forward_losses(t.img_l, t.img, t.img_r, d2, d2l, d2r, pose2, t.K)
            dg = torch.autograd.grad(losses.total_loss(dl, P.GEOM_WEIGHTS), d2 + d2l + d2r + [pose2], allow_unused=True)
            torch.cuda.synchronize()
            res[enabled] = ([total.detach()] + [x.detach() for x in g], [dl[k].detach() for k in sorted(dl) if dl[k].numel() > 2] + [x for x in dg if x is not None])
    finally:
        mode_steps._Side.enabled = saved
    for a, b in zip(res[True][0] + res[True][1], res[False][0] + res[False][1]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("B,H,W", [(2, 112, 208), (1, 76, 116)])     # second size: level 1 without TMA (width % 4 != 0), odd width at level 2
def test_geom_step_weights_form_equals_autograd_form(cuda_device, B, H, W):
    """GeometryLoss.forward_losses(step_weights=w): the flow branch runs as a fused training step (ugl_geom_flow_step: gradients written by
    the forward launches, no basis planes, no combine).  Same losses and masks bit for bit, same gradients up to the association of the
    four-term sum; the weighted total must then be formed with the same weights."""
    t = make_triplet(B, H, W, 4, 3, seed=101, flow_mode="rigid").to(cuda_device)
    mod = losses.GeometryLoss(3)
    res = {}
    for step in (False, True):
        ff, fb = _leaf_list(t.flows_fwd, cuda_device), _leaf_list(t.flows_bwd, cuda_device)
        disp, disp_l, disp_r = _leaf_list(t.disp, cuda_device), _leaf_list(t.disp_l, cuda_device), _leaf_list(t.disp_r, cuda_device)
        pose = t.pose.detach().clone().requires_grad_(True)
        loss, masks = mod.forward_losses(t.img_l, t.img, t.img_r, ff, fb, disp, disp_l, disp_r, pose, t.K, t.K_inv,
                                         step_weights=P.GEOM_WEIGHTS if step else None)
        total = losses.total_loss(loss, P.GEOM_WEIGHTS)
        g = torch.autograd.grad(total, ff[:3] + fb[:3] + disp + disp_l + disp_r + [pose])
        res[step] = (loss, masks, total, g)
        if step:
            other = dict(P.GEOM_WEIGHTS); other["loss_flow_ssim"] = 0.5
            with pytest.raises(ValueError):
                losses.total_loss(loss, other)
    a, b = res[True], res[False]
    for k in b[0]:
        assert torch.equal(a[0][k].detach(), b[0][k].detach()), k
    for key in ("occ_b", "occ_f", "valid_b", "valid_f", "dyn_b", "dyn_f", "tex_b", "tex_f"):
        for l in range(3):
            assert torch.equal(a[1][key][l], b[1][key][l]), (key, l)
    assert torch.equal(a[2].detach(), b[2].detach())
    for x, y in zip(a[3], b[3]):
        assert torch.isfinite(x).all() and rel_err(x, y) < 2e-6
