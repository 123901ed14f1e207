"""GPU parity tests of the fused flow-mode loss (through the C-ABI) against the oracle, the golden
fixtures generated from the reference, and size-independent properties at the BASELINE size."""
import pytest
import torch

from oracle import loss_port as P
from unsupervised_depth_opticalflow_egomotion_b200 import ops
from unsupervised_depth_opticalflow_egomotion_b200.synth import make_triplet
from util import load_golden, golden_triplet, rel_err, loss_rel_err, LOSS_RTOL, GRAD_RTOL

pytestmark = pytest.mark.gpu
KEYS = list(ops.FLOW_LOSS_KEYS)


def _cuda_flow(t, scales, gl, dev, mode="single_pass"):
    L = len(t.flows_fwd)
    t = t.to(dev)
    pl, pc, pr = (ops.image_pyramid(x, L, "box") for x in (t.img_l, t.img, t.img_r))
    ff = [f.detach().clone().requires_grad_(True) for f in t.flows_fwd]
    fb = [f.detach().clone().requires_grad_(True) for f in t.flows_bwd]
    loss = ops.flow_loss(pl, pc, pr, ff, fb, scales, as_matrix=True, mode=mode)
    g = torch.autograd.grad(loss, ff[:scales] + fb[:scales], grad_outputs=gl.to(dev))
    return loss.cpu(), [x.cpu() for x in g[:scales]], [x.cpu() for x in g[scales:]]


def _oracle_flow(t, scales, gl, dtype=torch.float32):
    cv = lambda x: x.detach().to(dtype)
    ff = [cv(f).requires_grad_(True) for f in t.flows_fwd]
    fb = [cv(f).requires_grad_(True) for f in t.flows_bwd]
    loss = P.flow_mode_loss(cv(t.img_l), cv(t.img), cv(t.img_r), ff, fb, scales)
    tot = sum((gl[k].to(dtype) * loss[KEYS[k]]).sum() for k in range(4))
    g = torch.autograd.grad(tot, ff[:scales] + fb[:scales])
    return loss, g[:scales], g[scales:]


def _assert_grad(name, cuda_g, ref32, ref64):
    """1e-4 of max|g| against the fp32 oracle; where the oracle's own fp32 rounding noise (measured against
    the same oracle in fp64) is that large — SSIM's E[x^2]-mu^2 cancellation — require being at least as
    close to the fp64 value as the fp32 oracle is."""
    e32 = rel_err(cuda_g, ref32)
    if e32 < GRAD_RTOL:
        return
    assert rel_err(cuda_g, ref64) <= 1.25 * rel_err(ref32, ref64), (name, e32)


@pytest.mark.parametrize("B,Hh,W,L,scales,px,oob,mode", [
    (2, 64, 208, 4, 4, 6.0, 0.0, "noise"), (1, 48, 80, 4, 4, 3.0, 0.3, "noise"), (2, 40, 72, 4, 3, 1.0, 0.0, "noise"),
    (1, 34, 50, 2, 2, 2.0, 0.1, "noise"), (2, 128, 416, 4, 4, 12.0, 0.0, "noise"), (1, 64, 208, 4, 4, 0.0, 0.0, "noise")])
def test_flow_loss_vs_oracle(cuda_device, B, Hh, W, L, scales, px, oob, mode):
    t = make_triplet(B, Hh, W, L, 1, seed=21, flow_px=px, oob_fraction=oob, flow_mode=mode)
    gl = torch.rand(4, B, generator=torch.Generator().manual_seed(1)) + 0.5
    ref, rf, rb = _oracle_flow(t, scales, gl)
    _, rf64, rb64 = _oracle_flow(t, scales, gl, torch.float64)
    for kernel_mode in ("single_pass", "recompute"):
        loss, gf, gb = _cuda_flow(t, scales, gl, cuda_device, kernel_mode)
        for k in range(4):
            assert loss_rel_err(loss[k], ref[KEYS[k]]) < LOSS_RTOL, (kernel_mode, KEYS[k])
        for l in range(scales):
            _assert_grad("%s fwd%d" % (kernel_mode, l), gf[l], rf[l], rf64[l])
            _assert_grad("%s bwd%d" % (kernel_mode, l), gb[l], rb[l], rb64[l])


@pytest.mark.parametrize("name,scales", [("flow_mode_s4", 4), ("flow_mode_s4_oob", 4), ("flow_mode_s3", 3), ("flow_mode_s4_mt", 4)])
def test_flow_loss_vs_reference_golden(cuda_device, name, scales):
    d = load_golden(name)
    t = golden_triplet(d)
    B = t.img.shape[0]
    w = (torch.tensor([P.FLOW_WEIGHTS[k] for k in KEYS]).view(4, 1).repeat(1, B) / B).contiguous()
    loss, gf, gb = _cuda_flow(t, scales, w, cuda_device)
    for k in range(4):
        assert loss_rel_err(loss[k], d["out_" + KEYS[k]]) < LOSS_RTOL, KEYS[k]
    for l in range(scales):
        assert rel_err(gf[l], d["grad_flows_fwd_%d" % l]) < GRAD_RTOL
        assert rel_err(gb[l], d["grad_flows_bwd_%d" % l]) < 1.5 * GRAD_RTOL


def test_full_size_b1_vs_oracle(cuda_device):
    """BASELINE config 1 shape: 256x832, 4 levels, batch 1 (the oracle needs ~1 s for it)."""
    t = make_triplet(1, 256, 832, 4, 1, seed=1234, flow_px=10.0)
    gl = torch.tensor([[0.15], [0.85], [10.0], [0.01]])
    loss, gf, gb = _cuda_flow(t, 4, gl, cuda_device)
    ref, rf, rb = _oracle_flow(t, 4, gl)
    _, rf64, rb64 = _oracle_flow(t, 4, gl, torch.float64)
    for k in range(4):
        assert loss_rel_err(loss[k], ref[KEYS[k]]) < LOSS_RTOL, KEYS[k]
    for l in range(4):
        _assert_grad("fwd%d" % l, gf[l], rf[l], rf64[l])
        _assert_grad("bwd%d" % l, gb[l], rb[l], rb64[l])


def test_high_res_b1_vs_oracle(cuda_device):
    """BASELINE config 5 resolution (384x1280), batch 1, 4 levels."""
    t = make_triplet(1, 384, 1280, 4, 1, seed=4321, flow_px=12.0, oob_fraction=0.02)
    gl = torch.tensor([[0.15], [0.85], [10.0], [0.01]])
    loss, gf, gb = _cuda_flow(t, 4, gl, cuda_device)
    ref, rf, rb = _oracle_flow(t, 4, gl)
    _, rf64, rb64 = _oracle_flow(t, 4, gl, torch.float64)
    for k in range(4):
        assert loss_rel_err(loss[k], ref[KEYS[k]]) < LOSS_RTOL, KEYS[k]
    for l in range(4):
        _assert_grad("fwd%d" % l, gf[l], rf[l], rf64[l])
        _assert_grad("bwd%d" % l, gb[l], rb[l], rb64[l])


def test_full_size_properties(cuda_device):
    """BASELINE config 2 shape (256x832, batch 8): determinism, batch-shard invariance, linearity of the
    backward pass in the upstream gradient, zero gradient for unused levels."""
    B = 8
    t = make_triplet(B, 256, 832, 4, 1, seed=77, flow_px=10.0)
    g1 = torch.rand(4, B, generator=torch.Generator().manual_seed(2)) + 0.1
    g2 = torch.rand(4, B, generator=torch.Generator().manual_seed(3)) + 0.1
    la, fa, ba = _cuda_flow(t, 4, g1, cuda_device)
    lb, fb_, bb = _cuda_flow(t, 4, g1, cuda_device)
    assert torch.equal(la, lb) and all(torch.equal(x, y) for x, y in zip(fa + ba, fb_ + bb))      # bit-reproducible
    assert torch.isfinite(la).all() and all(torch.isfinite(x).all() for x in fa + ba)
    # every reduction is per sample: a shard of the batch gives the same per-sample numbers, bit for bit
    sub = make_triplet(B, 256, 832, 4, 1, seed=77, flow_px=10.0)
    pick = lambda x: x[2:5].contiguous()
    sub.img_l, sub.img, sub.img_r = pick(sub.img_l), pick(sub.img), pick(sub.img_r)
    sub.flows_fwd, sub.flows_bwd = [pick(f) for f in sub.flows_fwd], [pick(f) for f in sub.flows_bwd]
    ls, fs, bs = _cuda_flow(sub, 4, g1[:, 2:5].contiguous(), cuda_device)
    assert torch.equal(ls, la[:, 2:5])
    assert all(torch.equal(x, y[2:5]) for x, y in zip(fs + bs, fa + ba))
    # backward is linear in grad_loss
    l2, f2, b2 = _cuda_flow(t, 4, g2, cuda_device)
    l3, f3, b3 = _cuda_flow(t, 4, g1 + g2, cuda_device)
    for x, y, z in zip(fa + ba, f2 + b2, f3 + b3):
        assert rel_err(x + y, z) < 1e-5


def test_single_pass_and_recompute_agree_at_full_size(cuda_device):
    t = make_triplet(4, 256, 832, 4, 1, seed=99, flow_px=10.0, oob_fraction=0.05)
    gl = torch.rand(4, 4, generator=torch.Generator().manual_seed(5)) + 0.1
    la, fa, ba = _cuda_flow(t, 4, gl, cuda_device, "single_pass")
    lb, fb_, bb = _cuda_flow(t, 4, gl, cuda_device, "recompute")
    assert loss_rel_err(la, lb) < 1e-6
    for x, y in zip(fa + ba, fb_ + bb):
        assert rel_err(x, y) < 2e-5


def test_unused_levels_get_no_gradient(cuda_device):
    t = make_triplet(1, 64, 208, 4, 1, seed=5).to(cuda_device)
    pl, pc, pr = (ops.image_pyramid(x, 4, "box") for x in (t.img_l, t.img, t.img_r))
    ff = [f.requires_grad_(True) for f in t.flows_fwd]
    fb = [f.requires_grad_(True) for f in t.flows_bwd]
    loss = ops.flow_loss(pl, pc, pr, ff, fb, 3)
    sum(v.sum() for v in loss.values()).backward()
    assert ff[3].grad is None and fb[3].grad is None and ff[2].grad is not None
    assert fb[0].grad.abs().max() > 0


def test_shape_errors(cuda_device):
    t = make_triplet(1, 32, 64, 2, 1, seed=5).to(cuda_device)
    pl, pc, pr = (ops.image_pyramid(x, 2, "box") for x in (t.img_l, t.img, t.img_r))
    with pytest.raises(ValueError):
        ops.flow_loss(pl, pc, pr, t.flows_fwd, [t.flows_bwd[0], t.flows_bwd[1][:, :, :-1]], 2)
    with pytest.raises(ValueError):
        ops.flow_loss(pl, pc, pr, t.flows_fwd, t.flows_bwd, 3)
