"""``torch.autograd.Function`` surface over the C-ABI kernels (``include/ugl.h``).

PyTorch is plumbing here: it owns device memory (caching allocator), the current stream and the
autograd graph.  Every function below launches hand-written sm_100a kernels from
``libugl_b200.so`` on ``torch.cuda.current_stream()``; there is no CPU path and no PyTorch
fallback — CPU tensors raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _cabi

Tensor = torch.Tensor

LAUNCH_COUNTER = {"n": 0}   # kernels launched through this module (bench.py reports it)


def _count(n: int) -> None:
    LAUNCH_COUNTER["n"] += n


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t: Tensor, name: str) -> Tensor:
    if not isinstance(t, Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: the B200 loss path has no CPU implementation" % (name, t.device))
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def _dev_u8(t: Tensor, name: str) -> Tensor:
    if not isinstance(t, Tensor) or not t.is_cuda or t.dtype != torch.uint8 or not t.is_contiguous():
        raise TypeError("%s must be a contiguous CUDA uint8 tensor" % name)
    return t


def _ptr(t) -> Optional[int]:
    return t.data_ptr() if isinstance(t, Tensor) else None


# ================================================================================================
# image pyramid (PY)
# ================================================================================================
def image_pyramid(img: Tensor, levels: int, mode: str) -> List[Tensor]:
    """``generate_img_pyramid`` — mode ``'box'``: model_flow.py:58-64 (adaptive average pooling, also
    the ``'area'`` resize of model_geometry.py:91); mode ``'bilinear'``: model_geometry.py:65-72.
    Level 0 is the input itself.  Not differentiable (images never require grad in the reference)."""
    img = _dev(img, "img").detach()
    B, Cc, H, W = img.shape
    if levels < 1 or levels > _cabi.MAX_LEVELS:
        raise ValueError("levels must be in [1, %d]" % _cabi.MAX_LEVELS)
    outs = [img] + [torch.empty((B, Cc, H >> l, W >> l), device=img.device, dtype=torch.float32) for l in range(1, levels)]
    if levels > 1:
        arr = (C.c_void_p * levels)(*[o.data_ptr() for o in outs])
        with torch.cuda.device_of(img):
            rc = _cabi.lib().ugl_image_pyramid(img.data_ptr(), B, Cc, H, W, levels, {"box": 0, "area": 0, "bilinear": 1}[mode],
                                               arr, _stream_ptr())
        _cabi.check(rc, "ugl_image_pyramid")
        _count(levels - 1)
    return outs


def image_pyramids(imgs: Sequence[Tensor], levels: int, modes: Sequence[str]):
    """The pyramids of up to three images in ONE launch.  ``modes[i]`` is a string or tuple of strings out of ``'box'`` /
    ``'area'`` / ``'bilinear'``; returns, per image, a dict ``mode -> [level 0 .. levels-1]`` (level 0 = the input).  Same values
    as :func:`image_pyramid` (which it falls back to for ``levels`` outside 2..4)."""
    imgs = [_dev(t, "img").detach() for t in imgs]
    modes = [(m,) if isinstance(m, str) else tuple(m) for m in modes]
    if not 1 <= len(imgs) <= _cabi.UglPyramidArgs.MAX_IMAGES or len(modes) != len(imgs):
        raise ValueError("image_pyramids: 1..3 images with one mode entry each")
    B, Cc, H, W = imgs[0].shape
    if any(tuple(t.shape) != (B, Cc, H, W) for t in imgs):
        raise ValueError("image_pyramids: images must share one shape")
    f = 1 << (levels - 1)
    if not 2 <= levels <= 4 or H % f or W % f or (H * W) % 4 or any(t.data_ptr() % 16 for t in imgs):
        return [{m: image_pyramid(t, levels, m) for m in ms} for t, ms in zip(imgs, modes)]
    a = _cabi.UglPyramidArgs()
    a.batch, a.channels, a.height, a.width, a.levels, a.images = B, Cc, H, W, levels, len(imgs)
    out = []
    for i, (t, ms) in enumerate(zip(imgs, modes)):
        a.img[i] = t.data_ptr()
        d = {}
        for m in ms:
            lv = [t] + [torch.empty((B, Cc, H >> l, W >> l), device=t.device, dtype=torch.float32) for l in range(1, levels)]
            slot = a.bil if m == "bilinear" else a.box
            if m not in ("box", "area", "bilinear"):
                raise ValueError("image_pyramids: unknown mode %r" % (m,))
            for l in range(1, levels):
                slot[i][l] = lv[l].data_ptr()
            d[m] = lv
        out.append(d)
    a.stream = torch.cuda.current_stream().cuda_stream
    with torch.cuda.device_of(imgs[0]):
        _cabi.check(_cabi.lib().ugl_image_pyramid_multi(C.byref(a)), "ugl_image_pyramid_multi")
    _count(1)
    return out


# ================================================================================================
# warp_flow (W1)
# ================================================================================================
class _WarpFlowFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, flow: Tensor, use_mask: bool):
        x, flow = _dev(x, "x"), _dev(flow, "flow")
        B, Cc, H, W = x.shape
        out = torch.empty_like(x)
        with torch.cuda.device_of(x):
            rc = _cabi.lib().ugl_warp_flow_forward(x.data_ptr(), flow.data_ptr(), B, Cc, H, W, int(use_mask), out.data_ptr(),
                                                   None, _stream_ptr())
        _cabi.check(rc, "ugl_warp_flow_forward")
        _count(1)
        ctx.save_for_backward(x, flow)
        ctx.use_mask = bool(use_mask)
        return out

    @staticmethod
    def backward(ctx, gout: Tensor):
        x, flow = ctx.saved_tensors
        B, Cc, H, W = x.shape
        gout = _dev(gout, "grad_output")
        need_x, need_f = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gflow = torch.empty_like(flow) if need_f else None
        gx = torch.empty_like(x) if need_x else None
        ws_bytes = int(_cabi.lib().ugl_warp_flow_backward_workspace_bytes(B, Cc, H, W, int(need_x)))
        ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.int64, device=x.device) if ws_bytes else None
        with torch.cuda.device_of(x):
            rc = _cabi.lib().ugl_warp_flow_backward_ex(x.data_ptr(), flow.data_ptr(), gout.data_ptr(), B, Cc, H, W,
                                                       int(ctx.use_mask), _ptr(gflow), _ptr(gx), _ptr(ws), ws_bytes,
                                                       _cabi.SCATTER_FORMS[SCATTER_FORM], _stream_ptr())
        _cabi.check(rc, "ugl_warp_flow_backward_ex")
        _count((1 if need_f else 0) + (4 if need_x else 0))
        return gx, gflow, None


def warp_flow(x: Tensor, flow: Tensor, use_mask: bool = False) -> Tensor:
    """Drop-in for ``warp_flow`` (structures/net_utils.py:16-54): same signature, same
    ``ValueError`` on a grid/flow shape mismatch (:35-36)."""
    B, Cc, H, W = x.size()
    if torch.Size((B, 2, H, W)) != flow.shape:
        raise ValueError("the shape of grid {0} is not equal to the shape of flow {1}.".format(
            torch.Size((B, 2, H, W)), flow.shape))
    return _WarpFlowFn.apply(x, flow, bool(use_mask))


# ================================================================================================
# PWC-Net cost volume (the caller of warp_flow inside the flow network)
# ================================================================================================
class _CostVolumeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f1: Tensor, f2: Tensor, d: int):
        f1, f2 = _dev(f1, "input1"), _dev(f2, "input2")
        B, Cc, H, W = f1.shape
        n = 2 * d + 1
        out = torch.empty((B, n * n, H, W), device=f1.device, dtype=torch.float32)
        with torch.cuda.device_of(f1):
            _call("ugl_cost_volume_forward", f1.data_ptr(), f2.data_ptr(), B, Cc, H, W, d, out.data_ptr(), _stream_ptr())
        ctx.save_for_backward(f1, f2)
        ctx.d = d
        return out

    @staticmethod
    def backward(ctx, g):
        f1, f2 = ctx.saved_tensors
        B, Cc, H, W = f1.shape
        g = _dev(g, "grad_out")
        g1 = torch.empty_like(f1) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(f2) if ctx.needs_input_grad[1] else None
        if g1 is not None or g2 is not None:
            with torch.cuda.device_of(f1):
                _call("ugl_cost_volume_backward", f1.data_ptr(), f2.data_ptr(), g.data_ptr(), B, Cc, H, W, ctx.d, _ptr(g1), _ptr(g2),
                      _stream_ptr(), launches=1)
        return g1, g2, None


def cost_volume(input1: Tensor, input2: Tensor, d: int = 4) -> Tensor:
    """Drop-in for ``PWC_tf.corr_naive(input1, input2, d=4)`` (structures/pwc_tf.py:97-106): ``(B,(2d+1)^2,H,W)`` channel-mean
    correlation of ``input1`` with the integer shifts of the zero-padded ``input2``; one kernel instead of 81 multiply / mean
    pairs and a cat.  Same ``AssertionError`` on a shape mismatch (:99).  Differentiable w.r.t. both inputs (deterministic)."""
    assert (input1.shape == input2.shape)
    if input1.dim() != 4:
        raise ValueError("cost_volume: inputs must be (B,C,H,W)")
    if not 1 <= int(d) <= 4:
        raise ValueError("cost_volume: d must be in [1, 4]")
    return _CostVolumeFn.apply(input1, input2, int(d))


# ================================================================================================
# fused flow-mode loss (T0 / flow)
# ================================================================================================
FLOW_LOSS_KEYS = ("loss_flow_pixel", "loss_flow_ssim", "loss_flow_smooth", "loss_flow_consis")

"""internal form of the single-pass forward (flow and geom modes): 'split' (photometry kernel + TMA-staged stencil kernel, the
default), 'split_plain' / 'split_tma' (force the staging form), 'fused' (the one-kernel form).  Same results; a test / profiling knob."""
SINGLE_PASS_VARIANT = "split"
SCATTER_FORM = "tile_local"
"""How warp_flow's grad_x and forward_splat accumulate (include/ugl.h UGL_SCATTER_*): ``tile_local`` (shared-memory window per CTA, each
global cell touched once) or ``global`` (one 64-bit global atomic per tap corner; the cross-check).  Same bits."""


def _variant() -> int:
    return _cabi.SINGLE_PASS_VARIANTS[SINGLE_PASS_VARIANT]



def _flow_args(img_l, img, img_r, ff, fb, scales, loss, stats, ws, gloss=None, gf=None, gb=None, basis=None) -> _cabi.UglFlowLossArgs:
    a = _cabi.UglFlowLossArgs()
    if basis is not None:
        for l in range(scales):
            a.basis[l] = basis[l].data_ptr()
    L = len(img)
    a.batch, a.levels, a.scales = img[0].shape[0], L, scales
    for l in range(L):
        a.height[l], a.width[l] = img[l].shape[2], img[l].shape[3]
        a.img_l[l], a.img[l], a.img_r[l] = img_l[l].data_ptr(), img[l].data_ptr(), img_r[l].data_ptr()
        a.flow_fwd[l], a.flow_bwd[l] = ff[l].data_ptr(), fb[l].data_ptr()
        if gf is not None and l < scales:
            a.grad_flow_fwd[l], a.grad_flow_bwd[l] = gf[l].data_ptr(), gb[l].data_ptr()
    a.loss, a.stats = _ptr(loss), _ptr(stats)
    a.grad_loss = _ptr(gloss)
    a.workspace, a.workspace_bytes = _ptr(ws), (ws.numel() * ws.element_size() if ws is not None else 0)
    a.stream = torch.cuda.current_stream().cuda_stream
    return a


def _alloc_basis(ff, scales):
    return [torch.empty((f.shape[0], _cabi.FLOW_BASIS_PLANES, f.shape[2], f.shape[3]), device=f.device, dtype=torch.float32)
            for f in ff[:scales]]


class _FlowLossFn(torch.autograd.Function):
    """mode 'single_pass' (default when a flow requires grad): the forward stencil kernel also writes the
    un-normalised gradient maps, backward is an element-wise combine.  mode 'recompute': nothing per-pixel is
    saved, the backward kernel recomputes the photometry (lower memory, ~1.4x the instructions)."""

    @staticmethod
    def forward(ctx, scales: int, L: int, mode: str, *ts: Tensor):
        ts = tuple(_dev(t, "input %d" % i) for i, t in enumerate(ts))
        img_l, img, img_r, ff, fb = (ts[k * L:(k + 1) * L] for k in range(5))
        B = img[0].shape[0]
        for l in range(L):
            h, w = img[l].shape[2], img[l].shape[3]
            for name, t, ch in (("img_l", img_l[l], 3), ("img", img[l], 3), ("img_r", img_r[l], 3), ("flow_fwd", ff[l], 2),
                                ("flow_bwd", fb[l], 2)):
                if tuple(t.shape) != (B, ch, h, w):
                    raise ValueError("flow_loss: %s[%d] has shape %s, expected %s" % (name, l, tuple(t.shape), (B, ch, h, w)))
        dev = img[0].device
        need_grad = any(ctx.needs_input_grad[3 + 3 * L:])
        single = need_grad and mode == "single_pass"
        loss = torch.empty((4, B), device=dev, dtype=torch.float32)
        stats = torch.empty((B, scales, _cabi.FLOW_NSTATS), device=dev, dtype=torch.float32)
        basis = _alloc_basis(ff, scales) if single else None
        a = _flow_args(img_l, img, img_r, ff, fb, scales, loss, stats, None, basis=basis)
        ws = torch.empty(max(int(_cabi.lib().ugl_flow_loss_workspace_bytes(C.byref(a))) // 4, 1), device=dev, dtype=torch.float32)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel() * 4
        with torch.cuda.device_of(img[0]):
            if single:
                _call("ugl_flow_loss_forward_grad_ex", C.byref(a), _variant(), launches=2 if SINGLE_PASS_VARIANT == "fused" else 3)
            else:
                _call("ugl_flow_loss_forward", C.byref(a), launches=2)
        if single:
            ctx.save_for_backward(stats, *ts, *basis)
        else:
            ctx.save_for_backward(stats, *ts)
        ctx.scales, ctx.L, ctx.single = scales, L, single
        return loss

    @staticmethod
    def backward(ctx, gloss: Tensor):
        stats, *rest = ctx.saved_tensors
        L, scales = ctx.L, ctx.scales
        ts, basis = rest[:5 * L], (rest[5 * L:] if ctx.single else None)
        img_l, img, img_r, ff, fb = (ts[k * L:(k + 1) * L] for k in range(5))
        gloss = _dev(gloss, "grad_loss")
        gf = [torch.empty_like(ff[l]) for l in range(scales)]
        gb = [torch.empty_like(fb[l]) for l in range(scales)]
        a = _flow_args(img_l, img, img_r, ff, fb, scales, None, stats, None, gloss, gf, gb, basis=basis)
        with torch.cuda.device_of(gloss):
            _call("ugl_flow_loss_combine" if ctx.single else "ugl_flow_loss_backward", C.byref(a))
        none_l = [None] * L
        pad = [None] * (L - scales)
        return (None, None, None, *none_l, *none_l, *none_l, *gf, *pad, *gb, *pad)


def flow_loss_step(img_l_pyr: Sequence[Tensor], img_pyr: Sequence[Tensor], img_r_pyr: Sequence[Tensor],
                   flows_fwd: Sequence[Tensor], flows_bwd: Sequence[Tensor], grad_loss: Tensor,
                   num_scales: Optional[int] = None, out: Optional[dict] = None, mode: str = "fused_step", phase: str = "both"):
    """Forward + backward of the fused flow-mode loss in one call, without the autograd engine: for
    the training step where the upstream gradient is known up front (``train.py:211-215``:
    ``d total / d loss_k[b] = w_k / B``).  Returns ``{'loss': (4,B), 'gf': [...], 'gb': [...], ...}``; pass the
    previous result as ``out`` to reuse its buffers (CUDA-graph friendly: no allocation).

    ``mode='fused_step'`` (default): ``ugl_flow_loss_step`` -- photometry kernel, weight sums, stencil kernel writing the
    flow gradients, finalize: 4 launches, nothing saved per pixel.  ``'single_pass'``: forward_grad + combine through the
    14 basis planes (what autograd uses, 4 launches).  ``'recompute'``: forward, finalize, recompute backward."""
    L = len(flows_fwd)
    scales = L if num_scales is None else int(num_scales)
    ts = [_dev(t, "input") for t in (*img_l_pyr[:L], *img_pyr[:L], *img_r_pyr[:L], *flows_fwd, *flows_bwd)]
    img_l, img, img_r, ff, fb = (ts[k * L:(k + 1) * L] for k in range(5))
    gloss = _dev(grad_loss, "grad_loss")
    B, dev = img[0].shape[0], img[0].device
    if out is None:
        out = {"loss": torch.empty((4, B), device=dev, dtype=torch.float32),
               "stats": torch.empty((B, scales, _cabi.FLOW_NSTATS), device=dev, dtype=torch.float32),
               "gf": [torch.empty_like(ff[l]) for l in range(scales)], "gb": [torch.empty_like(fb[l]) for l in range(scales)],
               "basis": _alloc_basis(ff, scales) if mode == "single_pass" else None, "mode": mode}
        a = _flow_args(img_l, img, img_r, ff, fb, scales, out["loss"], out["stats"], None)
        out["ws"] = torch.empty(max(int(_cabi.lib().ugl_flow_loss_workspace_bytes(C.byref(a))) // 4, 1), device=dev,
                                dtype=torch.float32)
    a = _flow_args(img_l, img, img_r, ff, fb, scales, out["loss"], out["stats"], out["ws"], gloss, out["gf"], out["gb"],
                   basis=out["basis"])
    fwd, bwd = phase in ("both", "forward"), phase in ("both", "backward")   # 'forward'/'backward': one half only (kernel timing)
    if out.get("mode", mode) != mode:
        raise ValueError("flow_loss_step: `out` was allocated for mode %r" % out.get("mode"))
    if mode == "fused_step":
        # phase: 'both' = the whole step; 'photo' / 'norm' / 'stencil' / 'finalize' = that kernel only (per-kernel timing)
        parts = _cabi.STEP_PARTS["all" if phase == "both" else phase]
        with torch.cuda.device_of(img[0]):
            _cabi.check(_cabi.lib().ugl_flow_loss_step_parts(C.byref(a), parts), "ugl_flow_loss_step")
        _count(bin(parts).count("1"))
        return out
    with torch.cuda.device_of(img[0]):
        if out["basis"] is not None:
            if fwd:
                _cabi.check(_cabi.lib().ugl_flow_loss_forward_grad_ex(C.byref(a), _variant()), "ugl_flow_loss_forward_grad")
            if bwd:
                _cabi.check(_cabi.lib().ugl_flow_loss_combine(C.byref(a)), "ugl_flow_loss_combine")
        else:
            if fwd:
                _cabi.check(_cabi.lib().ugl_flow_loss_forward(C.byref(a)), "ugl_flow_loss_forward")
            if bwd:
                _cabi.check(_cabi.lib().ugl_flow_loss_backward(C.byref(a)), "ugl_flow_loss_backward")
    _count(((3 if out["basis"] is not None and SINGLE_PASS_VARIANT != "fused" else 2) if fwd else 0) + (1 if bwd else 0))
    return out


def flow_loss(img_l_pyr: Sequence[Tensor], img_pyr: Sequence[Tensor], img_r_pyr: Sequence[Tensor],
              flows_fwd: Sequence[Tensor], flows_bwd: Sequence[Tensor], num_scales: Optional[int] = None,
              as_matrix: bool = False, mode: str = "single_pass"):
    """Fused loss body of ``Model_flow.forward`` (model_flow.py:232-254).

    Takes the three image pyramids (``generate_img_pyramid`` outputs) and the forward / backward flow
    pyramids, returns ``{'loss_flow_pixel', 'loss_flow_ssim', 'loss_flow_smooth', 'loss_flow_consis'}``,
    each a ``(B,)`` tensor differentiable w.r.t. the flows of levels ``< num_scales``
    (``as_matrix=True`` returns the underlying ``(4,B)`` tensor instead)."""
    L = len(flows_fwd)
    if not (len(img_l_pyr) >= L and len(img_pyr) >= L and len(img_r_pyr) >= L and len(flows_bwd) == L):
        raise ValueError("flow_loss: pyramids must have at least len(flows_fwd)=%d levels" % L)
    scales = L if num_scales is None else int(num_scales)
    if not 1 <= scales <= L or L > _cabi.MAX_LEVELS:
        raise ValueError("flow_loss: num_scales=%d outside [1, %d]" % (scales, L))
    if mode not in ("single_pass", "recompute"):
        raise ValueError("flow_loss: mode must be 'single_pass' or 'recompute'")
    loss = _FlowLossFn.apply(scales, L, mode, *img_l_pyr[:L], *img_pyr[:L], *img_r_pyr[:L], *flows_fwd, *flows_bwd)
    if as_matrix:
        return loss
    return {k: loss[i] for i, k in enumerate(FLOW_LOSS_KEYS)}


# ================================================================================================
# flow branch of the geom-mode loss (same single-pass kernel, Model_geometry's masks)
# ================================================================================================
MASK_VALID_BWD, MASK_VALID_FWD, MASK_OCC_BWD, MASK_OCC_FWD, MASK_DYN_BWD, MASK_DYN_FWD = 1, 2, 4, 8, 16, 32
MASK_ALL_BWD = MASK_VALID_BWD | MASK_OCC_BWD | MASK_DYN_BWD    # model_geometry.py:857
MASK_ALL_FWD = MASK_VALID_FWD | MASK_OCC_FWD | MASK_DYN_FWD    # model_geometry.py:858


def _geom_args(S, L, img_l, img, img_r, ff, fb, disp, Kinv, P_b, P_f, masks, alpha, beta, loss, stats, ws, basis, gloss=None, gf=None,
               gb=None) -> _cabi.UglGeomFlowArgs:
    g = _cabi.UglGeomFlowArgs()
    g.flow = _flow_args(img_l, img, img_r, ff, fb, S, loss, stats, ws, gloss, gf, gb, basis=basis)
    for l in range(S):
        g.disp[l], g.Kinv[l], g.P_bwd[l], g.P_fwd[l] = disp[l].data_ptr(), Kinv[l].data_ptr(), P_b[l].data_ptr(), P_f[l].data_ptr()
        g.mask_bytes[l] = masks[l].data_ptr()
    g.alpha, g.beta = float(alpha), float(beta)
    return g


class _GeomFlowLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, S: int, L: int, alpha: float, beta: float, *ts: Tensor):
        ts = tuple(_dev(t, "input %d" % i) for i, t in enumerate(ts))
        img_l, img, img_r, ff, fb = (ts[k * L:(k + 1) * L] for k in range(5))
        disp, Kinv, P_b, P_f = (ts[5 * L + k * S:5 * L + (k + 1) * S] for k in range(4))
        B, dev = img[0].shape[0], img[0].device
        for l in range(S):
            h, w = img[l].shape[2], img[l].shape[3]
            for name, t, shp in (("img_l", img_l[l], (B, 3, h, w)), ("img_r", img_r[l], (B, 3, h, w)), ("flow_fwd", ff[l], (B, 2, h, w)),
                                 ("flow_bwd", fb[l], (B, 2, h, w)), ("disp", disp[l], (B, 1, h, w)), ("Kinv", Kinv[l], (B, 3, 3)),
                                 ("P_bwd", P_b[l], (B, 3, 4)), ("P_fwd", P_f[l], (B, 3, 4))):
                if tuple(t.shape) != shp:
                    raise ValueError("geom_flow_loss: %s[%d] has shape %s, expected %s" % (name, l, tuple(t.shape), shp))
        loss = torch.empty((4, B), device=dev, dtype=torch.float32)
        stats = torch.empty((B, S, _cabi.GEOM_NSTATS), device=dev, dtype=torch.float32)
        masks = [torch.empty((B,) + tuple(img[l].shape[2:]), device=dev, dtype=torch.uint8) for l in range(S)]
        # Training-step form (mode_steps sets ``ctx.step_gloss``: the upstream gradient d total / d loss_k[b] = w_k / B is known before
        # the forward runs, train.py:211-215): ugl_geom_flow_step writes the flow gradients in the forward launch sequence -- no basis
        # planes, no combine launch; backward hands them out.
        step_gloss = getattr(ctx, "step_gloss", None)
        ctx.step_grads = None
        if step_gloss is not None:
            gl = _dev(step_gloss, "step grad_loss")
            if tuple(gl.shape) != (4, B):
                raise ValueError("geom_flow_loss: step grad_loss must be (4,%d), got %s" % (B, tuple(gl.shape)))
            gf = [torch.empty_like(ff[l]) for l in range(S)]
            gb = [torch.empty_like(fb[l]) for l in range(S)]
            g = _geom_args(S, L, img_l, img, img_r, ff, fb, disp, Kinv, P_b, P_f, masks, alpha, beta, loss, stats, None, None, gl, gf, gb)
            ws = torch.empty(max(int(_cabi.lib().ugl_flow_loss_workspace_bytes(C.byref(g.flow))) // 4, 1), device=dev, dtype=torch.float32)
            g.flow.workspace, g.flow.workspace_bytes = ws.data_ptr(), ws.numel() * 4
            after_photo = getattr(ctx, "after_photo", None)
            with torch.cuda.device_of(img[0]):
                if after_photo is None:
                    _call("ugl_geom_flow_step", C.byref(g), launches=4)
                else:       # mode_steps: the consumers of the mask bytes start on side streams as soon as the photometry kernel is queued
                    _call("ugl_geom_flow_step_parts", C.byref(g), _cabi.STEP_PARTS["photo"], launches=1)
                    after_photo(masks)
                    _call("ugl_geom_flow_step_parts", C.byref(g), _cabi.STEP_PARTS["all"] & ~_cabi.STEP_PARTS["photo"], launches=3)
            ctx.step_grads = (gf, gb)
            ctx.S, ctx.L, ctx.ab = S, L, (alpha, beta)
            ctx.mark_non_differentiable(*masks)
            ctx.set_materialize_grads(False)
            return (loss, *masks)
        basis = _alloc_basis(ff, S)
        g = _geom_args(S, L, img_l, img, img_r, ff, fb, disp, Kinv, P_b, P_f, masks, alpha, beta, loss, stats, None, basis)
        ws = torch.empty(max(int(_cabi.lib().ugl_flow_loss_workspace_bytes(C.byref(g.flow))) // 4, 1), device=dev, dtype=torch.float32)
        g.flow.workspace, g.flow.workspace_bytes = ws.data_ptr(), ws.numel() * 4
        with torch.cuda.device_of(img[0]):
            _call("ugl_geom_flow_forward_grad_ex", C.byref(g), _variant(), launches=2 if SINGLE_PASS_VARIANT == "fused" else 3)
        ctx.save_for_backward(stats, *ts, *basis, *masks)
        ctx.S, ctx.L, ctx.ab = S, L, (alpha, beta)
        ctx.mark_non_differentiable(*masks)
        ctx.set_materialize_grads(False)
        return (loss, *masks)

    @staticmethod
    def backward(ctx, gloss, *unused):
        S, L = ctx.S, ctx.L
        n_in = 5 * L + 4 * S
        if gloss is None:
            return (None,) * (4 + n_in)
        if getattr(ctx, "step_grads", None) is not None:      # training-step form: computed by the forward launches for ctx.step_gloss
            gf, gb = ctx.step_grads
            none_l, pad, none_s = [None] * L, [None] * (L - S), [None] * S
            return (None, None, None, None, *none_l, *none_l, *none_l, *gf, *pad, *gb, *pad, *none_s, *none_s, *none_s, *none_s)
        stats, *rest = ctx.saved_tensors
        ts, basis, masks = rest[:n_in], rest[n_in:n_in + S], rest[n_in + S:]
        img_l, img, img_r, ff, fb = (ts[k * L:(k + 1) * L] for k in range(5))
        disp, Kinv, P_b, P_f = (ts[5 * L + k * S:5 * L + (k + 1) * S] for k in range(4))
        gloss = _dev(gloss, "grad_loss")
        gf = [torch.empty_like(ff[l]) for l in range(S)]
        gb = [torch.empty_like(fb[l]) for l in range(S)]
        g = _geom_args(S, L, img_l, img, img_r, ff, fb, disp, Kinv, P_b, P_f, masks, *ctx.ab, None, stats, None, basis, gloss, gf, gb)
        with torch.cuda.device_of(gloss):
            _call("ugl_geom_flow_combine", C.byref(g))
        none_l, pad, none_s = [None] * L, [None] * (L - S), [None] * S
        return (None, None, None, None, *none_l, *none_l, *none_l, *gf, *pad, *gb, *pad, *none_s, *none_s, *none_s, *none_s)


def geom_flow_loss(img_l_pyr, img_pyr, img_r_pyr, flows_fwd, flows_bwd, disps, Kinv, P_bwd, P_fwd, alpha: float, beta: float,
                   num_scales: Optional[int] = None):
    """Flow branch of ``Model_geometry.forward``'s loss loop (model_geometry.py:845-919) in one stencil kernel: warps, hard
    occlusion + valid masks, rigid flow -> dynamic mask, the L1 terms split by the dynamic mask (weights 1 / 2), masked SSIM,
    second-order smoothness and direction consistency (mask ``1 - occ_fwd``), all ``num_scales`` levels.

    Returns ``(loss (4,B), mask_bytes[S])``: rows flow_pixel, flow_ssim, flow_smooth, flow_consis; ``mask_bytes[l]`` is a
    ``(B,h,w)`` uint8 map with the ``MASK_*`` bits (see :func:`unpack_mask`).  Differentiable w.r.t. the flows only (the
    reference detaches every mask)."""
    S = len(disps) if num_scales is None else int(num_scales)
    if not (1 <= S <= _cabi.MAX_LEVELS) or min(len(x) for x in (img_l_pyr, img_pyr, img_r_pyr, flows_fwd, flows_bwd, disps, Kinv, P_bwd,
                                                                 P_fwd)) < S:
        raise ValueError("geom_flow_loss: every pyramid / per-scale list needs at least num_scales=%d levels" % S)
    # only the first num_scales levels carry a loss (the flow pyramids of the reference have one more, unused here)
    out = _GeomFlowLossFn.apply(S, S, float(alpha), float(beta), *img_l_pyr[:S], *img_pyr[:S], *img_r_pyr[:S], *flows_fwd[:S],
                                *flows_bwd[:S], *disps[:S], *Kinv[:S], *P_bwd[:S], *P_fwd[:S])
    return out[0], list(out[1:])


class _GeomRigidFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow_b, flow_f, disp, mask_bytes, Kinv, P_b, P_f, F_b, F_f, need):
        ts = [_dev(t, n) for t, n in ((flow_b, "flow_bwd"), (flow_f, "flow_fwd"), (disp, "disp"), (Kinv, "Kinv"), (P_b, "P_bwd"),
                                      (P_f, "P_fwd"), (F_b, "F_bwd"), (F_f, "F_fwd"))]
        flow_b, flow_f, disp, Kinv, P_b, P_f, F_b, F_f = ts
        mask_bytes = _dev_u8(mask_bytes, "mask_bytes")
        B, _, H, W = flow_b.shape
        if (tuple(flow_f.shape) != (B, 2, H, W) or tuple(disp.shape) != (B, 1, H, W) or tuple(mask_bytes.shape) != (B, H, W)
                or any(tuple(t.shape) != (B, 3, 4) for t in (P_b, P_f)) or any(tuple(t.shape) != (B, 3, 3) for t in (Kinv, F_b, F_f))):
            raise ValueError("geom_rigid_terms: inconsistent shapes")
        dev = flow_b.device
        dfc, epi = torch.empty(B, device=dev, dtype=torch.float32), torch.empty(B, device=dev, dtype=torch.float32)
        den = torch.empty((B, 2), device=dev, dtype=torch.float32)
        a = _GeomRigidFn._args(flow_b, flow_f, disp, mask_bytes, Kinv, P_b, P_f, F_b, F_f, need, den)
        a.loss_dfc, a.loss_epi = dfc.data_ptr(), epi.data_ptr()
        with torch.cuda.device_of(flow_b):
            _call("ugl_geom_rigid_forward", C.byref(a), launches=2)
        ctx.save_for_backward(flow_b, flow_f, disp, mask_bytes, Kinv, P_b, P_f, F_b, F_f, den)
        ctx.need = need
        ctx.set_materialize_grads(False)
        return dfc, epi

    @staticmethod
    def _args(flow_b, flow_f, disp, mask_bytes, Kinv, P_b, P_f, F_b, F_f, need, den):
        a = _cabi.UglGeomRigidArgs()
        B, _, H, W = flow_b.shape
        a.batch, a.height, a.width = B, H, W
        a.need[0], a.need[1] = int(need[0]), int(need[1])
        a.flow_bwd, a.flow_fwd, a.disp, a.mask_bytes = flow_b.data_ptr(), flow_f.data_ptr(), disp.data_ptr(), mask_bytes.data_ptr()
        a.Kinv, a.P_bwd, a.P_fwd, a.F_bwd, a.F_fwd = Kinv.data_ptr(), P_b.data_ptr(), P_f.data_ptr(), F_b.data_ptr(), F_f.data_ptr()
        a.den = den.data_ptr()
        n = int(_cabi.lib().ugl_geom_rigid_workspace_bytes(B, H, W))
        a._ws = torch.empty((n + 7) // 8, dtype=torch.int64, device=flow_b.device)      # kept alive by the struct object
        a.workspace, a.workspace_bytes = a._ws.data_ptr(), _nbytes(a._ws)
        a.stream = torch.cuda.current_stream().cuda_stream
        return a

    @staticmethod
    def backward(ctx, g_dfc, g_epi):
        flow_b, flow_f, disp, mask_bytes, Kinv, P_b, P_f, F_b, F_f, den = ctx.saved_tensors
        if g_dfc is None and g_epi is None:
            return (None,) * 10
        B, dev = flow_b.shape[0], flow_b.device
        g_dfc = _dev(g_dfc, "grad_dfc") if g_dfc is not None else None
        g_epi = _dev(g_epi, "grad_epi") if g_epi is not None else None
        gfb, gff, gd = torch.empty_like(flow_b), torch.empty_like(flow_f), torch.empty_like(disp)
        gPb, gPf = torch.empty_like(P_b), torch.empty_like(P_f)
        gFb, gFf = torch.empty_like(F_b), torch.empty_like(F_f)
        a = _GeomRigidFn._args(flow_b, flow_f, disp, mask_bytes, Kinv, P_b, P_f, F_b, F_f, ctx.need, den)
        a.grad_dfc, a.grad_epi = _ptr(g_dfc), _ptr(g_epi)
        a.grad_flow_bwd, a.grad_flow_fwd, a.grad_disp = gfb.data_ptr(), gff.data_ptr(), gd.data_ptr()
        a.grad_P_bwd, a.grad_P_fwd, a.grad_F_bwd, a.grad_F_fwd = gPb.data_ptr(), gPf.data_ptr(), gFb.data_ptr(), gFf.data_ptr()
        with torch.cuda.device_of(flow_b):
            _call("ugl_geom_rigid_backward", C.byref(a), launches=2)
        return gfb, gff, gd, None, None, gPb, gPf, gFb, gFf, None


def geom_rigid_terms(flow_bwd: Tensor, flow_fwd: Tensor, disp: Tensor, mask_bytes: Tensor, Kinv: Tensor, P_bwd: Tensor, P_fwd: Tensor,
                     F_bwd: Tensor, F_fwd: Tensor, need=(MASK_ALL_BWD, MASK_ALL_FWD)):
    """``loss_depth_flow_consis`` and ``loss_epipolar`` of ``Model_geometry.forward`` (model_geometry.py:921-935) at level 0, both
    directions, one forward and one backward kernel -> ``(dfc (B,), epi (B,))``.  ``mask_bytes`` is level 0 of
    :func:`geom_flow_loss`; differentiable w.r.t. the flows, the disparity, ``P_*`` and ``F_*``."""
    return _GeomRigidFn.apply(flow_bwd, flow_fwd, disp, mask_bytes, Kinv, P_bwd, P_fwd, F_bwd, F_fwd, tuple(need))


def unpack_mask(mask_bytes: Tensor, bits: int, invert: bool = False) -> Tensor:
    """``(B,h,w)`` packed map -> ``(B,1,h,w)`` float {0,1} mask: 1 where every bit of ``bits`` is set (plain torch; off the
    hot path — the kernels consume the packed maps directly)."""
    m = (mask_bytes & bits) == bits
    if invert:
        m = ~m
    return m.unsqueeze(1).float()


# ================================================================================================
# stand-alone terms, masks and geometry (one kernel family per reference method)
# ================================================================================================
def _call(name: str, *args, launches: int = 1) -> None:
    _cabi.check(getattr(_cabi.lib(), name)(*args), name)
    _count(launches)


def _reduce_ws(B: int, H: int, W: int, device) -> Tensor:
    n = int(_cabi.lib().ugl_reduce_workspace_bytes(B, H, W))
    return torch.empty((n + 7) // 8, dtype=torch.int64, device=device)


def _nbytes(t: Tensor) -> int:
    return t.numel() * t.element_size()


class _MaskedMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, mask, mode):
        a = _dev(a, "a")
        b = _dev(b, "b") if b is not None else None
        mask = _dev(mask, "mask").detach() if mask is not None else None
        B, Cc, H, W = a.shape
        if b is not None and b.shape != a.shape:
            raise ValueError("masked_mean: a %s and b %s differ" % (tuple(a.shape), tuple(b.shape)))
        if mask is not None and tuple(mask.shape) != (B, 1, H, W):
            raise ValueError("masked_mean: mask must be (B,1,H,W), got %s" % (tuple(mask.shape),))
        out = torch.empty(B, device=a.device, dtype=torch.float32)
        den = torch.empty_like(out)
        ws = _reduce_ws(B, H, W, a.device)
        with torch.cuda.device_of(a):
            _call("ugl_masked_mean_forward", a.data_ptr(), _ptr(b), _ptr(mask), B, Cc, H, W, mode, out.data_ptr(), den.data_ptr(),
                  ws.data_ptr(), _nbytes(ws), _stream_ptr(), launches=2)
        ctx.save_for_backward(a, b, mask, den)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, gout):
        a, b, mask, den = ctx.saved_tensors
        B, Cc, H, W = a.shape
        gout = _dev(gout, "grad_out")
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and b is not None
        ga = torch.empty_like(a) if need_a else None
        gb = torch.empty_like(b) if need_b else None
        if ga is not None or gb is not None:
            with torch.cuda.device_of(a):
                _call("ugl_masked_mean_backward", a.data_ptr(), _ptr(b), _ptr(mask), den.data_ptr(), gout.data_ptr(), B, Cc, H, W,
                      ctx.mode, _ptr(ga), _ptr(gb), _stream_ptr())
        return ga, gb, None, None


def masked_l1(img: Tensor, warped: Tensor, mask: Optional[Tensor]) -> Tensor:
    """One level of ``compute_photometric_loss`` (model_geometry.py:143-153): P(|img - warped|, mask) -> (B,)."""
    return _MaskedMeanFn.apply(img, warped, mask, 0)


def masked_mean(diff: Tensor, mask: Optional[Tensor]) -> Tensor:
    """One level of ``compute_loss_with_mask`` / ``compute_depth_flow_consis_loss``: P(diff, mask) -> (B,);
    ``mask=None`` is the plain per-sample mean."""
    return _MaskedMeanFn.apply(diff, None, mask, 1)


class _SsimFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        x, y = _dev(x, "x"), _dev(y, "y")
        if x.shape != y.shape or x.dim() != 4:
            raise ValueError("SSIM: x %s and y %s must be equal 4-d shapes" % (tuple(x.shape), tuple(y.shape)))
        B, Cc, H, W = x.shape
        out = torch.empty_like(x)
        with torch.cuda.device_of(x):
            _call("ugl_ssim_forward", x.data_ptr(), y.data_ptr(), B, Cc, H, W, out.data_ptr(), _stream_ptr())
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, y = ctx.saved_tensors
        B, Cc, H, W = x.shape
        gout = _dev(gout, "grad_out")
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        with torch.cuda.device_of(x):
            _call("ugl_ssim_backward", x.data_ptr(), y.data_ptr(), gout.data_ptr(), B, Cc, H, W, _ptr(gx), _ptr(gy), _stream_ptr())
        return gx, gy


def ssim(x: Tensor, y: Tensor) -> Tensor:
    """Drop-in for ``SSIM(x, y)`` (pytorch_ssim/ssim.py:4-19)."""
    return _SsimFn.apply(x, y)


class _SsimLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, warped, mask):
        img, warped = _dev(img, "img"), _dev(warped, "warped")
        mask = _dev(mask, "mask").detach() if mask is not None else None
        B, Cc, H, W = img.shape
        out = torch.empty(B, device=img.device, dtype=torch.float32)
        den = torch.empty_like(out)
        n = int(_cabi.lib().ugl_ssim_loss_workspace_bytes(B, Cc, H, W))
        ws = torch.empty((n + 7) // 8, dtype=torch.int64, device=img.device)
        with torch.cuda.device_of(img):
            _call("ugl_ssim_loss_forward", img.data_ptr(), warped.data_ptr(), _ptr(mask), B, Cc, H, W, out.data_ptr(), den.data_ptr(),
                  ws.data_ptr(), _nbytes(ws), _stream_ptr(), launches=2)
        ctx.save_for_backward(img, warped, mask, den)
        return out

    @staticmethod
    def backward(ctx, gout):
        img, warped, mask, den = ctx.saved_tensors
        B, Cc, H, W = img.shape
        gout = _dev(gout, "grad_out")
        gi = torch.empty_like(img) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(warped) if ctx.needs_input_grad[1] else None
        if gi is not None or gw is not None:
            with torch.cuda.device_of(img):
                _call("ugl_ssim_loss_backward", img.data_ptr(), warped.data_ptr(), _ptr(mask), den.data_ptr(), gout.data_ptr(), B, Cc,
                      H, W, _ptr(gi), _ptr(gw), _stream_ptr())
        return gi, gw, None


def ssim_loss(img: Tensor, warped: Tensor, mask: Optional[Tensor]) -> Tensor:
    """One level of ``compute_ssim_loss`` / ``compute_loss_ssim`` (model_geometry.py:212-223) -> (B,)."""
    return _SsimLossFn.apply(img, warped, mask)


class _OccWeightsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, from_l, img, from_r, soft):
        from_l, img, from_r = _dev(from_l, "from_l"), _dev(img, "img"), _dev(from_r, "from_r")
        B, Cc, H, W = img.shape
        if Cc != 3:
            raise ValueError("occlusion weights expect 3-channel images")
        o = [torch.empty((B, 1, H, W), device=img.device, dtype=torch.float32) for _ in range(6)]
        with torch.cuda.device_of(img):
            _call("ugl_occlusion_weights", from_l.data_ptr(), img.data_ptr(), from_r.data_ptr(), B, H, W, int(soft),
                  *[t.data_ptr() for t in o], _stream_ptr())
        ctx.save_for_backward(from_l, img, from_r)
        ctx.mark_non_differentiable(*o[:4])
        ctx.set_materialize_grads(False)
        return tuple(o)

    @staticmethod
    def backward(ctx, g0, g1, g2, g3, g_db, g_df):
        from_l, img, from_r = ctx.saved_tensors
        B, Cc, H, W = img.shape
        gl = gr = None
        with torch.cuda.device_of(img):
            if ctx.needs_input_grad[0] and g_db is not None:
                gl = torch.empty_like(from_l)
                _call("ugl_channel_mean_abs_diff_backward", img.data_ptr(), from_l.data_ptr(), _dev(g_db, "g").data_ptr(), B, Cc, H, W,
                      gl.data_ptr(), _stream_ptr())
            if ctx.needs_input_grad[2] and g_df is not None:
                gr = torch.empty_like(from_r)
                _call("ugl_channel_mean_abs_diff_backward", img.data_ptr(), from_r.data_ptr(), _dev(g_df, "g").data_ptr(), B, Cc, H, W,
                      gr.data_ptr(), _stream_ptr())
        return gl, None, gr, None


def occlusion_weights(from_l: Tensor, img: Tensor, from_r: Tensor, soft: bool):
    """(w_bwd, w_fwd, valid_bwd, valid_fwd, diff_bwd, diff_fwd) for one level — ``compute_occ_weight``
    (soft=False, model_geometry.py:105-132) / ``compute_diff_weight`` (soft=True, model_flow.py:105-138).
    The weights / valid maps are constants; the diffs are differentiable w.r.t. the warped images."""
    return _OccWeightsFn.apply(from_l, img, from_r, bool(soft))


def texture_mask(img: Tensor, rec: Tensor, src: Tensor) -> Tensor:
    """One level of ``compute_texture_mask`` (model_geometry.py:134-140); constant for autograd."""
    img, rec, src = _dev(img, "img").detach(), _dev(rec, "rec").detach(), _dev(src, "src").detach()
    B, Cc, H, W = img.shape
    out = torch.empty((B, 1, H, W), device=img.device, dtype=torch.float32)
    with torch.cuda.device_of(img):
        _call("ugl_texture_mask", img.data_ptr(), rec.data_ptr(), src.data_ptr(), B, H, W, out.data_ptr(), _stream_ptr())
    return out


class _DynamicMaskFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, rflow, alpha, beta):
        flow, rflow = _dev(flow, "flow"), _dev(rflow, "rigid_flow")
        B, _, H, W = flow.shape
        fd = torch.empty_like(flow)
        dyn = torch.empty((B, 1, H, W), device=flow.device, dtype=torch.float32)
        score = torch.empty_like(dyn)
        with torch.cuda.device_of(flow):
            _call("ugl_dynamic_mask_forward", flow.data_ptr(), rflow.data_ptr(), B, H, W, float(alpha), float(beta), fd.data_ptr(),
                  dyn.data_ptr(), score.data_ptr(), _stream_ptr())
        ctx.save_for_backward(flow, rflow)
        ctx.mark_non_differentiable(dyn, score)
        ctx.set_materialize_grads(False)
        return fd, dyn, score

    @staticmethod
    def backward(ctx, g_fd, g_dyn, g_score):
        flow, rflow = ctx.saved_tensors
        if g_fd is None:
            return None, None, None, None
        g_fd = _dev(g_fd, "grad")
        gf = torch.empty_like(flow) if ctx.needs_input_grad[0] else None
        gr = torch.empty_like(rflow) if ctx.needs_input_grad[1] else None
        if gf is not None or gr is not None:
            with torch.cuda.device_of(flow):   # fd = |rflow - flow|
                _call("ugl_abs_diff_backward", rflow.data_ptr(), flow.data_ptr(), g_fd.data_ptr(), flow.numel(), _ptr(gr), _ptr(gf),
                      _stream_ptr())
        return gf, gr, None, None


def dynamic_mask(flow: Tensor, rigid_flow: Tensor, alpha: float, beta: float):
    """Body of ``compute_dynamic_mask`` for one level (model_geometry.py:698-711): (flow_diff, dyn_mask, score)."""
    return _DynamicMaskFn.apply(flow, rigid_flow, alpha, beta)


def mask_product(masks: Sequence[Tensor], invert: Optional[Sequence[bool]] = None) -> Tensor:
    """``fusion_mask*`` (model_geometry.py:735-765): product of up to four maps, ``invert[k]`` uses 1 - m."""
    ms = [_dev(m, "mask").detach() for m in masks]
    if not 1 <= len(ms) <= 4 or any(m.shape != ms[0].shape for m in ms):
        raise ValueError("mask_product: need 1..4 maps of one shape")
    out = torch.empty_like(ms[0])
    arr = (C.c_void_p * len(ms))(*[m.data_ptr() for m in ms])
    inv = (C.c_int32 * len(ms))(*[int(bool(v)) for v in (invert or [False] * len(ms))])
    with torch.cuda.device_of(out):
        _call("ugl_mask_product", arr, inv, len(ms), out.numel(), out.data_ptr(), _stream_ptr())
    return out


def rigid_mask(dist: Tensor, rigid_thres: float = 0.5, inlier_thres: float = 0.1):
    """``get_rigid_mask`` (model_geometry.py:420-425): (rigid, inlier, score), constants for autograd."""
    dist = _dev(dist, "dist").detach()
    o = [torch.empty_like(dist) for _ in range(3)]
    with torch.cuda.device_of(dist):
        _call("ugl_rigid_mask", dist.data_ptr(), dist.numel(), float(rigid_thres), float(inlier_thres), *[t.data_ptr() for t in o],
              _stream_ptr())
    return tuple(o)


class _FlowSmoothFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, img):
        flow, img = _dev(flow, "flow"), _dev(img, "img").detach()
        B, _, H, W = flow.shape
        out = torch.empty(B, device=flow.device, dtype=torch.float32)
        ws = _reduce_ws(B, H, W, flow.device)
        with torch.cuda.device_of(flow):
            _call("ugl_flow_smooth_forward", flow.data_ptr(), img.data_ptr(), B, H, W, out.data_ptr(), ws.data_ptr(), _nbytes(ws),
                  _stream_ptr(), launches=2)
        ctx.save_for_backward(flow, img)
        return out

    @staticmethod
    def backward(ctx, gout):
        flow, img = ctx.saved_tensors
        B, _, H, W = flow.shape
        g = torch.empty_like(flow)
        with torch.cuda.device_of(flow):
            _call("ugl_flow_smooth_backward", flow.data_ptr(), img.data_ptr(), _dev(gout, "g").data_ptr(), B, H, W, g.data_ptr(), _stream_ptr())
        return g, None


def flow_smooth(flow: Tensor, img: Tensor) -> Tensor:
    """``cal_grad2_error(flow/20, img)`` for one level (model_geometry.py:254-279) -> (B,)."""
    return _FlowSmoothFn.apply(flow, img)


class _FlowConsisFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fwd, bwd, occ):
        fwd, bwd, occ = _dev(fwd, "fwd"), _dev(bwd, "bwd").detach(), _dev(occ, "occ").detach()
        B, _, H, W = fwd.shape
        out = torch.empty(B, device=fwd.device, dtype=torch.float32)
        den = torch.empty_like(out)
        ws = _reduce_ws(B, H, W, fwd.device)
        with torch.cuda.device_of(fwd):
            _call("ugl_flow_consis_forward", fwd.data_ptr(), bwd.data_ptr(), occ.data_ptr(), B, H, W, out.data_ptr(), den.data_ptr(),
                  ws.data_ptr(), _nbytes(ws), _stream_ptr(), launches=2)
        ctx.save_for_backward(fwd, bwd, occ, den)
        return out

    @staticmethod
    def backward(ctx, gout):
        fwd, bwd, occ, den = ctx.saved_tensors
        B, _, H, W = fwd.shape
        g = torch.empty_like(fwd)
        with torch.cuda.device_of(fwd):
            _call("ugl_flow_consis_backward", fwd.data_ptr(), bwd.data_ptr(), occ.data_ptr(), den.data_ptr(), _dev(gout, "g").data_ptr(),
                  B, H, W, g.data_ptr(), _stream_ptr())
        return g, None, None


def flow_consis(fwd: Tensor, bwd: Tensor, occ: Tensor) -> Tensor:
    """One level of ``compute_loss_flow_consis`` (model_geometry.py:195-210) -> (B,)."""
    return _FlowConsisFn.apply(fwd, bwd, occ)


class _DepthDiffFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, comp, proj):
        comp, proj = _dev(comp, "computed_depth"), _dev(proj, "predicted_depth")
        out = torch.empty_like(comp)
        with torch.cuda.device_of(comp):
            _call("ugl_depth_diff_forward", comp.data_ptr(), proj.data_ptr(), comp.numel(), out.data_ptr(), _stream_ptr())
        ctx.save_for_backward(comp, proj)
        return out

    @staticmethod
    def backward(ctx, g):
        comp, proj = ctx.saved_tensors
        gc = torch.empty_like(comp) if ctx.needs_input_grad[0] else None
        gp = torch.empty_like(proj) if ctx.needs_input_grad[1] else None
        if gc is not None or gp is not None:
            with torch.cuda.device_of(comp):
                _call("ugl_depth_diff_backward", comp.data_ptr(), proj.data_ptr(), _dev(g, "g").data_ptr(), comp.numel(), _ptr(gc), _ptr(gp),
                      _stream_ptr())
        return gc, gp


def depth_diff(comp: Tensor, proj: Tensor) -> Tensor:
    """clamp(|comp - proj| / |comp + proj|, 0, 1) of ``compute_consis_loss`` (model_geometry.py:186-188)."""
    return _DepthDiffFn.apply(comp, proj)


class _DispSmoothFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, *disps):
        img = _dev(img, "img").detach()
        disps = [_dev(d, "disp") for d in disps]
        B, _, H, W = img.shape
        L = len(disps)
        out = torch.empty(B, device=img.device, dtype=torch.float32)
        ws = _reduce_ws(B, H, W, img.device)
        arr = (C.c_void_p * L)(*[d.data_ptr() for d in disps])
        hs = (C.c_int32 * L)(*[d.shape[2] for d in disps])
        wsz = (C.c_int32 * L)(*[d.shape[3] for d in disps])
        with torch.cuda.device_of(img):
            _call("ugl_disp_smooth_forward", img.data_ptr(), arr, hs, wsz, L, B, H, W, out.data_ptr(), ws.data_ptr(), _nbytes(ws),
                  _stream_ptr(), launches=2)
        ctx.save_for_backward(img, *disps)
        return out

    @staticmethod
    def backward(ctx, gout):
        img, *disps = ctx.saved_tensors
        B, _, H, W = img.shape
        L = len(disps)
        grads = [torch.empty_like(d) for d in disps]
        n = int(_cabi.lib().ugl_disp_smooth_backward_workspace_bytes(B, H, W))
        ws = torch.empty((n + 7) // 8, dtype=torch.int64, device=img.device)
        arr = (C.c_void_p * L)(*[d.data_ptr() for d in disps])
        garr = (C.c_void_p * L)(*[g.data_ptr() for g in grads])
        hs = (C.c_int32 * L)(*[d.shape[2] for d in disps])
        wsz = (C.c_int32 * L)(*[d.shape[3] for d in disps])
        with torch.cuda.device_of(img):
            _call("ugl_disp_smooth_backward", img.data_ptr(), arr, hs, wsz, L, _dev(gout, "g").data_ptr(), B, H, W, garr, ws.data_ptr(),
                  _nbytes(ws), _stream_ptr(), launches=2 * L - 1)
        return (None, *grads)


def _disp_smooth_args(imgs, disps, out, G, gout=None, gdisp=None) -> _cabi.UglDispSmoothArgs:
    a = _cabi.UglDispSmoothArgs()
    nl, L = len(imgs), len(disps[0])
    a.batch, a.lists, a.levels, a.height, a.width = imgs[0].shape[0], nl, L, imgs[0].shape[2], imgs[0].shape[3]
    for l in range(L):
        a.lheight[l], a.lwidth[l] = disps[0][l].shape[2], disps[0][l].shape[3]
    for i in range(nl):
        a.img[i] = _ptr(imgs[i])
        for l in range(L):
            a.disp[i][l] = _ptr(disps[i][l])
            if G is not None:
                a.G[i][l] = G[i][l].data_ptr()
            if gdisp is not None:
                a.grad_disp[i][l] = gdisp[i][l].data_ptr()
    a.out, a.grad_out = _ptr(out), _ptr(gout)
    a.stream = torch.cuda.current_stream().cuda_stream
    return a


class _DispSmoothMultiFn(torch.autograd.Function):
    """single-pass form: forward also writes G = d loss / d (up-sampled disparity); backward = transpose-gather of G"""

    @staticmethod
    def forward(ctx, nl, L, *ts):
        imgs = [_dev(t, "img").detach() for t in ts[:nl]]
        disps = [[_dev(t, "disp") for t in ts[nl + i * L:nl + (i + 1) * L]] for i in range(nl)]
        B, _, H, W = imgs[0].shape
        dev = imgs[0].device
        for i in range(nl):
            if tuple(imgs[i].shape) != (B, 3, H, W):
                raise ValueError("disp_smooth: images must share one (B,3,H,W) shape")
            for l in range(L):
                if tuple(disps[i][l].shape) != tuple(disps[0][l].shape) or disps[i][l].shape[:2] != (B, 1):
                    raise ValueError("disp_smooth: disparity pyramids of all lists must share their (B,1,h,w) shapes")
        need = any(ctx.needs_input_grad[2 + nl:])
        out = torch.empty((nl, B), device=dev, dtype=torch.float32)
        G = [[torch.empty((B, 1, H, W), device=dev, dtype=torch.float32) for _ in range(L)] for _ in range(nl)] if need else None
        a = _disp_smooth_args(imgs, disps, out, G)
        n = int(_cabi.lib().ugl_disp_smooth_fused_workspace_bytes(C.byref(a)))
        ws = torch.empty((n + 7) // 8, dtype=torch.int64, device=dev)
        a.workspace, a.workspace_bytes = ws.data_ptr(), _nbytes(ws)
        with torch.cuda.device_of(imgs[0]):
            _call("ugl_disp_smooth_forward_grad", C.byref(a), launches=2)
        if need:
            ctx.save_for_backward(*[g for row in G for g in row])
        ctx.nl, ctx.L, ctx.shapes, ctx.full = nl, L, [tuple(d.shape) for d in disps[0]], (B, H, W)
        return out

    @staticmethod
    def backward(ctx, gout):
        nl, L = ctx.nl, ctx.L
        Gs = ctx.saved_tensors
        if not Gs:
            return (None,) * (2 + nl + nl * L)
        G = [list(Gs[i * L:(i + 1) * L]) for i in range(nl)]
        gout = _dev(gout, "grad_out")             # (nl,B), or ONE (B,) row shared by all lists (mode_steps: no expand + copy launch)
        dev = gout.device
        gd = [[torch.empty(ctx.shapes[l], device=dev, dtype=torch.float32) for l in range(L)] for _ in range(nl)]

        class _Shape:   # shapes only: combine reads G, not the images / disparities
            def __init__(self, shape):
                self.shape = shape

        B, H, W = ctx.full
        a = _disp_smooth_args([_Shape((B, 3, H, W))] * nl, [[_Shape(s) for s in ctx.shapes]] * nl, None, G, gout, gd)
        a.grad_out_shared = 1 if gout.dim() == 1 else 0
        with torch.cuda.device_of(gout):
            _call("ugl_disp_smooth_combine", C.byref(a))
        return (None, None, *([None] * nl), *[g for row in gd for g in row])


def disp_smooth_multi(imgs: Sequence[Tensor], disps: Sequence[Sequence[Tensor]]) -> Tensor:
    """``compute_smooth_loss`` for up to three (image, disparity pyramid) lists in one launch -> (lists, B)
    (model_geometry.py:938-940 calls it for the centre, left and right frames back to back)."""
    nl = len(imgs)
    if not 1 <= nl <= _cabi.UglDispSmoothArgs.MAX_LISTS or len(disps) != nl or any(len(d) != len(disps[0]) for d in disps):
        raise ValueError("disp_smooth_multi: 1..3 lists with equally many levels each")
    L = len(disps[0])
    return _DispSmoothMultiFn.apply(nl, L, *imgs, *[d for row in disps for d in row])


def disp_smooth(img: Tensor, disps: Sequence[Tensor], mode: str = "single_pass") -> Tensor:
    """``compute_smooth_loss(img, disps)`` (model_geometry.py:225-252) for ``len(disps)`` levels -> (B,).
    ``mode='recompute'`` keeps nothing per pixel between forward and backward (the per-method kernels)."""
    if mode == "recompute":
        return _DispSmoothFn.apply(img, *disps)
    return disp_smooth_multi([img], [list(disps)])[0]


class _ReprojectFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, depth, ref_depth, Kinv, P):
        img, depth, ref_depth = _dev(img, "img"), _dev(depth, "depth"), _dev(ref_depth, "ref_depth")
        Kinv, P = _dev(Kinv, "Kinv").detach(), _dev(P, "P")
        B, Cc, H, W = img.shape
        out = torch.empty_like(img)
        valid, proj, comp = (torch.empty((B, 1, H, W), device=img.device, dtype=torch.float32) for _ in range(3))
        with torch.cuda.device_of(img):
            _call("ugl_reproject_forward", img.data_ptr(), depth.data_ptr(), ref_depth.data_ptr(), Kinv.data_ptr(), P.data_ptr(), B, Cc, H,
                  W, out.data_ptr(), valid.data_ptr(), proj.data_ptr(), comp.data_ptr(), _stream_ptr())
        ctx.save_for_backward(img, depth, ref_depth, Kinv, P)
        ctx.mark_non_differentiable(valid)
        ctx.set_materialize_grads(False)     # unused outputs arrive as None, not as zero maps to scatter
        return out, valid, proj, comp

    @staticmethod
    def backward(ctx, g_img, g_valid, g_proj, g_comp):
        img, depth, ref_depth, Kinv, P = ctx.saved_tensors
        B, Cc, H, W = img.shape
        need_img, need_depth, need_ref, _, need_P = ctx.needs_input_grad
        if g_img is None and g_proj is None and g_comp is None:
            return None, None, None, None, None
        need_img = need_img and g_img is not None          # d/d img only flows through the sampled image,
        need_ref = need_ref and g_proj is not None         # d/d ref_depth only through the projected depth
        g_img = _dev(g_img, "g") if g_img is not None else None
        g_proj = _dev(g_proj, "g") if g_proj is not None else None
        g_comp = _dev(g_comp, "g") if g_comp is not None else None
        gd = torch.empty_like(depth) if need_depth else None
        gP = torch.empty((B, 3, 4), device=img.device, dtype=torch.float32)
        gi = torch.empty_like(img) if need_img else None
        gr = torch.empty_like(ref_depth) if need_ref else None
        n = int(_cabi.lib().ugl_reproject_backward_workspace_bytes(B, Cc, H, W, int(need_img), int(need_ref)))
        ws = torch.empty((n + 7) // 8, dtype=torch.int64, device=img.device)
        with torch.cuda.device_of(img):
            _call("ugl_reproject_backward", img.data_ptr(), depth.data_ptr(), ref_depth.data_ptr(), Kinv.data_ptr(), P.data_ptr(),
                  _ptr(g_img), _ptr(g_proj), _ptr(g_comp), B, Cc, H, W, _ptr(gd), gP.data_ptr(), _ptr(gi), _ptr(gr), ws.data_ptr(),
                  _nbytes(ws), _stream_ptr(), launches=2 + (5 if (need_img or need_ref) else 0))
        return gi, gd, gr, None, (gP if need_P else None)


def reproject(img: Tensor, depth: Tensor, ref_depth: Tensor, Kinv: Tensor, P: Tensor):
    """Core of ``inverse_warp2`` (structures/inverse_warp.py:284-303) given K^-1 and P = K [R|t]:
    (projected_img, valid_mask, projected_depth, computed_depth)."""
    return _ReprojectFn.apply(img, depth, ref_depth, Kinv, P)


class _RigidFlowFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, Kinv, P):
        depth, Kinv, P = _dev(depth, "depth"), _dev(Kinv, "Kinv").detach(), _dev(P, "P")
        B, _, H, W = depth.shape
        out = torch.empty((B, 2, H, W), device=depth.device, dtype=torch.float32)
        with torch.cuda.device_of(depth):
            _call("ugl_rigid_flow_forward", depth.data_ptr(), Kinv.data_ptr(), P.data_ptr(), B, H, W, out.data_ptr(), _stream_ptr())
        ctx.save_for_backward(depth, Kinv, P)
        return out

    @staticmethod
    def backward(ctx, g):
        depth, Kinv, P = ctx.saved_tensors
        B, _, H, W = depth.shape
        gd = torch.empty_like(depth) if ctx.needs_input_grad[0] else None
        gP = torch.empty((B, 3, 4), device=depth.device, dtype=torch.float32)
        ws = _reduce_ws(B, H, W, depth.device)
        with torch.cuda.device_of(depth):
            _call("ugl_rigid_flow_backward", depth.data_ptr(), Kinv.data_ptr(), P.data_ptr(), _dev(g, "g").data_ptr(), B, H, W, _ptr(gd),
                  gP.data_ptr(), ws.data_ptr(), _nbytes(ws), _stream_ptr(), launches=2)
        return gd, None, (gP if ctx.needs_input_grad[2] else None)


def rigid_flow(depth: Tensor, Kinv: Tensor, P: Tensor) -> Tensor:
    """Core of ``calculate_rigid_flow`` (structures/inverse_warp.py:329-342) given K^-1 and P."""
    return _RigidFlowFn.apply(depth, Kinv, P)


class _PoseSetupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, K, K_inv, downs, want_F):
        pose, K = _dev(pose, "pose"), _dev(K, "K").detach()
        K_inv = _dev(K_inv, "K_inv").detach() if K_inv is not None else None
        if pose.dim() != 3 or pose.shape[2] != 6 or not 1 <= pose.shape[1] <= 2:
            raise ValueError("pose_setup: pose must be (B,n,6) with n <= 2, got %s" % (tuple(pose.shape),))
        B, n, _ = pose.shape
        if tuple(K.shape) != (B, 3, 3) or (K_inv is not None and tuple(K_inv.shape) != (B, 3, 3)):
            raise ValueError("pose_setup: K / K_inv must be (B,3,3)")
        if want_F and K_inv is None:
            raise ValueError("pose_setup: the fundamental matrices need K_inv")
        S, dev = len(downs), pose.device
        new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
        Kinv = [new(B, 3, 3) for _ in range(S)]
        Ps = [new(B, 3, 4) for _ in range(n * S)]
        Fs = [new(B, 3, 3) for _ in range(n)] if want_F else []
        darr = (C.c_float * S)(*[float(d) for d in downs])
        arr = lambda ts: (C.c_void_p * max(len(ts), 1))(*[t.data_ptr() for t in ts])
        with torch.cuda.device_of(pose):
            _call("ugl_pose_setup_forward", pose.data_ptr(), K.data_ptr(), _ptr(K_inv), darr, B, n, S, arr(Kinv), arr(Ps),
                  arr(Fs) if want_F else None, _stream_ptr())
        ctx.save_for_backward(pose, K, K_inv)
        ctx.downs, ctx.want_F = tuple(float(d) for d in downs), want_F
        ctx.mark_non_differentiable(*Kinv)
        ctx.set_materialize_grads(False)
        return (*Kinv, *Ps, *Fs)

    @staticmethod
    def backward(ctx, *grads):
        pose, K, K_inv = ctx.saved_tensors
        B, n, _ = pose.shape
        S = len(ctx.downs)
        gP, gF = grads[S:S + n * S], grads[S + n * S:]
        if all(g is None for g in (*gP, *gF)):
            return None, None, None, None, None
        gP = [(_dev(g, "grad_P") if g is not None else None) for g in gP]
        gF = [(_dev(g, "grad_F") if g is not None else None) for g in gF]
        ptrs = lambda gs: (C.c_void_p * max(len(gs), 1))(*[_ptr(g) for g in gs])
        gpose = torch.empty_like(pose)
        darr = (C.c_float * S)(*ctx.downs)
        with torch.cuda.device_of(pose):
            _call("ugl_pose_setup_backward", pose.data_ptr(), K.data_ptr(), _ptr(K_inv), darr, B, n, S, ptrs(gP),
                  ptrs(gF) if ctx.want_F else None, gpose.data_ptr(), _stream_ptr())
        return gpose, None, None, None, None


def pose_setup(pose: Tensor, K: Tensor, downscales: Sequence[float], K_inv: Optional[Tensor] = None, fundamental: bool = False):
    """Every 3x3 / 3x4 matrix a depth / geom step needs, in one launch: ``Kinv[s]`` (B,3,3) = inverse of the level's
    intrinsics, ``P[k][s]`` (B,3,4) = K_s [R|t] of ``pose[:,k]`` and, with ``fundamental=True``, ``F[k]`` (B,3,3) =
    K^-T [t]x R K^-1 (inverse_warp.py:110-145, 172-187, 284-289, 354-364; model_geometry.py:92-93).  Differentiable
    w.r.t. ``pose`` (analytic backward, one launch)."""
    n, S = pose.shape[1], len(downscales)
    out = _PoseSetupFn.apply(pose, K, K_inv, tuple(downscales), bool(fundamental))
    Kinv = list(out[:S])
    P = [list(out[S + k * S:S + (k + 1) * S]) for k in range(n)]
    F = list(out[S + n * S:]) if fundamental else None
    return Kinv, P, F


class _EpipolarFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, Fm):
        flow, Fm = _dev(flow, "flow"), _dev(Fm, "F")
        B, _, H, W = flow.shape
        out = torch.empty((B, 1, H, W), device=flow.device, dtype=torch.float32)
        with torch.cuda.device_of(flow):
            _call("ugl_epipolar_forward", flow.data_ptr(), Fm.data_ptr(), B, H, W, out.data_ptr(), _stream_ptr())
        ctx.save_for_backward(flow, Fm)
        return out

    @staticmethod
    def backward(ctx, g):
        flow, Fm = ctx.saved_tensors
        B, _, H, W = flow.shape
        gf = torch.empty_like(flow) if ctx.needs_input_grad[0] else None
        gF = torch.empty((B, 3, 3), device=flow.device, dtype=torch.float32)
        ws = _reduce_ws(B, H, W, flow.device)
        with torch.cuda.device_of(flow):
            _call("ugl_epipolar_backward", flow.data_ptr(), Fm.data_ptr(), _dev(g, "g").data_ptr(), B, H, W, _ptr(gf), gF.data_ptr(),
                  ws.data_ptr(), _nbytes(ws), _stream_ptr(), launches=2)
        return gf, (gF if ctx.needs_input_grad[1] else None)


def epipolar_distance(flow: Tensor, Fm: Tensor) -> Tensor:
    """Core of ``compute_epipolar_map`` (model_geometry.py:381-391) given the fundamental matrix F (B,3,3)."""
    return _EpipolarFn.apply(flow, Fm)


def forward_splat(x: Tensor, flow: Tensor, clamp01: bool = False) -> Tensor:
    """EXTENSION — ``transformerFwd`` (undefined in the reference, model_flow.py:36): forward-splat ``x`` (B,C,H,W) by ``flow``
    (B,2,H,W, pixels) with bilinear weights, out-of-range corners dropped.  Constant for autograd, deterministic."""
    x, flow = _dev(x, "x").detach(), _dev(flow, "flow").detach()
    B, Cc, H, W = x.shape
    if tuple(flow.shape) != (B, 2, H, W):
        raise ValueError("forward_splat: flow must be (B,2,H,W), got %s" % (tuple(flow.shape),))
    out = torch.empty_like(x)
    n = int(_cabi.lib().ugl_forward_splat_workspace_bytes(B, Cc, H, W))
    ws = torch.empty((n + 7) // 8, dtype=torch.int64, device=x.device)
    with torch.cuda.device_of(x):
        _call("ugl_forward_splat_ex", x.data_ptr(), flow.data_ptr(), B, Cc, H, W, int(clamp01), out.data_ptr(), ws.data_ptr(), _nbytes(ws),
              _cabi.SCATTER_FORMS[SCATTER_FORM], _stream_ptr(), launches=3)
    return out


def frames_from_u8(frames: Sequence[Tensor], out: Optional[Sequence[Tensor]] = None) -> List[Tensor]:
    """uint8 frames (same shape, up to four) -> fp32 frames in [0,1] in ONE launch: the reference dataset's ``img / 255.0``
    (core/dataset/kitti_prepared.py:89) evaluated on the device, bit-identical to it for every byte value."""
    frames = [_dev_u8(f, "frame") for f in frames]
    if not frames or len(frames) > 4:
        raise ValueError("frames_from_u8 takes 1..4 frames")
    for f in frames:
        if f.shape != frames[0].shape:
            raise ValueError("frames_from_u8: frames must have one shape")
    outs = list(out) if out is not None else [torch.empty(f.shape, dtype=torch.float32, device=f.device) for f in frames]
    src = (C.c_void_p * len(frames))(*[f.data_ptr() for f in frames])
    dst = (C.c_void_p * len(frames))(*[o.data_ptr() for o in outs])
    with torch.cuda.device_of(frames[0]):
        _call("ugl_frames_u8_to_float", src, dst, len(frames), frames[0].numel(), _stream_ptr(), launches=1)
    return outs


def selftest_packed_pairs(device, blocks: int = 64, windows_per_thread: int = 500) -> Tensor:
    """Device self-test of the packed fp32 pair arithmetic of the single-pass kernels (``ugl_selftest_packed_pairs``):
    returns a (2, 14) int64 tensor of bit-mismatch counts against the scalar SSIM functions; all zeros is the contract."""
    out = torch.zeros(2, 14, dtype=torch.int64, device=device)
    with torch.cuda.device(out.device):
        _call("ugl_selftest_packed_pairs", out.data_ptr(), int(blocks), int(windows_per_thread), _stream_ptr(), launches=1)
    return out


# ================================================================================================
# fused reprojection-photometric term (depth / geom modes)
# ================================================================================================
def _depth_photo_args(S, img, area, bil, disp, Kinv, P, ext, loss, den, ws, valid_out=None, tex_out=None, gloss=None, gdisp=None, gP=None,
                      ext_bytes=None, ext_need=(0, 0)):
    a = _cabi.UglDepthPhotoArgs()
    a.batch, a.scales = img[0].shape[0], S
    a.ext_need[0], a.ext_need[1] = int(ext_need[0]), int(ext_need[1])
    for l in range(S):
        if ext_bytes is not None:
            a.ext_bytes[l] = ext_bytes[l].data_ptr()
        a.height[l], a.width[l] = img[l].shape[2], img[l].shape[3]
        a.img[l], a.disp[l], a.Kinv[l] = img[l].data_ptr(), disp[l].data_ptr(), Kinv[l].data_ptr()
        for d in range(2):
            a.src_area[d][l], a.src_bil[d][l], a.P[d][l] = area[d][l].data_ptr(), bil[d][l].data_ptr(), P[d][l].data_ptr()
            if ext is not None:
                a.ext_mask[d][l] = ext[d][l].data_ptr()
            if valid_out is not None:
                a.valid_out[d][l], a.tex_out[d][l] = valid_out[d][l].data_ptr(), tex_out[d][l].data_ptr()
            if gP is not None:
                a.grad_P[d][l] = gP[d][l].data_ptr()
        if gdisp is not None:
            a.grad_disp[l] = gdisp[l].data_ptr()
    a.loss, a.den, a.grad_loss = _ptr(loss), _ptr(den), _ptr(gloss)
    a.workspace, a.workspace_bytes = _ptr(ws), (_nbytes(ws) if ws is not None else 0)
    a.stream = torch.cuda.current_stream().cuda_stream
    return a


DEPTH_PHOTO_SINGLE_PASS = True
"""reprojection-photometric term: True = the forward launch also emits the gradient basis and backward is an element-wise combine
(ugl_depth_photo_forward_grad / _combine); False = backward recomputes the gathers (ugl_depth_photo_backward; saves nothing per pixel)."""


class _DepthPhotoFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, S, has_ext, ext_need, *ts):
        # has_ext: 0 none, 1 float masks (ext_b[S], ext_f[S]), 2 packed uint8 masks (bytes[S]) + ext_need
        nb = S if has_ext == 2 else 0
        ext_bytes = [_dev_u8(t, "ext_bytes") for t in ts[len(ts) - nb:]] if nb else None
        ts = [_dev(t, "input %d" % i) for i, t in enumerate(ts[:len(ts) - nb])]
        # layout: img[S], area_b[S], area_f[S], bil_b[S], bil_f[S], disp[S], Kinv[S], P_b[S], P_f[S], (ext_b[S], ext_f[S])
        g = lambda k: ts[k * S:(k + 1) * S]
        img, area, bil, disp, Kinv, P = g(0), (g(1), g(2)), (g(3), g(4)), g(5), g(6), (g(7), g(8))
        ext = (g(9), g(10)) if has_ext == 1 else None
        B, dev = img[0].shape[0], img[0].device
        for l in range(S):
            h, w = img[l].shape[2:]
            if tuple(disp[l].shape) != (B, 1, h, w) or tuple(area[0][l].shape) != (B, 3, h, w) or tuple(P[0][l].shape) != (B, 3, 4):
                raise ValueError("depth_photo_loss: inconsistent shapes at level %d" % l)
        loss = torch.empty(B, device=dev, dtype=torch.float32)
        den = torch.empty((B, S, 2), device=dev, dtype=torch.float32)
        mk = lambda: [[torch.empty((B, 1) + tuple(img[l].shape[2:]), device=dev, dtype=torch.float32) for l in range(S)] for _ in range(2)]
        valid_out, tex_out = mk(), mk()
        a = _depth_photo_args(S, img, area, bil, disp, Kinv, P, ext, loss, den, None, valid_out, tex_out, ext_bytes=ext_bytes, ext_need=ext_need)
        # a gradient will be asked for: single-pass variant (the forward launch also emits the un-normalised derivatives while
        # its taps are in registers; backward = an element-wise combine instead of a second gather kernel)
        single = DEPTH_PHOTO_SINGLE_PASS and any(ctx.needs_input_grad[3:])
        ctx.single = single
        if single:
            g = _cabi.UglDepthPhotoGradArgs()
            g.photo = a
            basis = [torch.empty((B, 2) + tuple(img[l].shape[2:]), device=dev, dtype=torch.float32) for l in range(S)]
            psum = torch.empty((B, S, 2, 12), device=dev, dtype=torch.float32)
            for l in range(S):
                g.basis[l] = basis[l].data_ptr()
            g.psum = psum.data_ptr()
            n = int(_cabi.lib().ugl_depth_photo_grad_workspace_bytes(C.byref(g)))
            ws = torch.empty((n + 7) // 8, dtype=torch.int64, device=dev)
            g.photo.workspace, g.photo.workspace_bytes = ws.data_ptr(), _nbytes(ws)
            with torch.cuda.device_of(img[0]):
                _call("ugl_depth_photo_forward_grad", C.byref(g), launches=2)
            ctx.save_for_backward(den, psum, *basis, *disp, *P[0], *P[1])
        else:
            n = int(_cabi.lib().ugl_depth_photo_workspace_bytes(C.byref(a)))
            ws = torch.empty((n + 7) // 8, dtype=torch.int64, device=dev)
            a.workspace, a.workspace_bytes = ws.data_ptr(), _nbytes(ws)
            with torch.cuda.device_of(img[0]):
                _call("ugl_depth_photo_forward", C.byref(a), launches=2)
            if any(ctx.needs_input_grad[3:]):            # recompute mode: backward re-runs the gathers from the inputs
                ctx.save_for_backward(den, *ts, *(ext_bytes or []))
        ctx.S, ctx.has_ext, ctx.n_in, ctx.ext_need = S, has_ext, len(ts) + nb, ext_need
        masks = [m for grp in (valid_out, tex_out) for d in grp for m in d]
        ctx.mark_non_differentiable(*masks)
        ctx.set_materialize_grads(False)
        return (loss, *masks)

    @staticmethod
    def backward(ctx, gloss, *unused):
        S = ctx.S
        if gloss is None or not ctx.saved_tensors:
            return (None,) * (3 + ctx.n_in)
        if not ctx.single:
            return _DepthPhotoFn._backward_recompute(ctx, gloss)
        den, psum, *rest = ctx.saved_tensors
        basis, disp, P = rest[0:S], rest[S:2 * S], (rest[2 * S:3 * S], rest[3 * S:4 * S])
        B, dev = disp[0].shape[0], disp[0].device
        gloss = _dev(gloss, "grad_loss")
        gdisp = [torch.empty_like(d) for d in disp]
        gP = [[torch.empty((B, 3, 4), device=dev, dtype=torch.float32) for _ in range(S)] for _ in range(2)]
        g = _cabi.UglDepthPhotoGradArgs()
        a = g.photo
        a.batch, a.scales = B, S
        for l in range(S):
            a.height[l], a.width[l] = disp[l].shape[2], disp[l].shape[3]
            a.grad_disp[l] = gdisp[l].data_ptr()
            g.basis[l] = basis[l].data_ptr()
            for d in range(2):
                a.grad_P[d][l] = gP[d][l].data_ptr()
        a.den, a.grad_loss, g.psum, a.stream = den.data_ptr(), gloss.data_ptr(), psum.data_ptr(), _stream_ptr()
        with torch.cuda.device_of(gloss):
            _call("ugl_depth_photo_combine", C.byref(g), launches=1)
        none = [None] * S
        out = [None, None, None, *none, *none, *none, *none, *none, *gdisp, *none, *gP[0], *gP[1]]
        if ctx.has_ext == 1:
            out += none + none
        elif ctx.has_ext == 2:
            out += none
        return tuple(out)


def _depth_photo_backward_recompute(ctx, gloss):
    den, *ts = ctx.saved_tensors
    S = ctx.S
    ext_bytes = ts[len(ts) - S:] if ctx.has_ext == 2 else None
    g = lambda k: ts[k * S:(k + 1) * S]
    img, area, bil, disp, Kinv, P = g(0), (g(1), g(2)), (g(3), g(4)), g(5), g(6), (g(7), g(8))
    ext = (g(9), g(10)) if ctx.has_ext == 1 else None
    B, dev = img[0].shape[0], img[0].device
    gloss = _dev(gloss, "grad_loss")
    gdisp = [torch.empty_like(d) for d in disp]
    gP = [[torch.empty((B, 3, 4), device=dev, dtype=torch.float32) for _ in range(S)] for _ in range(2)]
    a = _depth_photo_args(S, img, area, bil, disp, Kinv, P, ext, None, den, None, gloss=gloss, gdisp=gdisp, gP=gP,
                          ext_bytes=ext_bytes, ext_need=ctx.ext_need)
    n = int(_cabi.lib().ugl_depth_photo_workspace_bytes(C.byref(a)))
    ws = torch.empty((n + 7) // 8, dtype=torch.int64, device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), _nbytes(ws)
    with torch.cuda.device_of(gloss):
        _call("ugl_depth_photo_backward", C.byref(a), launches=2)
    none = [None] * S
    out = [None, None, None, *none, *none, *none, *none, *none, *gdisp, *none, *gP[0], *gP[1]]
    if ctx.has_ext == 1:
        out += none + none
    elif ctx.has_ext == 2:
        out += none
    return tuple(out)


_DepthPhotoFn._backward_recompute = staticmethod(_depth_photo_backward_recompute)


class _DepthConsisFn(torch.autograd.Function):
    @staticmethod
    def _args(S, disp, ref, Kinv, P, loss=None, gloss=None, gdisp=None, gref=None, gP=None):
        a = _cabi.UglDepthConsisArgs()
        a.batch, a.scales = disp[0].shape[0], S
        for l in range(S):
            a.height[l], a.width[l] = disp[l].shape[2], disp[l].shape[3]
            a.disp[l], a.Kinv[l] = disp[l].data_ptr(), Kinv[l].data_ptr()
            if gdisp is not None:
                a.grad_disp[l] = gdisp[l].data_ptr()
            for d in range(2):
                a.ref_disp[d][l], a.P[d][l] = ref[d][l].data_ptr(), P[d][l].data_ptr()
                if gref is not None:
                    a.grad_ref[d][l], a.grad_P[d][l] = gref[d][l].data_ptr(), gP[d][l].data_ptr()
        a.loss, a.grad_loss = _ptr(loss), _ptr(gloss)
        n = int(_cabi.lib().ugl_depth_consis_workspace_bytes(C.byref(a)))
        a._ws = torch.empty((n + 7) // 8, dtype=torch.int64, device=disp[0].device)
        a.workspace, a.workspace_bytes = a._ws.data_ptr(), _nbytes(a._ws)
        a.stream = torch.cuda.current_stream().cuda_stream
        return a

    @staticmethod
    def forward(ctx, S, *ts):
        ts = [_dev(t, "input %d" % i) for i, t in enumerate(ts)]
        g_ = lambda k: ts[k * S:(k + 1) * S]
        disp, ref, Kinv, P = g_(0), (g_(1), g_(2)), g_(3), (g_(4), g_(5))
        B = disp[0].shape[0]
        for l in range(S):
            if tuple(ref[0][l].shape) != tuple(disp[l].shape) or tuple(ref[1][l].shape) != tuple(disp[l].shape) or tuple(P[0][l].shape) != (B, 3, 4):
                raise ValueError("depth_consis_loss: inconsistent shapes at level %d" % l)
        loss = torch.empty(B, device=disp[0].device, dtype=torch.float32)
        a = _DepthConsisFn._args(S, disp, ref, Kinv, P, loss=loss)
        with torch.cuda.device_of(loss):
            _call("ugl_depth_consis_forward", C.byref(a), launches=2)
        ctx.save_for_backward(*ts)
        ctx.S = S
        return loss

    @staticmethod
    def backward(ctx, gloss):
        ts, S = ctx.saved_tensors, ctx.S
        g_ = lambda k: ts[k * S:(k + 1) * S]
        disp, ref, Kinv, P = g_(0), (g_(1), g_(2)), g_(3), (g_(4), g_(5))
        gloss = _dev(gloss, "grad_loss")
        gdisp = [torch.empty_like(d) for d in disp]
        gref = [[torch.empty_like(d) for d in disp] for _ in range(2)]
        gP = [[torch.empty_like(p) for p in P[d]] for d in range(2)]
        a = _DepthConsisFn._args(S, disp, ref, Kinv, P, gloss=gloss, gdisp=gdisp, gref=gref, gP=gP)
        with torch.cuda.device_of(gloss):
            _call("ugl_depth_consis_backward", C.byref(a), launches=5)
        return (None, *gdisp, *gref[0], *gref[1], *([None] * S), *gP[0], *gP[1])


def depth_consis_loss(disps, ref_disps, Kinv, P):
    """``loss_depth_consis`` of the depth mode (``compute_consis_loss`` on the projected / computed depths of both
    ``reconstruction`` calls, model_depth_texture.py:289-292, 308-309) for all levels in one forward and one backward pass.
    ``ref_disps`` / ``P``: pairs ``(left, right)`` of per-level lists.  Differentiable w.r.t. ``disps``, ``ref_disps`` (deterministic
    scatter) and ``P``."""
    S = len(disps)
    return _DepthConsisFn.apply(S, *disps, *ref_disps[0][:S], *ref_disps[1][:S], *Kinv[:S], *P[0][:S], *P[1][:S])


class _DepthSsimFn(torch.autograd.Function):
    """depth mode with the SSIM term: single-pass tile kernel (reprojection warps) + combine"""

    @staticmethod
    def _args(S, img, area, bil, disp, Kinv, P, loss4, stats, basis, ws, valid_out=None, tex_out=None, gloss=None, gdisp=None, gP=None):
        g = _cabi.UglDepthSsimArgs()
        a = g.photo
        a.batch, a.scales = disp[0].shape[0], S
        for l in range(S):
            a.height[l], a.width[l] = disp[l].shape[2], disp[l].shape[3]
            a.img[l], a.disp[l], a.Kinv[l] = _ptr(img[l]), disp[l].data_ptr(), Kinv[l].data_ptr()
            g.basis[l] = basis[l].data_ptr()
            for d in range(2):
                a.src_area[d][l], a.src_bil[d][l], a.P[d][l] = _ptr(area[d][l]), _ptr(bil[d][l]), P[d][l].data_ptr()
                if valid_out is not None:
                    a.valid_out[d][l], a.tex_out[d][l] = valid_out[d][l].data_ptr(), tex_out[d][l].data_ptr()
                if gP is not None:
                    a.grad_P[d][l] = gP[d][l].data_ptr()
            if gdisp is not None:
                a.grad_disp[l] = gdisp[l].data_ptr()
        g.loss4, g.stats, g.grad_loss4 = _ptr(loss4), stats.data_ptr(), _ptr(gloss)
        a.workspace, a.workspace_bytes = _ptr(ws), (_nbytes(ws) if ws is not None else 0)
        a.stream = torch.cuda.current_stream().cuda_stream
        return g

    @staticmethod
    def forward(ctx, S, *ts):
        ts = [_dev(t, "input %d" % i) for i, t in enumerate(ts)]
        g_ = lambda k: ts[k * S:(k + 1) * S]
        img, area, bil, disp, Kinv, P = g_(0), (g_(1), g_(2)), (g_(3), g_(4)), g_(5), g_(6), (g_(7), g_(8))
        B, dev = img[0].shape[0], img[0].device
        for l in range(S):
            h, w = img[l].shape[2:]
            if tuple(disp[l].shape) != (B, 1, h, w) or tuple(area[0][l].shape) != (B, 3, h, w) or tuple(P[0][l].shape) != (B, 3, 4):
                raise ValueError("depth_ssim_loss: inconsistent shapes at level %d" % l)
        loss4 = torch.empty((4, B), device=dev, dtype=torch.float32)
        stats = torch.empty((B, S, _cabi.GEOM_NSTATS), device=dev, dtype=torch.float32)
        basis = [torch.empty((B, _cabi.DEPTH_BASIS_PLANES) + tuple(d.shape[2:]), device=dev, dtype=torch.float32) for d in disp]
        mk = lambda: [[torch.empty_like(disp[l]) for l in range(S)] for _ in range(2)]
        valid_out, tex_out = mk(), mk()
        g = _DepthSsimFn._args(S, img, area, bil, disp, Kinv, P, loss4, stats, basis, None, valid_out, tex_out)
        n = int(_cabi.lib().ugl_depth_ssim_workspace_bytes(C.byref(g)))
        ws = torch.empty((n + 7) // 8, dtype=torch.int64, device=dev)
        g.photo.workspace, g.photo.workspace_bytes = ws.data_ptr(), _nbytes(ws)
        with torch.cuda.device_of(img[0]):
            _call("ugl_depth_ssim_forward_grad", C.byref(g), launches=2)
        ctx.save_for_backward(stats, *disp, *Kinv, *P[0], *P[1], *basis)
        ctx.S = S
        masks = [m for grp in (valid_out, tex_out) for d in grp for m in d]
        ctx.mark_non_differentiable(*masks)
        ctx.set_materialize_grads(False)
        return (loss4[:2], *masks)

    @staticmethod
    def backward(ctx, gloss, *unused):
        stats, *ts = ctx.saved_tensors
        S = ctx.S
        if gloss is None:
            return (None,) * (1 + 9 * S)
        g_ = lambda k: ts[k * S:(k + 1) * S]
        disp, Kinv, P, basis = g_(0), g_(1), (g_(2), g_(3)), g_(4)
        B, dev = disp[0].shape[0], disp[0].device
        gloss = _dev(gloss, "grad_loss")
        gdisp = [torch.empty_like(d) for d in disp]
        gP = [[torch.empty((B, 3, 4), device=dev, dtype=torch.float32) for _ in range(S)] for _ in range(2)]
        none_l = [None] * S
        g = _DepthSsimFn._args(S, none_l, (none_l, none_l), (none_l, none_l), disp, Kinv, P, None, stats, basis, None, gloss=gloss,
                               gdisp=gdisp, gP=gP)
        n = int(_cabi.lib().ugl_depth_ssim_workspace_bytes(C.byref(g)))
        ws = torch.empty((n + 7) // 8, dtype=torch.int64, device=dev)
        g.photo.workspace, g.photo.workspace_bytes = ws.data_ptr(), _nbytes(ws)
        with torch.cuda.device_of(gloss):
            _call("ugl_depth_ssim_combine", C.byref(g), launches=2)
        return (None, *none_l, *none_l, *none_l, *none_l, *none_l, *gdisp, *none_l, *gP[0], *gP[1])


def depth_ssim_loss(img_pyr, src_area, src_bil, disps, Kinv, P):
    """``loss_depth_pixel`` and ``loss_depth_ssim`` of the depth mode with the SSIM term (model_depth_texture.py:296-301) in the
    single-pass tile kernel: reprojection of both source frames, valid + texture masks, masked L1 and masked 3x3 SSIM, all levels.
    Same argument convention as :func:`depth_photo_loss`.  Returns ``(loss (2,B) = [pixel, ssim], valid[2][S], tex[2][S])``,
    differentiable w.r.t. ``disps`` and ``P``."""
    S = len(disps)
    flat = [*img_pyr[:S], *src_area[0][:S], *src_area[1][:S], *src_bil[0][:S], *src_bil[1][:S], *disps, *Kinv[:S], *P[0][:S], *P[1][:S]]
    out = _DepthSsimFn.apply(S, *flat)
    loss, masks = out[0], out[1:]
    return loss, [list(masks[0:S]), list(masks[S:2 * S])], [list(masks[2 * S:3 * S]), list(masks[3 * S:4 * S])]


def depth_photo_loss(img_pyr, src_area, src_bil, disps, Kinv, P, ext_mask=None, ext_bytes=None, ext_need=(0, 0)):
    """Fused ``loss_depth_pixel`` of the depth / geom modes (reconstruction + texture mask + mask fusion +
    ``compute_photometric_loss`` for both directions and all ``len(disps)`` levels).

    ``src_area`` / ``src_bil`` / ``P`` / ``ext_mask`` are pairs ``(left|bwd, right|fwd)`` of per-level lists.  Returns
    ``(loss (B,), valid[2][S], tex[2][S])``; the loss is differentiable w.r.t. ``disps`` and ``P``.
    ``ext_bytes`` (per-level uint8 maps of :func:`geom_flow_loss`) + ``ext_need`` (bit pattern per direction) replace
    ``ext_mask`` without unpacking the masks to float maps."""
    S = len(disps)
    flat = [*img_pyr[:S], *src_area[0][:S], *src_area[1][:S], *src_bil[0][:S], *src_bil[1][:S], *disps, *Kinv[:S], *P[0][:S], *P[1][:S]]
    if ext_mask is not None and ext_bytes is not None:
        raise ValueError("depth_photo_loss: give ext_mask or ext_bytes, not both")
    if ext_mask is not None:
        flat += [*ext_mask[0][:S], *ext_mask[1][:S]]
    if ext_bytes is not None:
        flat += [*ext_bytes[:S]]
    out = _DepthPhotoFn.apply(S, 1 if ext_mask is not None else (2 if ext_bytes is not None else 0), tuple(ext_need), *flat)
    loss, masks = out[0], out[1:]
    valid = [list(masks[0:S]), list(masks[S:2 * S])]
    tex = [list(masks[2 * S:3 * S]), list(masks[3 * S:4 * S])]
    return loss, valid, tex
