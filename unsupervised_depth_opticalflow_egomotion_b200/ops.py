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


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


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
            rc = _cabi.lib().ugl_warp_flow_backward(x.data_ptr(), flow.data_ptr(), gout.data_ptr(), B, Cc, H, W,
                                                    int(ctx.use_mask), _ptr(gflow), _ptr(gx), _ptr(ws), ws_bytes, _stream_ptr())
        _cabi.check(rc, "ugl_warp_flow_backward")
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
# fused flow-mode loss (T0 / flow)
# ================================================================================================
FLOW_LOSS_KEYS = ("loss_flow_pixel", "loss_flow_ssim", "loss_flow_smooth", "loss_flow_consis")


def _flow_args(img_l, img, img_r, ff, fb, scales, loss, stats, ws, gloss=None, gf=None, gb=None) -> _cabi.UglFlowLossArgs:
    a = _cabi.UglFlowLossArgs()
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


class _FlowLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scales: int, L: int, *ts: Tensor):
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
        loss = torch.empty((4, B), device=dev, dtype=torch.float32)
        stats = torch.empty((B, scales, _cabi.FLOW_NSTATS), device=dev, dtype=torch.float32)
        a = _flow_args(img_l, img, img_r, ff, fb, scales, loss, stats, None)
        ws = torch.empty(max(int(_cabi.lib().ugl_flow_loss_workspace_bytes(C.byref(a))) // 4, 1), device=dev, dtype=torch.float32)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel() * 4
        with torch.cuda.device_of(img[0]):
            rc = _cabi.lib().ugl_flow_loss_forward(C.byref(a))
        _cabi.check(rc, "ugl_flow_loss_forward")
        _count(2)
        ctx.save_for_backward(stats, *ts)
        ctx.scales, ctx.L = scales, L
        return loss

    @staticmethod
    def backward(ctx, gloss: Tensor):
        stats, *ts = ctx.saved_tensors
        L, scales = ctx.L, ctx.scales
        img_l, img, img_r, ff, fb = (ts[k * L:(k + 1) * L] for k in range(5))
        gloss = _dev(gloss, "grad_loss")
        gf = [torch.empty_like(ff[l]) for l in range(scales)]
        gb = [torch.empty_like(fb[l]) for l in range(scales)]
        a = _flow_args(img_l, img, img_r, ff, fb, scales, None, stats, None, gloss, gf, gb)
        with torch.cuda.device_of(gloss):
            rc = _cabi.lib().ugl_flow_loss_backward(C.byref(a))
        _cabi.check(rc, "ugl_flow_loss_backward")
        _count(1)
        none_l = [None] * L
        pad = [None] * (L - scales)
        return (None, None, *none_l, *none_l, *none_l, *gf, *pad, *gb, *pad)


def flow_loss_step(img_l_pyr: Sequence[Tensor], img_pyr: Sequence[Tensor], img_r_pyr: Sequence[Tensor],
                   flows_fwd: Sequence[Tensor], flows_bwd: Sequence[Tensor], grad_loss: Tensor,
                   num_scales: Optional[int] = None, out: Optional[dict] = None):
    """Forward + backward of the fused flow-mode loss in one call, without the autograd engine: for
    the training step where the upstream gradient is known up front (``train.py:211-215``:
    ``d total / d loss_k[b] = w_k / B``).  Returns ``(loss (4,B), grads_fwd, grads_bwd)``; pass the
    previous result as ``out`` to reuse its buffers (CUDA-graph friendly: 3 kernel launches, no
    allocation)."""
    L = len(flows_fwd)
    scales = L if num_scales is None else int(num_scales)
    ts = [_dev(t, "input") for t in (*img_l_pyr[:L], *img_pyr[:L], *img_r_pyr[:L], *flows_fwd, *flows_bwd)]
    img_l, img, img_r, ff, fb = (ts[k * L:(k + 1) * L] for k in range(5))
    gloss = _dev(grad_loss, "grad_loss")
    B, dev = img[0].shape[0], img[0].device
    if out is None:
        out = {"loss": torch.empty((4, B), device=dev, dtype=torch.float32),
               "stats": torch.empty((B, scales, _cabi.FLOW_NSTATS), device=dev, dtype=torch.float32),
               "gf": [torch.empty_like(ff[l]) for l in range(scales)], "gb": [torch.empty_like(fb[l]) for l in range(scales)]}
        a = _flow_args(img_l, img, img_r, ff, fb, scales, out["loss"], out["stats"], None)
        out["ws"] = torch.empty(max(int(_cabi.lib().ugl_flow_loss_workspace_bytes(C.byref(a))) // 4, 1), device=dev,
                                dtype=torch.float32)
    a = _flow_args(img_l, img, img_r, ff, fb, scales, out["loss"], out["stats"], out["ws"], gloss, out["gf"], out["gb"])
    with torch.cuda.device_of(img[0]):
        _cabi.check(_cabi.lib().ugl_flow_loss_forward(C.byref(a)), "ugl_flow_loss_forward")
        _cabi.check(_cabi.lib().ugl_flow_loss_backward(C.byref(a)), "ugl_flow_loss_backward")
    _count(3)
    return out


def flow_loss(img_l_pyr: Sequence[Tensor], img_pyr: Sequence[Tensor], img_r_pyr: Sequence[Tensor],
              flows_fwd: Sequence[Tensor], flows_bwd: Sequence[Tensor], num_scales: Optional[int] = None,
              as_matrix: bool = False):
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
    loss = _FlowLossFn.apply(scales, L, *img_l_pyr[:L], *img_pyr[:L], *img_r_pyr[:L], *flows_fwd, *flows_bwd)
    if as_matrix:
        return loss
    return {k: loss[i] for i, k in enumerate(FLOW_LOSS_KEYS)}
