"""One ``torch.autograd.Function`` per mode assembly (T0): the fused loss bodies of ``Model_geometry.forward``
(model_geometry.py:777-951) and ``Model_depth.forward`` (model_depth.py:281-335, model_depth_texture.py:296-311) as ONE autograd
node each.

The per-term kernels are the ones behind ``ops.geom_flow_loss`` / ``depth_photo_loss`` / ``geom_rigid_terms`` /
``disp_smooth_multi`` / ``depth_ssim_loss`` / ``depth_consis_loss`` / ``pose_setup``; here their ``forward`` / ``backward`` halves are
run by hand, outside the autograd engine, so that what the engine would do between them with library kernels -- one ``add`` launch
per tensor that feeds several terms (disparities, level-0 flows, K[R|t]), zero fills, the stack / mean / multiply / sum of
``train.py:211-214`` -- becomes one multi-tensor accumulate launch (``ugl_accumulate_multi``) and one weighted-total launch
(``ugl_weighted_total_*``).  Same kernels, same numbers: tests compare this path with the op-by-op composition bit for bit.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _cabi, ops


class _Ctx:
    """Stands in for autograd's ``ctx`` when a Function's static ``forward`` / ``backward`` is called directly."""

    def __init__(self, needs: Sequence[bool]):
        self.needs_input_grad = tuple(bool(n) for n in needs)
        self.saved_tensors: Tuple[Tensor, ...] = ()

    def save_for_backward(self, *ts):
        self.saved_tensors = tuple(ts)

    def mark_non_differentiable(self, *ts):
        pass

    def set_materialize_grads(self, flag):
        pass


def _run(fn, args, grad_from: int, need: bool, **ctx_attrs):
    """``fn.forward`` with a stand-in ctx: inputs from position ``grad_from`` on count as requiring grad iff ``need``; ``ctx_attrs`` are
    set on the ctx before the call (``step_gloss``: the training-step form of ``ops._GeomFlowLossFn``)."""
    ctx = _Ctx([False] * grad_from + [need] * (len(args) - grad_from))
    for k, v in ctx_attrs.items():
        setattr(ctx, k, v)
    with torch.no_grad():
        out = fn.forward(ctx, *args)
    return ctx, out


class _Side:
    """Independent branches of a mode step on side streams: ``with _Side(k):`` runs its body on the device's k-th side stream, ordered
    after everything issued so far on the current stream; ``_Side.join()`` makes the current stream wait for every side stream used
    since the last join.  Fork / join by stream waits only (no host synchronisation): captures into a CUDA graph as parallel branches.
    Tensors allocated inside a branch are only used on the main stream after the join, and every later use of a side stream starts
    with a wait on the main stream, so the caching allocator's per-stream reuse stays ordered.  ``UGL_MODE_STREAMS=0`` disables it
    (everything on the current stream, as before)."""
    _streams: Dict[Tuple[int, int], "torch.cuda.Stream"] = {}
    _open: Dict[int, "torch.cuda.Stream"] = {}
    enabled = os.environ.get("UGL_MODE_STREAMS", "1") != "0"

    def __init__(self, k: int):
        self.k = k

    def __enter__(self):
        if not _Side.enabled:
            self.ctx = None
            return self
        main = torch.cuda.current_stream()
        key = (main.device.index if main.device.index is not None else torch.cuda.current_device(), self.k)
        st = _Side._streams.get(key)
        if st is None:
            st = _Side._streams[key] = torch.cuda.Stream(device=main.device)
        st.wait_stream(main)
        _Side._open[self.k] = st
        self.ctx = torch.cuda.stream(st)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False

    @staticmethod
    def join(*ks: int):
        """the current stream waits for side streams ``ks`` (default: every one used since its last join)"""
        main = torch.cuda.current_stream()
        for k in (ks or tuple(_Side._open)):
            st = _Side._open.pop(k, None)
            if st is not None:
                main.wait_stream(st)


def _accumulate(pairs: List[Tuple[Tensor, Tensor]]) -> None:
    """dst += src for every (dst, src) pair, one launch (ugl_accumulate_multi).  A destination may appear in several pairs (up to
    three): its contributions are grouped and added by one CTA row, in the order given."""
    groups: Dict[int, Tuple[Tensor, List[Tensor]]] = {}
    for d, s in pairs:
        if d is None or s is None:
            continue
        if d.shape != s.shape or not d.is_contiguous() or not s.is_contiguous():
            raise ValueError("accumulate: mismatched / non-contiguous gradient tensors")
        groups.setdefault(d.data_ptr(), (d, []))[1].append(s)
    if not groups:
        return
    items = list(groups.values())
    if any(len(srcs) > 3 for _, srcs in items):
        raise ValueError("accumulate: more than three contributions to one tensor")
    n = len(items)
    dst = (C.c_void_p * n)(*[d.data_ptr() for d, _ in items])
    src = (C.c_void_p * (3 * n))(*[(srcs[j].data_ptr() if j < len(srcs) else None) for _, srcs in items for j in range(3)])
    num = (C.c_int64 * n)(*[d.numel() for d, _ in items])
    with torch.cuda.device_of(items[0][0]):
        ops._call("ugl_accumulate_multi", dst, src, num, n, ops._stream_ptr())


def _assemble(rows: List[Tuple[Tensor, int]], B: int, device) -> Tensor:
    """(tensor, n) pairs -> (len(rows), B) matrix, row i = sum of the first n consecutive (B,) rows of tensor i; one launch."""
    n = len(rows)
    for t, _ in rows:
        if not t.is_contiguous() or t.dtype != torch.float32:
            raise ValueError("assemble: rows must be contiguous fp32")
    mat = torch.empty((n, B), device=device, dtype=torch.float32)
    src = (C.c_void_p * n)(*[t.data_ptr() for t, _ in rows])
    ns = (C.c_int32 * n)(*[k for _, k in rows])
    with torch.cuda.device(device):
        ops._call("ugl_assemble_rows", src, ns, n, B, mat.data_ptr(), ops._stream_ptr())
    return mat


class _WeightedTotalFn(torch.autograd.Function):
    """``sum_k w_k * mean_b loss[k][b]`` (train.py:211-214) over a (K,B) matrix: one launch forward, one backward."""

    @staticmethod
    def forward(ctx, loss: Tensor, weights: Tensor):
        loss = ops._dev(loss, "loss matrix")
        out = torch.empty((), device=loss.device, dtype=torch.float32)
        with torch.cuda.device_of(loss):
            ops._call("ugl_weighted_total_forward", loss.data_ptr(), weights.data_ptr(), loss.shape[0], loss.shape[1], out.data_ptr(),
                      ops._stream_ptr())
        ctx.save_for_backward(weights)
        ctx.shape = tuple(loss.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        weights, = ctx.saved_tensors
        K, B = ctx.shape
        g = g.contiguous()
        grad = torch.empty((K, B), device=g.device, dtype=torch.float32)
        with torch.cuda.device_of(g):
            ops._call("ugl_weighted_total_backward", g.data_ptr(), weights.data_ptr(), K, B, grad.data_ptr(), ops._stream_ptr())
        return grad, None


class LossPack(dict):
    """The loss dict of a fused mode step: ``(B,)`` rows of ONE (K,B) matrix (``matrix``, row order ``keys``) plus the reference's
    constant ``zeros([2])`` placeholders.  ``losses.total_loss`` recognises it and evaluates ``sum_k w_k mean(loss_k)`` on the
    matrix in one launch."""

    def __init__(self, matrix: Tensor, keys: Sequence[str], placeholders: Dict[str, Tensor], step_weights: Optional[Dict[str, float]] = None):
        super().__init__({k: matrix[i] for i, k in enumerate(keys)})
        self.update(placeholders)
        self.matrix, self.keys_live = matrix, tuple(keys)
        # training-step form: part of the backward was evaluated in the forward for these weights; total() must be called with the same
        self.step_weights = None if step_weights is None else {k: float(step_weights[k]) for k in keys}

    _wcache: Dict[tuple, Tensor] = {}

    def total(self, weights: Dict[str, float]) -> Tensor:
        if self.step_weights is not None and any(float(weights[k]) != w for k, w in self.step_weights.items()):
            raise ValueError("LossPack.total: this pack was produced with step_weights=%r; the weighted total must use the same weights"
                             % (self.step_weights,))
        key = (self.keys_live, tuple(float(weights[k]) for k in self.keys_live), str(self.matrix.device))
        w = LossPack._wcache.get(key)
        if w is None:       # built once per (keys, weights, device): no host-to-device copy inside a captured step
            w = LossPack._wcache[key] = torch.tensor(key[1], dtype=torch.float32, device=self.matrix.device)
        return _WeightedTotalFn.apply(self.matrix, w)


GEOM_KEYS = ("loss_depth_pixel", "loss_depth_smooth", "loss_flow_pixel", "loss_flow_ssim", "loss_flow_smooth", "loss_flow_consis",
             "loss_depth_flow_consis", "loss_epipolar")


def step_grad_matrix(keys: Sequence[str], weights: Dict[str, float], B: int, device) -> Tensor:
    """(K,B) upstream gradient of ``sum_k w_k mean_b loss_k[b]`` (train.py:211-215): rows ``fl(w_k / B)``, what ``_WeightedTotalFn.backward``
    produces for an upstream gradient of 1.  Built once per (keys, weights, B, device)."""
    key = ("step", tuple(keys), tuple(float(weights[k]) for k in keys), int(B), str(device))
    m = LossPack._wcache.get(key)
    if m is None:
        w = torch.tensor(key[2], dtype=torch.float32)
        m = LossPack._wcache[key] = (w / float(B)).view(-1, 1).repeat(1, B).contiguous().to(device)
    return m


class _GeomStepFn(torch.autograd.Function):
    """inputs: S, L, alpha, beta, step_gmat ((8,B) upstream gradient known in advance, or None), then img_l, img, img_r, ff[L], fb[L],
    disp[S], disp_l[S], disp_r[S], pose, K, K_inv.
    outputs: loss matrix (8,B) in GEOM_KEYS order; then (non-differentiable) mask bytes[S], val_l[S], val_r[S], tex_b[S], tex_f[S],
    F_bwd, F_fwd."""

    @staticmethod
    def forward(ctx, S: int, L: int, alpha: float, beta: float, step_gmat: Optional[Tensor], *ts: Tensor):
        img_l, img, img_r = ts[0:3]
        ff, fb = list(ts[3:3 + L]), list(ts[3 + L:3 + 2 * L])
        o = 3 + 2 * L
        disp, disp_l, disp_r = list(ts[o:o + S]), list(ts[o + S:o + 2 * S]), list(ts[o + 2 * S:o + 3 * S])
        pose, K, K_inv = ts[o + 3 * S:o + 3 * S + 3]
        needs = ctx.needs_input_grad[5:]
        need_flow, need_disp = any(needs[3:3 + 2 * L]), any(needs[o:o + 3 * S])
        need_pose = needs[o + 3 * S]
        need_any = need_flow or need_disp or need_pose
        H = img.shape[2]
        downs = tuple(H / d.shape[2] for d in disp)
        with _Side(0):      # the disparity smoothness needs nothing the other branches produce
            c_smooth, sm3 = _run(ops._DispSmoothMultiFn, (3, S, img, img_l, img_r, *disp, *disp_l, *disp_r), 2 + 3, need_disp)
        with _Side(1):
            c_pose, out = _run(ops._PoseSetupFn, (pose, K, K_inv, downs, True), 0, need_pose)
        Kinv, P_b, P_f, Fm = list(out[:S]), list(out[S:2 * S]), list(out[2 * S:3 * S]), list(out[3 * S:])
        pyr = ops.image_pyramids((img, img_l, img_r), S, ("bilinear", ("bilinear", "area"), ("bilinear", "area")))
        _Side.join(1)
        pc, pl, pr = (d["bilinear"] for d in pyr)
        area = (pyr[1]["area"], pyr[2]["area"])
        # training-step form: the flow branch's gradients come out of its forward launches (rows 2..5 of the upstream gradient)
        def rigid_terms(masks):      # the level-0 rigid terms and the reprojection term both start from the flow branch's mask bytes
            return _run(ops._GeomRigidFn, (fb[0], ff[0], disp[0], masks[0], Kinv[0], P_b[0], P_f[0], Fm[0], Fm[1],
                                           (ops.MASK_ALL_BWD, ops.MASK_ALL_FWD)), 0, need_any)

        def reprojection(masks):
            return _run(ops._DepthPhotoFn, (S, 2, (ops.MASK_ALL_BWD, ops.MASK_ALL_FWD), *pc[:S], *area[0][:S], *area[1][:S], *pl[:S],
                                            *pr[:S], *disp, *Kinv, *P_b, *P_f, *masks), 3, need_disp or need_pose)

        early = {}

        def after_photo(masks):      # training-step form: called between the photometry kernel and the rest of the flow branch
            with _Side(1):
                early["rigid"] = rigid_terms(masks)
            with _Side(2):
                early["photo"] = reprojection(masks)

        step = {} if step_gmat is None or not need_flow else {"step_gloss": step_gmat[2:6]}
        if step and _Side.enabled:
            step["after_photo"] = after_photo
        c_flow, out = _run(ops._GeomFlowLossFn, (S, S, float(alpha), float(beta), *pl[:S], *pc[:S], *pr[:S], *ff[:S], *fb[:S], *disp, *Kinv,
                                                  *P_b, *P_f), 4 + 3 * S, need_flow, **step)
        flow4, mbytes = out[0], list(out[1:])
        if early:
            (c_rigid, out_r), (c_photo, out) = early["rigid"], early["photo"]
        else:
            with _Side(1):
                c_rigid, out_r = rigid_terms(mbytes)
            c_photo, out = reprojection(mbytes)
        depth_pixel, pmasks = out[0], out[1:]
        dfc, epi = out_r
        _Side.join()
        B = img.shape[0]
        mat = _assemble([(depth_pixel, 1), (sm3, 3), (flow4[0], 1), (flow4[1], 1), (flow4[2], 1), (flow4[3], 1), (dfc, 1), (epi, 1)], B, img.device)
        ctx.sub = (c_pose, c_flow, c_photo, c_rigid, c_smooth)
        ctx.S, ctx.L = S, L
        ctx.flags = (need_flow, need_disp, need_pose)
        nd = [*mbytes, *pmasks, *Fm]
        ctx.mark_non_differentiable(*nd)
        ctx.set_materialize_grads(False)
        return (mat, *nd)

    @staticmethod
    def backward(ctx, gmat, *unused):
        S, L = ctx.S, ctx.L
        n_in = 3 + 2 * L + 3 * S + 3
        if gmat is None:
            return (None,) * (5 + n_in)
        c_pose, c_flow, c_photo, c_rigid, c_smooth = ctx.sub
        need_flow, need_disp, need_pose = ctx.flags
        gmat = gmat.contiguous()
        with torch.no_grad():
            with _Side(0):
                gsm = ops._DispSmoothMultiFn.backward(c_smooth, gmat[1])
            with _Side(1):
                g_rigid = ops._GeomRigidFn.backward(c_rigid, gmat[6].contiguous(), gmat[7].contiguous())
            g_photo = ops._DepthPhotoFn.backward(c_photo, gmat[0].contiguous())
            g_flow = ops._GeomFlowLossFn.backward(c_flow, gmat[2:6].contiguous())
            _Side.join()
            # unpack by the sub-Functions' own input layouts
            gfb0, gff0, gd0, _, _, gPb0, gPf0, gFb, gFf, _ = g_rigid
            gdisp = list(g_photo[3 + 5 * S:3 + 6 * S])
            gPb, gPf = list(g_photo[3 + 7 * S:3 + 8 * S]), list(g_photo[3 + 8 * S:3 + 9 * S])
            gf, gb = list(g_flow[4 + 3 * S:4 + 4 * S]), list(g_flow[4 + 4 * S:4 + 5 * S])
            gsm_c, gsm_l, gsm_r = list(gsm[5:5 + S]), list(gsm[5 + S:5 + 2 * S]), list(gsm[5 + 2 * S:5 + 3 * S])
            pairs = [(gf[0], gff0), (gb[0], gfb0), (gdisp[0], gd0), (gPb[0], gPb0), (gPf[0], gPf0)]
            pairs += [(gdisp[l], gsm_c[l]) for l in range(S)]
            _accumulate(pairs)
            gpose = None
            if need_pose:
                gpose = ops._PoseSetupFn.backward(c_pose, *([None] * S), *gPb, *gPf, gFb, gFf)[0]
        pad = [None] * (L - S)
        if not need_flow:
            gf, gb = [None] * S, [None] * S
        if not need_disp:
            gdisp, gsm_l, gsm_r = [None] * S, [None] * S, [None] * S
        return (None, None, None, None, None, None, None, None, *gf, *pad, *gb, *pad, *gdisp, *gsm_l, *gsm_r, gpose, None, None)


def geom_step(S: int, alpha: float, beta: float, img_l, img, img_r, flows_fwd, flows_bwd, disp, disp_l, disp_r, pose, K, K_inv,
              step_weights: Optional[Dict[str, float]] = None):
    """-> (LossPack-ready (8,B) matrix, mask bytes[S], (val_l, val_r), (tex_b, tex_f), (F_bwd, F_fwd)).
    ``step_weights``: the weights of the weighted total that will be taken of these losses (train.py:211-215), declared up front: the flow
    branch then runs as a fused training step (``ugl_geom_flow_step``: no basis planes, no combine launch).  The caller must form the total
    with the same weights (``LossPack.total`` checks) and back-propagate an upstream gradient of 1."""
    L = len(flows_fwd)
    gmat = None if step_weights is None else step_grad_matrix(GEOM_KEYS, step_weights, img.shape[0], img.device)
    out = _GeomStepFn.apply(S, L, float(alpha), float(beta), gmat, img_l, img, img_r, *flows_fwd, *flows_bwd, *disp[:S], *disp_l[:S], *disp_r[:S],
                            pose, K, K_inv)
    mat, rest = out[0], out[1:]
    mbytes = list(rest[0:S])
    val = (list(rest[S:2 * S]), list(rest[2 * S:3 * S]))
    tex = (list(rest[3 * S:4 * S]), list(rest[4 * S:5 * S]))
    return mat, mbytes, val, tex, (rest[5 * S], rest[5 * S + 1])


DEPTH_KEYS = {"live": ("loss_depth_pixel", "loss_depth_smooth"),
              "ssim": ("loss_depth_pixel", "loss_depth_ssim", "loss_depth_smooth"),
              "texture": ("loss_depth_pixel", "loss_depth_ssim", "loss_depth_smooth", "loss_depth_consis")}


class _DepthStepFn(torch.autograd.Function):
    """inputs: S, variant, then img_l, img, img_r, disp[S], disp_l[S], disp_r[S], pose, K.
    outputs: loss matrix (len(DEPTH_KEYS[variant]), B); then valid_l[S], valid_r[S], tex_b[S], tex_f[S] (non-differentiable)."""

    @staticmethod
    def forward(ctx, S: int, variant: str, *ts: Tensor):
        img_l, img, img_r = ts[0:3]
        disp, disp_l, disp_r = list(ts[3:3 + S]), list(ts[3 + S:3 + 2 * S]), list(ts[3 + 2 * S:3 + 3 * S])
        pose, K = ts[3 + 3 * S], ts[4 + 3 * S]
        needs = ctx.needs_input_grad[2:]
        need_disp, need_pose = any(needs[3:3 + 3 * S]), needs[3 + 3 * S]
        H = img.shape[2]
        downs = tuple(H / d.shape[2] for d in disp)
        with _Side(0):      # the disparity smoothness needs nothing the other terms produce
            c_smooth, sm3 = _run(ops._DispSmoothMultiFn, (3, S, img, img_l, img_r, *disp, *disp_l, *disp_r), 2 + 3, need_disp)
        with _Side(1):
            c_pose, out = _run(ops._PoseSetupFn, (pose, K, None, downs, False), 0, need_pose)
        Kinv, P_b, P_f = list(out[:S]), list(out[S:2 * S]), list(out[2 * S:3 * S])
        pyr = ops.image_pyramids((img, img_l, img_r), S, ("bilinear", ("bilinear", "area"), ("bilinear", "area")))
        _Side.join(1)
        pc, pl, pr = (d["bilinear"] for d in pyr)
        area = (pyr[1]["area"], pyr[2]["area"])
        flat = (*pc[:S], *area[0][:S], *area[1][:S], *pl[:S], *pr[:S], *disp, *Kinv, *P_b, *P_f)
        B = img.shape[0]
        c_consis, cons, rows = None, None, []
        if variant == "texture":      # the depth-consistency term only needs the disparities and the matrices: its own branch
            with _Side(1):
                c_consis, cons = _run(ops._DepthConsisFn, (S, *disp, *disp_l, *disp_r, *Kinv, *P_b, *P_f), 1, need_disp or need_pose)
        if variant == "live":
            c_photo, out = _run(ops._DepthPhotoFn, (S, 0, (0, 0), *flat), 3, need_disp or need_pose)
            rows.append((out[0], 1))
        else:
            c_photo, out = _run(ops._DepthSsimFn, (S, *flat), 1, need_disp or need_pose)
            rows += [(out[0][0], 1), (out[0][1], 1)]
        masks = out[1:]
        rows.append((sm3, 3))
        if variant == "texture":
            rows.append((cons, 1))
        _Side.join()
        mat = _assemble(rows, B, img.device)
        ctx.sub = (c_pose, c_photo, c_consis, c_smooth)
        ctx.S, ctx.variant, ctx.flags = S, variant, (need_disp, need_pose)
        ctx.mark_non_differentiable(*masks)
        ctx.set_materialize_grads(False)
        return (mat, *masks)

    @staticmethod
    def backward(ctx, gmat, *unused):
        S, variant = ctx.S, ctx.variant
        n_in = 3 + 3 * S + 2
        if gmat is None:
            return (None,) * (2 + n_in)
        c_pose, c_photo, c_consis, c_smooth = ctx.sub
        need_disp, need_pose = ctx.flags
        keys = DEPTH_KEYS[variant]
        gmat = gmat.contiguous()
        with torch.no_grad():
            with _Side(0):
                gsm = ops._DispSmoothMultiFn.backward(c_smooth, gmat[keys.index("loss_depth_smooth")])
            g_c = None
            if c_consis is not None:
                with _Side(1):
                    g_c = ops._DepthConsisFn.backward(c_consis, gmat[3].contiguous())
            if variant == "live":
                g_photo = ops._DepthPhotoFn.backward(c_photo, gmat[0].contiguous())
                base = 3
            else:
                g_photo = ops._DepthSsimFn.backward(c_photo, gmat[0:2].contiguous())
                base = 1
            gdisp = list(g_photo[base + 5 * S:base + 6 * S])
            gPb, gPf = list(g_photo[base + 7 * S:base + 8 * S]), list(g_photo[base + 8 * S:base + 9 * S])
            gsm_c, gsm_l, gsm_r = list(gsm[5:5 + S]), list(gsm[5 + S:5 + 2 * S]), list(gsm[5 + 2 * S:5 + 3 * S])
            pairs = [(gdisp[l], gsm_c[l]) for l in range(S)]
            _Side.join()
            if g_c is not None:
                # (None, gdisp[S], gref_l[S], gref_r[S], None[S], gP_b[S], gP_f[S])
                pairs += [(gdisp[l], g_c[1 + l]) for l in range(S)]
                pairs += [(gsm_l[l], g_c[1 + S + l]) for l in range(S)] + [(gsm_r[l], g_c[1 + 2 * S + l]) for l in range(S)]
                pairs += [(gPb[l], g_c[1 + 4 * S + l]) for l in range(S)] + [(gPf[l], g_c[1 + 5 * S + l]) for l in range(S)]
            _accumulate(pairs)
            gpose = ops._PoseSetupFn.backward(c_pose, *([None] * S), *gPb, *gPf)[0] if need_pose else None
        if not need_disp:
            gdisp, gsm_l, gsm_r = [None] * S, [None] * S, [None] * S
        return (None, None, None, None, None, *gdisp, *gsm_l, *gsm_r, gpose, None)


def depth_step(S: int, variant: str, img_l, img, img_r, disp, disp_l, disp_r, pose, K):
    """-> ((K,B) loss matrix in DEPTH_KEYS[variant] order, (valid_l, valid_r), (tex_b, tex_f))"""
    out = _DepthStepFn.apply(S, variant, img_l, img, img_r, *disp[:S], *disp_l[:S], *disp_r[:S], pose, K)
    mat, rest = out[0], out[1:]
    return mat, (list(rest[0:S]), list(rest[S:2 * S])), (list(rest[2 * S:3 * S]), list(rest[3 * S:4 * S]))
