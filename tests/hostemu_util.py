"""TEST INFRASTRUCTURE — builds and drives tests/hostemu/libhostemu.so (the host emulator of the
CUDA kernels' tile logic) with the same ctypes structs as the real library.  CPU tensors only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import torch

from unsupervised_depth_opticalflow_egomotion_b200 import _cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "hostemu", "hostemu.cpp")
_LIB = os.path.join(_HERE, "hostemu", "libhostemu.so")
_CSRC = os.path.join(_HERE, "..", "unsupervised_depth_opticalflow_egomotion_b200", "csrc")

_emu = None


def emu():
    global _emu
    if _emu is None:
        deps = [_SRC] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
        if not os.path.exists(_LIB) or any(os.path.getmtime(d) > os.path.getmtime(_LIB) for d in deps):
            subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", _LIB, _SRC])
        _emu = C.CDLL(_LIB)
    return _emu


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def flow_loss_args(img_l, img, img_r, flows_fwd, flows_bwd, scales, loss, stats, gloss=None, gf=None, gb=None):
    a = _cabi.UglFlowLossArgs()
    a.batch, a.levels, a.scales = img[0].shape[0], len(img), scales
    for l in range(len(img)):
        a.height[l], a.width[l] = img[l].shape[2], img[l].shape[3]
        a.img_l[l], a.img[l], a.img_r[l] = img_l[l].data_ptr(), img[l].data_ptr(), img_r[l].data_ptr()
        a.flow_fwd[l], a.flow_bwd[l] = flows_fwd[l].data_ptr(), flows_bwd[l].data_ptr()
        if gf is not None and l < scales:
            a.grad_flow_fwd[l], a.grad_flow_bwd[l] = gf[l].data_ptr(), gb[l].data_ptr()
    a.loss, a.stats = loss.data_ptr(), stats.data_ptr()
    a.grad_loss = gloss.data_ptr() if gloss is not None else None
    return a


def emu_flow_loss(img_l, img, img_r, flows_fwd, flows_bwd, scales, gloss):
    """returns loss (4,B), grads fwd list, grads bwd list — all via the host emulator"""
    B = img[0].shape[0]
    loss = torch.zeros(4, B)
    stats = torch.zeros(B, scales, _cabi.FLOW_NSTATS)
    gf = [torch.zeros_like(f) for f in flows_fwd[:scales]]
    gb = [torch.zeros_like(f) for f in flows_bwd[:scales]]
    a = flow_loss_args(img_l, img, img_r, flows_fwd, flows_bwd, scales, loss, stats, gloss, gf, gb)
    emu().emu_flow_loss_forward(C.byref(a))
    emu().emu_flow_loss_backward(C.byref(a))
    return loss, gf, gb, stats


def emu_flow_loss_single_pass(img_l, img, img_r, flows_fwd, flows_bwd, scales, gloss):
    """single-pass mode (forward_grad + combine) via the host emulator"""
    B = img[0].shape[0]
    loss = torch.zeros(4, B)
    stats = torch.zeros(B, scales, _cabi.FLOW_NSTATS)
    gf = [torch.zeros_like(f) for f in flows_fwd[:scales]]
    gb = [torch.zeros_like(f) for f in flows_bwd[:scales]]
    basis = [torch.zeros(B, _cabi.FLOW_BASIS_PLANES, f.shape[2], f.shape[3]) for f in flows_fwd[:scales]]
    a = flow_loss_args(img_l, img, img_r, flows_fwd, flows_bwd, scales, loss, stats, gloss, gf, gb)
    for l in range(scales):
        a.basis[l] = basis[l].data_ptr()
    emu().emu_flow_loss_forward_grad(C.byref(a))
    emu().emu_flow_loss_combine(C.byref(a))
    return loss, gf, gb, stats


def emu_flow_loss_split(img_l, img, img_r, flows_fwd, flows_bwd, scales, gloss):
    """split form (photometry pixels -> photometry planes -> stencil tiles) + combine via the host emulator"""
    B = img[0].shape[0]
    loss = torch.zeros(4, B)
    stats = torch.zeros(B, scales, _cabi.FLOW_NSTATS)
    gf = [torch.zeros_like(f) for f in flows_fwd[:scales]]
    gb = [torch.zeros_like(f) for f in flows_bwd[:scales]]
    basis = [torch.zeros(B, _cabi.FLOW_BASIS_PLANES, f.shape[2], f.shape[3]) for f in flows_fwd[:scales]]
    a = flow_loss_args(img_l, img, img_r, flows_fwd, flows_bwd, scales, loss, stats, gloss, gf, gb)
    for l in range(scales):
        a.basis[l] = basis[l].data_ptr()
    emu().emu_flow_loss_split_forward_grad(C.byref(a))
    emu().emu_flow_loss_combine(C.byref(a))
    return loss, gf, gb, stats


def emu_flow_loss_step(img_l, img, img_r, flows_fwd, flows_bwd, scales, gloss):
    """the fused training step (stencil tiles write the flow gradients) via the host emulator"""
    B = img[0].shape[0]
    loss = torch.zeros(4, B)
    stats = torch.zeros(B, scales, _cabi.FLOW_NSTATS)
    gf = [torch.zeros_like(f) for f in flows_fwd[:scales]]
    gb = [torch.zeros_like(f) for f in flows_bwd[:scales]]
    a = flow_loss_args(img_l, img, img_r, flows_fwd, flows_bwd, scales, loss, stats, gloss, gf, gb)
    emu().emu_flow_loss_step(C.byref(a))
    return loss, gf, gb, stats


def emu_geom_flow(img_l, img, img_r, flows_fwd, flows_bwd, disp, Kinv, P_b, P_f, alpha, beta, scales, gloss, split=False):
    """geom-mode flow branch (forward_grad + combine) via the host emulator -> loss (4,B), grads, mask bytes"""
    B = img[0].shape[0]
    loss = torch.zeros(4, B)
    stats = torch.zeros(B, scales, _cabi.GEOM_NSTATS)
    gf = [torch.zeros_like(f) for f in flows_fwd[:scales]]
    gb = [torch.zeros_like(f) for f in flows_bwd[:scales]]
    basis = [torch.zeros(B, _cabi.FLOW_BASIS_PLANES, f.shape[2], f.shape[3]) for f in flows_fwd[:scales]]
    masks = [torch.zeros(B, f.shape[2], f.shape[3], dtype=torch.uint8) for f in flows_fwd[:scales]]
    g = _cabi.UglGeomFlowArgs()
    g.flow = flow_loss_args(img_l, img, img_r, flows_fwd, flows_bwd, scales, loss, stats, gloss, gf, gb)
    for l in range(scales):
        g.flow.basis[l] = basis[l].data_ptr()
        g.disp[l], g.Kinv[l], g.P_bwd[l], g.P_fwd[l] = disp[l].data_ptr(), Kinv[l].data_ptr(), P_b[l].data_ptr(), P_f[l].data_ptr()
        g.mask_bytes[l] = masks[l].data_ptr()
    g.alpha, g.beta = alpha, beta
    if split == "step":       # the fused training step: no basis planes, no combine
        emu().emu_geom_flow_step(C.byref(g))
        return loss, gf, gb, masks
    (emu().emu_geom_flow_split_forward_grad if split else emu().emu_geom_flow_forward_grad)(C.byref(g))
    emu().emu_geom_flow_combine(C.byref(g))
    return loss, gf, gb, masks


def emu_depth_ssim(img, area, bil, disp, Kinv, P, gloss):
    """depth-mode single-pass kernel (forward_grad + combine) via the host emulator.  area / bil / P: pairs (left, right) of
    per-level lists; gloss (2,B).  Returns loss4 (4,B), grad_disp[S], grad_P[2][S], valid[2][S], tex[2][S]"""
    S, B = len(disp), img[0].shape[0]
    loss4 = torch.zeros(4, B)
    stats = torch.zeros(B, S, _cabi.GEOM_NSTATS)
    basis = [torch.zeros(B, _cabi.DEPTH_BASIS_PLANES, d.shape[2], d.shape[3]) for d in disp]
    gdisp = [torch.zeros_like(d) for d in disp]
    gP = [[torch.zeros(B, 3, 4) for _ in range(S)] for _ in range(2)]
    valid = [[torch.zeros_like(d) for d in disp] for _ in range(2)]
    tex = [[torch.zeros_like(d) for d in disp] for _ in range(2)]
    g = _cabi.UglDepthSsimArgs()
    a = g.photo
    a.batch, a.scales = B, S
    for l in range(S):
        a.height[l], a.width[l] = disp[l].shape[2], disp[l].shape[3]
        a.img[l], a.disp[l], a.Kinv[l], a.grad_disp[l] = img[l].data_ptr(), disp[l].data_ptr(), Kinv[l].data_ptr(), gdisp[l].data_ptr()
        g.basis[l] = basis[l].data_ptr()
        for d in range(2):
            a.src_area[d][l], a.src_bil[d][l], a.P[d][l] = area[d][l].data_ptr(), bil[d][l].data_ptr(), P[d][l].data_ptr()
            a.valid_out[d][l], a.tex_out[d][l], a.grad_P[d][l] = valid[d][l].data_ptr(), tex[d][l].data_ptr(), gP[d][l].data_ptr()
    g.loss4, g.stats, g.grad_loss4 = loss4.data_ptr(), stats.data_ptr(), gloss.data_ptr()
    emu().emu_depth_ssim_forward_grad(C.byref(g))
    emu().emu_depth_ssim_combine(C.byref(g))
    return loss4, gdisp, gP, valid, tex
