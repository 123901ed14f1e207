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


def emu_geom_flow(img_l, img, img_r, flows_fwd, flows_bwd, disp, Kinv, P_b, P_f, alpha, beta, scales, gloss):
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
    emu().emu_geom_flow_forward_grad(C.byref(g))
    emu().emu_geom_flow_combine(C.byref(g))
    return loss, gf, gb, masks
