"""ORACLE-SIDE (test infrastructure) — runs the UNMODIFIED reference loss functions in the build
container.  It exists only where ``/root/reference`` is mounted (never on the GPU box) and is
used by ``oracle/validate_against_reference.py`` and ``tests/golden/make_golden.py``.

No reference source is copied or edited: the reference modules are imported from where they lie,
with the two shims SURVEY §8(c) lists:
  (1) ``torch.Tensor.get_device`` returns ``.device`` for CPU tensors (the reference uses
      ``get_device()`` as a device handle, e.g. structures/net_utils.py:38,48), and
  (2) loss *methods* are called unbound on a plain namespace carrying the hyper-parameters, so
      no network is constructed (``Model_flow.__init__`` is broken as shipped).
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import torch

REFERENCE_ROOT = os.environ.get("UGL_REFERENCE_ROOT", "/root/reference")
_NETWORKS = os.path.join(REFERENCE_ROOT, "core", "networks")


def available() -> bool:
    return os.path.isdir(_NETWORKS)


_loaded = None


def load():
    """Import the reference hot-path modules; returns a namespace of them."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    warnings.filterwarnings("ignore", message=".*align_corners.*")
    warnings.filterwarnings("ignore", message=".*grid_sample.*")
    orig = torch.Tensor.get_device
    if not getattr(torch.Tensor.get_device, "_ugl_shim", False):
        def get_device(self):
            return self.device if not self.is_cuda else orig(self)
        get_device._ugl_shim = True
        torch.Tensor.get_device = get_device
    if _NETWORKS not in sys.path:
        sys.path.insert(0, _NETWORKS)
    import structures, pytorch_ssim, model_flow, model_geometry, model_depth, model_depth_texture  # noqa
    _loaded = types.SimpleNamespace(
        structures=structures, pytorch_ssim=pytorch_ssim, model_flow=model_flow,
        model_geometry=model_geometry, model_depth=model_depth, model_depth_texture=model_depth_texture)
    return _loaded


class _Bound:
    """Calls a reference class's methods unbound on a hyper-parameter namespace."""

    def __init__(self, cls, **hp):
        object.__setattr__(self, "_cls", cls)
        object.__setattr__(self, "_ns", types.SimpleNamespace(**hp))

    def __getattr__(self, name):
        cls, ns = object.__getattribute__(self, "_cls"), object.__getattribute__(self, "_ns")
        if hasattr(ns, name):
            return getattr(ns, name)
        fn = getattr(cls, name)
        return lambda *a, **k: fn(self, *a, **k)


def flow_model(num_scales: int):
    return _Bound(load().model_flow.Model_flow, num_scales=num_scales)


def geom_model(num_scales: int, alpha: float = 0.01, beta: float = 0.5):
    return _Bound(load().model_geometry.Model_geometry, num_scales=num_scales, flow_consist_alpha=alpha,
                  flow_consist_beta=beta, rigid_thres=0.5, inlier_thres=0.1)


def depth_model(num_scales: int, texture_variant: bool = False):
    ref = load()
    cls = ref.model_depth_texture.Model_depth if texture_variant else ref.model_depth.Model_depth
    return _Bound(cls, num_scales=num_scales)


# ------------------------------------------------------------------------------------------------
# per-mode assembly executed with the reference's own methods (mirrors the bodies of the three
# ``forward`` functions after the networks have produced flows / disparities / poses)
# ------------------------------------------------------------------------------------------------
def reference_flow_mode(t, scales: int):
    """model_flow.py:232-254 with the network outputs supplied by the caller."""
    m = flow_model(scales)
    L = len(t.flows_fwd)
    pl, pc, pr = m.generate_img_pyramid(t.img_l, L), m.generate_img_pyramid(t.img, L), m.generate_img_pyramid(t.img_r, L)
    from_l = m.warp_flow_pyramid(pl, t.flows_bwd)
    from_r = m.warp_flow_pyramid(pr, t.flows_fwd)
    diff_bwd, diff_fwd, w_bwd, w_fwd = m.compute_diff_weight(from_l, pc, from_r)
    loss = {
        "loss_flow_pixel": m.compute_loss_with_mask(diff_fwd, w_fwd) + m.compute_loss_with_mask(diff_bwd, w_bwd),
        "loss_flow_ssim": m.compute_loss_ssim(pc, from_r, w_fwd) + m.compute_loss_ssim(pc, from_l, w_bwd),
        "loss_flow_smooth": m.compute_loss_flow_smooth(t.flows_fwd, pc) + m.compute_loss_flow_smooth(t.flows_bwd, pc),
        "loss_flow_consis": m.compute_loss_flow_consis(t.flows_fwd, t.flows_bwd, w_fwd),
    }
    aux = dict(w_bwd=w_bwd, w_fwd=w_fwd, diff_bwd=diff_bwd, diff_fwd=diff_fwd, from_l=from_l, from_r=from_r)
    return loss, aux


def reference_depth_mode(t, scales: int, texture_variant: bool = False):
    """model_depth.py:281-335 (live) or model_depth_texture.py:262-311 (SSIM + consistency enabled)."""
    m = depth_model(scales, texture_variant)
    pl, pc, pr = (m.generate_img_pyramid(x, scales) for x in (t.img_l, t.img, t.img_r))
    rec_l, val_l, proj_l, comp_l = m.reconstruction(t.img_l, t.K, t.disp, t.disp_l, t.pose[:, 0])
    rec_r, val_r, proj_r, comp_r = m.reconstruction(t.img_r, t.K, t.disp, t.disp_r, t.pose[:, 1])
    if texture_variant:
        loss = {
            "loss_depth_pixel": m.compute_photometric_depth_loss(pc, rec_l, pl, val_l)
                                + m.compute_photometric_depth_loss(pc, rec_r, pr, val_r),
            "loss_depth_ssim": m.compute_ssim_loss(pc, rec_l, val_l) + m.compute_ssim_loss(pc, rec_r, val_r),
            "loss_depth_smooth": m.compute_smooth_loss(t.img, t.disp) + m.compute_smooth_loss(t.img_l, t.disp_l)
                                 + m.compute_smooth_loss(t.img_r, t.disp_r),
            "loss_depth_consis": m.compute_consis_loss(proj_l, comp_l) + m.compute_consis_loss(proj_r, comp_r),
        }
        tex_b = tex_f = None
    else:
        tex_b = m.compute_texture_mask(pc, rec_l, pl)
        tex_f = m.compute_texture_mask(pc, rec_r, pr)
        loss = {
            "loss_depth_pixel": m.compute_photometric_loss(pc, rec_l, m.fusion_mask(val_l, tex_b))
                                + m.compute_photometric_loss(pc, rec_r, m.fusion_mask(val_r, tex_f)),
            "loss_depth_smooth": m.compute_smooth_loss(t.img, t.disp) + m.compute_smooth_loss(t.img_l, t.disp_l)
                                 + m.compute_smooth_loss(t.img_r, t.disp_r),
        }
    aux = dict(valid_l=val_l, valid_r=val_r, tex_b=tex_b, tex_f=tex_f, rec_l=rec_l, rec_r=rec_r)
    return loss, aux


def reference_geom_mode(t, scales: int, alpha: float = 0.01, beta: float = 0.5):
    """model_geometry.py:777-951 with the network outputs supplied by the caller (the zero
    placeholders and the dead sample_match call are omitted)."""
    m = geom_model(scales, alpha, beta)
    K, K_inv = t.K, t.K_inv
    pc, pl, pr = (m.generate_img_pyramid(x, scales) for x in (t.img, t.img_l, t.img_r))
    pb, pf = t.pose[:, 0], t.pose[:, 1]
    rec_l, val_l, _, _ = m.reconstruction(t.img_l, K, t.disp, t.disp_l, pb)
    rec_r, val_r, _, _ = m.reconstruction(t.img_r, K, t.disp, t.disp_r, pf)
    tex_b = m.compute_texture_mask(pc, rec_l, pl)
    tex_f = m.compute_texture_mask(pc, rec_r, pr)
    from_l = m.warp_flow_pyramid(pl, t.flows_bwd)
    from_r = m.warp_flow_pyramid(pr, t.flows_fwd)
    occ_b, occ_f, valid_b, valid_f = m.compute_occ_weight(from_l, pc, from_r)
    fd_b, dyn_b, _ = m.compute_dynamic_mask(K, t.disp, pb, t.flows_bwd)
    fd_f, dyn_f, _ = m.compute_dynamic_mask(K, t.disp, pf, t.flows_fwd)
    dist_b = m.compute_epipolar_map(pb, t.flows_bwd[0], K, K_inv)
    dist_f = m.compute_epipolar_map(pf, t.flows_fwd[0], K, K_inv)
    fwd_mask = m.fusion_mask(valid_f, occ_f, dyn_f)
    bwd_mask = m.fusion_mask(valid_b, occ_b, dyn_b)
    fwd_mask_tex = m.fusion_mask_2item(fwd_mask, tex_f)
    bwd_mask_tex = m.fusion_mask_2item(bwd_mask, tex_b)
    fwd_vo = m.fusion_mask_2item(valid_f, occ_f)
    bwd_vo = m.fusion_mask_2item(valid_b, occ_b)
    fwd_vo_rigid = m.fusion_mask_2item(fwd_vo, dyn_f)
    bwd_vo_rigid = m.fusion_mask_2item(bwd_vo, dyn_b)
    fwd_vo_dyna = m.fusion_mask_2item(fwd_vo, [1 - k for k in dyn_f])
    bwd_vo_dyna = m.fusion_mask_2item(bwd_vo, [1 - k for k in dyn_b])
    loss = {
        "loss_depth_pixel": m.compute_photometric_loss(pc, rec_l, bwd_mask_tex) + m.compute_photometric_loss(pc, rec_r, fwd_mask_tex),
        "loss_depth_smooth": m.compute_smooth_loss(t.img, t.disp) + m.compute_smooth_loss(t.img_l, t.disp_l)
                             + m.compute_smooth_loss(t.img_r, t.disp_r),
        "loss_flow_pixel": m.compute_photometric_loss(pc, from_l, bwd_vo_rigid) + m.compute_photometric_loss(pc, from_r, fwd_vo_rigid)
                           + 2 * m.compute_photometric_loss(pc, from_l, bwd_vo_dyna) + 2 * m.compute_photometric_loss(pc, from_r, fwd_vo_dyna),
        "loss_flow_ssim": m.compute_ssim_loss(pc, from_l, bwd_vo) + m.compute_ssim_loss(pc, from_r, fwd_vo),
        "loss_flow_smooth": m.compute_loss_flow_smooth(t.flows_fwd, pc) + m.compute_loss_flow_smooth(t.flows_bwd, pc),
        "loss_flow_consis": m.compute_loss_flow_consis(t.flows_fwd, t.flows_bwd, occ_f),
        "loss_depth_flow_consis": m.compute_depth_flow_consis_loss(fd_b, bwd_mask, 1) + m.compute_depth_flow_consis_loss(fd_f, fwd_mask, 1),
        "loss_epipolar": m.compute_epipolar_loss(dist_b, dyn_b[0]) + m.compute_epipolar_loss(dist_f, dyn_f[0]),
    }
    aux = dict(occ_b=occ_b, occ_f=occ_f, valid_b=valid_b, valid_f=valid_f, dyn_b=dyn_b, dyn_f=dyn_f, tex_b=tex_b,
               tex_f=tex_f, val_l=val_l, val_r=val_r, dist_b=dist_b, dist_f=dist_f)
    return loss, aux
