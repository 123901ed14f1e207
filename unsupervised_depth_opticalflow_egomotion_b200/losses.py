"""The reference's loss *methods* (``Model_flow`` / ``Model_depth`` / ``Model_geometry`` in
``core/networks``), re-hosted on the sm_100a kernels with the same names, argument meaning and
return shapes, plus ``forward_losses`` = the loss body of each ``forward`` given the network outputs.

The networks (PWC-Net, ResNet depth net, PoseCNN) are not part of this package: a trainer keeps the
reference's modules and swaps its loss methods for these.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .structures import (calculate_rigid_flow, compute_essential_matrix, inv3x3, inverse_warp2, projection_pyramid,
                         scaled_intrinsics, warp_flow)

Tensor = torch.Tensor


_ZEROS2: Dict[str, Tensor] = {}


def _zeros2(like: Tensor) -> Tensor:
    """The reference's disabled-term placeholder ``torch.zeros([2]).to(device).requires_grad_()``: a fresh leaf per call over one
    constant buffer per device (no fill launch inside a captured step)."""
    z = _ZEROS2.get(str(like.device))
    if z is None:
        z = _ZEROS2[str(like.device)] = torch.zeros([2], device=like.device)
    return z.detach().requires_grad_()


class _LossBase:
    def __init__(self, num_scales: int):
        self.num_scales = int(num_scales)

    # ---- shared helpers (identical bodies in the three reference classes) --------------------------------
    def warp_flow_pyramid(self, img_pyramid, flow_pyramid):
        """model_geometry.py:74-78 / model_flow.py:66-70"""
        return [warp_flow(img, flow, use_mask=True) for img, flow in zip(img_pyramid, flow_pyramid)]

    def compute_photometric_loss(self, img_list, img_warped_list, mask_list):
        """model_geometry.py:143-153 / model_depth.py:92-103 / model_flow.py:72-81 (compute_loss_pixel)"""
        return sum(ops.masked_l1(img_list[s], img_warped_list[s], mask_list[s]) for s in range(self.num_scales))

    compute_loss_pixel = compute_photometric_loss

    def compute_ssim_loss(self, img_list, img_warped_list, mask_list):
        """model_geometry.py:212-223 / model_flow.py:141-152"""
        return sum(ops.ssim_loss(img_list[s], img_warped_list[s], mask_list[s]) for s in range(self.num_scales))

    compute_loss_ssim = compute_ssim_loss

    def compute_loss_flow_smooth(self, optical_flows, img_pyramid):
        """model_geometry.py:271-279 / model_flow.py:173-181"""
        return sum(ops.flow_smooth(optical_flows[s], img_pyramid[s]) for s in range(self.num_scales))

    def compute_loss_flow_consis(self, fwd_flow_pyramid, bwd_flow_pyramid, occ_mask_list):
        """model_geometry.py:195-210 / model_flow.py:184-199"""
        return sum(ops.flow_consis(fwd_flow_pyramid[s], bwd_flow_pyramid[s], occ_mask_list[s]) for s in range(self.num_scales))

    def compute_smooth_loss(self, img, disps, mode: str = "single_pass"):
        """model_geometry.py:225-252 / model_depth.py:220-247"""
        return ops.disp_smooth(img, list(disps[:self.num_scales]), mode=mode)

    def compute_smooth_loss3(self, imgs, disp_lists):
        """the three back-to-back ``compute_smooth_loss`` calls (model_geometry.py:938-940 / model_depth.py:281-283), one launch"""
        return ops.disp_smooth_multi(list(imgs), [list(d[:self.num_scales]) for d in disp_lists]).sum(0)

    def compute_texture_mask(self, img_list, img_warped_list, img_list_source):
        """model_geometry.py:134-140 / model_depth.py:84-90"""
        return [ops.texture_mask(img_list[s], img_warped_list[s], img_list_source[s]) for s in range(self.num_scales)]

    def reconstruction(self, ref_img, intrinsics, depth, depth_ref, pose, padding_mode="zeros"):
        """model_geometry.py:80-103 / model_depth.py:59-82: area-resized source, K rows 0-1 / downscale, inverse_warp2."""
        rec, valid, proj, comp = [], [], [], []
        H = ref_img.size(2)
        area = ops.image_pyramid(ref_img, self.num_scales, "area")
        for s in range(self.num_scales):
            h, w = depth[s].shape[2:]
            if (h, w) != tuple(area[s].shape[2:]):
                raise ValueError("reconstruction: depth level %d is %dx%d, expected %s" % (s, h, w, tuple(area[s].shape[2:])))
            Ks = scaled_intrinsics(intrinsics, H / h)
            a, b, c, d = inverse_warp2(area[s], depth[s], depth_ref[s], pose, Ks, padding_mode)
            rec.append(a); valid.append(b); proj.append(c); comp.append(d)
        return rec, valid, proj, comp


    # ---- batched variants used by forward_losses: identical numbers, the 3x3 glue computed once per step ------------
    def _projections(self, ref_h: int, intrinsics, depth, poses):
        downs = [ref_h / depth[s].shape[2] for s in range(self.num_scales)]
        return projection_pyramid(intrinsics, poses, downs)

    def _pose_setup(self, ref_h: int, intrinsics, depth, pose_vectors, K_inv=None, fundamental=False):
        """one-launch form of :meth:`_projections` (+ the fundamental matrices of compute_epipolar_map): pose index 0 = centre->left
        (bwd), 1 = centre->right (fwd)"""
        downs = [ref_h / depth[s].shape[2] for s in range(self.num_scales)]
        return ops.pose_setup(pose_vectors, intrinsics, downs, K_inv, fundamental)

    def _reconstruction_with(self, ref_img, depth, depth_ref, Kinv, P):
        rec, valid, proj, comp = [], [], [], []
        area = ops.image_pyramid(ref_img, self.num_scales, "area")
        for s in range(self.num_scales):
            if tuple(depth[s].shape[2:]) != tuple(area[s].shape[2:]):
                raise ValueError("reconstruction: depth level %d has shape %s, expected %s" % (s, tuple(depth[s].shape[2:]), tuple(area[s].shape[2:])))
            a, b, c, d = ops.reproject(area[s], depth[s], depth_ref[s], Kinv[s], P[s])
            rec.append(a); valid.append(b); proj.append(c); comp.append(d)
        return rec, valid, proj, comp


class FlowLoss(_LossBase):
    """Loss methods of ``Model_flow`` (model_flow.py)."""

    def generate_img_pyramid(self, img, num_pyramid):
        """model_flow.py:58-64 (adaptive average pooling, detached)"""
        return ops.image_pyramid(img, num_pyramid, "box")

    def get_occlusion_mask_from_flow(self, tensor_size, flow):
        """model_flow.py:33-39 — EXTENSION: the reference calls an undefined ``transformerFwd`` here (dead code); this is the
        upstream TrianFlow forward splat of a ones map, clamped to [0,1]."""
        ones = torch.ones(tuple(tensor_size), device=flow.device, dtype=torch.float32)
        return ops.forward_splat(ones, flow, clamp01=True)

    def compute_diff_weight(self, img_pyramid_from_l, img_pyramid, img_pyramid_from_r):
        """model_flow.py:105-138 -> (diff_bwd, diff_fwd, weight_bwd, weight_fwd)"""
        out = [ops.occlusion_weights(img_pyramid_from_l[s], img_pyramid[s], img_pyramid_from_r[s], soft=True)
               for s in range(self.num_scales)]
        return [o[4] for o in out], [o[5] for o in out], [o[0] for o in out], [o[1] for o in out]

    def compute_loss_with_mask(self, diff_list, occ_mask_list):
        """model_flow.py:94-103"""
        return sum(ops.masked_mean(diff_list[s], occ_mask_list[s]) for s in range(self.num_scales))

    def forward_losses(self, imgl, img, imgr, optical_flows_fwd, optical_flows_bwd, fused: bool = True) -> Dict[str, Tensor]:
        """Loss body of ``Model_flow.forward`` (model_flow.py:232-254).  ``fused=True`` runs the single fused
        kernel pair; ``fused=False`` composes the per-method kernels exactly like the reference does."""
        L = len(optical_flows_fwd)
        if fused:   # the three box pyramids in one launch, then the fused kernels
            pl, pc, pr = (d["box"] for d in ops.image_pyramids((imgl, img, imgr), L, ("box", "box", "box")))
            return ops.flow_loss(pl, pc, pr, optical_flows_fwd, optical_flows_bwd, self.num_scales)
        pl, pc, pr = (self.generate_img_pyramid(x, L) for x in (imgl, img, imgr))
        from_l = self.warp_flow_pyramid(pl, optical_flows_bwd)
        from_r = self.warp_flow_pyramid(pr, optical_flows_fwd)
        diff_bwd, diff_fwd, w_bwd, w_fwd = self.compute_diff_weight(from_l, pc, from_r)
        return {
            "loss_flow_pixel": self.compute_loss_with_mask(diff_fwd, w_fwd) + self.compute_loss_with_mask(diff_bwd, w_bwd),
            "loss_flow_ssim": self.compute_loss_ssim(pc, from_r, w_fwd) + self.compute_loss_ssim(pc, from_l, w_bwd),
            "loss_flow_smooth": self.compute_loss_flow_smooth(optical_flows_fwd, pc) + self.compute_loss_flow_smooth(optical_flows_bwd, pc),
            "loss_flow_consis": self.compute_loss_flow_consis(optical_flows_fwd, optical_flows_bwd, w_fwd),
        }


class DepthLoss(_LossBase):
    """Loss methods of ``Model_depth`` (model_depth.py).  ``variant='live'``: model_depth.py:281-335 (L1 + smoothness);
    ``'texture'``: model_depth_texture.py:296-311 (L1 + SSIM + smoothness + unmasked depth consistency); ``'ssim'``: the same file
    without its consistency term (:296-307) -- BASELINE configs[2] as SURVEY 8(d) spells it out: reprojection + SSIM + smoothness."""

    def __init__(self, num_scales: int, variant: str = "live"):
        super().__init__(num_scales)
        if variant not in ("live", "texture", "ssim"):
            raise ValueError("variant must be 'live', 'texture' or 'ssim'")
        self.variant = variant

    def generate_img_pyramid(self, img, num_pyramid):
        """model_depth.py:44-50 (bilinear)"""
        return ops.image_pyramid(img, num_pyramid, "bilinear")

    def fusion_mask(self, valid_mask, texture_mask):
        """model_depth.py:262-269"""
        return [ops.mask_product([valid_mask[s], texture_mask[s]]) for s in range(self.num_scales)]

    def compute_consis_loss(self, predicted_depth_list, computed_depth_list):
        """model_depth.py:154-163 (unmasked)"""
        return sum(ops.masked_mean(ops.depth_diff(computed_depth_list[s], predicted_depth_list[s]), None) for s in range(self.num_scales))

    def forward_losses(self, img_l, img, img_r, disp_list, disp_l_list, disp_r_list, pose_vectors, K, fused=True) -> Tuple[Dict[str, Tensor], Dict]:
        """Loss body of ``Model_depth.forward`` (model_depth.py:281-335) / model_depth_texture.py:262-311.  ``fused=True``: the fused
        kernels as one autograd node; ``fused="ops"``: the same kernels as separate autograd Functions; ``fused=False``: one kernel per
        reference method."""
        S = self.num_scales
        pose_fwd, pose_bwd = pose_vectors[:, 1, :], pose_vectors[:, 0, :]
        if fused is True:
            # the whole loss body as ONE autograd node (mode_steps.py): the fused kernels below, their gradient accumulation in one
            # launch, no library kernels in between.  fused="ops" keeps the same kernels as separate autograd Functions.
            from . import mode_steps
            mat, valid, tex = mode_steps.depth_step(S, self.variant, img_l, img, img_r, disp_list, disp_l_list, disp_r_list, pose_vectors, K)
            keys = mode_steps.DEPTH_KEYS[self.variant]
            ph = {k: _zeros2(img) for k in ("loss_depth_pixel", "loss_depth_ssim", "loss_depth_consis", "loss_depth_smooth") if k not in keys}
            return mode_steps.LossPack(mat, keys, ph), dict(valid_l=valid[0], valid_r=valid[1], tex_b=tex[0], tex_f=tex[1])
        if fused:
            Kinv, (P_b, P_f), _ = self._pose_setup(img.size(2), K, disp_list, pose_vectors)
            pyr = ops.image_pyramids((img, img_l, img_r), S, ("bilinear", ("bilinear", "area"), ("bilinear", "area")))   # one launch
            pc, pl, pr = (d["bilinear"] for d in pyr)
        else:
            Kinv, (P_b, P_f) = self._projections(img.size(2), K, disp_list, [pose_bwd, pose_fwd])
            pl, pc, pr = (self.generate_img_pyramid(x, S) for x in (img_l, img, img_r))
        if fused and self.variant == "live":
            area = (pyr[1]["area"], pyr[2]["area"])
            pix, valid, tex = ops.depth_photo_loss(pc, area, (pl, pr), list(disp_list[:S]), Kinv, (P_b, P_f))
            loss = {"loss_depth_pixel": pix, "loss_depth_ssim": _zeros2(img), "loss_depth_consis": _zeros2(img),
                    "loss_depth_smooth": self.compute_smooth_loss3((img, img_l, img_r), (disp_list, disp_l_list, disp_r_list))}
            return loss, dict(valid_l=valid[0], valid_r=valid[1], tex_b=tex[0], tex_f=tex[1])
        if fused and self.variant in ("texture", "ssim"):
            # L1 + SSIM of both source frames and all levels: the single-pass tile kernel in depth mode; the depth-consistency term
            # (disabled in the live Model_depth, model_depth.py:333-335): one forward / backward kernel set for all levels
            area = (pyr[1]["area"], pyr[2]["area"])
            l2, valid, tex = ops.depth_ssim_loss(pc, area, (pl, pr), list(disp_list[:S]), Kinv, (P_b, P_f))
            consis = (ops.depth_consis_loss(list(disp_list[:S]), (list(disp_l_list[:S]), list(disp_r_list[:S])), Kinv, (P_b, P_f))
                      if self.variant == "texture" else _zeros2(img))
            loss = {"loss_depth_pixel": l2[0], "loss_depth_ssim": l2[1], "loss_depth_consis": consis,
                    "loss_depth_smooth": self.compute_smooth_loss3((img, img_l, img_r), (disp_list, disp_l_list, disp_r_list))}
            return loss, dict(valid_l=valid[0], valid_r=valid[1], tex_b=tex[0], tex_f=tex[1])
        rec_l, val_l, proj_l, comp_l = self._reconstruction_with(img_l, disp_list, disp_l_list, Kinv, P_b)
        rec_r, val_r, proj_r, comp_r = self._reconstruction_with(img_r, disp_list, disp_r_list, Kinv, P_f)
        tex_b = self.compute_texture_mask(pc, rec_l, pl)
        tex_f = self.compute_texture_mask(pc, rec_r, pr)
        m_b, m_f = self.fusion_mask(val_l, tex_b), self.fusion_mask(val_r, tex_f)
        loss = {"loss_depth_pixel": self.compute_photometric_loss(pc, rec_l, m_b) + self.compute_photometric_loss(pc, rec_r, m_f)}
        if self.variant in ("texture", "ssim"):
            loss["loss_depth_ssim"] = self.compute_ssim_loss(pc, rec_l, val_l) + self.compute_ssim_loss(pc, rec_r, val_r)
            loss["loss_depth_consis"] = (self.compute_consis_loss(proj_l, comp_l) + self.compute_consis_loss(proj_r, comp_r)
                                         if self.variant == "texture" else _zeros2(img))
        else:
            loss["loss_depth_ssim"] = _zeros2(img)
            loss["loss_depth_consis"] = _zeros2(img)
        if fused:
            loss["loss_depth_smooth"] = self.compute_smooth_loss3((img, img_l, img_r), (disp_list, disp_l_list, disp_r_list))
        else:
            loss["loss_depth_smooth"] = (self.compute_smooth_loss(img, disp_list, "recompute") + self.compute_smooth_loss(img_l, disp_l_list, "recompute")
                                         + self.compute_smooth_loss(img_r, disp_r_list, "recompute"))
        masks = dict(valid_l=val_l, valid_r=val_r, tex_b=tex_b, tex_f=tex_f)
        return loss, masks


_WEIGHT_CACHE: Dict[tuple, Tensor] = {}


def total_loss(loss_pack: Dict[str, Tensor], weights: Dict[str, float]) -> Tensor:
    """``train.py:211-214``: ``sum_k weights[k] * loss_pack[k].mean()`` as four launches instead of four per key: the live
    ``(B,)`` terms are stacked, averaged, weighted and summed at once.  The reference's constant ``zeros([2])`` placeholders add
    exactly 0 and carry no gradient, so they are skipped.  Gradients are bit-identical to the per-key loop
    (``d total / d loss_k[b] = fl(w_k / B)``)."""
    if hasattr(loss_pack, "total"):          # mode_steps.LossPack: the live terms are rows of one matrix -> one launch
        return loss_pack.total(weights)
    items = list(loss_pack.items())
    B = max(v.numel() for _, v in items)
    live = [(k, v) for k, v in items if v.numel() == B]          # B == 2: the placeholders stack like any other term
    if any(v.numel() not in (B, 2) for _, v in items):
        return sum(weights[k] * v.mean() for k, v in items)
    key = (tuple(k for k, _ in live), tuple(float(weights[k]) for k, _ in live), str(live[0][1].device))
    w = _WEIGHT_CACHE.get(key)
    if w is None:       # built once per (keys, weights, device): no host-to-device copy inside a captured step
        w = _WEIGHT_CACHE[key] = torch.tensor(key[1], dtype=torch.float32, device=live[0][1].device)
    return (torch.stack([v for _, v in live]).mean(1) * w).sum()


class GeomMasks(dict):
    """``mask_pack`` of the fused geom path (model_geometry.py:871-880): the float maps the kernels wrote are plain dict
    entries; the flow-branch masks (``occ_b/f, valid_b/f, dyn_b/f, fwd_mask, bwd_mask, rigid_f, inlier_f``) are unpacked from
    the packed byte maps only when asked for — the training step never reads them."""

    _BITS = dict(valid_b=ops.MASK_VALID_BWD, valid_f=ops.MASK_VALID_FWD, occ_b=ops.MASK_OCC_BWD, occ_f=ops.MASK_OCC_FWD,
                 dyn_b=ops.MASK_DYN_BWD, dyn_f=ops.MASK_DYN_FWD, bwd_mask=ops.MASK_ALL_BWD, fwd_mask=ops.MASK_ALL_FWD)

    def __init__(self, mask_bytes, maps, owner, epi_inputs):
        super().__init__(maps)
        self.mask_bytes, self._owner, self._epi = mask_bytes, owner, epi_inputs

    def __missing__(self, key):
        if key in self._BITS:
            val = [ops.unpack_mask(m, self._BITS[key]) for m in self.mask_bytes]
        elif key in ("dist_b", "dist_f"):        # the epipolar distance maps are only materialised for mask_pack consumers
            with torch.no_grad():
                fb, ff, Fm = self._epi
                self["dist_b"], self["dist_f"] = ops.epipolar_distance(fb, Fm[0]), ops.epipolar_distance(ff, Fm[1])
            return self[key]
        elif key in ("rigid_f", "inlier_f"):
            rigid, inlier, _ = self._owner.get_rigid_mask(self["dist_f"])
            self["rigid_f"], self["inlier_f"] = rigid, inlier
            return self[key]
        else:
            raise KeyError(key)
        self[key] = val
        return val


class GeometryLoss(_LossBase):
    """Loss methods of ``Model_geometry`` (model_geometry.py)."""

    def __init__(self, num_scales: int, flow_consist_alpha: float = 0.01, flow_consist_beta: float = 0.5,
                 rigid_thres: float = 0.5, inlier_thres: float = 0.1):
        super().__init__(num_scales)
        self.flow_consist_alpha, self.flow_consist_beta = flow_consist_alpha, flow_consist_beta
        self.rigid_thres, self.inlier_thres = rigid_thres, inlier_thres

    def generate_img_pyramid(self, img, num_pyramid):
        """model_geometry.py:65-72 (bilinear)"""
        return ops.image_pyramid(img, num_pyramid, "bilinear")

    def compute_occ_weight(self, img_pyramid_from_l, img_pyramid, img_pyramid_from_r):
        """model_geometry.py:105-132 -> (weight_bwd, weight_fwd, valid_bwd, valid_fwd)"""
        out = [ops.occlusion_weights(img_pyramid_from_l[s], img_pyramid[s], img_pyramid_from_r[s], soft=False)
               for s in range(self.num_scales)]
        return [o[0] for o in out], [o[1] for o in out], [o[2] for o in out], [o[3] for o in out]

    def compute_consis_loss(self, predicted_depth_list, computed_depth_list, mask_list):
        """model_geometry.py:182-193 (masked)"""
        return sum(ops.masked_mean(ops.depth_diff(computed_depth_list[s], predicted_depth_list[s]), mask_list[s])
                   for s in range(self.num_scales))

    def compute_dynamic_mask(self, intrinsics, depth, pose, flow):
        """model_geometry.py:685-713 -> (flow_diffs, dynamic_masks, flow_diff_scores)"""
        diffs, masks, scores = [], [], []
        H0 = depth[0].size(2)
        for s in range(self.num_scales):
            h = depth[s].size(2)
            Ks = scaled_intrinsics(intrinsics, H0 / h)
            rf = calculate_rigid_flow(depth[s], pose, Ks)
            fd, dyn, score = ops.dynamic_mask(flow[s], rf, self.flow_consist_alpha, self.flow_consist_beta)
            diffs.append(fd); masks.append(dyn); scores.append(score)
        return diffs, masks, scores

    def _dynamic_mask_with(self, depth, flow, Kinv, P):
        diffs, masks, scores = [], [], []
        for s in range(self.num_scales):
            fd, dyn, score = ops.dynamic_mask(flow[s], ops.rigid_flow(depth[s], Kinv[s], P[s]), self.flow_consist_alpha, self.flow_consist_beta)
            diffs.append(fd); masks.append(dyn); scores.append(score)
        return diffs, masks, scores

    def compute_depth_flow_consis_loss(self, flow_diffs, masks=None, scales=3):
        """model_geometry.py:716-732"""
        return sum(ops.masked_mean(flow_diffs[s], None if masks is None else masks[s]) for s in range(scales))

    def compute_epipolar_map(self, pose, flow, intrinsics, intrinsics_inverse):
        """model_geometry.py:355-403 -> (B,1,h,w) point-to-epipolar-line distance"""
        E = compute_essential_matrix(pose)
        Fm = intrinsics_inverse.transpose(1, 2).bmm(E.bmm(intrinsics_inverse))
        return ops.epipolar_distance(flow, Fm.contiguous())

    def compute_epipolar_loss(self, dist_map, rigid_mask):
        """model_geometry.py:413-418: the masked value is overwritten by the plain mean (:416)."""
        return ops.masked_mean(dist_map, None)

    def get_rigid_mask(self, dist_map):
        """model_geometry.py:420-425"""
        return ops.rigid_mask(dist_map, self.rigid_thres, self.inlier_thres)

    def fusion_mask(self, valid_mask, occ_mask, dynamic_mask):
        """model_geometry.py:735-745"""
        return [ops.mask_product([valid_mask[s], occ_mask[s], dynamic_mask[s]]) for s in range(self.num_scales)]

    def fusion_mask_4item(self, valid_mask, occ_mask, dynamic_mask, texture_mask):
        """model_geometry.py:747-756"""
        return [ops.mask_product([valid_mask[s], occ_mask[s], dynamic_mask[s], texture_mask[s]]) for s in range(self.num_scales)]

    def fusion_mask_2item(self, valid_mask, occ_mask, invert_second: bool = False):
        """model_geometry.py:757-765 (``invert_second`` fuses the ``[1-mask for mask in ...]`` of :863-864)"""
        return [ops.mask_product([valid_mask[s], occ_mask[s]], [False, invert_second]) for s in range(self.num_scales)]

    def _forward_losses_fused(self, img, img_l, img_r, pyr, flows_fwd, flows_bwd, disp_list, disp_l_list, disp_r_list,
                              Fm, Kinv, P_b, P_f):
        """Two stencil kernels carry the loss loop: the geom-mode single-pass flow kernel (warps, every mask, the four flow
        terms; masks leave as one packed byte map per level) and the reprojection-photometric kernel (reads the byte maps).
        The level-0 point-wise terms (depth-flow consistency, epipolar) and the disparity smoothness stay on their ops."""
        S = self.num_scales
        pc, pl, pr = (d["bilinear"] for d in pyr)
        flow4, mbytes = ops.geom_flow_loss(pl, pc, pr, list(flows_fwd), list(flows_bwd), list(disp_list[:S]), Kinv, P_b, P_f,
                                           self.flow_consist_alpha, self.flow_consist_beta, S)
        area = (pyr[1]["area"], pyr[2]["area"])
        depth_pixel, (val_l, val_r), (tex_b, tex_f) = ops.depth_photo_loss(
            pc, area, (pl, pr), list(disp_list[:S]), Kinv, (P_b, P_f), ext_bytes=mbytes, ext_need=(ops.MASK_ALL_BWD, ops.MASK_ALL_FWD))
        # level 0: |rigid flow - flow| under valid * occ * dyn (:921-926) and the epipolar distance means (:928-935), one kernel
        dfc, epi = ops.geom_rigid_terms(flows_bwd[0], flows_fwd[0], disp_list[0], mbytes[0], Kinv[0], P_b[0], P_f[0], Fm[0], Fm[1])
        loss = {
            "loss_depth_pixel": depth_pixel,
            "loss_depth_ssim": _zeros2(img),
            "loss_depth_smooth": self.compute_smooth_loss3((img, img_l, img_r), (disp_list, disp_l_list, disp_r_list)),
            "loss_depth_consis": _zeros2(img),
            "loss_flow_pixel": flow4[0], "loss_flow_ssim": flow4[1], "loss_flow_smooth": flow4[2], "loss_flow_consis": flow4[3],
            "loss_depth_flow_consis": dfc,
            "loss_epipolar": epi,
            "loss_triangle": _zeros2(img), "loss_pnp": _zeros2(img), "loss_eight_point": _zeros2(img),
        }
        return loss, GeomMasks(mbytes, dict(tex_b=tex_b, tex_f=tex_f, val_l=val_l, val_r=val_r), self, (flows_bwd[0], flows_fwd[0], Fm))

    def forward_losses(self, img_l, img, img_r, optical_flows_fwd, optical_flows_bwd, disp_list, disp_l_list, disp_r_list,
                       pose_vectors, K, K_inv, fused=True, step_weights: Optional[Dict[str, float]] = None) -> Tuple[Dict[str, Tensor], Dict]:
        """Loss body of ``Model_geometry.forward`` (model_geometry.py:777-951) given the network outputs.
        The second return value holds the device-side masks (the reference's ``mask_pack`` without its
        unconditional D2H copies, :871-880).  ``step_weights`` (``fused=True`` only): the loss weights of ``train.py:211-215`` declared
        up front -- the flow branch then evaluates its gradients in its forward launches (``mode_steps.geom_step``); ``total_loss``
        must be taken with the same weights."""
        S = self.num_scales
        pose_fwd, pose_bwd = pose_vectors[:, 1, :], pose_vectors[:, 0, :]
        if fused is True:
            from . import mode_steps
            mat, mbytes, (val_l, val_r), (tex_b, tex_f), Fm = mode_steps.geom_step(
                S, self.flow_consist_alpha, self.flow_consist_beta, img_l, img, img_r, list(optical_flows_fwd), list(optical_flows_bwd),
                disp_list, disp_l_list, disp_r_list, pose_vectors, K, K_inv, step_weights=step_weights)
            ph = {k: _zeros2(img) for k in ("loss_depth_ssim", "loss_depth_consis", "loss_triangle", "loss_pnp", "loss_eight_point")}
            return (mode_steps.LossPack(mat, mode_steps.GEOM_KEYS, ph, step_weights),
                    GeomMasks(mbytes, dict(tex_b=tex_b, tex_f=tex_f, val_l=val_l, val_r=val_r), self,
                              (optical_flows_bwd[0], optical_flows_fwd[0], list(Fm))))
        if fused:
            Kinv, (P_b, P_f), Fm = self._pose_setup(img.size(2), K, disp_list, pose_vectors, K_inv, fundamental=True)
            pyr = ops.image_pyramids((img, img_l, img_r), S, ("bilinear", ("bilinear", "area"), ("bilinear", "area")))   # one launch
            return self._forward_losses_fused(img, img_l, img_r, pyr, optical_flows_fwd, optical_flows_bwd, disp_list, disp_l_list,
                                              disp_r_list, Fm, Kinv, P_b, P_f)
        pc, pl, pr = (self.generate_img_pyramid(x, S) for x in (img, img_l, img_r))
        Kinv, (P_b, P_f) = self._projections(img.size(2), K, disp_list, [pose_bwd, pose_fwd])
        # composed path: one kernel per reference method (kept as the cross-check of the fused kernels)
        rec_l, val_l, _, _ = self._reconstruction_with(img_l, disp_list, disp_l_list, Kinv, P_b)
        rec_r, val_r, _, _ = self._reconstruction_with(img_r, disp_list, disp_r_list, Kinv, P_f)
        tex_b = self.compute_texture_mask(pc, rec_l, pl)
        tex_f = self.compute_texture_mask(pc, rec_r, pr)
        from_l = self.warp_flow_pyramid(pl, optical_flows_bwd)
        from_r = self.warp_flow_pyramid(pr, optical_flows_fwd)
        occ_b, occ_f, valid_b, valid_f = self.compute_occ_weight(from_l, pc, from_r)
        fd_b, dyn_b, _ = self._dynamic_mask_with(disp_list, optical_flows_bwd, Kinv, P_b)
        fd_f, dyn_f, _ = self._dynamic_mask_with(disp_list, optical_flows_fwd, Kinv, P_f)
        dist_b = self.compute_epipolar_map(pose_bwd, optical_flows_bwd[0], K, K_inv)
        dist_f = self.compute_epipolar_map(pose_fwd, optical_flows_fwd[0], K, K_inv)
        rigid_f, inlier_f, _ = self.get_rigid_mask(dist_f)

        fwd_mask = self.fusion_mask(valid_f, occ_f, dyn_f)
        bwd_mask = self.fusion_mask(valid_b, occ_b, dyn_b)
        fwd_mask_tex = self.fusion_mask_2item(fwd_mask, tex_f)
        bwd_mask_tex = self.fusion_mask_2item(bwd_mask, tex_b)
        fwd_vo = self.fusion_mask_2item(valid_f, occ_f)
        bwd_vo = self.fusion_mask_2item(valid_b, occ_b)
        fwd_vo_rigid = self.fusion_mask_2item(fwd_vo, dyn_f)
        bwd_vo_rigid = self.fusion_mask_2item(bwd_vo, dyn_b)
        fwd_vo_dyna = self.fusion_mask_2item(fwd_vo, dyn_f, invert_second=True)
        bwd_vo_dyna = self.fusion_mask_2item(bwd_vo, dyn_b, invert_second=True)

        P = self.compute_photometric_loss
        loss = {
            "loss_depth_pixel": P(pc, rec_l, bwd_mask_tex) + P(pc, rec_r, fwd_mask_tex),
            "loss_depth_ssim": _zeros2(img),
            "loss_depth_smooth": self.compute_smooth_loss(img, disp_list, "recompute") + self.compute_smooth_loss(img_l, disp_l_list, "recompute")
                                 + self.compute_smooth_loss(img_r, disp_r_list, "recompute"),
            "loss_depth_consis": _zeros2(img),
            "loss_flow_pixel": P(pc, from_l, bwd_vo_rigid) + P(pc, from_r, fwd_vo_rigid)
                               + 2 * P(pc, from_l, bwd_vo_dyna) + 2 * P(pc, from_r, fwd_vo_dyna),
            "loss_flow_ssim": self.compute_ssim_loss(pc, from_l, bwd_vo) + self.compute_ssim_loss(pc, from_r, fwd_vo),
            "loss_flow_smooth": self.compute_loss_flow_smooth(optical_flows_fwd, pc) + self.compute_loss_flow_smooth(optical_flows_bwd, pc),
            "loss_flow_consis": self.compute_loss_flow_consis(optical_flows_fwd, optical_flows_bwd, occ_f),
            "loss_depth_flow_consis": self.compute_depth_flow_consis_loss(fd_b, bwd_mask, 1)
                                      + self.compute_depth_flow_consis_loss(fd_f, fwd_mask, 1),
            "loss_epipolar": self.compute_epipolar_loss(dist_b, dyn_b[0]) + self.compute_epipolar_loss(dist_f, dyn_f[0]),
            "loss_triangle": _zeros2(img), "loss_pnp": _zeros2(img), "loss_eight_point": _zeros2(img),
        }
        masks = dict(occ_b=occ_b, occ_f=occ_f, valid_b=valid_b, valid_f=valid_f, dyn_b=dyn_b, dyn_f=dyn_f, tex_b=tex_b, tex_f=tex_f,
                     val_l=val_l, val_r=val_r, dist_b=dist_b, dist_f=dist_f, rigid_f=rigid_f, inlier_f=inlier_f, fwd_mask=fwd_mask,
                     bwd_mask=bwd_mask)
        return loss, masks
