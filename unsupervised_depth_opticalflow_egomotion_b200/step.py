"""Host-buffer entry points: one training-step worth of the loss path, starting from HOST memory.

``FlowLossStep`` is what a data-parallel trainer (``train.py:171-215`` in the reference: H2D copy of the
batch, loss forward, ``sum_k w_k * mean(loss_k)``, ``backward()``) calls when the frames and the
network outputs live in pinned host buffers: it stages them to the device with ONE copy, builds the image
pyramids, runs the fused forward + backward step (``ugl_flow_loss_step``) and brings the per-sample losses back.
It owns all its buffers (allocated once) and never synchronises except for the final D2H read.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops

FLOW_WEIGHTS = {"loss_flow_pixel": 0.15, "loss_flow_ssim": 0.85, "loss_flow_smooth": 10.0, "loss_flow_consis": 0.01}
"""config/kitti.yaml:18-21"""


def weight_matrix(weights: Dict[str, float], keys: Sequence[str], batch: int, device) -> torch.Tensor:
    """d(sum_k w_k * mean_B(loss_k)) / d loss  as a (K,B) matrix (train.py:211-214)."""
    w = torch.tensor([weights[k] for k in keys], dtype=torch.float32).view(-1, 1).repeat(1, batch) / float(batch)
    return w.to(device).contiguous()


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class _Slot:
    """One staging slot: a pinned host buffer and a device buffer with the SAME byte layout
    ``[img_l | img | img_r | flow_fwd[0..L) | flow_bwd[0..L)]`` (every tensor 256-byte aligned), so a step's inputs cross PCIe
    in one ``cudaMemcpyAsync`` instead of 3 + 2 L; typed views into both; the step's output buffers; the events that order reuse."""

    def __init__(self, batch, height, width, levels, device, frame_dtype):
        fsz = 1 if frame_dtype == torch.uint8 else 4
        shapes = [("img_l", (batch, 3, height, width), frame_dtype, fsz), ("img", (batch, 3, height, width), frame_dtype, fsz),
                  ("img_r", (batch, 3, height, width), frame_dtype, fsz)]
        shapes += [("ff%d" % l, (batch, 2, height >> l, width >> l), torch.float32, 4) for l in range(levels)]
        shapes += [("fb%d" % l, (batch, 2, height >> l, width >> l), torch.float32, 4) for l in range(levels)]
        off, layout = 0, []
        for name, shp, dt, sz in shapes:
            n = sz
            for s in shp:
                n *= s
            layout.append((name, shp, dt, off, n))
            off = _align(off + n)
        self.nbytes = off
        self.payload_bytes = sum(n for *_, n in layout)
        self.h_raw = torch.empty(self.nbytes, dtype=torch.uint8).pin_memory()
        self.d_raw = torch.empty(self.nbytes, dtype=torch.uint8, device=device)

        def views(raw):
            return {name: raw[o:o + n].view(dt).view(shp) for name, shp, dt, o, n in layout}
        self.h, self.d = views(self.h_raw), views(self.d_raw)
        self.imgs_in = [self.d["img_l"], self.d["img"], self.d["img_r"]]          # as staged (f32 or u8)
        self.imgs = (self.imgs_in if frame_dtype == torch.float32 else
                     [torch.empty((batch, 3, height, width), device=device, dtype=torch.float32) for _ in range(3)])
        self.ff = [self.d["ff%d" % l] for l in range(levels)]
        self.fb = [self.d["fb%d" % l] for l in range(levels)]
        self.h_loss = torch.empty((4, batch), dtype=torch.float32).pin_memory()
        self.h2d_done = torch.cuda.Event()
        self.compute_done = torch.cuda.Event()
        self.out: Optional[dict] = None       # ops.flow_loss_step buffers of this slot (loss, gf, gb, workspace)
        self.busy = False                     # a step was submitted on this slot (its events are recorded)
        self.pending = False                  # ... and its result has not been read yet


class FlowLossStep:
    """Flow-mode loss step (Model_flow.forward loss body, model_flow.py:232-254) from host buffers.

    Two staging slots and a copy stream: a step's inputs are written (by the data loader, the network's host-side output, ...)
    into the slot's pinned views ``host_views(slot)``; ``submit_staged`` enqueues ONE H2D copy of the slot on the copy stream and
    the pyramids + fused forward/backward step + D2H of the losses on the compute stream, so the copy of step k+1 overlaps the
    kernels of step k.  ``submit(tensors...)`` is the convenience form for inputs that live elsewhere on the host (it pays a host
    memcpy into the pinned slot).  ``result(slot)`` waits for that step only and returns a private copy of its losses; a slot must
    be read before it is submitted again (enforced).  ``d total / d flow`` of the last read step stays on the device in
    ``grads(slot)`` until the slot is re-submitted.

    ``frame_dtype=torch.uint8``: the three frames arrive as bytes (what the image decoder produces) and the dataset's
    ``img / 255.0`` (core/dataset/kitti_prepared.py:89) runs on the device, bit-identical to the host division: the frames then
    cost a quarter of the PCIe traffic (the step from host buffers is PCIe-bound)."""

    def __init__(self, batch: int, height: int, width: int, levels: int = 4, num_scales: Optional[int] = None,
                 weights: Optional[Dict[str, float]] = None, device="cuda:0", frame_dtype=torch.float32):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FlowLossStep needs a CUDA device: the loss path has no CPU implementation")
        self.B, self.H, self.W, self.L = batch, height, width, levels
        self.scales = levels if num_scales is None else num_scales
        with torch.cuda.device(self.device):
            if frame_dtype not in (torch.float32, torch.uint8):
                raise TypeError("frame_dtype must be torch.float32 or torch.uint8")
            self.frame_dtype = frame_dtype
            self.slots = [_Slot(batch, height, width, levels, self.device, frame_dtype) for _ in range(2)]
            self.copy_stream = torch.cuda.Stream()
        self.wmat = weight_matrix(weights or FLOW_WEIGHTS, ops.FLOW_LOSS_KEYS, batch, self.device)
        self.h2d_bytes = self.slots[0].payload_bytes       # bytes of the step's input tensors (the copy moves nbytes incl. padding)
        self.h2d_copies_per_step = 1
        self.d2h_bytes = self.slots[0].h_loss.numel() * 4
        self._next = 0

    # ---- staging -------------------------------------------------------------------------------------------------
    def next_slot(self) -> int:
        return self._next

    def host_views(self, slot: int) -> Dict[str, torch.Tensor]:
        """Pinned host tensors of a slot to fill in place: ``img_l, img, img_r`` (B,3,H,W) and ``ff<l>, fb<l>`` (B,2,H>>l,W>>l)."""
        return self.slots[slot].h

    def submit_staged(self, slot: Optional[int] = None) -> int:
        """Enqueue the step whose inputs are in ``host_views(slot)`` (default: the next slot in turn); returns the slot."""
        idx = self._next if slot is None else slot
        sl = self.slots[idx]
        if sl.pending:
            raise RuntimeError("FlowLossStep: slot %d is re-submitted before result(%d) of its previous step was read" % (idx, idx))
        self._next = idx ^ 1
        main = torch.cuda.current_stream(self.device)
        if sl.busy:
            self.copy_stream.wait_event(sl.compute_done)          # the slot's previous step must have consumed its buffers
        with torch.cuda.stream(self.copy_stream):
            sl.d_raw.copy_(sl.h_raw, non_blocking=True)            # the one H2D of this step
            sl.h2d_done.record(self.copy_stream)
        main.wait_event(sl.h2d_done)
        if self.frame_dtype == torch.uint8:
            ops.frames_from_u8(sl.imgs_in, out=sl.imgs)            # one launch for the three frames
        pl, pc, pr = (d["box"] for d in ops.image_pyramids(sl.imgs, self.L, ("box", "box", "box")))     # one launch
        sl.out = ops.flow_loss_step(pl, pc, pr, sl.ff, sl.fb, self.wmat, self.scales, out=sl.out)       # 4 launches: losses + d total / d flow
        sl.h_loss.copy_(sl.out["loss"], non_blocking=True)
        sl.compute_done.record(main)
        sl.busy = sl.pending = True
        return idx

    def submit(self, h_img_l: torch.Tensor, h_img: torch.Tensor, h_img_r: torch.Tensor, h_flows_fwd: Sequence[torch.Tensor],
               h_flows_bwd: Sequence[torch.Tensor]) -> int:
        """Convenience: copy host tensors into the next slot's pinned views (a host memcpy), then ``submit_staged``."""
        idx = self._next
        sl = self.slots[idx]
        if sl.pending:
            raise RuntimeError("FlowLossStep: slot %d is re-submitted before result(%d) of its previous step was read" % (idx, idx))
        if sl.busy:
            sl.h2d_done.synchronize()                              # the previous H2D of this slot must have read the pinned buffer
        for name, src in zip(("img_l", "img", "img_r"), (h_img_l, h_img, h_img_r)):
            if src.dtype != sl.h[name].dtype:
                raise TypeError("FlowLossStep(frame_dtype=%s) got %s frames" % (sl.h[name].dtype, src.dtype))
            sl.h[name].copy_(src)
        for l, (f, b) in enumerate(zip(h_flows_fwd, h_flows_bwd)):
            sl.h["ff%d" % l].copy_(f.detach())
            sl.h["fb%d" % l].copy_(b.detach())
        return self.submit_staged(idx)

    # ---- results ---------------------------------------------------------------------------------------------------
    def result(self, slot: int) -> torch.Tensor:
        """Per-sample losses (4,B) of a submitted step, on the host (a private copy: the slot's pinned buffer is reused)."""
        sl = self.slots[slot]
        if not sl.busy:
            raise RuntimeError("FlowLossStep: nothing was submitted on slot %d" % slot)
        sl.compute_done.synchronize()
        sl.pending = False
        return sl.h_loss.clone()

    def grads(self, slot: int) -> List[torch.Tensor]:
        """d total / d flow of the step last run on ``slot`` (forward levels, then backward levels), on the device; valid after
        ``result(slot)`` and until the slot is submitted again."""
        out = self.slots[slot].out
        return [] if out is None else list(out["gf"]) + list(out["gb"])

    def __call__(self, h_img_l, h_img, h_img_r, h_flows_fwd, h_flows_bwd, sync: bool = True):
        """``sync=True``: the step's losses (4,B) on the host.  ``sync=False``: the slot handle to pass to ``result`` later."""
        slot = self.submit(h_img_l, h_img, h_img_r, h_flows_fwd, h_flows_bwd)
        return self.result(slot) if sync else slot
