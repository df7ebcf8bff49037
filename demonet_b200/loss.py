"""Training-side operators of the SSD detectors on the B200 kernels.

Names and argument meaning follow the reference:
  SSDMatcher(threshold)(match_quality_matrix)        demonet/models/_utils.py:350-362 (Matcher.__call__ :283-323)
  match_targets(targets, anchors, iou_thresh)        the matching loop of SSD.forward, generalized_ssd.py:326-335
  compute_loss(targets, head_outputs, anchors, matched_idxs, neg_to_pos_ratio)
                                                     SSD.compute_loss, generalized_ssd.py:210-269
  check_targets / resize_targets                     the target checks and resize_boxes of SSD.forward / the transform
                                                     (generalized_ssd.py:275-307, transform.py:278-292)
compute_loss is differentiable with respect to head_outputs['cls_logits'] and head_outputs['bbox_regression'] (closed-form
gradients written by the same kernel sequence), so it can sit behind any head that produces those tensors.  Everything
runs on CUDA tensors through the C ABI (dn_ssd_match, dn_match_quality, dn_ssd_loss); there is no CPU path.
"""
import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _C

BOX_WEIGHTS = (10.0, 10.0, 5.0, 5.0)             # BoxCoder(weights=(10., 10., 5., 5.)), generalized_ssd.py:169


def _stream(t: Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("demonet_b200 operators run on CUDA tensors only (no CPU fallback)")


def _pack_targets(targets: Sequence[Dict[str, Tensor]], device) -> Tuple[Tensor, Tensor, Tensor, int]:
    """boxes [G,4] fp32, labels [G] int64, offsets [B+1] int32 (device), G."""
    counts = [int(t["boxes"].shape[0]) for t in targets]
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    G = offs[-1]
    if G:
        boxes = torch.cat([t["boxes"].reshape(-1, 4) for t in targets]).to(device, torch.float32).contiguous()
        labels = torch.cat([t["labels"].reshape(-1) for t in targets]).to(device, torch.int64).contiguous()
    else:
        boxes = torch.zeros((0, 4), dtype=torch.float32, device=device)
        labels = torch.zeros((0,), dtype=torch.int64, device=device)
    offsets = torch.tensor(offs, dtype=torch.int32).to(device)
    return boxes, labels, offsets, G


def _anchors_2d(anchors, device) -> Tensor:
    # the reference passes one [P,4] tensor per image, all equal (anchor_utils.py:110-126)
    a = anchors[0] if isinstance(anchors, (list, tuple)) else anchors
    if a.dim() == 3:
        a = a[0]
    return a.to(device, torch.float32).contiguous()


def check_targets(targets) -> None:
    """The target checks of SSD.forward (generalized_ssd.py:275-286, 298-307), same exceptions and messages."""
    for target in targets:
        boxes = target["boxes"]
        if isinstance(boxes, torch.Tensor):
            if len(boxes.shape) != 2 or boxes.shape[-1] != 4:
                raise ValueError("Expected target boxes to be a tensor"
                                 "of shape [N, 4], got {:}.".format(boxes.shape))
        else:
            raise ValueError("Expected target boxes to be of type "
                             "Tensor, got {:}.".format(type(boxes)))
    for target_idx, target in enumerate(targets):
        boxes = target["boxes"]
        degenerate_boxes = boxes[:, 2:] <= boxes[:, :2]
        if degenerate_boxes.any():
            bb_idx = torch.where(degenerate_boxes.any(dim=1))[0][0]
            degen_bb: List[float] = boxes[bb_idx].tolist()
            raise ValueError("All bounding boxes should have positive height and width."
                             " Found invalid box {} for target at index {}."
                             .format(degen_bb, target_idx))


def resize_targets(targets, original_sizes: Sequence[Tuple[int, int]], new_size: Tuple[int, int]):
    """resize_boxes (transform.py:278-292) of every target for the fixed-size resize: x * (new_w / w), y * (new_h / h)."""
    out = []
    for t, (h, w) in zip(targets, original_sizes):
        if (h, w) == tuple(new_size):
            out.append(t)
            continue
        b = t["boxes"]
        rh = torch.tensor(new_size[0], dtype=torch.float32, device=b.device) / torch.tensor(h, dtype=torch.float32, device=b.device)
        rw = torch.tensor(new_size[1], dtype=torch.float32, device=b.device) / torch.tensor(w, dtype=torch.float32, device=b.device)
        xmin, ymin, xmax, ymax = b.unbind(1)
        nt = dict(t)
        nt["boxes"] = torch.stack((xmin * rw, ymin * rh, xmax * rw, ymax * rh), dim=1)
        out.append(nt)
    return out


class SSDMatcher(object):
    """_utils.SSDMatcher: Matcher(threshold, threshold, allow_low_quality_matches=False) plus the forced match of every
    ground-truth element to its best prediction.  `__call__(match_quality_matrix [M,N]) -> matches int64 [N]`."""

    BELOW_LOW_THRESHOLD = -1
    BETWEEN_THRESHOLDS = -2

    def __init__(self, threshold: float):
        self.high_threshold = threshold
        self.low_threshold = threshold
        self.allow_low_quality_matches = False

    def __call__(self, match_quality_matrix: Tensor) -> Tensor:
        _require_cuda(match_quality_matrix)
        if match_quality_matrix.dim() != 2:
            raise ValueError("match_quality_matrix must be MxN")
        q = match_quality_matrix.detach().to(torch.float32).contiguous()
        M, N = q.shape
        lib = _C.lib()
        out = torch.empty((N,), dtype=torch.int64, device=q.device)
        ws = torch.empty(max(4 * M, 4), dtype=torch.uint8, device=q.device)
        with torch.cuda.device(q.device):
            _C.check(lib.dn_match_quality(q.data_ptr(), M, N, float(self.low_threshold), out.data_ptr(), ws.data_ptr(),
                                          ws.numel(), _stream(q)), lib)
        return out


def match_targets(targets: Sequence[Dict[str, Tensor]], anchors, iou_thresh: float = 0.5) -> Tensor:
    """box_iou + SSDMatcher for every image of the batch in one launch: matched_idxs int64 [B,P] (generalized_ssd.py:326-335;
    an image without boxes gets -1 everywhere)."""
    device = anchors[0].device if isinstance(anchors, (list, tuple)) else anchors.device
    a = _anchors_2d(anchors, device)
    _require_cuda(a)
    boxes, _, offsets, G = _pack_targets(targets, device)
    B, P = len(targets), a.shape[0]
    lib = _C.lib()
    out = torch.empty((B, P), dtype=torch.int64, device=device)
    ws = torch.empty(lib.dn_ssd_loss_workspace_bytes(B, P, G), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _C.check(lib.dn_ssd_match(boxes.data_ptr(), offsets.data_ptr(), a.data_ptr(), B, P, G, float(iou_thresh), out.data_ptr(),
                                  ws.data_ptr(), ws.numel(), _stream(a)), lib)
    return out


class _SSDLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls_logits, bbox_regression, anchors, boxes, labels, offsets, matched, G, ratio, weights):
        B, P, K = cls_logits.shape
        lib = _C.lib()
        dev = cls_logits.device
        cls = cls_logits.detach().to(torch.float32).contiguous()
        reg = bbox_regression.detach().to(torch.float32).contiguous()
        need_grad = cls_logits.requires_grad or bbox_regression.requires_grad
        losses = torch.empty(3, dtype=torch.float32, device=dev)
        gcls = torch.empty_like(cls) if need_grad else None
        greg = torch.empty_like(reg) if need_grad else None
        ws = torch.empty(lib.dn_ssd_loss_workspace_bytes(B, P, G), dtype=torch.uint8, device=dev)
        w4 = (ctypes.c_float * 4)(*[float(w) for w in weights])
        with torch.cuda.device(dev):
            _C.check(lib.dn_ssd_loss(cls.data_ptr(), reg.data_ptr(), anchors.data_ptr(), boxes.data_ptr(), labels.data_ptr(),
                                     offsets.data_ptr(), matched.data_ptr(), B, P, K, G, float(ratio), w4, losses.data_ptr(),
                                     gcls.data_ptr() if need_grad else None, greg.data_ptr() if need_grad else None,
                                     ws.data_ptr(), ws.numel(), _stream(cls)), lib)
        if need_grad:
            ctx.save_for_backward(gcls, greg)
        ctx.in_dtypes = (cls_logits.dtype, bbox_regression.dtype)
        bbox, cls_l, n = losses[0].clone(), losses[1].clone(), losses[2].clone()
        ctx.mark_non_differentiable(n)
        return bbox, cls_l, n

    @staticmethod
    def backward(ctx, g_bbox, g_cls, _g_n):
        gcls, greg = ctx.saved_tensors
        d_cls = (gcls * g_cls).to(ctx.in_dtypes[0]) if ctx.needs_input_grad[0] else None
        d_reg = (greg * g_bbox).to(ctx.in_dtypes[1]) if ctx.needs_input_grad[1] else None
        return d_cls, d_reg, None, None, None, None, None, None, None, None


def compute_loss(targets: Sequence[Dict[str, Tensor]], head_outputs: Dict[str, Tensor], anchors, matched_idxs,
                 neg_to_pos_ratio: float = 3.0, box_weights: Sequence[float] = BOX_WEIGHTS) -> Dict[str, Tensor]:
    """SSD.compute_loss (generalized_ssd.py:210-269).  `matched_idxs`: the list of per-image int64 [P] tensors the reference
    passes, or one [B,P] tensor (match_targets).  Returns {'bbox_regression', 'classification'} (0-d tensors)."""
    cls_logits = head_outputs["cls_logits"]
    bbox_regression = head_outputs["bbox_regression"]
    _require_cuda(cls_logits, bbox_regression)
    if cls_logits.dim() != 3 or bbox_regression.shape != cls_logits.shape[:2] + (4,):
        raise ValueError("cls_logits must be [B,P,K] and bbox_regression [B,P,4]")
    device = cls_logits.device
    B, P, _ = cls_logits.shape
    if len(targets) != B:
        raise ValueError("one target per image is expected")
    if isinstance(matched_idxs, (list, tuple)):
        matched_idxs = torch.stack(list(matched_idxs))
    matched = matched_idxs.to(device, torch.int64).contiguous()
    if matched.shape != (B, P):
        raise ValueError("matched_idxs must be [B,P]")
    a = _anchors_2d(anchors, device)
    boxes, labels, offsets, G = _pack_targets(targets, device)
    bbox, cls, _ = _SSDLoss.apply(cls_logits, bbox_regression, a, boxes, labels, offsets, matched, G, float(neg_to_pos_ratio),
                                  tuple(box_weights))
    return {"bbox_regression": bbox, "classification": cls}


def detector_losses(model, images: Tensor, targets, original_sizes: Optional[Sequence[Tuple[int, int]]] = None,
                    iou_thresh: float = 0.5, neg_to_pos_ratio: float = 3.0) -> Dict[str, Tensor]:
    """The training branch of SSD.forward (generalized_ssd.py:271-337) for a detector of this package: target checks,
    head outputs of the [B,3,S,S] batch, matching, loss.  `model.head_outputs(batch)` provides (cls_logits, bbox_regression)."""
    check_targets(targets)
    S = tuple(model.size)
    if original_sizes is not None:
        targets = resize_targets(targets, original_sizes, S)
    cls, reg = model.head_outputs(images)[:2]
    anchors = model.anchors(images.device)
    matched = match_targets(targets, anchors, iou_thresh)
    return compute_loss(targets, {"cls_logits": cls, "bbox_regression": reg}, anchors, matched, neg_to_pos_ratio)
