"""`torch.library` registration of the hot-path operators.

The reference treats scriptability / exportability as its main test (test/test_model.py:62-119,
test/tracing/test_demonet_tracing.cpp, export/onnx_export.py).  A ctypes-loaded library is invisible to
`torch.export` / `torch.compile`, so the stage-level entry points are registered here as custom operators with
fake (shape-only) implementations:

    torch.ops.demonet_b200.nms(boxes, scores, iou_threshold)                    -> int64 [m]   (data dependent)
    torch.ops.demonet_b200.batched_nms(boxes, scores, idxs, iou_threshold)      -> int64 [m]   (data dependent)
    torch.ops.demonet_b200.postprocess(cls_logits, bbox_regression, anchors, image_h, image_w, score_thresh,
                                       nms_thresh, detections_per_img, topk_candidates, min_box_size)
                                                                      -> boxes [B,D,4], scores [B,D], labels [B,D], counts [B]
    torch.ops.demonet_b200.ssdlite_forward(images, engine_id)         -> the same four padded tensors

`ExportableSSDLite` wraps the last one for `torch.export` / `torch.compile(fullgraph=True)`, `ScriptableSSDLite` for
`torch.jit.script` (List[Tensor] in, (losses, List[Dict]) out like the scripted reference).

`postprocess` and `ssdlite_forward` have static output shapes (padded to detections_per_img + counts), which is what
makes the exported graph free of data-dependent shapes.  `ssdlite_forward` addresses a live engine through the integer
id `SSDLiteB200.export_handle()` returns (an exported program captures it as a constant; the module must outlive it).
The CUDA implementations are the ones of `demonet_b200.ops`; there is no CPU kernel.
"""
from typing import Dict, List, Tuple

import torch
from torch import Tensor

from . import ops as _ops

_ENGINES: Dict[int, object] = {}          # engine_id -> (module, device, batch)


@torch.library.custom_op("demonet_b200::batched_nms", mutates_args=(), device_types="cuda")
def batched_nms(boxes: Tensor, scores: Tensor, idxs: Tensor, iou_threshold: float) -> Tensor:
    return _ops.batched_nms(boxes, scores, idxs, iou_threshold).clone()


@batched_nms.register_fake
def _(boxes, scores, idxs, iou_threshold):
    m = torch.library.get_ctx().new_dynamic_size()
    return boxes.new_empty((m,), dtype=torch.int64)


@torch.library.custom_op("demonet_b200::nms", mutates_args=(), device_types="cuda")
def nms(boxes: Tensor, scores: Tensor, iou_threshold: float) -> Tensor:
    return _ops.nms(boxes, scores, iou_threshold).clone()


@nms.register_fake
def _(boxes, scores, iou_threshold):
    m = torch.library.get_ctx().new_dynamic_size()
    return boxes.new_empty((m,), dtype=torch.int64)


@torch.library.custom_op("demonet_b200::postprocess", mutates_args=(), device_types="cuda")
def postprocess(cls_logits: Tensor, bbox_regression: Tensor, anchors: Tensor, image_h: int, image_w: int,
                score_thresh: float, nms_thresh: float, detections_per_img: int, topk_candidates: int,
                min_box_size: float) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    return _ops.postprocess_padded(cls_logits, bbox_regression, anchors, (image_h, image_w), score_thresh, nms_thresh,
                                   detections_per_img, topk_candidates, min_box_size)


@postprocess.register_fake
def _(cls_logits, bbox_regression, anchors, image_h, image_w, score_thresh, nms_thresh, detections_per_img,
      topk_candidates, min_box_size):
    B, D = cls_logits.shape[0], detections_per_img
    return (cls_logits.new_empty((B, D, 4), dtype=torch.float32), cls_logits.new_empty((B, D), dtype=torch.float32),
            cls_logits.new_empty((B, D), dtype=torch.int64), cls_logits.new_empty((B,), dtype=torch.int32))


@torch.library.custom_op("demonet_b200::ssdlite_forward", mutates_args=(), device_types="cuda")
def ssdlite_forward(images: Tensor, engine_id: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    model = _ENGINES.get(engine_id)
    if model is None:
        raise RuntimeError("demonet_b200::ssdlite_forward: unknown engine id %d (the module was deleted?)" % engine_id)
    return model.forward_padded(images)


@ssdlite_forward.register_fake
def _(images, engine_id):
    model = _ENGINES.get(engine_id)
    D = model.detections_per_img if model is not None else 300
    B = images.shape[0]
    return (images.new_empty((B, D, 4), dtype=torch.float32), images.new_empty((B, D), dtype=torch.float32),
            images.new_empty((B, D), dtype=torch.int64), images.new_empty((B,), dtype=torch.int32))


def register_engine(model) -> int:
    """Give `model` (an SSDLiteB200) an integer id usable as the `engine_id` argument of ssdlite_forward."""
    eid = id(model)
    _ENGINES[eid] = model
    return eid


class ExportableSSDLite(torch.nn.Module):
    """`forward(images [B,3,S,S]) -> (boxes [B,D,4], scores [B,D], labels [B,D], counts [B])` through the registered
    operator: the form of the detector that `torch.export.export` / `torch.compile(fullgraph=True)` can capture."""

    def __init__(self, model):
        super().__init__()
        self.engine_id = register_engine(model)
        self._model = [model]                   # keeps the engine alive without registering it as a submodule

    def forward(self, images: Tensor):
        return torch.ops.demonet_b200.ssdlite_forward(images, self.engine_id)


class ScriptableSSDLite(torch.nn.Module):
    """The detector in the form `torch.jit.script` accepts — the reference's own main test scripts its model and compares
    the scripted with the eager detections (test/test_model.py:85-119).  The arithmetic lives behind the C ABI, so the
    scripted graph is: stack the images, ONE call of the registered operator `demonet_b200::ssdlite_forward`, slice the
    padded outputs by the per-image counts.  Like the scripted reference (generalized_ssd.py:336-349, `eager_outputs` under
    scripting) the module returns `(losses, detections)` with an empty loss dict, scripted or not.
    Images must already have the network's S x S size (the resize of GeneralizedSSDTransform is done by
    `SSDLiteB200.forward`; a scripted caller resizes beforehand)."""

    def __init__(self, model):
        super().__init__()
        self.engine_id: int = register_engine(model)
        self.size: int = int(model.plan.size)
        self._model = [model]                   # keeps the engine alive without registering it as a submodule

    def forward(self, images: List[Tensor]) -> Tuple[Dict[str, Tensor], List[Dict[str, Tensor]]]:
        for img in images:
            if img.dim() != 3 or img.shape[0] != 3 or img.shape[1] != self.size or img.shape[2] != self.size:
                raise ValueError("ScriptableSSDLite expects images of shape [3, S, S] with S the network input size")
        batch = torch.stack(images, 0).to(torch.float32)
        boxes, scores, labels, counts = torch.ops.demonet_b200.ssdlite_forward(batch, self.engine_id)
        n: List[int] = counts.to(torch.int64).tolist()
        detections: List[Dict[str, Tensor]] = []
        for i in range(len(images)):
            k = n[i]
            detections.append({"boxes": boxes[i, :k], "scores": scores[i, :k], "labels": labels[i, :k]})
        losses: Dict[str, Tensor] = {}
        return losses, detections
