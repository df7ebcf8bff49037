"""Stage-level operators on CUDA tensors, each a thin wrapper over one C-ABI entry point.

They keep the reference's names / argument meaning where the reference has a counterpart:
  batched_nms(boxes, scores, idxs, iou_threshold)   torchvision.ops.batched_nms as called at
                                                    generalized_ssd.py:389, box_head.py:374
  nms(boxes, scores, iou_threshold)                 torchvision.ops.nms
  postprocess_detections(head_outputs, anchors, image_shape, ...)   SSD.postprocess_detections
  PostProcess(...)                                   box_head.PostProcess (legacy V2 flavour)
  DefaultBoxGenerator                                anchor_utils.DefaultBoxGenerator (table)
  resize_bilinear / u8_to_f32 / rescale_boxes_       GeneralizedRCNNTransform resize, ToTensor, resize_boxes (transform.py)
  detections_to_coco / coco_results                  CocoEvaluator.prepare_for_coco_detection (data/coco_eval.py:76-98)
Everything raises if the CUDA library is missing or the tensors are not on a CUDA device.
"""
import ctypes
from typing import Dict, List, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _C
from .module import make_post_params, rescale_boxes_, resize_bilinear, u8_to_f32      # noqa: F401 (re-exported)


def _stream(t: Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


class _CL:
    """Call proxy: the library built for the activation dtype of `x`, with the status check on the same handle."""

    def __init__(self, x):
        self._h = _C.lib(_C.dtype_name(x.dtype))

    def __getattr__(self, name):
        fn, h = getattr(self._h, name), self._h

        def call(*args):
            _C.check(fn(*args), h)
        return call


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("demonet_b200 operators run on CUDA tensors only (no CPU fallback)")


# ---- NMS ---------------------------------------------------------------------------------------
def batched_nms(boxes: Tensor, scores: Tensor, idxs: Tensor, iou_threshold: float) -> Tensor:
    """Per-class NMS; returns int64 indices of kept boxes sorted by decreasing score.  Bit-exact
    against torchvision's CPU kernel with `_batched_nms_vanilla` semantics."""
    _require_cuda(boxes, scores, idxs)
    if boxes.dim() != 2 or boxes.shape[1] != 4:
        raise ValueError("boxes should be a 2d tensor of shape [N, 4], got {}".format(tuple(boxes.shape)))
    n = boxes.shape[0]
    if scores.shape != (n,) or idxs.shape != (n,):
        raise ValueError("scores and idxs must have shape [N]")
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    boxes = boxes.detach().float().contiguous()
    scores = scores.detach().float().contiguous()
    idxs = idxs.detach().to(torch.int64).contiguous()
    lib = _C.lib()
    ws = torch.empty(lib.dn_batched_nms_workspace_bytes(n), dtype=torch.uint8, device=boxes.device)
    keep = torch.empty(n, dtype=torch.int64, device=boxes.device)
    nkeep = torch.zeros(1, dtype=torch.int64, device=boxes.device)
    with torch.cuda.device(boxes.device):
        _C.check(lib.dn_batched_nms(boxes.data_ptr(), scores.data_ptr(), idxs.data_ptr(), n, float(iou_threshold),
                                    ws.data_ptr(), ws.numel(), keep.data_ptr(), nkeep.data_ptr(), _stream(boxes)))
    k = int(nkeep.item())
    if k == -1:
        raise ValueError("batched_nms: class indices must lie in [0, 4096)")
    if k == -2:
        raise NotImplementedError("batched_nms: more than 4095 boxes kept in one class")
    return keep[:k]


def nms(boxes: Tensor, scores: Tensor, iou_threshold: float) -> Tensor:
    return batched_nms(boxes, scores, torch.zeros(boxes.shape[0], dtype=torch.int64, device=boxes.device),
                       iou_threshold)


# ---- post-processing ---------------------------------------------------------------------------
def _postprocess(cls_logits, bbox_regression, anchors, image_shape, score_thresh, nms_thresh, detections_per_img,
                 topk_candidates, min_box_size):
    _require_cuda(cls_logits, bbox_regression, anchors)
    if cls_logits.dim() != 3 or bbox_regression.dim() != 3 or bbox_regression.shape[-1] != 4:
        raise ValueError("expected cls_logits [B,P,K] and bbox_regression [B,P,4]")
    B, P, K = cls_logits.shape
    if bbox_regression.shape[:2] != (B, P) or tuple(anchors.shape) != (P, 4):
        raise ValueError("shape mismatch between logits, box regression and anchors")
    cls_logits = cls_logits.detach().float().contiguous()
    bbox_regression = bbox_regression.detach().float().contiguous()
    anchors = anchors.detach().float().contiguous()
    dev = cls_logits.device
    prm = make_post_params(P, K, int(image_shape[0]), int(image_shape[1]), score_thresh, nms_thresh, topk_candidates,
                           detections_per_img, min_box_size)
    lib = _C.lib()
    D = detections_per_img
    ws = torch.empty(lib.dn_postprocess_workspace_bytes(B, ctypes.byref(prm)), dtype=torch.uint8, device=dev)
    boxes = torch.empty(B, D, 4, dtype=torch.float32, device=dev)
    scores = torch.empty(B, D, dtype=torch.float32, device=dev)
    labels = torch.empty(B, D, dtype=torch.int64, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _C.check(lib.dn_postprocess(cls_logits.data_ptr(), bbox_regression.data_ptr(), anchors.data_ptr(), B,
                                    ctypes.byref(prm), ws.data_ptr(), ws.numel(), boxes.data_ptr(), scores.data_ptr(),
                                    labels.data_ptr(), counts.data_ptr(), _stream(cls_logits)))
    return boxes, scores, labels, counts


def postprocess_padded(cls_logits, bbox_regression, anchors, image_shape, score_thresh=0.001, nms_thresh=0.55,
                       detections_per_img=300, topk_candidates=300, min_box_size=-1.0):
    """Fixed-shape outputs (boxes [B,D,4], scores [B,D], labels [B,D], counts [B]); no host sync."""
    return _postprocess(cls_logits, bbox_regression, anchors, image_shape, score_thresh, nms_thresh,
                        detections_per_img, topk_candidates, min_box_size)


def postprocess_scored(scores, boxes, image_shape, score_thresh=0.001, nms_thresh=0.55, detections_per_img=300,
                       topk_candidates=300, min_box_size=-1.0):
    """SSD.postprocess_detections from the point where the reference holds softmax `scores` [B,P,K] and decoded, clipped
    `boxes` [B,P,4] (generalized_ssd.py:361-396), through the kernels the engine runs (class sort, lazy warp / CTA NMS,
    top-D merge).  Returns padded (boxes [B,D,4], scores [B,D], labels [B,D], counts [B], priors int32 [B,D], rounds
    int32 [B]): (priors, labels) is the (anchor, class) identity of keep[:detections_per_img]."""
    _require_cuda(scores, boxes)
    if scores.dim() != 3 or boxes.dim() != 3 or boxes.shape[-1] != 4 or boxes.shape[:2] != scores.shape[:2]:
        raise ValueError("expected scores [B,P,K] and boxes [B,P,4]")
    B, P, K = scores.shape
    scores = scores.detach().float().contiguous()
    boxes = boxes.detach().float().contiguous()
    dev = scores.device
    prm = make_post_params(P, K, int(image_shape[0]), int(image_shape[1]), score_thresh, nms_thresh, topk_candidates,
                           detections_per_img, min_box_size)
    lib = _C.lib()
    D = detections_per_img
    ws = torch.empty(lib.dn_postprocess_workspace_bytes(B, ctypes.byref(prm)), dtype=torch.uint8, device=dev)
    out_boxes = torch.empty(B, D, 4, dtype=torch.float32, device=dev)
    out_scores = torch.empty(B, D, dtype=torch.float32, device=dev)
    out_labels = torch.empty(B, D, dtype=torch.int64, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    priors = torch.empty(B, D, dtype=torch.int32, device=dev)
    rounds = torch.zeros(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _C.check(lib.dn_postprocess_scored(scores.data_ptr(), boxes.data_ptr(), B, ctypes.byref(prm), ws.data_ptr(),
                                           ws.numel(), out_boxes.data_ptr(), out_scores.data_ptr(), out_labels.data_ptr(),
                                           counts.data_ptr(), priors.data_ptr(), rounds.data_ptr(), _stream(scores)))
    return out_boxes, out_scores, out_labels, counts, priors, rounds


def _to_dicts(boxes, scores, labels, counts) -> List[Dict[str, Tensor]]:
    out = []
    for i, n in enumerate(counts.tolist()):
        out.append({"boxes": boxes[i, :n], "scores": scores[i, :n], "labels": labels[i, :n]})
    return out


def postprocess_detections(head_outputs: Dict[str, Tensor], image_anchors, image_shapes, score_thresh=0.001,
                           nms_thresh=0.55, detections_per_img=300, topk_candidates=300) -> List[Dict[str, Tensor]]:
    """SSD.postprocess_detections (generalized_ssd.py:351-397): same inputs (head_outputs dict,
    per-image anchors list or one [P,4] table, image_shapes) and the same list-of-dicts result."""
    anchors = image_anchors[0] if isinstance(image_anchors, (list, tuple)) else image_anchors
    shape = image_shapes[0] if isinstance(image_shapes, (list, tuple)) and isinstance(image_shapes[0], (list, tuple)) \
        else image_shapes
    return _to_dicts(*_postprocess(head_outputs["cls_logits"], head_outputs["bbox_regression"], anchors, shape,
                                   score_thresh, nms_thresh, detections_per_img, topk_candidates, -1.0))


class PostProcess(torch.nn.Module):
    """box_head.PostProcess (box_head.py:298-381): legacy V2 post-processor -- no per-class top-k,
    remove_small_boxes(min_size=1e-2).  `variances` (0.1, 0.2) == BoxCoder weights (10,10,5,5)."""

    def __init__(self, variances: Tuple[float, float] = (0.1, 0.2), score_thresh: float = 0.5,
                 nms_thresh: float = 0.45, detections_per_img: int = 100):
        super().__init__()
        if tuple(variances) != (0.1, 0.2):
            raise NotImplementedError("only variances (0.1, 0.2) are supported")
        self.score_thresh, self.nms_thresh, self.detections_per_img = score_thresh, nms_thresh, detections_per_img

    def forward(self, pred_logits: Tensor, pred_boxes: Tensor, priors: Tensor, image_shapes) -> List[Dict[str, Tensor]]:
        shape = image_shapes[0] if isinstance(image_shapes[0], (list, tuple, torch.Size)) else image_shapes
        return _to_dicts(*_postprocess(pred_logits, pred_boxes, priors, shape, self.score_thresh, self.nms_thresh,
                                       self.detections_per_img, 0, 1e-2))


class DefaultBoxGenerator(torch.nn.Module):
    """anchor_utils.DefaultBoxGenerator (anchor_utils.py:10-126) as a per-(grid sizes, image size)
    cached table: the boxes depend only on static shapes, so they are computed once on the host
    with the reference's fp32 operation order and uploaded."""

    def __init__(self, aspect_ratios, min_ratio=0.15, max_ratio=0.9, clip=True):
        super().__init__()
        self.aspect_ratios, self.min_ratio, self.max_ratio, self.clip = aspect_ratios, min_ratio, max_ratio, clip
        self._cache = {}

    def num_anchors_per_location(self):
        return [2 + 2 * len(r) for r in self.aspect_ratios]

    def table(self, grid_sizes, image_size, device) -> Tensor:
        from . import plan as _plan
        key = (tuple(map(tuple, grid_sizes)), tuple(image_size), str(device))
        if key not in self._cache:
            if image_size[0] != image_size[1]:
                raise NotImplementedError("square inputs only")
            p = _plan.Plan("anchors", int(image_size[0]), 2, 0.0)
            for i, (h, w) in enumerate(grid_sizes):
                p.tensors["f%d" % i] = (int(h), int(w), 0)
                p.feature_names.append("f%d" % i)
            t = _plan.default_boxes(p, self.aspect_ratios, self.min_ratio, self.max_ratio, self.clip)
            self._cache[key] = torch.from_numpy(t).to(device)
        return self._cache[key]

    def forward(self, image_list, feature_maps: List[Tensor]) -> List[Tensor]:
        grid_sizes = [tuple(f.shape[-2:]) for f in feature_maps]
        image_size = tuple(image_list.tensors.shape[-2:])
        t = self.table(grid_sizes, image_size, feature_maps[0].device)
        return [t for _ in image_list.image_sizes]


# ---- conv stages (NHWC fp16 / bf16: the library is chosen by the dtype of x) -----------------
def dwconv(x: Tensor, w: Tensor, bias: Tensor, k: int, stride: int, act: str) -> Tensor:
    """x fp16/bf16 [B,H,W,C]; w fp32 [k*k,C]; bias fp32 [C] -> fp16/bf16 [B,Ho,Wo,C]."""
    _require_cuda(x, w, bias)
    B, H, W, C = x.shape
    pad = (k - 1) // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.empty(B, Ho, Wo, C, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _CL(x).dn_dwconv(x.contiguous().data_ptr(), w.contiguous().data_ptr(), bias.contiguous().data_ptr(),
                                    y.data_ptr(), B, H, W, C, k, stride, _C.ACT[act], _stream(x))
    return y


def pwconv(x: Tensor, w: Tensor, bias: Tensor, act: str = "none", residual: Tensor = None, out_fp32: bool = False,
           impl: int = 0) -> Tensor:
    """x fp16/bf16 [M,K]; w fp16/bf16 [N,K]; bias fp32 [N]; residual fp16/bf16 [M,N] -> [M,N] fp16/bf16 (or fp32)."""
    _require_cuda(x, w, bias, residual)
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty(M, N, dtype=torch.float32 if out_fp32 else x.dtype, device=x.device)
    x, w, bias = x.contiguous(), w.contiguous(), bias.contiguous()
    res = residual.contiguous() if residual is not None else None
    with torch.cuda.device(x.device):
        _CL(x).dn_pwconv(x.data_ptr(), w.data_ptr(), bias.data_ptr(), res.data_ptr() if res is not None else None,
                                    y.data_ptr(), M, K, N, _C.ACT[act], int(out_fp32), M, 0, N, impl, _stream(x))
    return y


def pwdw_fused(x: Tensor, w_pw: Tensor, b_pw: Tensor, w_dw: Tensor, b_dw: Tensor, k: int, stride: int, act_pw: str,
               act_dw: str) -> Tensor:
    """x fp16/bf16 [B,H,W,K] -> act_dw(dw(act_pw(x . w_pw^T + b_pw))) fp16/bf16 [B,Ho,Wo,N] with the expanded tensor kept on chip.
    w_pw fp16/bf16 [N,K]; w_dw fp32 [k*k,N]."""
    _require_cuda(x, w_pw, b_pw, w_dw, b_dw)
    B, H, W, K = x.shape
    N = w_pw.shape[0]
    Ho, Wo = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
    y = torch.empty(B, Ho, Wo, N, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _CL(x).dn_pwdw_fused(x.contiguous().data_ptr(), w_pw.contiguous().data_ptr(), b_pw.contiguous().data_ptr(),
                                        w_dw.contiguous().data_ptr(), b_dw.contiguous().data_ptr(), y.data_ptr(), B, H, W, K, N,
                                        k, stride, _C.ACT[act_pw], _C.ACT[act_dw], _stream(x))
    return y


def dwpw_fused(x: Tensor, w_dw: Tensor, b_dw: Tensor, w_pw: Tensor, b_pw: Tensor, k: int, stride: int, act_dw: str,
               residual: bool) -> Tensor:
    """x fp16/bf16 [B,H,W,C] -> (act_dw(dw(x)) . w_pw^T + b_pw) (+ x) fp16/bf16 [B,H,W,N] with the depthwise output kept on chip."""
    _require_cuda(x, w_dw, b_dw, w_pw, b_pw)
    B, H, W, C = x.shape
    N = w_pw.shape[0]
    y = torch.empty(B, H, W, N, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _CL(x).dn_dwpw_fused(x.contiguous().data_ptr(), w_dw.contiguous().data_ptr(), b_dw.contiguous().data_ptr(),
                                        w_pw.contiguous().data_ptr(), b_pw.contiguous().data_ptr(), y.data_ptr(), B, H, W, C, N,
                                        k, stride, _C.ACT[act_dw], int(residual), _stream(x))
    return y


def stem_conv(images: Tensor, w: Tensor, bias: Tensor, mean, std, act: str, act_dtype: str = None) -> Tensor:
    """images fp32 [B,3,H,W]; w fp32 [27,Cout]; -> fp16 / fp16/bf16 [B,Ho,Wo,Cout] (normalise + 3x3 s2 + act)."""
    _require_cuda(images, w, bias)
    B, _, H, W = images.shape
    Cout = w.shape[1]
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty(B, Ho, Wo, Cout, dtype=_C.torch_dtype(act_dtype), device=images.device)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    with torch.cuda.device(images.device):
        _CL(y).dn_stem_conv(images.contiguous().data_ptr(), w.contiguous().data_ptr(),
                                       bias.contiguous().data_ptr(), m, s, y.data_ptr(), B, H, W, Cout, _C.ACT[act],
                                       _stream(images))
    return y


def se_inplace(x: Tensor, w1: Tensor, b1: Tensor, w2t: Tensor, b2: Tensor) -> Tensor:
    """x fp16/bf16 [B,HW,C] scaled in place; w1 fp32 [Cs,C]; w2t fp32 [Cs,C] (fc2 transposed)."""
    _require_cuda(x, w1, b1, w2t, b2)
    B, HW, C = x.shape
    ws_bytes = _C.lib().dn_se_workspace_bytes(B, HW, C)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _CL(x).dn_se_inplace(x.data_ptr(), w1.contiguous().data_ptr(), b1.contiguous().data_ptr(),
                                        w2t.contiguous().data_ptr(), b2.contiguous().data_ptr(), B, HW, C, w1.shape[0],
                                        ws.data_ptr(), ws_bytes, _stream(x))
    return x


def se_project(x: Tensor, w1: Tensor, b1: Tensor, w2t: Tensor, b2: Tensor, w_pw: Tensor, b_pw: Tensor,
               residual: Tensor = None) -> Tensor:
    """Squeeze-excitation of x [B,HW,C] folded into the project GEMM behind it: y [B*HW,N] = (x * scale) . w_pw^T + b_pw
    (+ residual), the scaling done on the GEMM's A operand in shared memory; x is not modified.  Bit-identical to
    se_inplace followed by pwconv."""
    _require_cuda(x, w1, b1, w2t, b2, w_pw, b_pw, residual)
    B, HW, C = x.shape
    N = w_pw.shape[0]
    y = torch.empty(B * HW, N, dtype=x.dtype, device=x.device)
    ws_bytes = _C.lib().dn_se_workspace_bytes(B, HW, C)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
    res = residual.contiguous() if residual is not None else None
    with torch.cuda.device(x.device):
        _CL(x).dn_se_project(x.contiguous().data_ptr(), w1.contiguous().data_ptr(), b1.contiguous().data_ptr(),
                             w2t.contiguous().data_ptr(), b2.contiguous().data_ptr(), w_pw.contiguous().data_ptr(),
                             b_pw.contiguous().data_ptr(), res.data_ptr() if res is not None else None, y.data_ptr(), B, HW, C,
                             w1.shape[0], N, ws.data_ptr(), ws_bytes, _stream(x))
    return y


def dwconv_se(x: Tensor, w: Tensor, bias: Tensor, k: int, stride: int, act: str, w1: Tensor, b1: Tensor, w2t: Tensor,
              b2: Tensor):
    """Depthwise conv + squeeze-excitation of its output (the middle of an InvertedResidual with use_se,
    mobilenetv3.py:43-96).  Returns (y fp16/bf16 [B,Ho,Wo,C], pooled): pooled tells whether the depthwise launch produced the
    SE channel sums itself (stride-1 row stream, large batch) or the SE ran its own pooling pass."""
    _require_cuda(x, w, bias, w1, b1, w2t, b2)
    B, H, W, C = x.shape
    pad = (k - 1) // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.empty(B, Ho, Wo, C, dtype=x.dtype, device=x.device)
    ws_bytes = _C.lib().dn_se_workspace_bytes(B, Ho * Wo, C)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
    pooled = ctypes.c_int(0)
    with torch.cuda.device(x.device):
        _CL(x).dn_dwconv_se(x.contiguous().data_ptr(), w.contiguous().data_ptr(), bias.contiguous().data_ptr(),
                                       y.data_ptr(), B, H, W, C, k, stride, _C.ACT[act], w1.contiguous().data_ptr(),
                                       b1.contiguous().data_ptr(), w2t.contiguous().data_ptr(), b2.contiguous().data_ptr(),
                                       w1.shape[0], ws.data_ptr(), ws_bytes, ctypes.byref(pooled), _stream(x))
    return y, bool(pooled.value)


# ---- dense convolutions and the small operators of ssd300_vgg16 ------------------------------------
def conv3x3(x: Tensor, w: Tensor, bias: Tensor, stride: int = 1, padding: int = 1, dilation: int = 1, act: str = "relu",
            out: Tensor = None, out_batch_stride: int = 0, out_row_stride: int = 0) -> Tensor:
    """Dense 3x3 convolution on the tensor cores.  x 16-bit [B,H,W,C] NHWC (C % 64 == 0); w 16-bit [9,N,C] (tap-major:
    w[kh*3+kw, n, c] = conv.weight[n, c, kh, kw]); bias fp32 [N] -> act(conv) as 16-bit [B,Ho,Wo,N].  With `out` (an fp32
    tensor) the result is written with the SSD head addressing (see pwconv / dn_pwconv) and `out` is returned."""
    _require_cuda(x, w, bias, out)
    B, H, W, C = x.shape
    N = w.shape[1]
    Ho = (H + 2 * padding - 2 * dilation - 1) // stride + 1
    Wo = (W + 2 * padding - 2 * dilation - 1) // stride + 1
    y = out if out is not None else torch.empty(B, Ho, Wo, N, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _CL(x).dn_conv3x3(x.contiguous().data_ptr(), w.contiguous().data_ptr(), bias.contiguous().data_ptr(), y.data_ptr(), B, H, W, C,
                          N, stride, padding, dilation, _C.ACT[act], int(out is not None), out_batch_stride, out_row_stride,
                          _stream(x))
    return y


def conv3x3_first(images: Tensor, w: Tensor, bias: Tensor, mean, std, act_dtype: str = None) -> Tensor:
    """normalise + conv 3 -> 64 (3x3, stride 1, padding 1) + ReLU.  images fp32 [B,3,H,W]; w fp32 [27,64]."""
    _require_cuda(images, w, bias)
    B, _, H, W = images.shape
    y = torch.empty(B, H, W, w.shape[1], dtype=_C.torch_dtype(act_dtype), device=images.device)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    with torch.cuda.device(images.device):
        _CL(y).dn_conv3x3_first(images.contiguous().data_ptr(), w.contiguous().data_ptr(), bias.contiguous().data_ptr(), m, s,
                                y.data_ptr(), B, H, W, w.shape[1], _stream(images))
    return y


def im2col3x3_first(images: Tensor, mean, std, act_dtype: str = None) -> Tensor:
    """Normalised 3x3 neighbourhoods of a fp32 [B,3,H,W] batch as 16-bit rows [B*H*W, 32] (27 taps + 5 zeros)."""
    _require_cuda(images)
    B, _, H, W = images.shape
    cols = torch.empty(B * H * W, 32, dtype=_C.torch_dtype(act_dtype), device=images.device)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    with torch.cuda.device(images.device):
        _CL(cols).dn_im2col3x3_first(images.contiguous().data_ptr(), m, s, cols.data_ptr(), B, H, W, _stream(images))
    return cols


def maxpool2d(x: Tensor, k: int, stride: int, padding: int = 0, ceil_mode: bool = False) -> Tensor:
    """nn.MaxPool2d on 16-bit NHWC activations."""
    _require_cuda(x)
    B, H, W, C = x.shape

    def out(n):
        o = (n + 2 * padding - k + (stride - 1 if ceil_mode else 0)) // stride + 1
        if ceil_mode and (o - 1) * stride >= n + padding:
            o -= 1
        return o
    y = torch.empty(B, out(H), out(W), C, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _CL(x).dn_maxpool2d(x.contiguous().data_ptr(), y.data_ptr(), B, H, W, C, k, stride, padding, int(ceil_mode), _stream(x))
    return y


def l2norm_scale(x: Tensor, scale: Tensor) -> Tensor:
    """scale * F.normalize(x) over the last (channel) dimension of 16-bit NHWC activations."""
    _require_cuda(x, scale)
    C = x.shape[-1]
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _CL(x).dn_l2norm_scale(x.contiguous().data_ptr(), scale.float().contiguous().data_ptr(), y.data_ptr(), x.numel() // C, C,
                               _stream(x))
    return y


# ---- detection sink ---------------------------------------------------------------
def detections_to_coco(boxes: Tensor, scores: Tensor, labels: Tensor, counts: Tensor, image_ids) -> Dict[str, Tensor]:
    """Padded detections (boxes [B,D,4] xyxy, scores [B,D], labels [B,D], counts [B], as written by the engine) ->
    compact COCO rows in image order: {"image_id" i64[n], "category_id" i64[n], "bbox" f32[n,4] xywh, "score" f32[n]}.
    The device-side CocoEvaluator.prepare_for_coco_detection (demonet/data/coco_eval.py:76-98)."""
    _require_cuda(boxes, scores, labels, counts)
    B, D = boxes.shape[0], boxes.shape[1]
    dev = boxes.device
    ids = torch.as_tensor(image_ids, dtype=torch.int64).to(dev).contiguous()
    if ids.shape != (B,):
        raise ValueError("image_ids must have one id per image")
    out = {"image_id": torch.empty(B * D, dtype=torch.int64, device=dev),
           "category_id": torch.empty(B * D, dtype=torch.int64, device=dev),
           "bbox": torch.empty(B * D, 4, dtype=torch.float32, device=dev),
           "score": torch.empty(B * D, dtype=torch.float32, device=dev)}
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    if B == 0:
        return {k: v[:0] for k, v in out.items()}
    with torch.cuda.device(dev):
        _C.check(_C.lib().dn_detections_to_coco(
            boxes.contiguous().data_ptr(), scores.contiguous().data_ptr(), labels.to(torch.int64).contiguous().data_ptr(),
            counts.to(torch.int32).contiguous().data_ptr(), ids.data_ptr(), B, D, out["image_id"].data_ptr(),
            out["category_id"].data_ptr(), out["bbox"].data_ptr(), out["score"].data_ptr(), total.data_ptr(), _stream(boxes)))
    n = int(total.item())
    return {k: v[:n] for k, v in out.items()}


def coco_results(rows: Dict[str, Tensor]) -> List[dict]:
    """The list of dicts pycocotools' loadRes takes, from the rows of detections_to_coco (one D2H copy per column)."""
    ids, cats = rows["image_id"].tolist(), rows["category_id"].tolist()
    boxes, scores = rows["bbox"].tolist(), rows["score"].tolist()
    return [{"image_id": ids[k], "category_id": cats[k], "bbox": boxes[k], "score": scores[k]} for k in range(len(ids))]
