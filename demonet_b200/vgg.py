"""ssd300_vgg16 on the B200 kernels: the other detector of the reference that shares `SSD`, the
default boxes, the box decoding and the NMS with the SSDLite path (demonet/models/ssd_vgg16.py:30-213,
demonet/models/generalized_ssd.py:25-92).

Same constructor arguments, `state_dict` keys (71) and `forward(images) -> List[Dict[boxes, scores, labels]]` contract as the
reference, inference only.  Every convolution runs on the tcgen05 tensor cores -- the dense 3x3 ones as an implicit GEMM
(`dn_conv3x3`), the 1x1 ones on the pointwise GEMM of the SSDLite path (`dn_pwconv`) -- with 16-bit NHWC activations and
fp32 accumulation; the first convolution (3 -> 64) is `dn_im2col3x3_first` (normalisation folded in, 27 taps padded to one
32-wide k-block) + the pointwise GEMM, or the exact-fp32 SIMT kernel `dn_conv3x3_first` with `first_conv="direct"`; max-pooling and the
L2-normalisation of conv4_3 are small memory-bound kernels; the SSD heads write fp32 logits / box regression straight
into the [B, 8732, K] layout, and the post-processing is `dn_postprocess`, the kernels of the SSDLite path.
Unlike the SSDLite engine this model is driven layer by layer from Python (25 launches per batch, no CUDA graph): it is
tensor-bound (31 GMAC per image), not launch-bound.
"""
import os
import warnings
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _C, ops, plan as _plan
from .module import rescale_boxes_, resize_bilinear

# (kind, state_dict prefix, cin, cout, stride, padding, dilation) in execution order -- vgg16 `features` (cfg D) up to conv4_3,
# then SSDFeatureExtractorVGG.extra (ssd_vgg16.py:48-95); "tap" marks the six feature maps the heads read
_VGG = [("first", "backbone.features.0", 3, 64), ("conv", "backbone.features.2", 64, 64, 1, 1, 1), ("pool", 2, 2, 0, False),
        ("conv", "backbone.features.5", 64, 128, 1, 1, 1), ("conv", "backbone.features.7", 128, 128, 1, 1, 1), ("pool", 2, 2, 0, False),
        ("conv", "backbone.features.10", 128, 256, 1, 1, 1), ("conv", "backbone.features.12", 256, 256, 1, 1, 1),
        ("conv", "backbone.features.14", 256, 256, 1, 1, 1), ("pool", 2, 2, 0, True),          # ceil_mode patched in, ssd_vgg16.py:36-37
        ("conv", "backbone.features.17", 256, 512, 1, 1, 1), ("conv", "backbone.features.19", 512, 512, 1, 1, 1),
        ("conv", "backbone.features.21", 512, 512, 1, 1, 1), ("tap_l2norm",),                   # conv4_3: scale_weight * normalize(x)
        ("pool", 2, 2, 0, False),
        ("conv", "backbone.extra.0.1", 512, 512, 1, 1, 1), ("conv", "backbone.extra.0.3", 512, 512, 1, 1, 1),
        ("conv", "backbone.extra.0.5", 512, 512, 1, 1, 1), ("pool", 3, 1, 1, False),             # modified pool5, ssd_vgg16.py:84
        ("conv", "backbone.extra.0.7.1", 512, 1024, 1, 6, 6),                                    # fc6, atrous
        ("pw", "backbone.extra.0.7.3", 1024, 1024), ("tap",),                                    # fc7
        ("pw", "backbone.extra.1.0", 1024, 256), ("conv", "backbone.extra.1.2", 256, 512, 2, 1, 1), ("tap",),
        ("pw", "backbone.extra.2.0", 512, 128), ("conv", "backbone.extra.2.2", 128, 256, 2, 1, 1), ("tap",),
        ("pw", "backbone.extra.3.0", 256, 128), ("conv", "backbone.extra.3.2", 128, 256, 1, 0, 1), ("tap",),
        ("pw", "backbone.extra.4.0", 256, 128), ("conv", "backbone.extra.4.2", 128, 256, 1, 0, 1), ("tap",)]
_FEATURE_CHANNELS = [512, 1024, 512, 256, 256, 256]
_ASPECT_RATIOS = [[2], [2, 3], [2, 3], [2, 3], [2], [2]]
_SCALES = [0.07, 0.15, 0.33, 0.51, 0.69, 0.87, 1.05]
_STEPS = [8, 16, 32, 64, 100, 300]
_GRIDS = [(38, 38), (19, 19), (10, 10), (5, 5), (3, 3), (1, 1)]


class SSD300VGG16B200(nn.Module):
    """SSD300 / VGG16 on the B200 kernels; arguments mirror `SSD.__init__` (generalized_ssd.py:154-163).  Eval mode returns
    detections; training mode evaluates the loss branch (generalized_ssd.py:321-337) on the head outputs (`demonet_b200.loss`)."""

    size = (300, 300)

    def __init__(self, num_classes: int = 91, score_thresh: float = 0.01, nms_thresh: float = 0.45, detections_per_img: int = 200,
                 topk_candidates: int = 400, image_mean=None, image_std=None, act_dtype: Optional[str] = None, first_conv: Optional[str] = None,
                 iou_thresh: float = 0.5, positive_fraction: float = 0.25):
        super().__init__()
        self.iou_thresh = iou_thresh                                                                        # generalized_ssd.py:184
        self.neg_to_pos_ratio = (1.0 - positive_fraction) / positive_fraction                                # generalized_ssd.py:197
        self.first_conv = first_conv or os.environ.get("DN_VGG_FIRST", "gemm")
        if self.first_conv not in ("gemm", "direct"):
            raise ValueError("first_conv must be 'gemm' or 'direct'")
        self.num_classes = num_classes
        self.score_thresh, self.nms_thresh = score_thresh, nms_thresh
        self.detections_per_img, self.topk_candidates = detections_per_img, topk_candidates
        self.image_mean = list(image_mean) if image_mean is not None else [0.48235, 0.45882, 0.40784]      # ssd_vgg16.py:198
        self.image_std = list(image_std) if image_std is not None else [1.0 / 255.0] * 3                    # ssd_vgg16.py:199
        self.act_dtype = act_dtype or _C.DEFAULT_ACT_DTYPE
        if self.act_dtype not in _C.ACT_DTYPES:
            raise ValueError("act_dtype must be one of %s" % (_C.ACT_DTYPES,))
        self.num_anchors = [2 + 2 * len(r) for r in _ASPECT_RATIOS]                                          # [4, 6, 6, 6, 4, 4]
        self.num_priors = sum(h * w * a for (h, w), a in zip(_GRIDS, self.num_anchors))                      # 8732
        g = torch.Generator().manual_seed(0)
        self._param("backbone.scale_weight", torch.ones(512) * 20)                                          # ssd_vgg16.py:40
        for op in _VGG:
            if op[0] in ("first", "conv", "pw"):
                k = 1 if op[0] == "pw" else 3
                self._param(op[1] + ".weight", torch.randn(op[3], op[2], k, k, generator=g) * (2.0 / (op[2] * k * k)) ** 0.5)
                self._param(op[1] + ".bias", torch.zeros(op[3]))
        for name, cols in (("classification_head", num_classes), ("regression_head", 4)):
            for l, (c, a) in enumerate(zip(_FEATURE_CHANNELS, self.num_anchors)):
                self._param("head.%s.module_list.%d.weight" % (name, l), torch.randn(a * cols, c, 3, 3, generator=g) * (2.0 / (c * 9)) ** 0.5)
                self._param("head.%s.module_list.%d.bias" % (name, l), torch.zeros(a * cols))
        self._weights_epoch = 0
        self._packed: Dict[str, Tuple[int, dict]] = {}
        self._anchors: Dict[str, Tensor] = {}
        self.eval()

    def _param(self, key: str, value: Tensor):
        parts = key.split(".")
        mod = self
        for name in parts[:-1]:
            if name not in mod._modules:
                mod.add_module(name, nn.Module())
            mod = mod._modules[name]
        mod.register_parameter(parts[-1], nn.Parameter(value, requires_grad=False))

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._weights_epoch += 1
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._weights_epoch += 1
        return out

    def refresh_weights(self):
        self._weights_epoch += 1

    # ---- weights in kernel layout, cached per device until they change -----------------------------------------------
    def _weights(self, device) -> dict:
        key = str(device)
        hit = self._packed.get(key)
        if hit is not None and hit[0] == self._weights_epoch:
            return hit[1]
        sd = self.state_dict()
        h16 = _C.torch_dtype(self.act_dtype)
        out = {}

        def conv(prefix):
            w = sd[prefix + ".weight"].detach().float()
            n, c, k, _ = w.shape
            if c == 3:                                   # first layer: fp32 [27][64], (ci*3+kh)*3+kw major ...
                wk = w.permute(1, 2, 3, 0).reshape(27, n)
                w32 = torch.zeros(n, 32)                 # ... and its GEMM form, 16-bit [64][32] K-major (27 taps + 5 zeros)
                w32[:, :27] = w.reshape(n, 27)
                out[prefix + "/gemm"] = w32.to(h16).contiguous().to(device)
            elif k == 3:                                 # [9][N][C], tap = kh*3+kw major
                wk = w.permute(2, 3, 0, 1).reshape(9, n, c).to(h16)
            else:
                wk = w.reshape(n, c).to(h16)
            out[prefix] = (wk.contiguous().to(device), sd[prefix + ".bias"].detach().float().contiguous().to(device))
        for op in _VGG:
            if op[0] in ("first", "conv", "pw"):
                conv(op[1])
        for name in ("classification_head", "regression_head"):
            for l in range(6):
                conv("head.%s.module_list.%d" % (name, l))
        out["scale"] = sd["backbone.scale_weight"].detach().float().contiguous().to(device)
        self._packed[key] = (self._weights_epoch, out)
        return out

    def anchors(self, device) -> Tensor:
        key = str(device)
        if key not in self._anchors:
            t = _plan.default_boxes_for(_GRIDS, 300, _ASPECT_RATIOS, scales=_SCALES, steps=_STEPS)          # ssd_vgg16.py:193-195
            self._anchors[key] = torch.from_numpy(t).to(device)
        return self._anchors[key]

    # ---- the network: [B,3,300,300] fp32 CUDA batch -> (cls_logits [B,8732,K], bbox_regression [B,8732,4]) fp32 ----------
    def head_outputs(self, images: Tensor, return_features: bool = False):
        if not images.is_cuda:
            raise RuntimeError("demonet_b200 operators run on CUDA tensors only (no CPU fallback)")
        W = self._weights(images.device)
        B, K, P = images.shape[0], self.num_classes, self.num_priors
        x, feats = None, []
        for op in _VGG:
            kind = op[0]
            if kind == "first" and self.first_conv == "direct":
                x = ops.conv3x3_first(images, *W[op[1]], self.image_mean, self.image_std, act_dtype=self.act_dtype)
            elif kind == "first":
                cols = ops.im2col3x3_first(images, self.image_mean, self.image_std, act_dtype=self.act_dtype)
                x = ops.pwconv(cols, W[op[1] + "/gemm"], W[op[1]][1], act="relu").view(B, images.shape[2], images.shape[3], op[3])
            elif kind == "conv":
                x = ops.conv3x3(x, *W[op[1]], stride=op[4], padding=op[5], dilation=op[6], act="relu")
            elif kind == "pw":
                b, h, w_, c = x.shape
                x = ops.pwconv(x.view(b * h * w_, c), *W[op[1]], act="relu").view(b, h, w_, op[3])
            elif kind == "pool":
                x = ops.maxpool2d(x, op[1], op[2], op[3], op[4])
            elif kind == "tap_l2norm":
                feats.append(ops.l2norm_scale(x, W["scale"]))
            else:
                feats.append(x)
        cls = torch.empty(B, P, K, dtype=torch.float32, device=images.device)
        reg = torch.empty(B, P, 4, dtype=torch.float32, device=images.device)
        off = 0
        for l, (f, a) in enumerate(zip(feats, self.num_anchors)):
            # SSDScoringHead (generalized_ssd.py:60-74): conv3x3 -> view(N,A,K,H,W).permute(0,3,4,1,2).reshape(N,HWA,K), cat over
            # levels -- channel a*K+k of an NHWC row IS (a,k)-ordered, so the GEMM epilogue writes the final layout
            wq, bq = W["head.classification_head.module_list.%d" % l]
            ops.conv3x3(f, wq, bq, 1, 1, 1, "none", out=cls[:, off:], out_batch_stride=P * K, out_row_stride=a * K)
            wq, bq = W["head.regression_head.module_list.%d" % l]
            ops.conv3x3(f, wq, bq, 1, 1, 1, "none", out=reg[:, off:], out_batch_stride=P * 4, out_row_stride=a * 4)
            off += f.shape[1] * f.shape[2] * a
        return (cls, reg, feats) if return_features else (cls, reg)

    def forward(self, images: List[Tensor], targets: Optional[List[Dict[str, Tensor]]] = None):
        if isinstance(images, Tensor):
            if images.dim() != 4:
                raise ValueError("images is expected to be a list of 3d tensors of shape [C, H, W] "
                                 "or a batched 4d tensor, got {}".format(images.shape))
            images = list(images.unbind(0))
        if len(images) == 0:
            return []
        original_sizes = []
        for img in images:
            if img.dim() != 3:
                raise ValueError("images is expected to be a list of 3d tensors "
                                 "of shape [C, H, W], got {}".format(img.shape))       # transform.py:110-112
            if not img.is_floating_point():
                raise TypeError("Expected input images to be of floating type (in range [0, 1]), "
                                f"but found type {img.dtype} instead")                # transform.py:130-134
            original_sizes.append((int(img.shape[-2]), int(img.shape[-1])))
        if self.training and targets is None:
            raise ValueError("In training mode, targets should be passed")                 # generalized_ssd.py:273-274
        if self.training:
            from . import loss as _loss
            _loss.check_targets(targets)
        if not torch.cuda.is_available():
            raise RuntimeError("demonet_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        in_dev = images[0].device
        device = in_dev if in_dev.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        S = self.size[0]
        resized = any(sz != (S, S) for sz in original_sizes)
        if resized:
            batch = torch.empty(len(images), 3, S, S, dtype=torch.float32, device=device)
            for i, img in enumerate(images):
                if original_sizes[i] != (S, S):
                    resize_bilinear(img.to(device), (S, S), out=batch[i])          # transform.py:27-53
                else:
                    batch[i].copy_(img)
        else:
            batch = torch.stack([im.to(device, torch.float32) for im in images], 0)
        if self.training:
            # loss branch (generalized_ssd.py:321-337); parameters carry no gradients: loss evaluation (see demonet_b200.loss)
            targets = [{k: v.to(device) for k, v in t.items()} for t in targets]
            return _loss.detector_losses(self, batch, targets, original_sizes, self.iou_thresh, self.neg_to_pos_ratio)
        cls, reg = self.head_outputs(batch)
        boxes, scores, labels, counts = ops.postprocess_padded(cls, reg, self.anchors(device), self.size, self.score_thresh,
                                                               self.nms_thresh, self.detections_per_img, self.topk_candidates)
        if resized:
            rescale_boxes_(boxes, original_sizes, self.size)                       # transform.py:228-292
        out = []
        for i, n in enumerate(counts.tolist()):
            det = {"boxes": boxes[i, :n], "scores": scores[i, :n], "labels": labels[i, :n]}
            out.append({k: v.to(in_dev) for k, v in det.items()} if in_dev != device else det)
        return out


def ssd300_vgg16(pretrained: bool = False, progress: bool = True, num_classes: int = 91, pretrained_backbone: bool = False,
                 trainable_backbone_layers: Optional[int] = None, **kwargs: Any) -> SSD300VGG16B200:
    """SSD300 with a VGG16 backbone (ssd_vgg16.py:139-213) on the B200 kernels.  Defaults as in the reference: score_thresh
    0.01, nms_thresh 0.45, detections_per_img 200, topk_candidates 400 (SSD.__init__), image_mean (0.48235, 0.45882, 0.40784),
    image_std 1/255.  `trainable_backbone_layers` only affects training and is ignored; `pretrained_backbone=True` would
    download amdegroot's VGG16 features (ssd_vgg16.py:24-27) -- load a state_dict instead."""
    if "size" in kwargs:
        warnings.warn("The size of the model is already fixed; ignoring the argument.")
        kwargs.pop("size")
    if pretrained_backbone and not pretrained:
        raise NotImplementedError("pretrained_backbone=True needs a download; load the reference's state_dict instead")
    allowed = {"iou_thresh", "positive_fraction", "score_thresh", "nms_thresh", "detections_per_img", "topk_candidates", "image_mean", "image_std", "act_dtype", "first_conv"}
    unknown = sorted(set(kwargs) - allowed)
    if unknown:
        raise TypeError("SSD.__init__() got an unexpected keyword argument '%s'" % unknown[0])
    model = SSD300VGG16B200(num_classes=num_classes, **kwargs)
    if pretrained:
        state_dict = torch.hub.load_state_dict_from_url(
            "https://download.pytorch.org/models/ssd300_vgg16_coco-b556d3b4.pth", progress=progress)
        model.load_state_dict(state_dict)
    return model
