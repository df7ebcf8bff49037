"""Model builders with the reference's names and signatures.

  ssdlite320_mobilenet_v3_large  <- demonet/models/ssd_mobilenetv3.py:159-227
  ssd_lite_mobilenet_v2          <- hubconf.py:25-44 (the reference's own builder imports a module
                                    that no longer exists, hubconf.py:4; the assembly follows
                                    test/test_model.py:26-60 and SURVEY.md section 8(c))
"""
import warnings
from typing import Any, Optional

import torch

from . import plan as _plan
from .module import SSDLiteB200

__all__ = ["ssdlite320_mobilenet_v3_large", "ssd_lite_mobilenet_v2"]

model_urls = {
    "ssdlite320_mobilenet_v3_large_coco":
        "https://download.pytorch.org/models/ssdlite320_mobilenet_v3_large_coco-a79551df.pth",
    "mobilenet_v3_large": "https://download.pytorch.org/models/mobilenet_v3_large-8738ca79.pth",
    "ssd_lite_mobilenet_v2": "./checkpoints/mobilenet_v2/ssd_lite_mobilenet_v2_199.pth",
}

# kwargs the reference forwards to SSD.__init__ (generalized_ssd.py:154-163); the training-only ones configure the loss
# branch (SSDMatcher threshold, hard-negative ratio)
_SSD_KWARGS = {"score_thresh", "nms_thresh", "detections_per_img", "topk_candidates", "image_mean", "image_std"}
_SSD_TRAIN_KWARGS = {"iou_thresh", "positive_fraction"}
_ENGINE_KWARGS = {"gemm_impl", "use_cuda_graph", "keep_activations", "pipeline_slots", "act_dtype"}


def _reject_unknown(kwargs):
    """The reference forwards **kwargs both to the backbone builder, which swallows anything (mobilenetv3.py:111,190),
    and to SSD.__init__ (ssd_mobilenetv3.py:217-219), which does not: an unknown key ends in
    `TypeError: SSD.__init__() got an unexpected keyword argument 'bogus'` [probed on the unmodified reference].
    Same exception type and wording here."""
    unknown = sorted(set(kwargs) - _SSD_KWARGS - _SSD_TRAIN_KWARGS - _ENGINE_KWARGS)
    if unknown:
        raise TypeError("SSD.__init__() got an unexpected keyword argument '%s'" % unknown[0])


def ssdlite320_mobilenet_v3_large(pretrained: bool = False, progress: bool = True, num_classes: int = 91,
                                  pretrained_backbone: bool = False,
                                  trainable_backbone_layers: Optional[int] = None,
                                  norm_layer=None, **kwargs: Any) -> SSDLiteB200:
    """SSDlite 320x320 with a MobileNetV3-Large backbone (reduced tail), on the B200 engine.

    Same arguments and defaults as the reference (score_thresh 0.001, nms_thresh 0.55,
    detections_per_img 300, topk_candidates 300, image_mean = image_std = 0.5;
    ssd_mobilenetv3.py:207-217).  `trainable_backbone_layers` only affects training and is ignored.
    """
    if "size" in kwargs:
        warnings.warn("The size of the model is already fixed; ignoring the argument.")
        kwargs.pop("size")
    if norm_layer is not None:
        raise NotImplementedError("demonet_b200 folds BatchNorm2d(eps=0.001) into the convolutions; "
                                  "a custom norm_layer cannot be honoured")
    if kwargs.get("width_mult", 1.0) != 1.0 or kwargs.get("dilated", False):
        raise NotImplementedError("only width_mult=1.0, dilated=False are built")
    for k in ("width_mult", "dilated", "min_depth"):
        kwargs.pop(k, None)
    if pretrained_backbone and not pretrained:
        raise NotImplementedError("pretrained_backbone=True selects the non-reduced tail "
                                  "(ssd_mobilenetv3.py:192-193), which is not built; use pretrained=True")
    _reject_unknown(kwargs)
    defaults = {"score_thresh": 0.001, "nms_thresh": 0.55, "detections_per_img": 300, "topk_candidates": 300,
                "image_mean": [0.5, 0.5, 0.5], "image_std": [0.5, 0.5, 0.5]}
    cfg = {**defaults, **{k: v for k, v in kwargs.items() if k in _SSD_KWARGS | _SSD_TRAIN_KWARGS | _ENGINE_KWARGS}}
    model = SSDLiteB200(_plan.plan_ssdlite320_mobilenet_v3_large(num_classes, 320), postprocess="ssd", **cfg)
    if pretrained:
        state_dict = torch.hub.load_state_dict_from_url(model_urls["ssdlite320_mobilenet_v3_large_coco"],
                                                        progress=progress)
        model.load_state_dict(state_dict)
    return model


def ssd_lite_mobilenet_v2(pretrained: bool = False, image_size: int = 320, score_thresh: float = 0.5,
                          num_classes: int = 21, **kwargs: Any) -> SSDLiteB200:
    """SSDLite with the MobileNetV2 backbone + 4 extra inverted-residual blocks (hubconf.py:25-44).

    Defaults follow the reference's only surviving specification of this model,
    test/test_model.py:26-60: nms_thresh 0.45, detections_per_img 100, ImageNet mean/std, 6 anchors
    per location, and the legacy PostProcess semantics (box_head.py:323-381).  Pass
    postprocess="ssd" (+ topk_candidates) for the SSD.postprocess_detections flavour.
    Priors follow DefaultBoxGenerator([[2,3]]*6, 0.2, 0.95): the original AnchorGenerator has no
    surviving implementation.
    """
    flavour = kwargs.pop("postprocess", "legacy")
    _reject_unknown(kwargs)
    cfg = {"nms_thresh": 0.45, "detections_per_img": 100, "topk_candidates": 400,
           **{k: v for k, v in kwargs.items() if k in _SSD_KWARGS | _SSD_TRAIN_KWARGS | _ENGINE_KWARGS}}
    cfg["score_thresh"] = score_thresh
    model = SSDLiteB200(_plan.plan_ssd_lite_mobilenet_v2(num_classes, image_size), postprocess=flavour, **cfg)
    if pretrained:
        checkpoint = torch.load(model_urls["ssd_lite_mobilenet_v2"], map_location="cpu")
        model.load_state_dict(checkpoint)
    return model
