"""demonet_b200: B200-native (sm_100a) SSDLite inference hot path of zhiqwang/demonet.

Drop-in builders (`ssdlite320_mobilenet_v3_large`, `ssd_lite_mobilenet_v2`) returning an
`nn.Module` with the reference's state_dict keys and output contract; all arithmetic runs in
hand-written CUDA behind the C ABI of include/demonet_b200.h.
"""
from . import ops  # noqa: F401
from . import custom_ops  # noqa: F401  (registers torch.ops.demonet_b200.*)
from . import loss  # noqa: F401  (training side: SSDMatcher, match_targets, compute_loss)
from .models import ssd_lite_mobilenet_v2, ssdlite320_mobilenet_v3_large  # noqa: F401
from .module import SSDLiteB200  # noqa: F401
from .vgg import SSD300VGG16B200, ssd300_vgg16  # noqa: F401

__all__ = ["ssdlite320_mobilenet_v3_large", "ssd_lite_mobilenet_v2", "ssd300_vgg16", "SSDLiteB200", "SSD300VGG16B200", "ops",
           "custom_ops", "loss"]
