"""torch.hub entry points, mirroring the reference's hubconf.py:25-44 (`ssd_lite_mobilenet_v2`) and
additionally exposing `ssdlite320_mobilenet_v3_large` (the reference exports it from
demonet.models only, demonet/models/__init__.py:2)."""
from demonet_b200.models import ssd_lite_mobilenet_v2, ssdlite320_mobilenet_v3_large  # noqa: F401
from demonet_b200.vgg import ssd300_vgg16  # noqa: F401

dependencies = ["torch"]
