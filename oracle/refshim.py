"""Import the UNMODIFIED reference (zhiqwang/demonet at /root/reference) in-process.

TEST INFRASTRUCTURE ONLY. Used by tests/golden/make_golden.py (run in the build
container, where /root/reference exists) to pin the oracle restatements and to
generate the committed golden vectors. Nothing on the product path, in the
`-m gpu` tests, in smoke() or in bench.py imports this module: /root/reference
does not exist on the GPU box.

Two in-process shims, zero file edits:
  1. `torchvision.models.utils` was removed from modern torchvision; the
     reference imports `load_state_dict_from_url` from it
     (demonet/models/ssd_mobilenetv3.py:7, mobilenetv2.py:3, mobilenetv3.py:6).
  2. `demonet/__init__.py:3-5` pulls pycocotools through demonet/data/coco.py;
     pre-registering a bare `demonet` package skips that file.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DEMONET_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "demonet", "models"))


def install():
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch
    if "torchvision.models.utils" not in sys.modules:
        m = types.ModuleType("torchvision.models.utils")
        m.load_state_dict_from_url = torch.hub.load_state_dict_from_url
        sys.modules["torchvision.models.utils"] = m
    if "demonet" not in sys.modules:
        pkg = types.ModuleType("demonet")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "demonet")]
        sys.modules["demonet"] = pkg
    if "demonet.models" not in sys.modules:
        # demonet/models/__init__.py imports ssd_vgg16 etc.; all import fine under shim 1
        sub = types.ModuleType("demonet.models")
        sub.__path__ = [os.path.join(REFERENCE_ROOT, "demonet", "models")]
        sys.modules["demonet.models"] = sub


def ref_module(name: str):
    """e.g. ref_module('ssd_mobilenetv3') -> demonet.models.ssd_mobilenetv3"""
    install()
    return importlib.import_module("demonet.models." + name)
