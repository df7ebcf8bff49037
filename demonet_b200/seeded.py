"""Seeded, well-conditioned re-initialisation and synthetic inputs for parity runs and benchmarks.

The reference's default init (`_normal_init`, std 0.03,
demonet/models/ssd_mobilenetv3.py:57-62) collapses every softmax score to exactly 1/91
(SURVEY.md section 0.4), which makes top-k / sort order implementation-defined, so parity
and benchmark runs use this recipe instead, applied to a
state_dict so that the reference model, the oracle and the engine all load the same
tensors:
    conv weight  ~ N(0, 2/fan_in), fan_in = (Cin/groups)*k*k
    conv bias    ~ N(0, 0.5^2)
    BN gamma     ~ U(0.75, 1.25);  beta ~ N(0, 0.1^2)
    running_mean ~ N(0, 0.1^2);    running_var ~ U(0.75, 1.25)
Draw order = state_dict order, one torch.Generator (CPU), seed 1234.
"""
import math
from collections import OrderedDict

import torch

DEFAULT_SEED = 1234


def seeded_state_dict(template, seed=DEFAULT_SEED):
    """template: an (ordered) state_dict; returns a new OrderedDict of same keys/shapes/dtypes."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    keys = list(template.keys())
    keyset = set(keys)
    out = OrderedDict()
    for k in keys:
        t = template[k]
        shape = tuple(t.shape)
        stem = k.rsplit(".", 1)[0]
        leaf = k.rsplit(".", 1)[1]
        is_bn = (stem + ".running_mean") in keyset
        if leaf == "num_batches_tracked":
            v = torch.zeros(shape, dtype=t.dtype)
        elif is_bn and leaf == "weight":
            v = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif is_bn and leaf == "bias":
            v = torch.randn(shape, generator=g) * 0.1
        elif leaf == "running_mean":
            v = torch.randn(shape, generator=g) * 0.1
        elif leaf == "running_var":
            v = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif leaf == "scale_weight":            # ssd300_vgg16's L2-norm scale (ssd_vgg16.py:40: 20), spread a little
            v = 20.0 * (torch.rand(shape, generator=g) * 0.5 + 0.75)
        elif leaf == "weight" and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            v = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif leaf == "bias":
            v = torch.randn(shape, generator=g) * 0.5
        else:
            raise ValueError("unexpected state_dict entry %s %s" % (k, shape))
        out[k] = v.to(t.dtype)
    return out


def synthetic_images(batch, size, seed=1):
    """torch.rand(B,3,S,S) fp32 in [0,1), seed 1."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.rand(batch, 3, size, size, generator=g)


def seeded_vgg_state_dict(template, seed=DEFAULT_SEED):
    """The recipe above for ssd300_vgg16, with the SSD heads scaled down: the VGG input is normalised to +-128
    (image_std = 1/255, ssd_vgg16.py:199) and He-initialised ReLU layers keep that magnitude, so unscaled heads would give
    logits with a standard deviation in the hundreds and one-hot softmax scores -- a degenerate parity input.  With these
    factors the logits have std ~ 1.6 (no score saturates to exactly 1.0, which would make top-k / sort order a matter of
    tie-breaking) and the box regression std ~ 1.5."""
    sd = seeded_state_dict(template, seed)
    for k in sd:
        if k.startswith("head.classification_head") and k.endswith(".weight"):
            sd[k] = sd[k] * 0.008
        elif k.startswith("head.regression_head") and k.endswith(".weight"):
            sd[k] = sd[k] * 0.008
        elif k.startswith("head.") and k.endswith(".bias"):
            sd[k] = sd[k] * (1.0 if "classification" in k else 0.2)
    return sd
