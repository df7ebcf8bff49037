"""Host-side layer plans for the SSDLite detectors of demonet.

A plan is the static description the C engine executes: a list of layers (stem / depthwise /
pointwise / squeeze-excitation) with the state_dict keys that feed each one, the tensor each one
reads and writes, and the arena buffer assignment.  It mirrors the reference's module structure so
that the reference's state_dict loads unchanged:

  * V3: ssdlite320_mobilenet_v3_large              demonet/models/ssd_mobilenetv3.py:98-227,
        MobileNetV3-Large features (reduced tail)  demonet/models/mobilenetv3.py:61-99,198-215
  * V2: MobileNetWithExtraBlocks + MultiBoxLiteHead demonet/models/backbone.py:45-119,
        demonet/models/box_head.py:24-104, MobileNetV2 features demonet/models/mobilenetv2.py:62-171

BatchNorm folding (done once per weight update, float64 -> fp32):
    s = gamma / sqrt(running_var + eps);  W' = W * s;  b' = beta - running_mean * s (+ conv_bias * s)
Pointwise weights are then rounded to the 16-bit activation type (fp16 by default, bf16 optional: tensor-core operands); depthwise / stem / SE weights
stay fp32.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _C


@dataclass
class Layer:
    kind: str                       # 'stem' | 'dw' | 'pw' | 'se'
    conv: str                       # state_dict prefix of the conv ('.weight' [, '.bias'])
    bn: Optional[str]               # state_dict prefix of the BatchNorm, or None
    cin: int
    cout: int
    k: int = 1
    stride: int = 1
    act: str = "none"
    conv_bias: bool = False
    src: str = ""
    dst: str = ""
    res: Optional[str] = None
    se_mid: int = 0                 # SE squeeze width (kind == 'se'; conv = prefix of fc1/fc2)
    head: Optional[Tuple[str, int]] = None     # ('cls'|'reg', level) for the head 1x1 convs
    h_in: int = 0
    w_in: int = 0
    h_out: int = 0
    w_out: int = 0


@dataclass
class Plan:
    name: str
    size: int
    num_classes: int
    bn_eps: float
    layers: List[Layer] = field(default_factory=list)
    tensors: Dict[str, Tuple[int, int, int]] = field(default_factory=dict)      # name -> (H, W, C)
    feature_names: List[str] = field(default_factory=list)
    anchors_per_loc: int = 6
    param_specs: List[Tuple[str, Tuple[int, ...], str]] = field(default_factory=list)   # (key, shape, role)

    @property
    def grid_sizes(self):
        return [self.tensors[n][:2] for n in self.feature_names]

    @property
    def num_priors(self):
        return sum(h * w * self.anchors_per_loc for h, w in self.grid_sizes)


def _conv_out(n, k, s):
    p = (k - 1) // 2
    return (n + 2 * p - k) // s + 1


class _Builder:
    def __init__(self, plan: Plan):
        self.p = plan
        self.n = 0

    def _new(self, h, w, c, name=None):
        name = name or "t%d" % self.n
        self.n += 1
        self.p.tensors[name] = (h, w, c)
        return name

    def _params_conv(self, conv, shape, bias):
        self.p.param_specs.append((conv + ".weight", shape, "conv_w"))
        if bias:
            self.p.param_specs.append((conv + ".bias", (shape[0],), "conv_b"))

    def _params_bn(self, bn, c):
        for leaf, role in (("weight", "bn_w"), ("bias", "bn_b"), ("running_mean", "bn_m"), ("running_var", "bn_v")):
            self.p.param_specs.append((bn + "." + leaf, (c,), role))
        self.p.param_specs.append((bn + ".num_batches_tracked", (), "bn_n"))

    def stem(self, conv, bn, cout, act):
        s = self.p.size
        ho = _conv_out(s, 3, 2)
        dst = self._new(ho, ho, cout)
        self._params_conv(conv, (cout, 3, 3, 3), False)
        self._params_bn(bn, cout)
        self.p.layers.append(Layer("stem", conv, bn, 3, cout, 3, 2, act, False, "images", dst, None, 0, None, s, s, ho, ho))
        return dst

    def dw(self, src, conv, bn, k, stride, act, conv_bias=False):
        h, w, c = self.p.tensors[src]
        ho, wo = _conv_out(h, k, stride), _conv_out(w, k, stride)
        dst = self._new(ho, wo, c)
        self._params_conv(conv, (c, 1, k, k), conv_bias)
        self._params_bn(bn, c)
        self.p.layers.append(Layer("dw", conv, bn, c, c, k, stride, act, conv_bias, src, dst, None, 0, None, h, w, ho, wo))
        return dst

    def pw(self, src, conv, bn, cout, act, conv_bias=False, res=None, head=None):
        h, w, c = self.p.tensors[src]
        dst = None if head else self._new(h, w, cout)
        self._params_conv(conv, (cout, c, 1, 1), conv_bias)
        if bn:
            self._params_bn(bn, cout)
        self.p.layers.append(Layer("pw", conv, bn, c, cout, 1, 1, act, conv_bias, src, dst or "", res, 0, head, h, w, h, w))
        return dst

    def se(self, src, prefix, mid):
        h, w, c = self.p.tensors[src]
        self.p.param_specs += [(prefix + ".fc1.weight", (mid, c, 1, 1), "conv_w"), (prefix + ".fc1.bias", (mid,), "conv_b"),
                               (prefix + ".fc2.weight", (c, mid, 1, 1), "conv_w"), (prefix + ".fc2.bias", (c,), "conv_b")]
        self.p.layers.append(Layer("se", prefix, None, c, c, 1, 1, "none", True, src, src, None, mid, None, h, w, h, w))
        return src


def _make_divisible(v, divisor=8):
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


# (kernel, expanded, out, use_se, activation, stride): mobilenetv3.py:198-215, reduce_divider = 2
_V3_LARGE = [(3, 16, 16, False, "relu", 1), (3, 64, 24, False, "relu", 2), (3, 72, 24, False, "relu", 1),
             (5, 72, 40, True, "relu", 2), (5, 120, 40, True, "relu", 1), (5, 120, 40, True, "relu", 1),
             (3, 240, 80, False, "hardswish", 2), (3, 200, 80, False, "hardswish", 1),
             (3, 184, 80, False, "hardswish", 1), (3, 184, 80, False, "hardswish", 1),
             (3, 480, 112, True, "hardswish", 1), (3, 672, 112, True, "hardswish", 1),
             (5, 672, 80, True, "hardswish", 2), (5, 480, 80, True, "hardswish", 1), (5, 480, 80, True, "hardswish", 1)]


def plan_ssdlite320_mobilenet_v3_large(num_classes=91, size=320) -> Plan:
    p = Plan("ssdlite320_mobilenet_v3_large", size, num_classes, 1e-3)
    b = _Builder(p)
    f0 = "backbone.features.0."
    x = b.stem(f0 + "0.0", f0 + "0.1", 16, "hardswish")
    cin = 16
    feats = []
    for i, (k, cexp, cout, use_se, act, stride) in enumerate(_V3_LARGE):
        inp = x
        if i < 12:
            blk = f0 + "%d.block." % (i + 1)
        elif i == 12:
            blk = None                      # C4 is split after its expansion (ssd_mobilenetv3.py:104-108)
        else:
            blk = "backbone.features.1.%d.block." % (i - 12)
        j = 0
        if cexp != cin:
            pre = (f0 + "13") if i == 12 else (blk + str(j))
            x = b.pw(x, pre + ".0", pre + ".1", cexp, act)
            j += 1
            if i == 12:
                feats.append(x)
        if i == 12:
            blk = "backbone.features.1.0."
        pre = blk + str(j)
        x = b.dw(x, pre + ".0", pre + ".1", k, stride, act)
        j += 1
        if use_se:
            x = b.se(x, blk + str(j), _make_divisible(cexp // 4))
            j += 1
        pre = blk + str(j)
        x = b.pw(x, pre + ".0", pre + ".1", cout, "none", res=inp if (stride == 1 and cin == cout) else None)
        cin = cout
    x = b.pw(x, "backbone.features.1.3.0", "backbone.features.1.3.1", 6 * cin, "hardswish")
    feats.append(x)
    for e, cout in enumerate((512, 256, 256, 128)):       # _extra_block, ssd_mobilenetv3.py:39-54
        q = "backbone.extra.%d." % e
        x = b.pw(x, q + "0.0", q + "0.1", cout // 2, "relu6")
        x = b.dw(x, q + "1.0", q + "1.1", 3, 2, "relu6")
        x = b.pw(x, q + "2.0", q + "2.1", cout, "relu6")
        feats.append(x)
    p.feature_names = feats
    for name, cols, kind in (("classification_head", num_classes, "cls"), ("regression_head", 4, "reg")):
        for l, f in enumerate(feats):                     # _prediction_block, ssd_mobilenetv3.py:27-36
            q = "head.%s.module_list.%d." % (name, l)
            h = b.dw(f, q + "0.0", q + "0.1", 3, 1, "relu6")
            b.pw(h, q + "1", None, p.anchors_per_loc * cols, "none", conv_bias=True, head=(kind, l))
    return p


_V2_SETTING = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
_V2_EXTRAS = [(1280, 512, 0.2), (512, 256, 0.25), (256, 256, 0.5), (256, 64, 0.25)]     # backbone.py:54-58


def plan_ssd_lite_mobilenet_v2(num_classes=21, size=320) -> Plan:
    p = Plan("ssd_lite_mobilenet_v2", size, num_classes, 1e-5)
    b = _Builder(p)

    def inverted_residual(x, prefix, inp, oup, stride, hidden, expand):
        src = x
        j = 0
        if expand:
            x = b.pw(x, prefix + ".conv.%d.0" % j, prefix + ".conv.%d.1" % j, hidden, "relu6")
            j += 1
        x = b.dw(x, prefix + ".conv.%d.0" % j, prefix + ".conv.%d.1" % j, 3, stride, "relu6")
        j += 1
        return b.pw(x, prefix + ".conv.%d" % j, prefix + ".conv.%d" % (j + 1), oup, "none",
                    res=src if (stride == 1 and inp == oup) else None)

    x = b.stem("backbone.body.0.0", "backbone.body.0.1", 32, "relu6")
    cin, idx, feats = 32, 1, []
    for t, c, n, s in _V2_SETTING:
        for i in range(n):
            x = inverted_residual(x, "backbone.body.%d" % idx, cin, c, s if i == 0 else 1, int(round(cin * t)), t != 1)
            cin = c
            if idx == 13:
                feats.append(x)            # tap "13", backbone.py:52
            idx += 1
    x = b.pw(x, "backbone.body.18.0", "backbone.body.18.1", 1280, "relu6")
    feats.append(x)
    for e, (inp, oup, t) in enumerate(_V2_EXTRAS):
        x = inverted_residual(x, "backbone.extra_blocks.%d" % e, inp, oup, 2, int(round(inp * t)), True)
        feats.append(x)
    p.feature_names = feats
    for name, cols, kind in (("cls_logits", num_classes, "cls"), ("bbox_pred", 4, "reg")):
        for l, f in enumerate(feats):
            q = "head.%s.%d" % (name, l)
            if l < len(feats) - 1:         # SeperableConv2d, box_head.py:24-34 (biased depthwise)
                h = b.dw(f, q + ".0", q + ".1", 3, 1, "relu6", conv_bias=True)
                b.pw(h, q + ".3", None, p.anchors_per_loc * cols, "none", conv_bias=True, head=(kind, l))
            else:                          # plain 1x1, box_head.py:55-56
                b.pw(f, q, None, p.anchors_per_loc * cols, "none", conv_bias=True, head=(kind, l))
    # parameter order of the reference: backbone, then per-level interleaved cls/bbox is not needed --
    # state_dict loading is by key
    return p


# ---------------------------------------------------------------------------------------------
# weights: fold BN, lay out, pack into one blob
# ---------------------------------------------------------------------------------------------
def _align(n, a=256):
    return (n + a - 1) // a * a


def fold_layer(sd, layer: Layer, eps: float):
    """Returns (W' float64 [Cout, ...], b' float64 [Cout]) for a conv(+BN) layer."""
    w = sd[layer.conv + ".weight"].detach().double().cpu()
    cb = sd[layer.conv + ".bias"].detach().double().cpu() if layer.conv_bias else None
    if layer.bn is None:
        return w, (cb if cb is not None else torch.zeros(w.shape[0], dtype=torch.float64))
    g = sd[layer.bn + ".weight"].detach().double().cpu()
    beta = sd[layer.bn + ".bias"].detach().double().cpu()
    mean = sd[layer.bn + ".running_mean"].detach().double().cpu()
    var = sd[layer.bn + ".running_var"].detach().double().cpu()
    s = g / torch.sqrt(var + eps)
    wf = w * s.view(-1, 1, 1, 1)
    bf = beta - mean * s
    if cb is not None:
        bf = bf + cb * s
    return wf, bf


def pw_pack_factor(L: Layer) -> int:
    """Pixel packing for pointwise layers with very few input channels.

    A [M, K] NHWC activation with K <= 32 is also a [M/p, p*K] matrix (p consecutive pixels per row), and
    the [M, N] output is a [M/p, p*N] matrix, so the 1x1 conv equals one GEMM with the block-diagonal
    weight diag(W, ..., W).  The tensor cores do p times the MACs (irrelevant: these layers are
    bandwidth-bound) but every TMA row becomes a dense 128-byte line and there are p times fewer tiles.
    """
    if L.kind != "pw" or L.head or L.cin > 96:
        return 1
    hw = L.h_in * L.w_in

    def k_waste(k):            # padded K (64-wide k-blocks) per useful K
        return ((k + 63) // 64 * 64) / k
    for p in (4, 2):
        if p * L.cout <= (256 if p == 4 else 512) and hw % p == 0 and p * L.cin <= 192 and k_waste(p * L.cin) <= k_waste(L.cin):
            return p
    return 1


def pack_weights(plan: Plan, sd, act_dtype: str = None) -> Tuple[bytes, List[Dict[str, int]]]:
    """Fold + lay out every layer.  Returns (blob, per-layer offsets dict(w, b[, w2, b2])).  `act_dtype` ("fp16" /
    "bf16"): the 16-bit type of the tensor-core operands = the activation storage type of the library that will run."""
    chunks, offs, pos = [], [], 0
    h16 = _C.torch_dtype(act_dtype)

    def put(arr: np.ndarray):
        nonlocal pos
        raw = arr.tobytes()
        start = pos
        chunks.append(raw)
        pad = _align(len(raw)) - len(raw)
        if pad:
            chunks.append(b"\0" * pad)
        pos += len(raw) + pad
        return start

    for L in plan.layers:
        if L.kind == "se":
            w1 = sd[L.conv + ".fc1.weight"].detach().float().cpu().reshape(L.se_mid, L.cin)
            b1 = sd[L.conv + ".fc1.bias"].detach().float().cpu()
            w2 = sd[L.conv + ".fc2.weight"].detach().float().cpu().reshape(L.cin, L.se_mid)
            b2 = sd[L.conv + ".fc2.bias"].detach().float().cpu()
            offs.append({"w": put(w1.contiguous().numpy()), "b": put(b1.numpy()),
                         "w2": put(w2.t().contiguous().numpy()), "b2": put(b2.numpy())})
            continue
        wf, bf = fold_layer(sd, L, plan.bn_eps)
        bias = bf.float().numpy()
        if L.kind == "stem":      # [Cout,3,3,3] -> [(ci*3+kh)*3+kw][Cout] fp32
            w = wf.float().permute(1, 2, 3, 0).reshape(27, L.cout).contiguous().numpy()
        elif L.kind == "dw":      # [C,1,k,k] -> [k*k][C] fp32
            w = wf.float().reshape(L.cout, L.k * L.k).t().contiguous().numpy()
        else:                     # pw: [N,K,1,1] -> [N][K] fp16 / bf16 (block-diagonal when pixels are packed)
            w2 = wf.float().reshape(L.cout, L.cin).to(h16)
            pf = pw_pack_factor(L)
            raw_w, raw_b = w2.contiguous().view(torch.int16).numpy(), bias
            if pf > 1:
                w2 = torch.block_diag(*([w2.float()] * pf)).to(h16)
                bias = np.tile(bias, pf)
            w = w2.contiguous().view(torch.int16).numpy()
            entry = {"w": put(w), "b": put(bias)}
            # the fused expand + depthwise kernel takes the plain [N, K] weights (no pixel packing)
            entry["w_raw"], entry["b_raw"] = (put(raw_w), put(raw_b)) if pf > 1 else (entry["w"], entry["b"])
            offs.append(entry)
            continue
        offs.append({"w": put(w), "b": put(bias)})
    return b"".join(chunks), offs


# ---------------------------------------------------------------------------------------------
# arena buffers (liveness-based reuse) and the dn_op array
# ---------------------------------------------------------------------------------------------
def layer_lanes(plan: Plan) -> List[int]:
    """Launch lane of every layer: 0 = the backbone / extras chain, 1 + 2*level (+1) = the classification
    (regression) head of one feature level.  The 12 head branches (depthwise + 1x1 each) depend only on their
    feature map, so the engine runs them on side streams that fork off the main chain -- inside the CUDA graph
    they become parallel branches that overlap the tail of the backbone and each other."""
    lanes = [0] * len(plan.layers)
    by_dst = {L.dst: i for i, L in enumerate(plan.layers) if L.dst and L.kind != "se"}
    for i, L in enumerate(plan.layers):
        if L.head:
            kind, lvl = L.head
            lanes[i] = 1 + 2 * lvl + (0 if kind == "cls" else 1)
            j = by_dst.get(L.src)
            if j is not None and plan.layers[j].kind == "dw" and L.src not in plan.feature_names:
                lanes[j] = lanes[i]
    return lanes


def assign_buffers(plan: Plan, reuse: bool = True):
    """Returns (tensor -> buffer id, [(elems_per_image, elem_bytes)], logits_buf, bbox_buf).
    reuse=False gives every tensor its own buffer (used by the per-layer parity tests).
    Tensors written or read on a side lane (layer_lanes) keep a buffer of their own for the whole forward:
    list order is not execution order across lanes, so they can neither take over nor hand on a buffer."""
    lanes = layer_lanes(plan)
    pinned = set()
    last_use = {}
    for i, L in enumerate(plan.layers):
        for t in (L.src, L.res):
            if t and t != "images":
                last_use[t] = i
                if lanes[i]:
                    pinned.add(t)
        if lanes[i] and L.dst:
            pinned.add(L.dst)
    # A fused launch (pw+dw or dw+pw) reads layers[i].src while it writes layers[i+1].dst, and its CTAs read halo pixels
    # that other CTAs' output tiles cover: the input must stay live until the SECOND layer of the pair, or best-fit could
    # hand its buffer to the fused output and the stencil would run in place.
    for i in list(fused_pairs(plan)) + list(fused_dwpw_pairs(plan)):
        src = plan.layers[i].src
        if src and src != "images":
            last_use[src] = max(last_use.get(src, i), i + 1)
    bufs: List[List[int]] = []      # [elems, bytes]
    free: List[int] = []
    t2b: Dict[str, int] = {}
    for i, L in enumerate(plan.layers):
        if L.kind != "se" and L.dst:
            h, w, c = plan.tensors[L.dst]
            need = h * w * c
            best = None
            avail = [] if L.dst in pinned else free
            for bid in avail:
                if bufs[bid][0] >= need and (best is None or bufs[bid][0] < bufs[best][0]):
                    best = bid
            if best is None and avail:         # grow the largest free buffer instead of adding one
                best = max(avail, key=lambda q: bufs[q][0])
                bufs[best][0] = need
            if best is None:
                bufs.append([need, 2])
                best = len(bufs) - 1
            else:
                free.remove(best)
            t2b[L.dst] = best
        for t in {L.src, L.res}:
            if reuse and t and t != "images" and last_use.get(t) == i and t in t2b and t not in pinned:
                free.append(t2b[t])
    P, K = plan.num_priors, plan.num_classes
    bufs.append([P * K, 4])
    logits_buf = len(bufs) - 1
    bufs.append([P * 4, 4])
    bbox_buf = len(bufs) - 1
    return t2b, [tuple(b) for b in bufs], logits_buf, bbox_buf


def fused_pairs(plan: Plan) -> List[int]:
    """Indices i such that layers[i] (pointwise expand) and layers[i+1] (depthwise) run as ONE kernel with the expanded
    tensor kept on chip (csrc/pwdw_fused.cu).  Conditions: the expand output feeds only that depthwise layer, no
    residual / head, and the shape is one the kernel is built for (16 -> 64 channels, 3x3 stride 2: MobileNetV3
    block 2, the largest expanded tensor of the network)."""
    uses: Dict[str, int] = {}
    for L in plan.layers:
        for t in (L.src, L.res):
            if t:
                uses[t] = uses.get(t, 0) + 1
    out = []
    for i in range(len(plan.layers) - 1):
        a, b = plan.layers[i], plan.layers[i + 1]
        if a.kind != "pw" or b.kind != "dw" or a.head or a.res or b.src != a.dst or uses.get(a.dst, 0) != 1:
            continue
        if a.dst in plan.feature_names or a.conv_bias or b.conv_bias:
            continue
        if (a.cin, a.cout, b.k, b.stride) == (16, 64, 3, 2) and a.h_in >= 8 and a.w_in >= 8 and a.act == b.act and a.act != "none":
            out.append(i)
    return out


def dwpw_fusion_enabled() -> bool:
    """DN_FUSE_DWPW=0 keeps the depthwise + project pair of block 1 as two launches (measurement aid; r01: fused 0.199 ms,
    unfused 0.104 + 0.137 ms)."""
    import os
    return os.environ.get("DN_FUSE_DWPW", "1") != "0"


def fused_dwpw_pairs(plan: Plan) -> List[int]:
    """Indices i such that layers[i] (depthwise) and layers[i+1] (pointwise project, optionally `+= block input`) run as
    ONE kernel with the depthwise output kept on chip (csrc/dwpw_fused.cu): 16 channels, 3x3 stride 1, project 16 -> 16 --
    MobileNetV3 block 1.  The residual, if any, must be the depthwise layer's own input."""
    uses: Dict[str, int] = {}
    for L in plan.layers:
        for t in (L.src, L.res):
            if t:
                uses[t] = uses.get(t, 0) + 1
    taken = set(fused_pairs(plan))
    out = []
    for i in range(len(plan.layers) - 1):
        a, b = plan.layers[i], plan.layers[i + 1]
        if a.kind != "dw" or b.kind != "pw" or b.head or b.src != a.dst or uses.get(a.dst, 0) != 1 or a.conv_bias or b.conv_bias:
            continue
        if i in taken or i - 1 in taken or a.dst in plan.feature_names:
            continue
        if b.res is not None and b.res != a.src:
            continue
        if (a.cin, b.cout, a.k, a.stride) == (16, 16, 3, 1) and a.h_in >= 8 and a.w_in >= 16 and a.act != "none" and b.act == "none":
            out.append(i)
    return out


def se_fold_enabled() -> bool:
    """DN_SE_FOLD=0 keeps the in-place scaling pass of every squeeze-excitation (measurement aid)."""
    import os
    return os.environ.get("DN_SE_FOLD", "1") != "0"


def se_folded_layers(plan: Plan) -> List[int]:
    """Indices i of squeeze-excitation layers whose scaling pass is folded into the project GEMM layers[i + 1]: the SE
    output feeds only that 1x1 convolution (InvertedResidual, mobilenetv3.py:85-89), which has no activation."""
    uses: Dict[str, int] = {}
    for L in plan.layers:
        for t in (L.src, L.res):
            if t:
                uses[t] = uses.get(t, 0) + 1
    out = []
    for i in range(len(plan.layers) - 1):
        a, b = plan.layers[i], plan.layers[i + 1]
        # uses: the depthwise layer wrote it, the SE layer (src == dst) and the project GEMM read it
        if a.kind == "se" and b.kind == "pw" and b.src == a.src and b.act == "none" and not b.head and uses.get(a.src, 0) == 2 \
                and a.src not in plan.feature_names and b.res != a.src:
            out.append(i)
    return out


def build_ops(plan: Plan, offsets, t2b, logits_buf, bbox_buf, fuse: bool = True):
    """ctypes dn_op array for the engine (fuse=False keeps every layer a launch of its own)."""
    level_off, o = [], 0
    for h, w in plan.grid_sizes:
        level_off.append(o)
        o += h * w * plan.anchors_per_loc
    P, K = plan.num_priors, plan.num_classes
    ops = (_C.Op * len(plan.layers))()
    lanes = layer_lanes(plan)
    kinds = {"stem": _C.OP_STEM, "dw": _C.OP_DW, "pw": _C.OP_PW, "se": _C.OP_SE}
    fused = set(fused_pairs(plan)) if fuse else set()
    fused_dp = set(fused_dwpw_pairs(plan)) if fuse and dwpw_fusion_enabled() else set()
    folded = set(se_folded_layers(plan)) if fuse and se_fold_enabled() else set()
    for i, (L, off) in enumerate(zip(plan.layers, offsets)):
        op = ops[i]
        op.kind, op.act = kinds[L.kind], _C.ACT[L.act]
        op.se_fold = 1 if (i in folded or i - 1 in folded) else 0
        if i in fused_dp:               # depthwise + project (+ residual) in one launch
            Pj = plan.layers[i + 1]
            op.kind = _C.OP_DWPW
            op.in_buf, op.out_buf = t2b[L.src], t2b[Pj.dst]
            op.res_buf = t2b[Pj.res] if Pj.res else _C.BUF_NONE
            op.h_in, op.w_in, op.c_in = L.h_in, L.w_in, L.cin
            op.h_out, op.w_out, op.c_out = Pj.h_out, Pj.w_out, Pj.cout
            op.ksize, op.stride, op.lane = L.k, L.stride, lanes[i]
            op.w_off, op.b_off = off["w"], off["b"]
            op.w2_off, op.b2_off = offsets[i + 1]["w_raw"], offsets[i + 1]["b_raw"]
            assert op.in_buf != op.out_buf, "fused depthwise + project would run in place"
            continue
        if i - 1 in fused_dp:
            op.kind = _C.OP_NOP
            op.in_buf = op.out_buf = t2b[L.dst]
            op.res_buf = _C.BUF_NONE
            op.lane = lanes[i]
            continue
        if i in fused:                  # expand + depthwise in one launch; the depthwise slot becomes a no-op
            D = plan.layers[i + 1]
            op.kind, op.act2 = _C.OP_PWDW, _C.ACT[D.act]
            op.in_buf, op.out_buf, op.res_buf = t2b[L.src], t2b[D.dst], _C.BUF_NONE
            op.h_in, op.w_in, op.c_in = L.h_in, L.w_in, L.cin
            op.h_out, op.w_out, op.c_out = D.h_out, D.w_out, D.cout
            op.ksize, op.stride, op.lane = D.k, D.stride, lanes[i]
            op.w_off, op.b_off = off["w_raw"], off["b_raw"]
            op.w2_off, op.b2_off = offsets[i + 1]["w"], offsets[i + 1]["b"]
            assert op.in_buf != op.out_buf, "fused expand + depthwise would run in place"
            continue
        if i - 1 in fused:
            op.kind = _C.OP_NOP
            op.in_buf = op.out_buf = t2b[L.dst]
            op.res_buf = _C.BUF_NONE
            op.lane = lanes[i]
            continue
        op.in_buf = _C.BUF_IMAGES if L.src == "images" else t2b[L.src]
        op.res_buf = t2b[L.res] if L.res else _C.BUF_NONE
        op.h_in, op.w_in, op.c_in = L.h_in, L.w_in, L.cin
        op.h_out, op.w_out, op.c_out = L.h_out, L.w_out, L.cout
        pf = pw_pack_factor(L)
        if pf > 1:                # [M, K] x [K, N]  ==  [M/p, pK] x diag(W..W)  (same memory)
            op.h_in, op.w_in, op.c_in = (L.h_in * L.w_in) // pf, 1, pf * L.cin
            op.h_out, op.w_out, op.c_out = op.h_in, 1, pf * L.cout
        op.ksize, op.stride, op.c_mid = L.k, L.stride, L.se_mid
        op.lane = lanes[i]
        op.w_off, op.b_off = off["w"], off["b"]
        op.w2_off, op.b2_off = off.get("w2", 0), off.get("b2", 0)
        if L.head:
            kind, lvl = L.head
            cols = K if kind == "cls" else 4
            op.out_buf = logits_buf if kind == "cls" else bbox_buf
            op.out_fp32 = 1
            op.out_batch_stride = P * cols
            op.out_row_stride = plan.anchors_per_loc * cols
            op.out_offset = level_off[lvl] * cols
        else:
            op.out_buf = t2b[L.dst] if L.kind != "se" else t2b[L.src]
            op.out_fp32 = 0
    return ops


def default_boxes(plan: Plan, aspect_ratios=None, min_ratio=0.2, max_ratio=0.95, clip=True, scales=None, steps=None) -> np.ndarray:
    """The SSD default-box table, f32 [P,4] xyxy pixels -- DefaultBoxGenerator
    (demonet/models/anchor_utils.py:39-126) evaluated once per (model, size) instead of per forward.
    Same fp32 operation order as the reference so the table is bit-identical."""
    return default_boxes_for(plan.grid_sizes, plan.size, aspect_ratios, min_ratio, max_ratio, clip, scales, steps)


def default_boxes_for(grids, size, aspect_ratios=None, min_ratio=0.2, max_ratio=0.95, clip=True, scales=None, steps=None) -> np.ndarray:
    """default_boxes for explicit grid sizes [(H_k, W_k)] and a square image of `size` pixels.  `scales` / `steps`: the explicit
    form ssd300_vgg16 uses (ssd_vgg16.py:193-195; anchor_utils.py:29-50, 79-83)."""
    n = len(grids)
    if aspect_ratios is None:
        aspect_ratios = [[2, 3]] * n
    if scales is None:
        if n > 1:
            scales = [min_ratio + (max_ratio - min_ratio) * k / (n - 1.0) for k in range(n)] + [1.0]
        else:
            scales = [min_ratio, max_ratio]
    f32 = np.float32
    rows = []
    for k, (fh, fw) in enumerate(grids):
        s_k, s_p = scales[k], math.sqrt(scales[k] * scales[k + 1])
        wh = [[s_k, s_k], [s_p, s_p]]
        for ar in aspect_ratios[k]:
            r = math.sqrt(ar)
            wh += [[s_k * r, s_k / r], [s_k / r, s_k * r]]
        wh = np.asarray(wh, dtype=f32)
        if clip:
            wh = np.clip(wh, f32(0), f32(1))
        x_f, y_f = (size / steps[k], size / steps[k]) if steps is not None else (fw, fh)
        cx = (np.arange(fw).astype(f32) + f32(0.5)) / f32(x_f)
        cy = (np.arange(fh).astype(f32) + f32(0.5)) / f32(y_f)
        gy, gx = np.meshgrid(cy, cx, indexing="ij")
        ctr = np.repeat(np.stack([gx.reshape(-1), gy.reshape(-1)], -1), wh.shape[0], axis=0)
        rows.append(np.concatenate([ctr, np.tile(wh, (fh * fw, 1))], axis=1).astype(f32))
    t = np.concatenate(rows, 0)
    out = np.concatenate([t[:, :2] - f32(0.5) * t[:, 2:], t[:, :2] + f32(0.5) * t[:, 2:]], -1).astype(f32)
    out[:, 0::2] *= f32(size)
    out[:, 1::2] *= f32(size)
    return np.ascontiguousarray(out)
