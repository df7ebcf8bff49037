"""CPU restatement (PyTorch functional, fp32) of demonet's SSDLite forward, from a state_dict.

TEST INFRASTRUCTURE ONLY -- the parity oracle for the floating-point part of the hot
path and the `--impl reference` / cpu_baseline "port" timed by bench.py.  The product
(demonet_b200/) never imports it.

Parity status: PINNED for V3 against the unmodified reference imported from
/root/reference (tests/golden/make_golden.py asserts bit-equality of logits, box
regression, anchors and final detections in `fp32` mode, and commits the fixtures that
tests/test_oracle_cpu.py re-checks).  The V2 assembly (the reference's hubconf entry does
not import) is pinned against the reference's working parts
(backbone.MobileNetWithExtraBlocks + box_head.MultiBoxLiteHead) the same way.

Modes
  fp32 : the reference's own op sequence (conv2d -> batch_norm(eval) -> activation), i.e.
         what `SSD.forward` computes on CPU.
  bf16 : same graph with the numerics contract of the CUDA engine (DESIGN.md "Numerics"):
         BatchNorm folded into the conv in float64 then rounded to fp32, pointwise-GEMM
         weights rounded to bf16, every stored activation rounded to bf16, fp32 accumulate,
         head 1x1 outputs / softmax / decode kept in fp32.
  fp16 : the bf16 contract with IEEE half instead of bfloat16 (stored activations and pointwise-GEMM weights
         rounded to fp16: same bytes, 3 more mantissa bits, range 65504).
  wf16 : like w16 with fp16 GEMM weights.
  w16  : like bf16 but WITHOUT rounding the activations (folded BN + bf16 GEMM weights only).  The
         bf16 chain is chaotic on a random-weight network (a different summation order alone moves
         the logits by rms 0.14, see DESIGN.md "Numerics"), so plan-wiring tests compare in this
         mode, where only fp32 summation-order noise remains.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

# (kernel, expanded, out, use_se, activation, stride)  -- mobilenetv3.py:198-215 with
# reduced_tail=True (reduce_divider=2), as selected by ssd_mobilenetv3.py:192-199.
V3_LARGE_BNECK = [
    (3, 16, 16, False, "RE", 1),
    (3, 64, 24, False, "RE", 2),
    (3, 72, 24, False, "RE", 1),
    (5, 72, 40, True, "RE", 2),
    (5, 120, 40, True, "RE", 1),
    (5, 120, 40, True, "RE", 1),
    (3, 240, 80, False, "HS", 2),
    (3, 200, 80, False, "HS", 1),
    (3, 184, 80, False, "HS", 1),
    (3, 184, 80, False, "HS", 1),
    (3, 480, 112, True, "HS", 1),
    (3, 672, 112, True, "HS", 1),
    (5, 672, 80, True, "HS", 2),      # C4
    (5, 480, 80, True, "HS", 1),
    (5, 480, 80, True, "HS", 1),
]
V3_BN_EPS = 1e-3      # ssd_mobilenetv3.py:195-196
V2_BN_EPS = 1e-5      # nn.BatchNorm2d default (mobilenetv2.py:45, backbone.py:98, box_head.py:27)


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _fp16(x):
    return x.to(torch.float16).to(torch.float32)


def _ident(x):
    return x


def _act(x, act):
    if act == "RE":
        return F.relu(x)
    if act == "R6":
        return F.relu6(x)
    if act == "HS":
        return F.hardswish(x)
    if act == "ID":
        return x
    raise ValueError(act)


def fold_bn(w, gamma, beta, mean, var, eps, conv_bias=None):
    """W' = W*s, b' = beta - mean*s (+ conv_bias*s), s = g/sqrt(v+eps); float64 math, one rounding to fp32."""
    s = gamma.double() / torch.sqrt(var.double() + eps)
    wf = (w.double() * s.view(-1, 1, 1, 1)).float()
    bf = beta.double() - mean.double() * s
    if conv_bias is not None:
        bf = bf + conv_bias.double() * s
    return wf, bf.float()


class _Net:
    def __init__(self, sd, mode, eps):
        assert mode in ("fp32", "bf16", "w16", "fp16", "wf16")
        self.sd = sd
        self.mode = mode
        self.eps = eps
        # rounding of stored activations / of the tensor-core (pointwise GEMM) weights
        self.rnd = {"bf16": _bf16, "fp16": _fp16}.get(mode, _ident)
        self.wrnd = _fp16 if mode in ("fp16", "wf16") else _bf16

    # ConvBNActivation: conv (no bias) -> BN -> act      mobilenetv2.py:32-55
    def cba(self, x, prefix, stride, act, depthwise=False, conv_key=".0", bn_key=".1", conv_bias=False,
            residual=None):
        sd = self.sd
        w = sd[prefix + conv_key + ".weight"]
        k = w.shape[-1]
        pad = (k - 1) // 2
        groups = w.shape[0] if depthwise else 1
        cb = sd[prefix + conv_key + ".bias"] if conv_bias else None
        g, b = sd[prefix + bn_key + ".weight"], sd[prefix + bn_key + ".bias"]
        m, v = sd[prefix + bn_key + ".running_mean"], sd[prefix + bn_key + ".running_var"]
        if self.mode == "fp32":
            y = F.conv2d(x, w, cb, stride, pad, 1, groups)
            y = _act(F.batch_norm(y, m, v, g, b, False, 0.0, self.eps), act)
            return y if residual is None else y + residual          # mobilenetv3.py:95-99
        wf, bf = fold_bn(w, g, b, m, v, self.eps, cb)
        if k == 1 and not depthwise:
            wf = self.wrnd(wf)
        y = _act(F.conv2d(x, wf, bf, stride, pad, 1, groups), act)
        if residual is not None:        # the engine adds the residual in the GEMM epilogue: one rounding
            y = y + residual
        return self.rnd(y)

    # plain conv with bias, no BN (head 1x1): fp32 output in both modes
    def conv_bias(self, x, prefix):
        w, b = self.sd[prefix + ".weight"], self.sd[prefix + ".bias"]
        if self.mode != "fp32":
            w = self.wrnd(w)
        return F.conv2d(x, w, b)

    # SqueezeExcitation                                   mobilenetv3.py:22-40
    def se(self, x, prefix):
        sd = self.sd
        s = F.adaptive_avg_pool2d(x, 1)
        s = F.relu(F.conv2d(s, sd[prefix + ".fc1.weight"], sd[prefix + ".fc1.bias"]))
        s = F.hardsigmoid(F.conv2d(s, sd[prefix + ".fc2.weight"], sd[prefix + ".fc2.bias"]))
        return self.rnd(s * x)


def _score_layout(results, num_columns):
    """SSDScoringHead.forward, generalized_ssd.py:66-72: (N, A*K, H, W) -> (N, HWA, K)."""
    N, _, H, W = results.shape
    r = results.view(N, -1, num_columns, H, W).permute(0, 3, 4, 1, 2)
    return r.reshape(N, -1, num_columns)


def _tick(times, key, t0):
    """Accumulate wall time since t0 under times[key] (per-stage CPU timing of the baseline run); returns now."""
    import time
    t1 = time.perf_counter()
    if times is not None:
        times[key] = times.get(key, 0.0) + (t1 - t0)
    return t1


def v3_forward_raw(sd, images, mode="fp32", num_classes=91, image_mean=(0.5, 0.5, 0.5),
                   image_std=(0.5, 0.5, 0.5), return_features=False, times=None):
    """ssdlite320_mobilenet_v3_large up to the head outputs.

    images: f32[B,3,S,S] in [0,1] already at the model size (the transform's resize is then an
    identity, transform.py:150-160).  Returns (cls_logits f32[B,P,K], bbox_regression f32[B,P,4],
    [feature H,W per level]).  Follows generalized_ssd.py:296-313, ssd_mobilenetv3.py:121-132.
    """
    import time
    t0 = time.perf_counter()
    net = _Net(sd, mode, V3_BN_EPS)
    mean = torch.as_tensor(image_mean, dtype=torch.float32)[None, :, None, None]
    std = torch.as_tensor(image_std, dtype=torch.float32)[None, :, None, None]
    x = (images - mean) / std                                  # transform.py:129-138
    t0 = _tick(times, "transform", t0)
    p = "backbone.features.0."
    x = net.cba(x, p + "0", 2, "HS")                           # stem, mobilenetv3.py:141-142
    feats = []
    cin = 16
    for i, (k, cexp, cout, use_se, act, stride) in enumerate(V3_LARGE_BNECK):
        # InvertedResidual, mobilenetv3.py:61-99.  Block 12 (C4) is split after its expansion
        # (ssd_mobilenetv3.py:104-108): features.0.13 = expand, features.1.0 = the rest.
        if i < 12:
            bp = p + "%d.block." % (i + 1)
            j = 0
        elif i == 12:
            bp = None
        else:
            bp = "backbone.features.1.%d.block." % (i - 12)
            j = 0
        inp = x
        if i == 12:
            x = net.cba(x, p + "13", 1, act)                   # C4 expansion = feature level 0
            feats.append(x)
            q = "backbone.features.1.0."
            x = net.cba(x, q + "1", stride, act, depthwise=True)
            x = net.se(x, q + "2")
            x = net.cba(x, q + "3", 1, "ID")
        else:
            if cexp != cin:
                x = net.cba(x, bp + str(j), 1, act)
                j += 1
            x = net.cba(x, bp + str(j), stride, act, depthwise=True)
            j += 1
            if use_se:
                x = net.se(x, bp + str(j))
                j += 1
            x = net.cba(x, bp + str(j), 1, "ID",               # mobilenetv3.py:69,95-99
                        residual=inp if (stride == 1 and cin == cout) else None)
        cin = cout
    x = net.cba(x, "backbone.features.1.3", 1, "HS")           # last 1x1, mobilenetv3.py:149-152
    feats.append(x)
    for e in range(4):                                         # _extra_block, ssd_mobilenetv3.py:39-54
        q = "backbone.extra.%d." % e
        x = net.cba(x, q + "0", 1, "R6")
        x = net.cba(x, q + "1", 2, "R6", depthwise=True)
        x = net.cba(x, q + "2", 1, "R6")
        feats.append(x)
    t0 = _tick(times, "backbone", t0)
    cls, reg = [], []
    for l, f in enumerate(feats):                              # _prediction_block, :27-36
        for name, cols, dst in (("classification_head", num_classes, cls), ("regression_head", 4, reg)):
            q = "head.%s.module_list.%d." % (name, l)
            h = net.cba(f, q + "0", 1, "R6", depthwise=True)
            dst.append(_score_layout(net.conv_bias(h, q + "1"), cols))
    out = (torch.cat(cls, 1), torch.cat(reg, 1), [tuple(f.shape[-2:]) for f in feats])
    _tick(times, "head", t0)
    return out + (feats,) if return_features else out


# ---------------------------------------------------------------------------------------
# V2 assembly: backbone.MobileNetWithExtraBlocks (backbone.py:45-67) + box_head.MultiBoxLiteHead
# (box_head.py:37-104).  state_dict keys: backbone.body.*, backbone.extra_blocks.*,
# head.cls_logits.*, head.bbox_pred.*
# ---------------------------------------------------------------------------------------
V2_SETTING = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
              (6, 320, 1, 1)]                                  # mobilenetv2.py:138-147
V2_EXTRAS = [(1280, 512, 0.2), (512, 256, 0.25), (256, 256, 0.5), (256, 64, 0.25)]   # backbone.py:54-58


def _v2_ir(net, x, prefix, inp, oup, stride, hidden, expand):
    """InvertedResidual, mobilenetv2.py:62-100 / backbone.py:81-119 (keys `.conv.N`)."""
    y = x
    j = 0
    if expand:
        y = net.cba(y, prefix + ".conv.%d" % j, 1, "R6")
        j += 1
    y = net.cba(y, prefix + ".conv.%d" % j, stride, "R6", depthwise=True)
    j += 1
    # pw-linear: Conv2d + BN as siblings conv.j / conv.j+1
    return net.cba(y, prefix + ".conv", 1, "ID", conv_key=".%d" % j, bn_key=".%d" % (j + 1),
                   residual=x if (stride == 1 and inp == oup) else None)


def v2_forward_raw(sd, images, mode="fp32", num_classes=21, image_mean=(0.485, 0.456, 0.406),
                   image_std=(0.229, 0.224, 0.225), bp="backbone.", hp="head.", return_features=False, times=None):
    import time
    t0 = time.perf_counter()
    net = _Net(sd, mode, V2_BN_EPS)
    mean = torch.as_tensor(image_mean, dtype=torch.float32)[None, :, None, None]
    std = torch.as_tensor(image_std, dtype=torch.float32)[None, :, None, None]
    x = (images - mean) / std
    t0 = _tick(times, "transform", t0)
    x = net.cba(x, bp + "body.0", 2, "R6")                     # mobilenetv2.py:157
    feats = []
    idx = 1
    cin = 32
    for t, c, n, s in V2_SETTING:
        for i in range(n):
            stride = s if i == 0 else 1
            x = _v2_ir(net, x, bp + "body.%d" % idx, cin, c, stride, int(round(cin * t)), t != 1)
            cin = c
            if idx == 13:                                      # tap "13" -> "0", backbone.py:52
                feats.append(x)
            idx += 1
    x = net.cba(x, bp + "body.18", 1, "R6")                    # 320 -> 1280, mobilenetv2.py:166
    feats.append(x)
    for e, (inp, oup, t) in enumerate(V2_EXTRAS):
        x = _v2_ir(net, x, bp + "extra_blocks.%d" % e, inp, oup, 2, int(round(inp * t)), True)
        feats.append(x)
    t0 = _tick(times, "backbone", t0)
    cls, reg = [], []
    for l, f in enumerate(feats):
        for name, cols, dst in (("cls_logits", num_classes, cls), ("bbox_pred", 4, reg)):
            q = hp + "%s.%d" % (name, l)
            if l < len(feats) - 1:                             # SeperableConv2d, box_head.py:24-34
                h = net.cba(f, q, 1, "R6", depthwise=True, conv_key=".0", bn_key=".1", conv_bias=True)
                o = net.conv_bias(h, q + ".3")
            else:                                              # plain 1x1, box_head.py:55-56
                o = net.conv_bias(f, q)
            dst.append(_score_layout(o, cols))
    out = (torch.cat(cls, 1), torch.cat(reg, 1), [tuple(f.shape[-2:]) for f in feats])
    _tick(times, "head", t0)
    return out + (feats,) if return_features else out


# ---------------------------------------------------------------------------------------
# ssd300_vgg16: SSDFeatureExtractorVGG (ssd_vgg16.py:30-109) + SSDHead with dense 3x3 convolutions
# (generalized_ssd.py:25-92).  state_dict keys: backbone.scale_weight, backbone.features.N, backbone.extra.*, head.*
# ---------------------------------------------------------------------------------------
VGG_FEATURES = [("c", 0), ("c", 2), ("p", False), ("c", 5), ("c", 7), ("p", False), ("c", 10), ("c", 12), ("c", 14), ("p", True),
                ("c", 17), ("c", 19), ("c", 21)]          # torchvision vgg16 cfg "D" up to conv4_3; third pool with ceil_mode
VGG_ANCHORS_PER_LOC = [4, 6, 6, 6, 4, 4]


def vgg_forward_raw(sd, images, mode="fp32", num_classes=91, image_mean=(0.48235, 0.45882, 0.40784),
                    image_std=(1.0 / 255.0, 1.0 / 255.0, 1.0 / 255.0), return_features=False, times=None):
    """ssd300_vgg16 up to the head outputs: (cls_logits f32[B,8732,K], bbox_regression f32[B,8732,4], grid sizes).
    mode fp32 = the reference's op sequence; fp16 / bf16 = the engine's contract (16-bit weights for the tensor-core
    convolutions, every stored activation rounded, fp32 accumulate, fp32 head outputs; the first convolution keeps fp32
    weights like the SSDLite stem)."""
    import time
    t0 = time.perf_counter()
    rnd = {"bf16": _bf16, "fp16": _fp16}.get(mode, _ident)
    wr = rnd

    def conv(x, prefix, stride=1, padding=1, dilation=1, relu=True, round_w=True, out_round=True):
        w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
        y = F.conv2d(x, wr(w) if round_w else w, b, stride, padding, dilation)
        if relu:
            y = F.relu(y)
        return rnd(y) if out_round else y
    mean = torch.as_tensor(image_mean, dtype=torch.float32)[None, :, None, None]
    std = torch.as_tensor(image_std, dtype=torch.float32)[None, :, None, None]
    x = (images - mean) / std                                  # transform.py:129-138
    t0 = _tick(times, "transform", t0)
    for kind, arg in VGG_FEATURES:
        if kind == "c":
            x = conv(x, "backbone.features.%d" % arg, round_w=(arg != 0))
        else:
            x = F.max_pool2d(x, 2, 2, 0, ceil_mode=arg)
    feats = [rnd(sd["backbone.scale_weight"].view(1, -1, 1, 1) * F.normalize(x))]          # ssd_vgg16.py:98-100
    x = F.max_pool2d(x, 2, 2)                                   # backbone[maxpool4_pos:-1]: pool4, conv5_1..5_3
    for i in (1, 3, 5):
        x = conv(x, "backbone.extra.0.%d" % i)
    x = F.max_pool2d(x, 3, 1, 1)                                # modified pool5, ssd_vgg16.py:84
    x = conv(x, "backbone.extra.0.7.1", padding=6, dilation=6)  # fc6, atrous
    x = conv(x, "backbone.extra.0.7.3", padding=0)              # fc7
    feats.append(x)
    for e, (s, p) in zip((1, 2, 3, 4), ((2, 1), (2, 1), (1, 0), (1, 0))):                  # ssd_vgg16.py:48-73
        x = conv(x, "backbone.extra.%d.0" % e, padding=0)
        x = conv(x, "backbone.extra.%d.2" % e, stride=s, padding=p)
        feats.append(x)
    t0 = _tick(times, "backbone", t0)
    cls, reg = [], []
    for l, f in enumerate(feats):                               # SSDScoringHead.forward, generalized_ssd.py:60-74
        for name, cols, dst in (("classification_head", num_classes, cls), ("regression_head", 4, reg)):
            o = conv(f, "head.%s.module_list.%d" % (name, l), relu=False, out_round=False)
            dst.append(_score_layout(o, cols))
    out = (torch.cat(cls, 1), torch.cat(reg, 1), [tuple(f.shape[-2:]) for f in feats])
    _tick(times, "head", t0)
    return out + (feats,) if return_features else out


# ---------------------------------------------------------------------------------------
# Post-processing: torch port of the reference's CPU path (the CPU baseline "port")
# ---------------------------------------------------------------------------------------
def decode_boxes_torch(rel, anchors, weights=(10.0, 10.0, 5.0, 5.0), clip=4.135166556742356):
    """BoxCoder.decode_single (_utils.py:187-224) for one image: f32[P,4] x f32[P,4] -> f32[P,4]."""
    w = anchors[:, 2] - anchors[:, 0]
    h = anchors[:, 3] - anchors[:, 1]
    cx = anchors[:, 0] + 0.5 * w
    cy = anchors[:, 1] + 0.5 * h
    dx, dy = rel[:, 0] / weights[0], rel[:, 1] / weights[1]
    dw = torch.clamp(rel[:, 2] / weights[2], max=clip)
    dh = torch.clamp(rel[:, 3] / weights[3], max=clip)
    pcx, pcy = dx * w + cx, dy * h + cy
    pw, ph = torch.exp(dw) * w, torch.exp(dh) * h
    return torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=1)


def postprocess_detections_torch(cls_logits, bbox_regression, anchors, image_shape, score_thresh=0.001,
                                 nms_thresh=0.55, detections_per_img=300, topk_candidates=300):
    """SSD.postprocess_detections (generalized_ssd.py:351-397) with the same torch / torchvision
    calls and the same image x class Python loops, so that its CPU time is representative of the
    reference.  anchors: f32[P,4]; returns a list of dicts (boxes, scores, labels)."""
    from torchvision.ops import boxes as box_ops
    probs = F.softmax(cls_logits, dim=-1)
    K = probs.size(-1)
    results = []
    for rel, p in zip(bbox_regression, probs):
        boxes = box_ops.clip_boxes_to_image(decode_boxes_torch(rel, anchors), image_shape)
        bs, ss, ls = [], [], []
        for label in range(1, K):
            col = p[:, label]
            sel = col > score_thresh
            col, cand = col[sel], boxes[sel]
            col, order = col.topk(min(topk_candidates, col.size(0)))
            bs.append(cand[order])
            ss.append(col)
            ls.append(torch.full_like(col, fill_value=label, dtype=torch.int64))
        bs, ss, ls = torch.cat(bs), torch.cat(ss), torch.cat(ls)
        keep = box_ops.batched_nms(bs, ss, ls, nms_thresh)[:detections_per_img]
        results.append(OrderedDict(boxes=bs[keep], scores=ss[keep], labels=ls[keep]))
    return results
