"""The whole engine (SSD.forward eval branch) on the GPU against the oracle and the reference's golden
outputs: per-layer teacher-forced parity (tight), end-to-end logits (statistical -- the bf16 chain is
chaotic, see DESIGN.md "Numerics"), detections, API behaviour."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import demonet_b200
from demonet_b200 import plan as dplan
from demonet_b200 import _C
from oracle import boxes_np, net_ref, parity, weights

pytestmark = pytest.mark.gpu
ACTS = {"none": lambda v: v, "relu": F.relu, "relu6": F.relu6, "hardswish": F.hardswish}
DTYPES = ["fp16", "bf16"]
ULP = {"fp16": 2.0 ** -10, "bf16": 2.0 ** -7}
# Stated end-to-end tolerance vs the fp32 reference on the seeded random-weight network (logits std 4.1, ~60 layers of
# 16-bit storage; DESIGN.md "Numerics" has the measured values): (logits rel-rms, mean |score err|, 99.9th percentile
# |score err|, mean |box err| px, top-100 detection match)
E2E_TOL = {"fp16": dict(rel=0.02, score_mean=4e-4, score_p999=0.03, box_mean=0.6, top100=0.95),
           "bf16": dict(rel=0.12, score_mean=3e-3, score_p999=0.2, box_mean=4.0, top100=0.80)}


def _model(builder, **kw):
    m = builder(**kw)
    sd = weights.seeded_state_dict(m.state_dict())
    m.load_state_dict(sd)
    return m.cuda(), sd


def _layerwise_check(model, sd, x):
    """Teacher-forced: every layer's output buffer vs a torch fp32 evaluation of that layer on the
    ENGINE's own input buffer(s).  Needs keep_activations=True (one buffer per tensor)."""
    B = x.shape[0]
    model.head_outputs(x)
    eng = model._engine_for(x.device, B)
    plan = model.plan
    cls, reg = eng.head_outputs(B)
    mean = torch.tensor(model.image_mean, device="cuda")[None, :, None, None]
    std = torch.tensor(model.image_std, device="cuda")[None, :, None, None]
    level_off, o = [], 0
    for h, w in plan.grid_sizes:
        level_off.append(o)
        o += h * w * plan.anchors_per_loc
    layers = plan.layers
    sd = {k: v.cuda() for k, v in sd.items()}
    i = 0
    checked = 0
    while i < len(layers):
        L = layers[i]
        wf, bf = dplan.fold_layer(sd, L, plan.bn_eps) if L.kind != "se" else (None, None)
        h16 = _C.torch_dtype(model.act_dtype)
        rel_tol = ULP[model.act_dtype]           # 1 ulp of the storage type
        src = (x - mean) / std if L.kind == "stem" else eng.buffer(L.src, B)
        if L.kind == "stem":
            ref = ACTS[L.act](F.conv2d(src, wf.float().cuda(), bf.float().cuda(), 2, 1))
        elif L.kind == "dw":
            ref = ACTS[L.act](F.conv2d(src, wf.float().cuda(), bf.float().cuda(), L.stride, (L.k - 1) // 2, 1, L.cin))
            if i + 1 < len(layers) and layers[i + 1].kind == "se":       # SE runs in place on the dw output
                S = layers[i + 1]
                ref = ref.to(h16).float()
                s = F.adaptive_avg_pool2d(ref, 1)
                s = F.relu(F.conv2d(s, sd[S.conv + ".fc1.weight"], sd[S.conv + ".fc1.bias"]))
                s = F.hardsigmoid(F.conv2d(s, sd[S.conv + ".fc2.weight"], sd[S.conv + ".fc2.bias"]))
                ref = ref * s
                rel_tol = 2 * ULP[model.act_dtype]        # dw -> SE compounds two roundings
                i += 1
        else:
            w16 = wf.float().to(h16).float().cuda()
            ref = ACTS[L.act](F.conv2d(src, w16, bf.float().cuda()))
            if L.res:
                ref = ref + eng.buffer(L.res, B)
        if L.head:
            kind, lvl = L.head
            cols = plan.num_classes if kind == "cls" else 4
            got = (cls if kind == "cls" else reg)[:, level_off[lvl]:level_off[lvl] + L.h_in * L.w_in * plan.anchors_per_loc]
            want = ref.view(B, plan.anchors_per_loc, cols, L.h_in, L.w_in).permute(0, 3, 4, 1, 2).reshape(B, -1, cols)
            assert float((got - want).abs().max()) < 5e-3, (i, L.conv)
        else:
            got = eng.buffer(L.dst if L.kind != "se" else L.src, B)
            tol = ref.abs() * rel_tol + 4e-3
            bad = (got - ref).abs() > tol
            assert float(bad.float().mean()) < 1e-4, (i, L.conv, float((got - ref).abs().max()))
        checked += 1
        i += 1
    return checked


@pytest.mark.parametrize("act_dtype", DTYPES)
def test_v3_layerwise_simt_selfcheck(act_dtype):
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large, keep_activations=True, gemm_impl=1,
                       use_cuda_graph=False, act_dtype=act_dtype)
    x = weights.synthetic_images(2, 320).cuda()
    assert _layerwise_check(model, sd, x) == 82


@pytest.mark.parametrize("act_dtype", DTYPES)
def test_v3_layerwise(act_dtype):
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large, keep_activations=True, use_cuda_graph=False,
                       act_dtype=act_dtype)
    x = weights.synthetic_images(2, 320).cuda()
    assert _layerwise_check(model, sd, x) == 82


@pytest.mark.parametrize("act_dtype", DTYPES)
@pytest.mark.parametrize("S", [300, 512])
def test_v2_layerwise(S, act_dtype):
    model, sd = _model(demonet_b200.ssd_lite_mobilenet_v2, image_size=S, keep_activations=True, use_cuda_graph=False,
                       act_dtype=act_dtype)
    x = weights.synthetic_images(2, S).cuda()
    assert _layerwise_check(model, sd, x) > 80


def _rel_rms(a, b):
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt())


def _same_image_other_batch(a, b, frac=0.95):
    """Detections of one image computed inside batches of different sizes.  The squeeze-excitation pooling splits its fp32
    sums differently per batch size, the pooled means then differ in their last bits and the 16-bit storage amplifies
    that to a small fraction of its own rounding noise (DESIGN.md "Numerics"): the same detections (label, IoU >= 0.9)
    up to a few near-ties at the tail -- compared as sets, a single swap shifts every later position."""
    assert abs(a["scores"].numel() - b["scores"].numel()) <= 3
    assert parity.detection_match([a], [b], top=100, iou_thr=0.9) >= frac
    assert parity.detection_match([b], [a], top=100, iou_thr=0.9) >= frac
    assert float((a["scores"][:20] - b["scores"][:20]).abs().max()) < 2e-2


def _report(name, metrics):
    """Parity numbers in the north-star's units: printed (pytest -s / failure output) and collected under gpurun_out/."""
    import json
    print("[parity] %s: %s" % (name, parity.format_metrics(metrics)))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, **metrics}) + "\n")


def _check_e2e(name, act_dtype, cls, reg, dets, ref_cls, ref_reg, anchors, size=(320, 320), **post):
    """Engine head outputs + detections of some images against the fp32 oracle of the same images, in the north-star's
    units, against the stated tolerance of the storage type."""
    m = parity.head_metrics(cls, reg, ref_cls, ref_reg, anchors, size)
    m["top100_match"] = parity.detection_match(dets, parity.reference_detections(ref_cls, ref_reg, anchors, size, **post))
    _report("%s [%s]" % (name, act_dtype), m)
    tol = E2E_TOL[act_dtype]
    assert m["logits_rel_rms"] < tol["rel"] and m["bbox_rel_rms"] < tol["rel"], m
    assert m["score_mean_abs"] < tol["score_mean"] and m["score_p999_abs"] < tol["score_p999"], m
    assert m["box_mean_abs_px"] < tol["box_mean"], m
    assert m["top100_match"] >= tol["top100"], m
    return m


@pytest.mark.parametrize("act_dtype", DTYPES)
def test_v3_end_to_end_against_golden(golden_dir, act_dtype):
    """Head outputs vs the reference's golden tensors (rows the reference itself produced), then scores / decoded boxes /
    detections vs the fp32 oracle in the north-star's units (stated tolerance: E2E_TOL)."""
    g = np.load(os.path.join(golden_dir, "v3_ssdlite.npz"))
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large, act_dtype=act_dtype)
    x = weights.synthetic_images(2, 320)
    cls, reg = model.head_outputs(x.cuda())
    stride = int(g["row_stride"])
    ref_rows = torch.from_numpy(g["logits_rows"])
    got_rows = cls[:, ::stride].cpu()
    tol = E2E_TOL[act_dtype]
    assert _rel_rms(got_rows, ref_rows) < tol["rel"]
    if act_dtype == "bf16":
        assert _rel_rms(got_rows, torch.from_numpy(g["logits_bf16emu_rows"])) < 0.10
    assert _rel_rms(reg.cpu(), torch.from_numpy(g["bbox_regression"])) < tol["rel"]
    dets = model([x[0].cuda(), x[1].cuda()])
    for i, d in enumerate(dets):
        assert d["boxes"].shape == (300, 4) and d["labels"].dtype == torch.int64
        s = d["scores"]
        assert bool((s[:-1] >= s[1:]).all())
    gold = [{"boxes": g["det_boxes"][i], "labels": g["det_labels"][i]} for i in range(2)]
    assert parity.detection_match(dets, gold) >= tol["top100"]           # the reference's own detections
    with torch.no_grad():
        ocls, oreg, _ = net_ref.v3_forward_raw(sd, x, "fp32")
    assert _rel_rms(ocls[:, ::stride], ref_rows) < 1e-5                   # the oracle reproduces the golden rows here
    _check_e2e("v3 B=2 vs fp32 reference", act_dtype, cls.cpu(), reg.cpu(), dets, ocls, oreg, g["anchors"])


def test_fp16_is_closer_to_the_reference_than_bf16():
    """The default storage type must be the better one: fp16 logits at least 4x closer (rel-rms) than bf16."""
    x = weights.synthetic_images(4, 320, seed=21)
    err = {}
    for dt in DTYPES:
        model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large, act_dtype=dt)
        cls, reg = model.head_outputs(x.cuda())
        with torch.no_grad():
            ocls, oreg, _ = net_ref.v3_forward_raw(sd, x, "fp32")
        err[dt] = _rel_rms(cls.cpu(), ocls)
        assert torch.isfinite(cls).all() and torch.isfinite(reg).all()
    assert _C.DEFAULT_ACT_DTYPE == "fp16" or "DN_ACT_DTYPE" in os.environ
    assert err["fp16"] * 4 < err["bf16"], err


def test_benchmarked_configuration_against_oracle():
    """The configuration bench.py times -- batch 256, pipeline_slots=4, CUDA-graph replay, fused blocks, squeeze-excitation
    pooled by the depthwise row streams -- compared with the fp32 oracle on 16 sampled images, after asserting that those
    paths were really taken (dn_engine_get_stats).  Also: an image's detections do not depend on the batch it travels in."""
    B = 256
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large, pipeline_slots=4)
    x = weights.synthetic_images(B, 320, seed=77)
    xd = x.cuda()
    outs = list(model.forward_batches([xd] * 12))         # per slot: eager, capture, replay
    eng = model._engine_for(xd.device, B)
    st = eng.stats()
    assert st["pipeline_slots"] == 4 and st["graph_replays"] >= 4, st
    assert st["fused_pwdw"] == 1 and st["fused_dwpw"] == 1, st
    assert st["se_layers"] == 8 and st["se_pooled"] == 8 and st["se_folded"] == 8, st
    assert st["act_dtype"] == model.act_dtype
    dets = outs[-1]                                       # a replayed graph on the last slot
    for o in outs[:-1]:
        for a, b in zip(o[::37], dets[::37]):
            assert torch.equal(a["scores"], b["scores"]) and torch.equal(a["labels"], b["labels"])
    cls, reg = eng.head_outputs(B)                        # arena of the slot that ran last
    idx = list(range(0, B, 16))
    with torch.no_grad():
        ocls, oreg, _ = net_ref.v3_forward_raw(sd, x[idx], "fp32")
    anchors = dplan.default_boxes(model.plan)
    _check_e2e("v3 B=256 pipelined/graph/fused/pooled-SE, 16 sampled images", model.act_dtype, cls[idx].cpu(), reg[idx].cpu(),
               [dets[i] for i in idx], ocls, oreg, anchors)
    # post-processing of the engine's own head outputs equals the NumPy oracle's (same logits in, exp-ulp effects only)
    for i in idx[:4]:
        o = boxes_np.postprocess_detections(None, reg[i].cpu().numpy(), anchors, (320, 320),
                                            scores=torch.softmax(cls[i].cpu(), -1).numpy())
        same = (dets[i]["labels"].cpu().numpy() == o["labels"]) & (np.abs(dets[i]["scores"].cpu().numpy() - o["scores"]) < 3e-6)
        assert same.mean() >= 0.98
    # batch independence: the same images alone on a plain (unpipelined) engine of batch 16 -- a different split of the
    # SE pooling sums (fp32 sums of 16-bit values in another order) is the only difference, so the head outputs agree
    # to a small fraction of the storage rounding
    plain, _ = _model(demonet_b200.ssdlite320_mobilenet_v3_large)
    cls16, reg16 = plain.head_outputs(xd[idx])
    assert _rel_rms(cls16.cpu(), cls[idx].cpu()) < 0.25 * E2E_TOL[model.act_dtype]["rel"]


def test_v3_postprocess_exact_on_engine_logits():
    """The engine's own head outputs pushed through the NumPy oracle post-processing give the engine's
    detections (isolates post-processing from conv numerics at full model scale)."""
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large)
    x = weights.synthetic_images(3, 320).cuda()
    dets = model(list(x))
    cls, reg = model._engine_for(x.device, 3).head_outputs(3)
    anchors = dplan.default_boxes(model.plan)
    for i in range(3):
        o = boxes_np.postprocess_detections(None, reg[i].cpu().numpy(), anchors, (320, 320),
                                            scores=torch.softmax(cls[i].cpu(), -1).numpy())
        lab = dets[i]["labels"].cpu().numpy()
        same = (lab == o["labels"]) & (np.abs(dets[i]["scores"].cpu().numpy() - o["scores"]) < 3e-6)
        assert same.mean() >= 0.98


def test_graph_replay_is_deterministic_and_matches_eager():
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large)
    x = list(weights.synthetic_images(4, 320).cuda())
    outs = [model(x) for _ in range(4)]          # call 1 eager, call 2 captures, 3+ replay
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a["scores"], b["scores"]) and torch.equal(a["boxes"], b["boxes"]) and torch.equal(a["labels"], b["labels"])


def test_batch_independence_and_cpu_inputs():
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large)
    x = weights.synthetic_images(5, 320)
    full = model(list(x.cuda()))
    single = model([x[3].cuda()])
    _same_image_other_batch(full[3], single[0])
    host = model(list(x))                         # CPU tensors: pinned H2D -> forward -> D2H, results on the CPU
    assert host[0]["boxes"].device.type == "cpu"
    assert torch.equal(host[2]["scores"], full[2]["scores"].cpu())


def test_non_square_input_is_resized_and_boxes_rescaled():
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large)
    img = torch.rand(3, 256, 275, generator=torch.Generator().manual_seed(3)).cuda()     # test_demonet_tracing.cpp:31-33
    d = model([img])[0]
    assert float(d["boxes"][:, 0::2].max()) <= 275.0 + 1e-3 and float(d["boxes"][:, 1::2].max()) <= 256.0 + 1e-3


def test_weight_update_is_picked_up():
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large)
    x = list(weights.synthetic_images(1, 320).cuda())
    a = model(x)[0]["scores"].clone()
    sd2 = weights.seeded_state_dict(model.state_dict(), seed=99)
    model.load_state_dict(sd2)
    b = model(x)[0]["scores"]
    assert not torch.equal(a, b)


def test_v2_legacy_hub_entry():
    import hubconf
    m = hubconf.ssd_lite_mobilenet_v2(pretrained=False, image_size=300, score_thresh=0.05, num_classes=21)
    sd = weights.seeded_state_dict(m.state_dict())
    m.load_state_dict(sd)
    m = m.cuda()
    x = weights.synthetic_images(1, 300)
    d = m([x[0].cuda()])[0]
    assert d["boxes"].shape[0] <= 100 and d["labels"].numel() > 0
    cls, reg = m._engine_for(torch.device("cuda", 0), 1).head_outputs(1)
    priors = dplan.default_boxes(m.plan)
    o = boxes_np.legacy_postprocess(None, reg[0].cpu().numpy(), priors, (300, 300), score_thresh=0.05,
                                    scores=torch.softmax(cls[0].cpu(), -1).numpy())
    same = (d["labels"].cpu().numpy() == o["labels"]) & (np.abs(d["scores"].cpu().numpy() - o["scores"]) < 3e-6)
    assert same.mean() >= 0.97


def test_no_detections_and_batch_growth():
    """score_thresh = 1.0 can never be exceeded by a softmax score -> empty results (test_onnx.py:125-133
    'image with no detections'); a larger batch afterwards re-creates the engine transparently."""
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large, score_thresh=1.0)
    x = weights.synthetic_images(7, 320).cuda()
    out = model([x[0], x[1]])
    assert len(out) == 2
    for d in out:
        assert d["boxes"].shape == (0, 4) and d["scores"].shape == (0,) and d["labels"].shape == (0,)
        assert d["labels"].dtype == torch.int64
    model2, _ = _model(demonet_b200.ssdlite320_mobilenet_v3_large)
    a = model2([x[0]])
    b = model2(list(x))                              # max batch grows 1 -> 7
    assert len(b) == 7
    _same_image_other_batch(a[0], b[0])
    c = model2(x)                                    # a batched 4-D tensor is accepted like a list
    assert torch.equal(c[6]["labels"], b[6]["labels"])
    assert model2([]) == []


def test_default_box_generator_op():
    from demonet_b200 import ops

    class _IL:                                        # minimal ImageList stand-in (tensors, image_sizes)
        def __init__(self, t):
            self.tensors, self.image_sizes = t, [tuple(t.shape[-2:])] * t.shape[0]
    gen = ops.DefaultBoxGenerator([[2, 3]] * 6, min_ratio=0.2, max_ratio=0.95)
    feats = [torch.zeros(2, 8, s, s, device="cuda") for s in (20, 10, 5, 3, 2, 1)]
    out = gen(_IL(torch.zeros(2, 3, 320, 320, device="cuda")), feats)
    want = boxes_np.default_boxes([(s, s) for s in (20, 10, 5, 3, 2, 1)], (320, 320))
    assert len(out) == 2 and np.array_equal(out[0].cpu().numpy(), want)
    assert gen.num_anchors_per_location() == [6] * 6


@pytest.mark.parametrize("slots", [2, 3, 4])
def test_pipeline_mode_matches_plain_engine(slots):
    """pipeline_slots=n: consecutive batches go round n engine instances on engine-owned streams.  The detections must
    equal the plain engine's, batch by batch, for the streaming API (more and fewer batches than slots), single calls and
    host inputs."""
    plain, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large)
    piped, _ = _model(demonet_b200.ssdlite320_mobilenet_v3_large, pipeline_slots=slots)
    batches = [weights.synthetic_images(4, 320, seed=10 + i).cuda() for i in range(9)]
    want = [plain(list(b)) for b in batches]
    got = list(piped.forward_batches(batches)) + list(piped.forward_batches(batches[:2]))
    assert len(got) == len(want) + 2
    for g, w in zip(got, want + want[:2]):
        for a, b in zip(g, w):
            assert torch.equal(a["scores"], b["scores"]) and torch.equal(a["boxes"], b["boxes"]) and torch.equal(a["labels"], b["labels"])
    for _ in range(3):                                   # single calls join at once; slots keep alternating underneath
        one = piped(list(batches[2]))
        assert all(torch.equal(a["scores"], b["scores"]) for a, b in zip(one, want[2]))
    host = [piped(list(b.cpu())) for b in batches[:3]]    # pinned host path, also alternating slots
    for g, w in zip(host, want[:3]):
        assert all(torch.equal(a["boxes"], b["boxes"].cpu()) for a, b in zip(g, w))
    cls_p, _ = piped._engine_for(batches[0].device, 4).head_outputs(4)      # arena of the slot that ran last
    plain(list(batches[2]))
    cls_q, _ = plain._engine_for(batches[0].device, 4).head_outputs(4)
    assert torch.equal(cls_p, cls_q)


def test_scripted_module_matches_eager():
    """The reference's main test (test/test_model.py:85-119): torch.jit.script the detector, run two 3x320x320 images,
    scripted detections == eager detections.  Here the scripted graph calls the registered operator."""
    from demonet_b200 import custom_ops
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large)
    x = [t for t in weights.synthetic_images(2, 320, seed=5).cuda()]
    want = model(x)
    scripted = torch.jit.script(custom_ops.ScriptableSSDLite(model))
    assert "demonet_b200::ssdlite_forward" in str(scripted.graph)
    losses, got = scripted(x)
    assert losses == {} and len(got) == 2
    for g, w in zip(got, want):
        assert torch.equal(g["scores"], w["scores"]) and torch.equal(g["labels"], w["labels"]) and torch.equal(g["boxes"], w["boxes"])
    with pytest.raises(Exception):
        scripted([x[0][:, :100]])


def test_exported_program_runs_the_engine():
    """torch.export captures the detector through torch.ops.demonet_b200.ssdlite_forward (static, padded
    output shapes) and the exported program reproduces the module's detections; the post-processing op exports too."""
    from demonet_b200 import custom_ops, ops
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large)
    x = weights.synthetic_images(3, 320).cuda()
    want = model(list(x))
    wrapper = custom_ops.ExportableSSDLite(model)
    ep = torch.export.export(wrapper, (x,))
    assert "demonet_b200.ssdlite_forward" in ep.graph_module.code
    boxes, scores, labels, counts = ep.module()(x)
    for i, w in enumerate(want):
        n = int(counts[i])
        assert n == w["scores"].numel() and torch.equal(scores[i, :n], w["scores"]) and torch.equal(boxes[i, :n], w["boxes"])
        assert torch.equal(labels[i, :n], w["labels"])

    class Post(torch.nn.Module):
        def forward(self, lg, bb, an):
            return torch.ops.demonet_b200.postprocess(lg, bb, an, 320, 320, 0.001, 0.55, 300, 300, -1.0)

    cls, reg = model.head_outputs(x)
    anchors = torch.from_numpy(dplan.default_boxes(model.plan)).cuda()
    ep2 = torch.export.export(Post(), (cls, reg, anchors))
    b2, s2, l2, c2 = ep2.module()(cls, reg, anchors)
    assert torch.equal(c2, counts) and torch.equal(s2, scores) and torch.equal(b2, boxes)
    keep = torch.ops.demonet_b200.nms(boxes[0, :int(counts[0])], scores[0, :int(counts[0])], 0.5)
    assert torch.equal(keep, ops.nms(boxes[0, :int(counts[0])], scores[0, :int(counts[0])], 0.5))
