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
from oracle import boxes_np, net_ref, weights

pytestmark = pytest.mark.gpu
ACTS = {"none": lambda v: v, "relu": F.relu, "relu6": F.relu6, "hardswish": F.hardswish}


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
        rel_tol = 2.0 ** -7                      # 1 bf16 ulp
        src = (x - mean) / std if L.kind == "stem" else eng.buffer(L.src, B)
        if L.kind == "stem":
            ref = ACTS[L.act](F.conv2d(src, wf.float().cuda(), bf.float().cuda(), 2, 1))
        elif L.kind == "dw":
            ref = ACTS[L.act](F.conv2d(src, wf.float().cuda(), bf.float().cuda(), L.stride, (L.k - 1) // 2, 1, L.cin))
            if i + 1 < len(layers) and layers[i + 1].kind == "se":       # SE runs in place on the dw output
                S = layers[i + 1]
                ref = ref.bfloat16().float()
                s = F.adaptive_avg_pool2d(ref, 1)
                s = F.relu(F.conv2d(s, sd[S.conv + ".fc1.weight"], sd[S.conv + ".fc1.bias"]))
                s = F.hardsigmoid(F.conv2d(s, sd[S.conv + ".fc2.weight"], sd[S.conv + ".fc2.bias"]))
                ref = ref * s
                rel_tol = 2.0 ** -6              # dw -> SE compounds two bf16 roundings
                i += 1
        else:
            w16 = wf.float().bfloat16().float().cuda()
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


def test_v3_layerwise_simt_selfcheck():
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large, keep_activations=True, gemm_impl=1,
                       use_cuda_graph=False)
    x = weights.synthetic_images(2, 320).cuda()
    assert _layerwise_check(model, sd, x) == 82


def test_v3_layerwise():
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large, keep_activations=True, use_cuda_graph=False)
    x = weights.synthetic_images(2, 320).cuda()
    assert _layerwise_check(model, sd, x) == 82


@pytest.mark.parametrize("S", [300, 512])
def test_v2_layerwise(S):
    model, sd = _model(demonet_b200.ssd_lite_mobilenet_v2, image_size=S, keep_activations=True, use_cuda_graph=False)
    x = weights.synthetic_images(2, S).cuda()
    assert _layerwise_check(model, sd, x) > 80


def _rel_rms(a, b):
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt())


def test_v3_end_to_end_against_golden(golden_dir):
    """Stated tolerance (bf16 activations, ~60 layers): logits rel-rms <= 12 % of the fp32 reference
    (the bf16-emulating oracle itself sits at 5.2 %, and moves by 3.4 % when only its summation order
    changes); box regression likewise; detections: >= 80 % of the reference's top-100 detections are
    matched by a detection of the same label with IoU >= 0.5."""
    g = np.load(os.path.join(golden_dir, "v3_ssdlite.npz"))
    model, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large)
    x = weights.synthetic_images(2, 320)
    cls, reg = model.head_outputs(x.cuda())
    stride = int(g["row_stride"])
    ref_rows = torch.from_numpy(g["logits_rows"])
    emu_rows = torch.from_numpy(g["logits_bf16emu_rows"])
    got_rows = cls[:, ::stride].cpu()
    assert _rel_rms(got_rows, ref_rows) < 0.12
    assert _rel_rms(got_rows, emu_rows) < 0.10
    assert _rel_rms(reg.cpu(), torch.from_numpy(g["bbox_regression"])) < 0.12
    dets = model([x[0].cuda(), x[1].cuda()])
    from torchvision.ops import box_iou
    for i, d in enumerate(dets):
        assert d["boxes"].shape == (300, 4) and d["labels"].dtype == torch.int64
        s = d["scores"]
        assert bool((s[:-1] >= s[1:]).all())
        rb, rl = torch.from_numpy(g["det_boxes"][i][:100]), torch.from_numpy(g["det_labels"][i][:100])
        iou = box_iou(rb, d["boxes"].cpu())
        same = rl[:, None] == d["labels"].cpu()[None, :]
        matched = ((iou >= 0.5) & same).any(1).float().mean()
        assert float(matched) >= 0.80, float(matched)


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
    assert torch.equal(full[3]["scores"], single[0]["scores"]) and torch.equal(full[3]["labels"], single[0]["labels"])
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
    assert len(b) == 7 and torch.equal(a[0]["scores"], b[0]["scores"]) and torch.equal(a[0]["boxes"], b[0]["boxes"])
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


def test_pipeline_mode_matches_plain_engine():
    """pipeline_slots=2: consecutive batches alternate between two engine instances on engine-owned streams.  The
    detections must equal the plain engine's, batch by batch, for the streaming API, single calls and host inputs."""
    plain, sd = _model(demonet_b200.ssdlite320_mobilenet_v3_large)
    piped, _ = _model(demonet_b200.ssdlite320_mobilenet_v3_large, pipeline_slots=2)
    batches = [weights.synthetic_images(4, 320, seed=10 + i).cuda() for i in range(5)]
    want = [plain(list(b)) for b in batches]
    got = list(piped.forward_batches(batches))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert torch.equal(a["scores"], b["scores"]) and torch.equal(a["boxes"], b["boxes"]) and torch.equal(a["labels"], b["labels"])
    for _ in range(3):                                   # single calls join at once; slots keep alternating underneath
        one = piped(list(batches[2]))
        assert all(torch.equal(a["scores"], b["scores"]) for a, b in zip(one, want[2]))
    host = [piped(list(b.cpu())) for b in batches[:3]]    # pinned host path, also alternating slots
    for g, w in zip(host, want[:3]):
        assert all(torch.equal(a["boxes"], b["boxes"].cpu()) for a, b in zip(g, w))
    with pytest.raises(RuntimeError):
        piped._engine_for(batches[0].device, 4).head_outputs(4)


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
