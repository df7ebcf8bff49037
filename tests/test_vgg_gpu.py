"""ssd300_vgg16 on the GPU: head outputs and feature maps against the fp32 oracle (bit-equal to the
unmodified reference, see make_golden.py), detections against the reference's golden ones, the engine's NMS kernels fed
the reference's own scores and boxes (bit-exact), and the API behaviour (resize path, CPU inputs)."""
import os

import numpy as np
import pytest
import torch

import demonet_b200
from demonet_b200 import ops
from oracle import boxes_np, net_ref, parity, weights

pytestmark = pytest.mark.gpu
# stated tolerance vs the fp32 reference on the seeded network (logits std 1.6; 16 layers of 16-bit storage)
TOL = {"fp16": dict(rel=0.005, score_mean=2e-5, box_mean=0.05, top100=0.97), "bf16": dict(rel=0.03, score_mean=2e-4, box_mean=0.4, top100=0.9)}


def _model(**kw):
    m = demonet_b200.ssd300_vgg16(**kw)
    sd = weights.seeded_vgg_state_dict(m.state_dict())
    m.load_state_dict(sd)
    return m.cuda(), sd


def _rel(a, b):
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt())


@pytest.mark.parametrize("act_dtype,first_conv", [("fp16", "gemm"), ("fp16", "direct"), ("bf16", "gemm")])
def test_head_outputs_features_and_detections(golden_dir, act_dtype, first_conv):
    g = np.load(os.path.join(golden_dir, "ssd300_vgg16.npz"))
    model, sd = _model(act_dtype=act_dtype, first_conv=first_conv)
    x = weights.synthetic_images(2, 300)
    cls, reg, feats = model.head_outputs(x.cuda(), return_features=True)
    with torch.no_grad():
        ocls, oreg, _, ofeats = net_ref.vgg_forward_raw(sd, x, "fp32", return_features=True)
    tol = TOL[act_dtype]
    for l, (f, of) in enumerate(zip(feats, ofeats)):          # the six feature maps, NHWC 16-bit vs NCHW fp32
        assert f.shape[1:3] == of.shape[2:] and _rel(f.float().permute(0, 3, 1, 2).cpu(), of) < tol["rel"], l
    stride = int(g["row_stride"])
    assert _rel(cls[:, ::stride].cpu(), torch.from_numpy(g["logits_rows"])) < tol["rel"]      # rows the reference itself produced
    assert _rel(reg[:, ::stride].cpu(), torch.from_numpy(g["bbox_rows"])) < tol["rel"]
    anchors = model.anchors(torch.device("cuda")).cpu().numpy()
    dets = model([x[0].cuda(), x[1].cuda()])
    m = parity.head_metrics(cls.cpu(), reg.cpu(), ocls, oreg, anchors, (300, 300))
    gold = [{"boxes": g["det_boxes"][i], "labels": g["det_labels"][i]} for i in range(2)]
    m["top100_match"] = parity.detection_match(dets, gold)                                      # the reference's own detections
    print("[parity] ssd300_vgg16 B=2 vs fp32 reference [%s, first conv %s]: %s" % (act_dtype, first_conv, parity.format_metrics(m)))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        import json
        with open(os.path.join(out, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"case": "ssd300_vgg16 B=2 vs fp32 reference [%s, first conv %s]" % (act_dtype, first_conv), **m}) + "\n")
    assert m["logits_rel_rms"] < tol["rel"] and m["bbox_rel_rms"] < tol["rel"], m
    assert m["score_mean_abs"] < tol["score_mean"] and m["box_mean_abs_px"] < tol["box_mean"], m
    assert m["top100_match"] >= tol["top100"], m
    for d in dets:
        assert d["boxes"].shape == (200, 4) and d["labels"].dtype == torch.int64 and bool((d["scores"][:-1] >= d["scores"][1:]).all())


def test_nms_kernels_on_the_reference_scores_bit_exact(golden_dir):
    """The reference model's golden detections from the reference's own softmax scores and decoded boxes, through the
    engine's kernels: 8732 anchors, thr 0.01, top-k 400, NMS 0.45, D 200."""
    import hashlib
    from torchvision.ops import boxes as box_ops
    g = np.load(os.path.join(golden_dir, "ssd300_vgg16.npz"))
    sd = weights.seeded_vgg_state_dict(demonet_b200.ssd300_vgg16().state_dict())
    x = weights.synthetic_images(2, 300)
    with torch.no_grad():
        cls, reg, grids = net_ref.vgg_forward_raw(sd, x, "fp32")
    if hashlib.sha256(np.ascontiguousarray(cls.numpy()).tobytes()).hexdigest() != str(g["logits_sha256"]):
        pytest.skip("this host's ATen conv kernels do not reproduce the golden logits bit for bit")
    anchors = torch.from_numpy(boxes_np.default_boxes(grids, (300, 300), [[2], [2, 3], [2, 3], [2, 3], [2], [2]],
                                                      scales=[0.07, 0.15, 0.33, 0.51, 0.69, 0.87, 1.05], steps=[8, 16, 32, 64, 100, 300]))
    scores = torch.softmax(cls, -1)
    boxes = torch.stack([box_ops.clip_boxes_to_image(net_ref.decode_boxes_torch(r, anchors), (300, 300)) for r in reg])
    ob, os_, ol, oc, _, _ = [t.cpu().numpy() for t in ops.postprocess_scored(scores.cuda(), boxes.cuda(), (300, 300), 0.01, 0.45, 200, 400)]
    for i in range(2):
        n = g["det_scores"][i].shape[0]
        assert oc[i] == n and np.array_equal(ol[i, :n], g["det_labels"][i])
        assert np.array_equal(os_[i, :n].view(np.uint32), g["det_scores"][i].view(np.uint32))
        assert np.array_equal(ob[i, :n].view(np.uint32), g["det_boxes"][i].view(np.uint32))


def test_api_resize_cpu_inputs_and_weight_updates():
    model, sd = _model()
    gen = torch.Generator().manual_seed(3)
    imgs = [torch.rand(3, 300, 300, generator=gen), torch.rand(3, 500, 400, generator=gen)]       # the docstring's example sizes
    dev = model([i.cuda() for i in imgs])
    host = model(imgs)                                                                         # CPU tensors in -> CPU tensors out
    assert host[0]["boxes"].device.type == "cpu" and torch.equal(host[1]["scores"], dev[1]["scores"].cpu())
    assert float(dev[1]["boxes"][:, 0::2].max()) <= 400.0 + 1e-3 and float(dev[1]["boxes"][:, 1::2].max()) <= 500.0 + 1e-3
    again = model([imgs[0].cuda()])
    assert torch.equal(again[0]["scores"], dev[0]["scores"])                                    # independent of the batch
    model.load_state_dict(weights.seeded_vgg_state_dict(model.state_dict(), seed=7))
    assert not torch.equal(model([imgs[0].cuda()])[0]["scores"], dev[0]["scores"])
    with pytest.raises(ValueError):
        model([torch.rand(300, 300).cuda()])
    with pytest.raises(TypeError):
        model([torch.zeros(3, 300, 300, dtype=torch.uint8).cuda()])
    assert model([]) == []
