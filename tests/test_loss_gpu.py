"""Training-side kernels (dn_ssd_match, dn_match_quality, dn_ssd_loss) against the reference's own results (golden fixture
tests/golden/ssd_loss.npz, produced by the unmodified reference incl. its autograd gradients) and the NumPy oracle.
Matched indices: bit-exact.  Loss values / gradients: fp32 summation-order tolerance (stated per assert)."""
import os

import numpy as np
import pytest
import torch

import demonet_b200
from demonet_b200 import loss as dloss
from oracle import loss_np

pytestmark = pytest.mark.gpu


def _targets_cuda(targets):
    return [{"boxes": torch.from_numpy(b).cuda(), "labels": torch.from_numpy(l).cuda()} for b, l in targets]


@pytest.mark.parametrize("name", sorted(loss_np.LOSS_CASES))
def test_golden_matching_loss_and_gradients(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "ssd_loss.npz"))
    anchors, targets, cls, reg = loss_np.seeded_case(name)
    tt = _targets_cuda(targets)
    a = torch.from_numpy(anchors).cuda()
    matched = dloss.match_targets(tt, [a] * len(tt), 0.5)
    assert matched.dtype == torch.int64
    assert np.array_equal(matched.cpu().numpy(), g[name + "_matched"])                      # bit-exact indices
    # the reference's own call signature: SSDMatcher(thr)(box_iou(gt, anchors)) on its own IoU matrix
    for b, (boxes, _) in enumerate(targets):
        if len(boxes):
            q = torch.from_numpy(loss_np.box_iou(boxes, anchors)).cuda()
            assert np.array_equal(dloss.SSDMatcher(0.5)(q).cpu().numpy(), g[name + "_matched"][b])
    tc = torch.from_numpy(cls).cuda().requires_grad_(True)
    tr = torch.from_numpy(reg).cuda().requires_grad_(True)
    out = dloss.compute_loss(tt, {"cls_logits": tc, "bbox_regression": tr}, [a] * len(tt), list(matched.unbind(0)),
                             float(g["neg_to_pos_ratio"]))
    ref = g[name + "_losses"]
    with torch.no_grad():
        vb, vc = float(out["bbox_regression"]), float(out["classification"])
    assert abs(vb - ref[0]) <= 2e-6 * abs(ref[0])                 # fp32 summation order
    assert abs(vc - ref[1]) <= 2e-6 * abs(ref[1])
    (out["bbox_regression"] + out["classification"]).backward()
    gr, gc = tr.grad.cpu().numpy(), tc.grad.cpu().numpy()
    assert np.allclose(gr[:, ::7], g[name + "_grad_reg"], rtol=1e-5, atol=1e-8)
    assert np.allclose(gc[:, ::53], g[name + "_grad_cls"], rtol=2e-5, atol=1e-8)
    assert abs(np.abs(gc.astype(np.float64)).sum() - float(g[name + "_grad_cls_abs_sum"])) <= 1e-5 * float(g[name + "_grad_cls_abs_sum"])
    # deterministic: a second evaluation gives the same bits
    out2 = dloss.compute_loss(tt, {"cls_logits": tc.detach(), "bbox_regression": tr.detach()}, a, matched, float(g["neg_to_pos_ratio"]))
    assert float(out2["classification"]) == vc and float(out2["bbox_regression"]) == vb


def _random_case(seed, B, P, K, max_gt):
    rng = np.random.default_rng(seed)
    ctr = rng.uniform(20, 300, (P, 2))
    wh = rng.uniform(8, 120, (P, 2))
    anchors = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)
    nd = min(20, P // 2)
    anchors[P // 2:P // 2 + nd] = anchors[:nd]                                   # duplicated default boxes: arg-max ties over dim 1
    targets = []
    for b in range(B):
        m = int(rng.integers(0, max_gt + 1))
        c = rng.uniform(30, 290, (m, 2))
        s = rng.uniform(6, 150, (m, 2))
        boxes = np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32)
        if m > 2:
            boxes[2] = boxes[0]                                                   # duplicated ground truth: ties over dim 0
            boxes[1] = anchors[int(rng.integers(0, P))]
        targets.append((boxes, rng.integers(1, K, m).astype(np.int64)))
    cls = (rng.standard_normal((B, P, K)) * 3).astype(np.float32)
    reg = (rng.standard_normal((B, P, 4)) * 1.2).astype(np.float32)
    return anchors, targets, cls, reg


@pytest.mark.parametrize("seed,B,P,K,max_gt", [(0, 5, 1000, 21, 12), (1, 3, 33, 4, 3), (2, 2, 8732, 91, 40), (3, 7, 257, 2, 60),
                                               (4, 64, 3234, 91, 16)])
def test_random_cases_vs_oracle(seed, B, P, K, max_gt):
    anchors, targets, cls, reg = _random_case(seed, B, P, K, max_gt)
    tt = _targets_cuda(targets)
    a = torch.from_numpy(anchors).cuda()
    matched = dloss.match_targets(tt, a, 0.5).cpu().numpy()
    ref_m = np.stack([loss_np.match_image(b, anchors, 0.5) for b, _ in targets])
    assert np.array_equal(matched, ref_m)
    ref, d = loss_np.compute_loss(targets, cls, reg, anchors, ref_m, 3.0, return_details=True)
    out = dloss.compute_loss(tt, {"cls_logits": torch.from_numpy(cls).cuda(), "bbox_regression": torch.from_numpy(reg).cuda()},
                             a, torch.from_numpy(ref_m).cuda(), 3.0)
    for k in ref:
        assert abs(float(out[k]) - ref[k]) <= 3e-6 * max(1.0, abs(ref[k])), (k, float(out[k]), ref[k])


def test_gradients_vs_torch_autograd():
    """d loss / d head outputs against autograd of a plain torch fp32 restatement (same selection mask)."""
    anchors, targets, cls, reg = _random_case(7, 4, 600, 11, 9)
    tt = _targets_cuda(targets)
    a = torch.from_numpy(anchors).cuda()
    matched = dloss.match_targets(tt, a, 0.5)
    _, d = loss_np.compute_loss(targets, cls, reg, anchors, matched.cpu().numpy(), 3.0, return_details=True)
    tc = torch.from_numpy(cls).cuda().requires_grad_(True)
    tr = torch.from_numpy(reg).cuda().requires_grad_(True)
    out = dloss.compute_loss(tt, {"cls_logits": tc, "bbox_regression": tr}, a, matched, 3.0)
    (2.0 * out["bbox_regression"] + 0.5 * out["classification"]).backward()
    # torch restatement
    rc = torch.from_numpy(cls).cuda().requires_grad_(True)
    rr = torch.from_numpy(reg).cuda().requires_grad_(True)
    ct = torch.from_numpy(d["cls_targets"]).cuda()
    ce = torch.nn.functional.cross_entropy(rc.view(-1, cls.shape[-1]), ct.view(-1), reduction="none").view(ct.shape)
    w = torch.from_numpy(d["foreground"].astype(np.float32) + d["background"].astype(np.float32)).cuda()
    bl = 0
    for b, (boxes, _) in enumerate(targets):
        m = matched[b]
        fg = torch.where(m >= 0)[0]
        if fg.numel():
            t = torch.from_numpy(loss_np.encode_boxes(boxes[m[fg].cpu().numpy()], anchors[fg.cpu().numpy()])).cuda()
            bl = bl + torch.nn.functional.smooth_l1_loss(rr[b][fg], t, reduction="sum")
    (2.0 * bl / d["N"] + 0.5 * (ce * w).sum() / d["N"]).backward()
    assert torch.allclose(tc.grad, rc.grad, rtol=1e-4, atol=1e-7)
    assert torch.allclose(tr.grad, rr.grad, rtol=1e-4, atol=1e-7)


def test_matcher_errors_and_edges():
    with pytest.raises(ValueError, match="No ground-truth boxes"):
        dloss.SSDMatcher(0.5)(torch.zeros(0, 5, device="cuda"))
    q = torch.tensor([[0.9, 0.1], [0.9, 0.1]], device="cuda")                   # both ground-truth boxes claim prediction 0
    assert dloss.SSDMatcher(0.5)(q).tolist() == [1, -1]
    # no ground truth anywhere: every box is background, N = 1, no background sample (0 foreground)
    a = torch.tensor([[0.0, 0.0, 10.0, 10.0]] * 50, device="cuda")
    tt = [{"boxes": torch.zeros(0, 4, device="cuda"), "labels": torch.zeros(0, dtype=torch.int64, device="cuda")}] * 2
    m = dloss.match_targets(tt, a)
    assert (m == -1).all()
    out = dloss.compute_loss(tt, {"cls_logits": torch.randn(2, 50, 3, device="cuda"), "bbox_regression": torch.randn(2, 50, 4, device="cuda")}, a, m)
    assert float(out["bbox_regression"]) == 0.0 and float(out["classification"]) == 0.0
    with pytest.raises(RuntimeError):
        dloss.SSDMatcher(0.5)(torch.rand(2, 3))                                    # CPU tensor: no fallback


@pytest.mark.parametrize("builder,S", [("ssdlite320_mobilenet_v3_large", 320), ("ssd300_vgg16", 300)])
def test_training_mode_forward_returns_the_reference_losses(builder, S):
    """model.train(); model(images, targets) -> the loss dict of SSD.forward's training branch, evaluated on the engine's
    head outputs: equal to the oracle's compute_loss on those head outputs."""
    torch.manual_seed(0)
    model = getattr(demonet_b200, builder)(num_classes=21).cuda()
    images = [torch.rand(3, S, S, device="cuda") for _ in range(3)]
    targets = [{"boxes": torch.tensor([[30.0, 40.0, 200.0, 220.0], [100.0, 90.0, 140.0, 150.0]], device="cuda"),
                "labels": torch.tensor([3, 7], device="cuda")},
               {"boxes": torch.zeros(0, 4, device="cuda"), "labels": torch.zeros(0, dtype=torch.int64, device="cuda")},
               {"boxes": torch.tensor([[10.0, 10.0, 60.0, 90.0]], device="cuda"), "labels": torch.tensor([20], device="cuda")}]
    model.train()
    losses = model(images, targets)
    assert set(losses) == {"bbox_regression", "classification"}
    model.eval()
    cls, reg = model.head_outputs(torch.stack(images))[:2]
    anchors = model.anchors(torch.device("cuda", 0)).cpu().numpy()
    tn = [(t["boxes"].cpu().numpy(), t["labels"].cpu().numpy()) for t in targets]
    mm = np.stack([loss_np.match_image(b, anchors, 0.5) for b, _ in tn])
    ref = loss_np.compute_loss(tn, cls.cpu().numpy(), reg.cpu().numpy(), anchors, mm, 3.0)
    for k in ref:
        assert abs(float(losses[k]) - ref[k]) <= 5e-6 * max(1.0, abs(ref[k])), (k, float(losses[k]), ref[k])
    dets = model(images)                                                           # eval branch still returns detections
    assert len(dets) == 3 and set(dets[0]) == {"boxes", "scores", "labels"}
