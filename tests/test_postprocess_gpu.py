"""Fused post-processing (softmax + decode + threshold + top-k + NMS + top-D) on the GPU, entered with LOGITS, against
the reference's results (golden) and the NumPy/C oracle.  What is tested here is the floating-point front (softmax with
ex2.approx, decode with expf) -- the index-level, bit-exact parity of everything behind it (class sort, lazy warp / CTA
NMS, top-D merge) is tests/test_nms_engine_path_gpu.py, which feeds the reference's own scores and boxes."""
import os

import numpy as np
import pytest
import torch

from demonet_b200 import ops
from oracle import boxes_np, nms_c

pytestmark = pytest.mark.gpu
GRIDS_320 = [(20, 20), (10, 10), (5, 5), (3, 3), (2, 2), (1, 1)]


def _match(det, want_labels, want_scores, want_boxes, score_tol=3e-6, box_tol=1e-3, min_frac=0.99):
    """Detections must agree with the reference up to exp()-ulp effects: same count; position by position same label
    and score within score_tol (ex2.approx: <= 1e-7 absolute on a score, plus the reduction order of the row sum) for
    >= min_frac of the rows -- a last-bit score difference can swap two neighbours of a near-tie, or move a candidate
    across the threshold / top-k cut and shift the tail -- and boxes within box_tol = 1e-3 px on the matching rows
    (expf vs the host libm: 2 ulp of a <= 320 px coordinate is 8e-5 px)."""
    labels, scores, boxes = det["labels"].cpu().numpy(), det["scores"].cpu().numpy(), det["boxes"].cpu().numpy()
    assert labels.shape == want_labels.shape, (labels.shape, want_labels.shape)
    same = (labels == want_labels) & (np.abs(scores - want_scores) <= score_tol)
    assert same.mean() >= min_frac, same.mean()
    assert np.abs(boxes[same] - want_boxes[same]).max() <= box_tol
    assert np.all(scores[:-1] >= scores[1:])


def test_stress_golden(golden_dir):
    """Config 4 shape: 3234 priors x 91 classes, thr 0.001, top-k 400, NMS 0.55, D 300."""
    g = np.load(os.path.join(golden_dir, "postprocess_stress.npz"))
    B = int(g["batch"])
    gen = torch.Generator().manual_seed(int(g["seed"]))
    logits = torch.randn(B, 3234, 91, generator=gen) * 4.0
    bbox = torch.randn(B, 3234, 4, generator=gen) * 1.5
    anchors = torch.from_numpy(boxes_np.default_boxes(GRIDS_320, (320, 320)))
    dets = ops.postprocess_detections({"cls_logits": logits.cuda(), "bbox_regression": bbox.cuda()}, anchors.cuda(),
                                      (320, 320), score_thresh=0.001, nms_thresh=0.55, detections_per_img=300,
                                      topk_candidates=400)
    for i in range(B):
        _match(dets[i], g["det_labels"][i], g["det_scores"][i], g["det_boxes"][i])


@pytest.mark.parametrize("B,P,K,topk,D,thr", [(3, 3234, 91, 300, 300, 0.001), (2, 3000, 21, 400, 100, 0.01),
                                              (1, 8190, 21, 400, 200, 0.02), (5, 37, 3, 10, 7, 0.2), (2, 3234, 91, 300, 300, 0.9999),
                                              # > 2048 kept entries per image: the serial k-way merge instead of the sorted one
                                              (1, 8190, 21, 1000, 3000, 0.001)])
def test_vs_oracle(B, P, K, topk, D, thr):
    gen = torch.Generator().manual_seed(P + K)
    logits = torch.randn(B, P, K, generator=gen) * 3.0
    bbox = torch.randn(B, P, 4, generator=gen)
    a = torch.rand(P, 2, generator=gen) * 300
    anchors = torch.cat([a, a + torch.rand(P, 2, generator=gen) * 80 + 2], 1)
    dets = ops.postprocess_detections({"cls_logits": logits.cuda(), "bbox_regression": bbox.cuda()}, anchors.cuda(),
                                      (320, 320), score_thresh=thr, nms_thresh=0.5, detections_per_img=D,
                                      topk_candidates=topk)
    scores = torch.softmax(logits, -1).numpy()
    for i in range(B):
        o = boxes_np.postprocess_detections(None, bbox[i].numpy(), anchors.numpy(), (320, 320), score_thresh=thr,
                                            nms_thresh=0.5, detections_per_img=D, topk_candidates=topk, scores=scores[i])
        if o["labels"].shape[0] == 0:
            assert dets[i]["labels"].numel() == 0
            continue
        if D == 3000:
            assert o["labels"].shape[0] > 2048
        _match(dets[i], o["labels"], o["scores"], o["boxes"], min_frac=0.97)


def test_legacy_flavour_vs_oracle():
    """box_head.PostProcess semantics: no per-class top-k, remove_small_boxes(1e-2)."""
    gen = torch.Generator().manual_seed(11)
    B, P, K = 2, 3000, 21
    logits = torch.randn(B, P, K, generator=gen) * 3.0
    bbox = torch.randn(B, P, 4, generator=gen)
    bbox[:, ::7, 2:] = -60.0                       # degenerate (tiny) boxes that remove_small_boxes drops
    priors = torch.from_numpy(boxes_np.default_boxes([(19, 19), (10, 10), (5, 5), (3, 3), (2, 2), (1, 1)], (300, 300)))
    pp = ops.PostProcess((0.1, 0.2), 0.05, 0.45, 100)
    dets = pp(logits.cuda(), bbox.cuda(), priors.cuda(), [(300, 300)] * B)
    scores = torch.softmax(logits, -1).numpy()
    for i in range(B):
        o = boxes_np.legacy_postprocess(None, bbox[i].numpy(), priors.numpy(), (300, 300), score_thresh=0.05, scores=scores[i])
        _match(dets[i], o["labels"], o["scores"], o["boxes"], min_frac=0.97)
        w = dets[i]["boxes"][:, 2] - dets[i]["boxes"][:, 0]
        assert bool((w >= 1e-2).all())


def test_padded_outputs_and_counts():
    gen = torch.Generator().manual_seed(2)
    logits = torch.randn(4, 500, 6, generator=gen).cuda()
    logits[1] = -20.0
    logits[1, :, 0] = 20.0                          # image 1: everything is background -> no detections
    bbox = torch.randn(4, 500, 4, generator=gen).cuda()
    anchors = torch.tensor([[10., 10., 60., 60.]]).repeat(500, 1).cuda()
    boxes, scores, labels, counts = ops.postprocess_padded(logits, bbox, anchors, (100, 100), 0.05, 0.5, 50, 20)
    c = counts.tolist()
    assert c[1] == 0 and all(0 <= v <= 50 for v in c)
    for i in range(4):
        assert float(scores[i, c[i]:].abs().sum()) == 0.0 and int(labels[i, c[i]:].abs().sum()) == 0
        assert bool((labels[i, :c[i]] >= 1).all())
        assert float(boxes[i].min()) >= 0.0 and float(boxes[i].max()) <= 100.0       # clip_boxes_to_image


def test_full_size_properties():
    """BASELINE config 4 at full batch (1024 x 3234 x 91): size-independent invariants."""
    B = 1024
    gen = torch.Generator(device="cuda").manual_seed(7)
    logits = torch.randn(B, 3234, 91, generator=gen, device="cuda") * 4.0
    bbox = torch.randn(B, 3234, 4, generator=gen, device="cuda") * 1.5
    anchors = torch.from_numpy(boxes_np.default_boxes(GRIDS_320, (320, 320))).cuda()
    boxes, scores, labels, counts = ops.postprocess_padded(logits, bbox, anchors, (320, 320), 0.001, 0.55, 300, 400)
    assert int(counts.min()) == 300 and int(counts.max()) == 300
    assert bool((scores[:, :-1] >= scores[:, 1:]).all())
    assert int(labels.min()) >= 1 and int(labels.max()) <= 90
    # NMS invariant on a sample of images: within a class no two kept boxes overlap above the threshold,
    # checked with the C oracle (re-running NMS on the kept set keeps everything)
    for i in range(0, B, 97):
        keep = nms_c.batched_nms(boxes[i].cpu().numpy(), scores[i].cpu().numpy(), labels[i].cpu().numpy(), 0.55)
        assert keep.shape[0] == 300
    # permutation invariance over images: image b processed alone gives the same rows
    b1, s1, l1, c1 = ops.postprocess_padded(logits[5:6], bbox[5:6], anchors, (320, 320), 0.001, 0.55, 300, 400)
    assert torch.equal(s1[0], scores[5]) and torch.equal(l1[0], labels[5]) and torch.equal(b1[0], boxes[5])
