"""Bit-exact NMS on the GPU: kept indices vs the reference's own results (golden vectors produced by
torchvision's CPU kernel through tests/golden/make_golden.py) and vs the C oracle on fresh inputs."""
import os

import numpy as np
import pytest
import torch

from demonet_b200 import ops
from oracle import nms_c

pytestmark = pytest.mark.gpu


def test_golden_nms_cases_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "nms_cases.npz"))
    for c in range(int(g["n_cases"])):
        boxes = torch.from_numpy(g["c%d_boxes" % c]).cuda()
        scores = torch.from_numpy(g["c%d_scores" % c]).cuda()
        thr = float(g["c%d_thr" % c])
        idxs = g["c%d_idxs" % c]
        if idxs.shape[0] == 0:
            keep = ops.nms(boxes, scores, thr)
        else:
            keep = ops.batched_nms(boxes, scores, torch.from_numpy(idxs).cuda(), thr)
        assert keep.dtype == torch.int64
        assert np.array_equal(keep.cpu().numpy(), g["c%d_keep" % c]), "case %d (n=%d, thr=%g)" % (c, boxes.shape[0], thr)


def test_empty_and_single():
    z = ops.batched_nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda(), torch.zeros(0, dtype=torch.int64).cuda(), 0.5)
    assert z.shape == (0,) and z.dtype == torch.int64
    one = ops.nms(torch.tensor([[0., 0., 1., 1.]]).cuda(), torch.tensor([0.3]).cuda(), 0.5)
    assert one.tolist() == [0]


def test_errors():
    b = torch.rand(8, 4).cuda()
    with pytest.raises(ValueError):
        ops.batched_nms(b, torch.rand(8).cuda(), torch.full((8,), 5000, dtype=torch.int64).cuda(), 0.5)
    with pytest.raises(ValueError):
        ops.batched_nms(torch.rand(8, 5).cuda(), torch.rand(8).cuda(), torch.zeros(8, dtype=torch.int64).cuda(), 0.5)


@pytest.mark.parametrize("n,ncls,thr", [(5000, 1, 0.55), (20000, 33, 0.45), (36000, 90, 0.55), (70000, 200, 0.5)])
def test_random_vs_c_oracle(n, ncls, thr):
    g = torch.Generator().manual_seed(n + ncls)
    ctr = torch.rand(max(1, n // 10), 2, generator=g) * 320
    c = ctr[torch.randint(0, ctr.shape[0], (n,), generator=g)] + torch.randn(n, 2, generator=g) * 5
    wh = (torch.rand(n, 2, generator=g) * 0.5 + 0.25) * 70
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 320)
    scores = torch.rand(n, generator=g)                      # includes tied scores at these sizes
    idxs = torch.randint(0, ncls, (n,), generator=g)
    want = nms_c.batched_nms(boxes.numpy(), scores.numpy(), idxs.numpy(), thr)
    got = ops.batched_nms(boxes.cuda(), scores.cuda(), idxs.cuda(), thr).cpu().numpy()
    assert np.array_equal(got, want)


def test_idempotence_and_sortedness():
    """Size-independent properties: NMS of the kept set keeps everything; output is score-sorted."""
    g = torch.Generator().manual_seed(5)
    n = 30000
    boxes = torch.rand(n, 2, generator=g) * 300
    boxes = torch.cat([boxes, boxes + torch.rand(n, 2, generator=g) * 60 + 1], 1).cuda()
    scores = torch.rand(n, generator=g).cuda()
    idxs = torch.randint(0, 90, (n,), generator=g).cuda()
    keep = ops.batched_nms(boxes, scores, idxs, 0.55)
    s = scores[keep]
    assert bool((s[:-1] >= s[1:]).all())
    again = ops.batched_nms(boxes[keep], scores[keep], idxs[keep], 0.55)
    assert again.numel() == keep.numel()
