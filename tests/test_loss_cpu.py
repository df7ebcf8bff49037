"""Training-side oracle (oracle/loss_np.py) against the golden fixture the unmodified reference produced
(tests/golden/make_golden.py:gen_loss -> tests/golden/ssd_loss.npz).  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import loss_np


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()


@pytest.mark.parametrize("name", sorted(loss_np.LOSS_CASES))
def test_matching_and_loss_against_the_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "ssd_loss.npz"))
    anchors, targets, cls, reg = loss_np.seeded_case(name)
    assert str(g[name + "_inputs_sha256"]) == _sha(cls) + _sha(reg) + _sha(anchors), "seeded inputs differ from the fixture's"
    matched = np.stack([loss_np.match_image(b, anchors, 0.5) for b, _ in targets])
    assert np.array_equal(matched, g[name + "_matched"])                     # SSDMatcher indices, bit for bit
    assert (matched[1] == -1).all()                                         # the image without boxes
    out = loss_np.compute_loss(targets, cls, reg, anchors, matched, float(g["neg_to_pos_ratio"]))
    ref = g[name + "_losses"]
    assert abs(out["bbox_regression"] - ref[0]) <= 2e-6 * abs(ref[0])
    assert abs(out["classification"] - ref[1]) <= 2e-6 * abs(ref[1])


def test_matcher_semantics():
    # first maximum over the ground truth; the forced match overrides the threshold; the later ground-truth box wins a shared box
    q = np.array([[0.6, 0.2, 0.1, 0.3], [0.6, 0.7, 0.1, 0.3], [0.1, 0.1, 0.1, 0.3]], np.float32)
    assert loss_np.matcher(q, 0.5).tolist() == [0, 1, -1, -1]
    assert loss_np.ssd_matcher(q, 0.5).tolist() == [0, 1, -1, 2]
    q2 = np.array([[0.9, 0.1], [0.9, 0.1]], np.float32)                     # both boxes claim prediction 0
    assert loss_np.ssd_matcher(q2, 0.5).tolist() == [1, -1]
    with pytest.raises(ValueError, match="No ground-truth boxes"):
        loss_np.matcher(np.zeros((0, 5), np.float32), 0.5)


def test_hard_negative_count():
    # 2 foreground boxes -> ceil(3 * 2) = 6 background boxes, the largest losses among the negatives
    rng = np.random.default_rng(0)
    P, K = 40, 5
    cls = rng.standard_normal((1, P, K)).astype(np.float32)
    reg = np.zeros((1, P, 4), np.float32)
    anchors = np.tile(np.array([[0, 0, 10, 10]], np.float32), (P, 1))
    matched = np.full((1, P), -1, np.int64)
    matched[0, [3, 17]] = 0
    targets = [(np.array([[0, 0, 10, 10]], np.float32), np.array([2], np.int64))]
    _, d = loss_np.compute_loss(targets, cls, reg, anchors, matched, 3.0, return_details=True)
    assert d["foreground"].sum() == 2 and d["background"].sum() == 6 and not (d["background"] & d["foreground"]).any()
    neg = np.where(d["foreground"][0], -np.inf, d["ce"][0])
    assert set(np.argsort(-neg)[:6]) == set(np.nonzero(d["background"][0])[0])
