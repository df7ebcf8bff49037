"""NumPy restatement of the reference's training-side box matching and multibox loss.

TEST INFRASTRUCTURE ONLY (the checker of tests/, never imported by the product path).

Follows, line by line:
  box_iou            torchvision.ops.boxes.box_iou as called at demonet/models/generalized_ssd.py:334
  matcher            Matcher.__call__       demonet/models/_utils.py:283-323 (low == high threshold)
  ssd_matcher        SSDMatcher.__call__    demonet/models/_utils.py:350-362
  encode_boxes       encode_boxes           demonet/models/_utils.py:83-127
  compute_loss       SSD.compute_loss       demonet/models/generalized_ssd.py:210-269
Pinned against the unmodified reference by tests/golden/make_golden.py:gen_loss (indices bit for bit, loss values to
fp32 summation order) and against the committed fixture tests/golden/ssd_loss.npz by tests/test_oracle_cpu.py.
"""
import numpy as np

f32 = np.float32


def box_iou(boxes1, boxes2):
    """[M,4] x [N,4] -> [M,N], fp32 operation order of torchvision's box_iou."""
    b1 = np.asarray(boxes1, f32)
    b2 = np.asarray(boxes2, f32)
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = np.maximum(b1[:, None, :2], b2[None, :, :2])
    rb = np.minimum(b1[:, None, 2:], b2[None, :, 2:])
    wh = np.clip(rb - lt, 0, None).astype(f32)
    inter = wh[..., 0] * wh[..., 1]
    union = (area1[:, None] + area2[None, :]) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / union).astype(f32)


def matcher(quality, threshold):
    """Matcher(threshold, threshold, allow_low_quality_matches=False).__call__ (_utils.py:283-323)."""
    q = np.asarray(quality, f32)
    if q.size == 0:
        if q.shape[0] == 0:
            raise ValueError("No ground-truth boxes available for one of the images during training")
        raise ValueError("No proposal boxes available for one of the images during training")
    matches = q.argmax(axis=0).astype(np.int64)          # first maximum, like torch.max on the CPU
    vals = q.max(axis=0)
    matches[vals < f32(threshold)] = -1                   # BELOW_LOW_THRESHOLD; BETWEEN_THRESHOLDS is empty
    return matches


def ssd_matcher(quality, threshold):
    """SSDMatcher.__call__ (_utils.py:350-362): every ground-truth box also claims its best default box."""
    q = np.asarray(quality, f32)
    matches = matcher(q, threshold)
    best = q.argmax(axis=1)
    for g in range(q.shape[0]):                          # index_put_ on the CPU: later rows overwrite earlier ones
        matches[best[g]] = g
    return matches


def match_image(gt_boxes, anchors, threshold=0.5):
    """generalized_ssd.py:326-335 for one image."""
    gt_boxes = np.asarray(gt_boxes, f32).reshape(-1, 4)
    if gt_boxes.size == 0:
        return np.full((len(anchors),), -1, np.int64)
    return ssd_matcher(box_iou(gt_boxes, anchors), threshold)


def encode_boxes(reference_boxes, proposals, weights=(10.0, 10.0, 5.0, 5.0)):
    """_utils.py:83-127, fp32."""
    r = np.asarray(reference_boxes, f32)
    p = np.asarray(proposals, f32)
    wx, wy, ww, wh = (f32(w) for w in weights)
    ex_w = p[:, 2] - p[:, 0]
    ex_h = p[:, 3] - p[:, 1]
    ex_cx = p[:, 0] + f32(0.5) * ex_w
    ex_cy = p[:, 1] + f32(0.5) * ex_h
    gt_w = r[:, 2] - r[:, 0]
    gt_h = r[:, 3] - r[:, 1]
    gt_cx = r[:, 0] + f32(0.5) * gt_w
    gt_cy = r[:, 1] + f32(0.5) * gt_h
    dx = wx * (gt_cx - ex_cx) / ex_w
    dy = wy * (gt_cy - ex_cy) / ex_h
    dw = ww * np.log(gt_w / ex_w)
    dh = wh * np.log(gt_h / ex_h)
    return np.stack([dx, dy, dw, dh], axis=1).astype(f32)


def smooth_l1(d):
    a = np.abs(d)
    return np.where(a < 1.0, 0.5 * d * d, a - 0.5)


def cross_entropy_rows(logits, targets):
    """F.cross_entropy(reduction='none') in float64 (the comparison tolerance absorbs fp32 rounding)."""
    x = np.asarray(logits, np.float64)
    m = x.max(axis=-1, keepdims=True)
    lse = m[..., 0] + np.log(np.exp(x - m).sum(axis=-1))
    return lse - np.take_along_axis(x, targets[..., None], axis=-1)[..., 0]


def compute_loss(targets, cls_logits, bbox_regression, anchors, matched_idxs, neg_to_pos_ratio=3.0,
                 weights=(10.0, 10.0, 5.0, 5.0), return_details=False):
    """SSD.compute_loss (generalized_ssd.py:210-269).  targets: list of (boxes [M,4], labels [M]) per image.
    Ranking ties at the hard-negative cut go to the lower index (stable descending order)."""
    B, P, K = cls_logits.shape
    num_foreground = 0
    bbox_loss = 0.0
    cls_targets = np.zeros((B, P), np.int64)
    for b in range(B):
        boxes, labels = targets[b]
        m = matched_idxs[b]
        fg = np.nonzero(m >= 0)[0]
        num_foreground += fg.size
        if fg.size:
            t = encode_boxes(np.asarray(boxes, f32)[m[fg]], anchors[fg], weights)
            bbox_loss += smooth_l1(bbox_regression[b][fg].astype(np.float64) - t.astype(np.float64)).sum()
            cls_targets[b, fg] = np.asarray(labels, np.int64)[m[fg]]
    ce = cross_entropy_rows(cls_logits, cls_targets)
    foreground = cls_targets > 0
    num_negative = f32(neg_to_pos_ratio) * foreground.sum(1, keepdims=True).astype(f32)
    negative_loss = ce.astype(f32).copy()
    negative_loss[foreground] = -np.inf
    order = np.argsort(-negative_loss, axis=1, kind="stable")
    rank = np.argsort(order, axis=1, kind="stable")
    background = rank.astype(f32) < num_negative
    N = max(1, num_foreground)
    out = {"bbox_regression": bbox_loss / N, "classification": (ce[foreground].sum() + ce[background].sum()) / N}
    if return_details:
        return out, {"cls_targets": cls_targets, "foreground": foreground, "background": background, "N": N, "ce": ce}
    return out


# ---- seeded cases shared by tests/golden/make_golden.py:gen_loss (reference run) and the tests -----------------------
V3_GRIDS = [(20, 20), (10, 10), (5, 5), (3, 3), (2, 2), (1, 1)]                       # ssdlite320: P = 3234
VGG_GRIDS = [(38, 38), (19, 19), (10, 10), (5, 5), (3, 3), (1, 1)]                    # ssd300_vgg16: P = 8732
LOSS_CASES = {"v3": dict(seed=11, B=4, K=91, size=320), "vgg": dict(seed=12, B=3, K=21, size=300)}


def case_anchors(name):
    from . import boxes_np
    if name == "v3":
        return boxes_np.default_boxes(V3_GRIDS, (320, 320))
    ar = [[2], [2, 3], [2, 3], [2, 3], [2], [2]]
    return boxes_np.default_boxes(VGG_GRIDS, (300, 300), ar, scales=[0.07, 0.15, 0.33, 0.51, 0.69, 0.87, 1.05],
                                  steps=[8, 16, 32, 64, 100, 300])


def seeded_case(name):
    """Inputs of one loss case: anchors [P,4], targets [(boxes, labels)] * B, cls_logits [B,P,K], bbox_regression [B,P,4].
    Image 1 has no boxes; the others mix random boxes, exact copies of default boxes (IoU 1), duplicated ground-truth
    boxes (arg-max ties over the ground truth) and one tiny box that only the forced match can claim."""
    c = LOSS_CASES[name]
    rng = np.random.default_rng(c["seed"])
    anchors = case_anchors(name)
    P, S = len(anchors), float(c["size"])
    targets = []
    for b in range(c["B"]):
        m = 0 if b == 1 else int(rng.integers(2, 10))
        boxes = []
        for j in range(m):
            kind = j % 4
            if kind == 0 and boxes:
                boxes.append(boxes[int(rng.integers(0, len(boxes)))].copy())            # duplicate of an earlier box
            elif kind == 1:
                a = anchors[int(rng.integers(0, P))]
                boxes.append(np.clip(a, 0.0, S).astype(f32))                              # a default box itself
            elif kind == 2 and j == 2:
                cx, cy = rng.uniform(20, S - 20, 2)
                boxes.append(np.array([cx, cy, cx + 2.5, cy + 3.5], f32))                 # tiny
            else:
                cx, cy = rng.uniform(0.15 * S, 0.85 * S, 2)
                w, h = rng.uniform(0.05 * S, 0.6 * S, 2)
                boxes.append(np.clip(np.array([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2]), 0.0, S).astype(f32))
        boxes = np.stack(boxes).astype(f32) if boxes else np.zeros((0, 4), f32)
        keep = (boxes[:, 2] > boxes[:, 0]) & (boxes[:, 3] > boxes[:, 1])
        boxes = boxes[keep]
        labels = rng.integers(1, c["K"], len(boxes)).astype(np.int64)
        targets.append((boxes, labels))
    cls_logits = (rng.standard_normal((c["B"], P, c["K"])) * 2.0).astype(f32)
    bbox_regression = (rng.standard_normal((c["B"], P, 4)) * 1.5).astype(f32)
    return anchors, targets, cls_logits, bbox_regression
