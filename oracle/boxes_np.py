"""CPU restatement (NumPy, fp32) of the box arithmetic on demonet's SSDLite hot path.

TEST INFRASTRUCTURE ONLY -- this is the parity oracle, not the product. Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it. The product path (demonet_b200/) never does.

Parity status: PINNED. Every function here is checked against the unmodified
reference imported from /root/reference (oracle/refshim.py) by
tests/golden/make_golden.py, and against the committed outputs of that run
(tests/golden/*.npz) by tests/test_oracle_cpu.py. The reference itself ships no
golden vectors for this path.

Each function cites the reference lines it restates (paths relative to
/root/reference). torchvision.ops.{nms,batched_nms,clip_boxes_to_image,
remove_small_boxes} are third-party (torchvision 0.26.0, the version installed
in this image; the reference pins none -- requirements.txt:5-6 commented out);
their published algorithm is restated from the reference's call sites
generalized_ssd.py:363,389 and box_head.py:349,370,374.
"""
import math

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------
# Default boxes ("PriorBox")            demonet/models/anchor_utils.py:10-126
# --------------------------------------------------------------------------
def default_box_scales(num_outputs, min_ratio=0.15, max_ratio=0.9):
    """anchor_utils.py:38-46 (python-float arithmetic, as in the reference)."""
    if num_outputs > 1:
        range_ratio = max_ratio - min_ratio
        scales = [min_ratio + range_ratio * k / (num_outputs - 1.0) for k in range(num_outputs)]
        scales.append(1.0)
    else:
        scales = [min_ratio, max_ratio]
    return scales


def default_box_wh_pairs(aspect_ratios, scales):
    """anchor_utils.py:51-68: [s_k,s_k], [s'_k,s'_k], then (w,h),(h,w) per ratio; fp32 table."""
    out = []
    for k in range(len(aspect_ratios)):
        s_k = scales[k]
        s_prime_k = math.sqrt(scales[k] * scales[k + 1])
        wh = [[s_k, s_k], [s_prime_k, s_prime_k]]
        for ar in aspect_ratios[k]:
            sq_ar = math.sqrt(ar)
            w = scales[k] * sq_ar
            h = scales[k] / sq_ar
            wh.extend([[w, h], [h, w]])
        out.append(np.asarray(wh, dtype=F32))
    return out


def default_boxes(grid_sizes, image_size, aspect_ratios=None, min_ratio=0.2, max_ratio=0.95,
                  scales=None, clip=True, steps=None):
    """anchor_utils.py:75-100 + :111-126.  Returns f32[P,4] xyxy in pixels.

    grid_sizes: [(H_k, W_k)], image_size: (H, W).  Order: level, then cell
    (y-major, x fastest), then anchor (anchor_utils.py:88-94).
    """
    if aspect_ratios is None:
        aspect_ratios = [[2, 3] for _ in grid_sizes]
    if scales is None:
        scales = default_box_scales(len(aspect_ratios), min_ratio, max_ratio)
    wh_pairs = default_box_wh_pairs(aspect_ratios, scales)
    rows = []
    for k, (fh, fw) in enumerate(grid_sizes):
        # ((arange + 0.5) / f_k).to(float32): torch promotes int64 + 0.5 to fp32 (:85-86)
        # (:79-83) with `steps` the centres are tiled by image_size / step instead of the grid size (ssd300_vgg16)
        x_f, y_f = (image_size[1] / steps[k], image_size[0] / steps[k]) if steps is not None else (fw, fh)
        sx = (np.arange(fw).astype(F32) + F32(0.5)) / F32(x_f)
        sy = (np.arange(fh).astype(F32) + F32(0.5)) / F32(y_f)
        yy, xx = np.meshgrid(sy, sx, indexing="ij")
        xx = xx.reshape(-1)
        yy = yy.reshape(-1)
        A = wh_pairs[k].shape[0]
        shifts = np.stack([xx, yy], axis=-1)            # [HW,2]
        shifts = np.repeat(shifts, A, axis=0)           # (:91) stack(...)*A -> reshape(-1,2)
        wh = np.clip(wh_pairs[k], F32(0), F32(1)) if clip else wh_pairs[k]   # (:93)
        wh = np.tile(wh, (fh * fw, 1))                  # (:94) repeat((HW),1)
        rows.append(np.concatenate([shifts, wh], axis=1).astype(F32))
    db = np.concatenate(rows, axis=0)
    # forward (:119-124): cxcywh -> xyxy, then scale x by W and y by H, all fp32
    half = F32(0.5)
    out = np.concatenate([db[:, :2] - half * db[:, 2:], db[:, :2] + half * db[:, 2:]], axis=-1).astype(F32)
    out[:, 0::2] *= F32(image_size[1])
    out[:, 1::2] *= F32(image_size[0])
    return out


# --------------------------------------------------------------------------
# Box decode + clip                      demonet/models/_utils.py:187-224
# --------------------------------------------------------------------------
BBOX_XFORM_CLIP = math.log(1000.0 / 16)      # _utils.py:135


def decode_single(rel_codes, boxes, weights=(10.0, 10.0, 5.0, 5.0), clip=BBOX_XFORM_CLIP):
    """_utils.py:187-224, fp32, same operation order (true division by the weights)."""
    rel = np.asarray(rel_codes, dtype=F32)
    b = np.asarray(boxes, dtype=F32)
    widths = b[:, 2] - b[:, 0]
    heights = b[:, 3] - b[:, 1]
    ctr_x = b[:, 0] + F32(0.5) * widths
    ctr_y = b[:, 1] + F32(0.5) * heights
    wx, wy, ww, wh = (F32(w) for w in weights)
    dx = rel[:, 0] / wx
    dy = rel[:, 1] / wy
    dw = rel[:, 2] / ww
    dh = rel[:, 3] / wh
    dw = np.minimum(dw, F32(clip))
    dh = np.minimum(dh, F32(clip))
    pcx = dx * widths + ctr_x
    pcy = dy * heights + ctr_y
    pw = np.exp(dw).astype(F32) * widths
    ph = np.exp(dh).astype(F32) * heights
    half = F32(0.5)
    return np.stack([pcx - half * pw, pcy - half * ph, pcx + half * pw, pcy + half * ph], axis=1).astype(F32)


def clip_boxes_to_image(boxes, size):
    """torchvision.ops.boxes.clip_boxes_to_image (called generalized_ssd.py:363, box_head.py:349)."""
    h, w = size
    out = np.array(boxes, dtype=F32, copy=True)
    out[:, 0::2] = np.clip(out[:, 0::2], F32(0), F32(w))
    out[:, 1::2] = np.clip(out[:, 1::2], F32(0), F32(h))
    return out


def softmax_rows(logits):
    """F.softmax(x, -1) (generalized_ssd.py:354). Tolerance-only: exp differs by ulps across libms."""
    x = np.asarray(logits, dtype=F32)
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m).astype(F32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)


# --------------------------------------------------------------------------
# NMS                      torchvision::nms CPU kernel (called via batched_nms)
# --------------------------------------------------------------------------
def nms(boxes, scores, iou_threshold):
    """Greedy NMS, semantics pinned in SURVEY.md section 8(a) row N1.

    fp32 areas and IoU with no fused multiply-add, `inter / ((a_i + a_j) - inter)`,
    suppress iff double(iou) > iou_threshold (strict), order = stable descending
    score, NaN IoU never suppresses.  Returns int64 indices in kept order.
    """
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=F32).reshape(-1)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-scores, kind="stable") if not np.isnan(scores).any() else _stable_desc(scores)
    x1, y1, x2, y2 = (boxes[order, i] for i in range(4))
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, dtype=bool)
    thr = float(iou_threshold)
    zero = F32(0)
    keep = []
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(n):
            if suppressed[i]:
                continue
            keep.append(order[i])
            if i + 1 == n:
                break
            xx1 = np.maximum(x1[i], x1[i + 1:])
            yy1 = np.maximum(y1[i], y1[i + 1:])
            xx2 = np.minimum(x2[i], x2[i + 1:])
            yy2 = np.minimum(y2[i], y2[i + 1:])
            w = np.maximum(zero, xx2 - xx1)
            h = np.maximum(zero, yy2 - yy1)
            inter = w * h
            ovr = inter / ((areas[i] + areas[i + 1:]) - inter)
            suppressed[i + 1:] |= ovr.astype(np.float64) > thr
    return np.asarray(keep, dtype=np.int64)


def _stable_desc(scores):
    return np.asarray(sorted(range(len(scores)), key=lambda i: -scores[i]), dtype=np.int64)


def batched_nms_vanilla(boxes, scores, idxs, iou_threshold):
    """torchvision.ops.boxes._batched_nms_vanilla: nms() per class on raw coordinates,
    then all kept indices ordered by descending score.  torchvision's final sort is
    non-stable; here ties are broken by the lower candidate index (the order the CUDA
    path also produces), so results are comparable wherever scores are distinct.
    """
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=F32).reshape(-1)
    idxs = np.asarray(idxs).reshape(-1)
    keep_mask = np.zeros(scores.shape[0], dtype=bool)
    for c in np.unique(idxs):
        cur = np.nonzero(idxs == c)[0]
        k = nms(boxes[cur], scores[cur], iou_threshold)
        keep_mask[cur[k]] = True
    kept = np.nonzero(keep_mask)[0]
    return kept[np.argsort(-scores[kept], kind="stable")].astype(np.int64)


def batched_nms_coordinate_trick(boxes, scores, idxs, iou_threshold):
    """torchvision.ops.boxes._batched_nms_coordinate_trick (used by the reference when
    boxes.numel() <= 4000 on CPU): offsets boxes by idx*(max+1) in fp32, one nms() call."""
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    if boxes.size == 0:
        return np.zeros((0,), dtype=np.int64)
    mx = boxes.max()
    offsets = np.asarray(idxs).astype(F32) * (mx + F32(1))
    return nms(boxes + offsets[:, None], scores, iou_threshold)


# --------------------------------------------------------------------------
# SSD.postprocess_detections        demonet/models/generalized_ssd.py:351-397
# --------------------------------------------------------------------------
def select_candidates(scores, score_thresh, topk):
    """generalized_ssd.py:368-382 for one image.  scores f32[P,K] (softmax output).

    Per class 1..K-1: `score > score_thresh` compared in fp32, top-k by score
    (descending; ties -> lower anchor index, which torch.topk leaves unspecified).
    Returns (anchor_idx i64[n], score f32[n], label i64[n]) concatenated over classes.
    """
    P, K = scores.shape
    thr = F32(score_thresh)
    a_all, s_all, l_all = [], [], []
    for label in range(1, K):
        s = scores[:, label]
        idx = np.nonzero(s > thr)[0]
        order = np.argsort(-s[idx], kind="stable")[:topk]
        a_all.append(idx[order])
        s_all.append(s[idx[order]])
        l_all.append(np.full(order.shape[0], label, dtype=np.int64))
    return (np.concatenate(a_all).astype(np.int64), np.concatenate(s_all).astype(F32),
            np.concatenate(l_all))


def postprocess_detections(cls_logits, bbox_regression, anchors, image_shape, score_thresh=0.001,
                           nms_thresh=0.55, detections_per_img=300, topk_candidates=300, scores=None):
    """generalized_ssd.py:351-397 for ONE image; returns dict(boxes, scores, labels, anchor_idx).

    NMS strategy is always the per-class "vanilla" one (hard parts):
    that is what torchvision picks above 1000 candidates on CPU and what a
    per-(image,class) kernel reproduces exactly.
    """
    if scores is None:
        scores = softmax_rows(cls_logits)
    boxes = clip_boxes_to_image(decode_single(bbox_regression, anchors), image_shape)
    a, s, l = select_candidates(scores, score_thresh, topk_candidates)
    cand_boxes = boxes[a]
    keep = batched_nms_vanilla(cand_boxes, s, l, nms_thresh)[:detections_per_img]
    return {"boxes": cand_boxes[keep], "scores": s[keep], "labels": l[keep], "anchor_idx": a[keep],
            "keep": keep, "cand_boxes": cand_boxes, "cand_scores": s, "cand_labels": l}


# --------------------------------------------------------------------------
# Legacy PostProcess.forward (V2 flavour)   demonet/models/box_head.py:340-381
# --------------------------------------------------------------------------
def legacy_postprocess(pred_logits, pred_boxes, priors, image_shape, score_thresh=0.5,
                       nms_thresh=0.45, detections_per_img=100, min_size=1e-2, scores=None):
    """box_head.py:343-379 for ONE image: no per-class top-k; every (prior, class>0) pair is a
    candidate, flattened prior-major (box_head.py:350-365); score > thresh; remove_small_boxes
    (w >= min_size and h >= min_size, fp32); batched NMS; first detections_per_img."""
    if scores is None:
        scores = softmax_rows(pred_logits)
    P, K = scores.shape
    boxes = clip_boxes_to_image(decode_single(pred_boxes, priors), image_shape)
    flat_scores = scores[:, 1:].reshape(-1)
    flat_labels = np.tile(np.arange(1, K, dtype=np.int64), P)
    flat_prior = np.repeat(np.arange(P, dtype=np.int64), K - 1)
    inds = np.nonzero(flat_scores > F32(score_thresh))[0]
    b, s, l, a = boxes[flat_prior[inds]], flat_scores[inds], flat_labels[inds], flat_prior[inds]
    ws, hs = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    ok = np.nonzero((ws >= F32(min_size)) & (hs >= F32(min_size)))[0]
    b, s, l, a = b[ok], s[ok], l[ok], a[ok]
    keep = batched_nms_vanilla(b, s, l, nms_thresh)[:detections_per_img]
    return {"boxes": b[keep], "scores": s[keep], "labels": l[keep], "anchor_idx": a[keep],
            "keep": keep, "cand_boxes": b, "cand_scores": s, "cand_labels": l}
