"""Parity metrics in the north-star's units (TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke()
and nothing under demonet_b200/).

The north-star states the floating-point tolerance as "max abs 1e-2 on decoded box coordinates and 1e-3 on scores";
`head_metrics` measures exactly those quantities between an engine result and the fp32 oracle / reference
(generalized_ssd.py:354 softmax, :362-363 decode + clip), plus the relative rms of the raw head outputs, and
`detection_match` the fraction of the reference's best detections that the engine reproduces.
"""
import numpy as np
import torch

from . import boxes_np


def head_metrics(cls, reg, ref_cls, ref_reg, anchors, size):
    """cls / ref_cls f32 [B,P,K], reg / ref_reg f32 [B,P,4] (CPU tensors), anchors f32 [P,4] numpy.  Returns a dict of
    floats: logits_rel_rms, logit_max_abs, bbox_rel_rms, score_max_abs, score_mean_abs, score_p999_abs (99.9th
    percentile), box_max_abs_px, box_mean_abs_px, box_p999_abs_px."""
    cls, reg, ref_cls, ref_reg = (t.detach().float().cpu() for t in (cls, reg, ref_cls, ref_reg))

    def rel_rms(a, b):
        return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt())
    sc, rsc = torch.softmax(cls, -1), torch.softmax(ref_cls, -1)
    ds = (sc - rsc).abs()
    dec = np.stack([boxes_np.clip_boxes_to_image(boxes_np.decode_single(r.numpy(), anchors), size) for r in reg])
    rdec = np.stack([boxes_np.clip_boxes_to_image(boxes_np.decode_single(r.numpy(), anchors), size) for r in ref_reg])
    db = np.abs(dec - rdec)
    return {"logits_rel_rms": rel_rms(cls, ref_cls), "logit_max_abs": float((cls - ref_cls).abs().max()),
            "bbox_rel_rms": rel_rms(reg, ref_reg),
            "score_max_abs": float(ds.max()), "score_mean_abs": float(ds.mean()),
            "score_p999_abs": float(torch.quantile(ds.flatten()[:: max(1, ds.numel() // 4_000_000)], 0.999)),
            "box_max_abs_px": float(db.max()), "box_mean_abs_px": float(db.mean()),
            "box_p999_abs_px": float(np.quantile(db, 0.999))}


def detection_match(dets, ref_dets, top=100, iou_thr=0.5):
    """Mean over images of the fraction of the reference's `top` best detections matched by a detection of the same
    label with IoU >= iou_thr.  dets / ref_dets: lists of dicts with boxes [n,4], labels [n] (tensors or arrays)."""
    from torchvision.ops import box_iou
    def host(v):
        return torch.as_tensor(np.asarray(v.detach().cpu() if hasattr(v, "detach") else v))
    fr = []
    for d, r in zip(dets, ref_dets):
        rb, rl = host(r["boxes"])[:top].float(), host(r["labels"])[:top]
        if rb.shape[0] == 0:
            continue
        db, dl = host(d["boxes"]).float(), host(d["labels"])
        if db.shape[0] == 0:
            fr.append(0.0)
            continue
        ok = (box_iou(rb, db) >= iou_thr) & (rl[:, None] == dl[None, :])
        fr.append(float(ok.any(1).float().mean()))
    return float(np.mean(fr)) if fr else 1.0


def reference_detections(ref_cls, ref_reg, anchors, size, **post):
    """The oracle's SSD.postprocess_detections (NumPy restatement) on reference head outputs, per image."""
    sc = torch.softmax(ref_cls.float(), -1).numpy()
    return [boxes_np.postprocess_detections(None, ref_reg[i].numpy(), anchors, size, scores=sc[i], **post)
            for i in range(ref_cls.shape[0])]


def format_metrics(m):
    return ", ".join("%s %.3g" % (k, v) for k, v in m.items())
