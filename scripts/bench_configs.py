#!/usr/bin/env python
"""Measurements for the other BASELINE.json configurations (they are parity-test cases for bench.py's
contract, so their numbers are kept here, as JSON lines, for BASELINE.md section 4):

  --config 4   post-processing stress: 3234 priors x 91 classes, score_thresh 0.001, top-k 400, NMS 0.55,
               batch 1024 -> decode+NMS us/img and achieved GB/s on 1.229 MB/img algorithmic bytes
  --config 5   ssd_lite_mobilenet_v2, VOC 21 classes, 512x512, batch 512 -> img/s and the depthwise share
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import demonet_b200                                   # noqa: E402
from demonet_b200 import _C, plan as dplan            # noqa: E402
from demonet_b200.module import make_post_params      # noqa: E402
from demonet_b200 import seeded as weights            # noqa: E402


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def config4(args):
    B, P, K, D = args.batch or 1024, 3234, 91, 300
    g = torch.Generator(device="cuda").manual_seed(7)
    logits = torch.randn(B, P, K, generator=g, device="cuda") * 4.0
    bbox = torch.randn(B, P, 4, generator=g, device="cuda") * 1.5
    p = dplan.plan_ssdlite320_mobilenet_v3_large()
    anchors = torch.from_numpy(dplan.default_boxes(p)).cuda()
    prm = make_post_params(P, K, 320, 320, 0.001, 0.55, 400, D)
    lib = _C.lib()
    ws = torch.empty(lib.dn_postprocess_workspace_bytes(B, ctypes.byref(prm)), dtype=torch.uint8, device="cuda")
    boxes = torch.empty(B, D, 4, device="cuda")
    scores = torch.empty(B, D, device="cuda")
    labels = torch.empty(B, D, dtype=torch.int64, device="cuda")
    counts = torch.empty(B, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        _C.check(lib.dn_postprocess(logits.data_ptr(), bbox.data_ptr(), anchors.data_ptr(), B, ctypes.byref(prm),
                                    ws.data_ptr(), ws.numel(), boxes.data_ptr(), scores.data_ptr(), labels.data_ptr(),
                                    counts.data_ptr(), stream))
    ms = timed(step, args.steps, args.warmup)
    ms3 = (ctypes.c_float * 3)()
    _C.check(lib.dn_postprocess_profile(logits.data_ptr(), bbox.data_ptr(), anchors.data_ptr(), B, ctypes.byref(prm),
                                        ws.data_ptr(), ws.numel(), boxes.data_ptr(), scores.data_ptr(), labels.data_ptr(),
                                        counts.data_ptr(), 5, ms3, stream))
    alg = B * (P * K * 4 + P * 16) + P * 16
    return {"config": 4, "metric": "decode+NMS us/img", "value": ms * 1e3 / B, "unit": "us/img", "batch": B,
            "ms_per_batch": ms, "images_per_s": B / ms * 1e3, "algorithmic_bytes": alg, "achieved_GBps": alg / ms / 1e6,
            "phase_ms": {"softmax_decode_hist": ms3[0], "sort_nms_rounds": ms3[1], "merge": ms3[2]},
            "detections_per_img": int(counts.min()), "workload": "logits N(0,4^2) seed 7, bbox N(0,1.5^2), thr .001, "
            "top-k 400, NMS .55, D 300 (SURVEY 8(d) config 4)"}


def config5(args):
    B, S = args.batch or 512, 512
    model = demonet_b200.ssd_lite_mobilenet_v2(image_size=S, score_thresh=0.5, num_classes=21)
    model.load_state_dict(weights.seeded_state_dict(model.state_dict()))
    model = model.cuda()
    eng = model.reserve(B)
    imgs = weights.synthetic_images(B, S).cuda()
    io = model._io_buffers(torch.device("cuda", 0), B, False)
    ms = timed(lambda: eng.forward(imgs, io), args.steps, args.warmup)
    n = eng.launches_per_forward
    per = (ctypes.c_float * (len(model.plan.layers) + 3))()
    _C.check(_C.lib().dn_engine_profile(eng._handle, imgs.data_ptr(), B, 5, per, torch.cuda.current_stream().cuda_stream))
    kinds = {}
    bytes_dw = 0
    for L, t in zip(model.plan.layers, list(per)):
        kinds[L.kind] = kinds.get(L.kind, 0.0) + t
        if L.kind == "dw":
            bytes_dw += B * (L.h_in * L.w_in + L.h_out * L.w_out) * L.cin * 2
    post = sum(list(per)[len(model.plan.layers):])
    return {"config": 5, "metric": "ssd_lite_mobilenet_v2 512x512 images/sec", "value": B / ms * 1e3, "unit": "img/s",
            "batch": B, "ms_per_batch": ms, "launches": n, "kind_ms": {**{k: round(v, 4) for k, v in kinds.items()},
                                                                     "postprocess": round(post, 4)},
            "depthwise_GBps": bytes_dw / (kinds["dw"] * 1e-3) / 1e9, "depthwise_algorithmic_MB_per_img": bytes_dw / B / 1e6,
            "arena_GB": eng.device_bytes / 1e9,
            "workload": "V2 assembly, 21 classes, legacy PostProcess (thr 0.5, NMS 0.45, D 100), seeded weights"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[4, 5])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    print(json.dumps(config4(a) if a.config == 4 else config5(a)))
