"""Serialised engine files for callers outside Python.

The reference's deployment story is TorchScript: `trace_model.py` saves `ssd_lite_mobilenet_v2.pt` and a libtorch program
loads and runs it (test/tracing/trace_model.py:1-14, test/tracing/test_demonet_tracing.cpp:31-58).  The counterpart here is
a flat file with everything `dn_engine_create` + `dn_engine_load_weights` need -- the dn_model_desc, the dn_op / dn_buf
arrays, the default-box table and the folded weight blob -- which a C or C++ program reads straight into the structs of
include/demonet_b200.h and hands to the library it dlopen()ed (examples/cpp/dn_cpp_smoke.cpp).

Layout (little endian, no padding between sections):
    char[8]  magic "DNENGv1\\0"
    int32    abi_version, act_dtype (1 = fp16, 0 = bf16), sizeof(dn_model_desc), sizeof(dn_op), sizeof(dn_buf), n_priors
    int64    weight_blob_bytes
    dn_model_desc (pointer fields zero)
    dn_op[n_ops], dn_buf[n_bufs], float[n_priors * 4] anchors, uint8[weight_blob_bytes]
"""
import ctypes
import struct

import torch

from . import _C, plan as _plan
from .module import SSDLiteB200, _Engine

MAGIC = b"DNENGv1\0"


def save_engine_file(model: SSDLiteB200, path: str) -> dict:
    """Write the serialised engine of `model` (its current weights, thresholds and activation storage type)."""
    kw = dict(image_mean=model.image_mean, image_std=model.image_std, score_thresh=model.score_thresh,
              nms_thresh=model.nms_thresh, topk_candidates=model.topk_candidates,
              detections_per_img=model.detections_per_img, gemm_impl=0, use_cuda_graph=1, keep_activations=False,
              pipeline_slots=0, act_dtype=model.act_dtype,
              min_box_size=1e-2 if model.postprocess_flavour == "legacy" else -1.0)
    eng = _Engine(model.plan, kw, 1, torch.device("cpu"))           # never created: only describes
    blob, offsets = _plan.pack_weights(model.plan, model.state_dict(), model.act_dtype)
    d = eng.describe(offsets)
    n_ops, n_bufs, P = d.n_ops, d.n_bufs, model.plan.num_priors
    ops_bytes = bytes(eng._ops)
    bufs_bytes = bytes(eng._bufs_c)
    d.ops_host = ctypes.POINTER(_C.Op)()
    d.bufs_host = ctypes.POINTER(_C.Buf)()
    d.anchors_host = ctypes.POINTER(ctypes.c_float)()
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<6i", _C.ABI_VERSION, 1 if model.act_dtype == "fp16" else 0, ctypes.sizeof(_C.ModelDesc),
                            ctypes.sizeof(_C.Op), ctypes.sizeof(_C.Buf), P))
        f.write(struct.pack("<q", len(blob)))
        f.write(bytes(d))
        f.write(ops_bytes)
        f.write(bufs_bytes)
        f.write(eng.anchors.astype("<f4").tobytes())
        f.write(blob)
    return {"n_ops": n_ops, "n_bufs": n_bufs, "n_priors": P, "weight_bytes": len(blob), "act_dtype": model.act_dtype}
