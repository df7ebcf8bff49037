"""ctypes binding of libdemonet_b200.so (the C ABI declared in include/demonet_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# One library per activation storage type (csrc/build.sh).  fp16 is the default: same bytes and tensor-core rate as
# bf16, 6.6x closer to the fp32 reference end to end (DESIGN.md "Numerics"); DN_ACT_DTYPE=bf16 or the `act_dtype`
# argument of the model builders selects the bfloat16 build.
ACT_DTYPES = ("fp16", "bf16")
DEFAULT_ACT_DTYPE = os.environ.get("DN_ACT_DTYPE", "fp16")
# DN_LIB_DIR: a variant build of the library (csrc/build.sh with DN_LIB_OUT / DN_EXTRA_FLAGS), for A/B measurements
_LIB_DIR = os.environ.get("DN_LIB_DIR") or os.path.join(_HERE, "lib")
LIB_PATHS = {dt: os.path.join(_LIB_DIR, "libdemonet_b200_%s.so" % dt) for dt in ACT_DTYPES}
LIB_PATH = LIB_PATHS["fp16"]

ABI_VERSION = 4          # include/demonet_b200.h DN_ABI_VERSION
DN_OK = 0
DN_ERR_INVALID = -1
DN_ERR_CUDA = -2
DN_ERR_UNSUPPORTED = -3
DN_ERR_WORKSPACE = -4

ACT = {"none": 0, "relu": 1, "relu6": 2, "hardswish": 3}
OP_STEM, OP_DW, OP_PW, OP_SE, OP_PWDW, OP_NOP, OP_DWPW = 0, 1, 2, 3, 4, 5, 6
BUF_NONE, BUF_IMAGES = -1, -2

c_void_p, c_int, c_int32, c_int64, c_float, c_double, c_size_t = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_size_t)


class PostprocessParams(ctypes.Structure):
    _fields_ = [("num_priors", c_int32), ("num_classes", c_int32), ("image_h", c_int32), ("image_w", c_int32),
                ("score_thresh", c_float), ("nms_thresh", c_double), ("topk_candidates", c_int32),
                ("detections_per_img", c_int32), ("min_box_size", c_float), ("box_weights", c_float * 4),
                ("bbox_xform_clip", c_float)]


class Op(ctypes.Structure):
    _fields_ = [("kind", c_int32), ("act", c_int32), ("in_buf", c_int32), ("out_buf", c_int32), ("res_buf", c_int32),
                ("h_in", c_int32), ("w_in", c_int32), ("c_in", c_int32), ("h_out", c_int32), ("w_out", c_int32),
                ("c_out", c_int32), ("ksize", c_int32), ("stride", c_int32), ("c_mid", c_int32), ("out_fp32", c_int32),
                ("lane", c_int32), ("act2", c_int32), ("se_fold", c_int32), ("w_off", c_int64), ("b_off", c_int64), ("w2_off", c_int64), ("b2_off", c_int64),
                ("out_batch_stride", c_int64), ("out_row_stride", c_int64), ("out_offset", c_int64)]


class Buf(ctypes.Structure):
    _fields_ = [("elems_per_image", c_int64), ("elem_bytes", c_int32), ("reserved", c_int32)]


class EngineStats(ctypes.Structure):
    _fields_ = [("launches_per_forward", c_int32), ("fused_pwdw", c_int32), ("fused_dwpw", c_int32),
                ("se_layers", c_int32), ("se_pooled", c_int32), ("se_folded", c_int32), ("reserved", c_int32),
                ("pipeline_slots", c_int32), ("last_slot", c_int32),
                ("act_dtype", c_int32), ("forwards", c_int64), ("graph_replays", c_int64)]


class ModelDesc(ctypes.Structure):
    _fields_ = [("image_h", c_int32), ("image_w", c_int32), ("image_mean", c_float * 3), ("image_std", c_float * 3),
                ("n_ops", c_int32), ("n_bufs", c_int32), ("ops_host", ctypes.POINTER(Op)),
                ("bufs_host", ctypes.POINTER(Buf)), ("logits_buf", c_int32), ("bbox_buf", c_int32),
                ("anchors_host", ctypes.POINTER(c_float)), ("post", PostprocessParams), ("gemm_impl", c_int32),
                ("use_cuda_graph", c_int32), ("pipeline_slots", c_int32), ("reserved", c_int32)]


_SIGNATURES = {
    "dn_last_error": (ctypes.c_char_p, []),
    "dn_abi_version": (c_int, []),
    "dn_dwconv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dn_dwconv_plan_info": (c_int, [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_int32)]),
    "dn_pwconv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                          c_int64, c_int64, c_int, c_void_p]),
    "dn_pwconv_plan_info": (c_int, [ctypes.c_longlong, c_int, c_int, ctypes.POINTER(c_int32)]),
    "dn_pwdw_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                              c_int, c_int, c_int, c_int, c_void_p]),
    "dn_dwpw_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                              c_int, c_int, c_int, c_int, c_void_p]),
    "dn_conv3x3": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                           c_int, c_int64, c_int64, c_void_p]),
    "dn_conv3x3_first": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_float), ctypes.POINTER(c_float), c_void_p,
                                 c_int, c_int, c_int, c_int, c_void_p]),
    "dn_im2col3x3_first": (c_int, [c_void_p, ctypes.POINTER(c_float), ctypes.POINTER(c_float), c_void_p, c_int, c_int, c_int,
                                   c_void_p]),
    "dn_maxpool2d": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dn_l2norm_scale": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "dn_stem_conv": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_float), ctypes.POINTER(c_float), c_void_p,
                             c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dn_se_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "dn_se_inplace": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                              c_size_t, c_void_p]),
    "dn_se_project": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                              c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "dn_dwconv_se": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t, ctypes.POINTER(c_int), c_void_p]),
    "dn_postprocess_workspace_bytes": (c_size_t, [c_int, ctypes.POINTER(PostprocessParams)]),
    "dn_postprocess": (c_int, [c_void_p, c_void_p, c_void_p, c_int, ctypes.POINTER(PostprocessParams), c_void_p, c_size_t,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dn_postprocess_scored": (c_int, [c_void_p, c_void_p, c_int, ctypes.POINTER(PostprocessParams), c_void_p, c_size_t,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dn_postprocess_profile": (c_int, [c_void_p, c_void_p, c_void_p, c_int, ctypes.POINTER(PostprocessParams), c_void_p,
                                       c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_int, ctypes.POINTER(c_float),
                                       c_void_p]),
    "dn_batched_nms_workspace_bytes": (c_size_t, [c_int64]),
    "dn_batched_nms": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p, c_size_t, c_void_p, c_void_p,
                               c_void_p]),
    "dn_ssd_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "dn_ssd_match": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "dn_match_quality": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "dn_ssd_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                            c_float, ctypes.POINTER(c_float), c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "dn_engine_create": (c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(ModelDesc), c_int]),
    "dn_engine_destroy": (c_int, [c_void_p]),
    "dn_engine_load_weights": (c_int, [c_void_p, c_void_p, c_size_t]),
    "dn_engine_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dn_engine_join": (c_int, [c_void_p, c_void_p]),
    "dn_engine_join_previous": (c_int, [c_void_p, c_void_p]),
    "dn_engine_forward_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dn_engine_forward_host_u8": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dn_resize_bilinear": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "dn_u8_to_f32": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "dn_rescale_boxes": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dn_detections_to_coco": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "dn_engine_buffer": (c_int, [c_void_p, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_int64)]),
    "dn_engine_copy_buffer": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "dn_engine_profile": (c_int, [c_void_p, c_void_p, c_int, c_int, ctypes.POINTER(c_float), c_void_p]),
    "dn_engine_launches_per_forward": (c_int, [c_void_p]),
    "dn_engine_get_stats": (c_int, [c_void_p, ctypes.POINTER(EngineStats)]),
    "dn_engine_device_bytes": (c_size_t, [c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_libs = {}


def torch_dtype(act_dtype=None):
    import torch
    return {"fp16": torch.float16, "bf16": torch.bfloat16}[act_dtype or DEFAULT_ACT_DTYPE]


def dtype_name(torch_dt):
    """"fp16" / "bf16" for a torch activation dtype (the stage-level wrappers in ops.py pick the library by it)."""
    import torch
    if torch_dt == torch.float16:
        return "fp16"
    if torch_dt == torch.bfloat16:
        return "bf16"
    raise TypeError("activations must be float16 or bfloat16, got %s" % torch_dt)


def lib(act_dtype=None):
    """Load the shared library of one activation storage type (once).  Raises if it has not been built -- there is
    no CPU path."""
    dt = act_dtype or DEFAULT_ACT_DTYPE
    if dt not in LIB_PATHS:
        raise ValueError("act_dtype must be one of %s, got %r" % (ACT_DTYPES, dt))
    if dt not in _libs:
        path = LIB_PATHS[dt]
        if not os.path.exists(path):
            raise RuntimeError(
                "demonet_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or demonet_b200/csrc/build.sh); there is no CPU or PyTorch fallback." % path)
        handle = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        if handle.dn_abi_version() != ABI_VERSION:
            raise RuntimeError("demonet_b200: ABI version mismatch (library %d, binding %d); rebuild with `python -c 'import __graft_entry__ as g; g.build()'`" % (handle.dn_abi_version(), ABI_VERSION))
        _libs[dt] = handle
    return _libs[dt]


def check(rc, handle=None):
    """Map a dn_status to the exception type the reference would raise for the same mistake.  `handle`: the library
    the failing call went through (its thread-local message); default = the default-dtype library."""
    if rc == DN_OK:
        return
    msg = (handle or lib()).dn_last_error().decode("utf-8", "replace")
    if rc == DN_ERR_INVALID:
        raise ValueError("demonet_b200: " + msg)
    if rc == DN_ERR_UNSUPPORTED:
        raise NotImplementedError("demonet_b200: " + msg)
    raise RuntimeError("demonet_b200 (status %d): %s" % (rc, msg))
