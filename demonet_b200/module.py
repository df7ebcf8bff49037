"""Drop-in `nn.Module` for demonet's SSDLite detectors, executing on the B200 engine.

Same constructor arguments, `state_dict` keys, `forward(images, targets=None)` input contract and
`List[Dict[boxes, scores, labels]]` output contract as the reference's `SSD`
(demonet/models/generalized_ssd.py:95-397), inference only.  The arithmetic runs in
libdemonet_b200.so through ctypes; PyTorch only owns parameters, device memory and streams.
"""
import ctypes
import math
import os
import warnings
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from . import _C, plan as _plan


class _Engine:
    """Owns one dn_engine (one device, one max batch) and the weight blob uploaded to it."""

    def __init__(self, plan: _plan.Plan, desc_kwargs: dict, max_batch: int, device: torch.device):
        self.plan = plan
        self.device = device
        self.max_batch = max_batch
        self.t2b, self.bufs, self.logits_buf, self.bbox_buf = _plan.assign_buffers(
            plan, reuse=not desc_kwargs.get("keep_activations", False))
        self.anchors = _plan.default_boxes(plan)
        self._handle = ctypes.c_void_p()
        self._weights_token = None
        self._desc_kwargs = desc_kwargs
        self.act_dtype = desc_kwargs.get("act_dtype") or _C.DEFAULT_ACT_DTYPE
        self._lib = _C.lib(self.act_dtype)          # raises if that build is missing: no fallback
        self._ops = None
        self._created = False

    def _check(self, rc):
        _C.check(rc, self._lib)

    def describe(self, offsets):
        """The dn_model_desc the C engine is created from (ops / buffer arrays are kept alive on self)."""
        plan, kw = self.plan, self._desc_kwargs
        fuse = not kw.get("keep_activations", False) and int(os.environ.get("DN_FUSE", "1")) != 0 and int(kw.get("gemm_impl", 0)) == 0
        self._ops = _plan.build_ops(plan, offsets, self.t2b, self.logits_buf, self.bbox_buf, fuse=fuse)
        bufs = (_C.Buf * len(self.bufs))()
        for i, (elems, nbytes) in enumerate(self.bufs):
            bufs[i].elems_per_image, bufs[i].elem_bytes = elems, nbytes
        self._bufs_c = bufs
        d = _C.ModelDesc()
        d.image_h = d.image_w = plan.size
        d.image_mean = (ctypes.c_float * 3)(*kw["image_mean"])
        d.image_std = (ctypes.c_float * 3)(*kw["image_std"])
        d.n_ops, d.n_bufs = len(plan.layers), len(self.bufs)
        d.ops_host = ctypes.cast(self._ops, ctypes.POINTER(_C.Op))
        d.bufs_host = ctypes.cast(bufs, ctypes.POINTER(_C.Buf))
        d.logits_buf, d.bbox_buf = self.logits_buf, self.bbox_buf
        d.anchors_host = self.anchors.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        d.post = make_post_params(plan.num_priors, plan.num_classes, plan.size, plan.size, kw["score_thresh"],
                                  kw["nms_thresh"], kw["topk_candidates"], kw["detections_per_img"],
                                  kw.get("min_box_size", -1.0))
        d.gemm_impl = int(kw.get("gemm_impl", 0))
        d.use_cuda_graph = int(kw.get("use_cuda_graph", 1))
        d.pipeline_slots = int(kw.get("pipeline_slots", 0))
        return d

    def _create(self, offsets):
        d = self.describe(offsets)
        with torch.cuda.device(self.device):
            self._check(self._lib.dn_engine_create(ctypes.byref(self._handle), ctypes.byref(d), self.max_batch))
        self._created = True

    def load_weights(self, sd, token):
        blob, offsets = _plan.pack_weights(self.plan, sd, self.act_dtype)
        if not self._created:
            self._create(offsets)
        buf = ctypes.create_string_buffer(blob, len(blob))
        with torch.cuda.device(self.device):
            self._check(self._lib.dn_engine_load_weights(self._handle, buf, len(blob)))
        self._weights_token = token

    def forward(self, images: Tensor, out):
        B = images.shape[0]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self._lib.dn_engine_forward(self._handle, images.data_ptr(), B, out["boxes"].data_ptr(),
                                               out["scores"].data_ptr(), out["labels"].data_ptr(),
                                               out["counts"].data_ptr(), stream))

    @property
    def pipelined(self) -> bool:
        return self.n_slots >= 2

    @property
    def n_slots(self) -> int:
        """Forwards the engine keeps in flight (1 = plain stream semantics, 2..4 = pipeline mode)."""
        return max(1, int(self._desc_kwargs.get("pipeline_slots", 0)))

    def join(self):
        """Pipeline mode: order every forward issued so far before later work on the current stream."""
        with torch.cuda.device(self.device):
            self._check(self._lib.dn_engine_join(self._handle, torch.cuda.current_stream(self.device).cuda_stream))

    def join_previous(self):
        """Pipeline mode with n slots: order the OLDEST forward in flight (issued n - 1 calls before the most recent one, the
        one whose slot the next call reuses) before later work on the current stream; a no-op until n forwards were issued."""
        with torch.cuda.device(self.device):
            self._check(self._lib.dn_engine_join_previous(self._handle, torch.cuda.current_stream(self.device).cuda_stream))

    def forward_host(self, images: Tensor, out):
        B = images.shape[0]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self._lib.dn_engine_forward_host(self._handle, images.data_ptr(), B, out["boxes"].data_ptr(),
                                                    out["scores"].data_ptr(), out["labels"].data_ptr(),
                                                    out["counts"].data_ptr(), stream))

    def forward_host_u8(self, images_u8: Tensor, out):
        B = images_u8.shape[0]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self._lib.dn_engine_forward_host_u8(self._handle, images_u8.data_ptr(), B, out["boxes"].data_ptr(),
                                                       out["scores"].data_ptr(), out["labels"].data_ptr(),
                                                       out["counts"].data_ptr(), stream))

    def buffer(self, tensor_name: str, batch: int) -> Tensor:
        """Copy of an intermediate activation as fp32 NCHW (for stage-by-stage parity tests)."""
        h, w, c = self.plan.tensors[tensor_name]
        return self._read(self.t2b[tensor_name], batch, (h, w, c), _C.torch_dtype(self.act_dtype)).permute(0, 3, 1, 2).float()

    def head_outputs(self, batch: int) -> Tuple[Tensor, Tensor]:
        P, K = self.plan.num_priors, self.plan.num_classes
        return (self._read(self.logits_buf, batch, (P, K), torch.float32),
                self._read(self.bbox_buf, batch, (P, 4), torch.float32))

    def _read(self, buf_id, batch, shape, dtype):
        # pipeline mode: the C side reads the arena of the slot that ran the forward issued last
        n = batch * int(np.prod(shape))
        out = torch.empty(n, dtype=dtype, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self._lib.dn_engine_copy_buffer(self._handle, buf_id, out.data_ptr(), n * out.element_size(), stream))
        return out.view(batch, *shape)

    def stats(self) -> dict:
        """What the forward issued last ran: fused launches, pooled SE layers, graph replays, slots, storage type."""
        st = _C.EngineStats()
        self._check(self._lib.dn_engine_get_stats(self._handle, ctypes.byref(st)))
        out = {name: int(getattr(st, name)) for name, _ in _C.EngineStats._fields_}
        out["act_dtype"] = _C.ACT_DTYPES[1 - out["act_dtype"]]          # 1 = fp16, 0 = bf16
        return out

    @property
    def launches_per_forward(self):
        return self._lib.dn_engine_launches_per_forward(self._handle)

    @property
    def device_bytes(self):
        return self._lib.dn_engine_device_bytes(self._handle)

    def close(self):
        if self._created and self._handle:
            self._lib.dn_engine_destroy(self._handle)
            self._handle = ctypes.c_void_p()
            self._created = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_post_params(P, K, image_h, image_w, score_thresh, nms_thresh, topk_candidates, detections_per_img,
                     min_box_size=-1.0) -> _C.PostprocessParams:
    p = _C.PostprocessParams()
    p.num_priors, p.num_classes, p.image_h, p.image_w = P, K, image_h, image_w
    p.score_thresh = score_thresh
    p.nms_thresh = float(nms_thresh)
    p.topk_candidates = int(topk_candidates) if topk_candidates else 0
    p.detections_per_img = int(detections_per_img)
    p.min_box_size = float(min_box_size)
    p.box_weights = (ctypes.c_float * 4)(10.0, 10.0, 5.0, 5.0)      # generalized_ssd.py:170
    p.bbox_xform_clip = math.log(1000.0 / 16)                        # _utils.py:135
    return p


def resize_bilinear(image: Tensor, size: Tuple[int, int], out: Optional[Tensor] = None) -> Tensor:
    """Fixed-size bilinear resize of one CUDA [C,H,W] image (fp32 in [0,1], or uint8 -> x / 255 first) to fp32
    [C,size[0],size[1]] -- the interpolate call of _resize_image_and_masks (transform.py:27-53) on the device."""
    if not image.is_cuda:
        raise RuntimeError("demonet_b200 operators run on CUDA tensors only (no CPU fallback)")
    if image.dim() != 3:
        raise ValueError("expected a [C, H, W] image, got {}".format(tuple(image.shape)))
    if image.dtype not in (torch.float32, torch.uint8):
        image = image.float()
    image = image.contiguous()
    C, H, W = image.shape
    if out is None:
        out = torch.empty(C, size[0], size[1], dtype=torch.float32, device=image.device)
    with torch.cuda.device(image.device):
        _C.check(_C.lib().dn_resize_bilinear(image.data_ptr(), int(image.dtype == torch.uint8), C, H, W, out.data_ptr(),
                                             size[0], size[1], torch.cuda.current_stream(image.device).cuda_stream))
    return out


def u8_to_f32(src: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """ToTensor's conversion on the device: uint8 -> fp32 / 255 (exact fp32 division)."""
    if not src.is_cuda or src.dtype != torch.uint8:
        raise RuntimeError("u8_to_f32 takes a CUDA uint8 tensor (no CPU fallback)")
    src = src.contiguous()
    if out is None:
        out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    if src.numel() == 0:
        return out
    with torch.cuda.device(src.device):
        _C.check(_C.lib().dn_u8_to_f32(src.data_ptr(), out.data_ptr(), src.numel(),
                                       torch.cuda.current_stream(src.device).cuda_stream))
    return out


def rescale_boxes_(boxes: Tensor, original_sizes: List[Tuple[int, int]], new_size: Tuple[int, int]) -> Tensor:
    """In place boxes[b] *= (rw, rh, rw, rh) for a padded CUDA [B,D,4] batch: resize_boxes (transform.py:278-292) with
    the ratios formed in fp32 exactly as the reference forms them (orig / new as float32 tensors)."""
    if not boxes.is_cuda:
        raise RuntimeError("demonet_b200 operators run on CUDA tensors only (no CPU fallback)")
    B, D = boxes.shape[0], boxes.shape[1]
    orig = torch.tensor(original_sizes, dtype=torch.float32)
    ratios = (orig / torch.tensor(new_size, dtype=torch.float32)).contiguous()        # [B,2] = (rh, rw)
    ratios = ratios.to(boxes.device, non_blocking=False)
    with torch.cuda.device(boxes.device):
        _C.check(_C.lib().dn_rescale_boxes(boxes.data_ptr(), ratios.data_ptr(), B, D,
                                           torch.cuda.current_stream(boxes.device).cuda_stream))
    return boxes


class SSDLiteB200(nn.Module):
    """SSDLite detector on the B200 engine: the eval branch of SSD.forward (detections) and, in training mode, its loss
    branch evaluated on the engine's head outputs (generalized_ssd.py:321-337; `demonet_b200.loss`).

    Args mirror `SSD.__init__` (generalized_ssd.py:154-163): score_thresh, nms_thresh,
    detections_per_img, topk_candidates, image_mean, image_std.  `postprocess` selects the
    reference flavour: "ssd" = SSD.postprocess_detections (generalized_ssd.py:351-397),
    "legacy" = PostProcess.forward (box_head.py:323-381: no per-class top-k, remove_small_boxes).
    """

    def __init__(self, plan: _plan.Plan, score_thresh=0.01, nms_thresh=0.45, detections_per_img=200,
                 topk_candidates=400, image_mean=None, image_std=None, postprocess="ssd", init="normal",
                 gemm_impl=0, use_cuda_graph=True, keep_activations=False, pipeline_slots=0, act_dtype=None,
                 iou_thresh=0.5, positive_fraction=0.25):
        super().__init__()
        self.iou_thresh = iou_thresh                                          # SSDMatcher(iou_thresh), generalized_ssd.py:184
        self.neg_to_pos_ratio = (1.0 - positive_fraction) / positive_fraction  # generalized_ssd.py:197
        if postprocess not in ("ssd", "legacy"):
            raise ValueError("postprocess must be 'ssd' or 'legacy'")
        self.plan = plan
        self.size = (plan.size, plan.size)
        self.score_thresh = score_thresh
        self.nms_thresh = nms_thresh
        self.detections_per_img = detections_per_img
        self.topk_candidates = topk_candidates if postprocess == "ssd" else 0
        self.postprocess_flavour = postprocess
        self.image_mean = list(image_mean) if image_mean is not None else [0.485, 0.456, 0.406]
        self.image_std = list(image_std) if image_std is not None else [0.229, 0.224, 0.225]
        self._gemm_impl = gemm_impl
        self._use_cuda_graph = use_cuda_graph
        self._pipeline_slots = int(pipeline_slots)      # 2..4: consecutive batches overlap on that many engine instances
        if not 0 <= self._pipeline_slots <= 4:
            raise ValueError("pipeline_slots must be in 0..4 (got %d)" % self._pipeline_slots)
        self.act_dtype = act_dtype or _C.DEFAULT_ACT_DTYPE      # "fp16" (default) or "bf16": activation storage type
        if self.act_dtype not in _C.ACT_DTYPES:
            raise ValueError("act_dtype must be one of %s" % (_C.ACT_DTYPES,))
        self._keep_activations = keep_activations      # debug: one arena buffer per tensor
        self._engines: Dict[Tuple[str, int], _Engine] = {}
        self._io: Dict[Tuple[str, int], dict] = {}
        self._weights_epoch = 0                        # bumped whenever the parameters may have changed
        self._last_input_ptr = None
        self._register_parameters(init)
        self.eval()

    # ---- parameters: same keys / shapes as the reference's state_dict ----------------------
    def _register_parameters(self, init):
        g = torch.Generator().manual_seed(0)
        for key, shape, role in self.plan.param_specs:
            parts = key.split(".")
            mod = self
            for name in parts[:-1]:
                if name not in mod._modules:
                    mod.add_module(name, nn.Module())
                mod = mod._modules[name]
            if role == "conv_w":
                # _normal_init: N(0, 0.03) weights, zero biases (ssd_mobilenetv3.py:57-62)
                t = torch.randn(shape, generator=g) * 0.03 if init == "normal" else torch.zeros(shape)
                mod.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))
            elif role == "conv_b":
                mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape), requires_grad=False))
            elif role == "bn_w":
                mod.register_parameter(parts[-1], nn.Parameter(torch.ones(shape), requires_grad=False))
            elif role == "bn_b":
                mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape), requires_grad=False))
            elif role == "bn_m":
                mod.register_buffer(parts[-1], torch.zeros(shape))
            elif role == "bn_v":
                mod.register_buffer(parts[-1], torch.ones(shape))
            else:
                mod.register_buffer(parts[-1], torch.tensor(0, dtype=torch.long))

    # train(True) selects the LOSS branch of SSD.forward (generalized_ssd.py:321-337): forward(images, targets) then returns
    # {'bbox_regression', 'classification'} computed by dn_ssd_match / dn_ssd_loss on the engine's head outputs.  The engine
    # always runs with the folded running statistics of BatchNorm and produces no gradients for the parameters (they are
    # registered with requires_grad=False): this is loss EVALUATION; `demonet_b200.loss.compute_loss` is differentiable with
    # respect to head outputs for callers that train a torch head / backbone.

    def anchors(self, device) -> Tensor:
        """Default boxes [P,4] on `device` (DefaultBoxGenerator, anchor_utils.py:110-126)."""
        key = str(device)
        cache = self.__dict__.setdefault("_anchor_cache", {})
        if key not in cache:
            cache[key] = torch.from_numpy(_plan.default_boxes(self.plan)).to(device)
        return cache[key]

    def _losses(self, images, targets):
        from . import loss as _loss
        if targets is None:
            raise ValueError("In training mode, targets should be passed")                     # generalized_ssd.py:273-274
        if isinstance(images, Tensor):
            images = list(images.unbind(0))
        _loss.check_targets(targets)
        if not torch.cuda.is_available():
            raise RuntimeError("demonet_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        S = self.plan.size
        first = images[0]
        device = first.device if first.is_cuda else torch.device("cuda", torch.cuda.current_device())
        sizes = [(int(im.shape[-2]), int(im.shape[-1])) for im in images]
        batch = torch.empty(len(images), 3, S, S, dtype=torch.float32, device=device)
        for i, img in enumerate(images):
            if sizes[i] != (S, S):
                resize_bilinear(img.to(device), (S, S), out=batch[i])                           # transform.py:27-53
            else:
                batch[i].copy_(img)
        targets = [{k: v.to(device) for k, v in t.items()} for t in targets]
        return _loss.detector_losses(self, batch, targets, sizes, self.iou_thresh, self.neg_to_pos_ratio)

    # ---- engine management -----------------------------------------------------------------
    # The engine holds BN-folded, re-laid-out copies of the parameters.  They are refreshed when the weights change
    # through the nn.Module API (load_state_dict, .to() / .cuda() / .half() ... via _apply); after editing parameters in
    # place (p.data.mul_(...), optimiser steps) call refresh_weights().  A forward only compares one integer.
    def refresh_weights(self):
        """Fold and upload the current parameters again on the next forward."""
        self._weights_epoch += 1

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._weights_epoch += 1
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._weights_epoch += 1
        return out

    def _engine_for(self, device: torch.device, batch: int) -> _Engine:
        key = str(device)
        eng = self._engines.get(key)
        if eng is None or eng.max_batch < batch:
            if eng is not None:
                eng.close()
                self._io = {k: v for k, v in self._io.items() if k[0] != key}
            kw = dict(image_mean=self.image_mean, image_std=self.image_std, score_thresh=self.score_thresh,
                      nms_thresh=self.nms_thresh, topk_candidates=self.topk_candidates,
                      detections_per_img=self.detections_per_img, gemm_impl=self._gemm_impl,
                      use_cuda_graph=int(self._use_cuda_graph), keep_activations=self._keep_activations,
                      pipeline_slots=self._pipeline_slots, act_dtype=self.act_dtype,
                      min_box_size=1e-2 if self.postprocess_flavour == "legacy" else -1.0)
            eng = _Engine(self.plan, kw, max(batch, 1), device)
            self._engines[key] = eng
        if eng._weights_token != self._weights_epoch:
            eng.load_weights(self.state_dict(), self._weights_epoch)
        return eng

    def _io_buffers(self, device: torch.device, batch: int, host: bool):
        key = (str(device) + ("/host" if host else ""), batch)
        io = self._io.get(key)
        if io is None:
            D, S = self.detections_per_img, self.plan.size
            kw = dict(device="cpu", pin_memory=True) if host else dict(device=device)
            io = {"images": torch.empty(batch, 3, S, S, dtype=torch.float32, **kw),
                  "boxes": torch.empty(batch, D, 4, dtype=torch.float32, **kw),
                  "scores": torch.empty(batch, D, dtype=torch.float32, **kw),
                  "labels": torch.empty(batch, D, dtype=torch.int64, **kw),
                  "counts": torch.empty(batch, dtype=torch.int32, **kw)}
            self._io[key] = io
        return io

    def reserve(self, batch: int, device=None):
        """Pre-build the engine for `batch` images (otherwise done lazily on the first forward)."""
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        return self._engine_for(device, batch)

    # ---- forward ---------------------------------------------------------------------------
    def forward(self, images: List[Tensor], targets: Optional[List[Dict[str, Tensor]]] = None):
        if self.training:
            return self._losses(images, targets)
        if isinstance(images, Tensor):
            if images.dim() != 4:
                raise ValueError("images is expected to be a list of 3d tensors of shape [C, H, W] "
                                 "or a batched 4d tensor, got {}".format(images.shape))
            images = list(images.unbind(0))
        if len(images) == 0:
            return []
        B, S = len(images), self.plan.size
        first = images[0]
        step = 3 * S * S * 4
        # Fast path for the common case -- B float32 [3,S,S] CUDA tensors: torch.stack validates shapes, dtypes and devices
        # in C++ (one launch for the whole list), so the per-image Python checks below only run when something is off, to
        # raise the reference's own exceptions (transform.py:110-112, 130-134) or to take the resize / host paths.
        stacked = None
        if (isinstance(first, Tensor) and first.is_cuda and first.dim() == 3 and first.dtype == torch.float32
                and tuple(first.shape) == (3, S, S)):
            self._engine_for(first.device, B)         # (a larger batch re-creates the engine and its staging buffers first)
            base = first._base
            if (base is not None and base.dim() == 4 and tuple(base.shape) == (B, 3, S, S) and base.is_contiguous()
                    and base.dtype == torch.float32 and images[-1]._base is base
                    and all(im.data_ptr() == base.data_ptr() + i * step for i, im in enumerate(images))):
                stacked = base                                        # the list is the unbind() of one batch: no copy at all
            else:
                try:
                    stacked = torch.stack(images, 0, out=self._io_buffers(first.device, B, False)["images"])
                except (RuntimeError, TypeError):
                    stacked = None                                    # mixed shapes / dtypes / devices: the general path decides
        original_sizes: List[Tuple[int, int]] = []
        in_place = False
        resized = False
        if stacked is not None:
            original_sizes = [(S, S)] * B
            base = stacked.data_ptr()
            in_place = stacked is not self._io.get((str(first.device), B), {}).get("images")
        else:
            base = first.data_ptr() if first.dim() == 3 else 0
            in_place = first.dtype == torch.float32 and first.dim() == 3
            for i, img in enumerate(images):
                shp = img.shape
                if len(shp) != 3:
                    raise ValueError("images is expected to be a list of 3d tensors "
                                     "of shape [C, H, W], got {}".format(shp))                 # transform.py:110-112
                if not img.is_floating_point():
                    raise TypeError("Expected input images to be of floating type (in range [0, 1]), "
                                    f"but found type {img.dtype} instead")                # transform.py:130-134
                h, w = shp[1], shp[2]
                original_sizes.append((h, w))
                if h != S or w != S:
                    resized = True
                if in_place and (img.data_ptr() != base + i * step or img.dtype != torch.float32 or not img.is_contiguous()):
                    in_place = False
        if not torch.cuda.is_available():
            raise RuntimeError("demonet_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        in_dev = first.device
        host = in_dev.type != "cuda"
        device = torch.device("cuda", torch.cuda.current_device()) if host else in_dev
        eng = self._engine_for(device, B)
        # images that need the fixed-size resize (transform.py:27-53) go through the device kernel, so a host batch
        # with such images is assembled on the device; an all-S x S host batch takes the pinned-staging path
        io = self._io_buffers(device, B, host and not resized)
        batch = io["images"]
        if stacked is not None:
            if in_place:
                # a caller-owned batch is used in place only when the caller keeps reusing it (same address as in the
                # previous call): the engine keys its CUDA graphs by address, so a fresh address every call would mean an
                # eager run every call
                if self._last_input_ptr == base:
                    batch = stacked
                else:
                    batch.copy_(stacked)
                self._last_input_ptr = base
            # else: torch.stack has already filled the staging batch
        elif not resized:
            if any(im.shape[0] != 3 for im in images):
                raise ValueError("images must have 3 channels")
            in_place = (in_place and not host and first.untyped_storage().nbytes() - first.storage_offset() * 4 >= B * step)
            if in_place:
                whole = torch.as_strided(first, (B, 3, S, S), (3 * S * S, S * S, S, 1))
                if self._last_input_ptr == base:
                    batch = whole
                else:
                    batch.copy_(whole)
                self._last_input_ptr = base
            else:
                # one launch for the whole list instead of one copy per image
                torch.stack(images if first.dtype == torch.float32 else [im.float() for im in images], 0, out=batch)
        else:
            for i, img in enumerate(images):
                if tuple(img.shape[-2:]) != (S, S):
                    resize_bilinear(img.to(device), (S, S), out=batch[i])
                else:
                    batch[i].copy_(img)
        if host and not resized:
            eng.forward_host(batch, io)
            torch.cuda.current_stream(device).synchronize()
        else:
            eng.forward(batch, io)
            if eng.pipelined:
                eng.join()                        # a single call has nothing to overlap with: plain stream semantics
            if resized:                           # transform.postprocess / resize_boxes, transform.py:228-292
                rescale_boxes_(io["boxes"], original_sizes, (S, S))
        return self._detections(io, B, in_dev if host else None)

    def _detections(self, io, B, to_device=None):
        # The padded outputs are copied ONCE (three launches for the whole batch -- they are the engine's reusable output
        # buffers) and every image gets views of the copies: fresh tensors as the reference returns, without 3 B clones
        # and, for full rows, without 3 B slicing calls either.
        D = self.detections_per_img
        boxes, scores, labels = io["boxes"][:B].clone(), io["scores"][:B].clone(), io["labels"][:B].clone()
        counts = io["counts"][:B].tolist()        # the one host sync: data-dependent output shapes
        if to_device is not None:
            boxes, scores, labels = boxes.to(to_device), scores.to(to_device), labels.to(to_device)
        bl, sl, ll = boxes.unbind(0), scores.unbind(0), labels.unbind(0)
        return [{"boxes": bl[i], "scores": sl[i], "labels": ll[i]} if n == D else
                {"boxes": bl[i][:n], "scores": sl[i][:n], "labels": ll[i][:n]} for i, n in enumerate(counts)]

    def forward_uint8(self, images: Tensor):
        """uint8 ingest: `images` is a [B,3,S,S] uint8 batch as a decoder produces it; the ToTensor
        conversion x / 255 runs on the device (exact fp32 division), so the result equals
        `self(list(images.float() / 255))`.  A pinned host batch crosses PCIe at one byte per sample."""
        if images.dtype != torch.uint8 or images.dim() != 4 or tuple(images.shape[1:]) != (3, self.plan.size, self.plan.size):
            raise ValueError("forward_uint8 expects a uint8 [B,3,{0},{0}] batch, got {1} {2}".format(
                self.plan.size, images.dtype, tuple(images.shape)))
        if not torch.cuda.is_available():
            raise RuntimeError("demonet_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        B = images.shape[0]
        if B == 0:
            return []
        host = not images.is_cuda
        device = torch.device("cuda", torch.cuda.current_device()) if host else images.device
        eng = self._engine_for(device, B)
        io = self._io_buffers(device, B, host)
        if host:
            key = (str(device) + "/host_u8", B)
            stage = self._io.get(key)
            if stage is None:
                stage = self._io[key] = {"images": torch.empty(images.shape, dtype=torch.uint8, pin_memory=True)}
            stage["images"].copy_(images)
            eng.forward_host_u8(stage["images"], io)
            torch.cuda.current_stream(device).synchronize()
        else:
            u8_to_f32(images, out=io["images"])
            eng.forward(io["images"], io)
            if eng.pipelined:
                eng.join()
        return self._detections(io, B, images.device if host else None)

    def forward_padded(self, images: Tensor):
        """[B,3,S,S] fp32 CUDA batch -> fresh padded (boxes [B,D,4], scores [B,D], labels [B,D], counts [B]) with no host
        synchronisation and no data-dependent shape: the contract of torch.ops.demonet_b200.ssdlite_forward."""
        S = self.plan.size
        if images.dim() != 4 or tuple(images.shape[1:]) != (3, S, S) or not images.is_cuda:
            raise ValueError("forward_padded expects a CUDA batch of shape [B,3,%d,%d]" % (S, S))
        B = images.shape[0]
        eng = self._engine_for(images.device, B)
        io = self._io_buffers(images.device, B, False)
        io["images"].copy_(images)
        eng.forward(io["images"], io)
        if eng.pipelined:
            eng.join()
        return io["boxes"].clone(), io["scores"].clone(), io["labels"].clone(), io["counts"].clone()

    def forward_batches(self, batches):
        """Throughput path: iterate over [B,3,S,S] fp32 CUDA batches and yield their detections, keeping n batches in
        flight when the model was built with pipeline_slots=n (2..4; the post-processing tail of batch i overlaps the
        backbone of the batches behind it).  Every batch must stay alive and unchanged until its detections have been yielded."""
        pending = []                               # batches in flight, oldest first: (io, B, engine, source tensor)
        slot = 0
        for images in batches:
            if images.dim() != 4 or tuple(images.shape[1:]) != (3, self.plan.size, self.plan.size) or not images.is_cuda:
                raise ValueError("forward_batches expects fp32 CUDA batches of shape [B,3,%d,%d]" % (self.plan.size, self.plan.size))
            B = images.shape[0]
            cur = self._engines.get(str(images.device))
            if pending and cur is not None and cur.max_batch < B:
                # a larger batch re-creates the engine: finish and hand out what is still in flight on the old one first
                pending[0][2].join()
                for done in pending:
                    yield self._detections(done[0], done[1])
                pending = []
            eng = self._engine_for(images.device, B)
            key = (str(images.device) + "/slot%d" % slot, B)
            io = self._io.get(key)
            if io is None:
                io = self._io[key] = {k: v for k, v in self._io_buffers(images.device, B, False).items() if k == "images"}
                D = self.detections_per_img
                io.update(boxes=torch.empty(B, D, 4, device=images.device), scores=torch.empty(B, D, device=images.device),
                          labels=torch.empty(B, D, dtype=torch.int64, device=images.device),
                          counts=torch.empty(B, dtype=torch.int32, device=images.device))
            src = images if images.dtype == torch.float32 and images.is_contiguous() else images.float().contiguous()
            eng.forward(src, io)
            pending.append((io, B, eng, src))      # `src` (possibly a temporary) stays alive until its batch was joined
            depth = max(2, eng.n_slots)            # (one engine instance: still two sets of outputs, batch i - 1 is read while i runs)
            if len(pending) >= depth:              # every slot is busy: the oldest batch is handed out before its slot is reused
                if eng.pipelined:
                    eng.join_previous()
                done = pending.pop(0)
                yield self._detections(done[0], done[1])
            slot = (slot + 1) % depth
        if pending:
            pending[0][2].join()
            for done in pending:
                yield self._detections(done[0], done[1])

    def head_outputs(self, images: Tensor):
        """(cls_logits [B,P,K], bbox_regression [B,P,4]) of a [B,3,S,S] CUDA batch -- parity hook."""
        eng = self._engine_for(images.device, images.shape[0])
        io = self._io_buffers(images.device, images.shape[0], False)
        io["images"].copy_(images)
        eng.forward(io["images"], io)
        return eng.head_outputs(images.shape[0])
