"""Host-side logic on the CPU: state_dict contract, plans, weight packing, buffer assignment, and
the C ABI surface (symbols only -- no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import demonet_b200
from demonet_b200 import _C, plan as dplan
from oracle import boxes_np, net_ref, weights
from tests.plan_interp import run_plan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_v3_state_dict_contract(golden_dir):
    g = np.load(os.path.join(golden_dir, "v3_ssdlite.npz"))
    model = demonet_b200.ssdlite320_mobilenet_v3_large()
    sd = model.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["state_dict_keys"]]          # 476 keys, reference order
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in g["state_dict_shapes"]]
    assert sum(p.numel() for p in model.parameters()) == 3440060
    # reference defaults, ssd_mobilenetv3.py:207-217
    assert (model.score_thresh, model.nms_thresh, model.detections_per_img, model.topk_candidates) == (0.001, 0.55, 300, 300)
    assert model.image_mean == [0.5] * 3 and model.image_std == [0.5] * 3


def test_v2_state_dict_contract(golden_dir):
    g = np.load(os.path.join(golden_dir, "v2_ssdlite.npz"))
    model = demonet_b200.ssd_lite_mobilenet_v2()
    sd = model.state_dict()
    assert set(sd.keys()) == set(str(k) for k in g["state_dict_keys"])
    ref_shapes = dict(zip((str(k) for k in g["state_dict_keys"]), (str(s) for s in g["state_dict_shapes"])))
    assert all(str(tuple(v.shape)) == ref_shapes[k] for k, v in sd.items())
    assert (model.score_thresh, model.nms_thresh, model.detections_per_img) == (0.5, 0.45, 100)   # test_model.py:42-48


def test_builder_argument_errors():
    with pytest.raises(NotImplementedError):
        demonet_b200.ssdlite320_mobilenet_v3_large(norm_layer=torch.nn.BatchNorm2d)
    with pytest.raises(NotImplementedError):
        demonet_b200.ssdlite320_mobilenet_v3_large(width_mult=0.5)
    with pytest.raises(TypeError):
        demonet_b200.ssdlite320_mobilenet_v3_large(bogus=1)
    m = demonet_b200.ssdlite320_mobilenet_v3_large(topk_candidates=400, score_thresh=0.01)      # kwargs quirk, SURVEY 8(b)
    assert m.topk_candidates == 400 and m.score_thresh == 0.01
    m.train()                                     # loss branch of SSD.forward (generalized_ssd.py:273-274)
    with pytest.raises(ValueError, match="In training mode, targets should be passed"):
        m([torch.rand(3, 320, 320)])
    with pytest.raises(ValueError, match="Expected target boxes to be a tensor"):
        m([torch.rand(3, 320, 320)], [{"boxes": torch.zeros(4), "labels": torch.zeros(1, dtype=torch.int64)}])
    with pytest.raises(ValueError, match="All bounding boxes should have positive height and width"):
        m([torch.rand(3, 320, 320)], [{"boxes": torch.tensor([[5.0, 5.0, 5.0, 9.0]]), "labels": torch.ones(1, dtype=torch.int64)}])
    m.eval()
    with pytest.raises(ValueError):
        m([torch.rand(320, 320)])                 # transform.py:110-112
    with pytest.raises(TypeError):
        m([torch.zeros(3, 320, 320, dtype=torch.uint8)])          # transform.py:130-134


def test_plan_shapes():
    p = dplan.plan_ssdlite320_mobilenet_v3_large()
    assert p.grid_sizes == [(20, 20), (10, 10), (5, 5), (3, 3), (2, 2), (1, 1)] and p.num_priors == 3234
    kinds = [L.kind for L in p.layers]
    assert kinds.count("dw") == 31 and kinds.count("pw") == 50 and kinds.count("se") == 8 and kinds.count("stem") == 1
    assert sum(1 for L in p.layers if L.kind == "dw" and L.k == 5) == 6
    for S, P in ((300, 3000), (320, 3234), (512, 8190)):
        assert dplan.plan_ssd_lite_mobilenet_v2(21, S).num_priors == P
    a = dplan.default_boxes(p)
    assert np.array_equal(a, boxes_np.default_boxes(p.grid_sizes, (320, 320)))


def _interp(plan, sd, x, mean, std, round_activations, act_dtype):
    blob, offs = dplan.pack_weights(plan, sd, act_dtype)
    t2b, bufs, lb, bb = dplan.assign_buffers(plan)
    ops = dplan.build_ops(plan, offs, t2b, lb, bb)
    return run_plan(plan, blob, ops, bufs, lb, bb, x, mean, std, round_activations, act_dtype)


def _check_plan(model, sd, x, mean, std, forward_raw, act_dtype="fp16"):
    """dn_op array + packed blob + arena reuse, interpreted on the CPU, against the oracle.
    Tight in `w16` mode (no activation rounding -> only fp32 summation noise); statistical in `bf16`
    mode, where rounding flips make the chain chaotic (the oracle itself moves by rms 0.14 on the
    logits when only its summation order changes)."""
    with torch.no_grad():
        cls, reg = _interp(model.plan, sd, x, mean, std, False, act_dtype)
        ocls, oreg, _ = forward_raw(sd, x, "w16" if act_dtype == "bf16" else "wf16")
        assert not torch.isnan(cls).any() and not torch.isnan(reg).any()
        assert (cls - ocls).abs().max() < 2e-3 and (reg - oreg).abs().max() < 2e-3
        cls, reg = _interp(model.plan, sd, x, mean, std, True, act_dtype)
        ocls, oreg, _ = forward_raw(sd, x, act_dtype)
        tol = 0.3 if act_dtype == "bf16" else 0.05      # fp16 storage: 8x finer rounding
        assert (cls - ocls).pow(2).mean().sqrt() < tol and (reg - oreg).pow(2).mean().sqrt() < tol
        assert ocls.std() > 3.0


@pytest.mark.parametrize("act_dtype", ["fp16", "bf16"])
def test_v3_plan_reproduces_oracle(act_dtype):
    model = demonet_b200.ssdlite320_mobilenet_v3_large()
    sd = weights.seeded_state_dict(model.state_dict())
    _check_plan(model, sd, weights.synthetic_images(1, 320), [0.5] * 3, [0.5] * 3, net_ref.v3_forward_raw, act_dtype)


@pytest.mark.parametrize("S", [300, 512])
def test_v2_plan_reproduces_oracle(S):
    model = demonet_b200.ssd_lite_mobilenet_v2(image_size=S)
    sd = weights.seeded_state_dict(model.state_dict())
    _check_plan(model, sd, weights.synthetic_images(1, S), [0.485, 0.456, 0.406], [0.229, 0.224, 0.225],
                net_ref.v2_forward_raw)


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads on a CPU-only box and exports everything include/*.h declares."""
    header = open(os.path.join(ROOT, "include", "demonet_b200.h")).read()
    declared = set(re.findall(r"\b(dn_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for dt in _C.ACT_DTYPES:              # one library per activation storage type, same ABI
        handle = ctypes.CDLL(_C.LIB_PATHS[dt])
        for name in declared:
            assert hasattr(handle, name), "missing export in the %s build: %s" % (dt, name)
        assert _C.lib(dt).dn_abi_version() == _C.ABI_VERSION == 4
    assert declared == set(_C.EXPORTED_SYMBOLS)


def test_c_abi_argument_validation_without_gpu():
    lib = _C.lib()
    assert lib.dn_dwconv(None, None, None, None, 1, 8, 8, 8, 3, 1, 0, None) == _C.DN_ERR_INVALID
    assert b"NULL" in lib.dn_last_error()
    assert lib.dn_batched_nms_workspace_bytes(36000) > 36000 * 8
    with pytest.raises(RuntimeError):
        demonet_b200.ops.nms(torch.zeros(4, 4), torch.zeros(4), 0.5)       # CPU tensors: no fallback


def test_default_box_generator_table_cpu():
    from demonet_b200 import ops
    gen = ops.DefaultBoxGenerator([[2, 3]] * 6, min_ratio=0.2, max_ratio=0.95)
    grids = [(19, 19), (10, 10), (5, 5), (3, 3), (2, 2), (1, 1)]
    t = gen.table(grids, (300, 300), "cpu")
    assert np.array_equal(t.numpy(), boxes_np.default_boxes(grids, (300, 300)))
    assert gen.table(grids, (300, 300), "cpu") is t          # cached


def test_pixel_packing_rule():
    p = dplan.plan_ssdlite320_mobilenet_v3_large()
    packed = {i: dplan.pw_pack_factor(L) for i, L in enumerate(p.layers) if L.kind == "pw"}
    assert packed[2] == 4 and packed[3] == 4          # 160x160, 16 input channels
    assert all(f == 1 for i, f in packed.items() if p.layers[i].head)       # strided head outputs are never packed
    for i, f in packed.items():
        L = p.layers[i]
        assert (L.h_in * L.w_in) % f == 0 and (f == 1 or f * L.cout <= 512)


def test_stem_division_shortcut_is_exact():
    """stem_tma.cu replaces (x - mean) / std by q0 = RN(d * r), e = fma(-q0, std, d), q = fma(e, r, q0) with
    r = RN(1 / std).  Emulated here with exact float64 / long double intermediates: bit-identical to the
    float32 division the reference performs (transform.py:137)."""
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.random(400_000, dtype=np.float32), (np.arange(256) / 255).astype(np.float32),
                         (rng.standard_normal(100_000) * 3).astype(np.float32)])
    for m, s in [(0.485, 0.229), (0.456, 0.224), (0.406, 0.225), (0.5, 0.5), (0.3, 0.77777)]:
        m, s = np.float32(m), np.float32(s)
        r = np.float32(1.0 / np.float64(s))
        d = (xs - m).astype(np.float32)
        q0 = (d * r).astype(np.float32)
        e = np.float64(d) - np.float64(q0) * np.float64(s)
        e32 = e.astype(np.float32)
        assert np.all(e32.astype(np.float64) == e)                    # the remainder is exactly representable
        q = (np.longdouble(q0) + np.longdouble(e32) * np.longdouble(r)).astype(np.float32)
        assert np.array_equal(q.view(np.int32), (d / s).astype(np.float32).view(np.int32))


@pytest.mark.parametrize("make", [dplan.plan_ssdlite320_mobilenet_v3_large, lambda: dplan.plan_ssd_lite_mobilenet_v2(21, 512)])
def test_side_lanes_never_share_buffers(make):
    """Head branches run on side streams (plan.layer_lanes): every tensor they read or write must own its arena
    buffer for the whole forward, and the main chain must not depend on a side lane."""
    p = make()
    lanes = dplan.layer_lanes(p)
    t2b, bufs, lb, bb = dplan.assign_buffers(p)
    assert lanes.count(0) == len(p.layers) - sum(1 for L in p.layers if L.head) - sum(
        1 for i, L in enumerate(p.layers) if L.kind == "dw" and lanes[i])
    assert len({lanes[i] for i, L in enumerate(p.layers) if L.head}) == 12       # one lane per (head, level)
    producer = {L.dst: i for i, L in enumerate(p.layers) if L.dst and L.kind != "se"}
    touched = set()
    for i, L in enumerate(p.layers):
        if lanes[i]:
            touched.update(t for t in (L.src, L.res, L.dst) if t and t != "images")
            if L.src in producer:                       # a side lane reads the main chain or itself, never another lane
                assert lanes[producer[L.src]] in (0, lanes[i])
        else:
            assert all(lanes[producer[t]] == 0 for t in (L.src, L.res) if t in producer)
    owners = {}
    for t, b in t2b.items():
        owners.setdefault(b, []).append(t)
    for t in touched:
        assert owners[t2b[t]] == [t], "buffer of %s is shared: %s" % (t, owners[t2b[t]])
    assert set(p.feature_names) <= touched


def test_custom_ops_are_registered_with_fake_kernels():
    """the hot-path operators are torch.library ops; shape propagation needs no GPU, and there is no
    CPU kernel to fall back to."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from torch.fx.experimental.symbolic_shapes import ShapeEnv
    assert "Tensor cls_logits" in str(torch.ops.demonet_b200.postprocess.default._schema)
    with pytest.raises(NotImplementedError):
        torch.ops.demonet_b200.nms(torch.zeros(4, 4), torch.zeros(4), 0.5)
    with FakeTensorMode(shape_env=ShapeEnv()):
        lg, bb, an = torch.empty(2, 3234, 91, device="cuda"), torch.empty(2, 3234, 4, device="cuda"), torch.empty(3234, 4, device="cuda")
        boxes, scores, labels, counts = torch.ops.demonet_b200.postprocess(lg, bb, an, 320, 320, 0.001, 0.55, 300, 300, -1.0)
        assert tuple(boxes.shape) == (2, 300, 4) and tuple(scores.shape) == (2, 300) and labels.dtype == torch.int64
        assert tuple(counts.shape) == (2,) and counts.dtype == torch.int32
        keep = torch.ops.demonet_b200.batched_nms(torch.empty(10, 4, device="cuda"), torch.empty(10, device="cuda"),
                                                  torch.empty(10, dtype=torch.int64, device="cuda"), 0.5)
        assert keep.dtype == torch.int64 and keep.dim() == 1          # data-dependent length


def test_se_pool_slot_arithmetic():
    """The integer arithmetic shared by the pooling depthwise row streams (dwconv_stream*.cu, POOL: slot = share index -
    first share of the image), se_fc1_kernel (slots per image) and the launchers (slot bound): restated here and checked
    by brute force over random (batch, rows per image, CTA shares) -- every image's shares are contiguous, the formulas
    name the first and the last one, and their number never exceeds the bound the launcher sizes the workspace by."""
    import random
    rnd = random.Random(0)
    for _ in range(3000):
        B, H, parts = rnd.randint(1, 300), rnd.randint(1, 80), rnd.randint(1, 600)
        T = B * H
        parts = min(parts, T)
        owner = [None] * T
        for p in range(parts):
            for r in range(T * p // parts, T * (p + 1) // parts):
                owner[r] = p
        rmin = T // parts
        bound = (H + rmin - 1) // rmin + 1
        for b in range(B):
            first = ((b * H + 1) * parts + T - 1) // T - 1
            last = ((b + 1) * H * parts + T - 1) // T - 1
            assert first == owner[b * H] and last == owner[(b + 1) * H - 1]
            assert last - first + 1 <= bound


def test_detector_wrapper_scripts_without_a_gpu():
    """torch.jit.script of the scriptable wrapper needs no device: the graph holds ONE call of the registered operator."""
    import torch
    import demonet_b200
    from demonet_b200 import custom_ops
    model = demonet_b200.ssdlite320_mobilenet_v3_large(num_classes=91)
    scripted = torch.jit.script(custom_ops.ScriptableSSDLite(model))
    g = str(scripted.graph)
    assert g.count("demonet_b200::ssdlite_forward") == 1


def _dw_plan(H, W, C, k, s):
    out = (ctypes.c_int32 * 8)()
    assert _C.lib().dn_dwconv_plan_info(H, W, C, k, s, out) == 0, _C.lib().dn_last_error()
    return dict(zip(("impl", "CB", "ncb", "nstrip", "TW", "threads", "stages", "smem"), list(out)))


def test_depthwise_row_stream_plans():
    """Host logic of the depthwise planner (no GPU): invariants of every row-stream plan over a sweep of shapes, and the
    choices DESIGN.md quotes for the wide maps of V2 @ 512 (column strips) and for the V3 @ 320 layers (whole rows)."""
    STREAM, STREAM2 = 4, 8
    for H in (8, 10, 19, 20, 38, 40, 64, 75, 80, 128, 150, 160, 256):
        for C in (8, 16, 24, 32, 64, 72, 96, 120, 144, 184, 200, 240, 480, 672, 960):
            for k in (3, 5):
                for s in (1, 2):
                    p = _dw_plan(H, H, C, k, s)
                    if p["impl"] not in (STREAM, STREAM2):
                        continue
                    Wo = H if s == 1 else (H + 2 * (k // 2) - k) // 2 + 1
                    assert C % p["CB"] == 0 and p["CB"] % 8 == 0
                    assert p["nstrip"] >= 1 and p["ncb"] * p["TW"] * p["nstrip"] >= Wo            # the strips cover the width
                    assert (p["nstrip"] - 1) * p["ncb"] * p["TW"] < Wo                            # and none of them is empty
                    staged = p["ncb"] * p["TW"] + k - 1 if s == 1 else 2 * p["ncb"] * p["TW"] + k - 2
                    assert staged <= 256                                                          # TMA box limit
                    assert 64 <= p["threads"] <= 288 and p["stages"] >= 2 and 0 < p["smem"] <= 200 * 1024
                    if p["nstrip"] > 1:
                        assert Wo >= 64 and p["threads"] <= 160
    # config 5 (V2 @ 512): the four largest depthwise layers
    assert _dw_plan(256, 256, 32, 3, 1) | {"smem": 0, "stages": 0} == dict(impl=STREAM, CB=32, ncb=16, nstrip=4, TW=4, threads=160, stages=0, smem=0)
    p = _dw_plan(256, 256, 96, 3, 2)
    assert (p["impl"], p["CB"], p["nstrip"]) == (STREAM2, 32, 4)
    p = _dw_plan(128, 128, 144, 3, 2)
    assert (p["impl"], p["CB"], p["nstrip"]) == (STREAM2, 48, 4)
    p = _dw_plan(128, 128, 144, 3, 1)
    assert (p["impl"], p["CB"], p["nstrip"], p["threads"]) == (STREAM, 16, 1, 160)              # whole rows: more consumers win
    # config 2 (V3 @ 320): no strips anywhere, three-CTA-per-SM variants except the 184 / 200-channel layers
    for (H, C, k, s) in [(80, 72, 3, 1), (80, 72, 5, 2), (40, 120, 5, 1), (40, 240, 3, 2), (20, 480, 3, 1), (20, 672, 3, 1),
                         (20, 672, 5, 2), (10, 480, 5, 1), (10, 480, 3, 1)]:
        p = _dw_plan(H, H, C, k, s)
        assert p["impl"] == (STREAM if s == 1 else STREAM2) and p["nstrip"] == 1 and p["threads"] == 160, (H, C, k, s, p)
    assert _dw_plan(20, 20, 200, 3, 1)["threads"] == 288 and _dw_plan(5, 5, 512, 3, 1)["impl"] == 1


def _pw_plan(M, K, N):
    out = (ctypes.c_int32 * 8)()
    assert _C.lib().dn_pwconv_plan_info(M, K, N, out) == 0, _C.lib().dn_last_error()
    return dict(zip(("block_n", "n_tiles", "stages", "tmem_cols", "smem", "w_stat", "pair", "per_sm"), list(out)))


def test_gemm_tile_plans(monkeypatch):
    """Host logic of the GEMM planner (no GPU): every V3 / V2 / head shape gets a tile plan that fits the SM (shared memory,
    512 TMEM columns, at least two ring stages, at least one resident CTA); weight-stationary exactly for K <= 128; the pair
    modes only where asked for, only for ring-carried weights on large maps, with a deeper ring for cta_group::2."""
    shapes = [(256 * hw, k, n) for hw, k, n in [(25600, 16, 16), (6400, 64, 24), (6400, 24, 72), (6400, 72, 24), (1600, 72, 40),
                                                (1600, 40, 120), (1600, 120, 40), (1600, 40, 240), (400, 240, 80), (400, 80, 200),
                                                (400, 200, 80), (400, 80, 184), (400, 184, 80), (400, 80, 480), (400, 480, 112),
                                                (400, 112, 672), (400, 672, 112), (100, 672, 80), (100, 80, 480), (100, 480, 80),
                                                (100, 480, 256), (25, 256, 512), (25, 512, 128), (9, 128, 256), (4, 256, 64),
                                                (1, 64, 128), (400, 672, 546), (100, 480, 546), (25, 512, 546), (400, 672, 24),
                                                (1024, 96, 576), (256, 320, 1280), (256, 1280, 126)]]
    for M, K, N in shapes:
        p = _pw_plan(M, K, N)
        assert p["block_n"] % 16 == 0 and 16 <= p["block_n"] <= 256 and p["block_n"] * p["n_tiles"] >= N, (M, K, N, p)
        assert 2 * p["block_n"] <= p["tmem_cols"] <= 512 and p["tmem_cols"] & (p["tmem_cols"] - 1) == 0
        assert 2 <= p["stages"] <= 4 and 1 <= p["per_sm"] <= 4 and p["per_sm"] * p["tmem_cols"] <= 512
        assert 0 < p["smem"] <= 227 * 1024 and p["w_stat"] == (1 if K <= 128 else 0) and p["pair"] == 0
    assert _pw_plan(102400, 672, 546) | {"smem": 0} == dict(block_n=192, n_tiles=3, stages=4, tmem_cols=512, smem=0, w_stat=0, pair=0, per_sm=1)
    monkeypatch.setenv("DN_PW_PAIR", "1")
    assert _pw_plan(102400, 672, 546)["pair"] == 1 and _pw_plan(102400, 112, 672)["pair"] == 0        # weight-stationary: no pair
    assert _pw_plan(25600, 672, 80)["pair"] == 0                                                          # small map: no pair
    monkeypatch.setenv("DN_PW_PAIR", "2")
    p = _pw_plan(102400, 672, 546)
    assert p["pair"] == 2 and 4 < p["stages"] <= 8 and p["smem"] <= 227 * 1024                            # half-size weight stages
