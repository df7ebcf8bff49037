"""Each CUDA kernel against a plain PyTorch fp32 reference of the same op on the same 16-bit inputs
(C-ABI stage entry points, shapes from SURVEY.md Appendix C), for both builds of the library: fp16
activation storage (the default) and bf16.  Tolerance: the output is rounded to the storage type, so
the bar is 1 ulp of the reference value (2^-10 relative for fp16, 2^-7 for bf16) plus fp32 summation noise."""
import pytest
import torch
import torch.nn.functional as F

from demonet_b200 import ops

pytestmark = [pytest.mark.gpu, pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])]
ULP = {torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}
ACTS = {"none": lambda v: v, "relu": F.relu, "relu6": F.relu6, "hardswish": F.hardswish}


def _close_h16(dt, got, want, extra_abs=1e-3):
    assert got.dtype == dt
    got, want = got.float(), want.float()
    tol = want.abs() * ULP[dt] + extra_abs
    bad = (got - want).abs() > tol
    assert not bool(bad.any()), "max err %g at %d elements" % (float((got - want).abs().max()), int(bad.sum()))


# (C, k, stride, H) from Appendix C.1 / C.3 plus ragged sizes
DW_CASES = [(16, 3, 1, 160), (64, 3, 2, 160), (72, 5, 2, 80), (120, 5, 1, 40), (240, 3, 2, 40), (672, 3, 1, 20),
            (672, 5, 2, 20), (480, 5, 1, 10), (256, 3, 2, 10), (128, 3, 2, 5), (128, 3, 2, 3), (64, 3, 2, 2),
            (128, 3, 1, 1), (96, 3, 1, 19), (144, 3, 2, 75), (32, 3, 1, 150), (1280, 3, 1, 16), (8, 5, 2, 7),
            (200, 3, 1, 20), (72, 3, 1, 80), (24, 5, 1, 33), (8, 3, 1, 9), (184, 3, 1, 20),
            # wide maps: the row streams cut the width into column strips (config 5's largest layers, V2 @ 300, ragged widths)
            (32, 3, 1, 256), (96, 3, 2, 256), (144, 3, 1, 128), (144, 3, 2, 128), (96, 3, 2, 150), (24, 5, 1, 70),
            (48, 5, 2, 131), (16, 3, 1, 97)]


@pytest.mark.parametrize("C,k,s,H", DW_CASES)
@pytest.mark.parametrize("act", ["relu6", "hardswish"])
def test_dwconv(C, k, s, H, act, dt):
    g = torch.Generator().manual_seed(C * 100 + k * 10 + s + H)
    B = 3
    x = (torch.randn(B, H, H, C, generator=g) * 2).to(dt).cuda()
    w = (torch.randn(k * k, C, generator=g) / k).cuda()
    b = torch.randn(C, generator=g).cuda()
    y = ops.dwconv(x, w, b, k, s, act)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.t().reshape(C, 1, k, k), b, s, (k - 1) // 2, 1, C)
    ref = ACTS[act](ref).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    _close_h16(dt, y, ref)


@pytest.mark.parametrize("B,H,W,C,k,s", [(37, 12, 27, 40, 3, 1), (37, 12, 27, 40, 5, 1), (300, 8, 8, 64, 3, 1), (5, 64, 9, 16, 5, 1),
                                         (37, 12, 27, 40, 3, 2), (37, 13, 27, 40, 5, 2), (300, 8, 8, 64, 3, 2), (5, 64, 9, 16, 5, 2),
                                         (64, 80, 80, 72, 5, 2), (9, 33, 47, 24, 3, 2)])
def test_dwconv_rect_many_images(B, H, W, C, k, s, dt):
    """Non-square maps and enough images that a CTA's share of the row stream (dwconv_stream.cu, dwconv_stream2.cu)
    starts and ends in the middle of images and spans several of them."""
    g = torch.Generator().manual_seed(B + H * 3 + W * 5 + C * 7 + k + s)
    x = (torch.randn(B, H, W, C, generator=g) * 2).to(dt).cuda()
    w = (torch.randn(k * k, C, generator=g) / k).cuda()
    b = torch.randn(C, generator=g).cuda()
    y = ops.dwconv(x, w, b, k, s, "hardswish")
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.t().reshape(C, 1, k, k), b, s, (k - 1) // 2, 1, C)
    ref = ACTS["hardswish"](ref).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    _close_h16(dt, y, ref)


@pytest.mark.parametrize("B,H,W,C,k,s", [(20, 9, 130, 48, 3, 1), (20, 17, 131, 48, 5, 2), (3, 256, 256, 32, 3, 1), (3, 128, 128, 144, 3, 2),
                                         (7, 12, 70, 24, 5, 1)])
def test_dwconv_column_strips_bit_identical(B, H, W, C, k, s, dt, monkeypatch):
    """Column strips (wide maps) only change which CTA computes an output, not the order of its multiply-adds: the result
    equals the whole-row plan bit for bit where that plan exists, and fp32 torch within 1 ulp either way."""
    g = torch.Generator().manual_seed(B + H * 3 + W * 5 + C * 7 + k + s)
    x = (torch.randn(B, H, W, C, generator=g) * 2).to(dt).cuda()
    w = (torch.randn(k * k, C, generator=g) / k).cuda()
    b = torch.randn(C, generator=g).cuda()
    y = ops.dwconv(x, w, b, k, s, "relu6")
    monkeypatch.setenv("DN_DW_STRIPS", "0")
    y0 = ops.dwconv(x, w, b, k, s, "relu6")
    monkeypatch.delenv("DN_DW_STRIPS")
    assert torch.equal(y, y0)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.t().reshape(C, 1, k, k), b, s, (k - 1) // 2, 1, C)
    _close_h16(dt, y, F.relu6(ref).permute(0, 2, 3, 1))


# (M, K, N): Appendix C GEMM shapes (per-image M times a small batch), incl. N not multiple of 16 / > 256
PW_CASES = [(25600, 16, 16), (25600, 16, 64), (6400, 64, 24), (6400, 24, 72), (1600, 72, 40), (1600, 40, 240),
            (400, 80, 184), (400, 200, 80), (400, 112, 672), (400, 672, 112), (400, 672, 546), (400, 672, 24),
            (100, 480, 546), (100, 480, 256), (25, 512, 546), (9, 256, 128), (4, 256, 64), (1, 128, 546),
            (1024, 96, 576), (256, 320, 1280), (256, 1280, 126), (1, 64, 24), (130, 8, 8), (257, 200, 1000)]


@pytest.mark.parametrize("M,K,N,res,fp32", [(40960, 672, 546, False, True), (38017, 200, 80, True, False), (40960, 480, 112, False, False),
                                            (51200, 240, 1000, False, False)])
@pytest.mark.parametrize("mode", [1, 2])
def test_pwconv_pair_mode_is_bit_identical(M, K, N, res, fp32, mode, dt, monkeypatch):
    """PAIR modes compute exactly what the single-CTA kernel computes (same products, same accumulation order along K).
    DN_PW_PAIR=1: clusters of two CTAs, each loads half of every weight k-block and multicasts it into both CTAs' ring stage,
    stages released by both CTAs' MMAs.  DN_PW_PAIR=2: tcgen05 cta_group::2 -- the pair's leader issues 256 x N UMMAs, each CTA
    keeps its 128 A rows and its half of B, both CTAs' loads complete on the leader's barrier, commits are multicast.
    Odd M-tile counts (a cluster whose second CTA has no rows), residual / fp32 / multi-N-tile outputs."""
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).to(dt).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dt).cuda()
    b = torch.randn(N, generator=g).cuda()
    r = torch.randn(M, N, generator=g).to(dt).cuda() if res else None
    y0 = ops.pwconv(x, w, b, "relu6" if not res and not fp32 else "none", r, fp32)
    monkeypatch.setenv("DN_PW_PAIR", str(mode))
    y1 = ops.pwconv(x, w, b, "relu6" if not res and not fp32 else "none", r, fp32)
    monkeypatch.delenv("DN_PW_PAIR")
    assert torch.equal(y0, y1)
    ref = x.float() @ w.float().t() + b
    if res:
        ref = ref + r.float()
    elif not fp32:
        ref = F.relu6(ref)
    if fp32:
        assert float((y1 - ref).abs().max()) < 1e-3
    else:
        _close_h16(dt, y1, ref, extra_abs=2e-3)


@pytest.mark.parametrize("M,K,N", PW_CASES)
def test_pwconv_simt(M, K, N, dt):
    _check_pwconv(M, K, N, 1, dt)


@pytest.mark.parametrize("M,K,N", PW_CASES)
def test_pwconv_tc(M, K, N, dt):
    _check_pwconv(M, K, N, 0, dt)


def _check_pwconv(M, K, N, impl, dt):
    g = torch.Generator().manual_seed(M + K * 7 + N * 13)
    batch = 3 if M < 30000 else 1
    Mt = M * batch
    x = torch.randn(Mt, K, generator=g).to(dt).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dt).cuda()
    b = torch.randn(N, generator=g).cuda()
    ref = x.float() @ w.float().t() + b
    for act, res, fp32 in (("none", False, False), ("hardswish", False, False), ("none", True, False), ("none", False, True),
                           ("relu6", False, False)):
        r = torch.randn(Mt, N, generator=g).to(dt).cuda() if res else None
        y = ops.pwconv(x, w, b, act, r, fp32, impl)
        want = ACTS[act](ref) + (r.float() if res else 0)
        if fp32:
            assert y.dtype == torch.float32
            assert float((y - want).abs().max()) < 2e-3
        else:
            _close_h16(dt, y, want)


def test_pwconv_tc_equals_simt_large(dt):
    """tcgen05 kernel vs the independent SIMT kernel at a full config-2 layer size (B=256, cls head L0)."""
    g = torch.Generator().manual_seed(0)
    M, K, N = 256 * 400, 672, 546
    x = torch.randn(M, K, generator=g).to(dt).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dt).cuda()
    b = torch.randn(N, generator=g).cuda()
    y0 = ops.pwconv(x, w, b, "none", None, True, 0)
    y1 = ops.pwconv(x, w, b, "none", None, True, 1)
    assert float((y0 - y1).abs().max()) < 1e-3


def test_pwconv_head_addressing(dt):
    """Strided fp32 output: row (b, hw) of a level lands at b*P*K + (off + hw*A)*K  (generalized_ssd.py:66-74)."""
    from demonet_b200 import _C
    g = torch.Generator().manual_seed(1)
    B, HW, K, A, cols, P, off = 3, 25, 512, 6, 91, 3234, 3000
    N = A * cols
    x = torch.randn(B * HW, K, generator=g).to(dt).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dt).cuda()
    b = torch.randn(N, generator=g).cuda()
    out = torch.zeros(B, P, cols, device="cuda")
    base = out.data_ptr() + off * cols * 4
    lib = _C.lib(_C.dtype_name(dt))
    _C.check(lib.dn_pwconv(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, base, B * HW, K, N, 0, 1, HW, P * cols,
                           N, 0, torch.cuda.current_stream().cuda_stream), lib)
    ref = (x.float() @ w.float().t() + b).view(B, HW * A, cols)
    assert float((out[:, off:off + HW * A] - ref).abs().max()) < 2e-3
    assert float(out[:, :off].abs().sum()) == 0 and float(out[:, off + HW * A:].abs().sum()) == 0


@pytest.mark.parametrize("impl", ["tc", "simt"])          # tensor-core im2col GEMM (default) / fp32 SIMT tiles (DN_STEM=simt)
@pytest.mark.parametrize("Cout,S,act,B", [(16, 320, "hardswish", 3), (32, 300, "relu6", 3), (32, 512, "relu6", 3), (16, 33, "hardswish", 3),
                                          (16, 324, "relu", 3), (32, 36, "hardswish", 3),
                                          (16, 320, "hardswish", 40), (32, 300, "relu6", 37)])      # many tiles per CTA: the window ring wraps
def test_stem(Cout, S, act, B, dt, impl, monkeypatch):
    monkeypatch.setenv("DN_STEM", impl)
    g = torch.Generator().manual_seed(S)
    img = torch.rand(B, 3, S, S, generator=g).cuda()
    w = (torch.randn(Cout, 3, 3, 3, generator=g) * 0.3).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    y = ops.stem_conv(img, w.permute(1, 2, 3, 0).reshape(27, Cout).contiguous(), b, mean, std, act,
                      act_dtype={torch.float16: "fp16", torch.bfloat16: "bf16"}[dt])
    m = torch.tensor(mean, device="cuda")[None, :, None, None]
    s = torch.tensor(std, device="cuda")[None, :, None, None]
    ref = ACTS[act](F.conv2d((img - m) / s, w, b, 2, 1)).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    _close_h16(dt, y, ref)


@pytest.mark.parametrize("C,Cs,HW", [(72, 24, 1600), (120, 32, 1600), (480, 120, 400), (672, 168, 400), (672, 168, 100),
                                     (480, 120, 100), (16, 8, 9)])
def test_se(C, Cs, HW, dt):
    g = torch.Generator().manual_seed(C + HW)
    B = 3
    x = (torch.randn(B, HW, C, generator=g)).to(dt).cuda()
    w1 = (torch.randn(Cs, C, generator=g) / C ** 0.5).cuda()
    b1 = torch.randn(Cs, generator=g).cuda() * 0.5
    w2 = (torch.randn(C, Cs, generator=g) / Cs ** 0.5).cuda()
    b2 = torch.randn(C, generator=g).cuda() * 0.5
    xf = x.float()
    scale = F.hardsigmoid(F.relu(xf.mean(1) @ w1.t() + b1) @ w2.t() + b2)
    ref = xf * scale[:, None, :]
    y = ops.se_inplace(x.clone(), w1, b1, w2.t().contiguous(), b2)
    _close_h16(dt, y, ref, extra_abs=2e-3)


@pytest.mark.parametrize("B,C,Cs,HW", [(3, 72, 24, 1600), (37, 480, 120, 400), (64, 672, 168, 100), (17, 16, 8, 9), (50, 120, 32, 64)])
def test_se_cluster_form_is_bit_identical(B, C, Cs, HW, dt, monkeypatch):
    """fc1 + fc2 as ONE launch (thread-block clusters of eight CTAs, hidden activations exchanged through distributed shared
    memory) against the two-launch form: same arithmetic, same order, identical bits; several image groups, a ragged last one."""
    g = torch.Generator().manual_seed(B + C + HW)
    x = (torch.randn(B, HW, C, generator=g)).to(dt).cuda()
    w1 = (torch.randn(Cs, C, generator=g) / C ** 0.5).cuda()
    b1 = torch.randn(Cs, generator=g).cuda() * 0.5
    w2t = (torch.randn(Cs, C, generator=g) / Cs ** 0.5).cuda()
    b2 = torch.randn(C, generator=g).cuda() * 0.5
    y0 = ops.se_inplace(x.clone(), w1, b1, w2t, b2)
    monkeypatch.setenv("DN_SE_CLUSTER", "1")
    y = ops.se_inplace(x.clone(), w1, b1, w2t, b2)
    monkeypatch.delenv("DN_SE_CLUSTER")
    assert torch.equal(y, y0)
    xf = x.float()
    scale = F.hardsigmoid(F.relu(xf.mean(1) @ w1.t() + b1) @ w2t + b2)
    _close_h16(dt, y, xf * scale[:, None, :], extra_abs=2e-3)


@pytest.mark.parametrize("B,H,W,C,Cs,k,stride,act,want_pooled", [
    (64, 40, 40, 120, 32, 5, 1, "relu", True),            # V3 blocks 4-5 at batch size
    (48, 20, 20, 480, 120, 3, 1, "hardswish", True),      # block 10
    (37, 20, 20, 672, 168, 3, 1, "hardswish", True),      # block 11; CTA shares cut the images at irregular rows
    (96, 24, 22, 40, 16, 3, 1, "relu", True),             # width not a multiple of the 4-column thread tile
    (16, 10, 10, 480, 120, 5, 1, "hardswish", True),      # blocks 13-14
    (2, 10, 10, 480, 120, 5, 1, "hardswish", True),       # one row per CTA share, ten shares per image
    (32, 80, 80, 72, 24, 5, 2, "relu", True),             # block 3 (stride-2 row stream)
    (48, 20, 20, 672, 168, 5, 2, "hardswish", True),      # block 12 (stride-2 row stream, 2 columns per thread)
    (61, 40, 40, 240, 64, 3, 2, "hardswish", True),       # stride 2, k 3, odd batch
    (3, 40, 40, 120, 32, 5, 1, "relu", False),            # small batch: > 16 shares per image, SE pools by itself
    (8, 40, 40, 240, 64, 3, 2, "hardswish", False),       # the same on the stride-2 stream
    (16, 6, 6, 480, 120, 5, 1, "hardswish", False),       # maps under 8 rows run on the direct kernel
])
def test_dwconv_se_pooled_by_the_row_stream(B, H, W, C, Cs, k, stride, act, want_pooled, dt):
    """dn_dwconv_se (depthwise + SE; the row stream leaves the SE channel sums of the values it stores) against the two
    separate calls and against fp32 PyTorch.  Only the order of the fp32 summation differs from the separate pooling
    pass, and sums of bf16 values are all but exact in fp32: the two results agree bit for bit nearly everywhere."""
    g = torch.Generator().manual_seed(B * 1000 + C)
    x = torch.randn(B, H, W, C, generator=g).to(dt).cuda()
    w = (torch.randn(C, k, k, generator=g) * 0.3).cuda()
    b = torch.randn(C, generator=g).cuda() * 0.5
    w1 = (torch.randn(Cs, C, generator=g) / C ** 0.5).cuda()
    b1 = torch.randn(Cs, generator=g).cuda() * 0.5
    w2 = (torch.randn(C, Cs, generator=g) / Cs ** 0.5).cuda()
    b2 = torch.randn(C, generator=g).cuda() * 0.5
    wk = w.permute(1, 2, 0).reshape(k * k, C).contiguous()
    w2t = w2.t().contiguous()
    y, pooled = ops.dwconv_se(x, wk, b, k, stride, act, w1, b1, w2t, b2)
    assert pooled == want_pooled
    mid = ops.dwconv(x, wk, b, k, stride, act)
    two = ops.se_inplace(mid.clone().view(B, -1, C), w1, b1, w2t, b2).view_as(mid)
    assert y.shape == two.shape
    d = (y.float() - two.float()).abs()
    assert float((d / two.float().abs().clamp_min(1e-2)).max()) <= ULP[dt], "more than one ulp from the unfused pair"
    assert float((d > 0).float().mean()) < 1e-3                      # and nearly all elements identical
    conv = ACTS[act](F.conv2d(x.float().permute(0, 3, 1, 2), w[:, None], b, stride, (k - 1) // 2, 1, C))
    scale = F.hardsigmoid(F.relu(conv.mean((2, 3)) @ w1.t() + b1) @ w2.t() + b2)
    ref = (conv * scale[:, :, None, None]).permute(0, 2, 3, 1)
    _close_h16(dt, y, ref, extra_abs=4e-3)
    # deterministic: fixed summation order, no atomics
    y2, _ = ops.dwconv_se(x, wk, b, k, stride, act, w1, b1, w2t, b2)
    assert torch.equal(y, y2)


@pytest.mark.parametrize("B,H,W,act", [(3, 160, 160, "relu"), (2, 40, 40, "relu6"), (5, 50, 38, "hardswish"), (1, 17, 9, "relu"),
                                       (40, 32, 32, "relu")])
def test_pwdw_fused_matches_two_kernels(B, H, W, act, dt):
    """Fused expand (16 -> 64) + depthwise 3x3 s2 against the unfused pair: the expand GEMM is the same tensor-core
    product with the same single rounding, so only the fp32 summation order of the stencil may differ (<= 1 bf16 ulp);
    and against a plain fp32 PyTorch evaluation of the two convolutions."""
    g = torch.Generator().manual_seed(B * 7 + H + W)
    K, N = 16, 64
    x = (torch.randn(B, H, W, K, generator=g)).to(dt).cuda()
    w_pw = (torch.randn(N, K, generator=g) / 4).to(dt).cuda()
    b_pw = torch.randn(N, generator=g).cuda()
    w_dw = (torch.randn(9, N, generator=g) / 3).cuda()
    b_dw = torch.randn(N, generator=g).cuda()
    y = ops.pwdw_fused(x, w_pw, b_pw, w_dw, b_dw, 3, 2, act, act)
    mid = ops.pwconv(x.reshape(-1, K), w_pw, b_pw, act).reshape(B, H, W, N)
    two = ops.dwconv(mid, w_dw, b_dw, 3, 2, act)
    assert y.shape == two.shape
    _close_h16(dt, y, two.float())
    mid_ref = ACTS[act](F.conv2d(x.float().permute(0, 3, 1, 2), w_pw.float().reshape(N, K, 1, 1), b_pw)).to(dt).float()
    ref = ACTS[act](F.conv2d(mid_ref, w_dw.t().reshape(N, 1, 3, 3), b_dw, 2, 1, 1, N)).permute(0, 2, 3, 1)
    _close_h16(dt, y, ref, extra_abs=2e-2)


@pytest.mark.parametrize("B,H,W,act,res", [(3, 160, 160, "relu", True), (2, 40, 48, "relu6", False), (5, 50, 38, "hardswish", True),
                                           (1, 9, 17, "relu", True), (70, 16, 16, "relu", True)])
def test_dwpw_fused_matches_two_kernels(B, H, W, act, res, dt):
    """Fused depthwise 3x3 s1 (16 ch) + project 16 -> 16 (+ residual) against the unfused pair and against fp32 PyTorch."""
    g = torch.Generator().manual_seed(B * 11 + H + W)
    C = 16
    x = (torch.randn(B, H, W, C, generator=g)).to(dt).cuda()
    w_dw = (torch.randn(9, C, generator=g) / 3).cuda()
    b_dw = torch.randn(C, generator=g).cuda()
    w_pw = (torch.randn(C, C, generator=g) / 4).to(dt).cuda()
    b_pw = torch.randn(C, generator=g).cuda()
    y = ops.dwpw_fused(x, w_dw, b_dw, w_pw, b_pw, 3, 1, act, res)
    mid = ops.dwconv(x, w_dw, b_dw, 3, 1, act)
    two = ops.pwconv(mid.reshape(-1, C), w_pw, b_pw, "none", residual=x.reshape(-1, C) if res else None).reshape(B, H, W, C)
    _close_h16(dt, y, two.float(), extra_abs=2e-2)
    mid_ref = ACTS[act](F.conv2d(x.float().permute(0, 3, 1, 2), w_dw.t().reshape(C, 1, 3, 3), b_dw, 1, 1, 1, C)).to(dt).float()
    ref = F.conv2d(mid_ref, w_pw.float().reshape(C, C, 1, 1), b_pw)
    if res:
        ref = ref + x.float().permute(0, 3, 1, 2)
    _close_h16(dt, y, ref.permute(0, 2, 3, 1), extra_abs=2e-2)


@pytest.mark.parametrize("B,HW,C,Cs,N,res", [(3, 1600, 120, 32, 40, True), (3, 1600, 72, 24, 40, False), (5, 400, 480, 120, 112, False),
                                             (4, 400, 672, 168, 112, True), (7, 100, 672, 168, 80, False), (6, 100, 480, 120, 80, True),
                                             (2, 9, 16, 8, 24, False), (33, 130, 40, 16, 264, True)])
def test_se_folded_into_the_project_gemm(B, HW, C, Cs, N, res, dt):
    """dn_se_project: the squeeze-excitation scaling applied to the GEMM's A operand in shared memory (pwconv_tc.cu,
    SCALE_A) must equal the in-place scaling pass followed by the plain GEMM BIT FOR BIT (same products, same single
    rounding of x * scale to the storage type), leave x untouched, and match fp32 PyTorch within one ulp."""
    g = torch.Generator().manual_seed(B * 100 + C + N)
    x = torch.randn(B, HW, C, generator=g).to(dt).cuda()
    w1 = (torch.randn(Cs, C, generator=g) / C ** 0.5).cuda()
    b1 = torch.randn(Cs, generator=g).cuda() * 0.5
    w2 = (torch.randn(C, Cs, generator=g) / Cs ** 0.5).cuda()
    b2 = torch.randn(C, generator=g).cuda() * 0.5
    w_pw = (torch.randn(N, C, generator=g) / C ** 0.5).to(dt).cuda()
    b_pw = torch.randn(N, generator=g).cuda()
    r = torch.randn(B * HW, N, generator=g).to(dt).cuda() if res else None
    w2t = w2.t().contiguous()
    x0 = x.clone()
    y = ops.se_project(x, w1, b1, w2t, b2, w_pw, b_pw, r)
    assert torch.equal(x, x0)                                           # the scaled tensor never exists in HBM
    xs = ops.se_inplace(x.clone(), w1, b1, w2t, b2)
    two = ops.pwconv(xs.view(B * HW, C), w_pw, b_pw, "none", r)
    assert torch.equal(y, two)
    xf = x.float()
    scale = F.hardsigmoid(F.relu(xf.mean(1) @ w1.t() + b1) @ w2.t() + b2)
    ref = (xf * scale[:, None, :]).to(dt).float().view(B * HW, C) @ w_pw.float().t() + b_pw + (r.float() if res else 0)
    _close_h16(dt, y, ref, extra_abs=4e-3)
