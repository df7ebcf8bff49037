"""The operators ssd300_vgg16 adds against plain fp32 PyTorch on the same 16-bit inputs, both builds:
dense 3x3 convolution on the tensor cores (stride 1 / 2, padding 0 / 1, dilation 6, ragged maps, the head's fp32 strided
output), the first 3-channel convolution with the normalisation folded in, max-pooling (ceil mode, the 3x3 s1 pool5) and
the L2-normalisation + scale of conv4_3."""
import pytest
import torch
import torch.nn.functional as F

from demonet_b200 import ops

pytestmark = [pytest.mark.gpu, pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])]
ULP = {torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}


def _close(dt, got, want, extra_abs):
    got, want = got.float(), want.float()
    bad = (got - want).abs() > want.abs() * ULP[dt] + extra_abs
    assert not bool(bad.any()), "max err %g at %d of %d elements" % (float((got - want).abs().max()), int(bad.sum()), bad.numel())


# (B, H, W, C, N, stride, pad, dil): VGG / SSD300 layer shapes (small batches) + ragged ones
CONV_CASES = [(2, 150, 150, 64, 128, 1, 1, 1), (2, 75, 75, 128, 256, 1, 1, 1), (3, 38, 38, 256, 512, 1, 1, 1),
              (3, 19, 19, 512, 1024, 1, 6, 6), (3, 19, 19, 256, 512, 2, 1, 1), (5, 10, 10, 128, 256, 2, 1, 1),
              (7, 5, 5, 128, 256, 1, 0, 1), (9, 3, 3, 128, 256, 1, 0, 1), (2, 38, 38, 512, 364, 1, 1, 1),
              (2, 19, 19, 1024, 546, 1, 1, 1), (40, 1, 1, 256, 16, 1, 1, 1), (3, 13, 29, 64, 24, 1, 1, 1),
              (2, 300, 300, 64, 64, 1, 1, 1), (130, 3, 3, 256, 364, 1, 1, 1),
              (2, 20, 9, 64, 40, 1, 1, 1), (1, 8, 8, 128, 16, 1, 1, 1), (3, 77, 51, 64, 96, 1, 1, 1), (2, 10, 10, 512, 546, 1, 1, 1)]


@pytest.mark.parametrize("B,H,W,C,N,s,p,d", CONV_CASES)
def test_conv3x3_tensor_core(B, H, W, C, N, s, p, d, dt):
    g = torch.Generator().manual_seed(B + H * 3 + C + N)
    x = torch.randn(B, H, W, C, generator=g).to(dt).cuda()
    w = (torch.randn(N, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(dt).cuda()
    b = torch.randn(N, generator=g).cuda()
    wt = w.permute(2, 3, 0, 1).reshape(9, N, C).contiguous()
    y = ops.conv3x3(x, wt, b, s, p, d, "relu")
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, s, p, d)).permute(0, 2, 3, 1)
    assert y.shape == ref.shape and y.dtype == dt
    _close(dt, y, ref, 2e-3)
    y2 = ops.conv3x3(x, wt, b, s, p, d, "none")
    _close(dt, y2, F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, s, p, d).permute(0, 2, 3, 1), 2e-3)


def test_conv3x3_head_addressing(dt):
    """fp32 output written where SSDScoringHead's view / permute / reshape / cat puts it (generalized_ssd.py:60-74)."""
    g = torch.Generator().manual_seed(5)
    B, H, C, A, K, P, off = 3, 10, 512, 6, 91, 8732, 7942
    N = A * K
    x = torch.randn(B, H, H, C, generator=g).to(dt).cuda()
    w = (torch.randn(N, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(dt).cuda()
    b = torch.randn(N, generator=g).cuda()
    out = torch.zeros(B, P, K, device="cuda")
    ops.conv3x3(x, w.permute(2, 3, 0, 1).reshape(9, N, C).contiguous(), b, 1, 1, 1, "none", out=out[:, off:], out_batch_stride=P * K,
                out_row_stride=N)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, 1, 1)
    want = ref.view(B, A, K, H, H).permute(0, 3, 4, 1, 2).reshape(B, -1, K)
    assert float((out[:, off:off + H * H * A] - want).abs().max()) < 3e-3
    assert float(out[:, :off].abs().sum()) == 0 and float(out[:, off + H * H * A:].abs().sum()) == 0


@pytest.mark.parametrize("S", [300, 37])
def test_first_conv_with_normalisation(S, dt):
    g = torch.Generator().manual_seed(S)
    img = torch.rand(2, 3, S, S, generator=g).cuda()
    w = (torch.randn(64, 3, 3, 3, generator=g) * 0.05).cuda()
    b = torch.randn(64, generator=g).cuda()
    mean, std = [0.48235, 0.45882, 0.40784], [1.0 / 255.0] * 3             # ssd_vgg16.py:198-199
    y = ops.conv3x3_first(img, w.permute(1, 2, 3, 0).reshape(27, 64).contiguous(), b, mean, std,
                          act_dtype={torch.float16: "fp16", torch.bfloat16: "bf16"}[dt])
    m = torch.tensor(mean, device="cuda")[None, :, None, None]
    sd = torch.tensor(std, device="cuda")[None, :, None, None]
    ref = F.relu(F.conv2d((img - m) / sd, w, b, 1, 1)).permute(0, 2, 3, 1)
    _close(dt, y, ref, 2e-2)                                              # inputs are ~ +-128: fp32 summation noise ~ 1e-3 relative


@pytest.mark.parametrize("S", [300, 37])
def test_first_conv_as_im2col_plus_gemm(S, dt):
    """The tensor-core form of the first layer: the im2col rows are the normalised taps of F.unfold bit for bit (after the
    one rounding to the activation type), and rows x the 32-wide K-major filter through dn_pwconv is the convolution."""
    g = torch.Generator().manual_seed(S + 1)
    img = torch.rand(2, 3, S, S, generator=g).cuda()
    w = (torch.randn(64, 3, 3, 3, generator=g) * 0.05).cuda()
    b = torch.randn(64, generator=g).cuda()
    mean, std = [0.48235, 0.45882, 0.40784], [1.0 / 255.0] * 3
    name = {torch.float16: "fp16", torch.bfloat16: "bf16"}[dt]
    cols = ops.im2col3x3_first(img, mean, std, act_dtype=name)
    assert cols.shape == (2 * S * S, 32) and cols.dtype == dt
    m = torch.tensor(mean, device="cuda")[None, :, None, None]
    rs = (1.0 / torch.tensor(std, dtype=torch.float32, device="cuda"))[None, :, None, None]          # fp32 division, like the entry
    norm = (img - m) * rs                                                  # the kernel's own two fp32 roundings
    want = F.unfold(norm, 3, padding=1).view(2, 27, S * S).permute(0, 2, 1).reshape(-1, 27).to(dt)
    assert torch.equal(cols[:, :27], want) and float(cols[:, 27:].abs().sum()) == 0
    w32 = torch.zeros(64, 32, device="cuda")
    w32[:, :27] = w.reshape(64, 27)
    y = ops.pwconv(cols, w32.to(dt), b, act="relu").view(2, S, S, 64)
    sd = torch.tensor(std, device="cuda")[None, :, None, None]
    ref = F.relu(F.conv2d((img - m) / sd, w, b, 1, 1)).permute(0, 2, 3, 1)
    _close(dt, y, ref, 6e-2 if dt == torch.float16 else 0.5)               # inputs ~ +-128 rounded to the activation type


@pytest.mark.parametrize("H,W,C,k,s,p,ceil", [(300, 300, 64, 2, 2, 0, False), (75, 75, 256, 2, 2, 0, True), (19, 19, 512, 3, 1, 1, False),
                                              (38, 38, 512, 2, 2, 0, False), (7, 9, 8, 2, 2, 0, True)])
def test_maxpool(H, W, C, k, s, p, ceil, dt):
    g = torch.Generator().manual_seed(H + C)
    x = torch.randn(2, H, W, C, generator=g).to(dt).cuda()
    y = ops.maxpool2d(x, k, s, p, ceil)
    ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), k, s, p, ceil_mode=ceil).permute(0, 2, 3, 1)
    assert y.shape == ref.shape and torch.equal(y.float(), ref)


def test_l2norm_scale(dt):
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(2, 38, 38, 512, generator=g) * 3).to(dt).cuda()
    x[0, 0, 0] = 0                                                         # a zero vector: eps path of F.normalize
    scale = (torch.rand(512, generator=g) * 20 + 5).cuda()
    y = ops.l2norm_scale(x, scale)
    ref = (scale.view(1, -1, 1, 1) * F.normalize(x.float().permute(0, 3, 1, 2))).permute(0, 2, 3, 1)
    _close(dt, y, ref, 1e-4)
