"""Diagnostic: dn_dwconv_se (pooled by the row stream) against dn_dwconv + dn_se_inplace, repeated, at small batches."""
import sys
import torch
from demonet_b200 import ops

def case(B, H, W, C, Cs, k, act, reps=10):
    g = torch.Generator().manual_seed(B * 1000 + C)
    x = torch.randn(B, H, W, C, generator=g).bfloat16().cuda()
    w = (torch.randn(C, k, k, generator=g) * 0.3).cuda()
    b = torch.randn(C, generator=g).cuda() * 0.5
    w1 = (torch.randn(Cs, C, generator=g) / C ** 0.5).cuda()
    b1 = torch.randn(Cs, generator=g).cuda() * 0.5
    w2 = (torch.randn(C, Cs, generator=g) / Cs ** 0.5).cuda()
    b2 = torch.randn(C, generator=g).cuda() * 0.5
    wk = w.permute(1, 2, 0).reshape(k * k, C).contiguous()
    w2t = w2.t().contiguous()
    mid = ops.dwconv(x, wk, b, k, 1, act)
    two = ops.se_inplace(mid.clone().view(B, -1, C), w1, b1, w2t, b2).view_as(mid).float()
    worst, nbad, pooled = 0.0, 0, None
    for r in range(reps):
        y, pooled = ops.dwconv_se(x, wk, b, k, 1, act, w1, b1, w2t, b2)
        d = (y.float() - two).abs() / two.abs().clamp_min(1e-2)
        worst = max(worst, float(d.max()))
        bad = d > 2.0 ** -7 * 1.01
        nbad += int(bad.sum())
        if bool(bad.any()) and r == 0:
            idx = bad.nonzero()
            print("   bad (b, ch) pairs:", sorted(set((int(i[0]), int(i[3])) for i in idx))[:12])
    print("B=%d %dx%d C=%d k%d pooled=%s: worst rel diff %.4g (1 ulp = %.4g), elements beyond 1 ulp over %d reps: %d, differing at all: %.4f"
          % (B, H, W, C, k, pooled, worst, 2.0 ** -7, reps, nbad, float((y.float() != two).float().mean())))

for B in (1, 2, 3, 7, 16):
    case(B, 10, 10, 480, 120, 5, "hardswish")
for B in (7, 12, 48):
    case(B, 20, 20, 672, 168, 3, "hardswish")
case(64, 40, 40, 120, 32, 5, "relu")
