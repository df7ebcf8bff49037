"""PAIR mode of the GEMM (clusters of two CTAs, weight k-blocks by TMA multicast): correctness vs torch and time, one
process per DN_PW_PAIR setting."""
import os, sys
sys.path.insert(0, ".")
import torch
from demonet_b200 import ops

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (M, K, N, out_fp32, res) in [(102400, 672, 546, True, False), (102400, 672, 112, False, False), (102400, 200, 80, False, True),
                                 (102400 - 77, 480, 112, False, False), (409600, 240, 80, False, False)]:
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).half().cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).half().cuda()
    b = torch.randn(N, generator=g).cuda()
    r = torch.randn(M, N, generator=g).half().cuda() if res else None
    y = ops.pwconv(x, w, b, "none", r, out_fp32)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + b
    if res:
        ref = ref + r.float()
    err = (y.float() - ref).abs().max().item()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.pwconv(x, w, b, "none", r, out_fp32); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("PAIR=%s M=%d K=%d N=%d fp32=%d res=%d: max err %.3g, %.4f ms" % (os.environ.get("DN_PW_PAIR", "1"), M, K, N, out_fp32, res, err, ts[5]), flush=True)
