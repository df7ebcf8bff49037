"""Experiment: is the cls-head level-0 GEMM bound by the L2 -> shared-memory operand stream?  Same GEMM, tile width capped
by DN_PW_BN_MAX (one process per setting): narrower N tiles re-fetch the A tile more often."""
import os, sys
sys.path.insert(0, ".")
import torch
from demonet_b200 import ops

M, K, N = 102400, 672, 546
g = torch.Generator().manual_seed(0)
x = torch.randn(M, K, generator=g).half().cuda()
w = (torch.randn(N, K, generator=g) / K ** 0.5).half().cuda()
b = torch.randn(N, generator=g).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for out_fp32 in (True, False):
    for _ in range(3):
        ops.pwconv(x, w, b, "none", None, out_fp32)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.pwconv(x, w, b, "none", None, out_fp32); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("BN_MAX=%s out_fp32=%d: %.4f ms (median of 10, L2 flushed)" % (os.environ.get("DN_PW_BN_MAX", "256"), out_fp32, ts[5]), flush=True)
