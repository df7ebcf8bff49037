"""Plain project GEMM vs the SE-folded one (A operand rescaled in shared memory) on the SE layer shapes of V3 at batch 256."""
import sys
import torch
sys.path.insert(0, ".")
from demonet_b200 import ops

def timeit(fn, n=20):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B = 256
for HW, K, N, res in [(1600, 120, 40, True), (1600, 72, 40, False), (400, 480, 112, False), (400, 672, 112, True), (100, 672, 80, False), (100, 480, 80, True)]:
    xs = [torch.randn(B, HW, K, device="cuda").half() for _ in range(3)]
    rs = [torch.randn(B * HW, N, device="cuda").half() for _ in range(3)] if res else [None] * 3
    w = (torch.randn(N, K, device="cuda") * 0.05).half()
    b = torch.randn(N, device="cuda")
    cs = max(8, K // 4)
    w1 = torch.randn(cs, K, device="cuda") * 0.05; b1 = torch.zeros(cs, device="cuda")
    w2t = torch.randn(cs, K, device="cuda") * 0.05; b2 = torch.zeros(K, device="cuda")
    t_plain = timeit(lambda i: ops.pwconv(xs[i % 3].view(B * HW, K), w, b, "none", rs[i % 3]))
    t_se = timeit(lambda i: ops.se_project(xs[i % 3], w1, b1, w2t, b2, w, b, rs[i % 3]))
    t_fc = timeit(lambda i: ops.se_inplace(xs[i % 3], w1, b1, w2t, b2))
    print("HW %4d K %3d N %3d res %d: plain GEMM %.4f ms, se_project (pool + fc + folded GEMM) %.4f ms, se_inplace %.4f ms" % (HW, K, N, res, t_plain, t_se, t_fc))
