"""CPU interpreter of the engine's layer plan (test helper).

Executes exactly what the C engine is given -- the dn_op array, the packed weight blob and the arena
buffer assignment -- with plain torch ops on the CPU, including the buffer reuse, so that plan wiring
bugs (wrong key, wrong residual, wrong head offset, a buffer freed too early) are caught without a GPU.
Numerics follow the engine contract (fp16 or bf16 storage, fp32 accumulate)."""
import numpy as np
import torch
import torch.nn.functional as F

from demonet_b200 import _C


def _act(x, a):
    return {0: lambda v: v, 1: F.relu, 2: F.relu6, 3: F.hardswish}[a](x)


def run_plan(plan, blob, ops, bufs, logits_buf, bbox_buf, images, mean, std, round_activations=True, act_dtype=None):
    B = images.shape[0]
    h16 = _C.torch_dtype(act_dtype)
    _bf16 = (lambda v: v.to(h16).to(torch.float32)) if round_activations else (lambda v: v)
    raw = np.frombuffer(blob, dtype=np.uint8)
    arena = [torch.full((B * e,), float("nan")) for e, _ in bufs]

    def f32(off, n):
        return torch.from_numpy(raw[off:off + 4 * n].view(np.float32).copy())

    def bf16w(off, n):
        return torch.from_numpy(raw[off:off + 2 * n].view(np.int16).copy()).view(h16).float()

    for op in ops:
        hi, wi, ci, ho, wo, co = op.h_in, op.w_in, op.c_in, op.h_out, op.w_out, op.c_out
        if op.kind == _C.OP_STEM:
            m = torch.tensor(mean)[None, :, None, None]
            s = torch.tensor(std)[None, :, None, None]
            w = f32(op.w_off, 27 * co).view(3, 3, 3, co).permute(3, 0, 1, 2)
            y = _act(F.conv2d((images - m) / s, w, f32(op.b_off, co), 2, 1), op.act)
            arena[op.out_buf][:B * ho * wo * co] = _bf16(y).permute(0, 2, 3, 1).reshape(-1)
        elif op.kind == _C.OP_DW:
            x = arena[op.in_buf][:B * hi * wi * ci].view(B, hi, wi, ci).permute(0, 3, 1, 2)
            k = op.ksize
            w = f32(op.w_off, k * k * ci).view(k, k, ci).permute(2, 0, 1)[:, None]
            y = _act(F.conv2d(x, w, f32(op.b_off, ci), op.stride, (k - 1) // 2, 1, ci), op.act)
            assert y.shape[-2:] == (ho, wo)
            arena[op.out_buf][:B * ho * wo * co] = _bf16(y).permute(0, 2, 3, 1).reshape(-1)
        elif op.kind == _C.OP_NOP:
            continue
        elif op.kind == _C.OP_DWPW:
            x = arena[op.in_buf][:B * hi * wi * ci].view(B, hi, wi, ci)
            k = op.ksize
            w = f32(op.w_off, k * k * ci).view(k, k, ci).permute(2, 0, 1)[:, None]
            mid = _bf16(_act(F.conv2d(x.permute(0, 3, 1, 2), w, f32(op.b_off, ci), op.stride, (k - 1) // 2, 1, ci), op.act))
            y = mid.permute(0, 2, 3, 1).reshape(-1, ci) @ bf16w(op.w2_off, co * ci).view(co, ci).t() + f32(op.b2_off, co)
            if op.res_buf >= 0:
                y = y + arena[op.res_buf][:B * hi * wi * co].view(-1, co)
            arena[op.out_buf][:B * ho * wo * co] = _bf16(y).reshape(-1)
        elif op.kind == _C.OP_PWDW:
            hw = hi * wi
            x = arena[op.in_buf][:B * hw * ci].view(B * hw, ci)
            mid = _bf16(_act(x @ bf16w(op.w_off, co * ci).view(co, ci).t() + f32(op.b_off, co), op.act))
            k = op.ksize
            w = f32(op.w2_off, k * k * co).view(k, k, co).permute(2, 0, 1)[:, None]
            y = _act(F.conv2d(mid.view(B, hi, wi, co).permute(0, 3, 1, 2), w, f32(op.b2_off, co), op.stride, (k - 1) // 2, 1, co), op.act2)
            assert y.shape[-2:] == (ho, wo)
            arena[op.out_buf][:B * ho * wo * co] = _bf16(y).permute(0, 2, 3, 1).reshape(-1)
        elif op.kind == _C.OP_SE:
            x = arena[op.in_buf][:B * hi * wi * ci].view(B, hi * wi, ci)
            cs = op.c_mid
            pooled = x.mean(1)
            hid = F.relu(pooled @ f32(op.w_off, cs * ci).view(cs, ci).t() + f32(op.b_off, cs))
            sc = F.hardsigmoid(hid @ f32(op.w2_off, cs * ci).view(cs, ci) + f32(op.b2_off, ci))
            arena[op.in_buf][:B * hi * wi * ci] = _bf16(x * sc[:, None, :]).reshape(-1)
        else:
            hw = hi * wi
            x = arena[op.in_buf][:B * hw * ci].view(B * hw, ci)
            assert not torch.isnan(x).any(), "PW reads an unwritten buffer"
            y = _act(x @ bf16w(op.w_off, co * ci).view(co, ci).t() + f32(op.b_off, co), op.act)
            if op.res_buf >= 0:
                y = y + arena[op.res_buf][:B * hw * co].view(B * hw, co)
            if op.out_fp32:
                per_img = bufs[op.out_buf][0]
                dst = arena[op.out_buf].view(B, per_img)
                rows = torch.arange(hw)[:, None] * op.out_row_stride + op.out_offset + torch.arange(co)[None, :]
                dst[:, rows.reshape(-1)] = y.view(B, hw * co)
            else:
                arena[op.out_buf][:B * hw * co] = _bf16(y).reshape(-1)
    P, K = plan.num_priors, plan.num_classes
    return arena[logits_buf].view(B, P, K), arena[bbox_buf].view(B, P, 4)
