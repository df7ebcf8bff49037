#!/usr/bin/env python
"""Benchmark of the SSDLite inference hot path (BASELINE.json metric: SSDLite320 images/sec).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One "step" = forward + post-processing of one batch of synthetic 320x320 images per GPU
(config 2 of BASELINE.json: ssdlite320_mobilenet_v3_large, 91 classes, batch 256, bf16).
For N > 1 the driver launches this file under torch.distributed.run; the image batch is sharded
(256 per GPU, weak scaling), there is no inter-GPU traffic on the hot path and one NCCL all-gather
of the fixed-shape detections closes each step.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ssdlite320_mobilenet_v3_large images/sec (forward + postprocess)"
UNIT = "img/s"
S = 320
NUM_CLASSES = 91


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=16, help="images per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", action="store_true", help="also print the per-layer time table to stderr")
    ap.add_argument("--pipeline", type=int, default=2, choices=[1, 2],
                    help="batches in flight per GPU (2 = the engine's pipeline mode, 1 = one forward at a time)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU path: the oracle's torch port of the reference (conv stack through the same ATen CPU kernels
# the reference uses + SSD.postprocess_detections restated with the same torch/torchvision calls)
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(n_images, steps, warmup):
    import torch
    from oracle import net_ref              # the CPU leg is the one place bench.py executes oracle/
    from demonet_b200 import plan as dplan, seeded as weights
    import demonet_b200
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(cores)
    model = demonet_b200.ssdlite320_mobilenet_v3_large()          # only for the state_dict template / plan
    sd = weights.seeded_state_dict(model.state_dict())
    anchors = torch.from_numpy(dplan.default_boxes(model.plan))
    x = weights.synthetic_images(n_images, S)

    def step():
        with torch.no_grad():
            cls, reg, _ = net_ref.v3_forward_raw(sd, x, "fp32", NUM_CLASSES)
            return net_ref.postprocess_detections_torch(cls, reg, anchors, (S, S))
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n_images * steps / dt, dt / steps * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, ms, cores = cpu_reference_rate(args.cpu_sample, args.steps, args.warmup)
    sample = "%d images/step, fp32, %d threads; oracle port of SSD.forward (reference is pure Python/PyTorch)" % (
        args.cpu_sample, cores)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ssdlite320_mobilenet_v3_large 91 classes 320x320 forward+postprocess on the host CPU",
                       "batch_per_step": args.cpu_sample, "weights": "seeded re-init 1234", "images": "torch.rand seed 1"},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_claim_stdout(), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region.  NVML is polled in-process every 2 ms (a timed
    region of a few tens of milliseconds still gets samples); `nvidia-smi -lms` is the fallback when the NVML
    binding is not importable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.sm, self.mx, self.reasons, self.power = [], None, set(), []
        self.nvml, self.handle, self.stop_flag, self.thread = None, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown if hasattr(n, "nvmlClocksEventReasonHwSlowdown")
                else n.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown",
                                               getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0)),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown",
                                               getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0)),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap",
                                        getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0))}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.sm.append(int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                r = int(get_reasons(self.handle))
                for nm, b in bits.items():
                    if b and (r & b):
                        self.reasons.add(nm)
                self.power.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml:
            self.stop_flag = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            if self.thread:
                self.thread.join(timeout=1.0)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(sm), "power_w_max": round(max(self.power), 1) if self.power else None,
                    "source": "nvml, 2 ms poll inside the timed region"}
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------------
# roofline bookkeeping: algorithmic bytes / flops of every launch of the plan
# ------------------------------------------------------------------------------------------------
def layer_costs(plan, B):
    """[(kernel_name, bytes, flops)] per launch, in plan order, then the 3 post-processing kernels.
    Algorithmic bytes = input + output activations at their stored width + weights."""
    out = []
    P, K = plan.num_priors, plan.num_classes
    from demonet_b200 import plan as dplan
    fused = set(dplan.fused_pairs(plan)) if int(os.environ.get("DN_FUSE", "1")) else set()
    fused_dp = set(dplan.fused_dwpw_pairs(plan)) if int(os.environ.get("DN_FUSE", "1")) and dplan.dwpw_fusion_enabled() else set()
    for i, L in enumerate(plan.layers):
        hi, wi, ho, wo = L.h_in, L.w_in, L.h_out, L.w_out
        if i in fused:            # expand + depthwise in one launch: the expanded tensor is not algorithmic traffic any more
            D = plan.layers[i + 1]
            out.append(("pwdw_fused_kernel", B * (hi * wi * L.cin + D.h_out * D.w_out * D.cout) * 2 + L.cin * L.cout * 2 + D.k * D.k * D.cout * 4,
                        2 * B * (hi * wi * L.cin * L.cout + D.h_out * D.w_out * D.cout * D.k * D.k)))
            continue
        if i - 1 in fused or i - 1 in fused_dp:
            out.append(("(fused into the previous launch)", 0, 0))
            continue
        if i in fused_dp:         # depthwise + project (+ residual): input read once, output written once
            Pj = plan.layers[i + 1]
            out.append(("dwpw_fused_kernel", B * hi * wi * (L.cin + Pj.cout) * 2 + L.k * L.k * L.cin * 4 + L.cin * Pj.cout * 2,
                        2 * B * hi * wi * (L.cin * L.k * L.k + L.cin * Pj.cout)))
            continue
        if L.kind == "stem":
            out.append(("stem_conv_kernel", B * (3 * hi * wi * 4 + ho * wo * L.cout * 2), 2 * B * ho * wo * L.cout * 27))
        elif L.kind == "dw":
            out.append(("dwconv_kernel", B * (hi * wi + ho * wo) * L.cin * 2 + L.k * L.k * L.cin * 4,
                        2 * B * ho * wo * L.cin * L.k * L.k))
        elif L.kind == "se":
            out.append(("se_pool+fc1+fc2+scale kernels", B * hi * wi * L.cin * 2 * 2 + 2 * L.cin * L.se_mid * 4,
                        2 * B * 2 * L.cin * L.se_mid))
        else:
            ob = 4 if L.head else 2
            res = B * hi * wi * L.cout * 2 if L.res else 0
            out.append(("pwconv_tc_kernel", B * hi * wi * (L.cin * 2 + L.cout * ob) + res + L.cin * L.cout * 2,
                        2 * B * hi * wi * L.cin * L.cout))
    out.append(("softmax_decode_kernel", B * (P * K * 4 + P * 16) + P * 16 + B * (P * (K - 1) * 4 + P * 16), 0))
    out.append(("class_sort+select+nms kernels", B * (K - 1) * P * 4, 0))
    out.append(("merge_topd_kernel", B * plan_D(plan) * 28, 0))
    return out


def plan_D(plan):
    return 300


_JSON_OUT = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line: keep a private handle on it and point fd 1 at stderr, so that whatever
    a library prints (NCCL's version banner under NCCL_DEBUG, warnings of C++ runtimes) cannot get in front of it."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _JSON_OUT


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import demonet_b200
    from demonet_b200 import _C
    from demonet_b200 import seeded as weights      # seeded re-init recipe + synthetic images
    import ctypes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL_DEBUG=VERSION/INFO prints to stdout by default; stdout carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    B, D = args.batch, 300

    # --pipeline 2 (default): two batches in flight on two engine instances (dn_model_desc.pipeline_slots): the
    # latency-bound tail of batch i (small layers, NMS rounds) overlaps the bandwidth-bound head of batch i+1
    model = demonet_b200.ssdlite320_mobilenet_v3_large(num_classes=NUM_CLASSES, pipeline_slots=args.pipeline)
    model.load_state_dict(weights.seeded_state_dict(model.state_dict()))
    model = model.to(dev)
    eng = model.reserve(B, dev)
    lib = _C.lib()
    stream = torch.cuda.current_stream(dev)
    nslot = 2 if args.pipeline == 2 else 1

    # inputs resident in HBM (value) and in pinned host memory (e2e); different images per rank
    imgs_host = weights.synthetic_images(B, S, seed=1 + rank).pin_memory()
    imgs = imgs_host.to(dev)
    # one packed output buffer per rank and slot so that the gather of a batch is ONE collective
    from demonet_b200 import dist as ddist
    packed_det = [ddist.PackedDetections(B, D, dev) for _ in range(nslot)]
    io = [p.as_io() for p in packed_det]
    gathered = torch.empty(world * packed_det[0].buffer.numel(), dtype=torch.uint8, device=dev) if world > 1 else None
    host_outs = [{"boxes": torch.empty(B, D, 4, dtype=torch.float32).pin_memory(),
                  "scores": torch.empty(B, D, dtype=torch.float32).pin_memory(),
                  "labels": torch.empty(B, D, dtype=torch.int64).pin_memory(),
                  "counts": torch.empty(B, dtype=torch.int32).pin_memory()} for _ in range(nslot)]
    host_out = host_outs[0]
    tick = {"dev": 0, "host": 0, "u8": 0, "gather_owed": False}

    def step_device():
        # forward of batch i on slot i % nslot; with N > 1 every step closes with ONE all-gather of detections: those
        # of this batch, or in pipeline mode those of the previous batch (its forward is joined, this one keeps running)
        k = tick["dev"] % nslot
        tick["dev"] += 1
        eng.forward(imgs, io[k])
        if world > 1:
            if nslot == 1:
                ddist.gather_detections(packed_det[0], gathered)
            else:
                if tick["gather_owed"]:
                    eng.join_previous()
                    ddist.gather_detections(packed_det[k ^ 1], gathered)
                tick["gather_owed"] = True

    def finish_device():
        # drain the pipeline inside the timed region: join the last forward(s) and gather the batch still owed
        if nslot == 2:
            eng.join()
            if world > 1 and tick["gather_owed"]:
                ddist.gather_detections(packed_det[(tick["dev"] - 1) % nslot], gathered)
                tick["gather_owed"] = False

    def step_host():
        k = tick["host"] % nslot
        tick["host"] += 1
        eng.forward_host(imgs_host, host_outs[k])   # results land in each rank's own host memory: no gather

    imgs_u8_host = (imgs_host * 255).round().to(torch.uint8).pin_memory()      # what an image decoder hands over

    def step_host_u8():
        k = tick["u8"] % nslot
        tick["u8"] += 1
        eng.forward_host_u8(imgs_u8_host, host_outs[k])

    def timed(fn, steps, warmup, sampler=None, finish=None):
        for _ in range(warmup):
            fn()
        if finish:
            finish()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        if finish:
            finish()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks

    warm = max(args.warmup, 3)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    total_ms, clocks = timed(step_device, args.steps, warm, sampler, finish_device)
    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms * 1e-3)

    e2e_ms, _ = timed(step_host, args.steps, warm)
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    counts_ok = all(int(h["counts"].min()) >= 0 for h in host_outs)
    u8_ms, _ = timed(step_host_u8, args.steps, warm)
    u8_value = world * B * args.steps / (u8_ms * 1e-3)

    # per-launch device times, live, with CUDA events on the launching stream
    n_launch = eng.launches_per_forward
    ms = (ctypes.c_float * (len(model.plan.layers) + 3))()
    with torch.cuda.device(dev):
        _C.check(lib.dn_engine_profile(eng._handle, imgs.data_ptr(), B, 10, ms, stream.cuda_stream))
    costs = layer_costs(model.plan, B)
    per_kernel = {}
    for (name, nbytes, flops), t in zip(costs, list(ms)):
        if nbytes == 0 and flops == 0 and name.startswith("(fused"):
            continue
        k = per_kernel.setdefault(name, {"ms": 0.0, "bytes": 0, "flops": 0, "launches": 0})
        k["ms"] += t; k["bytes"] += nbytes; k["flops"] += flops; k["launches"] += 1
    sum_ms = sum(k["ms"] for k in per_kernel.values())
    dominant = max(per_kernel, key=lambda n: per_kernel[n]["ms"])
    peaks = {"hbm_gbs": 6650.0, "src": "fallback"}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks = {"hbm_gbs": float(mp["hbm_gbs"]), "bf16_tflops": float(mp.get("bf16_tflops_sustained", mp["bf16_tflops"])),
                 "src": "measured"}
    except Exception:
        pass
    dk = per_kernel[dominant]
    achieved = dk["bytes"] / (dk["ms"] * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel from the committed ncu launch list of this same command
    # (profiles/r01_ncu_launches_*.json, dram__bytes_read.sum + dram__bytes_write.sum), per launch like `achieved`
    traffic, traffic_src = None, None
    try:
        import glob
        import re

        def _ver(path):                         # r01_ncu_launches_v14.json -> (1, 14): newest round / version last
            m = re.search(r"r(\d+)_ncu_launches_v(\d+)", os.path.basename(path))
            return (int(m.group(1)), int(m.group(2))) if m else (0, 0)
        cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_launches_*.json")), key=_ver)
        if cands and B == 256:
            nl = json.load(open(cands[-1]))["kernels"]
            fam = {"dwconv_kernel": ["dwconv_kernel", "dwconv_tma_kernel", "dwconv_stream_kernel", "dwconv_stream2_kernel"],
                   "pwconv_tc_kernel": ["pwconv_tc_kernel"], "pwdw_fused_kernel": ["pwdw_fused_kernel"],
                   "dwpw_fused_kernel": ["dwpw_fused_kernel"], "stem_conv_kernel": ["stem_tma_kernel", "stem_conv_kernel"],
                   "se_pool+fc1+fc2+scale kernels": ["se_pool_kernel", "se_fc1_kernel", "se_fc2_kernel", "se_scale_kernel"]}
            names = fam.get(dominant, [dominant])
            names = names + ["dn::" + n for n in names]              # ncu reports the name with or without the namespace
            tb = sum(nl[n]["dram_bytes"] for n in names if n in nl)
            tl = sum(nl[n]["launches"] for n in names if n in nl)
            if dominant.startswith("se_pool"):
                tl //= 4                                             # four launches per squeeze-excitation layer
            if tl:
                traffic, traffic_src = tb / tl, os.path.basename(cands[-1])
    except Exception:
        pass
    roofline = {"kernel": dominant, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write)",
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": dk["bytes"] / dk["launches"], "peak_source": peaks["src"] + " (MEASURED_PEAKS.json hbm_gbs)",
                "launches_per_step": dk["launches"], "avg_launch_ms": dk["ms"] / dk["launches"],
                "share_of_step": dk["ms"] / sum_ms,
                "tensor_tflops": dk["flops"] / (dk["ms"] * 1e-3) / 1e12 if dk["flops"] else None,
                "per_kernel": {n: {"ms": round(k["ms"], 4), "share": round(k["ms"] / sum_ms, 4),
                                   "GBps": round(k["bytes"] / (k["ms"] * 1e-3) / 1e9, 1),
                                   "TFLOPs": round(k["flops"] / (k["ms"] * 1e-3) / 1e12, 2), "launches": k["launches"]}
                               for n, k in per_kernel.items()}}
    if args.layers and rank == 0:
        for i, ((name, nbytes, flops), t) in enumerate(zip(costs, list(ms))):
            L = model.plan.layers[i] if i < len(model.plan.layers) else None
            desc = "%s %dx%d c%d->%d k%d s%d" % (L.kind, L.h_in, L.w_in, L.cin, L.cout, L.k, L.stride) if L else name
            tt = max(t, 1e-9)
            if name == "pwdw_fused_kernel":
                desc = "pw+dw fused " + desc[3:]
            if name == "dwpw_fused_kernel":
                desc = "dw+pw fused " + desc[3:]
            print("%3d %-34s %8.4f ms %8.1f GB/s %7.2f TF/s" % (i, desc, t, nbytes / (tt * 1e-3) / 1e9,
                                                             flops / (tt * 1e-3) / 1e12), file=sys.stderr)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "ssdlite320_mobilenet_v3_large, 91 classes, 320x320, batch %d per GPU, bf16 activations, "
                                   "forward + softmax/decode/top-k/NMS/top-300 (BASELINE.json configs[1])" % B,
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d (image shards, final NCCL "
                       "all-gather of detections)" % world, "weights": "seeded re-init 1234 (demonet_b200/seeded.py)",
                       "images": "torch.rand seed 1+rank", "l2": "inputs larger than L2 (%.0f MB fp32 images per step; "
                       "activation arena %.1f GB)" % (B * 3 * S * S * 4 / 1e6, eng.device_bytes / 1e9),
                       "cuda_graph": True,
                       "batches_in_flight": nslot},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": B * 3 * S * S * 4, "d2h_bytes_per_step": B * D * 28 + B * 4,
                    "api": "dn_engine_forward_host (pinned fp32 images in, detections out)", "ok": counts_ok},
            "e2e_u8": {"value": u8_value, "unit": UNIT, "ms_per_step": u8_ms / args.steps,
                       "h2d_bytes_per_step": B * 3 * S * S, "d2h_bytes_per_step": B * D * 28 + B * 4,
                       "api": "dn_engine_forward_host_u8 (pinned uint8 pixels in, x/255 on the device, detections out); "
                              "the e2e entry above keeps the reference's fp32 input contract and is PCIe-bound"},
            "gpu_launches": n_launch * args.steps,
            "gpu_launches_per_step": n_launch,
            "roofline": roofline}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cpu_ms, cores = cpu_reference_rate(args.cpu_sample, 3, 1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d images x 3 steps, fp32, oracle torch port of SSD.forward (%.0f ms/step)"
                                          % (args.cpu_sample, cpu_ms)}
    if rank == 0:
        print(json.dumps(line), file=_claim_stdout(), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
