#!/usr/bin/env python
"""Benchmark of the SSDLite inference hot path (BASELINE.json metric: SSDLite320 images/sec at 1/2/4/8 B200 vs the
host-CPU reference; decode+NMS us/img).

  python bench.py --gpus N --steps K --warmup W              # this repo's CUDA engine, BASELINE.json configs[1]
  python bench.py --config {2,3,4,5} ...                     # the other BASELINE.json configurations, same JSON schema
  python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path (oracle port) on the host cores

One "step" = forward + post-processing of one batch of synthetic images per GPU.
  config 2 (default)  ssdlite320_mobilenet_v3_large, 91 classes, batch 256 per GPU (weak scaling over N GPUs)
  config 3            the same network, GLOBAL batch 2048 split over the N GPUs (strong scaling) + detection gather
  config 4            post-processing stress: 3234 anchors x 91 classes, thr .001, top-k 400, NMS .55, batch 1024 (us/img)
  config 5            ssd_lite_mobilenet_v2, 21 classes, 512x512, batch 512 per GPU
For N > 1 the driver launches this file under torch.distributed.run: images are sharded, there is no inter-GPU traffic on
the hot path, and ONE NCCL all-gather of the fixed-shape detections closes each step.  Prints ONE JSON line on rank 0.
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

CONFIGS = {
    2: dict(model="v3", S=320, K=91, batch=256, scaling="weak", D=300, unit="img/s", hib=True,
            metric="ssdlite320_mobilenet_v3_large images/sec (forward + postprocess)",
            name="ssdlite320_mobilenet_v3_large, 91 classes, 320x320, batch %d per GPU, %s activations, forward + "
                 "softmax/decode/top-k/NMS/top-300 (BASELINE.json configs[1])"),
    3: dict(model="v3", S=320, K=91, global_batch=2048, scaling="strong", D=300, unit="img/s", hib=True,
            metric="ssdlite320_mobilenet_v3_large images/sec (forward + postprocess)",
            name="ssdlite320_mobilenet_v3_large, 91 classes, 320x320, GLOBAL batch 2048 sharded (%d per GPU), %s activations, "
                 "forward + postprocess + NCCL detection gather (BASELINE.json configs[2])"),
    4: dict(model="post", S=320, K=91, batch=1024, scaling="weak", D=300, unit="us/img", hib=False,
            metric="decode+NMS microseconds per image (softmax + decode + threshold + top-k + batched NMS + top-300)",
            name="postprocess stress: 3234 anchors x 91 classes, logits N(0,4^2) seed 7, score_thresh 0.001, top-k 400, "
                 "NMS IoU 0.55, D 300, batch %d per GPU (BASELINE.json configs[3])"),
    5: dict(model="v2", S=512, K=21, batch=512, scaling="weak", D=100, unit="img/s", hib=True,
            metric="ssd_lite_mobilenet_v2 512x512 images/sec (forward + postprocess)",
            name="ssd_lite_mobilenet_v2, VOC 21 classes, 512x512, batch %d per GPU, %s activations, forward + legacy PostProcess "
                 "(thr 0.5, NMS 0.45, top-100) (BASELINE.json configs[4])"),
    6: dict(model="vgg", S=300, K=91, batch=64, scaling="weak", D=200, unit="img/s", hib=True,
            metric="ssd300_vgg16 images/sec (forward + postprocess)",
            name="ssd300_vgg16, 91 classes, 300x300, batch %d per GPU, %s activations, forward + postprocess (SURVEY.md 8(f4); not a "
                 "BASELINE.json configuration)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configuration (1-based)")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: the configuration's)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images per CPU-baseline step (default 32, BASELINE.md section 3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the e2e / API / per-kernel legs (timing runs only)")
    ap.add_argument("--layers", action="store_true", help="also print the per-layer time table to stderr")
    ap.add_argument("--act-dtype", default=None, choices=["fp16", "bf16"],
                    help="activation storage type (default: the library default, fp16)")
    ap.add_argument("--gather-every", type=int, default=1,
                    help="N > 1: all-gather the detections of this many consecutive batches in ONE collective (1 = a collective "
                         "per step; every detection is gathered inside the timed region either way)")
    ap.add_argument("--pipeline", type=int, default=4, choices=[1, 2, 3, 4],
                    help="batches in flight per GPU (2 = the engine's pipeline mode, 1 = one forward at a time)")
    return ap.parse_args()


def bind_to_gpu_numa_node(index):
    """Run this rank on the CPUs NVML reports as local to its GPU (nvmlDeviceGetCpuAffinity), so that the pinned staging
    buffers it allocates afterwards are first-touched on the NUMA node the GPU's PCIe root hangs off: with 8 ranks the
    host-buffer path otherwise funnels every H2D copy through whichever node the allocations happened to land on.
    Returns (original affinity, description); the original set is restored for the CPU baseline."""
    try:
        import pynvml
        orig = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1} & orig
        if not cpus or cpus == orig:
            return orig, "all %d cpus (NVML reports no narrower GPU-local set)" % len(orig)
        os.sched_setaffinity(0, cpus)
        return orig, "%d of %d cpus local to GPU %d (nvmlDeviceGetCpuAffinity)" % (len(cpus), len(orig), phys)
    except Exception as e:          # no NVML / not permitted: run unbound
        try:
            return os.sched_getaffinity(0), "unbound (%s)" % type(e).__name__
        except Exception:
            return None, "unbound"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
# CPU path: the oracle's torch port of the reference (conv stack through the same ATen CPU kernels the reference
# uses + SSD.postprocess_detections restated with the same torch / torchvision calls and Python loops).
# Protocol of BASELINE.md section 3: one process, all host cores, batch 32, warm-up + timed iterations, per-stage
# milliseconds; plus the reference-as-shipped evaluation setting (1 thread, batch 1: engine.py:73-75, train.py:141-144).
# ------------------------------------------------------------------------------------------------
def cpu_reference(cfg_id, n_images, steps, warmup, threads=None):
    import numpy as np
    import torch
    from oracle import boxes_np, net_ref     # the CPU leg is the one place bench.py executes oracle/
    from demonet_b200 import plan as dplan, seeded as weights
    import demonet_b200
    cfg = CONFIGS[cfg_id]
    cores = threads or host_cores()
    torch.set_num_threads(cores)
    S, K = cfg["S"], cfg["K"]
    stages = {}
    if cfg["model"] == "vgg":
        model = demonet_b200.ssd300_vgg16(num_classes=K)              # state_dict template only
        sd = weights.seeded_vgg_state_dict(model.state_dict())
        x = weights.synthetic_images(n_images, S)
        anchors = torch.from_numpy(dplan.default_boxes_for([(38, 38), (19, 19), (10, 10), (5, 5), (3, 3), (1, 1)], 300,
                                                           [[2], [2, 3], [2, 3], [2, 3], [2], [2]],
                                                           scales=[0.07, 0.15, 0.33, 0.51, 0.69, 0.87, 1.05], steps=[8, 16, 32, 64, 100, 300]))

        def step():
            with torch.no_grad():
                cls, reg, _ = net_ref.vgg_forward_raw(sd, x, "fp32", K, times=stages)
                t0 = time.perf_counter()
                net_ref.postprocess_detections_torch(cls, reg, anchors, (S, S), 0.01, 0.45, 200, 400)
                stages["postprocess"] = stages.get("postprocess", 0.0) + time.perf_counter() - t0
    elif cfg["model"] == "post":
        g = torch.Generator().manual_seed(7)
        logits = torch.randn(n_images, 3234, K, generator=g) * 4.0
        bbox = torch.randn(n_images, 3234, 4, generator=g) * 1.5
        anchors = torch.from_numpy(dplan.default_boxes(dplan.plan_ssdlite320_mobilenet_v3_large()))

        def step():
            t0 = time.perf_counter()
            net_ref.postprocess_detections_torch(logits, bbox, anchors, (S, S), 0.001, 0.55, 300, 400)
            stages["postprocess"] = stages.get("postprocess", 0.0) + time.perf_counter() - t0
    else:
        if cfg["model"] == "v3":
            model = demonet_b200.ssdlite320_mobilenet_v3_large(num_classes=K)      # state_dict template / plan only
            fwd = net_ref.v3_forward_raw
        else:
            model = demonet_b200.ssd_lite_mobilenet_v2(image_size=S, num_classes=K)
            fwd = net_ref.v2_forward_raw
        sd = weights.seeded_state_dict(model.state_dict())
        x = weights.synthetic_images(n_images, S)

        def step():
            with torch.no_grad():
                cls, reg, _ = fwd(sd, x, "fp32", K, times=stages)
                t0 = time.perf_counter()
                anchors = dplan.default_boxes(model.plan)          # DefaultBoxGenerator runs every forward in the reference
                stages["anchors"] = stages.get("anchors", 0.0) + time.perf_counter() - t0
                t0 = time.perf_counter()
                if cfg["model"] == "v3":
                    net_ref.postprocess_detections_torch(cls, reg, torch.from_numpy(anchors), (S, S))
                else:
                    sc = torch.softmax(cls, -1).numpy()
                    for i in range(cls.shape[0]):
                        boxes_np.legacy_postprocess(None, reg[i].numpy(), anchors, (S, S), score_thresh=0.5, scores=sc[i])
                stages["postprocess"] = stages.get("postprocess", 0.0) + time.perf_counter() - t0
    for _ in range(warmup):
        step()
    stages.clear()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    per_img = {k: round(v / (steps * n_images) * 1e3, 3) for k, v in stages.items()}
    return {"img_per_s": n_images * steps / dt, "ms_per_step": dt / steps * 1e3, "cores": cores,
            "stage_ms_per_img": per_img, "batch": n_images, "steps": steps}


def cpu_value(cfg_id, r):
    """the configuration's metric from a cpu_reference() result"""
    return 1e6 / r["img_per_s"] if CONFIGS[cfg_id]["unit"] == "us/img" else r["img_per_s"]


def cpu_baseline_block(cfg_id, n_images, steps, warmup, as_shipped=True):
    r = cpu_reference(cfg_id, n_images, steps, warmup)
    blk = {"value": cpu_value(cfg_id, r), "unit": CONFIGS[cfg_id]["unit"], "cores": r["cores"], "kind": "port",
           "sample": "%d images/step x %d timed steps (+%d warm-up), fp32, %d threads, one process; oracle torch port of "
                     "SSD.forward (the reference is pure Python/PyTorch: same ATen CPU kernels, same image x class loops); "
                     "%.0f ms/step" % (n_images, steps, warmup, r["cores"], r["ms_per_step"]),
           "stage_ms_per_img": r["stage_ms_per_img"]}
    if as_shipped:
        # the reference's own evaluation loop pins ONE thread and feeds batch 1 (engine.py:73-75, train.py:141-144)
        n1 = 3 if CONFIGS[cfg_id]["model"] != "post" else 2
        r1 = cpu_reference(cfg_id, 1, n1, 1, threads=1)
        blk["as_shipped"] = {"value": cpu_value(cfg_id, r1), "unit": CONFIGS[cfg_id]["unit"], "cores": 1,
                             "sample": "batch 1, 1 thread, %d timed images (reference-as-shipped: engine.py:73-75, "
                                       "train.py:141-144)" % n1, "stage_ms_per_img": r1["stage_ms_per_img"]}
    return blk


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    n = args.cpu_sample or (8 if cfg["model"] in ("post", "v2", "vgg") else 32)
    steps = max(1, min(args.steps, 10 if cfg["model"] == "v3" else 4))     # bounded: the whole run ends within minutes
    warm = max(1, min(args.warmup, 3))
    blk = cpu_baseline_block(args.config, n, steps, warm)
    workload = {"v3": "ssdlite320_mobilenet_v3_large 91 classes 320x320 forward+postprocess on the host CPU",
                "v2": "ssd_lite_mobilenet_v2 21 classes 512x512 forward+legacy PostProcess on the host CPU",
                "post": "postprocess stress (3234 anchors x 91 classes, thr .001, top-k 400, NMS .55) on the host CPU",
                "vgg": "ssd300_vgg16 91 classes 300x300 forward+postprocess on the host CPU"}[cfg["model"]]
    line = {"impl": "reference", "metric": cfg["metric"], "value": blk["value"], "unit": cfg["unit"], "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": None, "higher_is_better": cfg["hib"], "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "baseline_config": args.config, "batch_per_step": n,
                       "weights": "seeded re-init 1234", "images": "torch.rand seed 1",
                       "protocol": "BASELINE.md section 3 (bounded: %d timed steps)" % steps},
            "cpu_baseline": blk,
            "e2e": {"value": blk["value"], "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["ms_per_step"] = 1e3 * n / blk["value"] if cfg["unit"] == "img/s" else blk["value"] * n / 1e3
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
def layer_costs(plan, B, D):
    """[(kernel_name, bytes, flops)] per launch, in plan order, then the 3 post-processing entries of dn_engine_profile.
    Algorithmic bytes = input + output activations at their stored width + weights."""
    out = []
    P, K = plan.num_priors, plan.num_classes
    from demonet_b200 import plan as dplan
    fuse = int(os.environ.get("DN_FUSE", "1"))
    fused = set(dplan.fused_pairs(plan)) if fuse else set()
    fused_dp = set(dplan.fused_dwpw_pairs(plan)) if fuse and dplan.dwpw_fusion_enabled() else set()
    for i, L in enumerate(plan.layers):
        hi, wi, ho, wo = L.h_in, L.w_in, L.h_out, L.w_out
        if i in fused:            # expand + depthwise in one launch: the expanded tensor is not algorithmic traffic any more
            Dn = plan.layers[i + 1]
            out.append(("pwdw_fused_kernel", B * (hi * wi * L.cin + Dn.h_out * Dn.w_out * Dn.cout) * 2 + L.cin * L.cout * 2 + Dn.k * Dn.k * Dn.cout * 4,
                        2 * B * (hi * wi * L.cin * L.cout + Dn.h_out * Dn.w_out * Dn.cout * Dn.k * Dn.k)))
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
            out.append(("stem_tc_kernel", B * (3 * hi * wi * 4 + ho * wo * L.cout * 2), 2 * B * ho * wo * L.cout * 27))
        elif L.kind == "dw":
            out.append(("dwconv kernels (stream / stream2 / tma / direct)", B * (hi * wi + ho * wo) * L.cin * 2 + L.k * L.k * L.cin * 4,
                        2 * B * ho * wo * L.cin * L.k * L.k))
        elif L.kind == "se":
            out.append(("se kernels (fc1 + fc2 + scale)", B * hi * wi * L.cin * 2 * 2 + 2 * L.cin * L.se_mid * 4,
                        2 * B * 2 * L.cin * L.se_mid))
        else:
            ob = 4 if L.head else 2
            res = B * hi * wi * L.cout * 2 if L.res else 0
            out.append(("pwconv_tc_kernel", B * hi * wi * (L.cin * 2 + L.cout * ob) + res + L.cin * L.cout * 2,
                        2 * B * hi * wi * L.cin * L.cout))
    out.append(("softmax_decode_kernel (+ histogram, thresholds)", B * (P * K * 4 + P * 16) + P * 16 + B * (P * (K - 1) * 4 + P * 16), 0))
    out.append(("class_sort + nms + merge kernels (lazy rounds)", B * (K - 1) * P * 4 + B * D * 28, 0))
    out.append(("(empty event bracket)", 0, 0))
    return out


NCU_FAMILIES = {"dwconv kernels (stream / stream2 / tma / direct)": ["dwconv_kernel", "dwconv_tma_kernel", "dwconv_stream_kernel", "dwconv_stream2_kernel"],
                "pwconv_tc_kernel": ["pwconv_tc_kernel"], "pwdw_fused_kernel": ["pwdw_fused_kernel"],
                "dwpw_fused_kernel": ["dwpw_fused_kernel"], "stem_tc_kernel": ["stem_tc_kernel", "stem_tma_kernel", "stem_conv_kernel"],
                "se kernels (fc1 + fc2 + scale)": ["se_pool_kernel", "se_fc1_kernel", "se_fc2_kernel", "se_scale_kernel", "se_fc_kernel"],
                "softmax_decode_kernel (+ histogram, thresholds)": ["softmax_decode_kernel", "pick_thresholds_kernel"],
                "class_sort + nms + merge kernels (lazy rounds)": ["class_sort_kernel", "class_nms_warp_kernel", "class_nms_cta_kernel", "merge_topd_kernel"]}


def ncu_traffic(family, B):
    """DRAM bytes per launch of a kernel family from the newest committed ncu launch list of this command
    (profiles/rNN_ncu_launches_vM.json: dram__bytes_read.sum + dram__bytes_write.sum), or (None, None)."""
    try:
        import glob
        import re

        def _ver(path):
            m = re.search(r"r(\d+)_ncu_launches_v(\d+)", os.path.basename(path))
            return (int(m.group(1)), int(m.group(2))) if m else (0, 0)
        cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_launches_*.json")), key=_ver)
        if not cands or B != 256:
            return None, None
        doc = json.load(open(cands[-1]))
        nl = doc["kernels"]
        names = NCU_FAMILIES.get(family, [family])
        names = names + ["dn::" + n for n in names]
        tb = sum(nl[n]["dram_bytes"] for n in names if n in nl)
        tl = sum(nl[n]["launches"] for n in names if n in nl)
        return (tb / tl, os.path.basename(cands[-1])) if tl else (None, None)
    except Exception:
        return None, None


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


def load_peaks():
    peaks = {"hbm_gbs": 6650.0, "src": "fallback (B200_PROFILING.md)"}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks = {"hbm_gbs": float(mp["hbm_gbs"]), "bf16_tflops": float(mp.get("bf16_tflops_sustained", mp["bf16_tflops"])),
                 "src": "measured (MEASURED_PEAKS.json hbm_gbs)"}
    except Exception:
        pass
    return peaks


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import ctypes
    import torch
    import torch.distributed as dist
    import demonet_b200
    from demonet_b200 import _C, dist as ddist, plan as dplan
    from demonet_b200 import seeded as weights      # seeded re-init recipe + synthetic images
    from demonet_b200.module import make_post_params

    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU path)")
    orig_affinity, affinity_note = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    S, K, D = cfg["S"], cfg["K"], cfg["D"]
    if "global_batch" in cfg:
        lo, hi = ddist.shard_range(cfg["global_batch"], rank, world)
        B = args.batch or (hi - lo)
    else:
        B = args.batch or cfg["batch"]
    stream = torch.cuda.current_stream(dev)
    nslot = args.pipeline if args.pipeline >= 2 else 1
    peaks = load_peaks()
    warm = max(args.warmup, 3)
    sampler = ClockSampler(local_rank) if rank == 0 else None

    def timed(fn, steps, warmup, sampler=None, finish=None, prime=0):
        # `prime` untimed calls in front of the W warm-up steps: engine set-up, not warm-up -- every engine instance runs a
        # new (batch, input buffer, output buffers) combination eagerly once, captures its CUDA graph on the second call and
        # replays it from the third, so 3 calls per slot (6 on the host legs, which alternate two staging buffers per slot)
        # put the graph instantiation outside the timed region whatever W is
        for _ in range(prime + warmup):
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

    def to_metric(total_ms, steps):
        """the configuration's metric over all ranks from a timed region"""
        return (total_ms * 1e3 / (steps * B)) if cfg["unit"] == "us/img" else (world * B * steps / (total_ms * 1e-3))

    line = {"metric": cfg["metric"], "unit": cfg["unit"], "n_gpus": world, "steps": args.steps, "warmup": warm,
            "higher_is_better": cfg["hib"], "scaling": cfg["scaling"], "vs_baseline": None, "data": "synthetic"}

    if cfg["model"] == "vgg":
        # ---------------- ssd300_vgg16, layer by layer from Python (tensor-bound) -----------------------
        from demonet_b200 import ops, vgg as dvgg
        model = demonet_b200.ssd300_vgg16(num_classes=K, act_dtype=args.act_dtype)
        model.load_state_dict(weights.seeded_vgg_state_dict(model.state_dict()))
        model = model.to(dev)
        imgs_host = weights.synthetic_images(B, S, seed=1 + rank).pin_memory()
        imgs = imgs_host.to(dev)
        anchors = model.anchors(dev)
        keep = {}

        def step_device():
            cls, reg = model.head_outputs(imgs)
            keep["out"] = ops.postprocess_padded(cls, reg, anchors, (S, S), model.score_thresh, model.nms_thresh,
                                                 model.detections_per_img, model.topk_candidates)

        stage = torch.empty_like(imgs)

        def step_host():
            stage.copy_(imgs_host, non_blocking=True)
            cls, reg = model.head_outputs(stage)
            out = ops.postprocess_padded(cls, reg, anchors, (S, S), model.score_thresh, model.nms_thresh,
                                         model.detections_per_img, model.topk_candidates)
            keep["host"] = [t.to("cpu", non_blocking=True) for t in out]
        total_ms, clocks = timed(step_device, args.steps, warm, sampler)
        # algorithmic flops of the convolutions (2 * MAC, unpadded channels)
        flops, hw, launches = 0, S, 0
        for op in dvgg._VGG:
            if op[0] == "first":
                flops += 2 * hw * hw * 27 * 64
                launches += 1
            elif op[0] == "conv":
                ho = (hw + 2 * op[5] - 2 * op[6] - 1) // op[4] + 1
                flops += 2 * ho * ho * op[2] * op[3] * 9
                hw = ho
                launches += 1
            elif op[0] == "pw":
                flops += 2 * hw * hw * op[2] * op[3]
                launches += 1
            elif op[0] == "pool":
                k, st_, p, ceil = op[1:]
                hw = -(-(hw + 2 * p - k) // st_) + 1 if ceil else (hw + 2 * p - k) // st_ + 1
                launches += 1
            elif op[0] == "tap_l2norm":
                launches += 1
        for (gh, gw), c, a in zip(dvgg._GRIDS, dvgg._FEATURE_CHANNELS, model.num_anchors):
            flops += 2 * gh * gw * c * a * (K + 4) * 9
            launches += 2
        launches += 14
        step_ms = total_ms / args.steps
        tf = flops * B / (step_ms * 1e-3) / 1e12
        peak_tf = None
        try:
            peak_tf = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
        except Exception:
            peak_tf = 1392.3
        line.update(value=to_metric(total_ms, args.steps), ms_per_step=step_ms, dtype=model.act_dtype, clocks=clocks,
                    config={"workload": cfg["name"] % (B, model.act_dtype), "baseline_config": None, "batch_per_gpu": B,
                            "weights": "seeded re-init 1234 with scaled heads (demonet_b200/seeded.py)", "images": "torch.rand seed 1+rank",
                            "l2": "activations larger than L2 (%.0f MB after conv1_2)" % (B * S * S * 64 * 2 / 1e6),
                            "cuda_graph": False, "gflop_per_image": flops / 1e9},
                    gpu_launches=launches * args.steps, gpu_launches_per_step=launches)
        if not args.no_extras:
            e2e_ms, _ = timed(step_host, args.steps, warm, prime=6 * nslot)
            line["e2e"] = {"value": to_metric(e2e_ms, args.steps), "unit": cfg["unit"], "ms_per_step": e2e_ms / args.steps,
                           "h2d_bytes_per_step": B * 3 * S * S * 4, "d2h_bytes_per_step": B * D * 28 + B * 4,
                           "api": "pinned fp32 images in (H2D inside the timed region), layer-by-layer C-ABI calls, detections out"}
            line["roofline"] = {"kernel": "conv3x3_tc_kernel + pwconv_tc_kernel (whole forward)", "bound": "tensor", "achieved": tf,
                                "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf, "traffic": None,
                                "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained)",
                                "note": "algorithmic conv flops of the whole step over the whole step time (pools, L2-norm, "
                                        "post-processing included in the time)"}
    elif cfg["model"] == "post":
        # ---------------- config 4: the post-processing entry point alone --------------------------------------
        P = 3234
        g = torch.Generator().manual_seed(7 + rank)
        logits_h = (torch.randn(B, P, K, generator=g) * 4.0).pin_memory()
        bbox_h = (torch.randn(B, P, 4, generator=g) * 1.5).pin_memory()
        logits, bbox = logits_h.to(dev), bbox_h.to(dev)
        anchors = torch.from_numpy(dplan.default_boxes(dplan.plan_ssdlite320_mobilenet_v3_large())).to(dev)
        prm = make_post_params(P, K, S, S, 0.001, 0.55, 400, D)
        lib = _C.lib()
        ws = torch.empty(lib.dn_postprocess_workspace_bytes(B, ctypes.byref(prm)), dtype=torch.uint8, device=dev)
        out = ddist.PackedDetections(B, D, dev)
        host = {k: torch.empty_like(v, device="cpu").pin_memory() for k, v in out.as_io().items()}
        st_logits, st_bbox = torch.empty_like(logits), torch.empty_like(bbox)

        def post(lg, bb):
            with torch.cuda.device(dev):
                _C.check(lib.dn_postprocess(lg.data_ptr(), bb.data_ptr(), anchors.data_ptr(), B, ctypes.byref(prm), ws.data_ptr(),
                                            ws.numel(), out.boxes.data_ptr(), out.scores.data_ptr(), out.labels.data_ptr(),
                                            out.counts.data_ptr(), stream.cuda_stream), lib)

        def step_device():
            post(logits, bbox)

        def step_host():          # pinned host head outputs in, detections out
            st_logits.copy_(logits_h, non_blocking=True)
            st_bbox.copy_(bbox_h, non_blocking=True)
            post(st_logits, st_bbox)
            for k2, v in out.as_io().items():
                host[k2].copy_(v, non_blocking=True)

        total_ms, clocks = timed(step_device, args.steps, warm, sampler)
        value = to_metric(total_ms, args.steps)
        alg = B * (P * K * 4 + P * 16) + P * 16
        line.update(value=value, ms_per_step=total_ms / args.steps, dtype="f32", clocks=clocks,
                    config={"workload": cfg["name"] % B, "baseline_config": args.config, "batch_per_gpu": B,
                            "l2": "inputs larger than L2 (%.0f MB of logits per step)" % (B * P * K * 4 / 1e6),
                            "detections_per_img_min": int(out.counts.min()), "images_per_s": world * B * args.steps / (total_ms * 1e-3)},
                    gpu_launches=14 * args.steps, gpu_launches_per_step=14)
        if not args.no_extras:
            e2e_ms, _ = timed(step_host, args.steps, warm)
            line["e2e"] = {"value": to_metric(e2e_ms, args.steps), "unit": cfg["unit"], "ms_per_step": e2e_ms / args.steps,
                           "h2d_bytes_per_step": B * P * (K + 4) * 4, "d2h_bytes_per_step": B * D * 28 + B * 4,
                           "api": "dn_postprocess with pinned host head outputs in (H2D inside the timed region), detections out"}
            ms3 = (ctypes.c_float * 3)()
            with torch.cuda.device(dev):
                _C.check(lib.dn_postprocess_profile(logits.data_ptr(), bbox.data_ptr(), anchors.data_ptr(), B, ctypes.byref(prm),
                                                    ws.data_ptr(), ws.numel(), out.boxes.data_ptr(), out.scores.data_ptr(),
                                                    out.labels.data_ptr(), out.counts.data_ptr(), 5, ms3, stream.cuda_stream), lib)
            step_ms = total_ms / args.steps
            line["roofline"] = {"kernel": "postprocess kernels (softmax_decode + class_sort + class_nms_warp/cta + merge_topd)",
                                "bound": "hbm", "achieved": alg / (step_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": alg / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                                "algorithmic_bytes_per_launch": alg, "peak_source": peaks["src"],
                                "phase_ms": {"softmax_decode_hist": ms3[0], "sort_nms_rounds": ms3[1], "merge": ms3[2]},
                                "note": "1.229 MB algorithmic bytes per image (SURVEY 8(d) config 4) over the whole step time"}
    else:
        # ---------------- configs 2 / 3 / 5: the whole detector -------------------------------------------------
        if cfg["model"] == "v3":
            model = demonet_b200.ssdlite320_mobilenet_v3_large(num_classes=K, pipeline_slots=args.pipeline, act_dtype=args.act_dtype)
        else:
            model = demonet_b200.ssd_lite_mobilenet_v2(image_size=S, num_classes=K, score_thresh=0.5,
                                                       pipeline_slots=args.pipeline, act_dtype=args.act_dtype)
        model.load_state_dict(weights.seeded_state_dict(model.state_dict()))
        model = model.to(dev)
        eng = model.reserve(B, dev)
        lib = eng._lib
        # inputs resident in HBM (value) and in pinned host memory (e2e); different images per rank
        imgs_host = weights.synthetic_images(B, S, seed=1 + rank).pin_memory()
        imgs = imgs_host.to(dev)
        # one packed output buffer per rank and slot so that the gather of a batch is ONE collective
        packed_det = [ddist.PackedDetections(B, D, dev) for _ in range(nslot)]
        io = [p.as_io() for p in packed_det]
        # N > 1: the detections of G consecutive batches are gathered by ONE all_gather_into_tensor.  A collective per step makes
        # every step a cross-rank barrier (the slowest rank of each step paces all of them: 0.968 efficiency at 8 GPUs in r01);
        # with G = 4 the ranks only meet every fourth step.  Batch j is copied (2 MB, device to device) into position j % 2G of a
        # ring, so the G batches of a group are contiguous and the next group fills the other half while this one is in flight.
        G = max(1, args.gather_every) if world > 1 else 1
        nb = packed_det[0].buffer.numel()
        ring = torch.zeros(2 * G * nb, dtype=torch.uint8, device=dev) if world > 1 else None
        gathered = torch.empty(world * G * nb, dtype=torch.uint8, device=dev) if world > 1 else None
        host_outs = [{"boxes": torch.empty(B, D, 4, dtype=torch.float32).pin_memory(),
                      "scores": torch.empty(B, D, dtype=torch.float32).pin_memory(),
                      "labels": torch.empty(B, D, dtype=torch.int64).pin_memory(),
                      "counts": torch.empty(B, dtype=torch.int32).pin_memory()} for _ in range(nslot)]
        tick = {"dev": 0, "calls": 0, "host": 0, "u8": 0, "gathers": 0, "base": 0}

        def collect(j, k):
            """batch j (computed on slot k, already joined) -> ring; close the group when it is complete"""
            if G == 1:
                # one collective per batch: gather straight from the slot's packed output buffer, no ring copy (the buffer is
                # rewritten by the forward issued n_slots calls later on the same stream order, i.e. behind this collective)
                dist.all_gather_into_tensor(gathered, packed_det[k].buffer)
                tick["gathers"] += 1
                return
            pos = j % (2 * G)
            ring[pos * nb:(pos + 1) * nb].copy_(packed_det[k].buffer, non_blocking=True)
            if (j + 1) % G == 0:
                g0 = (j + 1 - G) % (2 * G)
                dist.all_gather_into_tensor(gathered, ring[g0 * nb:(g0 + G) * nb])
                tick["gathers"] += 1

        def step_device():
            # forward of batch i on slot i % nslot; the previous batch (its forward is joined, this one keeps running) goes
            # into the gather ring
            i = tick["dev"]                 # logical batch index (position in the gather ring)
            k = tick["calls"] % nslot       # the engine goes round its instances call by call: keep output buffers paired with them
            tick["dev"] += 1
            tick["calls"] += 1
            eng.forward(imgs, io[k])
            if world > 1:
                if nslot == 1:
                    collect(i, 0)
                elif i - (nslot - 1) >= tick["base"]:      # every slot is busy: the oldest batch in flight is joined and handed over
                    eng.join_previous()
                    collect(i - (nslot - 1), (k + 1) % nslot)

        def finish_device():
            # drain inside the timed region: join the last forward, hand its batch over and gather the (partial) last group
            if nslot >= 2:
                eng.join()
            if world > 1:
                i = tick["dev"]
                if nslot >= 2:
                    for j in range(max(tick["base"], i - (nslot - 1)), i):
                        collect(j, (tick["calls"] - (i - j)) % nslot)
                if i % G != 0:
                    g0 = (i - i % G) % (2 * G)
                    dist.all_gather_into_tensor(gathered, ring[g0 * nb:(g0 + G) * nb])
                    tick["gathers"] += 1
                tick["base"] = tick["dev"] = (i + G - 1) // G * G      # the next timed region starts on a group boundary

        def step_host():
            k = tick["host"] % nslot
            tick["host"] += 1
            eng.forward_host(imgs_host, host_outs[k])   # results land in each rank's own host memory: no gather

        imgs_u8_host = (imgs_host * 255).round().to(torch.uint8).pin_memory()      # what an image decoder hands over

        def step_host_u8():
            k = tick["u8"] % nslot
            tick["u8"] += 1
            eng.forward_host_u8(imgs_u8_host, host_outs[k])

        total_ms, clocks = timed(step_device, args.steps, warm, sampler, finish_device, prime=3 * nslot)
        n_launch = eng.launches_per_forward
        st = eng.stats()
        line.update(value=to_metric(total_ms, args.steps), ms_per_step=total_ms / args.steps, dtype=model.act_dtype, clocks=clocks,
                    config={"workload": cfg["name"] % (B, model.act_dtype), "baseline_config": args.config, "batch_per_gpu": B,
                            "global_batch": B * world,
                            "parallelism": "dp%d (image shards, no hot-path traffic; NCCL all-gather of the detections of every %d "
                                           "batches in one collective)" % (world, G),
                            "weights": "seeded re-init 1234 (demonet_b200/seeded.py)", "images": "torch.rand seed 1+rank",
                            "l2": "inputs larger than L2 (%.0f MB fp32 images per step; device memory %.1f GB)"
                                  % (B * 3 * S * S * 4 / 1e6, eng.device_bytes / 1e9),
                            "cuda_graph": True, "batches_in_flight": nslot,
                            "graph_priming": "%d untimed calls in front of the warm-up steps (eager, capture, first replay per engine "
                                             "instance: set-up, not counted as warm-up)" % (3 * nslot), "host_affinity": affinity_note, "engine": st},
                    gpu_launches=n_launch * args.steps, gpu_launches_per_step=n_launch)
        if not args.no_extras:
            e2e_ms, _ = timed(step_host, args.steps, warm, prime=6 * nslot)
            counts_ok = all(int(h["counts"].min()) >= 0 for h in host_outs)
            u8_ms, _ = timed(step_host_u8, args.steps, warm, prime=6 * nslot)
            line["e2e"] = {"value": to_metric(e2e_ms, args.steps), "unit": cfg["unit"], "ms_per_step": e2e_ms / args.steps,
                           "h2d_bytes_per_step": B * 3 * S * S * 4, "d2h_bytes_per_step": B * D * 28 + B * 4,
                           "api": "dn_engine_forward_host (pinned fp32 images in, detections out: the reference's input contract)",
                           "ok": counts_ok,
                           "u8": {"value": to_metric(u8_ms, args.steps), "unit": cfg["unit"], "ms_per_step": u8_ms / args.steps,
                                  "h2d_bytes_per_step": B * 3 * S * S, "d2h_bytes_per_step": B * D * 28 + B * 4,
                                  "api": "dn_engine_forward_host_u8 (pinned uint8 pixels in, x/255 on the device, detections out)"}}
            line["e2e_u8"] = line["e2e"]["u8"]
            # the reference's Python API itself: model(List[Tensor[3,S,S]]) -> List[Dict], one call (and one host sync) per step
            img_list = [t.clone() for t in imgs]            # B separate device tensors, as a data loader hands them over
            holder = {}

            def step_api():
                holder["d"] = model(img_list)
            api_ms, _ = timed(step_api, args.steps, warm, prime=3 * nslot)
            one = demonet_b200.ssdlite320_mobilenet_v3_large(num_classes=K, act_dtype=args.act_dtype) if cfg["model"] == "v3" else \
                demonet_b200.ssd_lite_mobilenet_v2(image_size=S, num_classes=K, score_thresh=0.5, act_dtype=args.act_dtype)
            one.load_state_dict(weights.seeded_state_dict(one.state_dict()))
            one = one.to(dev)
            eng1 = one.reserve(B, dev)
            io1 = ddist.PackedDetections(B, D, dev).as_io()

            def step_plain():
                eng1.forward(imgs, io1)
                stream.synchronize()
            plain_ms, _ = timed(step_plain, args.steps, warm)
            line["api_list"] = {"value": to_metric(api_ms, args.steps), "unit": cfg["unit"], "ms_per_step": api_ms / args.steps,
                                "api": "SSDLiteB200.forward(List[Tensor]) -> List[Dict] on %d separate CUDA tensors (one call + one host "
                                       "sync per step, like the reference's model(images))" % B,
                                "engine_forward_synchronous": {"value": to_metric(plain_ms, args.steps), "ms_per_step": plain_ms / args.steps,
                                                               "api": "dn_engine_forward + stream sync per step, one batch in flight"},
                                "detections_first_image": int(holder["d"][0]["scores"].numel())}
            del one, eng1
            # per-launch device times measured inside a CUDA graph (event-record nodes between consecutive ops)
            ms = (ctypes.c_float * (len(model.plan.layers) + 3))()
            with torch.cuda.device(dev):
                _C.check(lib.dn_engine_profile(eng._handle, imgs.data_ptr(), B, 10, ms, stream.cuda_stream), lib)
            costs = layer_costs(model.plan, B, D)
            per_kernel = {}
            for (name, nbytes, flops), t in zip(costs, list(ms)):
                if name.startswith("("):
                    continue
                k = per_kernel.setdefault(name, {"ms": 0.0, "bytes": 0, "flops": 0, "launches": 0})
                k["ms"] += t; k["bytes"] += nbytes; k["flops"] += flops; k["launches"] += 1
            sum_ms = sum(k["ms"] for k in per_kernel.values())
            dominant = max(per_kernel, key=lambda n: per_kernel[n]["ms"])
            dk = per_kernel[dominant]
            achieved = dk["bytes"] / (dk["ms"] * 1e-3) / 1e9
            traffic, traffic_src = ncu_traffic(dominant, B)
            total_bytes = sum(k["bytes"] for k in per_kernel.values())
            line["roofline"] = {
                "kernel": dominant, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write)",
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": dk["bytes"] / dk["launches"],
                "peak_source": peaks["src"], "launches_per_step": dk["launches"], "avg_launch_ms": dk["ms"] / dk["launches"],
                "share_of_step": dk["ms"] / sum_ms,
                "tensor_tflops": dk["flops"] / (dk["ms"] * 1e-3) / 1e12 if dk["flops"] else None,
                "timing": "in-graph: one CUDA graph of the plan on one stream with event-record nodes between consecutive ops, "
                          "10 replays, empty-bracket cost (%.1f us) subtracted (dn_engine_profile); sum over ops %.3f ms vs %.3f ms "
                          "per pipelined step" % (list(ms)[-1] * 1e3, sum_ms, total_ms / args.steps),
                "whole_step": {"algorithmic_GB": total_bytes / 1e9, "GBps": total_bytes / (total_ms / args.steps * 1e-3) / 1e9,
                               "frac": total_bytes / (total_ms / args.steps * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                "per_kernel": {n: {"ms": round(k["ms"], 4), "share": round(k["ms"] / sum_ms, 4),
                                   "GBps": round(k["bytes"] / (max(k["ms"], 1e-9) * 1e-3) / 1e9, 1),
                                   "frac_of_hbm_peak": round(k["bytes"] / (max(k["ms"], 1e-9) * 1e-3) / 1e9 / peaks["hbm_gbs"], 3),
                                   "TFLOPs": round(k["flops"] / (max(k["ms"], 1e-9) * 1e-3) / 1e12, 2), "launches": k["launches"]}
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

    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_extras:
        # bounded sample of the same workload on the host cores (BASELINE.md section 3 protocol, fewer timed steps)
        if orig_affinity:
            try:
                os.sched_setaffinity(0, orig_affinity)          # the CPU baseline gets every core again
            except Exception:
                pass
        n = args.cpu_sample or (4 if cfg["model"] in ("post", "v2", "vgg") else 32)
        line["cpu_baseline"] = cpu_baseline_block(args.config, n, 3, 1)
    if rank == 0:
        print(json.dumps(line), file=_claim_stdout(), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
