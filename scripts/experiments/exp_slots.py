"""Experiment: more than two batches in flight (n engines x two slots each, called round-robin)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import demonet_b200
from demonet_b200 import dist as ddist, seeded as weights

B, S, K, D = 256, 320, 91, 300
dev = torch.device("cuda:0")
for n_eng, slots in [(1, 4), (2, 3), (2, 4), (3, 4), (1, 4)]:
    engs, ios = [], []
    for _ in range(n_eng):
        m = demonet_b200.ssdlite320_mobilenet_v3_large(num_classes=K, pipeline_slots=slots)
        m.load_state_dict(weights.seeded_state_dict(m.state_dict()))
        m = m.to(dev)
        e = m.reserve(B, dev)
        engs.append((m, e))
        ios.append([ddist.PackedDetections(B, D, dev).as_io() for _ in range(max(slots, 1))])
    imgs = weights.synthetic_images(B, S, seed=1).to(dev)
    streams = [torch.cuda.Stream() for _ in range(n_eng)]
    calls = [0] * n_eng

    def step(i):
        j = i % n_eng
        with torch.cuda.stream(streams[j]):
            engs[j][1].forward(imgs, ios[j][calls[j] % max(slots, 1)])
        calls[j] += 1

    def drain():
        for j in range(n_eng):
            with torch.cuda.stream(streams[j]):
                if slots >= 2:
                    engs[j][1].join()
        torch.cuda.synchronize()

    for i in range(12):
        step(i)
    drain()
    steps = 60
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    for i in range(steps):
        step(i)
    for j in range(n_eng):
        with torch.cuda.stream(streams[j]):
            if slots >= 2:
                engs[j][1].join()
        torch.cuda.current_stream().wait_stream(streams[j])
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    print("engines %d x slots %d: %.3f ms / step, %.0f img/s" % (n_eng, slots, ms, B / ms * 1e3), flush=True)
    del engs, ios
    torch.cuda.empty_cache()
