"""Experiment: two engines (own arenas + graphs) on two streams, batches alternating, vs one engine on one stream."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import demonet_b200
from demonet_b200 import seeded as weights
from demonet_b200 import dist as ddist
B, D, S = 256, 300, 320
dev = torch.device("cuda", 0)
models, engs, ios, imgs, streams = [], [], [], [], []
for i in range(3):
    m = demonet_b200.ssdlite320_mobilenet_v3_large(num_classes=91)
    m.load_state_dict(weights.seeded_state_dict(m.state_dict()))
    m = m.to(dev)
    models.append(m); engs.append(m.reserve(B, dev))
    ios.append(ddist.PackedDetections(B, D, dev).as_io())
    imgs.append(weights.synthetic_images(B, S, seed=1 + i).to(dev))
    streams.append(torch.cuda.Stream(dev))
def run(n_streams, steps=40, warm=6):
    def one(i):
        k = i % n_streams
        with torch.cuda.stream(streams[k]):
            engs[k].forward(imgs[k], ios[k])
    for i in range(warm): one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for st in streams[1:]:
        st.wait_event(e0)
    for i in range(steps): one(i)
    for st in streams[1:]:
        ev = torch.cuda.Event(); ev.record(st); streams[0].wait_event(ev)
    e1.record(streams[0])
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return ms, B / ms * 1e3
for n in (1, 2, 3, 2, 3):
    ms, rate = run(n)
    print("streams %d: %.3f ms/step  %.0f img/s" % (n, ms, rate))
