"""How much of a pipelined step is the post-processing?  Same model, score_thresh 0.001 (default) vs 0.999 (no candidates: the
sort / NMS / merge kernels find nothing to do), two batches in flight, device-resident inputs."""
import sys, time
import torch
sys.path.insert(0, ".")
import demonet_b200
from demonet_b200 import seeded

B = 256
imgs = [seeded.synthetic_images(B, 320).cuda() for _ in range(2)]
for thr in (0.001, 0.999):
    m = demonet_b200.ssdlite320_mobilenet_v3_large(pipeline_slots=2, score_thresh=thr)
    m.load_state_dict(seeded.seeded_state_dict(m.state_dict()))
    m = m.cuda()
    def run(n):
        for _ in m.forward_batches(imgs[i & 1] for i in range(n)):
            pass
    run(6)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(40); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 40
    print("score_thresh %.3f: %.3f ms / step, %.0f img/s" % (thr, ms, B / ms * 1e3))
