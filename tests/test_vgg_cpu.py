"""ssd300_vgg16 without a GPU: the oracle against the reference's golden outputs
(tests/golden/ssd300_vgg16.npz, produced by the unmodified reference through make_golden.py), the default-box table, and
the host side of the drop-in (state_dict contract, builder arguments)."""
import hashlib
import os

import numpy as np
import pytest
import torch

import demonet_b200
from demonet_b200 import plan as dplan
from oracle import boxes_np, net_ref, weights

AR, SC, ST = [[2], [2, 3], [2, 3], [2, 3], [2], [2]], [0.07, 0.15, 0.33, 0.51, 0.69, 0.87, 1.05], [8, 16, 32, 64, 100, 300]
GRIDS = [(38, 38), (19, 19), (10, 10), (5, 5), (3, 3), (1, 1)]


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_state_dict_contract_matches_the_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ssd300_vgg16.npz"))
    m = demonet_b200.ssd300_vgg16()
    sd = m.state_dict()
    assert sorted(sd.keys()) == sorted(g["state_dict_keys"].tolist()) and len(sd) == 71
    shapes = dict(zip(g["state_dict_keys"].tolist(), g["state_dict_shapes"].tolist()))
    assert all(str(tuple(v.shape)) == shapes[k] for k, v in sd.items())
    assert m.num_priors == 8732 and m.num_anchors == [4, 6, 6, 6, 4, 4]
    assert (m.score_thresh, m.nms_thresh, m.detections_per_img, m.topk_candidates) == (0.01, 0.45, 200, 400)
    with pytest.raises(TypeError):
        demonet_b200.ssd300_vgg16(bogus=1)
    m.train()
    with pytest.raises(ValueError, match="In training mode, targets should be passed"):
        m([torch.rand(3, 300, 300)])
    m.eval()
    m2 = demonet_b200.ssd300_vgg16(score_thresh=0.3, iou_thresh=0.4, positive_fraction=0.5)      # training-side kwargs of SSD.__init__
    assert (m2.score_thresh, m2.iou_thresh, m2.neg_to_pos_ratio) == (0.3, 0.4, 1.0)
    import hubconf
    assert hubconf.ssd300_vgg16 is demonet_b200.ssd300_vgg16


def test_default_boxes_with_scales_and_steps(golden_dir):
    g = np.load(os.path.join(golden_dir, "ssd300_vgg16.npz"))
    a = boxes_np.default_boxes(GRIDS, (300, 300), AR, scales=SC, steps=ST)
    b = dplan.default_boxes_for(GRIDS, 300, AR, scales=SC, steps=ST)
    assert a.shape == (8732, 4) and _sha(a) == str(g["anchors_sha256"]) and np.array_equal(a, b)


def test_oracle_reproduces_the_reference_outputs(golden_dir):
    g = np.load(os.path.join(golden_dir, "ssd300_vgg16.npz"))
    sd = weights.seeded_vgg_state_dict(demonet_b200.ssd300_vgg16().state_dict())
    x = weights.synthetic_images(2, 300)
    with torch.no_grad():
        cls, reg, grids = net_ref.vgg_forward_raw(sd, x, "fp32")
        cls16, _, _ = net_ref.vgg_forward_raw(sd, x, "fp16")
    assert grids == GRIDS and cls.shape == (2, 8732, 91) and reg.shape == (2, 8732, 4)
    stride = int(g["row_stride"])
    rel = lambda a, b: float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt())          # noqa: E731
    assert rel(cls[:, ::stride], torch.from_numpy(g["logits_rows"])) < 1e-5
    assert rel(reg[:, ::stride], torch.from_numpy(g["bbox_rows"])) < 1e-5
    assert rel(cls16[:, ::stride], torch.from_numpy(g["logits_fp16emu_rows"])) < 1e-3
    assert 1e-4 < rel(cls16, cls) < 5e-3                      # the fp16 contract is a real, small perturbation
    if _sha(cls.numpy()) == str(g["logits_sha256"]):          # bit-identical conv kernels on this host: detections must be too
        a = boxes_np.default_boxes(GRIDS, (300, 300), AR, scales=SC, steps=ST)
        sc = torch.softmax(cls, -1).numpy()
        for i in range(2):
            o = boxes_np.postprocess_detections(None, reg[i].numpy(), a, (300, 300), score_thresh=0.01, nms_thresh=0.45,
                                                detections_per_img=200, topk_candidates=400, scores=sc[i])
            assert np.array_equal(o["labels"], g["det_labels"][i]) and np.array_equal(o["scores"], g["det_scores"][i])
