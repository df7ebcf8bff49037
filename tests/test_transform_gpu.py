"""Transforms either side of the path on the GPU against torch on the CPU:
bilinear resize = interpolate(align_corners=False) as called by transform.py:27-53, the ToTensor conversion
of the uint8 ingest, resize_boxes (transform.py:278-292), and the module paths that use them."""
import pytest
import torch
import torch.nn.functional as F

import demonet_b200
from demonet_b200 import ops
from oracle import weights

pytestmark = pytest.mark.gpu


def _model():
    m = demonet_b200.ssdlite320_mobilenet_v3_large()
    m.load_state_dict(weights.seeded_state_dict(m.state_dict()))
    return m.cuda()


@pytest.mark.parametrize("hw,size", [((256, 275), (320, 320)), ((480, 640), (320, 320)), ((1080, 1920), (320, 320)),
                                     ((100, 37), (320, 320)), ((320, 320), (320, 320)), ((333, 500), (512, 512)),
                                     ((1, 1), (8, 8)), ((7, 5), (3, 2))])
def test_resize_bilinear_matches_interpolate(hw, size):
    g = torch.Generator().manual_seed(hw[0] * 7 + hw[1])
    img = torch.rand(3, *hw, generator=g)
    want = F.interpolate(img[None], size=size, mode="bilinear", align_corners=False)[0]
    got = ops.resize_bilinear(img.cuda(), size).cpu()
    assert got.shape == want.shape
    # fp32 tolerance: ATen's CPU kernel groups the four products differently (values are in [0, 1])
    assert float((got - want).abs().max()) <= 2e-6
    if hw == size:
        assert torch.equal(got, img)              # the identity case the synthetic configs rely on
    u8 = (img * 255).round().to(torch.uint8)
    want8 = F.interpolate((u8.float() / 255)[None], size=size, mode="bilinear", align_corners=False)[0]
    got8 = ops.resize_bilinear(u8.cuda(), size).cpu()
    assert float((got8 - want8).abs().max()) <= 2e-6


@pytest.mark.parametrize("n,offset", [(0, 0), (1, 0), (15, 0), (16, 0), (3 * 320 * 320 * 2, 0), (1000003, 0), (4099, 1), (70000, 3)])
def test_u8_to_f32_is_exact(n, offset):
    g = torch.Generator().manual_seed(n + offset)
    base = torch.randint(0, 256, (n + offset,), dtype=torch.uint8, generator=g)
    src = base.cuda()[offset:]                    # offset > 0: source not 16-byte aligned
    got = ops.u8_to_f32(src).cpu()
    assert torch.equal(got, base[offset:].float() / 255)
    if n >= 256:
        assert torch.equal(ops.u8_to_f32(torch.arange(256, dtype=torch.uint8).cuda()).cpu(), torch.arange(256).float() / 255)


def test_rescale_boxes_matches_resize_boxes():
    g = torch.Generator().manual_seed(5)
    B, D = 7, 300
    boxes = torch.rand(B, D, 4, generator=g) * 320
    sizes = [(256, 275), (480, 640), (320, 320), (1080, 1920), (33, 1000), (320, 100), (719, 1279)]
    got = ops.rescale_boxes_(boxes.clone().cuda(), sizes, (320, 320)).cpu()
    for b, (oh, ow) in enumerate(sizes):          # transform.py:278-292, verbatim arithmetic
        rh = torch.tensor(oh, dtype=torch.float32) / torch.tensor(320, dtype=torch.float32)
        rw = torch.tensor(ow, dtype=torch.float32) / torch.tensor(320, dtype=torch.float32)
        x0, y0, x1, y1 = boxes[b].unbind(1)
        want = torch.stack((x0 * rw, y0 * rh, x1 * rw, y1 * rh), dim=1)
        assert torch.equal(got[b], want)


def test_forward_resizes_on_device_and_rescales_boxes():
    """model([img]) with a non-S x S image == resize kernel -> model -> resize_boxes, for CUDA and for CPU inputs."""
    model = _model()
    g = torch.Generator().manual_seed(3)
    imgs = [torch.rand(3, 256, 275, generator=g), torch.rand(3, 320, 320, generator=g), torch.rand(3, 400, 300, generator=g)]
    dets = model([i.cuda() for i in imgs])
    pre = [ops.resize_bilinear(i.cuda(), (320, 320)) for i in imgs]
    ref = model(pre)
    for d, r, i in zip(dets, ref, imgs):
        oh, ow = i.shape[-2:]
        rh = torch.tensor(oh, dtype=torch.float32) / torch.tensor(320, dtype=torch.float32)
        rw = torch.tensor(ow, dtype=torch.float32) / torch.tensor(320, dtype=torch.float32)
        want = r["boxes"].cpu() * torch.stack([rw, rh, rw, rh])
        assert torch.equal(d["boxes"].cpu(), want) and torch.equal(d["scores"], r["scores"]) and torch.equal(d["labels"], r["labels"])
    host = model(imgs)                            # CPU inputs of mixed sizes
    assert host[0]["boxes"].device.type == "cpu"
    for d, h in zip(dets, host):
        assert torch.equal(d["boxes"].cpu(), h["boxes"]) and torch.equal(d["scores"].cpu(), h["scores"])


def test_forward_uint8_equals_float_forward():
    model = _model()
    g = torch.Generator().manual_seed(11)
    u8 = torch.randint(0, 256, (5, 3, 320, 320), dtype=torch.uint8, generator=g)
    want = model(list((u8.float() / 255).cuda()))
    for got in (model.forward_uint8(u8.cuda()), model.forward_uint8(u8), model.forward_uint8(u8)):      # device, host, host again
        assert len(got) == 5
        for a, b in zip(got, want):
            assert torch.equal(a["scores"].cpu(), b["scores"].cpu()) and torch.equal(a["boxes"].cpu(), b["boxes"].cpu())
            assert torch.equal(a["labels"].cpu(), b["labels"].cpu())
    assert model.forward_uint8(u8)[0]["boxes"].device.type == "cpu"
    with pytest.raises(ValueError):
        model.forward_uint8(torch.zeros(2, 3, 300, 320, dtype=torch.uint8))
    with pytest.raises(TypeError):
        model([torch.zeros(3, 320, 320, dtype=torch.uint8).cuda()])          # the float contract of forward() is unchanged


def _reference_coco(predictions):
    """CocoEvaluator.prepare_for_coco_detection + convert_to_xywh, demonet/data/coco_eval.py:76-98,162-164 (restated:
    the module itself needs pycocotools to import)."""
    out = []
    for original_id, p in predictions.items():
        if len(p) == 0:
            continue
        xmin, ymin, xmax, ymax = p["boxes"].unbind(1)
        boxes = torch.stack((xmin, ymin, xmax - xmin, ymax - ymin), dim=1).tolist()
        scores, labels = p["scores"].tolist(), p["labels"].tolist()
        out.extend([{"image_id": original_id, "category_id": labels[k], "bbox": box, "score": scores[k]}
                    for k, box in enumerate(boxes)])
    return out


@pytest.mark.parametrize("B,D", [(1, 1), (7, 300), (300, 100), (1500, 5)])
def test_detections_to_coco_matches_reference(B, D):
    g = torch.Generator().manual_seed(B * 1000 + D)
    xy = torch.rand(B, D, 2, generator=g) * 300
    boxes = torch.cat([xy, xy + torch.rand(B, D, 2, generator=g) * 80], -1)
    scores = torch.rand(B, D, generator=g)
    labels = torch.randint(1, 91, (B, D), generator=g)
    counts = torch.randint(0, D + 1, (B,), generator=g, dtype=torch.int32)
    counts[0] = 0 if B > 1 else counts[0]
    ids = torch.randperm(100000, generator=g)[:B]
    rows = ops.detections_to_coco(boxes.cuda(), scores.cuda(), labels.cuda(), counts.cuda(), ids)
    got = ops.coco_results(rows)
    want = _reference_coco({int(ids[b]): {"boxes": boxes[b, :counts[b]], "scores": scores[b, :counts[b]],
                                         "labels": labels[b, :counts[b]]} for b in range(B)})
    assert got == want                        # bit-exact: same fp32 subtraction, same order


def test_model_detections_to_coco_rows():
    model = _model()
    x = weights.synthetic_images(3, 320).cuda()
    dets = model(list(x))
    eng = model._engine_for(x.device, 3)
    io = model._io_buffers(x.device, 3, False)
    rows = ops.detections_to_coco(io["boxes"], io["scores"], io["labels"], io["counts"], [11, 5, 7])
    want = _reference_coco({i: {k: v.cpu() for k, v in d.items()} for i, d in zip([11, 5, 7], dets)})
    assert ops.coco_results(rows) == want and eng is not None
