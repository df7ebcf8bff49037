"""The N > 1 path on the CPU: world_size-2 gloo processes shard a batch and gather packed detections."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from demonet_b200 import dist as ddist


def test_shard_range():
    for gb, world in ((2048, 8), (10, 3), (1, 2), (7, 7)):
        spans = [ddist.shard_range(gb, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == gb
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        ddist.shard_range(8, 2, 2)


def _fake_detections(rank, B, D):
    g = torch.Generator().manual_seed(100 + rank)
    counts = torch.randint(0, D + 1, (B,), generator=g, dtype=torch.int32)
    counts[0] = 0                      # an image with no detections (test_onnx.py:125-133)
    return torch.rand(B, D, 4, generator=g), torch.rand(B, D, generator=g), \
        torch.randint(1, 91, (B, D), generator=g), counts


def _worker(rank, world, port, B, D, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        packed = ddist.PackedDetections(B, D, "cpu")
        boxes, scores, labels, counts = _fake_detections(rank, B, D)
        packed.boxes.copy_(boxes); packed.scores.copy_(scores); packed.labels.copy_(labels); packed.counts.copy_(counts)
        gathered = ddist.gather_detections(packed)
        dets = ddist.unpack_gathered(packed, gathered)
        ok = len(dets) == world * B
        for r in range(world):
            eb, es, el, ec = _fake_detections(r, B, D)
            for i in range(B):
                d = dets[r * B + i]
                n = int(ec[i])
                ok &= d["boxes"].shape == (n, 4) and torch.equal(d["boxes"], eb[i, :n])
                ok &= torch.equal(d["scores"], es[i, :n]) and torch.equal(d["labels"], el[i, :n])
        # image ids travel the same way; the sampler's padding duplicates are dropped like coco_eval.merge does
        ids = torch.tensor([rank * 3 + i for i in range(B)], dtype=torch.int64)       # ranks overlap on purpose
        all_ids = ddist.gather_image_ids(ids)
        ok &= all_ids.tolist() == [r * 3 + i for r in range(world) for i in range(B)]
        keep = ddist.first_occurrence_sorted(all_ids)
        import numpy as np
        want_ids, want_idx = np.unique(all_ids.numpy(), return_index=True)
        ok &= keep.tolist() == want_idx.tolist() and all_ids[keep].tolist() == want_ids.tolist()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_detections_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def test_first_occurrence_sorted_matches_numpy_unique():
    import numpy as np
    g = torch.Generator().manual_seed(3)
    for n in (0, 1, 17, 1000):
        ids = torch.randint(0, max(1, n // 2), (n,), generator=g, dtype=torch.int64)
        keep = ddist.first_occurrence_sorted(ids)
        want_ids, want_idx = np.unique(ids.numpy(), return_index=True)
        assert keep.tolist() == want_idx.tolist() and ids[keep].tolist() == want_ids.tolist()
