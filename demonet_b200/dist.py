"""Multi-GPU plumbing for the SSDLite hot path: shard the image batch, gather the detections.

The path shards by image with no cross-image dependency; the reference's analog is a
`DistributedSampler` over the eval set plus a pickled `all_gather` of per-rank results
(demonet/train.py:125, demonet/util/misc.py:75-115).  Here every rank produces fixed-shape padded detections,
packed into ONE contiguous byte buffer, and a single `all_gather_into_tensor` closes the step -- no pickling,
no size exchange, no host round trip.  Works with NCCL (GPU) and gloo (CPU, used by the tests).
"""
from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a global batch owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside [0, %d)" % (rank, world))
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class PackedDetections:
    """boxes f32[B,D,4] | scores f32[B,D] | labels i64[B,D] | counts i32[B] in one uint8 buffer
    (every section 8-byte aligned so that the typed views are legal)."""

    def __init__(self, batch: int, detections_per_img: int, device):
        B, D = batch, detections_per_img
        self.batch, self.D = B, D
        sizes = [B * D * 16, B * D * 4, B * D * 8, B * 4]
        self.offsets, off = [], 0
        for s in sizes:
            self.offsets.append(off)
            off += (s + 7) // 8 * 8
        self.nbytes = off
        self.buffer = torch.zeros(off, dtype=torch.uint8, device=device)
        self.boxes, self.scores, self.labels, self.counts = self.views(self.buffer)

    def views(self, buf: torch.Tensor):
        B, D, o = self.batch, self.D, self.offsets
        return (buf[o[0]:o[0] + B * D * 16].view(torch.float32).view(B, D, 4),
                buf[o[1]:o[1] + B * D * 4].view(torch.float32).view(B, D),
                buf[o[2]:o[2] + B * D * 8].view(torch.int64).view(B, D),
                buf[o[3]:o[3] + B * 4].view(torch.int32))

    def as_io(self) -> Dict[str, torch.Tensor]:
        return {"boxes": self.boxes, "scores": self.scores, "labels": self.labels, "counts": self.counts}


def gather_detections(packed: PackedDetections, out: torch.Tensor = None, group=None) -> torch.Tensor:
    """All-gather the packed detections of every rank (equal per-rank batch).  Returns the
    [world * nbytes] uint8 buffer; `unpack_gathered` turns it into per-image dicts in global image order."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty(world * packed.nbytes, dtype=torch.uint8, device=packed.buffer.device)
    dist.all_gather_into_tensor(out, packed.buffer, group=group)
    return out


def unpack_gathered(packed: PackedDetections, gathered: torch.Tensor) -> List[Dict[str, torch.Tensor]]:
    world = gathered.numel() // packed.nbytes
    dets = []
    for r in range(world):
        boxes, scores, labels, counts = packed.views(gathered[r * packed.nbytes:(r + 1) * packed.nbytes])
        for i, n in enumerate(counts.tolist()):
            dets.append({"boxes": boxes[i, :n], "scores": scores[i, :n], "labels": labels[i, :n]})
    return dets


def gather_image_ids(image_ids: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather the per-rank image ids (int64 [B], equal B per rank) -> [world * B] in rank-major order, the order of
    `unpack_gathered`.  The reference pickles python lists through all_gather (coco_eval.py:167-186, misc.py:75-115)."""
    world = dist.get_world_size(group)
    ids = image_ids.to(torch.int64).contiguous()
    out = torch.empty(world * ids.numel(), dtype=torch.int64, device=ids.device)
    dist.all_gather_into_tensor(out, ids, group=group)
    return out


def first_occurrence_sorted(image_ids: torch.Tensor) -> torch.Tensor:
    """Indices that keep every image once, in ascending id order, choosing the first occurrence -- np.unique(ids,
    return_index=True) as used by coco_eval.merge (coco_eval.py:181-183) to drop the images a DistributedSampler
    repeats to pad the last batch."""
    ids = image_ids.detach().cpu()
    order = torch.argsort(ids, stable=True)
    sorted_ids = ids[order]
    keep = torch.ones_like(sorted_ids, dtype=torch.bool)
    keep[1:] = sorted_ids[1:] != sorted_ids[:-1]
    return order[keep]
