"""Batch sharding of the hot path across ranks (one process per GPU).

Images are independent in every stage, so the batch dimension is split into
contiguous shards with NO data-path collective; the only communication is the final gather of
the padded detections (``[B/G, K]`` scores / classes / boxes + ``count[B/G]``) and of the
per-image losses.  The functions work with any ``torch.distributed`` backend: NCCL over
NVLink on the GPU box, gloo on CPU in the tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of ``batch`` images over ``world`` ranks; the first ``batch % world``
    ranks take one extra image.  Returns [start, stop)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_levels(x, world: int, rank: int):
    """Slice (cls_list, cnt_list, reg_list) along the batch dimension for this rank."""
    lo, hi = shard_bounds(x[0][0].shape[0], world, rank)
    return [[t[lo:hi] for t in part] for part in x]


def gather_detections(scores: Tensor, classes: Tensor, boxes: Tensor, counts: Tensor, batch: int,
                      group=None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """All-gather padded per-shard detections back to the full batch, in image order.

    Shards may differ by one image, so every rank pads its shard to ``ceil(batch / world)`` rows;
    one collective per tensor, no host synchronisation.
    """
    world = dist.get_world_size(group)
    rows = -(-batch // world)

    def pad(t: Tensor) -> Tensor:
        if t.shape[0] == rows:
            return t.contiguous()
        out = t.new_zeros((rows,) + tuple(t.shape[1:]))
        out[: t.shape[0]] = t
        return out

    gathered = []
    for t in (scores, classes, boxes, counts):
        full = t.new_empty((world * rows,) + tuple(t.shape[1:]))
        if full.is_cuda:
            dist.all_gather_into_tensor(full, pad(t), group=group)
        else:                                             # gloo has no all_gather_into_tensor
            dist.all_gather(list(full.chunk(world, dim=0)), pad(t), group=group)
        keep = []
        for r in range(world):
            lo, hi = shard_bounds(batch, world, r)
            keep.append(full[r * rows: r * rows + (hi - lo)])
        gathered.append(torch.cat(keep, dim=0))
    return tuple(gathered)


def gather_packed(packed: Tensor, out: Tensor | None = None, group=None) -> Tensor:
    """ONE collective for a whole shard: all-gather the packed detection buffer of
    ``ops.postprocess`` (scores, boxes, classes, keep, counts share one allocation: the scores
    tensor's storage).  Every rank must hold the same shard size.  Returns ``[world, bytes]`` uint8;
    ``ops.detection_views(out[r], batch_per_rank, k)`` gives rank r's typed tensors."""
    world = dist.get_world_size(group)
    if out is None:
        out = packed.new_empty((world, packed.numel()))
    if packed.is_cuda:
        dist.all_gather_into_tensor(out, packed, group=group)
    else:
        dist.all_gather(list(out.unbind(0)), packed, group=group)
    return out


class PeerGather:
    """The final detection gather without a collective kernel: every rank PUSHES its packed buffer into its
    slot of each peer's receive buffer through NVLink peer memory (``cudaMemcpyAsync`` onto symmetric memory,
    i.e. the copy engines — no SM is taken from the HBM-bound kernels that keep running), then one small
    device-side barrier on the signal pads tells every rank that all slots have landed.

    Measured against NCCL ``all_gather`` in ``bench.py`` (the NCCL kernel's duration is exposed: its CTAs wait
    for SM resources behind the saturating post-process kernels).  ``slots`` receive buffers rotate so that a
    consumer may read slot q while later gathers fill the others (hence ``slots >= 2``); consume a slot in
    stream order before ``gather`` is called ``slots - 1`` more times.  Needs P2P-capable GPUs of one node (NVLink / NVSwitch) and a
    ``torch.distributed`` process group with one rank per GPU; raises if symmetric memory is unavailable
    (callers fall back to ``gather_packed``).  All ranks must call ``gather`` in the same order.
    """

    def __init__(self, nbytes: int, slots: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.nbytes, self.slots = int(nbytes), int(slots)
        if self.slots < 2:
            # with one slot a peer's next push is ordered after this rank REACHED the previous barrier, not after it
            # finished reading the slot: the returned view could be overwritten while it is read
            raise ValueError("PeerGather needs slots >= 2 (a slot is reused only after `slots` further gathers)")
        shape = (self.slots, self.world, self.nbytes)
        self.recv = symm_mem.empty(shape, dtype=torch.uint8, device=device)
        self.handle = symm_mem.rendezvous(self.recv, self.group)
        self.peers = [self.handle.get_buffer(p, shape, torch.uint8) for p in range(self.world)]
        self.handle.barrier(channel=0)

    def gather(self, slot: int, packed: Tensor, root: int | None = None) -> Tensor:
        """Push ``packed`` (uint8 [nbytes], on this device) into ``slot`` of every rank (``root=None``, an
        all-gather) or of rank ``root`` only (a gather: the other ranks send one buffer and receive nothing).
        Returns this rank's ``[world, nbytes]`` view of the slot, complete — on the receiving ranks — once the
        current stream reaches this point."""
        if packed.numel() != self.nbytes or packed.dtype != torch.uint8:
            raise ValueError("PeerGather.gather: packed must be uint8 of the size given at construction")
        if root is None:
            for k in range(self.world):                    # start with the neighbour: no two ranks hit one peer first
                p = (self.rank + 1 + k) % self.world
                self.peers[p][slot, self.rank].copy_(packed, non_blocking=True)
        else:
            self.peers[root][slot, self.rank].copy_(packed, non_blocking=True)
        self.handle.barrier(channel=0)                     # also keeps the ranks within `slots` gathers of each other
        return self.recv[slot]


def packed_of(scores: Tensor) -> Tensor:
    """The packed uint8 buffer behind the outputs of ``ops.postprocess`` / ``FCOSHead.detect``: it starts at the
    scores tensor ([B, K] fp32, the first member of the layout) and spans ``ops.packed_nbytes(B, K)`` bytes — also
    when the caller placed it inside a larger allocation (``out_packed`` slices)."""
    from . import ops
    batch, k = scores.shape
    st = scores.untyped_storage()
    start = scores.storage_offset() * scores.element_size()
    n = ops.packed_nbytes(batch, k)
    if start + n > st.nbytes():
        raise ValueError("scores is not the first member of a packed detection buffer")
    return torch.empty(0, dtype=torch.uint8, device=scores.device).set_(st, start, (n,))


def reduce_image_losses(per_image: Sequence[Tensor], batch: int, group=None) -> List[Tensor]:
    """Batch means of per-image losses computed on shards: sum(shard) / batch, all-reduced.

    Exact with respect to the reference's ``.mean()`` over the batch up to fp32 summation order,
    because every loss is normalised per image before the mean (loss.py:26,57,139,209-213).
    """
    sums = torch.stack([t.to(torch.float32).sum() for t in per_image]) / batch      # ONE collective for all losses
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return list(sums.unbind(0))
