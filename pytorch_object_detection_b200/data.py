"""Device-side mirror of the datasets' ``collate_fn`` (``dataset/voc.py:141-173``, ``dataset/coco.py:135-165``).

The reference pads every image and every GT list on the host (B ``F.pad`` calls + ``Normalize`` + three
``torch.stack``) and uploads the padded batch.  Here the ragged data goes up as it is — ONE pinned staging
buffer for all GT rows, the unpadded images — and two kernels write the padded, normalised batch where
``FCOSGenTargets`` / ``FCOSTargetLoss`` read it.  Same return contract: ``(batch_imgs [B,3,H,W] f32,
batch_boxes [B,M,4] f32 padded with -1, batch_classes [B,M] i64 padded with -1)``.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import torch

from . import _lib
from .ops import _count, _need_cuda, _stream

Tensor = torch.Tensor


def pack_gt(boxes_list: Sequence[Tensor], classes_list: Sequence[Tensor], device) -> Tuple[Tensor, Tensor]:
    """Ragged per-image ``boxes [n_i,4]`` / ``classes [n_i]`` -> ``[B,M,4]`` f32 and ``[B,M]`` i64, M = max n_i,
    padded with -1.  CPU inputs travel in one pinned buffer and one H2D copy."""
    lib = _lib.load()
    assert len(boxes_list) == len(classes_list) and len(boxes_list) > 0
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.B200DetError("pack_gt writes a CUDA batch: b200det has no CPU path")
    counts = [int(b.shape[0]) for b in boxes_list]
    for b, c, n in zip(boxes_list, classes_list, counts):
        if (n and (b.dim() != 2 or b.shape[1] != 4)) or c.numel() != n:
            raise _lib.B200DetError("expected boxes [n,4] and classes [n] per image")
    batch, total, m = len(counts), sum(counts), max(counts)
    gt_boxes = torch.empty((batch, m, 4), dtype=torch.float32, device=dev)
    gt_labels = torch.empty((batch, m), dtype=torch.int64, device=dev)
    if m == 0:
        return gt_boxes, gt_labels
    offs = [0]
    for n in counts:
        offs.append(offs[-1] + n)
    # staging layout: boxes [total,4] f32 | labels [total] i64 | offsets [batch+1] i32, each 16-byte aligned
    o_lab = (total * 16 + 15) // 16 * 16
    o_off = o_lab + (total * 8 + 15) // 16 * 16
    nbytes = o_off + (batch + 1) * 4
    if all(b.is_cuda for b in boxes_list) and all(c.is_cuda for c in classes_list):
        stage = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        stage[:total * 16].view(torch.float32).view(total, 4).copy_(
            torch.cat([b.reshape(-1, 4).to(torch.float32) for b in boxes_list]))
        stage[o_lab:o_lab + total * 8].view(torch.int64).copy_(torch.cat([c.reshape(-1).to(torch.int64) for c in classes_list]))
        stage[o_off:].view(torch.int32).copy_(torch.tensor(offs, dtype=torch.int32), non_blocking=True)
    else:
        host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        torch.cat([b.reshape(-1, 4).to(torch.float32) for b in boxes_list], out=host[:total * 16].view(torch.float32).view(total, 4))
        torch.cat([c.reshape(-1).to(torch.int64) for c in classes_list], out=host[o_lab:o_lab + total * 8].view(torch.int64))
        host[o_off:].view(torch.int32).copy_(torch.tensor(offs, dtype=torch.int32))
        stage = host.to(dev, non_blocking=True)
    base = stage.data_ptr()
    with torch.cuda.device(dev):
        rc = lib.b200det_pack_gt(base, base + o_lab, base + o_off, batch, m, gt_boxes.data_ptr(), gt_labels.data_ptr(),
                                 _stream(stage))
    _lib.check(rc, "b200det_pack_gt")
    _count("pack_gt")
    return gt_boxes, gt_labels


def collate_images(imgs_list: Sequence[Tensor], mean: Sequence[float], std: Sequence[float]) -> Tensor:
    """CUDA images ``[C,h_i,w_i]`` f32 -> ``[B,C,H,W]`` with H, W the batch maxima: zero pad, then
    ``transforms.Normalize(mean, std)`` (padded pixels become ``-mean/std``, as in the reference)."""
    lib = _lib.load()
    assert len(imgs_list) > 0
    imgs = []
    for t in imgs_list:
        _need_cuda(t, "image")
        if t.dim() != 3 or t.shape[0] != len(mean) or len(mean) != len(std):
            raise _lib.B200DetError("expected images [C,h,w] with C = len(mean) = len(std)")
        imgs.append(t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous())
    dev = imgs[0].device
    batch, ch = len(imgs), imgs[0].shape[0]
    out_h = max(int(t.shape[1]) for t in imgs)
    out_w = max(int(t.shape[2]) for t in imgs)
    out = torch.empty((batch, ch, out_h, out_w), dtype=torch.float32, device=dev)
    ptrs = (C.c_void_p * batch)(*[t.data_ptr() for t in imgs])
    hw = (C.c_int32 * (2 * batch))(*[v for t in imgs for v in (int(t.shape[1]), int(t.shape[2]))])
    mean_a = (C.c_float * ch)(*[float(v) for v in mean])
    std_a = (C.c_float * ch)(*[float(v) for v in std])
    with torch.cuda.device(dev):
        rc = lib.b200det_collate_images(ptrs, hw, batch, ch, out_h, out_w, mean_a, std_a, out.data_ptr(), _stream(out))
    _lib.check(rc, "b200det_collate_images")
    _count("collate_images")
    return out


class DeviceCollate:
    """``collate_fn(data)`` of the reference's datasets with the batch assembled on the device.

    ``data`` is what a ``DataLoader`` hands over: a list of ``(img [3,h,w] f32, boxes [n,4] f32, classes [n] i64)``.

    It launches CUDA work, so it must run in the MAIN process: ``DataLoader(..., num_workers=0, pin_memory=False,
    collate_fn=DeviceCollate(...))`` (the batch it returns is already on the device).  With worker processes
    (the reference uses ``num_workers=4``, train.py:79-86) keep decoding in the workers with a trivial
    ``collate_fn=lambda data: data`` and call ``DeviceCollate`` on the list the loader yields; called inside a
    worker it raises instead of failing with "Cannot re-initialize CUDA in forked subprocess".
    """

    def __init__(self, mean: Sequence[float], std: Sequence[float], device="cuda"):
        self.mean, self.std, self.device = list(mean), list(std), torch.device(device)

    def __call__(self, data):
        import torch.utils.data as tud
        if tud.get_worker_info() is not None:
            raise _lib.B200DetError(
                "DeviceCollate launches CUDA kernels and cannot run inside a DataLoader worker process: use "
                "num_workers=0 (and pin_memory=False), or collate_fn=lambda d: d in the workers and call "
                "DeviceCollate on the yielded list in the main process")
        imgs_list, boxes_list, classes_list = zip(*data)
        assert len(imgs_list) == len(boxes_list) == len(classes_list)          # voc.py:143
        imgs = [t if t.is_cuda else t.pin_memory().to(self.device, non_blocking=True) for t in imgs_list]
        batch_imgs = collate_images(imgs, self.mean, self.std)
        batch_boxes, batch_classes = pack_gt(boxes_list, classes_list, self.device)
        return batch_imgs, batch_boxes, batch_classes
