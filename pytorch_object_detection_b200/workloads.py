"""Seeded synthetic inputs for the hot path.

Shapes follow the reference's networks: HISFCOS / MNFCOS produce five NCHW maps at
strides 8..128 (``model/od/HISFcos.py:131-136,152-153``), COCO images are resized to
800x1333 and padded to 832x1344 (``dataset/coco.py:110-114``), VOC to 512x512.
The distributions are the survey's: ``cls ~ N(-4.595, 1)`` (prior-bias init,
``HISFcos.py:208``), ``cnt ~ N(0, 1)``, ``reg = exp(N(mu, sigma))``.
Everything is drawn from a CPU ``torch.Generator`` so the golden fixtures, the tests and
the bench see the same numbers.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

STRIDES = [8, 16, 32, 64, 128]
COCO_HW = (832, 1344)
VOC_HW = (512, 512)
COCO_LEVELS = [(104, 168), (52, 84), (26, 42), (13, 21), (6, 10)]      # P = 23265
VOC_LEVELS = [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4)]            # P = 5456
HISFCOS_RANGES = [[-1, 32], [32, 96], [96, 192], [192, 384], [384, 9999999]]   # config/coco.yaml:52-58
FCOS_RANGES = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, 999999]]     # train.py:98-99


def num_points(levels: Sequence[Tuple[int, int]]) -> int:
    return sum(h * w for h, w in levels)


def head_outputs(batch: int, num_classes: int, levels: Sequence[Tuple[int, int]], seed: int,
                 crowded: bool = False, dtype=torch.float32):
    """(cls_list, cnt_list, reg_list) of contiguous NCHW CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    mu, sigma = (4.5, 0.5) if crowded else (3.0, 1.0)
    cls, cnt, reg = [], [], []
    for h, w in levels:
        cls.append((torch.randn(batch, num_classes, h, w, generator=g) - 4.595).to(dtype))
        cnt.append(torch.randn(batch, 1, h, w, generator=g).to(dtype))
        reg.append(torch.exp(torch.randn(batch, 4, h, w, generator=g) * sigma + mu).to(dtype))
    return cls, cnt, reg


def saturate_logits(x, seed: int):
    """Class logits that collapse in the fp32 sigmoid (in place on ``x[0]``; returns x): scaled so that ~15 % of
    them exceed 16.64, where sigmoid rounds to exactly 1.0f, and — on a quarter of the points — one class copied from
    another a few ulps higher.  The reference's argmax over sigmoid(cls) (head.py:57-62) then returns the FIRST of
    the equal values, which an argmax over the logits would not."""
    g = torch.Generator().manual_seed(seed)
    for t in x[0]:
        t.mul_(8.0).add_(45.0)
        c = t.shape[1]
        src, dst = c - 2, 1                                    # a later class and an earlier one
        ulps = torch.randint(1, 4, t[:, 0].shape, generator=g)
        bumped = (t[:, src].contiguous().view(torch.int32) + ulps.to(torch.int32)).view(torch.float32)
        pick = torch.rand(t[:, 0].shape, generator=g) < 0.25
        t[:, src] = torch.where(pick & (t[:, src] > 0), bumped, t[:, src])
        t[:, dst] = torch.where(pick & (t[:, src] > 0), t[:, src].view(torch.int32).sub(ulps.to(torch.int32)).view(torch.float32), t[:, dst])
    return x


def gt_boxes(batch: int, max_gt: int, img_hw: Tuple[int, int], num_classes: int, seed: int):
    """GT boxes [B, M, 4] fp32 and labels [B, M] int64, padded with -1 like the
    reference's collate functions (``dataset/voc.py:164-167``, ``dataset/coco.py:157-158``)."""
    g = torch.Generator().manual_seed(seed)
    h, w = img_hw
    boxes = torch.full((batch, max_gt, 4), -1.0)
    labels = torch.full((batch, max_gt), -1, dtype=torch.int64)
    for b in range(batch):
        n = int(torch.randint(1, max_gt + 1, (1,), generator=g))
        cx = torch.rand(n, generator=g) * w
        cy = torch.rand(n, generator=g) * h
        bw = torch.exp(torch.rand(n, generator=g) * 5 + 2).clamp(max=w)
        bh = torch.exp(torch.rand(n, generator=g) * 5 + 2).clamp(max=h)
        x0 = (cx - bw / 2).clamp(0, w - 1)
        y0 = (cy - bh / 2).clamp(0, h - 1)
        x1 = (cx + bw / 2).clamp(0, w - 1)
        y1 = (cy + bh / 2).clamp(0, h - 1)
        boxes[b, :n] = torch.stack([x0, y0, x1, y1], dim=-1)
        labels[b, :n] = torch.randint(1, num_classes + 1, (n,), generator=g)
    return boxes, labels


def crowd_candidates(n: int, num_classes: int, seed: int, clusters: int = 50, spread: float = 20.0,
                     img_hw: Tuple[int, int] = COCO_HW):
    """Dense-crowd NMS stage input (config 4): boxes from ``clusters`` tight clusters so that
    many pair IoUs straddle the threshold.  Returns boxes [n,4] f32, scores [n] f32 (unsorted),
    classes [n] int64 (1-based)."""
    g = torch.Generator().manual_seed(seed)
    h, w = img_hw
    ccx = torch.rand(clusters, generator=g) * w
    ccy = torch.rand(clusters, generator=g) * h
    cw = torch.exp(torch.rand(clusters, generator=g) * 2 + 3.5)
    ch = torch.exp(torch.rand(clusters, generator=g) * 2 + 3.5)
    ccls = torch.randint(1, num_classes + 1, (clusters,), generator=g)
    which = torch.randint(0, clusters, (n,), generator=g)
    cx = ccx[which] + torch.randn(n, generator=g) * spread
    cy = ccy[which] + torch.randn(n, generator=g) * spread
    bw = cw[which] * torch.exp(torch.randn(n, generator=g) * 0.15)
    bh = ch[which] * torch.exp(torch.randn(n, generator=g) * 0.15)
    boxes = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], dim=-1)
    # 80 % of a cluster shares its class, the rest is random: same-place, different-class pairs
    rnd = torch.randint(1, num_classes + 1, (n,), generator=g)
    classes = torch.where(torch.rand(n, generator=g) < 0.8, ccls[which], rnd)
    scores = torch.rand(n, generator=g) * 0.9 + 0.05
    return boxes.contiguous(), scores.contiguous(), classes.contiguous()


def collate_case(seed: int, sizes: Sequence[Tuple[int, int]], counts: Sequence[int]):
    """What a DataLoader hands to ``collate_fn``: a list of (img [3,h,w] f32 in [0,1), boxes [n,4] f32, classes [n] i64)."""
    g = torch.Generator().manual_seed(seed)
    data = []
    for (h, w), n in zip(sizes, counts):
        data.append((torch.rand(3, h, w, generator=g), torch.rand(n, 4, generator=g) * 50,
                     torch.randint(1, 21, (n,), generator=g)))
    return data


def fingerprint(tensors: Sequence[torch.Tensor]) -> List[float]:
    """Order-sensitive checksum used by the golden fixtures to detect RNG drift."""
    out = []
    for t in tensors:
        f = t.detach().double().reshape(-1)
        out.append(float((f * torch.arange(1, f.numel() + 1, dtype=torch.float64).remainder(97.0)).sum()))
    return out
