"""CPU restatement of the HISFCOS detection hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to /root/reference).
It is written with torch CPU tensor ops because the reference itself is torch: sigmoid,
sqrt, min/max(dim) and topk then round exactly like the reference's CPU path.  The NMS
arithmetic lives in torchvision 0.26.0 (third party, not in the reference tree); it is
restated here in numpy float32 (one IEEE op per step, no FMA) from the published
``nms_kernel_impl`` algorithm and pinned against the installed CPU op by the tests.

Differences from the reference that are deliberate and documented:
  * batch > 1 returns a ragged python list (the reference's final ``torch.stack`` raises,
    head.py:99-101); per image the arithmetic is identical.
  * target assignment is evaluated image by image ([HW, M] temporaries, not
    [B, HW, M, 4]); every element goes through the same fp32 operations.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor

AREA_SENTINEL = 99999999  # head.py:285
CPU_TRICK_MAX_NUMEL = 4000  # torchvision ops/boxes.py batched_nms, CPU threshold


# --------------------------------------------------------------------------------------
# grid / layout                                                        utill/utills.py:58-73
# --------------------------------------------------------------------------------------
def grid_points(h: int, w: int, stride: int, device=None) -> Tensor:
    """Row-major point grid (x, y) = (j*s + s//2, i*s + s//2) as fp32 [h*w, 2] (``device``: where the
    reference builds it, head.py:22 — only bench.py's informative eager-on-GPU leg passes a CUDA device)."""
    xs = torch.arange(0, w * stride, stride, dtype=torch.float32, device=device)
    ys = torch.arange(0, h * stride, stride, dtype=torch.float32, device=device)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([xx.reshape(-1), yy.reshape(-1)], dim=-1) + stride // 2


def flatten_levels(levels: Sequence[Tensor], strides: Sequence[int]) -> Tuple[Tensor, Tensor]:
    """NCHW level list -> ([B, P, C], [P, 2]); zip() truncation kept.      head.py:8-26"""
    flat, pts = [], []
    b, c = levels[0].shape[0], levels[0].shape[1]
    for lvl, s in zip(levels, strides):
        nhwc = lvl.permute(0, 2, 3, 1)
        pts.append(grid_points(nhwc.shape[1], nhwc.shape[2], s, lvl.device))
        flat.append(nhwc.reshape(b, -1, c))
    return torch.cat(flat, dim=1), torch.cat(pts, dim=0)


def decode_boxes(points: Tensor, ltrb: Tensor) -> Tensor:
    """(x - l, y - t, x + r, y + b); no stride factor.                      head.py:29-38"""
    return torch.cat([points[None] - ltrb[..., :2], points[None] + ltrb[..., 2:]], dim=-1)


# --------------------------------------------------------------------------------------
# inference head                                                          head.py:52-102
# --------------------------------------------------------------------------------------
def score_points(x, strides: Sequence[int]) -> Tuple[Tensor, Tensor, Tensor]:
    """Per-point score sqrt(max_c sig(cls) * sig(cnt)), class argmax+1, box.  head.py:53-66"""
    cls, pts = flatten_levels(x[0], strides)
    cnt, _ = flatten_levels(x[1], strides)
    reg, _ = flatten_levels(x[2], strides)
    best, arg = torch.max(torch.sigmoid(cls), dim=-1)
    score = torch.sqrt(best * torch.sigmoid(cnt).squeeze(-1))
    return score, arg + 1, decode_boxes(pts, reg)


def select_topk(score: Tensor, classes: Tensor, boxes: Tensor, max_box: int):
    """Global (all levels) per-image top-k, sorted descending.             head.py:69-80"""
    k = min(max_box, score.shape[-1])
    idx = torch.topk(score, k, dim=1, largest=True, sorted=True)[1]
    take = lambda t: torch.stack([t[b][idx[b]] for b in range(t.shape[0])], dim=0)
    return take(score), take(classes), take(boxes), idx


def nms_greedy(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    """torchvision 0.26.0 CPU ``nms_kernel_impl<float>`` restated.

    stable descending score order; area = (x2-x1)*(y2-y1); for each kept i, every later
    unsuppressed j with  inter / (area_i + area_j - inter) > (double)thr  is suppressed,
    inter = max(0, min(x2)-max(x1)) * max(0, min(y2)-max(y1)); fp32, one rounding per op.
    """
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-scores.astype(np.float64), kind="stable")  # stable, descending
    x1, y1, x2, y2 = (boxes[order, i] for i in range(4))
    area = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, dtype=bool)
    keep = []
    thr = float(iou_threshold)
    zero = np.float32(0)
    for i in range(n):
        if dead[i]:
            continue
        keep.append(order[i])
        if i + 1 == n:
            break
        w = np.maximum(zero, np.minimum(x2[i], x2[i + 1:]) - np.maximum(x1[i], x1[i + 1:]))
        h = np.maximum(zero, np.minimum(y2[i], y2[i + 1:]) - np.maximum(y1[i], y1[i + 1:]))
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (area[i] + area[i + 1:] - inter)
        dead[i + 1:] |= ovr.astype(np.float64) > thr
    return np.asarray(keep, dtype=np.int64)


def batched_nms(boxes: Tensor, scores: Tensor, classes: Tensor, iou_threshold: float) -> Tensor:
    """torchvision 0.26.0 ``batched_nms`` on CPU (call site head.py:94).

    numel <= 4000: coordinate trick — boxes + class * (max_coord + 1) in fp32, one NMS
    over all boxes (so cross-class suppression of negative-coordinate boxes is kept).
    numel  > 4000: per-class NMS on raw boxes, survivors re-sorted by score (unstable
    sort in torch; here: descending score, ascending index on ties).
    """
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    b = boxes.detach().to(torch.float32).numpy()
    s = scores.detach().to(torch.float32).numpy()
    c = classes.detach().numpy()
    if boxes.numel() <= CPU_TRICK_MAX_NUMEL:
        span = b.max() + np.float32(1)
        off = c.astype(np.float32) * span
        return torch.from_numpy(nms_greedy(b + off[:, None], s, iou_threshold))
    alive = np.zeros(b.shape[0], dtype=bool)
    for cid in np.unique(c):
        members = np.nonzero(c == cid)[0]
        alive[members[nms_greedy(b[members], s[members], iou_threshold)]] = True
    kept = np.nonzero(alive)[0]
    return torch.from_numpy(kept[np.argsort(-s[kept].astype(np.float64), kind="stable")])


def post_process_image(score_k: Tensor, class_k: Tensor, box_k: Tensor,
                       score_thr: float, nms_thr: float, nms_fn=batched_nms):
    """threshold -> batched NMS -> gather for ONE image.                   head.py:89-98"""
    m = score_k >= score_thr
    s, c, b = score_k[m], class_k[m], box_k[m]
    keep = nms_fn(b, s, c, nms_thr)
    return s[keep], c[keep], b[keep], keep


def detect(x, score_thr: float, nms_thr: float, max_box: int, strides: Sequence[int], nms_fn=batched_nms):
    """FCOSHead.forward for any batch size; returns a per-image list.      head.py:52-102"""
    score, classes, boxes = score_points(x, strides)
    s_k, c_k, b_k, _ = select_topk(score, classes, boxes, max_box)
    return [post_process_image(s_k[i], c_k[i], b_k[i], score_thr, nms_thr, nms_fn)[:3]
            for i in range(s_k.shape[0])]


def clip_boxes_(boxes: Tensor, img_h: int, img_w: int) -> Tensor:
    """In-place clamp to [0, w-1] x [0, h-1], returns the same tensor.     head.py:152-162"""
    boxes.clamp_(min=0)
    boxes[..., 0::2] = boxes[..., 0::2].clamp(max=img_w - 1)
    boxes[..., 1::2] = boxes[..., 1::2].clamp(max=img_h - 1)
    return boxes


# --------------------------------------------------------------------------------------
# training targets                                                        head.py:211-316
# --------------------------------------------------------------------------------------
def assign_level(h: int, w: int, gt: Tensor, labels: Tensor, stride: int,
                 lim: Sequence[float], radius: float = 1.5):
    """generate_target for one level (only the level's h, w are used).     head.py:235-316

    Returns cls_t [B, hw, 1] int64, cnt_t [B, hw, 1] f32, reg_t [B, hw, 4] f32 and, for
    tests, the chosen GT index [B, hw] (argmin of masked area, -1 where negative).
    """
    pts = grid_points(h, w, stride, gt.device)
    px, py = pts[:, 0:1], pts[:, 1:2]                                    # [hw, 1]
    cls_o, cnt_o, reg_o, idx_o = [], [], [], []
    for b in range(gt.shape[0]):
        g = gt[b]                                                        # [M, 4]
        l = px - g[None, :, 0]                                           # head.py:261-264
        t = py - g[None, :, 1]
        r = g[None, :, 2] - px
        bt = g[None, :, 3] - py
        off = torch.stack([l, t, r, bt], dim=-1)                         # [hw, M, 4]
        area = (l + r) * (t + bt)                                        # head.py:268
        omin = off.min(dim=-1)[0]
        omax = off.max(dim=-1)[0]
        in_box = omin > 0                                                # head.py:272
        in_lvl = (omax > lim[0]) & (omax <= lim[1])                      # head.py:273
        cx = (g[:, 0] + g[:, 2]) / 2                                     # head.py:276-277
        cy = (g[:, 1] + g[:, 3]) / 2
        cmax = torch.stack([px - cx[None], py - cy[None], cx[None] - px, cy[None] - py], -1).max(-1)[0]
        pos = in_box & in_lvl & (cmax < stride * radius)                 # head.py:275-283
        area = torch.where(pos, area, torch.full_like(area, AREA_SENTINEL))  # head.py:285
        pick = area.min(dim=-1)[1]                                       # first index on ties
        rows = torch.arange(pick.numel(), device=pick.device)
        reg = off[rows, pick]                                            # head.py:287-288
        cls = labels[b][pick]                                            # head.py:290-292
        lr_min, lr_max = torch.min(reg[:, 0], reg[:, 2]), torch.max(reg[:, 0], reg[:, 2])
        tb_min, tb_max = torch.min(reg[:, 1], reg[:, 3]), torch.max(reg[:, 1], reg[:, 3])
        cnt = ((lr_min * tb_min) / (lr_max * tb_max + 1e-10)).sqrt()     # head.py:298-299
        any_pos = pos.long().sum(-1) >= 1                                # head.py:308-310
        cls = torch.where(any_pos, cls, torch.zeros_like(cls))           # head.py:312-314
        cnt = torch.where(any_pos, cnt, torch.full_like(cnt, -1))
        reg = torch.where(any_pos[:, None], reg, torch.full_like(reg, -1))
        cls_o.append(cls[:, None]); cnt_o.append(cnt[:, None]); reg_o.append(reg)
        idx_o.append(torch.where(any_pos, pick, torch.full_like(pick, -1)))
    return torch.stack(cls_o), torch.stack(cnt_o), torch.stack(reg_o), torch.stack(idx_o)


def assign_targets(level_hw: Sequence[Tuple[int, int]], gt: Tensor, labels: Tensor,
                   strides: Sequence[int], limit_ranges: Sequence[Sequence[float]]):
    """FCOSGenTargets.forward: level-major concat of per-level targets.    head.py:218-232"""
    assert len(strides) == len(level_hw)
    parts = [assign_level(h, w, gt, labels, s, lim) for (h, w), s, lim in zip(level_hw, strides, limit_ranges)]
    return tuple(torch.cat([p[i] for p in parts], dim=1) for i in range(4))


# --------------------------------------------------------------------------------------
# losses                                                                  model/loss.py
# --------------------------------------------------------------------------------------
def _flat(levels: Sequence[Tensor], c: int) -> Tensor:
    b = levels[0].shape[0]
    return torch.cat([p.permute(0, 2, 3, 1).reshape(b, -1, c) for p in levels], dim=1)


def focal_sum(logits: Tensor, onehot: Tensor, gamma: float = 2.0, alpha: float = 0.25) -> Tensor:
    """loss.py:180-193 — sigmoid, clip to [5e-6, 0.99999999995], alpha/gamma focal, summed."""
    p = torch.clip(logits.sigmoid(), min=0.000005, max=0.99999999995)
    pt = p * onehot + (1.0 - p) * (1.0 - onehot)
    wgt = alpha * onehot + (1.0 - alpha) * (1.0 - onehot)
    return (-wgt * torch.pow(1.0 - pt, gamma) * pt.log()).sum()


def iou_sum(p: Tensor, t: Tensor) -> Tensor:
    """loss.py:142-152 — ltrb IoU, -log(clamp(iou, 1e-6)) summed."""
    wh = torch.clamp(torch.min(p[:, 2:], t[:, 2:]) + torch.min(p[:, :2], t[:, :2]), min=0)
    inter = wh[:, 0] * wh[:, 1]
    a1 = (p[:, 2] + p[:, 0]) * (p[:, 3] + p[:, 1])
    a2 = (t[:, 2] + t[:, 0]) * (t[:, 3] + t[:, 1])
    return (-(inter / (a1 + a2 - inter)).clamp(min=1e-6).log()).sum()


def giou_sum(p: Tensor, t: Tensor) -> Tensor:
    """loss.py:155-177 — 1 - (iou - (G - union) / clamp(G, 1e-10)) summed."""
    wh = torch.clamp(torch.min(p[:, 2:], t[:, 2:]) + torch.min(p[:, :2], t[:, :2]), min=0)
    inter = wh[:, 0] * wh[:, 1]
    a1 = (p[:, 2] + p[:, 0]) * (p[:, 3] + p[:, 1])
    a2 = (t[:, 2] + t[:, 0]) * (t[:, 3] + t[:, 1])
    union = a1 + a2 - inter
    iou = inter / union
    whg = torch.clamp(torch.max(p[:, 2:], t[:, 2:]) + torch.max(p[:, :2], t[:, :2]), min=0)
    g = whg[:, 0] * whg[:, 1]
    return (1.0 - (iou - (g - union) / g.clamp(1e-10))).sum()


def cls_loss(cls_levels, cls_t: Tensor, mask: Tensor) -> Tensor:
    """compute_cls_loss -> [B]: focal over ALL points / clamp(num_pos, 1).   loss.py:6-26"""
    c = cls_levels[0].shape[1]
    logits = _flat(cls_levels, c)
    assert logits.shape[:2] == cls_t.shape[:2]
    npos = mask.sum(dim=1).clamp(min=1).float()
    ids = torch.arange(1, c + 1, device=cls_t.device)[None, :]
    out = [focal_sum(logits[b], (ids == cls_t[b]).float()).view(1) for b in range(logits.shape[0])]
    return torch.cat(out) / npos


def cnt_loss(cnt_levels, cnt_t: Tensor, mask: Tensor) -> Tensor:
    """compute_cnt_loss -> [B]: BCE-with-logits over positives / num_pos.   loss.py:29-57"""
    logits = _flat(cnt_levels, cnt_t.shape[-1])
    assert logits.shape == cnt_t.shape
    npos = mask.sum(dim=1).clamp(min=1).float()
    bce = torch.nn.functional.binary_cross_entropy_with_logits
    out = [bce(logits[b][mask[b]].reshape(-1), cnt_t[b][mask[b]].reshape(-1), reduction="sum").view(1)
           for b in range(logits.shape[0])]
    return torch.cat(out) / npos


def reg_loss(reg_levels, reg_t: Tensor, mask: Tensor, mode: str = "iou") -> Tensor:
    """compute_reg_loss -> [B]: IoU / GIoU over positives / num_pos.      loss.py:116-139"""
    if mode not in ("iou", "giou"):
        raise NotImplementedError("reg loss only implemented ['iou','giou']")
    pred = _flat(reg_levels, reg_t.shape[-1])
    assert pred.shape == reg_t.shape
    npos = mask.sum(dim=1).clamp(min=1).float()
    fn = iou_sum if mode == "iou" else giou_sum
    out = [fn(pred[b][mask[b]], reg_t[b][mask[b]]).view(1) for b in range(pred.shape[0])]
    return torch.cat(out) / npos


def fcos_loss(pred, target, mode: str = "giou"):
    """FCOSLoss.forward -> (cls, cnt, reg, total) 0-dim tensors.          loss.py:201-215"""
    cls_l, cnt_l, reg_l = pred
    cls_t, cnt_t, reg_t = target
    mask = (cnt_t > -1).squeeze(-1)
    a = cls_loss(cls_l, cls_t, mask).mean()
    b = cnt_loss(cnt_l, cnt_t, mask).mean()
    c = reg_loss(reg_l, reg_t, mask, mode).mean()
    return a, b, c, a + b + c
