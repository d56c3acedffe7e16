"""Device-side mirror of the reference's VOC evaluation (``test.py:15-162``): ``sort_by_score`` and ``eval_ap_2d``.

``eval_ap_batched`` is the native entry: it takes the padded detections exactly as ``FCOSHead.detect()`` leaves
them in HBM (scores descending per image) plus the padded GT batch, so an evaluation epoch concatenates device
tensors and needs ONE device->host copy of ``num_cls`` doubles at its end.  ``eval_ap_2d`` keeps the reference's
list-of-arrays signature and return value (``{label: ap}``).  Nothing here computes on the CPU.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _lib
from .data import pack_gt
from .ops import _count, _need_cuda, _stream

Tensor = torch.Tensor


def eval_ap_batched(scores: Tensor, classes: Tensor, boxes: Tensor, counts: Tensor, gt_boxes: Tensor,
                    gt_labels: Tensor, iou_thread: float, num_cls: int) -> Tensor:
    """AP per class as a CUDA f64 tensor [num_cls] (index 0 = background = 0).

    scores [N,K] f32, classes [N,K] i64, boxes [N,K,4] f32, counts [N] i32: detections of N images in the order
    they are to be matched (``detect()`` emits them by descending score); gt_boxes [N,M,4], gt_labels [N,M] (-1 pad).
    """
    lib = _lib.load()
    for t, what in ((scores, "scores"), (classes, "classes"), (boxes, "boxes"), (counts, "counts"),
                    (gt_boxes, "gt_boxes"), (gt_labels, "gt_labels")):
        _need_cuda(t, what)
    dev = scores.device
    n, k = scores.shape
    m = gt_labels.shape[1]
    if boxes.shape != (n, k, 4) or classes.shape != (n, k) or counts.shape != (n,) or gt_boxes.shape != (n, m, 4):
        raise _lib.B200DetError("expected scores/classes [N,K], boxes [N,K,4], counts [N], gt_boxes [N,M,4], gt_labels [N,M]")
    ap = torch.zeros((int(num_cls),), dtype=torch.float64, device=dev)
    if n == 0 or k == 0:
        return ap
    scores = scores.to(torch.float32).contiguous()
    classes = classes.to(torch.int64).contiguous()
    boxes = boxes.to(torch.float32).contiguous()
    counts = counts.to(torch.int32).contiguous()
    gt_boxes = gt_boxes.to(torch.float32).contiguous()
    gt_labels = gt_labels.to(torch.int64).contiguous()
    ws_bytes = lib.b200det_eval_ap_workspace_bytes(n, k, int(num_cls))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.b200det_eval_ap(n, k, m, int(num_cls), scores.data_ptr(), classes.data_ptr(), boxes.data_ptr(),
                                 counts.data_ptr(), gt_boxes.data_ptr() if m else None,
                                 gt_labels.data_ptr() if m else None, float(iou_thread), ws.data_ptr(), ws_bytes,
                                 ap.data_ptr(), _stream(ws))
    _lib.check(rc, "b200det_eval_ap")
    _count("eval_ap")
    return ap


def _as_cuda(x, dtype, device) -> Tensor:
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    return t.to(device=device, dtype=dtype, non_blocking=True)


def _pad_rows(rows: Sequence[Tensor], width: int, fill: float, dtype, device) -> Tensor:
    k = max([int(r.shape[0]) for r in rows] + [1])
    shape = (len(rows), k, width) if width else (len(rows), k)
    out = torch.full(shape, fill, dtype=dtype, device=device)
    for i, r in enumerate(rows):
        if r.shape[0]:
            out[i, :r.shape[0]] = r
    return out


def sort_by_score(pred_boxes, pred_labels, pred_scores, device="cuda"):
    """Per image, order detections by descending score (test.py:15-20).  Returns three lists of CUDA tensors;
    equal scores keep their input order (numpy's argsort leaves them unspecified)."""
    b, l, s = [], [], []
    for pb, pl, ps in zip(pred_boxes, pred_labels, pred_scores):
        ps = _as_cuda(ps, torch.float32, device)
        order = torch.argsort(ps, descending=True, stable=True)
        b.append(_as_cuda(pb, torch.float32, device).reshape(-1, 4)[order])
        l.append(_as_cuda(pl, torch.int64, device)[order])
        s.append(ps[order])
    return b, l, s


def eval_ap_2d(gt_boxes, gt_labels, pred_boxes, pred_labels, pred_scores, iou_thread: float, num_cls: int,
               device="cuda") -> Dict[int, float]:
    """``{label: average precision}`` for label 1..num_cls-1 — the reference's signature and result (test.py:86-162).
    Lists of per-image arrays (numpy or torch); detections are matched in the order given, as in the reference."""
    n = len(gt_boxes)
    assert n == len(gt_labels) == len(pred_boxes) == len(pred_labels) == len(pred_scores)
    if n == 0:
        return {c: 0.0 for c in range(1, num_cls)}
    dev = torch.device(device)
    gb, gl = pack_gt([_as_cuda(x, torch.float32, dev).reshape(-1, 4) for x in gt_boxes],
                     [_as_cuda(x, torch.int64, dev).reshape(-1) for x in gt_labels], dev)
    pb = [_as_cuda(x, torch.float32, dev).reshape(-1, 4) for x in pred_boxes]
    boxes = _pad_rows(pb, 4, 0.0, torch.float32, dev)
    classes = _pad_rows([_as_cuda(x, torch.int64, dev).reshape(-1) for x in pred_labels], 0, -1, torch.int64, dev)
    scores = _pad_rows([_as_cuda(x, torch.float32, dev).reshape(-1) for x in pred_scores], 0, 0.0, torch.float32, dev)
    counts = torch.tensor([int(x.shape[0]) for x in pb], dtype=torch.int32).to(dev, non_blocking=True)
    ap = eval_ap_batched(scores, classes, boxes, counts, gb, gl, iou_thread, num_cls).cpu().numpy()   # the one D2H
    return {c: float(ap[c]) for c in range(1, num_cls)}


def coco_results(scores: Tensor, classes: Tensor, boxes: Tensor, counts: Tensor, scales: Tensor, image_ids: Sequence,
                 id2category, threshold: float = 0.05) -> List[dict]:
    """The ``results`` list ``evaluate_coco`` hands to pycocotools (Test_coco.py:144-168) for a padded batch of
    detections: boxes / scale -> (x, y, w, h), detections up to the first score below ``threshold``.  The box
    arithmetic and the cut run on the device; one D2H copy per batch instead of three per image."""
    lib = _lib.load()
    for t, what in ((scores, "scores"), (classes, "classes"), (boxes, "boxes"), (counts, "counts"), (scales, "scales")):
        _need_cuda(t, what)
    n, k = scores.shape
    boxes = boxes.to(torch.float32).contiguous()
    scores = scores.to(torch.float32).contiguous()
    counts = counts.to(torch.int32).contiguous()
    scales = scales.to(torch.float32).reshape(n).contiguous()
    xywh = torch.empty_like(boxes)
    keep = torch.empty((n,), dtype=torch.int32, device=boxes.device)
    with torch.cuda.device(boxes.device):
        rc = lib.b200det_coco_boxes(n, k, boxes.data_ptr(), scores.data_ptr(), counts.data_ptr(), scales.data_ptr(),
                                    float(threshold), xywh.data_ptr(), keep.data_ptr(), _stream(boxes))
    _lib.check(rc, "b200det_coco_boxes")
    _count("coco_boxes")
    h_keep, h_box, h_sc, h_cl = keep.cpu().tolist(), xywh.cpu().numpy(), scores.cpu().numpy(), classes.cpu().numpy()
    results = []
    for i in range(n):
        for j in range(h_keep[i]):
            results.append({"image_id": image_ids[i], "category_id": id2category[int(h_cl[i, j])],
                            "score": float(h_sc[i, j]), "bbox": h_box[i, j].tolist()})
    return results
