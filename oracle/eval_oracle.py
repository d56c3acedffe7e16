"""CPU restatement of the reference's VOC average-precision evaluation — TEST INFRASTRUCTURE ONLY.

Only ``tests/`` may import this module; the product path (``pytorch_object_detection_b200/eval.py``) runs on
the GPU and never falls back to it.  Pinned against outputs of the reference's own ``eval_ap_2d`` on seeded
inputs (``tests/golden/eval_ap.npz``, made by ``tests/golden/make_golden_eval.py``).

Follows ``test.py`` of the reference: ``iou_2d`` (test.py:24-55), ``_compute_ap`` (test.py:58-83),
``eval_ap_2d`` (test.py:86-162).
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np


def iou_gt_vs_det(gts: np.ndarray, det: np.ndarray) -> np.ndarray:
    """IoU of every GT box [n,4] with one detection [4], fp32, operands in the reference's order (test.py:35-55)."""
    f = np.float32
    w = np.maximum(f(0), np.minimum(gts[:, 2], det[2]) - np.maximum(gts[:, 0], det[0]))
    h = np.maximum(f(0), np.minimum(gts[:, 3], det[3]) - np.maximum(gts[:, 1], det[1]))
    overlap = w * h
    area_g = (gts[:, 2] - gts[:, 0]) * (gts[:, 3] - gts[:, 1])
    area_d = (det[2] - det[0]) * (det[3] - det[1])
    with np.errstate(invalid="ignore", divide="ignore"):
        return overlap / (area_g + area_d - overlap)


def average_precision(recall: np.ndarray, precision: np.ndarray) -> float:
    """Area under the precision envelope where recall changes (test.py:58-83)."""
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([0.0], precision, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]            # envelope: running maximum from the end
    idx = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[idx + 1] - mrec[idx]) * mpre[idx + 1]))


def eval_ap(gt_boxes: Sequence[np.ndarray], gt_labels: Sequence[np.ndarray], det_boxes: Sequence[np.ndarray],
            det_labels: Sequence[np.ndarray], det_scores: Sequence[np.ndarray], iou_thr: float,
            num_cls: int) -> Dict[int, float]:
    """{label: AP} for label in 1..num_cls-1; detections are taken in the order given per image (test.py:86-162)."""
    out = {}
    for label in range(1, num_cls):
        tps, scores, total_gts = [], [], 0
        for g_box, g_lab, d_box, d_lab, d_sc in zip(gt_boxes, gt_labels, det_boxes, det_labels, det_scores):
            gts = np.asarray(g_box, dtype=np.float32)[np.asarray(g_lab) == label]
            sel = np.asarray(d_lab) == label
            dets = np.asarray(d_box, dtype=np.float32)[sel]
            total_gts += len(gts)
            taken = set()
            for det, sc in zip(dets, np.asarray(d_sc)[sel]):
                scores.append(sc)
                hit = False
                if len(gts):
                    iou = iou_gt_vs_det(gts, det)
                    j = int(np.argmax(iou))                       # first index of the maximum (NaN counts as maximum)
                    if iou[j] >= iou_thr and j not in taken:
                        taken.add(j)
                        hit = True
                tps.append(1.0 if hit else 0.0)
        tp = np.asarray(tps, dtype=np.float64)
        order = np.argsort(-np.asarray(scores, dtype=np.float64), kind="stable")
        tp = np.cumsum(tp[order])
        fp = np.cumsum(1.0 - np.asarray(tps, dtype=np.float64)[order]) if len(tps) else tp
        with np.errstate(invalid="ignore", divide="ignore"):
            recall = tp / total_gts
            precision = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
        out[label] = average_precision(recall, precision)
    return out
