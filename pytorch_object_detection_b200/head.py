"""Drop-in mirrors of the reference's ``model/modules/head.py`` classes, backed by the CUDA library.

Same constructor / ``forward`` signatures, argument meaning and error behaviour as
``FCOSHead`` (head.py:41-102), ``ClipBoxes`` (head.py:152-162) and ``FCOSGenTargets``
(head.py:211-316), so ``test.py:191-207``, ``Test_coco.py:135-142`` and ``train.py:98-99,177``
can import them unchanged.  Nothing here computes on the CPU.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops

Tensor = torch.Tensor


def _ragged_error(counts: Sequence[int]) -> RuntimeError:
    # the reference's final torch.stack raises exactly this for ragged batches (head.py:99-101)
    return RuntimeError(
        f"stack expects each tensor to be equal size, but images kept {list(counts)} boxes; "
        "the reference's FCOSHead only supports batch 1 — use FCOSHead.detect() for the padded "
        "[B, max_box] + count[B] outputs")


class FCOSHead(nn.Module):
    """Inference post-processing: score, global top-k, threshold, batched NMS (head.py:41-102)."""

    def __init__(self, score_threshold: float, nms_threshold: float, max_detection_box: int, strides: List[int],
                 reg_exp_scales: Sequence[Tensor] | None = None):
        """``reg_exp_scales`` (extension, default off): the head's per-level ``ScaleExp.scale`` parameters
        (HISFcos.py:209,228).  When given, ``x[2]`` must hold the RAW ``reg_pred`` convolution outputs and the
        ``exp(x * scale)`` pass of the model is evaluated inside the decode for the selected points only."""
        super().__init__()
        self.score = score_threshold
        self.nms_threshold = nms_threshold
        self.max_box = max_detection_box
        self.strides = strides
        self.reg_exp_scales = reg_exp_scales

    # -- batched, padded contract (new: the reference cannot return ragged batches) ------------
    def detect(self, x, clip_hw: Tuple[int, int] | None = None, out_packed: Tensor | None = None):
        """x = (cls_list, cnt_list, reg_list).  Returns scores [B,K], classes [B,K] i64,
        boxes [B,K,4], counts [B] i32 (device tensors, no host sync); boxes are clipped to
        ``clip_hw`` = (H, W) when given (ClipBoxes fused into the writer).  ``out_packed`` places
        the outputs in a caller-owned buffer (see ``ops.postprocess``)."""
        if out_packed is not None or self.reg_exp_scales is not None:
            s, c, b, _, n = ops.postprocess(x[0], x[1], x[2], self.strides, self.score, self.nms_threshold,
                                            self.max_box, clip_hw, out_packed, self.reg_exp_scales)
            return s, c, b, n
        s, c, b, _, n = torch.ops.b200det.postprocess(
            list(x[0]), list(x[1]), list(x[2]), [int(v) for v in self.strides], float(self.score),
            float(self.nms_threshold), int(self.max_box), *((int(clip_hw[0]), int(clip_hw[1])) if clip_hw else (0, 0)))
        return s, c, b, n

    @staticmethod
    def _exact(scores: Tensor, classes: Tensor, boxes: Tensor, counts: Tensor):
        n = counts.tolist()           # the one host sync; the reference syncs here too (boolean indexing)
        if any(v != n[0] for v in n):
            raise _ragged_error(n)
        k = n[0]
        return scores[:, :k], classes[:, :k], boxes[:, :k]

    def forward(self, x):
        scores, classes, boxes, counts = self.detect(x)
        return self._exact(scores, classes, boxes, counts)

    def post_process(self, preds_top_k: List[Tensor]):
        """threshold -> batched NMS -> gather on already selected candidates (head.py:84-102)."""
        cls_score_top_k, cls_class_top_k, box_top_k = preds_top_k
        s, c, b, _, n = torch.ops.b200det.batched_nms(box_top_k, cls_score_top_k, cls_class_top_k,
                                                       float(self.score), float(self.nms_threshold), 0, 0)
        return self._exact(s, c, b, n)


class ClipBoxes(nn.Module):
    """In-place clamp of boxes to the image (head.py:152-162).  Returns the same tensor."""

    def __init__(self):
        super().__init__()

    @staticmethod
    def forward(batch_imgs: Tensor, batch_boxes: Tensor) -> Tensor:
        h, w = batch_imgs.shape[2:]
        return torch.ops.b200det.clip_boxes_(batch_boxes, int(h), int(w))


class FCOSGenTargets(nn.Module):
    """Training target assignment (head.py:211-316)."""

    def __init__(self, strides: List[int], limit_range: List[List[int]]):
        super().__init__()
        self.stride = strides
        self.lim_range = limit_range
        assert len(strides) == len(limit_range)

    def forward(self, x):
        cls_logit, center_logit, reg_logit = x[0]
        gt_box = x[1]
        labels = x[2]
        assert len(self.stride) == len(cls_logit)
        level_hw = [v for t in cls_logit for v in (int(t.shape[2]), int(t.shape[3]))]
        return torch.ops.b200det.assign_targets(
            level_hw, [int(s) for s in self.stride], [float(r[0]) for r in self.lim_range],
            [float(r[1]) for r in self.lim_range], gt_box, labels, 1.5)

    @staticmethod
    def generate_target(lv_out, gt_box: Tensor, labels: Tensor, stride: int, lim_range: Sequence[float],
                        sample_radio_ratio: float = 1.5):
        """One level (head.py:235-316); only the level's spatial size is read from ``lv_out``."""
        cls_logit = lv_out[0]
        return torch.ops.b200det.assign_targets(
            [int(cls_logit.shape[2]), int(cls_logit.shape[3])], [int(stride)], [float(lim_range[0])],
            [float(lim_range[1])], gt_box, labels, float(sample_radio_ratio))
