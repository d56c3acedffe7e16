"""Drop-in mirrors of the reference's ``model/loss.py``, backed by the CUDA library.

``FCOSLoss`` (loss.py:196-215) and the free functions ``compute_cls_loss`` (loss.py:6-26),
``compute_cnt_loss`` (loss.py:29-57), ``compute_reg_loss`` (loss.py:116-139), ``iou_loss``
(loss.py:142-152), ``giou_loss`` (loss.py:155-177) and ``focal_loss_from_logits``
(loss.py:180-193) keep their signatures, return shapes and error behaviour.  Each is one fused
forward kernel and one fused backward kernel that read / write the per-level NCHW maps in place;
autograd sees them through ``torch.autograd.Function``.  Nothing here computes on the CPU.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

from . import ops

Tensor = torch.Tensor
_MODES = {"iou": 0, "giou": 1}


def _mask_source(mask: Tensor) -> Tensor:
    """bool [B,P] -> float [B,P] with positives > -1 (what the kernels test, loss.py:205)."""
    return torch.where(mask, 0.0, -1.0).to(torch.float32)


class _BoxLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mask_src: Tensor, reg_t: Tensor, mode: int, *reg: Tensor):
        loss, npos = ops.box_loss_fwd(reg, mask_src, reg_t, mode)
        ctx.save_for_backward(mask_src, reg_t, npos, *reg)
        ctx.mode = mode
        return loss

    @staticmethod
    def backward(ctx, grad_loss: Tensor):
        mask_src, reg_t, npos, *reg = ctx.saved_tensors
        grads = ops.box_loss_bwd(reg, mask_src, reg_t, ctx.mode, grad_loss, npos)
        return (None, None, None, *[g.to(r.dtype) for g, r in zip(grads, reg)])


class _CntLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mask_src: Tensor, cnt_t: Tensor, *cnt: Tensor):
        loss, npos = ops.cnt_loss_fwd(cnt, mask_src, cnt_t)
        ctx.save_for_backward(mask_src, cnt_t, npos, *cnt)
        return loss

    @staticmethod
    def backward(ctx, grad_loss: Tensor):
        mask_src, cnt_t, npos, *cnt = ctx.saved_tensors
        grads = ops.cnt_loss_bwd(cnt, mask_src, cnt_t, grad_loss, npos)
        return (None, None, *[g.to(c.dtype) for g, c in zip(grads, cnt)])


class _ClsLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mask_src: Tensor, cls_t: Tensor, *cls: Tensor):
        loss, npos = ops.cls_loss_fwd(cls, mask_src, cls_t)
        ctx.save_for_backward(cls_t, npos, *cls)
        return loss

    @staticmethod
    def backward(ctx, grad_loss: Tensor):
        cls_t, npos, *cls = ctx.saved_tensors
        grads = ops.cls_loss_bwd(cls, cls_t, grad_loss, npos)
        return (None, None, *[g.to(c.dtype) for g, c in zip(grads, cls)])


class _Upstream:
    """The upstream gradient a loss output is ASSUMED to receive, kept on the device.

    The step kernels write final gradients during the forward pass, so they must know dL/d(loss) in advance.
    It starts at 1 (``total_loss.backward()``); every backward overwrites it with the value that actually
    arrived (b200det_rescale_maps), and rescales the stored gradients only when the assumption THAT forward read
    (each forward keeps its own copy) was wrong.
    Under ``torch.cuda.amp.GradScaler`` (train.py:127,180) the upstream gradient is the loss scale, which is
    constant for thousands of steps: from the second step on the rescale pass never runs.  No host sync."""

    def __init__(self):
        self._value = {}
        self.observed = False        # a backward has stored the upstream gradient that really arrives

    def on(self, device: torch.device) -> Tensor:
        key = (device.type, device.index)
        v = self._value.get(key)
        if v is None:
            v = torch.zeros(2, dtype=torch.float32, device=device)               # {assumed, ticket}
            v[:1].fill_(1.0)                                                       # (fill kernels only: capturable)
            if not torch.cuda.is_current_stream_capturing():
                self._value[key] = v
        return v


def _as_scalar(g: Tensor) -> Tensor:
    return g.detach().to(torch.float32).reshape(1).contiguous()


class _ClsLossStep(torch.autograd.Function):
    """Batch-mean focal loss whose gradient is written by the forward kernel (one read of the logits), for the
    upstream gradient assumed in ``up`` (see _Upstream); backward rescales only on a wrong assumption.  fp16 /
    bf16 logits are read as they are and the gradient is written in their type.  One backward per forward."""

    @staticmethod
    def forward(ctx, cls_t: Tensor, mask_src, num_pos, up: _Upstream, *cls: Tensor):
        state = up.on(cls[0].device)
        loss, mean, npos, grads = ops.cls_loss_step(cls, cls_t, mask_src=mask_src, num_pos=num_pos, up_mean=state)
        # mean[1] = the upstream gradient THIS forward read: the shared state may have moved on by the time its
        # backward runs (two forwards outstanding, a GradScaler scale change, per-call loss weights)
        ctx.save_for_backward(state, mean[1:2], *grads)
        ctx.up = up
        ctx.dtypes = [t.dtype for t in cls]
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(loss, npos)
        return mean[0], loss, npos

    @staticmethod
    def backward(ctx, g_mean, *_):
        if getattr(ctx, "consumed", False):
            raise RuntimeError("the fused focal step supports a single backward per forward")
        ctx.consumed = True
        if g_mean is None:
            return (None, None, None, None, *[None] * len(ctx.dtypes))
        state, assumed, *grads = ctx.saved_tensors
        ops.rescale_maps_(grads, [_as_scalar(g_mean)] * len(grads), [state] * len(grads), [assumed] * len(grads))
        ctx.up.observed = True
        return (None, None, None, None, *[g.to(dt) for g, dt in zip(grads, ctx.dtypes)])


class _TapUpstream(torch.autograd.Function):
    """Identity whose backward stores the gradient that arrives into an _Upstream (first step of half-precision
    logits: fp16 gradients written for a wrongly assumed loss scale would have underflowed)."""

    @staticmethod
    def forward(ctx, x: Tensor, up: _Upstream):
        ctx.up = up
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if g is not None:
            word = ctx.up.on(g.device)[:1]
            arrived = _as_scalar(g)
            word.copy_(torch.where((arrived != 0) & torch.isfinite(arrived), arrived, word))
            ctx.up.observed = True
        return g, None


def _cls_loss_mean(preds: List[Tensor], cls_t: Tensor, mask_src: Tensor | None, num_pos: Tensor | None,
                   up: _Upstream, per_image: dict | None = None) -> Tensor:
    """``compute_cls_loss(...).mean()`` (loss.py:210).  When a gradient will be asked for, the forward kernel
    writes it too (238 MB read once instead of twice at COCO batch 32); otherwise the forward kernel alone runs.
    ``per_image["cls"]`` receives the detached per-image losses [B] (what a multi-rank run reduces)."""
    _check_points(preds, cls_t)                                    # loss.py:18
    if torch.is_grad_enabled() and any(t.requires_grad for t in preds):
        half = preds[0].dtype in (torch.float16, torch.bfloat16)
        if not half or up.observed:
            mean, loss, _ = _ClsLossStep.apply(cls_t, mask_src, num_pos, up, *preds)
            if per_image is not None:
                per_image["cls"] = loss
            return mean
        # half-precision logits, first step: the loss scale is not known yet -> forward and backward kernels
        if mask_src is not None:
            loss = _ClsLoss.apply(mask_src, cls_t, *preds)
            if per_image is not None:
                per_image["cls"] = loss.detach()
            return _TapUpstream.apply(loss.mean(), up)
    if mask_src is None:
        raise ValueError("the forward-only focal loss needs the positive-mask source (cnt_t)")
    loss = _ClsLoss.apply(mask_src, cls_t, *preds)
    if per_image is not None:
        per_image["cls"] = loss.detach()
    return loss.mean()


def _check_points(preds: Sequence[Tensor], target: Tensor) -> None:
    p = sum(int(t.shape[2]) * int(t.shape[3]) for t in preds)
    assert p == target.shape[1] and preds[0].shape[0] == target.shape[0], \
        f"predictions cover [{preds[0].shape[0]}, {p}] points, target {tuple(target.shape[:2])}"


def compute_cls_loss(preds: List[Tensor], target: Tensor, mask: Tensor, _mask_src: Tensor | None = None) -> Tensor:
    """Focal loss over all points / clamp(num_pos, 1) -> [B] (loss.py:6-26)."""
    _check_points(preds, target)                                   # loss.py:18
    return _ClsLoss.apply(_mask_source(mask) if _mask_src is None else _mask_src, target, *preds)


def compute_cnt_loss(preds: List[Tensor], target: Tensor, mask: Tensor, _mask_src: Tensor | None = None) -> Tensor:
    """BCE-with-logits over positives / clamp(num_pos, 1) -> [B] (loss.py:29-57)."""
    _check_points(preds, target)                                   # loss.py:41
    assert target.shape[-1] == 1 and preds[0].shape[1] == 1
    return _CntLoss.apply(_mask_source(mask) if _mask_src is None else _mask_src, target, *preds)


def compute_reg_loss(preds: List[Tensor], target: Tensor, mask: Tensor, mode: str = "iou",
                     _mask_src: Tensor | None = None) -> Tensor:
    """IoU / GIoU loss over positives / clamp(num_pos, 1) -> [B] (loss.py:116-139)."""
    _check_points(preds, target)                                   # loss.py:127
    assert target.shape[-1] == 4 and preds[0].shape[1] == 4
    if mode not in _MODES:
        raise NotImplementedError("reg loss only implemented ['iou','giou']")   # loss.py:137-138
    return _BoxLoss.apply(_mask_source(mask) if _mask_src is None else _mask_src, target, _MODES[mode], *preds)


def _pairwise(preds: Tensor, targets: Tensor, mode: int) -> Tensor:
    # [n,4] pairs as a single one-image, one-level problem: preds -> [1,4,n,1] map, every row positive
    n = preds.shape[0]
    if n == 0:
        return preds.sum() * 0.0
    level = preds.t().reshape(1, 4, n, 1)
    mask_src = torch.zeros((1, n), dtype=torch.float32, device=preds.device)
    loss = _BoxLoss.apply(mask_src, targets.reshape(1, n, 4), mode, level)
    return (loss * float(n)).reshape(())          # kernel divides by num_pos = n; the reference returns the sum


def iou_loss(preds: Tensor, targets: Tensor) -> Tensor:
    """-log(clamp(iou, 1e-6)) summed over [n,4] ltrb pairs (loss.py:142-152)."""
    return _pairwise(preds, targets, 0)


def giou_loss(preds: Tensor, targets: Tensor) -> Tensor:
    """(1 - giou) summed over [n,4] ltrb pairs (loss.py:155-177)."""
    return _pairwise(preds, targets, 1)


def focal_loss_from_logits(preds: Tensor, targets: Tensor, gamma: float = 2.0, alpha: float = 0.25) -> Tensor:
    """Summed focal loss of logits [P,C] against one-hot targets [P,C] (loss.py:180-193)."""
    if gamma != 2.0 or alpha != 0.25:
        raise NotImplementedError("the CUDA focal loss is specialised for gamma=2.0, alpha=0.25 (the reference's call)")
    p, c = preds.shape
    hot = targets > 0.5
    # (checked on the device, asynchronously: no host synchronisation inside the loss)
    torch._assert_async((hot.sum(dim=1) <= 1).all(), "focal_loss_from_logits: targets must be one-hot rows (or all zero)")
    cls_t = torch.where(hot.any(dim=1), hot.float().argmax(dim=1) + 1, 0).reshape(1, p, 1)
    level = preds.t().reshape(1, c, p, 1)
    mask_src = torch.full((1, p), -1.0, dtype=torch.float32, device=preds.device)   # num_pos clamps to 1
    return _ClsLoss.apply(mask_src, cls_t, level).reshape(())


class FCOSLoss(nn.Module):
    """(cls_loss, cnt_loss, reg_loss, total_loss) as 0-dim tensors (loss.py:196-215)."""

    def __init__(self, mode: str = "giou"):
        super().__init__()
        self.mode = mode
        self._up_cls = _Upstream()

    def forward(self, x):
        pred, target = x
        cls_logit, cnt_logit, reg_logit = pred
        cls_target, cnt_target, reg_target = target
        mask_pos = None                    # loss.py:205: cnt_target > -1; the kernels test cnt_target directly
        src = cnt_target
        cls_loss = _cls_loss_mean(cls_logit, cls_target, src, None, self._up_cls)
        cnt_loss = compute_cnt_loss(cnt_logit, cnt_target, mask_pos, _mask_src=src).mean()
        reg_loss = compute_reg_loss(reg_logit, reg_target, mask_pos, self.mode, _mask_src=src).mean()
        total_loss = cls_loss + cnt_loss + reg_loss
        return cls_loss, cnt_loss, reg_loss, total_loss


class _FusedTargetLoss(torch.autograd.Function):
    """Targets + box / centerness losses + their gradients from one kernel (csrc/train_fused.cu).

    forward computes the gradients of the two batch means eagerly for the upstream gradients assumed in
    ``up_box`` / ``up_cnt`` (see _Upstream); backward only rescales them when an assumption was wrong (a no-op
    launch otherwise).  backward may run once per forward."""

    @staticmethod
    def forward(ctx, gt_boxes: Tensor, labels: Tensor, cfg, up_box: Tensor, up_cnt: Tensor, *maps: Tensor):
        strides, limit_range, mode, radius, n_reg, n_cnt = cfg
        reg, cnt, scales = maps[:n_reg], (maps[n_reg:n_reg + n_cnt] or None), (maps[n_reg + n_cnt:] or None)
        r = ops.assign_loss_fused(reg, cnt, strides, limit_range, gt_boxes, labels, mode, radius,
                                  reg_exp_scales=scales, up_box=up_box, up_cnt=up_cnt if cnt else None)
        ctx.n_reg, ctx.has_cnt = n_reg, cnt is not None
        ctx.dtypes = [t.dtype for t in maps]
        ctx.scale_shapes = [t.shape for t in scales] if scales else []
        mean = r["mean"]                          # [2], [3]: the upstream gradients THIS forward read (see _ClsLossStep)
        ctx.save_for_backward(up_box, up_cnt, mean[2:3], mean[3:4], *r["reg_grads"], *(r["cnt_grads"] or []),
                              *([r["scale_grad"]] if scales else []))
        ctx.set_materialize_grads(False)
        aux = [r["cls_t"], r["cnt_t"], r["reg_t"], r["box_loss"], r["num_pos"]] + ([r["cnt_loss"]] if cnt else [])
        ctx.mark_non_differentiable(*aux)
        return (mean[0], mean[1], *aux)

    @staticmethod
    def backward(ctx, g_box, g_cnt, *_):
        if getattr(ctx, "consumed", False):
            raise RuntimeError("the fused target/loss step supports a single backward per forward")
        ctx.consumed = True
        up_box, up_cnt, read_box, read_cnt, *grads = ctx.saved_tensors
        n = ctx.n_reg
        scale_grad = grads.pop() if ctx.scale_shapes else None
        todo, got, state, assumed = [], [], [], []
        g_box = None if g_box is None else _as_scalar(g_box)
        g_cnt = None if g_cnt is None else _as_scalar(g_cnt)
        for i, g in enumerate(grads):
            arrived, up, read = (g_box, up_box, read_box) if i < n else (g_cnt, up_cnt, read_cnt)
            if arrived is None:
                grads[i] = None
            else:
                todo.append(g); got.append(arrived); state.append(up); assumed.append(read)
        if scale_grad is not None and g_box is not None:
            todo.append(scale_grad); got.append(g_box); state.append(up_box); assumed.append(read_box)
        if todo:
            ops.rescale_maps_(todo, got, state, assumed)
        out = [g if g is None else g.to(dt) for g, dt in zip(grads, ctx.dtypes)]
        if scale_grad is not None:                  # one gradient per ScaleExp.scale parameter
            out += [None if g_box is None else scale_grad[i].reshape(shp) for i, shp in enumerate(ctx.scale_shapes)]
        return (None, None, None, None, None, *out)


class FCOSTargetLoss(nn.Module):
    """``FCOSGenTargets`` (head.py:211-316) followed by ``FCOSLoss`` (loss.py:196-215) as ONE module.

    ``forward([out, gt_boxes, labels])`` takes what the reference hands to ``FCOSGenTargets`` and returns
    what its ``FCOSLoss`` returns, ``(cls_loss, cnt_loss, reg_loss, total_loss)``; the targets of the step
    are kept in ``self.targets`` as ``(cls_t [B,P,1] i64, cnt_t [B,P,1], reg_t [B,P,4])``.  Target
    assignment, the box and centerness losses and their gradients are one kernel launch; the focal loss
    and its gradient are one more (``b200det_cls_loss_step``: the logits are read once).
    """

    def __init__(self, strides: Sequence[int], limit_range: Sequence[Sequence[float]], mode: str = "giou",
                 sample_radio_ratio: float = 1.5, reg_exp_scales: Sequence[Tensor] | None = None):
        """``reg_exp_scales`` (extension, default off): the head's per-level ``ScaleExp.scale`` parameters
        (HISFcos.py:209,228).  When given, ``reg_preds`` must be the RAW ``reg_pred`` convolution outputs:
        ``exp(x * scale)`` is evaluated at the positives inside the kernel, gradients flow to x and to the scales."""
        super().__init__()
        assert len(strides) == len(limit_range)                          # head.py:216
        if mode not in _MODES:
            raise NotImplementedError("reg loss only implemented ['iou','giou']")   # loss.py:137-138
        self.strides = [int(s) for s in strides]
        self.limit_range = [tuple(float(v) for v in r) for r in limit_range]
        self.mode = mode
        self.sample_radio_ratio = float(sample_radio_ratio)
        self.reg_exp_scales = reg_exp_scales
        self.targets = None
        self._up_cls, self._up_box, self._up_cnt = _Upstream(), _Upstream(), _Upstream()

    def box_cnt_losses(self, cnt_logits, reg_preds, gt_boxes: Tensor, labels: Tensor):
        """(reg_loss, cnt_loss) batch means + targets; ``cnt_logits`` may be None (box loss only)."""
        n = min(len(self.strides), len(reg_preds), len(cnt_logits) if cnt_logits is not None else len(reg_preds))
        n_cnt = n if cnt_logits is not None else 0
        cfg = (self.strides[:n], self.limit_range[:n], _MODES[self.mode], self.sample_radio_ratio, n, n_cnt)
        maps = list(reg_preds[:n]) + (list(cnt_logits[:n]) if cnt_logits is not None else [])
        if self.reg_exp_scales is not None:
            maps += list(self.reg_exp_scales)[:n]
        dev = maps[0].device
        out = _FusedTargetLoss.apply(gt_boxes, labels, cfg, self._up_box.on(dev), self._up_cnt.on(dev), *maps)
        self.targets = (out[2], out[3], out[4])
        self.per_image = {"reg": out[5], "num_pos": out[6], "cnt": out[7] if cnt_logits is not None else None}
        return out[0], (out[1] if cnt_logits is not None else None)

    def forward(self, x):
        (cls_logits, cnt_logits, reg_preds), gt_boxes, labels = x
        assert len(cls_logits) >= 1 and len(cnt_logits) == len(cls_logits) == len(reg_preds)
        reg_loss, cnt_loss = self.box_cnt_losses(cnt_logits, reg_preds, gt_boxes, labels)
        cls_t, cnt_t, _ = self.targets
        n = min(len(self.strides), len(cls_logits))
        cls_loss = _cls_loss_mean(list(cls_logits[:n]), cls_t, cnt_t, self.per_image["num_pos"], self._up_cls,
                                  self.per_image)
        total_loss = cls_loss + cnt_loss + reg_loss
        return cls_loss, cnt_loss, reg_loss, total_loss
