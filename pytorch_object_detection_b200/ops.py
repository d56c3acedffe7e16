"""torch.library ops of the ``b200det`` namespace — thin wrappers over the C ABI.

Every op takes CUDA tensors, allocates its outputs/workspace with torch (the library never
allocates), launches on the current CUDA stream and returns without synchronising.  There is
no CPU implementation and no other backend: a non-CUDA tensor raises.

Shapes use the reference's conventions (``model/modules/head.py``, ``model/loss.py``):
level lists are NCHW maps (fp32; the post-process also reads fp16 / bf16 as they are), points are numbered level-major / row-major (the order
``reshape_cat_out`` produces), ``zip(levels, strides)`` truncation included (head.py:20).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import torch

from . import _lib

Tensor = torch.Tensor
launch_count = 0          # kernels launched through this module (bench.py reports it)

# kernels per C entry point (see csrc/*.cu)
_LAUNCHES = {"postprocess": 4, "batched_nms": 3, "score_points": 1, "select_topk": 1, "clip_boxes": 1,
             "assign_targets": 1, "box_loss_fwd": 1, "box_loss_bwd": 1, "cnt_loss_fwd": 1, "cnt_loss_bwd": 1,
             "cls_loss_fwd": 2, "cls_loss_bwd": 1, "cls_loss_step": 2, "count_pos": 1, "assign_loss_fused": 2, "scale_maps": 1, "rescale_maps": 1,
             "pack_gt": 1, "collate_images": 1, "eval_ap": 2, "coco_boxes": 1}


def _count(name: str) -> None:
    global launch_count
    launch_count += _LAUNCHES[name]


def _stream(t: Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _need_cuda(t: Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.B200DetError(f"{what} must be a CUDA tensor: b200det has no CPU path "
                                f"(got {getattr(t, 'device', type(t))})")


def _f32c(t: Tensor, what: str) -> Tensor:
    _need_cuda(t, what)
    if t.dtype != torch.float32:
        t = t.float()         # AMP fp16/bf16 head outputs are evaluated in fp32
    return t if t.is_contiguous() else t.contiguous()


def _scale_ptrs(reg_exp_scales, n: int, keep: List[Tensor]):
    """Device pointers of the per-level ScaleExp.scale parameters (modules.py:170-176), or Nones."""
    if reg_exp_scales is None:
        return [None] * n
    if len(reg_exp_scales) < n:
        raise _lib.B200DetError("reg_exp_scales is shorter than the level list")
    out = []
    for t in list(reg_exp_scales)[:n]:
        _need_cuda(t, "reg_exp_scale")
        t = t.detach()
        if t.dtype != torch.float32 or t.numel() != 1:
            raise _lib.B200DetError("a reg_exp_scale must be ONE fp32 value (ScaleExp.scale)")
        t = t.contiguous()
        keep.append(t)
        out.append(t.data_ptr())
    return out


_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}       # b200det_dtype
NATIVE_HALF_POSTPROCESS = True      # K1 / K2 read fp16 / bf16 head outputs as they are (b200det_level.dtypes)


def _common_half(*lists, n: int):
    """The fp16 / bf16 dtype shared by the first n maps of every given list, else None (-> fp32 up-cast)."""
    dts = {t.dtype for lst in lists if lst is not None for t in list(lst)[:n]}
    if len(dts) == 1:
        dt = dts.pop()
        if dt in (torch.float16, torch.bfloat16):
            return dt
    return None


def _levels(cls: Sequence[Tensor] | None, cnt: Sequence[Tensor] | None, reg: Sequence[Tensor] | None,
            strides: Sequence[int], reg_exp_scales=None, keep_half_cls: bool = False, native_half: bool = False):
    """zip()-truncated level table.  Returns (ctypes array, kept tensors, P, batch, n_levels).  With
    ``keep_half_cls`` fp16 / bf16 class maps are passed as they are (entry points that take a cls_dtype); with
    ``native_half`` (the post-process entry points) cls + cnt maps that share a half dtype, and reg maps that do, are
    passed as they are and declared in ``b200det_level.dtypes`` — anything else is up-cast to fp32."""
    lists = [l for l in (cls, cnt, reg) if l is not None]
    n = min([len(strides)] + [len(l) for l in lists])
    if n == 0:
        raise _lib.B200DetError("no levels")
    keep: List[Tensor] = []
    entries = []
    batch = None
    p_total = 0
    half_cc = _common_half(cls, cnt, n=n) if native_half and (cls is not None or cnt is not None) else None
    half_reg = _common_half(reg, n=n) if native_half and reg is not None else None
    dtypes = _DTYPE_CODE[half_cc or torch.float32] | (_DTYPE_CODE[half_reg or torch.float32] << 4)
    for i in range(n):
        ptrs = []
        hw = None
        for lst, ch in ((cls, None), (cnt, 1), (reg, 4)):
            if lst is None:
                ptrs.append(0)
                continue
            as_is = (keep_half_cls and ch is None and lst[i].dtype in (torch.float16, torch.bfloat16)) or \
                    (half_cc is not None and ch != 4) or (half_reg is not None and ch == 4)
            if as_is:
                _need_cuda(lst[i], "level map")
                t = lst[i] if lst[i].is_contiguous() else lst[i].contiguous()
            else:
                t = _f32c(lst[i], "level map")
            if t.dim() != 4 or (ch is not None and t.shape[1] != ch):
                raise _lib.B200DetError(f"level {i}: expected [B,{ch or 'C'},h,w], got {tuple(t.shape)}")
            if hw is None:
                hw = (t.shape[2], t.shape[3])
            elif hw != (t.shape[2], t.shape[3]):
                raise _lib.B200DetError(f"level {i}: cls/cnt/reg spatial sizes differ")
            if batch is None:
                batch = t.shape[0]
            elif batch != t.shape[0]:
                raise _lib.B200DetError("batch sizes differ between level maps")
            keep.append(t)
            ptrs.append(t.data_ptr())
        entries.append((ptrs[0], ptrs[1], ptrs[2], hw[0], hw[1], int(strides[i])))
        p_total += hw[0] * hw[1]
    if reg_exp_scales is not None:
        n_maps = len(keep)
        scales = _scale_ptrs(reg_exp_scales, n, keep)
        entries = [e + (sp,) for e, sp in zip(entries, scales)]
        keep_maps, keep_scales = keep[:n_maps], keep[n_maps:]
        keep = keep_maps + keep_scales          # maps first: callers slice keep[:...] by map count
    return _lib.make_levels(entries, dtypes), keep, p_total, batch, n


# --------------------------------------------------------------------------------------------
# inference
# --------------------------------------------------------------------------------------------
def postprocess(cls: Sequence[Tensor], cnt: Sequence[Tensor], reg: Sequence[Tensor], strides: Sequence[int],
                score_thr: float, nms_thr: float, max_box: int, clip_hw: Tuple[int, int] | None = None,
                out_packed: Tensor | None = None, reg_exp_scales: Sequence[Tensor] | None = None):
    """FCOSHead.forward (+ ClipBoxes) for any batch size, padded outputs.

    With ``reg_exp_scales`` (one 1-element CUDA tensor per level: the head's ``ScaleExp.scale``), ``reg`` holds
    the RAW regression convolution outputs and ``exp(reg * scale)`` (HISFcos.py:228) is evaluated inside the
    decode, only for the <= max_box selected points.

    Returns scores [B,K] f32, classes [B,K] i64 (1-based), boxes [B,K,4] f32, keep [B,K] i64,
    counts [B] i32 with K = min(max_box, P); rows beyond counts[b] are unspecified.  All five share
    one allocation; ``out_packed`` (uint8, ``packed_nbytes(B, K)`` bytes, 256-byte aligned) lets the
    caller place it, e.g. inside a buffer that one collective gathers for several batches.
    """
    lib = _lib.load()
    lv, keep_alive, p_total, batch, n = _levels(cls, cnt, reg, strides, reg_exp_scales, native_half=True)
    dev = keep_alive[0].device
    k = min(int(max_box), p_total)
    if k > _lib.MAX_BOX:
        raise _lib.B200DetError(f"max_detection_box {k} > {_lib.MAX_BOX}")
    ncls = keep_alive[0].shape[1]
    ws_bytes = lib.b200det_postprocess_workspace_bytes(batch, p_total, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if out_packed is None:
        out_packed = packed_detections(batch, k, dev)
    elif out_packed.device != dev or out_packed.data_ptr() % 256:
        raise _lib.B200DetError("out_packed must live on the inputs' device and be 256-byte aligned")
    scores, classes, boxes, keep, counts = detection_views(out_packed, batch, k)
    ch, cw = (int(clip_hw[0]), int(clip_hw[1])) if clip_hw else (0, 0)
    with torch.cuda.device(dev):
        rc = lib.b200det_postprocess(lv, n, batch, ncls, float(score_thr), float(nms_thr), k, ch, cw,
                                     ws.data_ptr(), ws_bytes, scores.data_ptr(), classes.data_ptr(),
                                     boxes.data_ptr(), keep.data_ptr(), counts.data_ptr(), _stream(ws))
    _lib.check(rc, "b200det_postprocess")
    _count("postprocess")
    return scores, classes, boxes, keep, counts


def _packed_layout(batch: int, k: int):
    """Byte offsets of (scores f32, boxes f32x4, classes i64, keep i64, counts i32) in one buffer."""
    sizes = [batch * k * 4, batch * k * 16, batch * k * 8, batch * k * 8, batch * 4]
    offs, off = [], 0
    for sz in sizes:
        offs.append(off)
        off += (sz + 255) // 256 * 256
    return offs, sizes, off


def packed_nbytes(batch: int, k: int) -> int:
    """Bytes of the packed output buffer of a [batch, k] post-process call."""
    return _packed_layout(batch, k)[2]


def packed_detections(batch: int, k: int, device) -> Tensor:
    """One contiguous uint8 buffer holding every post-process output of a [batch, k] shard, so that a
    multi-rank gather is a single collective (see sharding.gather_packed)."""
    return torch.empty(_packed_layout(batch, k)[2], dtype=torch.uint8, device=device)


def detection_views(packed: Tensor, batch: int, k: int):
    """Typed views (scores, classes, boxes, keep, counts) into a packed_detections buffer (or into one
    rank's slice of a gathered buffer)."""
    offs, sizes, total = _packed_layout(batch, k)
    assert packed.numel() == total and packed.dtype == torch.uint8
    part = [packed[o:o + n] for o, n in zip(offs, sizes)]
    scores = part[0].view(torch.float32).view(batch, k)
    boxes = part[1].view(torch.float32).view(batch, k, 4)
    classes = part[2].view(torch.int64).view(batch, k)
    keep = part[3].view(torch.int64).view(batch, k)
    counts = part[4].view(torch.int32).view(batch)
    return scores, classes, boxes, keep, counts


def score_points(cls: Sequence[Tensor], cnt: Sequence[Tensor], strides: Sequence[int]):
    """K1 alone: score [B,P] f32, class argmax [B,P] i16 (0-based)."""
    lib = _lib.load()
    lv, keep_alive, p_total, batch, n = _levels(cls, cnt, None, strides, native_half=True)
    dev = keep_alive[0].device
    score = torch.empty((batch, p_total), dtype=torch.float32, device=dev)
    cls0 = torch.empty((batch, p_total), dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        rc = lib.b200det_score_points(lv, n, batch, keep_alive[0].shape[1], score.data_ptr(), cls0.data_ptr(),
                                      _stream(score))
    _lib.check(rc, "b200det_score_points")
    _count("score_points")
    return score, cls0


def select_topk(reg: Sequence[Tensor], strides: Sequence[int], score: Tensor, cls0: Tensor, score_thr: float,
                max_box: int):
    """K2 alone: (scores [B,K], classes [B,K] i32, boxes [B,K,4], points [B,K] i32, counts [B] i32)."""
    lib = _lib.load()
    lv, keep_alive, p_total, batch, n = _levels(None, None, reg, strides, native_half=True)
    dev = score.device
    k = min(int(max_box), p_total)
    assert score.shape == (batch, p_total) and score.is_contiguous() and cls0.is_contiguous()
    s = torch.empty((batch, k), dtype=torch.float32, device=dev)
    c = torch.empty((batch, k), dtype=torch.int32, device=dev)
    b = torch.empty((batch, k, 4), dtype=torch.float32, device=dev)
    p = torch.empty((batch, k), dtype=torch.int32, device=dev)
    cnt = torch.empty((batch,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.b200det_select_topk(lv, n, batch, score.data_ptr(), cls0.data_ptr(), float(score_thr), k,
                                     s.data_ptr(), c.data_ptr(), b.data_ptr(), p.data_ptr(), cnt.data_ptr(),
                                     _stream(score))
    _lib.check(rc, "b200det_select_topk")
    _count("select_topk")
    return s, c, b, p, cnt


def batched_nms(boxes: Tensor, scores: Tensor, classes: Tensor, score_thr: float, nms_thr: float,
                in_count: Tensor | None = None, clip_hw: Tuple[int, int] | None = None):
    """post_process (head.py:84-102) on [B,n] candidates: threshold -> torchvision-CPU-exact NMS.

    Returns scores [B,n], classes [B,n] i64, boxes [B,n,4], keep [B,n] i64 (index into the image's
    thresholded candidate list, as ``torchvision.ops.batched_nms`` returns), counts [B] i32.
    """
    lib = _lib.load()
    _need_cuda(boxes, "boxes")
    dev = boxes.device
    boxes = _f32c(boxes, "boxes")
    scores = _f32c(scores, "scores")
    _need_cuda(classes, "classes")
    classes = classes.to(torch.int64).contiguous()
    batch, n = scores.shape
    if boxes.shape != (batch, n, 4) or classes.shape != (batch, n):
        raise _lib.B200DetError("batched_nms expects boxes [B,n,4], scores [B,n], classes [B,n]")
    o_s = torch.empty((batch, n), dtype=torch.float32, device=dev)
    o_c = torch.empty((batch, n), dtype=torch.int64, device=dev)
    o_b = torch.empty((batch, n, 4), dtype=torch.float32, device=dev)
    o_k = torch.empty((batch, n), dtype=torch.int64, device=dev)
    counts = torch.zeros((batch,), dtype=torch.int32, device=dev)
    if n == 0 or batch == 0:
        return o_s, o_c, o_b, o_k, counts
    if n > _lib.MAX_BOX:
        raise _lib.B200DetError(f"{n} candidates per image > {_lib.MAX_BOX}")
    if in_count is not None:
        in_count = in_count.to(device=dev, dtype=torch.int32).contiguous()
    ws_bytes = lib.b200det_nms_workspace_bytes(batch, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    ch, cw = (int(clip_hw[0]), int(clip_hw[1])) if clip_hw else (0, 0)
    with torch.cuda.device(dev):
        rc = lib.b200det_batched_nms(batch, n, boxes.data_ptr(), scores.data_ptr(), classes.data_ptr(),
                                     in_count.data_ptr() if in_count is not None else None,
                                     float(score_thr), float(nms_thr), ch, cw, ws.data_ptr(), ws_bytes,
                                     o_s.data_ptr(), o_c.data_ptr(), o_b.data_ptr(), o_k.data_ptr(),
                                     counts.data_ptr(), _stream(ws))
    _lib.check(rc, "b200det_batched_nms")
    _count("batched_nms")
    return o_s, o_c, o_b, o_k, counts


def clip_boxes_(boxes: Tensor, img_h: int, img_w: int) -> Tensor:
    """ClipBoxes.forward: in-place, returns the same tensor (head.py:156-162)."""
    lib = _lib.load()
    _need_cuda(boxes, "boxes")
    if boxes.numel() == 0:
        return boxes
    if boxes.shape[-1] != 4:
        raise _lib.B200DetError("boxes must end in a dimension of 4")
    work = boxes
    direct = boxes.dtype == torch.float32 and boxes.is_contiguous() and boxes.data_ptr() % 16 == 0
    if not direct:
        work = boxes.float().contiguous()
    with torch.cuda.device(boxes.device):
        rc = lib.b200det_clip_boxes(work.data_ptr(), work.numel() // 4, int(img_h), int(img_w), _stream(work))
    _lib.check(rc, "b200det_clip_boxes")
    _count("clip_boxes")
    if not direct:
        boxes.copy_(work)
    return boxes


# --------------------------------------------------------------------------------------------
# training targets
# --------------------------------------------------------------------------------------------
def assign_targets(level_hw: Sequence[Tuple[int, int]], strides: Sequence[int],
                   limit_range: Sequence[Sequence[float]], gt_boxes: Tensor, labels: Tensor,
                   sample_radius: float = 1.5, want_index: bool = False):
    """FCOSGenTargets.forward: cls_t [B,P,1] i64, cnt_t [B,P,1] f32, reg_t [B,P,4] f32 (+ gt index [B,P] i32)."""
    lib = _lib.load()
    _need_cuda(gt_boxes, "gt_boxes")
    _need_cuda(labels, "labels")
    dev = gt_boxes.device
    gt = _f32c(gt_boxes, "gt_boxes")
    lab = labels.to(torch.int64).contiguous()
    if gt.dim() != 3 or gt.shape[-1] != 4 or lab.shape != gt.shape[:2]:
        raise _lib.B200DetError("expected gt_boxes [B,M,4] and labels [B,M]")
    batch, m = lab.shape
    n = len(level_hw)
    assert len(strides) == n and len(limit_range) == n
    p_total = sum(h * w for h, w in level_hw)
    hw_arr = (C.c_int32 * (2 * n))(*[v for hw in level_hw for v in hw])
    st_arr = (C.c_int32 * n)(*[int(s) for s in strides])
    lo_arr = (C.c_float * n)(*[float(r[0]) for r in limit_range])
    hi_arr = (C.c_float * n)(*[float(r[1]) for r in limit_range])
    ra_arr = (C.c_float * n)(*[float(s * sample_radius) for s in strides])
    cls_t = torch.empty((batch, p_total, 1), dtype=torch.int64, device=dev)
    cnt_t = torch.empty((batch, p_total, 1), dtype=torch.float32, device=dev)
    reg_t = torch.empty((batch, p_total, 4), dtype=torch.float32, device=dev)
    idx = torch.empty((batch, p_total), dtype=torch.int32, device=dev) if want_index else None
    with torch.cuda.device(dev):
        rc = lib.b200det_assign_targets(hw_arr, st_arr, lo_arr, hi_arr, ra_arr, n, batch, m, gt.data_ptr(),
                                        lab.data_ptr(), cls_t.data_ptr(), cnt_t.data_ptr(), reg_t.data_ptr(),
                                        idx.data_ptr() if idx is not None else None, _stream(gt))
    _lib.check(rc, "b200det_assign_targets")
    _count("assign_targets")
    return (cls_t, cnt_t, reg_t, idx) if want_index else (cls_t, cnt_t, reg_t)


# --------------------------------------------------------------------------------------------
# losses (forward / backward pairs; autograd glue lives in loss.py)
# --------------------------------------------------------------------------------------------
def _mask_src(mask_src: Tensor, batch: int, p_total: int) -> Tensor:
    m = _f32c(mask_src, "mask source").reshape(batch, -1)
    if m.shape[1] != p_total:
        raise AssertionError(f"targets cover {m.shape[1]} points, predictions {p_total}")
    return m


def _grad_ptrs(grads: Sequence[Tensor]):
    return (C.c_void_p * len(grads))(*[g.data_ptr() for g in grads])


def box_loss_fwd(reg: Sequence[Tensor], mask_src: Tensor, reg_t: Tensor, mode: int):
    lib = _lib.load()
    lv, keep_alive, p_total, batch, n = _levels(None, None, reg, [1] * len(reg))
    m = _mask_src(mask_src, batch, p_total)
    t = _f32c(reg_t, "reg target").reshape(batch, p_total, 4)
    loss = torch.empty((batch,), dtype=torch.float32, device=m.device)
    npos = torch.empty_like(loss)
    with torch.cuda.device(m.device):
        rc = lib.b200det_box_loss_fwd(lv, n, batch, m.data_ptr(), t.data_ptr(), mode, loss.data_ptr(),
                                      npos.data_ptr(), _stream(m))
    _lib.check(rc, "b200det_box_loss_fwd")
    _count("box_loss_fwd")
    return loss, npos


def box_loss_bwd(reg: Sequence[Tensor], mask_src: Tensor, reg_t: Tensor, mode: int, grad_loss: Tensor,
                 npos: Tensor) -> List[Tensor]:
    lib = _lib.load()
    lv, keep_alive, p_total, batch, n = _levels(None, None, reg, [1] * len(reg))
    m = _mask_src(mask_src, batch, p_total)
    t = _f32c(reg_t, "reg target").reshape(batch, p_total, 4)
    g = _f32c(grad_loss, "grad").reshape(batch)
    grads = [torch.empty_like(x) for x in keep_alive]
    with torch.cuda.device(m.device):
        rc = lib.b200det_box_loss_bwd(lv, _grad_ptrs(grads), n, batch, m.data_ptr(), t.data_ptr(), mode,
                                      g.data_ptr(), npos.data_ptr(), _stream(m))
    _lib.check(rc, "b200det_box_loss_bwd")
    _count("box_loss_bwd")
    return grads


def cnt_loss_fwd(cnt: Sequence[Tensor], mask_src: Tensor, cnt_t: Tensor):
    lib = _lib.load()
    lv, keep_alive, p_total, batch, n = _levels(None, cnt, None, [1] * len(cnt))
    m = _mask_src(mask_src, batch, p_total)
    t = _mask_src(cnt_t, batch, p_total)
    loss = torch.empty((batch,), dtype=torch.float32, device=m.device)
    npos = torch.empty_like(loss)
    with torch.cuda.device(m.device):
        rc = lib.b200det_cnt_loss_fwd(lv, n, batch, m.data_ptr(), t.data_ptr(), loss.data_ptr(), npos.data_ptr(),
                                      _stream(m))
    _lib.check(rc, "b200det_cnt_loss_fwd")
    _count("cnt_loss_fwd")
    return loss, npos


def cnt_loss_bwd(cnt: Sequence[Tensor], mask_src: Tensor, cnt_t: Tensor, grad_loss: Tensor, npos: Tensor):
    lib = _lib.load()
    lv, keep_alive, p_total, batch, n = _levels(None, cnt, None, [1] * len(cnt))
    m = _mask_src(mask_src, batch, p_total)
    t = _mask_src(cnt_t, batch, p_total)
    g = _f32c(grad_loss, "grad").reshape(batch)
    grads = [torch.empty_like(x) for x in keep_alive]
    with torch.cuda.device(m.device):
        rc = lib.b200det_cnt_loss_bwd(lv, _grad_ptrs(grads), n, batch, m.data_ptr(), t.data_ptr(), g.data_ptr(),
                                      npos.data_ptr(), _stream(m))
    _lib.check(rc, "b200det_cnt_loss_bwd")
    _count("cnt_loss_bwd")
    return grads


def cls_loss_fwd(cls: Sequence[Tensor], mask_src: Tensor, cls_t: Tensor):
    lib = _lib.load()
    lv, keep_alive, p_total, batch, n = _levels(cls, None, None, [1] * len(cls))
    m = _mask_src(mask_src, batch, p_total)
    _need_cuda(cls_t, "cls target")
    t = cls_t.to(torch.int64).reshape(batch, -1).contiguous()
    assert t.shape[1] == p_total
    ws_bytes = lib.b200det_cls_loss_workspace_bytes(batch, p_total, keep_alive[0].shape[1])
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=m.device)
    loss = torch.empty((batch,), dtype=torch.float32, device=m.device)
    npos = torch.empty_like(loss)
    with torch.cuda.device(m.device):
        rc = lib.b200det_cls_loss_fwd(lv, n, batch, keep_alive[0].shape[1], t.data_ptr(), m.data_ptr(),
                                      ws.data_ptr(), ws_bytes, loss.data_ptr(), npos.data_ptr(), _stream(m))
    _lib.check(rc, "b200det_cls_loss_fwd")
    _count("cls_loss_fwd")
    return loss, npos


def cls_loss_bwd(cls: Sequence[Tensor], cls_t: Tensor, grad_loss: Tensor, npos: Tensor):
    lib = _lib.load()
    lv, keep_alive, p_total, batch, n = _levels(cls, None, None, [1] * len(cls))
    t = cls_t.to(torch.int64).reshape(batch, -1).contiguous()
    g = _f32c(grad_loss, "grad").reshape(batch)
    grads = [torch.empty_like(x) for x in keep_alive]
    with torch.cuda.device(g.device):
        rc = lib.b200det_cls_loss_bwd(lv, _grad_ptrs(grads), n, batch, keep_alive[0].shape[1], t.data_ptr(),
                                      g.data_ptr(), npos.data_ptr(), _stream(g))
    _lib.check(rc, "b200det_cls_loss_bwd")
    _count("cls_loss_bwd")
    return grads


def cls_loss_step(cls: Sequence[Tensor], cls_t: Tensor, mask_src: Tensor | None = None,
                  num_pos: Tensor | None = None, grad_loss: Tensor | None = None, up_mean: Tensor | None = None):
    """compute_cls_loss forward AND backward from one read of the logits (b200det_cls_loss_step).

    ``num_pos`` [B] (from the fused assignment step) or ``mask_src`` (cnt_t, > -1 = positive) must be given.
    fp16 / bf16 class maps are read as they are and their gradients come back in the same type (fp32 arithmetic).
    Returns (loss [B], mean [2] = {batch mean, upstream gradient the gradients were written for}, num_pos [B], grads) with grads = d(sum_b grad_loss[b] * loss[b]) / d(cls maps),
    grad_loss defaulting to 1/B (the gradient of the batch mean); ``up_mean`` (ONE fp32 CUDA value, exclusive
    with grad_loss) is the upstream gradient of the batch mean, e.g. a GradScaler's loss scale: grad_loss[b] = up / B."""
    lib = _lib.load()
    half = all(x.dtype == cls[0].dtype for x in cls) and cls[0].dtype in (torch.float16, torch.bfloat16)
    lv, keep_alive, p_total, batch, n = _levels(cls, None, None, [1] * len(cls), keep_half_cls=half)
    dev = keep_alive[0].device
    dtype_code = _DTYPE_CODE[keep_alive[0].dtype]
    _need_cuda(cls_t, "cls target")
    t = cls_t.to(torch.int64).reshape(batch, -1).contiguous()
    assert t.shape[1] == p_total
    if num_pos is None and mask_src is None:
        raise _lib.B200DetError("cls_loss_step needs num_pos or the positive-mask source (cnt_t)")
    m = _mask_src(mask_src, batch, p_total) if num_pos is None else None
    ready = num_pos is not None
    npos = _f32c(num_pos, "num_pos").reshape(batch) if ready else torch.empty((batch,), dtype=torch.float32, device=dev)
    gl = _f32c(grad_loss, "grad_loss").reshape(batch) if grad_loss is not None else None
    if up_mean is not None:
        assert grad_loss is None and up_mean.is_cuda and up_mean.dtype == torch.float32 and up_mean.numel() >= 1
        gl = up_mean
    c = keep_alive[0].shape[1]
    ws_bytes = lib.b200det_cls_loss_workspace_bytes(batch, p_total, c)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    loss = torch.empty((batch,), dtype=torch.float32, device=dev)
    mean = torch.empty((2,), dtype=torch.float32, device=dev)     # {batch mean, the upstream gradient assumed}
    grads = [torch.empty_like(x) for x in keep_alive]
    ptr = lambda x: x.data_ptr() if x is not None else None
    with torch.cuda.device(dev):
        rc = lib.b200det_cls_loss_step(lv, _grad_ptrs(grads), dtype_code, n, batch, c, t.data_ptr(), ptr(m), ptr(gl),
                                       int(up_mean is not None), int(ready),
                                       ws.data_ptr(), ws_bytes, loss.data_ptr(), npos.data_ptr(), mean.data_ptr(),
                                       _stream(t))
    _lib.check(rc, "b200det_cls_loss_step")
    _count("cls_loss_step")
    if not ready:
        _count("count_pos")        # the positive count kernel ran first
    return loss, mean, npos, grads


# --------------------------------------------------------------------------------------------
# fused training step: targets + box / centerness loss forward and backward in one launch
# --------------------------------------------------------------------------------------------
_fused_ws = {}            # (device index, batch, P) -> workspace (tile partials)


def _fused_workspace(dev: torch.device, batch: int, p_total: int) -> Tensor:
    lib = _lib.load()
    n = int(lib.b200det_assign_loss_workspace_bytes(batch, p_total))
    key = (dev.index, batch, p_total)
    ws = _fused_ws.get(key)
    if ws is None:
        ws = torch.zeros(n, dtype=torch.uint8, device=dev)
        if not torch.cuda.is_current_stream_capturing():     # a tensor of a graph's private pool is not cached
            _fused_ws[key] = ws
    return ws


def assign_loss_fused(reg: Sequence[Tensor], cnt: Sequence[Tensor] | None, strides: Sequence[int],
                      limit_range: Sequence[Sequence[float]], gt_boxes: Tensor, labels: Tensor, mode: int,
                      sample_radius: float = 1.5, grad_box: Tensor | None = None, grad_cnt: Tensor | None = None,
                      want_mean: bool = True, workspace: Tensor | None = None,
                      reg_exp_scales: Sequence[Tensor] | None = None, up_box: Tensor | None = None,
                      up_cnt: Tensor | None = None):
    """FCOSGenTargets.forward + compute_reg_loss (+ compute_cnt_loss) forward AND backward, one kernel.

    Returns a dict: cls_t [B,P,1] i64, cnt_t [B,P,1], reg_t [B,P,4] (bit-identical to assign_targets),
    box_loss / cnt_loss / num_pos [B], mean [4] (batch means of box_loss, cnt_loss; the upstream gradients of the
    two means this call assumed), reg_grads /
    cnt_grads (lists shaped like the maps) = gradient of sum_b grad_*[b] * loss[b]; grad_* default to
    1/B, i.e. the gradient of the batch mean.  With ``reg_exp_scales`` (per level ONE fp32 CUDA value, the
    head's ScaleExp.scale) ``reg`` holds the raw regression outputs x, the distances are exp(x * scale),
    reg_grads are gradients w.r.t. x and ``scale_grad`` [n_levels] w.r.t. the scales.  ``up_box`` / ``up_cnt``
    (ONE fp32 CUDA value each, instead of grad_box / grad_cnt) are upstream gradients of the two batch MEANS.  Concurrent calls on different streams of one device need
    their own ``workspace`` (b200det_assign_loss_workspace_bytes(B, P) bytes).
    """
    lib = _lib.load()
    lv, keep_alive, p_total, batch, n = _levels(None, cnt, reg, strides, reg_exp_scales)
    if len(limit_range) < n:
        raise _lib.B200DetError("limit_range is shorter than the level list")
    _need_cuda(gt_boxes, "gt_boxes")
    _need_cuda(labels, "labels")
    dev = keep_alive[0].device
    gt = _f32c(gt_boxes, "gt_boxes")
    lab = labels.to(torch.int64).contiguous()
    if gt.dim() != 3 or gt.shape[-1] != 4 or lab.shape != gt.shape[:2] or gt.shape[0] != batch:
        raise _lib.B200DetError("expected gt_boxes [B,M,4] and labels [B,M] with the maps' batch size")
    m = lab.shape[1]
    lo_arr = (C.c_float * n)(*[float(r[0]) for r in limit_range[:n]])
    hi_arr = (C.c_float * n)(*[float(r[1]) for r in limit_range[:n]])
    ra_arr = (C.c_float * n)(*[float(s * sample_radius) for s in list(strides)[:n]])
    per = 2 if cnt is not None else 1
    maps = keep_alive[:per * n]                                         # (scale tensors follow the maps)
    maps_cnt = maps[0::per] if cnt is not None else None                # _levels appends (cnt, reg) per level
    maps_reg = maps[1::per] if cnt is not None else maps
    reg_grads = [torch.empty_like(x) for x in maps_reg]
    cnt_grads = [torch.empty_like(x) for x in maps_cnt] if cnt is not None else None
    cls_t = torch.empty((batch, p_total, 1), dtype=torch.int64, device=dev)
    cnt_t = torch.empty((batch, p_total, 1), dtype=torch.float32, device=dev)
    reg_t = torch.empty((batch, p_total, 4), dtype=torch.float32, device=dev)
    box_loss = torch.empty((batch,), dtype=torch.float32, device=dev)
    cnt_loss = torch.empty_like(box_loss) if cnt is not None else None
    num_pos = torch.empty_like(box_loss)
    mean = torch.empty((4,), dtype=torch.float32, device=dev) if want_mean else None   # means + assumed upstreams
    scale_grad = torch.empty((n,), dtype=torch.float32, device=dev) if reg_exp_scales is not None else None
    ws = workspace if workspace is not None else _fused_workspace(dev, batch, p_total)
    gb = _f32c(grad_box, "grad_box").reshape(batch) if grad_box is not None else None
    gc = _f32c(grad_cnt, "grad_cnt").reshape(batch) if grad_cnt is not None else None
    grad_mode = 0
    if up_box is not None or up_cnt is not None:
        assert grad_box is None and grad_cnt is None and up_box is not None and (cnt is None or up_cnt is not None)
        for u in (up_box, up_cnt):
            assert u is None or (u.is_cuda and u.dtype == torch.float32 and u.numel() >= 1)
        gb, gc, grad_mode = up_box, up_cnt, 1
    ptr = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(dev):
        rc = lib.b200det_assign_loss_fused(lv, _grad_ptrs(reg_grads), _grad_ptrs(cnt_grads) if cnt_grads else None, n,
                                           lo_arr, hi_arr, ra_arr, batch, m, gt.data_ptr(), lab.data_ptr(), int(mode),
                                           ptr(gb), ptr(gc), grad_mode, cls_t.data_ptr(), cnt_t.data_ptr(), reg_t.data_ptr(),
                                           box_loss.data_ptr(), ptr(cnt_loss), num_pos.data_ptr(), ptr(mean),
                                           ptr(scale_grad), ptr(ws), ws.numel(), _stream(gt))
    _lib.check(rc, "b200det_assign_loss_fused")
    _count("assign_loss_fused")
    return {"cls_t": cls_t, "cnt_t": cnt_t, "reg_t": reg_t, "box_loss": box_loss, "cnt_loss": cnt_loss,
            "num_pos": num_pos, "mean": mean, "reg_grads": reg_grads, "cnt_grads": cnt_grads,
            "scale_grad": scale_grad}


def scale_maps_(maps: Sequence[Tensor], factors: Sequence[Tensor]) -> None:
    """maps[i] *= factors[i] (0-dim fp32 CUDA tensors) in one launch; a factor of exactly 1 touches nothing."""
    lib = _lib.load()
    n = len(maps)
    assert n == len(factors) and n > 0
    for t, f in zip(maps, factors):
        _need_cuda(t, "map")
        assert t.dtype == torch.float32 and t.is_contiguous() and f.dtype == torch.float32 and f.numel() == 1
    m_arr = (C.c_void_p * n)(*[t.data_ptr() for t in maps])
    n_arr = (C.c_int64 * n)(*[t.numel() for t in maps])
    f_arr = (C.c_void_p * n)(*[f.data_ptr() for f in factors])
    with torch.cuda.device(maps[0].device):
        rc = lib.b200det_scale_maps(m_arr, n_arr, f_arr, n, _stream(maps[0]))
    _lib.check(rc, "b200det_scale_maps")
    _count("scale_maps")


def rescale_maps_(maps: Sequence[Tensor], got: Sequence[Tensor], state: Sequence[Tensor],
                  assumed: Sequence[Tensor] | None = None) -> None:
    """maps[i] *= got[i] / assumed[i] where the two differ (nothing is touched where they are equal), then
    state[i][0] <- got[i]: the backward of the steps whose forward wrote gradients for an assumed upstream
    gradient (b200det_rescale_maps, one launch).  got: 1-element, state: 2-element {next assumption, ticket},
    assumed: 1-element fp32 CUDA tensors — the value the forward being differentiated read (its own copy: the
    shared state may have moved on when several forwards were outstanding); None = the state's current word.
    Maps that share a state object share its got / assumed."""
    lib = _lib.load()
    n = len(maps)
    assert n == len(got) == len(state) and n > 0 and (assumed is None or len(assumed) == n)
    states, gots, assumes, index = [], [], [], []
    for i, (t, g, a) in enumerate(zip(maps, got, state)):
        _need_cuda(t, "map")
        assert t.dtype == maps[0].dtype and t.dtype in _DTYPE_CODE and t.is_contiguous()
        assert g.dtype == torch.float32 and g.numel() == 1 and g.is_cuda
        assert a.dtype == torch.float32 and a.numel() == 2 and a.is_cuda and a.is_contiguous()
        for k, s0 in enumerate(states):
            if s0.data_ptr() == a.data_ptr():
                index.append(k)
                break
        else:
            index.append(len(states))
            states.append(a)
            gots.append(g)
            if assumed is not None:
                u = assumed[i]
                assert u.dtype == torch.float32 and u.numel() == 1 and u.is_cuda
                assumes.append(u)
    m_arr = (C.c_void_p * n)(*[t.data_ptr() for t in maps])
    n_arr = (C.c_int64 * n)(*[t.numel() for t in maps])
    i_arr = (C.c_int32 * n)(*index)
    g_arr = (C.c_void_p * len(states))(*[g.data_ptr() for g in gots])
    s_arr = (C.c_void_p * len(states))(*[a.data_ptr() for a in states])
    a_arr = (C.c_void_p * len(states))(*[u.data_ptr() for u in assumes]) if assumed is not None else None
    with torch.cuda.device(maps[0].device):
        rc = lib.b200det_rescale_maps(m_arr, n_arr, i_arr, _DTYPE_CODE[maps[0].dtype], n, g_arr, a_arr, s_arr, len(states),
                                      _stream(maps[0]))
    _lib.check(rc, "b200det_rescale_maps")
    _count("rescale_maps")


# --------------------------------------------------------------------------------------------
# torch.library registration: every entry point as a dispatcher op of the ``b200det`` namespace — CUDA kernels
# only (a CPU tensor finds no kernel and raises) plus a fake (meta) implementation that knows the output shapes, so
# the ops can be traced / exported without running them.  Absent optional outputs come back as empty tensors.
# --------------------------------------------------------------------------------------------
_LIBDEF = torch.library.Library("b200det", "DEF")
OP_NAMES: List[str] = []


def _register(schema: str, impl, fake) -> None:
    name = schema.split("(", 1)[0]
    _LIBDEF.define(schema)
    _LIBDEF.impl(name, impl, "CUDA")
    torch.library.register_fake(f"b200det::{name}")(fake)
    OP_NAMES.append(name)


def _points(levels) -> int:
    return sum(int(t.shape[2]) * int(t.shape[3]) for t in levels)


def _zip_n(strides, *lists) -> int:
    return min([len(strides)] + [len(l) for l in lists if l is not None])


def _e(like: Tensor, shape, dtype=None) -> Tensor:
    return like.new_empty(tuple(shape), dtype=dtype if dtype is not None else like.dtype)


# ---- inference --------------------------------------------------------------------------------------------
def _op_postprocess(cls, cnt, reg, strides, score_thr, nms_thr, max_box, clip_h, clip_w):
    return postprocess(cls, cnt, reg, strides, score_thr, nms_thr, max_box, (clip_h, clip_w) if clip_h > 0 else None)


def _fake_postprocess(cls, cnt, reg, strides, score_thr, nms_thr, max_box, clip_h, clip_w):
    n = _zip_n(strides, cls, cnt, reg)
    b, k = cls[0].shape[0], min(int(max_box), _points(cls[:n]))
    x = cls[0]
    return (_e(x, (b, k), torch.float32), _e(x, (b, k), torch.int64), _e(x, (b, k, 4), torch.float32),
            _e(x, (b, k), torch.int64), _e(x, (b,), torch.int32))


def _op_batched_nms(boxes, scores, classes, score_thr, nms_thr, clip_h, clip_w):
    return batched_nms(boxes, scores, classes, score_thr, nms_thr, None, (clip_h, clip_w) if clip_h > 0 else None)


def _fake_batched_nms(boxes, scores, classes, score_thr, nms_thr, clip_h, clip_w):
    b, n = scores.shape
    return (_e(scores, (b, n), torch.float32), _e(scores, (b, n), torch.int64), _e(scores, (b, n, 4), torch.float32),
            _e(scores, (b, n), torch.int64), _e(scores, (b,), torch.int32))


def _fake_score_points(cls, cnt, strides):
    n = _zip_n(strides, cls, cnt)
    b, p = cls[0].shape[0], _points(cls[:n])
    return _e(cls[0], (b, p), torch.float32), _e(cls[0], (b, p), torch.int16)


def _fake_select_topk(reg, strides, score, cls0, score_thr, max_box):
    b, k = score.shape[0], min(int(max_box), _points(reg[:_zip_n(strides, reg)]))
    return (_e(score, (b, k), torch.float32), _e(score, (b, k), torch.int32), _e(score, (b, k, 4), torch.float32),
            _e(score, (b, k), torch.int32), _e(score, (b,), torch.int32))


# ---- training targets -------------------------------------------------------------------------------------
def _op_assign(level_hw, strides, limit_lo, limit_hi, gt_boxes, labels, sample_radius):
    hw = [(level_hw[2 * i], level_hw[2 * i + 1]) for i in range(len(strides))]
    return assign_targets(hw, strides, list(zip(limit_lo, limit_hi)), gt_boxes, labels, sample_radius)


def _fake_assign(level_hw, strides, limit_lo, limit_hi, gt_boxes, labels, sample_radius):
    b, p = gt_boxes.shape[0], sum(level_hw[2 * i] * level_hw[2 * i + 1] for i in range(len(strides)))
    return (_e(gt_boxes, (b, p, 1), torch.int64), _e(gt_boxes, (b, p, 1), torch.float32),
            _e(gt_boxes, (b, p, 4), torch.float32))


# ---- losses -----------------------------------------------------------------------------------------------
def _fake_loss_fwd(maps, *_):
    b = maps[0].shape[0]
    return _e(maps[0], (b,), torch.float32), _e(maps[0], (b,), torch.float32)


def _fake_loss_bwd(maps, *_):
    return [_e(t, t.shape, torch.float32) for t in maps]


def _op_cls_loss_step(cls, cls_t, mask_src, num_pos, grad_loss, up_mean):
    loss, mean, npos, grads = cls_loss_step(cls, cls_t, mask_src, num_pos, grad_loss, up_mean)
    return loss, mean, npos, grads


def _fake_cls_loss_step(cls, cls_t, mask_src, num_pos, grad_loss, up_mean):
    b = cls[0].shape[0]
    f = lambda shape: _e(cls[0], shape, torch.float32)                         # noqa: E731
    return f((b,)), f((2,)), f((b,)), [_e(t, t.shape) for t in cls]            # gradients in the logits' own dtype


def _op_assign_loss_fused(reg, cnt, strides, limit_lo, limit_hi, gt_boxes, labels, mode, sample_radius, up_box, up_cnt,
                          reg_exp_scales):
    r = assign_loss_fused(reg, cnt, strides, list(zip(limit_lo, limit_hi)), gt_boxes, labels, mode, sample_radius,
                          reg_exp_scales=reg_exp_scales, up_box=up_box, up_cnt=up_cnt)
    none = gt_boxes.new_empty((0,))
    return (r["cls_t"], r["cnt_t"], r["reg_t"], r["box_loss"], r["cnt_loss"] if r["cnt_loss"] is not None else none,
            r["num_pos"], r["mean"], r["reg_grads"], r["cnt_grads"] or [],
            r["scale_grad"] if r["scale_grad"] is not None else none)


def _fake_assign_loss_fused(reg, cnt, strides, limit_lo, limit_hi, gt_boxes, labels, mode, sample_radius, up_box, up_cnt,
                            reg_exp_scales):
    n = _zip_n(strides, reg, cnt)
    b, p = reg[0].shape[0], _points(reg[:n])
    f = lambda shape, dt=torch.float32: _e(reg[0], shape, dt)                  # noqa: E731
    return (f((b, p, 1), torch.int64), f((b, p, 1)), f((b, p, 4)), f((b,)), f((b,) if cnt is not None else (0,)), f((b,)),
            f((4,)), [f(t.shape) for t in reg[:n]], [f(t.shape) for t in cnt[:n]] if cnt is not None else [],
            f((n,) if reg_exp_scales is not None else (0,)))


def _op_rescale_maps(maps, got, state, assumed):
    rescale_maps_(maps, got, state, assumed)


_register("postprocess(Tensor[] cls, Tensor[] cnt, Tensor[] reg, int[] strides, float score_thr, float nms_thr, "
          "int max_box, int clip_h, int clip_w) -> (Tensor, Tensor, Tensor, Tensor, Tensor)", _op_postprocess, _fake_postprocess)
_register("batched_nms(Tensor boxes, Tensor scores, Tensor classes, float score_thr, float nms_thr, int clip_h, "
          "int clip_w) -> (Tensor, Tensor, Tensor, Tensor, Tensor)", _op_batched_nms, _fake_batched_nms)
_register("score_points(Tensor[] cls, Tensor[] cnt, int[] strides) -> (Tensor, Tensor)", score_points, _fake_score_points)
_register("select_topk(Tensor[] reg, int[] strides, Tensor score, Tensor cls0, float score_thr, int max_box) -> "
          "(Tensor, Tensor, Tensor, Tensor, Tensor)", select_topk, _fake_select_topk)
_register("clip_boxes_(Tensor(a!) boxes, int img_h, int img_w) -> Tensor(a!)", clip_boxes_, lambda boxes, h, w: boxes)
_register("assign_targets(int[] level_hw, int[] strides, float[] limit_lo, float[] limit_hi, Tensor gt_boxes, "
          "Tensor labels, float sample_radius) -> (Tensor, Tensor, Tensor)", _op_assign, _fake_assign)
_register("box_loss_fwd(Tensor[] reg, Tensor mask_src, Tensor reg_t, int mode) -> (Tensor, Tensor)", box_loss_fwd, _fake_loss_fwd)
_register("box_loss_bwd(Tensor[] reg, Tensor mask_src, Tensor reg_t, int mode, Tensor grad_loss, Tensor npos) -> Tensor[]",
          box_loss_bwd, _fake_loss_bwd)
_register("cnt_loss_fwd(Tensor[] cnt, Tensor mask_src, Tensor cnt_t) -> (Tensor, Tensor)", cnt_loss_fwd, _fake_loss_fwd)
_register("cnt_loss_bwd(Tensor[] cnt, Tensor mask_src, Tensor cnt_t, Tensor grad_loss, Tensor npos) -> Tensor[]",
          cnt_loss_bwd, _fake_loss_bwd)
_register("cls_loss_fwd(Tensor[] cls, Tensor mask_src, Tensor cls_t) -> (Tensor, Tensor)", cls_loss_fwd, _fake_loss_fwd)
_register("cls_loss_bwd(Tensor[] cls, Tensor cls_t, Tensor grad_loss, Tensor npos) -> Tensor[]", cls_loss_bwd, _fake_loss_bwd)
_register("cls_loss_step(Tensor[] cls, Tensor cls_t, Tensor? mask_src, Tensor? num_pos, Tensor? grad_loss, Tensor? up_mean)"
          " -> (Tensor, Tensor, Tensor, Tensor[])", _op_cls_loss_step, _fake_cls_loss_step)
_register("assign_loss_fused(Tensor[] reg, Tensor[]? cnt, int[] strides, float[] limit_lo, float[] limit_hi, "
          "Tensor gt_boxes, Tensor labels, int mode, float sample_radius, Tensor? up_box, Tensor? up_cnt, "
          "Tensor[]? reg_exp_scales) -> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor[], Tensor[], Tensor)",
          _op_assign_loss_fused, _fake_assign_loss_fused)
_register("rescale_maps_(Tensor(a!)[] maps, Tensor[] got, Tensor(b!)[] state, Tensor[]? assumed) -> ()", _op_rescale_maps,
          lambda maps, got, state, assumed: None)
