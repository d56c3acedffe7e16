"""CPU-side tests: the C-ABI library loads and exports every symbol of include/b200det.h, argument
validation happens before any CUDA call, the drop-in modules keep the reference's error behaviour,
and the multi-rank sharding / gather logic is exact (gloo, world_size 2).  No GPU needed."""
import ctypes as C
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pytorch_object_detection_b200 as P
from oracle import fcos_oracle as O
from pytorch_object_detection_b200 import _lib, sharding, workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "b200det.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200det_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/b200det.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototypes and header disagree"
    assert lib.b200det_abi_version() == _lib.ABI_VERSION
    assert lib.b200det_status_string(0) == b"ok"
    assert b"workspace" in lib.b200det_status_string(3)


def test_header_is_plain_c():
    import subprocess
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                          os.path.join(ROOT, "include", "b200det.h")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_workspace_sizes_are_host_only_and_monotone():
    lib = _lib.load()
    a = lib.b200det_postprocess_workspace_bytes(16, 23265, 1000)
    b = lib.b200det_postprocess_workspace_bytes(32, 23265, 1000)
    c = lib.b200det_postprocess_workspace_bytes(16, 23265, 5000)
    assert 0 < a < b and a < c
    assert a >= 16 * 23265 * 6 + 16 * 16 * 16 * 64 * 8          # scores + classes + mask tiles
    assert lib.b200det_postprocess_workspace_bytes(16, 23265, _lib.MAX_BOX + 1) == 0
    assert lib.b200det_nms_workspace_bytes(1, 0) == 0
    assert lib.b200det_nms_workspace_bytes(4, 1000) > 4 * 1000 * 44
    assert lib.b200det_cls_loss_workspace_bytes(32, 23265, 80) >= 32 * 46 * 5 * 4


def test_c_abi_rejects_bad_arguments_before_touching_the_gpu():
    lib = _lib.load()
    lv = _lib.make_levels([(0, 0, 0, 4, 4, 8)])
    # null outputs / workspace
    assert lib.b200det_postprocess(lv, 1, 1, 20, 0.05, 0.6, 100, 0, 0, None, 0, None, None, None, None, None, None) == 1
    assert lib.b200det_score_points(lv, 1, 1, 20, None, None, None) == 1
    assert lib.b200det_score_points(lv, 0, 1, 20, None, None, None) == 1
    assert lib.b200det_batched_nms(1, 10, None, None, None, None, 0.0, 0.5, 0, 0, None, 0, None, None, None, None,
                                   None, None) == 1
    assert lib.b200det_assign_targets(None, None, None, None, None, 1, 1, 1, None, None, None, None, None, None,
                                      None) == 1
    assert lib.b200det_clip_boxes(None, 5, 10, 10, None) == 1
    assert lib.b200det_clip_boxes(None, 0, 10, 10, None) == 0           # nothing to do is fine
    with pytest.raises(_lib.B200DetError):
        _lib.make_levels([(0, 0, 0, 1, 1, 1)] * 9)                      # more than B200DET_MAX_LEVELS
    with pytest.raises(_lib.B200DetError, match="bad argument"):
        _lib.check(1, "x")


def test_training_step_entry_points_validate_before_launching():
    """b200det_cls_loss_step / b200det_assign_loss_fused / b200det_rescale_maps: null pointers, unknown
    grad_mode / dtype, duplicate states and undersized workspaces are refused on the host."""
    lib = _lib.load()
    one = C.c_void_p(16)                                                  # a non-null, 16-byte aligned dummy address
    lv = _lib.make_levels([(16, 16, 16, 4, 4, 8)])
    grads = (C.c_void_p * 1)(16)
    ws_bytes = lib.b200det_cls_loss_workspace_bytes(1, 16, 20)
    step = lambda **kw: lib.b200det_cls_loss_step(                        # noqa: E731
        lv, kw.get("grads", grads), kw.get("dtype", 0), 1, 1, 20, one, one, None, kw.get("grad_mode", 0), 0,
        kw.get("ws", one), kw.get("ws_bytes", ws_bytes), one, one, None, None)
    assert step(grads=None) == 1
    assert step(grad_mode=2) == 1
    assert step(dtype=7) == 2                                             # B200DET_ERR_UNSUPPORTED
    assert step(ws=None) == 1
    assert step(ws_bytes=4) == 3                                          # B200DET_ERR_WORKSPACE
    maps = (C.c_void_p * 2)(16, 32)
    numel = (C.c_int64 * 2)(4, 4)
    idx = (C.c_int32 * 2)(0, 1)
    got = (C.c_void_p * 2)(16, 16)
    states = (C.c_void_p * 2)(64, 128)
    assert lib.b200det_rescale_maps(maps, numel, idx, 0, 0, got, None, states, 2, None) == 1        # no maps
    assert lib.b200det_rescale_maps(maps, numel, idx, 0, 17, got, None, states, 2, None) == 1       # too many maps
    assert lib.b200det_rescale_maps(maps, numel, idx, 9, 2, got, None, states, 2, None) == 2        # unknown dtype
    assert lib.b200det_rescale_maps(maps, numel, (C.c_int32 * 2)(0, 2), 0, 2, got, None, states, 2, None) == 1   # state index
    assert lib.b200det_rescale_maps(maps, numel, idx, 0, 2, got, got, (C.c_void_p * 2)(64, 64), 2, None) == 1   # duplicate state
    assert lib.b200det_rescale_maps(maps, numel, idx, 0, 2, got, None, states, 5, None) == 1        # too many states
    assert lib.b200det_assign_loss_workspace_bytes(0, 100) == 0
    lo = (C.c_float * 1)(-1.0)
    hi = (C.c_float * 1)(64.0)
    ra = (C.c_float * 1)(12.0)
    fused = lambda grad_mode, mode=1: lib.b200det_assign_loss_fused(       # noqa: E731
        lv, grads, None, 1, lo, hi, ra, 1, 0, None, None, mode, None, None, grad_mode, one, one, one, one, None, one,
        None, None, one, 1 << 20, None)
    assert fused(2) == 2 and fused(0, mode=5) == 2
    # the half-precision class maps of autocast keep their dtype through the Python level table only when asked
    from pytorch_object_detection_b200 import ops
    assert ops._DTYPE_CODE == {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def test_upstream_state_is_lazy_and_per_device():
    from pytorch_object_detection_b200.loss import _Upstream
    up = _Upstream()
    assert up.observed is False and up._value == {}
    step = P.FCOSTargetLoss(W.STRIDES, W.FCOS_RANGES, "giou")
    assert all(isinstance(u, _Upstream) for u in (step._up_cls, step._up_box, step._up_cnt))
    assert isinstance(P.FCOSLoss()._up_cls, _Upstream)


def test_modules_have_no_cpu_path_and_keep_reference_errors():
    x = W.head_outputs(1, 20, W.VOC_LEVELS, seed=3)
    head = P.FCOSHead(0.05, 0.6, 1000, W.STRIDES)
    assert (head.score, head.nms_threshold, head.max_box, head.strides) == (0.05, 0.6, 1000, W.STRIDES)
    with pytest.raises(Exception):
        head(x)                                                          # CPU tensors: refused, not emulated
    with pytest.raises(Exception):
        P.ClipBoxes()(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 4))
    gen = P.FCOSGenTargets(W.STRIDES, W.FCOS_RANGES)
    with pytest.raises(AssertionError):
        P.FCOSGenTargets(W.STRIDES, W.FCOS_RANGES[:4])                   # head.py:216
    with pytest.raises(AssertionError):
        gen([[x[0][:4], x[1][:4], x[2][:4]], torch.zeros(1, 2, 4), torch.zeros(1, 2, dtype=torch.long)])  # head.py:225
    with pytest.raises(Exception):
        gen([x, torch.zeros(1, 2, 4), torch.zeros(1, 2, dtype=torch.long)])
    with pytest.raises(NotImplementedError):                             # loss.py:137-138
        P.compute_reg_loss(x[2], torch.zeros(1, W.num_points(W.VOC_LEVELS), 4),
                           torch.zeros(1, W.num_points(W.VOC_LEVELS), dtype=torch.bool), mode="diou")
    with pytest.raises(AssertionError):                                  # loss.py:127 shape check
        P.compute_reg_loss(x[2], torch.zeros(1, 7, 4), torch.zeros(1, 7, dtype=torch.bool), mode="iou")
    assert P.FCOSLoss().mode == "giou" and P.FCOSLoss("iou").mode == "iou"


def test_shard_bounds_partition_the_batch():
    for batch in (1, 2, 7, 16, 255, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    x = W.head_outputs(batch, 5, W.VOC_LEVELS[2:], seed=9)            # small: 16x16 .. 4x4 levels
    strides = W.STRIDES[2:]
    mine = sharding.shard_levels(x, world, rank)
    k = 50
    dets = O.detect(mine, 0.05, 0.6, k, strides)                       # the CPU oracle stands in for the kernels
    nb = len(dets)
    scores = torch.zeros(nb, k)
    classes = torch.zeros(nb, k, dtype=torch.int64)
    boxes = torch.zeros(nb, k, 4)
    counts = torch.zeros(nb, dtype=torch.int32)
    for i, (s, c, b) in enumerate(dets):
        n = s.numel()
        scores[i, :n], classes[i, :n], boxes[i, :n], counts[i] = s, c, b, n
    full = sharding.gather_detections(scores, classes, boxes, counts, batch)
    per_image = [torch.arange(nb, dtype=torch.float32) + 10 * rank + 1]
    red = sharding.reduce_image_losses(per_image, batch)
    packed_all = None
    if batch % world == 0:                                            # single-collective packed gather
        from pytorch_object_detection_b200 import ops
        pk = ops.packed_detections(nb, k, "cpu")
        pk.zero_()
        for dst, src in zip(ops.detection_views(pk, nb, k), (scores, classes, boxes, classes, counts)):
            dst.copy_(src)
        assert sharding.packed_of(ops.detection_views(pk, nb, k)[0]).data_ptr() == pk.data_ptr()
        packed_all = sharding.gather_packed(pk)
    if rank == 0:
        torch.save({"full": full, "red": red, "packed": packed_all}, os.path.join(out_dir, "out.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [5, 6])
def test_two_rank_gather_matches_single_process(tmp_path, batch):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), batch, str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(str(tmp_path), "out.pt"))
    x = W.head_outputs(batch, 5, W.VOC_LEVELS[2:], seed=9)
    want = O.detect(x, 0.05, 0.6, 50, W.STRIDES[2:])
    scores, classes, boxes, counts = got["full"]
    assert scores.shape[0] == batch and counts.shape[0] == batch
    for i, (s, c, b) in enumerate(want):
        n = int(counts[i])
        assert n == s.numel()
        assert torch.equal(scores[i, :n], s) and torch.equal(classes[i, :n], c) and torch.equal(boxes[i, :n], b)
    if got["packed"] is not None:
        from pytorch_object_detection_b200 import ops
        nb = batch // world
        for r in range(world):
            ps, pc, pb, _, pn = ops.detection_views(got["packed"][r], nb, 50)
            for i in range(nb):
                n = int(pn[i])
                s, c, b = want[r * nb + i]
                assert n == s.numel() and torch.equal(ps[i, :n], s) and torch.equal(pc[i, :n], c)
                assert torch.equal(pb[i, :n], b)
    # per-image "losses": rank r holds arange(nb) + 10 r + 1
    sizes = [sharding.shard_bounds(batch, world, r) for r in range(world)]
    total = sum(float((torch.arange(hi - lo, dtype=torch.float32) + 10 * r + 1).sum()) for r, (lo, hi) in enumerate(sizes))
    assert float(got["red"][0]) == pytest.approx(total / batch, rel=1e-6)


def test_sharded_oracle_equals_full_batch():
    """Images are independent: running shards separately gives the same per-image result."""
    x = W.head_outputs(3, 20, W.VOC_LEVELS, seed=21)
    full = O.detect(x, 0.05, 0.6, 1000, W.STRIDES)
    for world in (2, 3):
        got = []
        for r in range(world):
            got += O.detect(sharding.shard_levels(x, world, r), 0.05, 0.6, 1000, W.STRIDES)
        for a, b in zip(got, full):
            assert all(torch.equal(p, q) for p, q in zip(a, b))


def test_every_dispatcher_op_has_a_cuda_kernel_and_a_fake_shape_function():
    """The b200det torch.library namespace: CUDA-only kernels — a CPU tensor finds none — and fake
    implementations that give the output shapes / dtypes without running anything (no GPU needed: FakeTensorMode
    fabricates the CUDA tensors)."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from pytorch_object_detection_b200 import ops
    assert len(ops.OP_NAMES) >= 15 and {"cls_loss_step", "assign_loss_fused", "rescale_maps_"} <= set(ops.OP_NAMES)
    with pytest.raises(NotImplementedError):
        torch.ops.b200det.score_points([torch.zeros(1, 2, 4, 4)], [torch.zeros(1, 1, 4, 4)], [8])
    with pytest.raises(NotImplementedError):
        torch.ops.b200det.box_loss_fwd([torch.zeros(1, 4, 4, 4)], torch.zeros(1, 16), torch.zeros(1, 16, 4), 1)
    p = W.num_points(W.VOC_LEVELS)
    with FakeTensorMode():
        dev = "cuda"
        x = [[torch.empty(2, c, h, w, device=dev) for h, w in W.VOC_LEVELS] for c in (20, 1, 4)]
        s, c, b, k, n = torch.ops.b200det.postprocess(x[0], x[1], x[2], W.STRIDES[:4], 0.05, 0.6, 300, 0, 0)
        assert s.shape == (2, 300) and c.dtype == torch.int64 and b.shape == (2, 300, 4) and n.dtype == torch.int32
        sc, c0 = torch.ops.b200det.score_points(x[0], x[1], W.STRIDES)
        assert sc.shape == (2, p) and c0.dtype == torch.int16
        top = torch.ops.b200det.select_topk(x[2], W.STRIDES, sc, c0, 0.05, 100000)
        assert top[0].shape == (2, p) and top[2].shape == (2, p, 4)
        gt = torch.empty(2, 7, 4, device=dev)
        lab = torch.empty(2, 7, dtype=torch.int64, device=dev)
        t = torch.ops.b200det.assign_targets([v for hw in W.VOC_LEVELS for v in hw], W.STRIDES, [-1.0] * 5, [64.0] * 5, gt, lab, 1.5)
        assert t[0].shape == (2, p, 1) and t[0].dtype == torch.int64 and t[2].shape == (2, p, 4)
        loss, npos = torch.ops.b200det.box_loss_fwd(x[2], t[1], t[2], 1)
        grads = torch.ops.b200det.box_loss_bwd(x[2], t[1], t[2], 1, loss, npos)
        assert loss.shape == (2,) and [g.shape for g in grads] == [m.shape for m in x[2]]
        half = [m.half() for m in x[0]]
        l2, mean, np2, g2 = torch.ops.b200det.cls_loss_step(half, t[0], t[1], None, None, None)
        assert mean.shape == (2,) and g2[0].dtype == torch.float16 and g2[4].shape == half[4].shape
        r = torch.ops.b200det.assign_loss_fused(x[2], None, W.STRIDES, [-1.0] * 5, [64.0] * 5, gt, lab, 1, 1.5, None, None, None)
        assert r[0].shape == (2, p, 1) and r[4].numel() == 0 and len(r[7]) == 5 and r[8] == [] and r[6].shape == (4,)
        r = torch.ops.b200det.assign_loss_fused(x[2], x[1], W.STRIDES, [-1.0] * 5, [64.0] * 5, gt, lab, 0, 1.5, None, None,
                                                [torch.empty(1, device=dev)] * 5)
        assert r[4].shape == (2,) and len(r[8]) == 5 and r[9].shape == (5,)


def test_iou_threshold_midpoint_test_equals_the_rounded_division():
    """csrc/nms_body.cuh iou_reaches(): `fl32(inter / uni) >= thr_up` (torchvision's nms_kernel_impl expression as the
    kernels used to evaluate it) is the same predicate as `double(inter) > mid * double(uni)` with mid the midpoint of
    thr_up and the float below it.  Checked on random pairs and on pairs built to sit within a few ulps of the
    threshold, for several thresholds including the smallest one (nms_thr = 0)."""
    rng = np.random.default_rng(7)
    for thr in (0.6, 0.5, 0.3, 0.05, 0.999, 1.0 / 3.0, 0.0, 1.5):
        f = np.float32(thr)
        if not float(f) > thr:
            f = np.nextafter(f, np.float32(np.inf), dtype=np.float32)
        below = (f.view(np.int32) - np.int32(1)).view(np.float32)
        mid = 0.5 * (np.float64(below) + np.float64(f))
        uni = np.exp(rng.uniform(-3, 12, 400000)).astype(np.float32)
        near = (uni.astype(np.float64) * float(f)).astype(np.float32)
        steps = rng.integers(-3, 4, near.shape)
        for _ in range(3):                                  # walk a few ulps either side of thr_up * uni
            near = np.where(steps > 0, np.nextafter(near, np.float32(np.inf)), np.where(steps < 0, np.nextafter(near, np.float32(-np.inf)), near)).astype(np.float32)
            steps = steps - np.sign(steps)
        inter = np.concatenate([near, (uni * rng.uniform(0, 1, uni.shape)).astype(np.float32)])
        uni2 = np.concatenate([uni, uni])
        special_i = np.array([0, 0, 1, np.inf, np.inf, np.nan, 1, 0, 1e-45], np.float32)
        special_u = np.array([0, 1, 0, np.inf, 1, 1, np.nan, np.inf, 1e-45], np.float32)
        inter = np.concatenate([inter, special_i])
        uni2 = np.concatenate([uni2, special_u])
        with np.errstate(all="ignore"):
            ref = (inter / uni2).astype(np.float32) >= f
            got = inter.astype(np.float64) > mid * uni2.astype(np.float64)
        # 1 / 0 = inf reaches every threshold in the division form; a positive intersection with an empty union cannot
        # come out of two overlapping boxes (uni >= the larger area >= inter > 0), so that pair is left out
        keep = ~((uni2 == 0) & (inter > 0))
        assert np.array_equal(ref[keep], got[keep]), thr
        assert ref[: near.size].any() and not ref[: near.size].all()
