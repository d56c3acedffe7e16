"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden
fixtures produced by the unmodified reference.  Run on the B200 box with ``-m gpu``.

Tolerances (BASELINE.json north_star): integers (keep indices, labels, GT indices) bit-exact on
identical stage inputs; boxes / scores / losses within 1e-5 relative.
"""
import numpy as np
import pytest
import torch

from oracle import fcos_oracle as O
from pytorch_object_detection_b200 import workloads as W
from helpers import (EVAL_CASES, REL_TOL, assert_ap_equal, assert_close, assert_detections_match, assert_equal_int,
                     head_inputs_from_meta, load_eval_case, load_golden, to_np)

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import pytorch_object_detection_b200 as P
    from pytorch_object_detection_b200 import _lib, ops
DEV = "cuda:0"


def cuda_levels(x):
    return [[t.to(DEV) for t in part] for part in x]


HEAD_CASES = {
    "head_voc_b1": (W.VOC_LEVELS, W.VOC_HW),
    "head_voc_4strides": (W.VOC_LEVELS, W.VOC_HW),
    "head_coco_b2": (W.COCO_LEVELS, W.COCO_HW),
    "head_coco_crowded": (W.COCO_LEVELS, W.COCO_HW),
    "head_voc_k300": (W.VOC_LEVELS, W.VOC_HW),
    "head_voc_saturated": (W.VOC_LEVELS, W.VOC_HW),      # class logits that collapse in the fp32 sigmoid
}


def head_case(name):
    levels, img_hw = HEAD_CASES[name]
    g = load_golden(name)
    x, batch, max_box, strides = head_inputs_from_meta(g, levels)
    return g, x, batch, max_box, strides, img_hw


# ------------------------------------------------------------------------------------------
# K1: score / argmax
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["head_voc_b1", "head_coco_b2", "head_voc_4strides", "head_voc_saturated"])
def test_score_points_matches_oracle(name):
    g, x, batch, max_box, strides, _ = head_case(name)
    want_s, want_c, _ = O.score_points(x, strides)
    xc = cuda_levels(x)
    got_s, got_c = ops.score_points(xc[0], xc[1], strides)
    assert got_s.shape == want_s.shape
    assert_close(to_np(got_s), to_np(want_s), REL_TOL, what="score")
    # the class is torch.max's FIRST index among equal fp32 sigmoid values (head.py:57-62): on the saturated case
    # more than half of the points would get another class from an argmax over the logits
    got = to_np(got_c).astype(np.int64) + 1
    want = to_np(want_c)
    if name == "head_voc_saturated":
        assert int(g["points_where_logit_argmax_differs"]) > 1000
        # torch's CPU sigmoid (vectorised exp) and expf can round a logit next to a collapse boundary differently:
        # such a point may legitimately take the neighbouring class of the SAME sigmoid value up to 1 ulp
        bad = np.nonzero(got != want)
        cls_flat = torch.cat([t.permute(0, 2, 3, 1).reshape(batch, -1, t.shape[1]) for t in x[0]], dim=1)
        sig = torch.sigmoid(cls_flat).numpy()
        for b_i, p_i in zip(*bad):
            a, w_ = sig[b_i, p_i, got[b_i, p_i] - 1], sig[b_i, p_i, want[b_i, p_i] - 1]
            assert abs(float(a) - float(w_)) <= 1.2e-7 * float(w_), "class differs beyond a sigmoid rounding boundary"
        assert bad[0].size <= 0.002 * got.size, f"{bad[0].size} class mismatches"
    else:
        assert_equal_int(got, want, what="class")


# ------------------------------------------------------------------------------------------
# K2: top-k on IDENTICAL scores (the oracle's), tie-aware order, boxes bit-exact
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["head_voc_b1", "head_coco_b2", "head_voc_k300"])
def test_select_topk_on_identical_scores(name):
    g, x, batch, max_box, strides, _ = head_case(name)
    score, classes, boxes = O.score_points(x, strides)
    s_k, c_k, b_k, idx = O.select_topk(score, classes, boxes, max_box)
    xc = cuda_levels(x)
    gs, gc, gb, gp, gn = ops.select_topk(xc[2], strides, score.to(DEV).contiguous(),
                                         (classes - 1).to(torch.int16).to(DEV).contiguous(), 0.05, max_box)
    for b in range(batch):
        m = s_k[b] >= 0.05
        n = int(m.sum())
        assert int(gn[b]) == n
        assert_detections_match((to_np(gs[b, :n]), to_np(gc[b, :n]), to_np(gb[b, :n])),
                                (to_np(s_k[b][m]), to_np(c_k[b][m]), to_np(b_k[b][m])), rel=0.0,
                                what=f"{name} top-k img {b}")
        # our tie rule: score desc, point index asc
        pts = to_np(gp[b, :n]).astype(np.int64)
        sc = to_np(gs[b, :n]).astype(np.float64)
        order = np.lexsort((pts, -sc))
        assert np.array_equal(order, np.arange(n))
        assert np.array_equal(np.sort(pts), np.sort(to_np(idx[b][m])))


def test_select_topk_all_equal_scores_takes_lowest_indices():
    levels = [(8, 8), (4, 4)]
    strides = [8, 16]
    p = W.num_points(levels)
    reg = [torch.rand(2, 4, h, w, device=DEV) * 20 + 1 for h, w in levels]
    score = torch.full((2, p), 0.25, device=DEV)
    score[1, 10:30] = 0.5
    cls0 = torch.zeros((2, p), dtype=torch.int16, device=DEV)
    gs, gc, gb, gp, gn = ops.select_topk(reg, strides, score, cls0, 0.05, 16)
    assert gn.tolist() == [16, 16]
    assert gp[0].tolist() == list(range(16))
    assert gp[1].tolist() == list(range(10, 26))
    gs, gc, gb, gp, gn = ops.select_topk(reg, strides, score, cls0, 0.3, 50)
    assert gn.tolist() == [0, 20]
    assert gp[1, :20].tolist() == list(range(10, 30))
    # k larger than P: everything above the threshold, sorted
    gs, gc, gb, gp, gn = ops.select_topk(reg, strides, score, cls0, 0.05, 1000)
    assert gn.tolist() == [p, p]
    assert gp[1, :20].tolist() == list(range(10, 30))


# ------------------------------------------------------------------------------------------
# K3: NMS keep indices bit-exact against torchvision's CPU op (golden) on identical inputs
# ------------------------------------------------------------------------------------------
def run_nms(boxes, scores, classes, thr, score_thr=-1e30, clip=None):
    s, c, b, k, n = ops.batched_nms(boxes[None].to(DEV), scores[None].to(DEV), classes[None].to(DEV), score_thr, thr,
                                    None, clip)
    n = int(n[0])
    return to_np(s[0, :n]), to_np(c[0, :n]), to_np(b[0, :n]), to_np(k[0, :n])


def test_nms_matches_torchvision_golden():
    g = load_golden("nms_cases")
    names = sorted({k.rsplit("_", 1)[0] for k in g.files if k.endswith("_keep")})
    for nme in names:
        boxes = torch.from_numpy(g[nme + "_boxes"])
        scores = torch.from_numpy(g[nme + "_scores"])
        classes = torch.from_numpy(g[nme + "_classes"]).long()
        s, c, b, keep = run_nms(boxes, scores, classes, float(g[nme + "_thr"]))
        want = g[nme + "_keep"].astype(np.int64)
        if nme in ("crowd5000", "crowd1001"):   # vanilla branch: torch's final sort is unstable on ties
            assert np.array_equal(np.sort(keep), np.sort(want)), nme
            assert_close(scores.numpy()[keep], scores.numpy()[want], rel=0.0, what=nme)
        else:
            assert_equal_int(keep, want, what=nme)
        assert np.array_equal(s, scores.numpy()[keep])
        assert np.array_equal(c, classes.numpy()[keep])
        assert np.array_equal(b, boxes.numpy()[keep])


@pytest.mark.parametrize("seed,n,ncls", [(1, 300, 4), (2, 999, 80), (3, 1000, 20), (4, 1500, 10), (5, 64, 2),
                                         (6, 65, 2), (7, 2048, 3), (8, 5000, 80),
                                         # per-class path: one class larger than a warp handles (dense fallback),
                                         # a single class just above / below that limit, many tiny classes
                                         (9, 3000, 2), (10, 1200, 1), (11, 1001, 1), (12, 4100, 1000), (13, 8192, 80)])
def test_nms_matches_oracle_random(seed, n, ncls):
    boxes, scores, classes = W.crowd_candidates(n, ncls, seed=seed, clusters=8)
    boxes[::7] -= 600.0          # negative coordinates: cross-class suppression on the trick branch
    want = O.batched_nms(boxes, scores, classes, 0.6).numpy()
    _, _, _, keep = run_nms(boxes, scores, classes, 0.6)
    assert_equal_int(keep, want, what=f"n={n}")


@pytest.mark.parametrize("case", ["two_batches_dense_ids", "three_batches_many_tiny_classes", "zero_iou_suppresses",
                                  "threshold_zero", "ids_256_to_16383", "ids_above_the_histogram", "id_255_and_256",
                                  "sizes_63_to_66_and_129", "improper_boxes"])
def test_nms_per_class_kernel_branches(case):
    """nms_class.cu: every way through the per-class kernel (> 1000 candidates) against the oracle — the table-driven
    class split (ids < 256) and the counting / bitonic sorts behind it, one and several tile batches per CTA, ordered
    and round-robin dealing, the ZERO_SUP pair test, the smallest threshold, class sizes around the 64-box block."""
    g = torch.Generator().manual_seed(77)
    thr = 0.6
    if case == "two_batches_dense_ids":            # 9 classes x 900: 120 tiles each, the CTA with two of them needs 2 batches
        boxes, scores, _ = W.crowd_candidates(8100, 9, seed=21, clusters=12)
        classes = torch.arange(8100) % 9 + 1
    elif case == "three_batches_many_tiny_classes":   # ~2 900 classes: > 136 one-tile classes per CTA
        boxes, scores, _ = W.crowd_candidates(8192, 9, seed=22, clusters=6)
        classes = torch.randint(5000, 8000, (8192,), generator=g)
    elif case == "zero_iou_suppresses":
        boxes, scores, classes = W.crowd_candidates(3000, 40, seed=23, clusters=30)
        thr = -0.5
    elif case == "threshold_zero":
        boxes, scores, classes = W.crowd_candidates(3000, 40, seed=24, clusters=300, spread=60.0)
        thr = 0.0
    elif case == "ids_256_to_16383":
        boxes, scores, classes = W.crowd_candidates(4000, 60, seed=25, clusters=20)
        classes = classes * 250 + 300
    elif case == "ids_above_the_histogram":
        boxes, scores, classes = W.crowd_candidates(2500, 30, seed=26, clusters=20)
        classes = classes * 3000 + 20000
    elif case == "id_255_and_256":
        boxes, scores, classes = W.crowd_candidates(2400, 3, seed=27, clusters=10)
        classes = classes + 253
    elif case == "sizes_63_to_66_and_129":
        sizes = [63, 64, 65, 66, 128, 129, 1, 2, 700]
        classes = torch.cat([torch.full((m,), i + 1) for i, m in enumerate(sizes)])
        classes = classes[torch.randperm(classes.numel(), generator=g)]
        boxes, scores, _ = W.crowd_candidates(classes.numel(), 3, seed=28, clusters=3)
    else:                                          # inverted and empty boxes never suppress and are never suppressed
        boxes, scores, classes = W.crowd_candidates(2000, 6, seed=29, clusters=6)
        boxes[::5] = boxes[::5][:, [2, 3, 0, 1]]
        boxes[1::11, 2:] = boxes[1::11, :2]
    want = O.batched_nms(boxes, scores, classes, thr).numpy()
    _, _, _, keep = run_nms(boxes, scores, classes.long(), thr)
    assert_equal_int(keep, want, what=case)
    assert 0 < keep.size < boxes.shape[0] or case == "zero_iou_suppresses"


def test_nms_per_class_kernel_pairs_at_the_threshold():
    """The per-class kernel decides most pairs from two rounded products and leaves a band of one or two ulps around
    the threshold to the exact test (csrc/nms_body.cuh mask_rows2_part).  Here hundreds of same-class pairs sit IN that
    band: a box and a box it contains, whose IoU is h / side with h walked a few ulps around 0.6 * side, on three
    scales; the kept set must be the oracle's (torchvision's fp32 quotient compared with the double threshold)."""
    g = torch.Generator().manual_seed(5)
    boxes, scores, classes = [], [], []
    k = 0
    for side in (10.0, 20.0, 37.0):
        for step in range(-4, 5):
            for rep in range(12):
                h = np.float32(0.6) * np.float32(side)
                for _ in range(abs(step)):
                    h = np.nextafter(h, np.float32(np.inf if step > 0 else -np.inf), dtype=np.float32)
                x0, y0 = 64.0 * (k % 40), 64.0 * (k // 40)
                boxes += [[x0, y0, x0 + side, y0 + side], [x0, y0, x0 + side, y0 + float(h)]]
                scores += [0.9 - 1e-4 * k, 0.5 - 1e-4 * k]
                classes += [1 + k % 3, 1 + k % 3]
                k += 1
    fb, fs, fc = W.crowd_candidates(900, 3, seed=31, clusters=40)
    fb[:, [0, 2]] += 3000.0                                       # the crowd sits beside the constructed pairs
    boxes = torch.cat([torch.tensor(boxes, dtype=torch.float32), fb])
    scores = torch.cat([torch.tensor(scores, dtype=torch.float32), fs * 0.4])
    classes = torch.cat([torch.tensor(classes), fc])
    perm = torch.randperm(boxes.shape[0], generator=g)
    boxes, scores, classes = boxes[perm], scores[perm], classes[perm]
    assert boxes.shape[0] > 1000                                   # per-class branch
    want = O.batched_nms(boxes, scores, classes, 0.6).numpy()
    _, _, _, keep = run_nms(boxes, scores, classes.long(), 0.6)
    assert_equal_int(keep, want, what="pairs at the threshold")
    inside = np.isin(np.nonzero((perm < 2 * k).numpy())[0], want)  # both outcomes occur among the constructed boxes
    assert inside.sum() > k and inside.sum() < 2 * k


@pytest.mark.parametrize("n,ncls,clusters,thr", [(1000, 80, 50, 0.6), (1024, 3, 4, 0.5), (700, 1, 2, 0.3), (64, 2, 1, 0.6),
                                                 (1, 1, 1, 0.6), (999, 20, 300, 0.0)])
def test_standalone_nms_bucket_path_equals_dense_chain(n, ncls, clusters, thr, monkeypatch):
    """b200det_batched_nms with <= 1024 candidates runs the NMS half of the fused head kernel on the prepared set;
    B200DET_DENSE_NMS=1 forces the dense mask + scan kernels.  Same outputs, and the oracle's keep indices; boxes with
    negative corners (cross-class suppression under the coordinate trick) included."""
    boxes, scores, classes = zip(*[W.crowd_candidates(n, ncls, seed=700 + i, clusters=clusters) for i in range(3)])
    boxes, scores, classes = torch.stack(boxes), torch.stack(scores), torch.stack(classes)
    boxes[:, ::9] -= 500.0
    in_count = torch.tensor([n, max(n // 2, 1), n], dtype=torch.int32)
    args = (boxes.to(DEV), scores.to(DEV), classes.to(DEV), 0.1, thr, in_count.to(DEV), (832, 1344))
    got = ops.batched_nms(*args)
    monkeypatch.setenv("B200DET_DENSE_NMS", "1")
    ref = ops.batched_nms(*args)
    assert torch.equal(got[4], ref[4])
    for i in range(3):
        m = int(ref[4][i])
        for a, b in zip(got[:4], ref[:4]):
            assert torch.equal(a[i, :m], b[i, :m])
        cnt = int(in_count[i])
        want = O.post_process_image(scores[i, :cnt], classes[i, :cnt], boxes[i, :cnt], 0.1, thr)
        assert_equal_int(to_np(got[3][i, :m]), to_np(want[3]), what=f"keep img {i}")


def test_nms_threshold_ragged_batch_and_clip():
    b0, s0, c0 = W.crowd_candidates(500, 5, seed=41, clusters=5)
    b1, s1, c1 = W.crowd_candidates(500, 5, seed=42, clusters=5)
    boxes, scores, classes = torch.stack([b0, b1]), torch.stack([s0, s1]), torch.stack([c0, c1])
    in_count = torch.tensor([500, 123], dtype=torch.int32)
    s, c, b, k, n = ops.batched_nms(boxes.to(DEV), scores.to(DEV), classes.to(DEV), 0.4, 0.5, in_count.to(DEV),
                                    (832, 1344))
    for i, cnt in enumerate([500, 123]):
        want = O.post_process_image(scores[i, :cnt], classes[i, :cnt], boxes[i, :cnt], 0.4, 0.5)
        m = int(n[i])
        assert m == want[0].numel()
        assert_equal_int(to_np(k[i, :m]), to_np(want[3]), what="keep")
        assert np.array_equal(to_np(s[i, :m]), to_np(want[0]))
        clipped = O.clip_boxes_(want[2].clone(), 832, 1344)
        assert np.array_equal(to_np(b[i, :m]), to_np(clipped))
    # nothing above the threshold
    s, c, b, k, n = ops.batched_nms(boxes.to(DEV), scores.to(DEV), classes.to(DEV), 2.0, 0.5)
    assert n.tolist() == [0, 0]


# ------------------------------------------------------------------------------------------
# whole head against the reference's outputs (golden)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(HEAD_CASES))
def test_head_matches_reference_golden(name):
    g, x, batch, max_box, strides, img_hw = head_case(name)
    head = P.FCOSHead(0.05, 0.6, max_box, strides)
    xc = cuda_levels(x)
    s, c, b, n = head.detect(xc)
    s2, c2, b2, n2 = head.detect(xc, clip_hw=img_hw)
    for i in range(batch):
        m = int(n[i])
        assert_detections_match((to_np(s[i, :m]), to_np(c[i, :m]), to_np(b[i, :m])),
                                (g[f"score_{i}"], g[f"class_{i}"], g[f"box_{i}"]), rel=REL_TOL, what=f"{name} img {i}")
        assert int(n2[i]) == m
        assert_detections_match((to_np(s2[i, :m]), to_np(c2[i, :m]), to_np(b2[i, :m])),
                                (g[f"score_{i}"], g[f"class_{i}"], g[f"clipped_{i}"]), rel=REL_TOL,
                                what=f"{name} clipped img {i}")
    if batch == 1:     # the drop-in forward + ClipBoxes, exactly as test.py:206-207 calls them
        fs, fc, fb = head(xc)
        assert fs.shape == (1, int(n[0])) and fc.dtype == torch.int64 and fb.shape == (1, int(n[0]), 4)
        imgs = torch.zeros(1, 3, *img_hw, device=DEV)
        out = P.ClipBoxes()(imgs, fb)
        assert out.data_ptr() == fb.data_ptr()
        assert_detections_match((to_np(fs[0]), to_np(fc[0]), to_np(fb[0])),
                                (g["score_0"], g["class_0"], g["clipped_0"]), rel=REL_TOL, what=name + " forward")


def test_head_ragged_batch_raises_like_reference_and_empty_result():
    x = cuda_levels(W.head_outputs(2, 20, W.VOC_LEVELS, seed=77))
    x[0][0][1] -= 3.0                          # image 1 keeps fewer boxes
    head = P.FCOSHead(0.05, 0.6, 1000, W.STRIDES)
    s, c, b, n = head.detect(x)
    assert n[0] != n[1]
    with pytest.raises(RuntimeError, match="stack expects each tensor to be equal size"):
        head(x)
    x1 = [[t[:1] for t in part] for part in x]
    fs, fc, fb = P.FCOSHead(0.99, 0.6, 1000, W.STRIDES)(x1)
    assert fs.shape == (1, 0) and fc.shape == (1, 0) and fc.dtype == torch.int64 and fb.shape == (1, 0, 4)
    with pytest.raises(Exception):
        head([[t.cpu() for t in part] for part in x1])     # no CPU path


def test_post_process_entry_matches_oracle():
    g, x, batch, max_box, strides, _ = head_case("head_voc_b1")
    score, classes, boxes = O.score_points(x, strides)
    s_k, c_k, b_k, _ = O.select_topk(score, classes, boxes, max_box)
    want = O.post_process_image(s_k[0], c_k[0], b_k[0], 0.05, 0.6)
    got = P.FCOSHead(0.05, 0.6, max_box, strides).post_process([s_k.to(DEV), c_k.to(DEV), b_k.to(DEV)])
    assert np.array_equal(to_np(got[0][0]), to_np(want[0]))
    assert np.array_equal(to_np(got[1][0]), to_np(want[1]))
    assert np.array_equal(to_np(got[2][0]), to_np(want[2]))


# ------------------------------------------------------------------------------------------
# K4a: target assignment, bit-exact against the reference (golden)
# ------------------------------------------------------------------------------------------
TRAIN_CASES = {
    "train_voc_b2": (W.VOC_LEVELS, W.VOC_HW),
    "train_coco_b2": (W.COCO_LEVELS, W.COCO_HW),
    "train_voc_dense": (W.VOC_LEVELS, W.VOC_HW),
}


def ieee_centerness(reg_t, cnt_ref):
    """Centerness recomputed from reg targets with numpy fp32 (every op correctly rounded, as CUDA's
    __fsqrt_rn/__fdiv_rn are).  torch's CPU sqrt goes through MKL VML and is off by 1 ulp on ~0.6 %
    of inputs, so the reference's CPU cnt_t is matched to 1 ulp and this IEEE value bit-exactly."""
    r = np.asarray(reg_t, dtype=np.float32)
    f = np.float32
    lr_min, lr_max = np.minimum(r[..., 0], r[..., 2]), np.maximum(r[..., 0], r[..., 2])
    tb_min, tb_max = np.minimum(r[..., 1], r[..., 3]), np.maximum(r[..., 1], r[..., 3])
    with np.errstate(invalid="ignore"):
        c = np.sqrt((lr_min * tb_min) / (lr_max * tb_max + f(1e-10)), dtype=np.float32)
    return np.where(np.asarray(cnt_ref)[..., 0] > -1, c, f(-1))[..., None]


def assert_cnt_matches(got, ref_cnt, ref_reg):
    got = to_np(got)
    assert np.array_equal(got > -1, np.asarray(ref_cnt) > -1), "positive sets differ"
    assert np.array_equal(got, ieee_centerness(ref_reg, ref_cnt)), "cnt_t differs from the IEEE evaluation"
    assert_close(got, ref_cnt, rel=2e-7, what="cnt_t vs reference CPU (1 ulp: MKL sqrt)")


def train_case(name):
    levels, img_hw = TRAIN_CASES[name]
    g = load_golden(name)
    batch, ncls, seed, max_gt = (int(v) for v in g["meta"][:4])
    gt, labels = W.gt_boxes(batch, max_gt, img_hw, ncls, seed)
    x = W.head_outputs(batch, ncls, levels, seed + 1)
    return g, x, gt, labels, g["ranges"].tolist(), levels


@pytest.mark.parametrize("name", sorted(TRAIN_CASES))
def test_assign_targets_bit_exact(name):
    g, x, gt, labels, ranges, levels = train_case(name)
    gen = P.FCOSGenTargets(W.STRIDES, ranges)
    cls_t, cnt_t, reg_t = gen([cuda_levels(x), gt.to(DEV), labels.to(DEV)])
    assert cls_t.dtype == torch.int64 and cls_t.shape == (gt.shape[0], W.num_points(levels), 1)
    assert_equal_int(to_np(cls_t), g["cls_t"], what="cls_t")
    assert np.array_equal(to_np(reg_t), g["reg_t"]), "reg_t not bit-exact"
    assert_cnt_matches(cnt_t, g["cnt_t"], g["reg_t"])
    # GT index against the oracle
    _, _, _, want_idx = O.assign_targets(levels, gt, labels, W.STRIDES, ranges)
    out = ops.assign_targets(levels, W.STRIDES, ranges, gt.to(DEV), labels.to(DEV), want_index=True)
    assert_equal_int(to_np(out[3]), to_np(want_idx), what="gt index")
    # single-level static entry point (head.py:235)
    lv = 1
    one = P.FCOSGenTargets.generate_target([t[lv].to(DEV) for t in x], gt.to(DEV), labels.to(DEV), W.STRIDES[lv],
                                           ranges[lv])
    o0 = sum(h * w for h, w in levels[:lv])
    o1 = o0 + levels[lv][0] * levels[lv][1]
    assert np.array_equal(to_np(one[0]), g["cls_t"][:, o0:o1].astype(np.int64))
    assert np.array_equal(to_np(one[2]), g["reg_t"][:, o0:o1])


def test_assign_targets_full_size_config3():
    """BASELINE config 3 at full size (B=32, P=23265, M<=100) against the oracle."""
    gt, labels = W.gt_boxes(32, 100, W.COCO_HW, 80, seed=301)
    want = O.assign_targets(W.COCO_LEVELS, gt, labels, W.STRIDES, W.HISFCOS_RANGES)
    got = ops.assign_targets(W.COCO_LEVELS, W.STRIDES, W.HISFCOS_RANGES, gt.to(DEV), labels.to(DEV), want_index=True)
    assert_equal_int(to_np(got[0]), to_np(want[0]), what="cls_t")
    assert np.array_equal(to_np(got[2]), to_np(want[2]))
    assert_cnt_matches(got[1], to_np(want[1]), to_np(want[2]))
    assert_equal_int(to_np(got[3]), to_np(want[3]), what="gt index")
    assert int((got[1] > -1).sum()) > 1000


def test_assign_targets_edge_cases():
    # no ground truth at all (all padding), one GT, and GT ties (identical boxes -> lowest index wins)
    gt = torch.full((3, 4, 4), -1.0)
    labels = torch.full((3, 4), -1, dtype=torch.int64)
    gt[1, 0] = torch.tensor([100.0, 120.0, 300.0, 260.0]); labels[1, 0] = 7
    gt[2, 1] = torch.tensor([50.0, 50.0, 250.0, 250.0]); labels[2, 1] = 3
    gt[2, 3] = torch.tensor([50.0, 50.0, 250.0, 250.0]); labels[2, 3] = 9
    want = O.assign_targets(W.VOC_LEVELS, gt, labels, W.STRIDES, W.FCOS_RANGES)
    got = ops.assign_targets(W.VOC_LEVELS, W.STRIDES, W.FCOS_RANGES, gt.to(DEV), labels.to(DEV), want_index=True)
    for j in (0, 2, 3):
        assert np.array_equal(to_np(got[j]).astype(np.float64), to_np(want[j]).astype(np.float64))
    assert_cnt_matches(got[1], to_np(want[1]), to_np(want[2]))
    assert int((got[0][0] != 0).sum()) == 0
    assert set(to_np(got[3][2]).tolist()) <= {-1, 1}


# ------------------------------------------------------------------------------------------
# K4b: losses and gradients against the reference (golden)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(TRAIN_CASES))
@pytest.mark.parametrize("mode", ["giou", "iou"])
def test_losses_and_gradients_match_reference(name, mode):
    g, x, gt, labels, ranges, levels = train_case(name)
    xc = cuda_levels(x)
    for part in xc:
        for t in part:
            t.requires_grad_(True)
    tgt = P.FCOSGenTargets(W.STRIDES, ranges)([xc, gt.to(DEV), labels.to(DEV)])
    losses = P.FCOSLoss(mode)([xc, tgt])
    assert all(v.dim() == 0 for v in losses)
    assert_close([float(v) for v in losses], g[f"loss_{mode}"], rel=REL_TOL, what=f"loss {mode}")
    losses[3].backward()
    for lv in range(len(levels)):
        assert_close(to_np(xc[2][lv].grad), g[f"g_reg_{mode}_{lv}"], rel=REL_TOL, abs_=1e-9, what=f"g_reg {lv}")
        assert_close(to_np(xc[1][lv].grad), g[f"g_cnt_{mode}_{lv}"], rel=REL_TOL, abs_=1e-9, what=f"g_cnt {lv}")
        if f"g_cls_{mode}_{lv}" in g.files:
            assert_close(to_np(xc[0][lv].grad), g[f"g_cls_{mode}_{lv}"], rel=REL_TOL, abs_=1e-12, what=f"g_cls {lv}")
        else:
            assert_close(to_np(xc[0][lv].grad.double().sum(dim=(2, 3))), g[f"g_cls_sum_{mode}_{lv}"], rel=1e-4,
                         abs_=1e-9, what=f"g_cls sums {lv}")


def test_known_answer_and_free_functions():
    """model/loss.py:219-221 prints tensor([0.3133, 0.3133]); analytic IoU/GIoU values."""
    got = P.compute_cnt_loss([torch.ones(2, 1, 4, 4, device=DEV)] * 5, torch.ones(2, 80, 1, device=DEV),
                             torch.ones(2, 80, dtype=torch.bool, device=DEV))
    assert_close(to_np(got), load_golden("known_answers")["cnt_loss_ones"], rel=1e-6)
    assert [round(float(v), 4) for v in got] == [0.3133, 0.3133]
    same = torch.tensor([[3.0, 4.0, 5.0, 6.0]], device=DEV)
    assert float(P.giou_loss(same, same)) == pytest.approx(0.0, abs=1e-6)
    inner = torch.tensor([[5.0, 5.0, 5.0, 5.0]], device=DEV, requires_grad=True)
    outer = torch.tensor([[10.0, 10.0, 10.0, 10.0]], device=DEV)
    gl = P.giou_loss(inner, outer)
    assert float(gl) == pytest.approx(0.75, rel=1e-6)
    assert float(P.iou_loss(inner, outer)) == pytest.approx(-np.log(0.25), rel=1e-6)
    gl.backward()
    ref_in = torch.tensor([[5.0, 5.0, 5.0, 5.0]], requires_grad=True)
    O.giou_sum(ref_in, outer.cpu()).backward()
    assert_close(to_np(inner.grad), to_np(ref_in.grad), rel=REL_TOL)
    with pytest.raises(NotImplementedError):
        P.compute_reg_loss([torch.ones(1, 4, 2, 2, device=DEV)], torch.ones(1, 4, 4, device=DEV),
                           torch.ones(1, 4, dtype=torch.bool, device=DEV), mode="diou")
    # focal_loss_from_logits on [P, C] logits / one-hot
    torch.manual_seed(0)
    logits = torch.randn(37, 6) - 2
    hot = torch.zeros(37, 6); hot[torch.arange(0, 37, 3), torch.arange(0, 37, 3) % 6] = 1
    assert_close(float(P.focal_loss_from_logits(logits.to(DEV), hot.to(DEV))), float(O.focal_sum(logits, hot)),
                 rel=REL_TOL)


def test_focal_wide_logit_range():
    """Focal loss and its gradient over logits in [-30, 14] (both branches of the in-kernel log, exp
    overflow, the clip at 5e-6) against the reference formula evaluated by torch on the CPU (loss.py:180-193)."""
    gen = torch.Generator().manual_seed(9)
    n, c = 4096, 8
    logits = torch.cat([torch.linspace(-30, 14, n * c // 2), (torch.rand(n * c // 2, generator=gen) - 0.5) * 16])
    logits = logits[torch.randperm(n * c, generator=gen)].reshape(n, c)
    hot = torch.zeros(n, c)
    rows = torch.arange(0, n, 5)
    hot[rows, rows % c] = 1
    a = logits.clone().to(DEV).requires_grad_(True)
    b = logits.clone().requires_grad_(True)
    la = P.focal_loss_from_logits(a, hot.to(DEV))
    lb = O.focal_sum(b, hot)
    assert_close(float(la), float(lb), rel=REL_TOL)
    la.backward()
    lb.backward()
    assert_close(to_np(a.grad), to_np(b.grad), rel=REL_TOL, abs_=1e-9)


def test_box_loss_tie_subgradients_match_autograd():
    """Exact min/max ties (pred == target component) split the gradient 1/2-1/2 like torch."""
    p = torch.tensor([[4.0, 6.0, 8.0, 3.0], [2.0, 2.0, 2.0, 2.0], [1.0, 9.0, 4.0, 4.0]])
    t = torch.tensor([[4.0, 5.0, 8.0, 7.0], [2.0, 2.0, 2.0, 2.0], [3.0, 9.0, 1.0, 4.0]])
    for fn_gpu, fn_cpu in ((P.giou_loss, O.giou_sum), (P.iou_loss, O.iou_sum)):
        a = p.clone().to(DEV).requires_grad_(True)
        b = p.clone().requires_grad_(True)
        fn_gpu(a, t.to(DEV)).backward()
        fn_cpu(b, t).backward()
        assert_close(to_np(a.grad), to_np(b.grad), rel=REL_TOL, abs_=1e-9)


# ------------------------------------------------------------------------------------------
# K4 fused: targets + box / centerness loss forward and backward in one launch
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(TRAIN_CASES))
@pytest.mark.parametrize("mode", ["giou", "iou"])
def test_fused_target_loss_matches_reference(name, mode):
    """FCOSTargetLoss([out, gt, labels]) == FCOSLoss([out, FCOSGenTargets(...)]) of the reference (golden)."""
    g, x, gt, labels, ranges, levels = train_case(name)
    xc = cuda_levels(x)
    for part in xc:
        for t in part:
            t.requires_grad_(True)
    step = P.FCOSTargetLoss(W.STRIDES, ranges, mode)
    losses = step([xc, gt.to(DEV), labels.to(DEV)])
    cls_t, cnt_t, reg_t = step.targets
    assert_equal_int(to_np(cls_t), g["cls_t"], what="cls_t")
    assert np.array_equal(to_np(reg_t), g["reg_t"]), "reg_t not bit-exact"
    assert_cnt_matches(cnt_t, g["cnt_t"], g["reg_t"])
    assert all(v.dim() == 0 for v in losses)
    assert_close([float(v) for v in losses], g[f"loss_{mode}"], rel=REL_TOL, what=f"loss {mode}")
    losses[3].backward()
    for lv in range(len(levels)):
        assert_close(to_np(xc[2][lv].grad), g[f"g_reg_{mode}_{lv}"], rel=REL_TOL, abs_=1e-9, what=f"g_reg {lv}")
        assert_close(to_np(xc[1][lv].grad), g[f"g_cnt_{mode}_{lv}"], rel=REL_TOL, abs_=1e-9, what=f"g_cnt {lv}")
        if f"g_cls_{mode}_{lv}" in g.files:
            assert_close(to_np(xc[0][lv].grad), g[f"g_cls_{mode}_{lv}"], rel=REL_TOL, abs_=1e-12, what=f"g_cls {lv}")
        else:
            assert_close(to_np(xc[0][lv].grad.double().sum(dim=(2, 3))), g[f"g_cls_sum_{mode}_{lv}"], rel=1e-4,
                         abs_=1e-9, what=f"g_cls sums {lv}")


def test_focal_step_wide_logit_range_odd_levels_and_upstream_scale():
    """b200det_cls_loss_step (loss + gradient from one read of the logits) over logits in [-30, 14], on levels
    that take the 128-bit path, the scalar path (13x21) and a ragged last tile, with C = 21 (class-chunk
    remainder of 5 = one 4-plane group + one slow plane), against the reference formula under CPU autograd
    (loss.py:6-26, 180-193); then FCOSLoss with a non-unit upstream gradient against the two-kernel path.

    Gradient floor: the reference forms om = 1 - fl(1 - p), a multiple of 2^-24.  A 1-ulp difference between
    torch's CPU sigmoid and the kernel's moves fl(1 - p) by one step on ~1 % of the elements, i.e. om by 6e-8
    ABSOLUTE; the gradient is ~2.25 om^3, so that step alone is 4e-7 om^2 — above 1e-5 relative for om < 0.05
    and below 1e-9 in absolute terms.  Same floor as test_focal_wide_logit_range, times the gradient scale."""
    gen = torch.Generator().manual_seed(17)
    B, C = 3, 21
    levels = [(40, 52), (13, 21), (7, 12), (1, 3)]
    P_total = sum(h * w for h, w in levels)
    cls = []
    for h, w in levels:
        n = B * C * h * w
        v = torch.cat([torch.linspace(-30, 14, n // 2), (torch.rand(n - n // 2, generator=gen) - 0.5) * 16])
        cls.append(v[torch.randperm(n, generator=gen)].reshape(B, C, h, w))
    cls_t = torch.zeros(B, P_total, 1, dtype=torch.int64)
    pos = torch.rand(B, P_total, generator=gen) < 0.03
    pos[2] = False                                                       # an image without positives: num_pos clamps to 1
    cls_t[pos] = torch.randint(1, C + 1, (int(pos.sum()), 1), generator=gen)
    cnt_t = torch.where(pos, 0.5, -1.0).reshape(B, P_total, 1)
    ref_in = [t.clone().requires_grad_(True) for t in cls]
    want = O.cls_loss(ref_in, cls_t, pos)
    want.mean().backward()
    a = [t.clone().to(DEV).requires_grad_(True) for t in cls]
    loss, mean, npos, grads = ops.cls_loss_step(a, cls_t.to(DEV), mask_src=cnt_t.to(DEV))
    assert_close(to_np(loss), to_np(want), rel=REL_TOL, what="per-image focal loss")
    assert_close(float(mean[0]), float(want.mean()), rel=REL_TOL)
    assert float(mean[1]) == 1.0                                  # no upstream gradient was given: 1 assumed
    assert_equal_int(to_np(npos), np.maximum(to_np(pos.sum(dim=1)), 1))
    for gpu, ref in zip(grads, ref_in):
        assert_close(to_np(gpu), to_np(ref.grad), rel=REL_TOL, abs_=1e-9 / B, what="focal step gradient")
    # num_pos handed in (the fused assignment made it) + per-image upstream gradients
    up = torch.tensor([0.25, 2.0, -1.0])
    _, _, _, grads2 = ops.cls_loss_step(a, cls_t.to(DEV), num_pos=npos, grad_loss=up.to(DEV))
    for t in ref_in:
        t.grad = None
    (O.cls_loss(ref_in, cls_t, pos) * up).sum().backward()
    for gpu, ref in zip(grads2, ref_in):
        assert_close(to_np(gpu), to_np(ref.grad), rel=REL_TOL, abs_=2e-9, what="focal step gradient, upstream")
    # through the modules: FCOSLoss (gradient written by the forward kernel, rescaled in backward) == two kernels
    cnt = [torch.randn(B, 1, h, w, generator=gen).to(DEV).requires_grad_(True) for h, w in levels]
    reg = [torch.exp(torch.randn(B, 4, h, w, generator=gen)).to(DEV).requires_grad_(True) for h, w in levels]
    reg_t = torch.where(pos[..., None], torch.rand(B, P_total, 4, generator=gen) * 50 + 1, -1.0)
    tgt = (cls_t.to(DEV), cnt_t.to(DEV), reg_t.to(DEV))
    losses = P.FCOSLoss("giou")([(a, cnt, reg), tgt])
    (2.5 * losses[3]).backward()
    two = P.compute_cls_loss([t.detach().clone().requires_grad_(True) for t in a], tgt[0], None, _mask_src=tgt[1])
    assert_close(float(losses[0]), float(two.mean()), rel=REL_TOL)
    b2 = [t.detach().clone().requires_grad_(True) for t in a]
    (2.5 * P.compute_cls_loss(b2, tgt[0], None, _mask_src=tgt[1]).mean()).backward()
    for x, y in zip(a, b2):
        assert_close(to_np(x.grad), to_np(y.grad), rel=REL_TOL, abs_=1e-12, what="FCOSLoss cls gradient")   # same sigmoid
    with torch.no_grad():                                                # no gradient asked for: forward kernel only
        l0 = P.FCOSLoss("giou")([(a, cnt, reg), tgt])[0]
    assert_close(float(l0), float(losses[0]), rel=REL_TOL)


def test_fused_target_loss_equals_unfused_kernels_full_size_and_upstream_scale():
    """Config 3 at full size: the fused launch against the separate assign / loss kernels — targets
    bit-identical, losses and gradients to 1e-5 — with a non-unit upstream gradient (loss scaling),
    with and without the centerness branch, repeated so the self-resetting ticket is exercised."""
    B, M = 32, 100
    gt, labels = W.gt_boxes(B, M, W.COCO_HW, 80, seed=411)
    gt, labels = gt.to(DEV), labels.to(DEV)
    gen = torch.Generator().manual_seed(412)
    reg = [torch.exp(torch.randn(B, 4, h, w, generator=gen) + 3).to(DEV).requires_grad_(True) for h, w in W.COCO_LEVELS]
    cnt = [torch.randn(B, 1, h, w, generator=gen).to(DEV).requires_grad_(True) for h, w in W.COCO_LEVELS]
    want_t = ops.assign_targets(W.COCO_LEVELS, W.STRIDES, W.HISFCOS_RANGES, gt, labels)
    want_reg = P.compute_reg_loss(reg, want_t[2], None, "giou", _mask_src=want_t[1])
    want_cnt = P.compute_cnt_loss(cnt, want_t[1], None, _mask_src=want_t[1])
    (3.0 * want_reg.mean() + 0.5 * want_cnt.mean()).backward()
    want_g = [t.grad.clone() for t in reg + cnt]
    step = P.FCOSTargetLoss(W.STRIDES, W.HISFCOS_RANGES, "giou")
    for rep in range(3):
        for t in reg + cnt:
            t.grad = None
        reg_loss, cnt_loss = step.box_cnt_losses(cnt, reg, gt, labels)
        for a, b in zip(step.targets, want_t):
            assert torch.equal(a, b)
        assert_close(to_np(step.per_image["reg"]), to_np(want_reg), rel=REL_TOL)
        assert_close(to_np(step.per_image["cnt"]), to_np(want_cnt), rel=REL_TOL)
        assert_close(float(reg_loss), float(want_reg.mean()), rel=REL_TOL)
        assert_close(float(cnt_loss), float(want_cnt.mean()), rel=REL_TOL)
        (3.0 * reg_loss + 0.5 * cnt_loss).backward()
        for t, w in zip(reg + cnt, want_g):
            assert_close(to_np(t.grad), to_np(w), rel=REL_TOL, abs_=1e-10)
    # box loss only (no centerness maps), unit upstream
    for t in reg:
        t.grad = None
    reg_only, none = step.box_cnt_losses(None, reg, gt, labels)
    assert none is None
    reg_only.backward()
    for t, w in zip(reg, want_g[:5]):
        assert_close(to_np(t.grad), to_np(w) / 3.0, rel=REL_TOL, abs_=1e-10)


def test_training_step_under_a_loss_scale_assumes_the_previous_upstream_gradient():
    """train.py:175-181 runs the step under GradScaler: total_loss arrives in backward with the loss scale as
    upstream gradient.  The step kernels write their gradients in the forward pass for an ASSUMED upstream
    (start: 1), backward rescales on a wrong assumption and stores what arrived.  Every step must give the
    reference's gradients times the scale — first step (wrong assumption), steady state, a scale change, a
    zero upstream (grads 0, assumption kept) — against CPU autograd of the oracle."""
    g, x, gt, labels, ranges, levels = train_case("train_voc_b2")
    xc = cuda_levels(x)
    flat = [t for part in xc for t in part]
    for t in flat:
        t.requires_grad_(True)
    ref = [[t.clone().requires_grad_(True) for t in part] for part in x]
    tgt = O.assign_targets(levels, gt, labels, W.STRIDES, ranges)
    O.fcos_loss(ref, tgt[:3], "giou")[3].backward()
    want = [t.grad.clone() for part in ref for t in part]
    step = P.FCOSTargetLoss(W.STRIDES, ranges, "giou")
    plain = P.FCOSLoss("giou")
    for scale in (1024.0, 1024.0, 512.0, 0.0, 512.0):
        for module, arg in ((step, [xc, gt.to(DEV), labels.to(DEV)]), (plain, None)):
            for t in flat:
                t.grad = None
            if arg is None:
                arg = [xc, P.FCOSGenTargets(W.STRIDES, ranges)([xc, gt.to(DEV), labels.to(DEV)])]
            (module(arg)[3] * scale).backward()
            for got, w in zip(flat, want):
                assert_close(to_np(got.grad), to_np(w) * scale, rel=REL_TOL, abs_=1e-9 * max(scale, 1.0),
                             what=f"{type(module).__name__} gradient at loss scale {scale}")
    for up in (step._up_cls, step._up_box, step._up_cnt, plain._up_cls):
        assert up.on(torch.device(DEV)).tolist() == [512.0, 0.0]  # the zero upstream was not adopted; ticket reset


def test_fused_step_recount_path_many_gt_and_multi_wave_grids():
    """The paths of the fill + patch kernel that a one-wave COCO batch of 32 never takes: (1) num_pos by the local
    recount (forced through B200DET_FORCE_RECOUNT=1: what a CTA does when the other tiles of its image are late) must
    give the counter's result bit for bit; (2) more ground-truth boxes than a chain thread stages in registers
    (M > 448) and more than the 48 KB default of shared memory (M ~ 1100); (3) a grid of several waves (B = 70 at the
    COCO size: 1 260 CTAs on 592 slots) against the separate kernels."""
    import os
    levels = W.VOC_LEVELS
    for batch, m, hw, lv, seed in ((3, 40, W.VOC_HW, W.VOC_LEVELS, 71), (2, 1100, W.VOC_HW, W.VOC_LEVELS, 72),
                                   (70, 30, W.COCO_HW, W.COCO_LEVELS, 73)):
        gt, labels = W.gt_boxes(batch, m, hw, 20, seed=seed)
        gt, labels = gt.to(DEV), labels.to(DEV)
        gen = torch.Generator(device=DEV).manual_seed(seed)
        reg = [torch.exp(torch.randn(batch, 4, h, w, device=DEV, generator=gen) + 3) for h, w in lv]
        cnt = [torch.randn(batch, 1, h, w, device=DEV, generator=gen) for h, w in lv]
        a = ops.assign_loss_fused(reg, cnt, W.STRIDES, W.FCOS_RANGES, gt, labels, 1)
        os.environ["B200DET_FORCE_RECOUNT"] = "1"
        try:
            r = ops.assign_loss_fused(reg, cnt, W.STRIDES, W.FCOS_RANGES, gt, labels, 1)
        finally:
            del os.environ["B200DET_FORCE_RECOUNT"]
        for key in ("cls_t", "cnt_t", "reg_t", "box_loss", "cnt_loss", "num_pos", "mean"):
            assert torch.equal(a[key], r[key]), f"{key}: recount path differs (B={batch}, M={m})"
        for x, y in zip(a["reg_grads"] + a["cnt_grads"], r["reg_grads"] + r["cnt_grads"]):
            assert torch.equal(x, y), "gradients: recount path differs"
        # against the separate kernels: targets bit-exact, losses and gradients of the batch means
        t = ops.assign_targets(lv, W.STRIDES, W.FCOS_RANGES, gt, labels)
        for key, w_ in zip(("cls_t", "cnt_t", "reg_t"), t):
            assert torch.equal(a[key], w_), key
        loss, npos = ops.box_loss_fwd(reg, t[1], t[2], 1)
        assert torch.equal(npos, a["num_pos"])
        assert_close(to_np(a["box_loss"]), to_np(loss), REL_TOL, what="box loss")
        gl = torch.full((batch,), 1.0 / batch, device=DEV)
        for x, y in zip(a["reg_grads"], ops.box_loss_bwd(reg, t[1], t[2], 1, gl, npos)):
            assert_close(to_np(x), to_np(y), REL_TOL, abs_=1e-12, what="box gradients")
    # the oracle on the many-GT case (labels, GT indices bit-exact)
    gt, labels = W.gt_boxes(2, 1100, W.VOC_HW, 20, seed=72)
    got = ops.assign_targets(levels, W.STRIDES, W.FCOS_RANGES, gt.to(DEV), labels.to(DEV), want_index=True)
    want = O.assign_targets(levels, gt, labels, W.STRIDES, W.FCOS_RANGES)
    assert_equal_int(to_np(got[0]), to_np(want[0]), what="labels, M=1100")
    assert_equal_int(to_np(got[3]), to_np(want[3]), what="GT index, M=1100")
    assert np.array_equal(to_np(got[2]), to_np(want[2]))


def test_fill_and_patch_kernel_stays_inside_its_buffers():
    """compute-sanitizer is closed on this pool, so the bulk-copy fill is checked with guard bands: every output of
    b200det_assign_targets / b200det_assign_loss_fused is placed (through the C ABI) in the middle of a larger buffer
    filled with a canary, at the odd point counts that make the fill's scalar heads / tails non-trivial (P odd, a
    13 x 21 level, images whose first point is not 16-byte aligned); the canaries must survive and the payload must equal
    the normal call's."""
    import ctypes as C
    lib = _lib.load()
    GUARD = 1024                                         # bytes on either side
    levels = [(13, 21), (7, 11), (5, 5), (3, 3), (1, 1)]            # P = 273 + 77 + 25 + 9 + 1 = 385 (odd)
    strides = W.STRIDES
    ranges = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, 999999]]
    hw_img = (13 * 8, 21 * 8)
    batch, m = 3, 9
    p_total = sum(h * w for h, w in levels)
    gt, labels = W.gt_boxes(batch, m, hw_img, 20, seed=81)
    gt, labels = gt.to(DEV), labels.to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(81)
    reg = [torch.exp(torch.randn(batch, 4, h, w, device=DEV, generator=gen) + 2) for h, w in levels]
    cnt = [torch.randn(batch, 1, h, w, device=DEV, generator=gen) for h, w in levels]
    want = ops.assign_loss_fused(reg, cnt, strides, ranges, gt, labels, 1)
    want_t = ops.assign_targets(levels, strides, ranges, gt, labels, want_index=True)

    def guarded(nbytes):
        buf = torch.full((nbytes + 2 * GUARD,), 0xA5, dtype=torch.uint8, device=DEV)
        return buf, buf.data_ptr() + GUARD

    def check(buf, nbytes, what):
        assert bool((buf[:GUARD] == 0xA5).all()) and bool((buf[GUARD + nbytes:] == 0xA5).all()), f"{what}: write outside the buffer"
        return buf[GUARD:GUARD + nbytes]

    n = len(levels)
    hw_arr = (C.c_int32 * (2 * n))(*[v for hw in levels for v in hw])
    st_arr = (C.c_int32 * n)(*strides)
    lo_arr = (C.c_float * n)(*[float(r[0]) for r in ranges])
    hi_arr = (C.c_float * n)(*[float(r[1]) for r in ranges])
    ra_arr = (C.c_float * n)(*[float(s_ * 1.5) for s_ in strides])
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    sizes = {"cls_t": batch * p_total * 8, "cnt_t": batch * p_total * 4, "reg_t": batch * p_total * 16, "idx": batch * p_total * 4}
    # ---- assign only -----------------------------------------------------------------------------------------
    bufs = {k: guarded(v) for k, v in sizes.items()}
    rc = lib.b200det_assign_targets(hw_arr, st_arr, lo_arr, hi_arr, ra_arr, n, batch, m, gt.data_ptr(), labels.data_ptr(),
                                    bufs["cls_t"][1], bufs["cnt_t"][1], bufs["reg_t"][1], bufs["idx"][1], stream)
    assert rc == 0
    torch.cuda.synchronize()
    for key, w_, dt in (("cls_t", want_t[0], torch.int64), ("cnt_t", want_t[1], torch.float32),
                        ("reg_t", want_t[2], torch.float32), ("idx", want_t[3], torch.int32)):
        got = check(bufs[key][0], sizes[key], key).view(dt)
        assert torch.equal(got, w_.reshape(-1)), key
    # ---- fused: targets + gradients -------------------------------------------------------------------------------
    bufs = {k: guarded(v) for k, v in sizes.items() if k != "idx"}
    greg = [guarded(t.numel() * 4) for t in reg]
    gcnt = [guarded(t.numel() * 4) for t in cnt]
    small = {k: guarded(batch * 4) for k in ("box_loss", "cnt_loss", "num_pos")}
    mean = guarded(16)
    lv = _lib.make_levels([(0, c_.data_ptr(), r_.data_ptr(), h, w, s_) for c_, r_, (h, w), s_ in zip(cnt, reg, levels, strides)])
    ws_bytes = lib.b200det_assign_loss_workspace_bytes(batch, p_total)
    ws = torch.zeros(ws_bytes + 2 * GUARD, dtype=torch.uint8, device=DEV)
    ws[:GUARD] = 0xA5
    ws[GUARD + ws_bytes:] = 0xA5
    rc = lib.b200det_assign_loss_fused(lv, (C.c_void_p * n)(*[g_[1] for g_ in greg]), (C.c_void_p * n)(*[g_[1] for g_ in gcnt]), n,
                                       lo_arr, hi_arr, ra_arr, batch, m, gt.data_ptr(), labels.data_ptr(), 1, None, None, 0,
                                       bufs["cls_t"][1], bufs["cnt_t"][1], bufs["reg_t"][1], small["box_loss"][1],
                                       small["cnt_loss"][1], small["num_pos"][1], mean[1], None, ws.data_ptr() + GUARD, ws_bytes,
                                       stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert bool((ws[:GUARD] == 0xA5).all()) and bool((ws[GUARD + ws_bytes:] == 0xA5).all()), "workspace overrun"
    for key, dt in (("cls_t", torch.int64), ("cnt_t", torch.float32), ("reg_t", torch.float32)):
        assert torch.equal(check(bufs[key][0], sizes[key], key).view(dt), want[key].reshape(-1)), key
    for (buf, _), w_ in zip(greg, want["reg_grads"]):
        assert torch.equal(check(buf, w_.numel() * 4, "reg grad").view(torch.float32), w_.reshape(-1))
    for (buf, _), w_ in zip(gcnt, want["cnt_grads"]):
        assert torch.equal(check(buf, w_.numel() * 4, "cnt grad").view(torch.float32), w_.reshape(-1))
    for key in ("box_loss", "cnt_loss", "num_pos"):
        assert torch.equal(check(small[key][0], batch * 4, key).view(torch.float32), want[key])
    assert torch.equal(check(mean[0], 16, "mean").view(torch.float32), want["mean"])


def test_two_outstanding_forwards_and_per_call_loss_weights():
    """Two forwards of ONE module before any backward — (lossA + lossB).backward() on the first GradScaler step,
    gradient accumulation with per-micro-batch weights: every forward keeps its own copy of the upstream gradient it
    assumed, so the second backward is not fooled by the state the first one has already updated.  Against CPU
    autograd of the oracle for each input set."""
    g, x, gt, labels, ranges, levels = train_case("train_voc_b2")
    x2 = W.head_outputs(x[0][0].shape[0], x[0][0].shape[1], levels, seed=4242)
    gt2, labels2 = W.gt_boxes(gt.shape[0], gt.shape[1], W.VOC_HW, x[0][0].shape[1], seed=4243)
    sets = [(x, gt, labels), (x2, gt2, labels2)]

    def reference(weights):
        out = []
        for (xs, g_, l_), w_ in zip(sets, weights):
            ref = [[t.clone().requires_grad_(True) for t in part] for part in xs]
            tgt = O.assign_targets(levels, g_, l_, W.STRIDES, ranges)
            (O.fcos_loss(ref, tgt[:3], "giou")[3] * w_).backward()
            out.append([t.grad.clone() for part in ref for t in part])
        return out

    for module_kind in ("fused", "plain"):
        module = P.FCOSTargetLoss(W.STRIDES, ranges, "giou") if module_kind == "fused" else P.FCOSLoss("giou")
        gen = P.FCOSGenTargets(W.STRIDES, ranges)
        for weights in ((1024.0, 1024.0), (1024.0, 1024.0), (2.0, 3.0), (3.0, 2.0), (1.0, 1.0)):
            want = reference(weights)
            dev_sets = [([[t.to(DEV).requires_grad_(True) for t in part] for part in xs], g_.to(DEV), l_.to(DEV))
                        for xs, g_, l_ in sets]
            losses = []
            for xs, g_, l_ in dev_sets:                         # BOTH forwards first
                arg = [xs, g_, l_] if module_kind == "fused" else [xs, gen([xs, g_, l_])]
                losses.append(module(arg)[3])
            (losses[0] * weights[0] + losses[1] * weights[1]).backward()
            for (xs, _, _), w_list, wt in zip(dev_sets, want, weights):
                for got, w in zip([t for part in xs for t in part], w_list):
                    assert_close(to_np(got.grad), to_np(w), rel=REL_TOL, abs_=1e-9 * max(wt, 1.0),
                                 what=f"{module_kind} gradients, two outstanding forwards, weights {weights}")


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_focal_step_reads_half_logits_and_writes_half_gradients(dtype):
    """Under autocast (train.py:175) the class logits are fp16 conv outputs.  b200det_cls_loss_step reads them as
    they are and writes the gradient in their type; the arithmetic is the fp32 kernel's: the loss must be
    BIT-IDENTICAL to the fp32 kernel on the up-cast logits and the gradient identical to its gradient rounded
    once.  Levels cover the staged path (hw % 8 == 0), the 4-element path (26 x 42, 7 x 12) and the scalar path."""
    gen = torch.Generator().manual_seed(23)
    B, C = 2, 21
    levels = [(40, 52), (26, 42), (13, 21), (7, 12), (1, 3)]
    P_total = sum(h * w for h, w in levels)
    cls = [((torch.randn(B, C, h, w, generator=gen) * 2.0 - 3.0).to(dtype)).to(DEV) for h, w in levels]
    cls_t = torch.zeros(B, P_total, 1, dtype=torch.int64)
    pos = torch.rand(B, P_total, generator=gen) < 0.03
    cls_t[pos] = torch.randint(1, C + 1, (int(pos.sum()), 1), generator=gen)
    cnt_t = torch.where(pos, 0.5, -1.0).reshape(B, P_total, 1).to(DEV)
    state = torch.tensor([1024.0, 0.0], device=DEV)                      # a loss scale: fp16 gradients need one
    loss_h, mean_h, npos_h, grads_h = ops.cls_loss_step(cls, cls_t.to(DEV), mask_src=cnt_t, up_mean=state)
    loss_f, mean_f, npos_f, grads_f = ops.cls_loss_step([t.float() for t in cls], cls_t.to(DEV), mask_src=cnt_t,
                                                        up_mean=state)
    assert torch.equal(loss_h, loss_f) and torch.equal(mean_h, mean_f) and torch.equal(npos_h, npos_f)
    for gh, gf in zip(grads_h, grads_f):
        assert gh.dtype == dtype and torch.equal(gh, gf.to(dtype))
    # against the reference formula on the CPU (fp32 on the up-cast logits)
    ref_in = [t.float().cpu().requires_grad_(True) for t in cls]
    want = O.cls_loss(ref_in, cls_t, pos)
    (want.mean() * 1024.0).backward()
    assert_close(to_np(loss_h), to_np(want), rel=REL_TOL, what="focal loss of half logits")
    half_ulp = 2.0 ** -11 if dtype == torch.float16 else 2.0 ** -8
    for gh, r in zip(grads_h, ref_in):
        assert_close(to_np(gh.float()), to_np(r.grad), rel=2.0 * half_ulp, abs_=1e-7, what="half focal gradient")
    # a wrong assumption is repaired in the half maps themselves (x 0.5: exact in fp16 above the subnormals)
    got = torch.tensor([512.0], device=DEV)
    ops.rescale_maps_(grads_h, [got] * len(grads_h), [state] * len(grads_h))
    assert state.tolist() == [512.0, 0.0]
    for gh, r in zip(grads_h, ref_in):
        assert_close(to_np(gh.float()), to_np(r.grad) * 0.5, rel=3.0 * half_ulp, abs_=1e-7, what="rescaled half gradient")


def test_training_step_with_autocast_style_half_logits():
    """FCOSTargetLoss on fp16 cls / cnt maps (what autocast hands over) under a loss scale, three steps: the first
    takes the two-kernel focal path and taps the upstream gradient, the next ones the step kernel on fp16 maps."""
    g, x, gt, labels, ranges, levels = train_case("train_voc_b2")
    cls32, cnt32, reg32 = x
    xh = ([t.half().to(DEV).requires_grad_(True) for t in cls32], [t.half().to(DEV).requires_grad_(True) for t in cnt32],
          [t.to(DEV).requires_grad_(True) for t in reg32])
    ref = ([t.half().float().requires_grad_(True) for t in cls32], [t.half().float().requires_grad_(True) for t in cnt32],
           [t.clone().requires_grad_(True) for t in reg32])
    tgt = O.assign_targets(levels, gt, labels, W.STRIDES, ranges)
    want_losses = O.fcos_loss(ref, tgt[:3], "giou")
    want_losses[3].backward()
    step = P.FCOSTargetLoss(W.STRIDES, ranges, "giou")
    for it, scale in enumerate((4096.0, 4096.0, 2048.0)):
        for part in xh:
            for t in part:
                t.grad = None
        losses = step([xh, gt.to(DEV), labels.to(DEV)])
        assert step._up_cls.observed == (it > 0)
        assert_close([float(v) for v in losses], [float(v) for v in want_losses], rel=REL_TOL, what="losses, half logits")
        (losses[3] * scale).backward()
        for part, rpart in zip(xh, ref):
            for t, r in zip(part, rpart):
                assert t.grad.dtype == t.dtype
                assert_close(to_np(t.grad.float()), to_np(r.grad) * scale, rel=2e-3, abs_=1e-6 * scale,
                             what=f"step {it} gradient ({t.dtype})")


def test_fused_target_loss_edge_cases():
    """No GT at all, M = 0, one image, a slice-boundary-heavy tiny pyramid, dense crowd (300 GT)."""
    step = P.FCOSTargetLoss(W.STRIDES, W.FCOS_RANGES, "giou")
    gen = torch.Generator().manual_seed(5)

    def maps(b, levels):
        reg = [torch.exp(torch.randn(b, 4, h, w, generator=gen) + 2).to(DEV).requires_grad_(True) for h, w in levels]
        cnt = [torch.randn(b, 1, h, w, generator=gen).to(DEV).requires_grad_(True) for h, w in levels]
        return reg, cnt

    def check(levels, gt, labels, ranges=W.FCOS_RANGES):
        st = P.FCOSTargetLoss(W.STRIDES[:len(levels)], ranges[:len(levels)], "giou")
        reg, cnt = maps(gt.shape[0], levels)
        a, c = st.box_cnt_losses(cnt, reg, gt.to(DEV), labels.to(DEV))
        (a + c).backward()
        got_g = [t.grad.clone() for t in reg + cnt]
        for t in reg + cnt:
            t.grad = None
        want_t = ops.assign_targets(levels, W.STRIDES[:len(levels)], ranges[:len(levels)], gt.to(DEV), labels.to(DEV))
        wr = P.compute_reg_loss(reg, want_t[2], None, "giou", _mask_src=want_t[1]).mean()
        wc = P.compute_cnt_loss(cnt, want_t[1], None, _mask_src=want_t[1]).mean()
        (wr + wc).backward()
        for x, y in zip(st.targets, want_t):
            assert torch.equal(x, y)
        assert_close(float(a), float(wr), rel=REL_TOL, abs_=1e-12)
        assert_close(float(c), float(wc), rel=REL_TOL, abs_=1e-12)
        for t, w in zip(reg + cnt, got_g):
            assert_close(to_np(w), to_np(t.grad), rel=REL_TOL, abs_=1e-10)
        return st

    # all padding rows
    check(W.VOC_LEVELS, torch.full((2, 3, 4), -1.0), torch.full((2, 3), -1, dtype=torch.int64))
    # M = 0
    st = check(W.VOC_LEVELS, torch.zeros((2, 0, 4)), torch.zeros((2, 0), dtype=torch.int64))
    assert float(st.per_image["num_pos"].max()) == 1.0
    # batch 1, tiny pyramid whose levels are smaller than a slice (several levels per CTA, empty CTAs)
    gt, labels = W.gt_boxes(1, 6, (64, 96), 5, seed=77)
    check([(8, 12), (4, 6), (2, 3), (1, 2), (1, 1)], gt, labels)
    # dense crowd: BASELINE config 4's assignment side
    gt, labels = W.gt_boxes(4, 300, W.COCO_HW, 80, seed=78)
    check(W.COCO_LEVELS, gt, labels, W.HISFCOS_RANGES)


# ------------------------------------------------------------------------------------------
# N2: ScaleExp (exp(x * scale), modules.py:170-176 / HISFcos.py:228) folded into the consumers
# ------------------------------------------------------------------------------------------
def _raw_reg_case(batch, levels, ncls, seed):
    """Head outputs whose regression branch is still RAW (before ScaleExp) plus per-level scales."""
    gen = torch.Generator().manual_seed(seed)
    x = W.head_outputs(batch, ncls, levels, seed)
    scales = [torch.tensor([1.2 - 0.07 * i]) for i in range(len(levels))]               # HISFcos.py:209: init 1.2
    raw = [torch.randn(batch, 4, h, w, generator=gen) * 0.6 + 2.2 for h, w in levels]   # exp(1.2 * 2.2) ~ 14 px
    return x, raw, scales


def test_head_with_folded_scale_exp_matches_reference_order():
    """FCOSHead(reg_exp_scales=...) on raw reg outputs == the reference head on exp(raw * scale)."""
    x, raw, scales = _raw_reg_case(2, W.COCO_LEVELS, 80, seed=91)
    reg_ref = [torch.exp(r * s) for r, s in zip(raw, scales)]                # what HISFCOSHead.forward emits
    want = O.detect((x[0], x[1], reg_ref), 0.05, 0.6, 1000, W.STRIDES)
    head = P.FCOSHead(0.05, 0.6, 1000, W.STRIDES, reg_exp_scales=[s.to(DEV) for s in scales])
    s_, c_, b_, n_ = head.detect(([t.to(DEV) for t in x[0]], [t.to(DEV) for t in x[1]], [r.to(DEV) for r in raw]))
    for i, (ws, wc, wb) in enumerate(want):
        k = int(n_[i])
        assert_detections_match((to_np(s_[i, :k]), to_np(c_[i, :k]), to_np(b_[i, :k])),
                                (to_np(ws), to_np(wc), to_np(wb)), what=f"image {i}")


@pytest.mark.parametrize("mode", ["giou", "iou"])
def test_fused_target_loss_with_folded_scale_exp(mode):
    """Gradients w.r.t. the raw regression outputs AND the ScaleExp scales against CPU autograd through
    exp(raw * scale) -> reference assignment -> reference losses."""
    B, M = 3, 12
    x, raw, scales = _raw_reg_case(B, W.COCO_LEVELS, 80, seed=92)
    gt, labels = W.gt_boxes(B, M, W.COCO_HW, 80, seed=93)
    # CPU reference
    raw_c = [r.clone().requires_grad_(True) for r in raw]
    sc_c = [s.clone().requires_grad_(True) for s in scales]
    cnt_c = [t.clone().requires_grad_(True) for t in x[1]]
    reg_c = [torch.exp(r * s) for r, s in zip(raw_c, sc_c)]
    tgt = O.assign_targets(W.COCO_LEVELS, gt, labels, W.STRIDES, W.HISFCOS_RANGES)
    mask = tgt[1].squeeze(-1) > -1
    want_reg = O.reg_loss(reg_c, tgt[2], mask, mode).mean()
    want_cnt = O.cnt_loss(cnt_c, tgt[1], mask).mean()
    (want_reg + 0.5 * want_cnt).backward()
    # fused CUDA path
    raw_g = [r.clone().to(DEV).requires_grad_(True) for r in raw]
    sc_g = [s.clone().to(DEV).requires_grad_(True) for s in scales]
    cnt_g = [t.clone().to(DEV).requires_grad_(True) for t in x[1]]
    step = P.FCOSTargetLoss(W.STRIDES, W.HISFCOS_RANGES, mode, reg_exp_scales=sc_g)
    got_reg, got_cnt = step.box_cnt_losses(cnt_g, raw_g, gt.to(DEV), labels.to(DEV))
    assert_close(float(got_reg), float(want_reg), rel=REL_TOL)
    assert_close(float(got_cnt), float(want_cnt), rel=REL_TOL)
    (got_reg + 0.5 * got_cnt).backward()
    for a, w in zip(raw_g + cnt_g, raw_c + cnt_c):
        assert_close(to_np(a.grad), to_np(w.grad), rel=REL_TOL, abs_=1e-9)
    for lv, (a, w) in enumerate(zip(sc_g, sc_c)):
        assert a.grad.shape == w.grad.shape
        assert_close(to_np(a.grad), to_np(w.grad), rel=2e-5, abs_=1e-7, what=f"scale grad level {lv}")
    # the stand-alone box-loss entry points refuse the folded form instead of mis-reading raw values
    lv, keep, *_ = ops._levels(None, None, raw_g, [1] * 5, sc_g)
    assert _lib.load().b200det_box_loss_fwd(lv, 5, B, 0, 0, 1, 0, 0, None) != 0


# ------------------------------------------------------------------------------------------
# N4: the datasets' collate_fn on the device
# ------------------------------------------------------------------------------------------
COLLATE_CASES = {"ragged3": (1, [(37, 53), (64, 41), (5, 64)], [3, 0, 7]), "same2": (2, [(32, 48), (32, 48)], [4, 4]),
                 "one": (3, [(1, 1)], [2]), "coco4": (4, [(96, 128), (80, 132), (100, 100), (64, 160)], [11, 1, 0, 25])}


@pytest.mark.parametrize("name", sorted(COLLATE_CASES))
def test_device_collate_matches_reference_collate(name):
    """DeviceCollate against the outputs of the reference's own collate_fn (dataset/voc.py:141-173 and
    dataset/coco.py:135-165, executed by tests/golden/make_golden_r2.py) on the same seeded ragged batch."""
    g = load_golden("collate")
    mean, std = [float(v) for v in g["mean"]], [float(v) for v in g["std"]]
    data = W.collate_case(*COLLATE_CASES[name])
    want = (g[name + "_imgs"], g[name + "_boxes"], g[name + "_classes"])
    got = P.DeviceCollate(mean, std, DEV)(data)
    assert tuple(got[0].shape) == want[0].shape and got[1].dtype == torch.float32 and got[2].dtype == torch.int64
    assert np.array_equal(to_np(got[0]), want[0]), "normalised, padded images are not bit-exact"
    assert np.array_equal(to_np(got[1]), want[1])
    assert np.array_equal(to_np(got[2]), want[2])
    # device-resident ragged lists take the same kernel
    got2 = P.pack_gt([d[1].to(DEV) for d in data], [d[2].to(DEV) for d in data], DEV)
    assert torch.equal(got2[0], got[1]) and torch.equal(got2[1], got[2])
    # the packed batch feeds the assignment directly
    if want[1].shape[1]:
        a = ops.assign_targets(W.VOC_LEVELS, W.STRIDES, W.FCOS_RANGES, got[1], got[2])
        b = O.assign_targets(W.VOC_LEVELS, torch.from_numpy(want[1]), torch.from_numpy(want[2]), W.STRIDES, W.FCOS_RANGES)
        assert_equal_int(to_np(a[0]), to_np(b[0]))


def test_device_collate_refuses_dataloader_workers():
    """A collate_fn runs inside the DataLoader's worker processes, where CUDA cannot be used: DeviceCollate says so
    instead of dying with 'Cannot re-initialize CUDA in forked subprocess'."""
    import torch.utils.data as tud
    data = W.collate_case(5, [(8, 8), (8, 8)], [1, 2])
    loader = tud.DataLoader(data, batch_size=2, num_workers=1, collate_fn=P.DeviceCollate([0.5] * 3, [0.25] * 3, DEV))
    with pytest.raises(Exception, match="num_workers=0"):
        next(iter(loader))
    main = tud.DataLoader(data, batch_size=2, num_workers=0, collate_fn=P.DeviceCollate([0.5] * 3, [0.25] * 3, DEV))
    imgs, boxes, classes = next(iter(main))
    assert imgs.is_cuda and boxes.shape == (2, 2, 4) and classes.shape == (2, 2)


def test_pack_gt_empty_batch():
    b, c = P.pack_gt([torch.zeros(0, 4), torch.zeros(0, 4)], [torch.zeros(0, dtype=torch.int64)] * 2, DEV)
    assert b.shape == (2, 0, 4) and c.shape == (2, 0)


# ------------------------------------------------------------------------------------------
# N3: VOC average precision on the device (test.py:15-162)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", EVAL_CASES)
def test_eval_ap_matches_reference_golden(name):
    """sort_by_score + eval_ap_2d through the CUDA kernels == the reference's numpy loops (golden)."""
    lists, num_cls, thr, want = load_eval_case(name)
    sb, sl, ss = P.sort_by_score(lists["det_boxes"], lists["det_labels"], lists["det_scores"], DEV)
    got = P.eval_ap_2d(lists["gt_boxes"], lists["gt_labels"], sb, sl, ss, thr, num_cls, DEV)
    assert sorted(got) == list(range(1, num_cls))
    assert_ap_equal([got[c] for c in range(1, num_cls)], want, what=name)


def test_eval_ap_on_detect_outputs_large_class_and_empty():
    """The batched entry on FCOSHead.detect() outputs against the oracle; one class with > 1024 detections
    (several scan chunks, a sort larger than a CTA); images without detections / without ground truth."""
    from oracle import eval_oracle as E
    x = W.head_outputs(6, 3, W.VOC_LEVELS, seed=123)
    head = P.FCOSHead(0.05, 0.6, 1000, W.STRIDES)
    s, c, b, n = head.detect(cuda_levels(x), clip_hw=W.VOC_HW)
    gt, labels = W.gt_boxes(6, 40, W.VOC_HW, 3, seed=124)
    labels[4] = -1                                   # an image without ground truth
    n = n.clone(); n[5] = 0                          # an image without detections
    ap = P.eval_ap_batched(s, c, b, n, gt.to(DEV), labels.to(DEV), 0.5, 4)
    cnt = to_np(n)
    want = E.eval_ap([to_np(gt[i])[to_np(labels[i]) > 0] for i in range(6)],
                     [to_np(labels[i])[to_np(labels[i]) > 0] for i in range(6)],
                     [to_np(b[i, :cnt[i]]) for i in range(6)], [to_np(c[i, :cnt[i]]) for i in range(6)],
                     [to_np(s[i, :cnt[i]]) for i in range(6)], 0.5, 4)
    assert int((c[:5] == 1).sum()) > 1024
    assert float(ap[0]) == 0.0
    assert_ap_equal(to_np(ap)[1:], [want[k] for k in (1, 2, 3)])


def test_coco_results_match_reference_golden():
    """coco_results against the rows the reference's own export loop produced (Test_coco.py:144-168, executed by
    tests/golden/make_golden_r2.py) — stage-wise on IDENTICAL inputs: the reference's FCOSHead + ClipBoxes detections
    stored in the fixture go up to the device, the rows that come back are the reference's rows exactly (the box
    arithmetic is fp32: divide by the scale, then w = x2 - x1, h = y2 - y1 in place; cut at the first score below the
    threshold)."""
    g = load_golden("coco_export")
    ncls, thr = int(g["meta"][1]), float(g["meta"][3])
    s = torch.from_numpy(g["det_scores"]).to(DEV)
    c = torch.from_numpy(g["det_classes"]).to(DEV)
    b = torch.from_numpy(g["det_boxes"]).to(DEV)
    n = torch.from_numpy(g["det_counts"]).to(DEV)
    ids, id2cat = [int(v) for v in g["ids"]], {k: 100 + k for k in range(1, ncls + 1)}
    got = P.coco_results(s, c, b, n, torch.tensor(g["scales"], dtype=torch.float32, device=DEV), ids, id2cat, threshold=thr)
    assert len(got) == len(g["score"]) > 0
    assert [r["image_id"] for r in got] == [int(v) for v in g["image_id"]]
    assert [r["category_id"] for r in got] == [int(v) for v in g["category_id"]]
    assert np.array_equal(np.array([r["score"] for r in got]), g["score"])
    assert np.array_equal(np.array([r["bbox"] for r in got]), g["bbox"]), "bbox rows are not bit-exact"
    # and end to end on our own detections of the same inputs: same number of rows per image
    x = W.head_outputs(int(g["meta"][0]), ncls, W.COCO_LEVELS, seed=int(g["meta"][2]))
    head = P.FCOSHead(0.05, 0.6, 1000, W.STRIDES)
    s2, c2, b2, n2 = head.detect(cuda_levels(x), clip_hw=W.COCO_HW)
    mine = P.coco_results(s2, c2, b2, n2, torch.tensor(g["scales"], dtype=torch.float32, device=DEV), ids, id2cat, threshold=thr)
    assert [sum(r["image_id"] == i for r in mine) for i in ids] == [int((g["image_id"] == i).sum()) for i in ids]


# ------------------------------------------------------------------------------------------
# full-size configs: properties that do not need the oracle at scale + oracle spot checks
# ------------------------------------------------------------------------------------------
def test_full_size_config2_postprocess_properties_and_oracle():
    """BASELINE config 2: COCO 832x1344, 80 classes, batch 16."""
    x = W.head_outputs(16, 80, W.COCO_LEVELS, seed=202)
    head = P.FCOSHead(0.05, 0.6, 1000, W.STRIDES)
    xc = cuda_levels(x)
    s, c, b, n = head.detect(xc)
    want = O.detect(x, 0.05, 0.6, 1000, W.STRIDES)
    for i in range(16):
        m = int(n[i])
        assert_detections_match((to_np(s[i, :m]), to_np(c[i, :m]), to_np(b[i, :m])),
                                tuple(to_np(t) for t in want[i]), rel=REL_TOL, what=f"img {i}")
    # idempotence: NMS of the survivors keeps everything, in the same order
    s2, c2, b2, k2, n2 = ops.batched_nms(b, s, c, 0.05, 0.6, n)
    assert torch.equal(n2, n)
    for i in range(16):
        m = int(n[i])
        assert torch.equal(k2[i, :m], torch.arange(m, device=DEV))
        assert torch.equal(b2[i, :m], b[i, :m])
    # per-image independence: image 5 alone gives the same answer as inside the batch
    s1, c1, b1, n1 = head.detect([[t[5:6] for t in part] for part in xc])
    m = int(n[5])
    assert int(n1[0]) == m and torch.equal(s1[0, :m], s[5, :m]) and torch.equal(b1[0, :m], b[5, :m])


def test_full_size_config4_dense_crowd():
    """BASELINE config 4: 5 x 1000 candidates per image (vanilla per-class branch), IoU 0.6."""
    boxes, scores, classes = zip(*[W.crowd_candidates(5000, 80, seed=400 + i) for i in range(4)])
    boxes, scores, classes = torch.stack(boxes), torch.stack(scores), torch.stack(classes)
    s, c, b, k, n = ops.batched_nms(boxes.to(DEV), scores.to(DEV), classes.to(DEV), 0.05, 0.6)
    for i in range(4):
        want = O.post_process_image(scores[i], classes[i], boxes[i], 0.05, 0.6)
        m = int(n[i])
        assert m == want[0].numel()
        assert_equal_int(to_np(k[i, :m]), to_np(want[3]), what=f"keep img {i}")
    # 300 GT boxes per image for assignment (dense-crowd training side)
    gt, labels = W.gt_boxes(8, 300, W.COCO_HW, 80, seed=404)
    want = O.assign_targets(W.COCO_LEVELS, gt, labels, W.STRIDES, W.HISFCOS_RANGES)
    got = ops.assign_targets(W.COCO_LEVELS, W.STRIDES, W.HISFCOS_RANGES, gt.to(DEV), labels.to(DEV), want_index=True)
    assert_equal_int(to_np(got[3]), to_np(want[3]), what="gt index")
    assert np.array_equal(to_np(got[2]), to_np(want[2]))


# ------------------------------------------------------------------------------------------
# fused K2+K3 kernel (class-bucketed sparse mask) against the dense three-kernel path
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ncls,crowded,seed,max_box", [(3, True, 51, 1000), (80, True, 52, 1000), (1, True, 53, 1000),
                                                       (20, False, 54, 1024), (5, True, 55, 300)])
def test_fused_postprocess_equals_dense_stage_chain(ncls, crowded, seed, max_box, monkeypatch):
    """Crowded boxes (many same-class overlaps) and large boxes around the top-left corner (negative
    x1, y1: the cross-class 'wildcard' case of the coordinate trick).  The fused kernel must give
    exactly what K1 -> K2 -> dense NMS (stand-alone entry, dense mask + scan kernels forced) and the oracle give."""
    monkeypatch.setenv("B200DET_DENSE_NMS", "1")
    x = W.head_outputs(2, ncls, W.VOC_LEVELS, seed=seed, crowded=crowded)
    for lv in range(len(x[2])):
        x[2][lv][:, :2, :3, :3] += 300.0           # l, t huge near the corner -> x1, y1 << -1 (wildcards)
        x[0][lv][:, :, :3, :3] += 3.0              # and make sure those points are selected
    xc = cuda_levels(x)
    head = P.FCOSHead(0.05, 0.6, max_box, W.STRIDES)
    s, c, b, n = head.detect(xc)
    score, cls0 = ops.score_points(xc[0], xc[1], W.STRIDES)
    ks, kc, kb, kp, kn = ops.select_topk(xc[2], W.STRIDES, score, cls0, 0.05, max_box)
    ds, dc, db, dk, dn = ops.batched_nms(kb, ks, kc.long(), 0.05, 0.6, kn)
    assert torch.equal(n, dn)
    want = O.detect(x, 0.05, 0.6, max_box, W.STRIDES)
    for i in range(2):
        m = int(n[i])
        assert torch.equal(s[i, :m], ds[i, :m]) and torch.equal(c[i, :m], dc[i, :m]) and torch.equal(b[i, :m], db[i, :m])
        assert_detections_match((to_np(s[i, :m]), to_np(c[i, :m]), to_np(b[i, :m])),
                                tuple(to_np(t) for t in want[i]), rel=REL_TOL, what=f"img {i}")
    assert int((kb[..., 0] < -1).logical_and(kb[..., 1] < -1).sum()) > 0     # wildcards were present


# ------------------------------------------------------------------------------------------
# input formats and the less-travelled paths of the head
# ------------------------------------------------------------------------------------------
def same_detections(got, ref):
    """padded outputs: rows beyond counts[b] are unspecified, compare the valid prefix only"""
    if not torch.equal(got[3], ref[3]):
        return False
    for i, n in enumerate(ref[3].tolist()):
        if not all(torch.equal(a[i, :n], b[i, :n]) for a, b in zip(got[:3], ref[:3])):
            return False
    return True


def test_head_accepts_half_and_channels_last_inputs():
    x = W.head_outputs(2, 20, W.VOC_LEVELS, seed=61)
    xc = cuda_levels(x)
    head = P.FCOSHead(0.05, 0.6, 1000, W.STRIDES)
    ref = head.detect(xc)
    # channels_last (what cuDNN convolutions often return): same values, different strides
    xl = [[t.contiguous(memory_format=torch.channels_last) for t in part] for part in xc]
    assert not xl[0][0].is_contiguous()
    got = head.detect(xl)
    assert same_detections(got, ref)
    # fp16 head outputs (AMP): evaluated in fp32 on the up-cast values
    xh = [[t.half() for t in part] for part in xc]
    got = head.detect(xh)
    want = head.detect([[t.float() for t in part] for part in xh])
    assert same_detections(got, want)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("levels,ncls", [(W.VOC_LEVELS, 20), (W.COCO_LEVELS[1:], 7)])
def test_postprocess_reads_half_maps_natively(dtype, levels, ncls):
    """K1 / K2 read the fp16 / bf16 head outputs of an autocast forward (train.py:175) as they are: bit-identical to
    the fp32 kernels on the up-cast maps — with every map half, with fp32 distances (autocast runs exp in fp32) and
    on levels whose planes are not 16-byte aligned (13 x 21) — and no up-cast copy is made."""
    strides = W.STRIDES[:len(levels)]
    x = W.head_outputs(3, ncls, levels, seed=62)
    x[2][0][0, :, :2, :3] = 60000.0 if dtype == torch.float16 else 3.0e38       # extreme but finite distances
    xh = [[t.to(dtype).to(DEV) for t in part] for part in x]
    up = [[t.float() for t in part] for part in xh]
    head = P.FCOSHead(0.05, 0.6, 1000, strides)
    want = head.detect(up)
    assert same_detections(head.detect(xh), want)                            # cls, cnt, reg all half
    assert same_detections(head.detect([xh[0], xh[1], up[2]]), want)          # half logits, fp32 distances
    assert same_detections(head.detect([up[0], up[1], xh[2]]), want)          # fp32 logits, half distances
    assert same_detections(head.detect([xh[0], up[1], xh[2]]), want)          # cls / cnt disagree: up-cast fallback
    s_h, c_h = ops.score_points(xh[0], xh[1], strides)
    s_f, c_f = ops.score_points(up[0], up[1], strides)
    assert torch.equal(s_h, s_f) and torch.equal(c_h, c_f)
    # the level table points at the caller's half tensors themselves
    lv, keep, *_ = ops._levels(xh[0], xh[1], xh[2], strides, native_half=True)
    assert lv[0].dtypes == ops._DTYPE_CODE[dtype] | (ops._DTYPE_CODE[dtype] << 4)
    assert keep[0].data_ptr() == xh[0][0].data_ptr() and keep[2].data_ptr() == xh[2][0].data_ptr()
    # entry points without half kernels refuse a declared half map instead of mis-reading it
    assert _lib.load().b200det_box_loss_fwd(lv, len(levels), 3, 0, 0, 1, 0, 0, None) == 2


@pytest.mark.parametrize("max_box,thr,nms_thr", [(2000, 0.05, 0.6), (1500, 0.0, 0.5), (1000, 0.05, -0.5),
                                                 (10000, 0.05, 0.6), (1, 0.05, 0.6)])
def test_head_large_k_and_odd_thresholds_match_oracle(max_box, thr, nms_thr):
    """k > 1024 takes the three-kernel chain (per-class branch of batched_nms above 1000 candidates);
    nms_thr < 0 (a zero IoU suppresses) takes the dense mask; k > P selects every point."""
    x = W.head_outputs(2, 20, W.VOC_LEVELS, seed=62, crowded=True)
    head = P.FCOSHead(thr, nms_thr, max_box, W.STRIDES)
    s, c, b, n = head.detect(cuda_levels(x))
    want = O.detect(x, thr, nms_thr, max_box, W.STRIDES)
    for i in range(2):
        m = int(n[i])
        assert_detections_match((to_np(s[i, :m]), to_np(c[i, :m]), to_np(b[i, :m])),
                                tuple(to_np(t) for t in want[i]), rel=REL_TOL, what=f"k={max_box} img {i}")


def test_single_level_and_tiny_maps():
    x = W.head_outputs(3, 7, [(5, 3)], seed=63)
    head = P.FCOSHead(0.05, 0.6, 1000, [16])
    s, c, b, n = head.detect(cuda_levels(x))
    want = O.detect(x, 0.05, 0.6, 1000, [16])
    for i in range(3):
        m = int(n[i])
        assert_detections_match((to_np(s[i, :m]), to_np(c[i, :m]), to_np(b[i, :m])),
                                tuple(to_np(t) for t in want[i]), rel=REL_TOL, what=f"img {i}")
    gt = torch.tensor([[[4.0, 4.0, 40.0, 70.0], [-1, -1, -1, -1]]] * 3)
    lab = torch.tensor([[3, -1]] * 3)
    got = ops.assign_targets([(5, 3)], [16], [[-1, 64]], gt.to(DEV), lab.to(DEV), want_index=True)
    ref = O.assign_targets([(5, 3)], gt, lab, [16], [[-1, 64]])
    assert_equal_int(to_np(got[0]), to_np(ref[0]), what="cls_t")
    assert np.array_equal(to_np(got[2]), to_np(ref[2]))
    assert_equal_int(to_np(got[3]), to_np(ref[3]), what="gt index")


def test_assign_large_batch_tile_shape():
    """B*P above the small/large tile switch (<128 threads x 8 points> kernel)."""
    gt, labels = W.gt_boxes(80, 40, W.COCO_HW, 80, seed=64)
    got = ops.assign_targets(W.COCO_LEVELS, W.STRIDES, W.HISFCOS_RANGES, gt.to(DEV), labels.to(DEV), want_index=True)
    sel = [0, 17, 79]
    ref = O.assign_targets(W.COCO_LEVELS, gt[sel], labels[sel], W.STRIDES, W.HISFCOS_RANGES)
    for j, i in enumerate(sel):
        assert_equal_int(to_np(got[0][i]), to_np(ref[0][j]), what="cls_t")
        assert np.array_equal(to_np(got[2][i]), to_np(ref[2][j]))
        assert_equal_int(to_np(got[3][i]), to_np(ref[3][j]), what="gt index")


# ------------------------------------------------------------------------------------------
# multi-GPU gather through peer memory (one rank here; bench.py under torchrun runs it at 2 / 8 GPUs)
# ------------------------------------------------------------------------------------------
def test_peer_gather_single_rank_roundtrip():
    """sharding.PeerGather on a one-rank group: symmetric-memory rendezvous, the copy-engine push into the
    rank's own slot, the device-side barrier and slot rotation.  Skipped where symmetric memory is unavailable
    (bench.py then falls back to the NCCL all_gather of sharding.gather_packed)."""
    import socket

    import torch.distributed as dist
    from pytorch_object_detection_b200.sharding import PeerGather

    created = False
    if not dist.is_initialized():
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                                device_id=torch.device(DEV))
        created = True
    try:
        try:
            pg = PeerGather(4096, 2, torch.device(DEV))
        except Exception as e:                                 # noqa: BLE001
            pytest.skip(f"symmetric memory unavailable: {type(e).__name__}: {e}")
        gen = torch.Generator(device=DEV).manual_seed(3)
        for slot, root in ((0, None), (1, 0), (0, 0)):
            src = torch.randint(0, 256, (4096,), dtype=torch.uint8, device=DEV, generator=gen)
            view = pg.gather(slot, src, root=root)
            torch.cuda.synchronize()
            assert view.shape == (1, 4096) and torch.equal(view[0], src)
        with pytest.raises(ValueError):
            pg.gather(0, torch.zeros(8, dtype=torch.uint8, device=DEV))
    finally:
        if created:
            dist.destroy_process_group()
