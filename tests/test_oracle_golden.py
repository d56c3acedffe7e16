"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/*.npz,
made by tests/golden/make_golden.py) and against the installed torchvision CPU NMS."""
import numpy as np
import pytest
import torch

from oracle import fcos_oracle as O
from pytorch_object_detection_b200 import workloads as W
from helpers import assert_close, assert_detections_match, assert_equal_int, head_inputs_from_meta, load_golden

HEAD_CASES = {
    "head_voc_b1": (W.VOC_LEVELS, W.VOC_HW),
    "head_voc_4strides": (W.VOC_LEVELS, W.VOC_HW),
    "head_coco_b2": (W.COCO_LEVELS, W.COCO_HW),
    "head_coco_crowded": (W.COCO_LEVELS, W.COCO_HW),
    "head_voc_k300": (W.VOC_LEVELS, W.VOC_HW),
    "head_voc_saturated": (W.VOC_LEVELS, W.VOC_HW),      # class logits that collapse in the fp32 sigmoid (first-index rule)
}


def head_inputs(g, levels):
    x, batch, max_box, strides = head_inputs_from_meta(g, levels)
    assert_close(W.fingerprint(x[0] + x[1] + x[2]), g["fingerprint"], rel=1e-12, what="input fingerprint")
    return x, batch, max_box, strides


@pytest.mark.parametrize("name", sorted(HEAD_CASES))
def test_head_oracle_matches_reference(name):
    levels, img_hw = HEAD_CASES[name]
    g = load_golden(name)
    x, batch, max_box, strides = head_inputs(g, levels)
    dets = O.detect(x, 0.05, 0.6, max_box, strides)
    score, classes, boxes = O.score_points(x, strides)
    s_k, c_k, b_k, _ = O.select_topk(score, classes, boxes, max_box)
    for b in range(batch):
        assert_detections_match((s_k[b], c_k[b], b_k[b]),
                                (g[f"topk_score_{b}"], g[f"topk_class_{b}"], g[f"topk_box_{b}"]),
                                rel=0.0, what=f"{name} top-k img {b}")
        assert_detections_match(dets[b], (g[f"score_{b}"], g[f"class_{b}"], g[f"box_{b}"]),
                                rel=0.0, what=f"{name} detections img {b}")
        clipped = O.clip_boxes_(dets[b][2].clone()[None], *img_hw)[0]
        assert_close(clipped, g[f"clipped_{b}"], rel=0.0, what=f"{name} clip img {b}")


def test_nms_oracle_matches_torchvision_golden():
    g = load_golden("nms_cases")
    names = sorted({k.rsplit("_", 1)[0] for k in g.files if k.endswith("_keep")})
    assert "crowd5000" in names and "neg_cross_class" in names
    for n in names:
        boxes = torch.from_numpy(g[n + "_boxes"])
        scores = torch.from_numpy(g[n + "_scores"])
        classes = torch.from_numpy(g[n + "_classes"]).long()
        keep = O.batched_nms(boxes, scores, classes, float(g[n + "_thr"])).numpy()
        want = g[n + "_keep"]
        if n == "crowd5000" or n == "crowd1001":   # vanilla branch: unstable final sort -> tie-aware
            assert np.array_equal(np.sort(keep), np.sort(want)), n
            assert_close(scores.numpy()[keep], scores.numpy()[want], rel=0.0, what=n)
        else:
            assert_equal_int(keep, want, what=n)
    # semantics the survey calls out
    assert g["neg_cross_class_keep"].tolist() == [0]
    assert g["iou_exact_0p6_keep"].tolist() == [0]
    assert g["equal_scores_keep"].tolist() == [0, 1, 3]


def test_nms_oracle_matches_installed_torchvision_live():
    import torchvision
    for seed, n, ncls in [(1, 300, 4), (2, 999, 80), (3, 1000, 20), (4, 1500, 10)]:
        boxes, scores, classes = W.crowd_candidates(n, ncls, seed=seed, clusters=8)
        boxes[::7] -= 600.0  # negative coordinates
        want = torchvision.ops.batched_nms(boxes, scores, classes, 0.6).numpy()
        got = O.batched_nms(boxes, scores, classes, 0.6).numpy()
        assert_equal_int(got, want, what=f"n={n}")


TRAIN_CASES = {
    "train_voc_b2": (W.VOC_LEVELS, W.VOC_HW),
    "train_coco_b2": (W.COCO_LEVELS, W.COCO_HW),
    "train_voc_dense": (W.VOC_LEVELS, W.VOC_HW),
}


def train_inputs(g, levels, img_hw):
    batch, ncls, seed, max_gt = (int(v) for v in g["meta"][:4])
    gt, labels = W.gt_boxes(batch, max_gt, img_hw, ncls, seed)
    x = W.head_outputs(batch, ncls, levels, seed + 1)
    assert_close(W.fingerprint([gt, labels.float()] + x[0] + x[1] + x[2]), g["fingerprint"], rel=1e-12,
                 what="input fingerprint")
    return x, gt, labels, g["ranges"].tolist()


@pytest.mark.parametrize("name", sorted(TRAIN_CASES))
def test_assign_and_loss_oracle_match_reference(name):
    levels, img_hw = TRAIN_CASES[name]
    g = load_golden(name)
    x, gt, labels, ranges = train_inputs(g, levels, img_hw)
    cls_t, cnt_t, reg_t, _ = O.assign_targets(levels, gt, labels, W.STRIDES, ranges)
    assert cls_t.dtype == torch.int64 and cls_t.shape[-1] == 1
    assert_equal_int(cls_t, g["cls_t"], what="cls_t")
    assert np.array_equal(cnt_t.numpy(), g["cnt_t"]), "cnt_t not bit-exact"
    assert np.array_equal(reg_t.numpy(), g["reg_t"]), "reg_t not bit-exact"
    assert (cnt_t > -1).sum() > 0
    for mode in ("giou", "iou"):
        for part in x:
            for t in part:
                t.grad = None
                t.requires_grad_(True)
        losses = O.fcos_loss(x, (cls_t, cnt_t, reg_t), mode)
        assert_close([float(v) for v in losses], g[f"loss_{mode}"], rel=1e-6, what=f"loss {mode}")
        losses[3].backward()
        for lv in range(len(levels)):
            assert_close(x[2][lv].grad, g[f"g_reg_{mode}_{lv}"], rel=1e-5, abs_=1e-9, what="g_reg")
            assert_close(x[1][lv].grad, g[f"g_cnt_{mode}_{lv}"], rel=1e-5, abs_=1e-9, what="g_cnt")
            if f"g_cls_{mode}_{lv}" in g.files:
                assert_close(x[0][lv].grad, g[f"g_cls_{mode}_{lv}"], rel=1e-5, abs_=1e-12, what="g_cls")


def test_known_answer_cnt_loss():
    """model/loss.py:219-221 prints tensor([0.3133, 0.3133]) = log(1 + e^-1)."""
    got = O.cnt_loss([torch.ones(2, 1, 4, 4)] * 5, torch.ones(2, 80, 1), torch.ones(2, 80, dtype=torch.bool))
    assert_close(got, load_golden("known_answers")["cnt_loss_ones"], rel=1e-6)
    assert_close(got, [np.log1p(np.exp(-1.0))] * 2, rel=1e-6)
    assert [round(float(v), 4) for v in got] == [0.3133, 0.3133]


def test_analytic_box_losses():
    same = torch.tensor([[3.0, 4.0, 5.0, 6.0]])
    assert float(O.giou_sum(same, same)) == pytest.approx(0.0, abs=1e-7)
    inner = torch.tensor([[5.0, 5.0, 5.0, 5.0]])     # 10x10 nested in 20x20, same centre
    outer = torch.tensor([[10.0, 10.0, 10.0, 10.0]])
    assert float(O.giou_sum(inner, outer)) == pytest.approx(0.75, rel=1e-6)
    assert float(O.iou_sum(inner, outer)) == pytest.approx(-np.log(0.25), rel=1e-6)
    with pytest.raises(NotImplementedError):
        O.reg_loss([torch.ones(1, 4, 2, 2)], torch.ones(1, 4, 4), torch.ones(1, 4, dtype=torch.bool), mode="diou")


# ------------------------------------------------------------------------------------------
# N3: the evaluation oracle against the reference's own eval_ap_2d (tests/golden/eval_ap.npz)
# ------------------------------------------------------------------------------------------
from helpers import EVAL_CASES, assert_ap_equal, load_eval_case  # noqa: E402
from oracle import eval_oracle as E  # noqa: E402


@pytest.mark.parametrize("name", EVAL_CASES)
def test_eval_oracle_matches_reference_ap(name):
    lists, num_cls, thr, want = load_eval_case(name)
    order = [np.argsort(-s, kind="stable") for s in lists["det_scores"]]            # sort_by_score (no ties in the cases)
    got = E.eval_ap(lists["gt_boxes"], lists["gt_labels"], [b[o] for b, o in zip(lists["det_boxes"], order)],
                    [l[o] for l, o in zip(lists["det_labels"], order)],
                    [s[o] for s, o in zip(lists["det_scores"], order)], thr, num_cls)
    assert_ap_equal([got[c] for c in range(1, num_cls)], want, what=name)


def test_focal_step_gradient_identity_matches_reference_autograd():
    """The focal step kernel writes the gradient of a non-target element as 0.75 om pr (om - 2 pt log pt) with
    pt = fl(1 - p), om = fl(1 - pt) (csrc/loss.cu, focal_neg_both_nb2): no reciprocal, every intermediate
    shared with the loss.  Restated here in fp32 numpy with exact transcendentals and checked against autograd
    of the reference's formula (loss.py:180-193) over logits in [-30, 14]: the identity itself is exact to fp32
    rounding, what remains for the GPU test is the accuracy of MUFU.EX2 / RCP / LG2."""
    f = np.float32
    x = np.linspace(-30, 14, 200001).astype(f)
    xt = torch.tensor(x, requires_grad=True)
    p = torch.sigmoid(xt).clip(min=0.000005, max=0.99999999995)
    pt = 1.0 - p                                                          # y = 0: pt = 1 - p, w = 0.75
    loss = -0.75 * torch.pow(1.0 - pt, 2.0) * torch.log(pt)
    loss.sum().backward()
    want_g = xt.grad.numpy().astype(np.float64)
    want_l = loss.detach().numpy().astype(np.float64)
    pr = torch.sigmoid(torch.tensor(x)).numpy()
    pc = np.maximum(pr, f(0.000005))
    ptn = (f(1) - pc).astype(f)
    om = (f(1) - ptn).astype(f)
    lg = np.log(ptn.astype(np.float64)).astype(f)
    t = (om.astype(np.float64) - 2.0 * lg.astype(np.float64) * ptn.astype(np.float64)).astype(f)   # one fma
    g = ((f(0.75) * (om * pr).astype(f)).astype(f) * t).astype(f)
    g = np.where(pr >= f(0.000005), g, f(0))
    l = ((f(-0.75) * (om * om).astype(f)).astype(f) * lg).astype(f)
    big = np.abs(want_g) > 1e-12
    assert np.max(np.abs(g[big] - want_g[big]) / np.abs(want_g[big])) < 2e-6
    assert np.all(g[~big & (pr < f(0.000005))] == 0)                      # clip: no gradient below 5e-6
    ok = np.abs(want_l) > 1e-30
    assert np.max(np.abs(l[ok] - want_l[ok]) / np.abs(want_l[ok])) < 2e-6
    # the series the kernel uses for log(pt) where om <= 1/16: -om * (1 + om/2 + ... + om^6/7)
    u = np.linspace(0, 0.0625, 10001)
    series = -u * sum(u ** k / (k + 1) for k in range(7))
    exact = np.log1p(-u)
    assert np.max(np.abs(series[1:] - exact[1:]) / np.abs(exact[1:])) < 2e-9
