"""Generate tests/golden/*.npz by running the UNMODIFIED reference on seeded inputs.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    CUDA_VISIBLE_DEVICES="" python tests/golden/make_golden.py

* ``head.py`` / ``loss.py`` are imported from /root/reference with a stub ``torchinfo``
  (``utill/utills.py:5`` imports it; it is not installed) — SURVEY.md §8(c).
* ``FCOSHead`` is called once per image (its final ``torch.stack`` raises for ragged
  batches, ``head.py:99-101``).
* NMS cases call the installed torchvision 0.26.0 CPU ``batched_nms`` directly, which is
  what ``head.py:94`` calls.
Inputs are NOT stored: they are regenerated from the seed by
``pytorch_object_detection_b200.workloads`` and checked through a fingerprint.
"""
import os
import sys
import types

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("B200DET_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
stub = types.ModuleType("torchinfo")
stub.summary = lambda *a, **k: None
sys.modules["torchinfo"] = stub

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torchvision  # noqa: E402

from model.modules.head import FCOSHead, ClipBoxes, FCOSGenTargets  # noqa: E402  (reference)
from model.loss import FCOSLoss, compute_cnt_loss  # noqa: E402  (reference)
from pytorch_object_detection_b200 import workloads as W  # noqa: E402

assert not torch.cuda.is_available()
torch.manual_seed(0)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def run_head(name, levels, img_hw, num_classes, batch, seed, strides, max_box=1000, crowded=False):
    x = W.head_outputs(batch, num_classes, levels, seed, crowded=crowded)
    head = FCOSHead(0.05, 0.6, max_box, strides)
    out = {"fingerprint": W.fingerprint(x[0] + x[1] + x[2]),
           "meta": np.array([batch, num_classes, seed, max_box, int(crowded), img_hw[0], img_hw[1]]),
           "strides": np.array(strides)}
    # stage capture: run the reference's forward with post_process intercepted
    for b in range(batch):
        xb = [[t[b:b + 1] for t in part] for part in x]
        captured = {}
        orig = head.post_process

        def spy(preds, _orig=orig, _cap=captured):
            _cap["topk"] = [p.clone() for p in preds]
            return _orig(preds)

        head.post_process = spy
        s, c, bx = head(xb)
        head.post_process = orig
        out[f"topk_score_{b}"] = captured["topk"][0][0].numpy()
        out[f"topk_class_{b}"] = captured["topk"][1][0].numpy().astype(np.int32)
        out[f"topk_box_{b}"] = captured["topk"][2][0].numpy()
        out[f"score_{b}"] = s[0].numpy()
        out[f"class_{b}"] = c[0].numpy().astype(np.int32)
        out[f"box_{b}"] = bx[0].numpy().copy()
        imgs = torch.zeros(1, 3, img_hw[0], img_hw[1])
        clipped = ClipBoxes()(imgs, bx)
        assert clipped.data_ptr() == bx.data_ptr()
        out[f"clipped_{b}"] = clipped[0].numpy()
    save(name, **out)


def run_nms(name, cases):
    out = {}
    for key, (boxes, scores, classes, thr) in cases.items():
        keep = torchvision.ops.batched_nms(boxes, scores, classes, thr)
        out[key + "_boxes"] = boxes.numpy()
        out[key + "_scores"] = scores.numpy()
        out[key + "_classes"] = classes.numpy().astype(np.int32)
        out[key + "_thr"] = np.float64(thr)
        out[key + "_keep"] = keep.numpy().astype(np.int32)
    save(name, **out)


def run_train(name, levels, img_hw, num_classes, batch, max_gt, seed, ranges, store_grads):
    gt, labels = W.gt_boxes(batch, max_gt, img_hw, num_classes, seed)
    x = W.head_outputs(batch, num_classes, levels, seed + 1)
    for part in x:
        for t in part:
            t.requires_grad_(True)
    gen = FCOSGenTargets(W.STRIDES, ranges)
    tgt = gen([x, gt, labels])
    out = {"fingerprint": W.fingerprint([gt, labels.float()] + x[0] + x[1] + x[2]),
           "meta": np.array([batch, num_classes, seed, max_gt, img_hw[0], img_hw[1]]),
           "ranges": np.array(ranges, dtype=np.int64),
           "cls_t": tgt[0].numpy().astype(np.int32), "cnt_t": tgt[1].numpy(), "reg_t": tgt[2].numpy()}
    for mode in ("giou", "iou"):
        for part in x:
            for t in part:
                t.grad = None
        losses = FCOSLoss(mode)([x, tgt])
        losses[3].backward()
        out[f"loss_{mode}"] = np.array([float(v) for v in losses], dtype=np.float64)
        out[f"loss32_{mode}"] = np.array([v.detach().numpy() for v in losses], dtype=np.float32)
        # gradients: reg / cnt grads are sparse (positives only) and compress well; cls grads are
        # dense, stored only for the small configuration
        for lv in range(len(levels)):
            out[f"g_reg_{mode}_{lv}"] = x[2][lv].grad.numpy()
            out[f"g_cnt_{mode}_{lv}"] = x[1][lv].grad.numpy()
            if store_grads:
                out[f"g_cls_{mode}_{lv}"] = x[0][lv].grad.numpy()
            else:
                out[f"g_cls_sum_{mode}_{lv}"] = x[0][lv].grad.double().sum(dim=(2, 3)).numpy()
    save(name, **out)


def main():
    # --- inference head -------------------------------------------------------------
    run_head("head_voc_b1", W.VOC_LEVELS, W.VOC_HW, 20, 1, seed=11, strides=W.STRIDES)            # config 1
    run_head("head_voc_4strides", W.VOC_LEVELS, W.VOC_HW, 20, 1, seed=12, strides=W.STRIDES[:4])  # zip truncation
    run_head("head_coco_b2", W.COCO_LEVELS, W.COCO_HW, 80, 2, seed=13, strides=W.STRIDES)         # config 2 @ b2
    run_head("head_coco_crowded", W.COCO_LEVELS, W.COCO_HW, 80, 1, seed=14, strides=W.STRIDES, crowded=True)
    run_head("head_voc_k300", W.VOC_LEVELS, W.VOC_HW, 20, 1, seed=15, strides=W.STRIDES, max_box=300)

    # --- NMS stage (torchvision CPU) ------------------------------------------------
    cases = {}
    b, s, c = W.crowd_candidates(1000, 80, seed=21)
    cases["crowd1000"] = (b, s, c, 0.6)                       # coordinate-trick branch, unsorted input
    b, s, c = W.crowd_candidates(5000, 80, seed=22)
    cases["crowd5000"] = (b, s, c, 0.6)                       # vanilla per-class branch (config 4)
    b, s, c = W.crowd_candidates(1001, 80, seed=23)
    cases["crowd1001"] = (b, s, c, 0.6)                       # first size on the vanilla branch
    b, s, c = W.crowd_candidates(700, 3, seed=24, clusters=6)
    cases["crowd700_thr05"] = (b, s, c, 0.5)
    # negative coordinates: the trick suppresses ACROSS classes
    cases["neg_cross_class"] = (torch.tensor([[100., 100, 300, 300], [-201, -201, -1, -1]]),
                                torch.tensor([0.9, 0.8]), torch.tensor([1, 2]), 0.6)
    # IoU exactly 0.6f is suppressed at threshold 0.6 (fp32 IoU vs double threshold, §0.13)
    cases["iou_exact_0p6"] = (torch.tensor([[0., 0, 10, 10], [0., 0, 10, 6]]),
                              torch.tensor([0.9, 0.8]), torch.tensor([1, 1]), 0.6)
    # equal scores keep input order (stable sort, §0.14)
    cases["equal_scores"] = (torch.tensor([[0., 0, 10, 10], [20., 20, 30, 30], [0., 0, 10, 9], [40., 40, 50, 50]]),
                             torch.tensor([0.5, 0.5, 0.5, 0.5]), torch.tensor([1, 1, 1, 1]), 0.6)
    b, s, c = W.crowd_candidates(400, 5, seed=25, clusters=4)
    s = (s * 8).round() / 8                                     # many exact score ties
    cases["tied400"] = (b, s, c, 0.6)
    cases["single"] = (torch.tensor([[1., 2, 3, 4]]), torch.tensor([0.3]), torch.tensor([7]), 0.6)
    run_nms("nms_cases", cases)

    # --- training targets + losses --------------------------------------------------
    run_train("train_voc_b2", W.VOC_LEVELS, W.VOC_HW, 20, 2, 8, seed=31, ranges=W.FCOS_RANGES, store_grads=True)
    run_train("train_coco_b2", W.COCO_LEVELS, W.COCO_HW, 80, 2, 100, seed=32, ranges=W.HISFCOS_RANGES,
              store_grads=False)
    run_train("train_voc_dense", W.VOC_LEVELS, W.VOC_HW, 20, 1, 300, seed=33, ranges=W.HISFCOS_RANGES,
              store_grads=False)

    # --- the reference's own known answer (model/loss.py:219-221) --------------------
    ka = compute_cnt_loss([torch.ones([2, 1, 4, 4])] * 5, torch.ones([2, 80, 1]), torch.ones([2, 80], dtype=torch.bool))
    save("known_answers", cnt_loss_ones=ka.numpy())
    print("cnt known answer:", ka)


if __name__ == "__main__":
    main()
