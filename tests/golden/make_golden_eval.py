"""Generate tests/golden/eval_ap.npz by running the UNMODIFIED reference's eval_ap_2d / sort_by_score.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    CUDA_VISIBLE_DEVICES="" python tests/golden/make_golden_eval.py

``test.py`` of the reference imports its model zoo (which needs packages that are not installed), so the four
pure-numpy functions this path consists of — ``sort_by_score``, ``iou_2d``, ``_compute_ap``, ``eval_ap_2d``
(test.py:15-162) — are compiled from the reference's own source file, untouched, with ``ast`` and executed here.
Inputs are stored (they are small); ragged lists are stored flat with offsets.
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B200DET_REFERENCE", "/root/reference")
WANT = {"sort_by_score", "iou_2d", "_compute_ap", "eval_ap_2d"}

tree = ast.parse(open(os.path.join(REF, "test.py")).read())
tree.body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANT]
ns = {"np": np}
exec(compile(tree, os.path.join(REF, "test.py"), "exec"), ns)
assert WANT <= set(ns)


def make_case(seed, n_img, num_cls, max_gt, max_det, img=300.0, jitter=2.5, dup=0.3):
    """Detections = jittered copies of GT boxes (some duplicated, some with a wrong class) + random boxes."""
    rng = np.random.default_rng(seed)
    gt_boxes, gt_labels, det_boxes, det_labels, det_scores = [], [], [], [], []
    for _ in range(n_img):
        g = int(rng.integers(0, max_gt + 1))
        xy = rng.uniform(0, img * 0.7, size=(g, 2))
        wh = rng.uniform(10, img * 0.3, size=(g, 2))
        gb = np.concatenate([xy, xy + wh], axis=1).astype(np.float32)
        gl = rng.integers(1, num_cls, size=g)
        boxes, labels = [], []
        for b, l in zip(gb, gl):
            for _r in range(1 + int(rng.random() < dup) + int(rng.random() < dup / 2)):
                boxes.append(b + rng.normal(0, jitter, size=4))
                labels.append(l if rng.random() > 0.1 else int(rng.integers(1, num_cls)))
        for _r in range(int(rng.integers(0, max(1, max_det - len(boxes))))):
            xy = rng.uniform(0, img * 0.7, size=2)
            boxes.append(np.concatenate([xy, xy + rng.uniform(10, img * 0.3, size=2)]))
            labels.append(int(rng.integers(1, num_cls)))
        boxes = np.asarray(boxes, dtype=np.float32).reshape(-1, 4)[:max_det]
        labels = np.asarray(labels, dtype=np.int64)[:max_det]
        scores = rng.permutation(np.linspace(0.05, 0.99, 4096))[:len(boxes)].astype(np.float32) \
            + np.float32(1e-4) * np.float32(len(det_boxes))       # distinct scores: no sort ties anywhere
        gt_boxes.append(gb)
        gt_labels.append(gl.astype(np.int64))
        det_boxes.append(boxes)
        det_labels.append(labels)
        det_scores.append(scores)
    return gt_boxes, gt_labels, det_boxes, det_labels, det_scores


def flat(lists, width=None):
    off = np.cumsum([0] + [len(x) for x in lists]).astype(np.int64)
    cat = np.concatenate([np.asarray(x).reshape(len(x), width) if width else np.asarray(x) for x in lists]) \
        if sum(len(x) for x in lists) else np.zeros((0, width) if width else (0,))
    return cat, off


out = {}
CASES = {"voc_like": (11, 40, 21, 6, 60, 0.5), "coco_like": (12, 24, 81, 12, 100, 0.5),
         "strict_iou": (13, 30, 5, 8, 50, 0.75), "sparse": (14, 12, 21, 1, 4, 0.5)}
for name, (seed, n_img, num_cls, max_gt, max_det, thr) in CASES.items():
    gb, gl, db, dl, ds = make_case(seed, n_img, num_cls, max_gt, max_det)
    sb, sl, ss = ns["sort_by_score"](db, dl, ds)                     # the reference sorts, then evaluates
    ap = ns["eval_ap_2d"](gb, gl, sb, sl, ss, thr, num_cls)
    out[name + "_meta"] = np.array([seed, n_img, num_cls, max_gt, max_det, thr])
    out[name + "_ap"] = np.array([ap[c] for c in range(1, num_cls)], dtype=np.float64)
    for key, lists, width in (("gt_boxes", gb, 4), ("gt_labels", gl, None), ("det_boxes", db, 4),
                              ("det_labels", dl, None), ("det_scores", ds, None)):
        cat, off = flat(lists, width)
        out[f"{name}_{key}"] = cat
        out[f"{name}_{key}_off"] = off
    print(name, "mAP", np.nanmean(out[name + "_ap"]), "NaN classes", int(np.isnan(out[name + "_ap"]).sum()))
np.savez_compressed(os.path.join(HERE, "eval_ap.npz"), **out)
print("eval_ap.npz", os.path.getsize(os.path.join(HERE, "eval_ap.npz")) / 1024, "KiB")
