"""Comparators shared by the parity tests.

* integers (keep indices, labels, GT indices): bit-exact.
* floats: relative 1e-5 (the tolerance BASELINE.json's north_star states), written here.
* top-k / equal-score order: tie-aware — torch.topk and torch.sort(stable=False) order
  equal scores arbitrarily on CPU, so equal-score runs are compared as sets.
"""
import os

import numpy as np
import torch

REL_TOL = 1e-5
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def assert_close(a, b, rel=REL_TOL, abs_=0.0, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b)
    lim = rel * np.maximum(np.abs(a), np.abs(b)) + abs_
    bad = err > lim
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} beyond rel {rel}; worst {err.max():.3e}"


def assert_equal_int(a, b, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert np.array_equal(a.astype(np.int64), b.astype(np.int64)), \
        f"{what}: {(a.astype(np.int64) != b.astype(np.int64)).sum()} mismatches of {a.size}"


def canon_rows(scores, *cols):
    """Sort rows by (score desc, then the remaining columns lexicographically) so that rows
    with equal scores compare as a set.  Returns one [n, k] float64 matrix."""
    mats = [np.asarray(scores, dtype=np.float64).reshape(len(scores), -1)]
    for c in cols:
        mats.append(np.asarray(c, dtype=np.float64).reshape(len(scores), -1))
    m = np.concatenate(mats, axis=1)
    keys = [m[:, j] for j in range(m.shape[1] - 1, 0, -1)] + [-m[:, 0]]
    return m[np.lexsort(keys)]


def assert_detections_match(got, want, rel=REL_TOL, what=""):
    """got / want: (scores[n], classes[n], boxes[n,4]).  Same count, descending scores, and the
    same rows up to a permutation inside equal-score runs; floats to `rel`, classes exact."""
    gs, gc, gb = (np.asarray(x) for x in got)
    ws, wc, wb = (np.asarray(x) for x in want)
    assert gs.shape == ws.shape, f"{what}: kept {gs.shape[0]} vs {ws.shape[0]}"
    assert np.all(np.diff(gs.astype(np.float64)) <= 0), f"{what}: scores not descending"
    assert_close(gs, ws, rel, what=what + " scores")
    if np.array_equal(gc.astype(np.int64), wc.astype(np.int64)):
        assert_close(gb, wb, rel, what=what + " boxes")
        return
    g = canon_rows(gs, gc, gb)
    w = canon_rows(ws, wc, wb)
    assert np.array_equal(g[:, 1], w[:, 1]), f"{what}: classes differ beyond tie permutation"
    assert_close(g[:, 2:], w[:, 2:], rel, what=what + " boxes (tie-canonical)")


def to_np(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def load_eval_case(name):
    """Ragged lists + reference APs of one case of tests/golden/eval_ap.npz (made by make_golden_eval.py)."""
    g = load_golden("eval_ap")

    def ragged(key):
        cat, off = g[f"{name}_{key}"], g[f"{name}_{key}_off"]
        return [cat[off[i]:off[i + 1]] for i in range(len(off) - 1)]

    meta = g[name + "_meta"]
    lists = {k: ragged(k) for k in ("gt_boxes", "gt_labels", "det_boxes", "det_labels", "det_scores")}
    lists["gt_boxes"] = [np.asarray(x, dtype=np.float32).reshape(-1, 4) for x in lists["gt_boxes"]]
    lists["det_boxes"] = [np.asarray(x, dtype=np.float32).reshape(-1, 4) for x in lists["det_boxes"]]
    lists["gt_labels"] = [np.asarray(x, dtype=np.int64) for x in lists["gt_labels"]]
    lists["det_labels"] = [np.asarray(x, dtype=np.int64) for x in lists["det_labels"]]
    lists["det_scores"] = [np.asarray(x, dtype=np.float32) for x in lists["det_scores"]]
    return lists, int(meta[2]), float(meta[5]), g[name + "_ap"]


EVAL_CASES = ["voc_like", "coco_like", "strict_iou", "sparse"]


def assert_ap_equal(got, want, what=""):
    """APs are fp64 sums of a few hundred terms: equal to 1e-12, NaN where the reference gives NaN."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: {got.shape} vs {want.shape}"
    assert np.array_equal(np.isnan(got), np.isnan(want)), f"{what}: NaN classes differ"
    ok = ~np.isnan(want)
    assert np.all(np.abs(got[ok] - want[ok]) <= 1e-12), f"{what}: worst {np.abs(got[ok] - want[ok]).max():.3e}"


def head_inputs_from_meta(g, levels):
    """Regenerate a head golden's inputs from its meta row (batch, classes, seed, max_box, flavour): flavour 1 =
    crowded boxes, 2 = saturated class logits (make_golden_r2.py)."""
    from pytorch_object_detection_b200 import workloads as W
    batch, ncls, seed, max_box, flavour = (int(v) for v in g["meta"][:5])
    x = W.head_outputs(batch, ncls, levels, seed, crowded=(flavour == 1))
    if flavour == 2:
        x = W.saturate_logits(x, seed + 1)
    return x, batch, max_box, [int(s) for s in g["strides"]]
