"""Random shapes through the per-class NMS kernel (> 1000 candidates) against the CPU oracle: class-size mixes that
exercise one / several batches per CTA, early and late greedy passes, ordered and round-robin dealing.

    python scripts/fuzz_nms_class.py [--cases N] [--seed S]      (run it under `timeout`)
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import fcos_oracle as O  # noqa: E402
from pytorch_object_detection_b200 import ops, workloads as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cases", type=int, default=60)
ap.add_argument("--seed", type=int, default=0)
args = ap.parse_args()
rng = np.random.default_rng(args.seed)
bad = 0
t0 = time.time()
for case in range(args.cases):
    n = int(rng.integers(1001, 8193))
    kind = int(rng.integers(0, 5))
    if kind == 0:      # a few big classes (several 64-box blocks each, early greedy passes)
        ncls = int(rng.integers(2, 12))
        classes = torch.from_numpy(rng.integers(1, ncls + 1, n))
    elif kind == 1:    # geometric sizes: one class of hundreds, a tail of tiny ones
        ncls = int(rng.integers(20, 200))
        p = 0.5 ** np.arange(ncls) + 1e-3
        classes = torch.from_numpy(rng.choice(ncls, n, p=p / p.sum()) + 1)
    elif kind == 2:    # many tiny classes with large ids (counting sort / round-robin dealing, several batches)
        classes = torch.from_numpy(rng.integers(300, 300 + int(rng.integers(300, 4000)), n))
    elif kind == 3:    # sizes right at the block borders
        sizes = []
        while sum(sizes) < n:
            sizes.append(int(rng.choice([1, 63, 64, 65, 127, 128, 129, 192, 193, 500, 1024])))
        classes = torch.cat([torch.full((m,), i + 1) for i, m in enumerate(sizes)])[:n]
        classes = classes[torch.from_numpy(rng.permutation(n))]
    else:              # the crowd generator's own classes
        classes = None
    clusters = int(rng.integers(3, 80))
    boxes, scores, cc = W.crowd_candidates(n, 80, seed=1000 + case, clusters=clusters, spread=float(rng.uniform(5, 60)))
    if classes is None:
        classes = cc
    if rng.random() < 0.3:
        scores = torch.round(scores * 50) / 50          # many equal scores
    thr = float(rng.choice([0.6, 0.5, 0.3, 0.75, 0.0, -1.0]))
    sizes = torch.bincount(classes.long())
    if int(sizes.max()) > 1024 and rng.random() < 0.7:
        continue                                        # (dense fallback: covered by the tests, slow on the oracle)
    want = O.batched_nms(boxes, scores, classes, thr).numpy()
    s, c, b, k, cnt = ops.batched_nms(boxes[None].cuda(), scores[None].cuda(), classes[None].long().cuda(), -1e30, thr)
    torch.cuda.synchronize()
    got = k[0, : int(cnt[0])].cpu().numpy()
    ok = got.shape == want.shape and np.array_equal(got, want)
    if not ok:   # torch's final sort of the kept set is unstable on equal scores: compare as sets, then the scores
        ok = np.array_equal(np.sort(got), np.sort(want)) and np.array_equal(scores.numpy()[got], scores.numpy()[want])
    bad += 0 if ok else 1
    print(f"case {case:3d} n={n:5d} kind={kind} classes={int((sizes > 0).sum()):5d} largest={int(sizes.max()):5d} thr={thr:5.2f} "
          f"kept={want.size:5d} {'ok' if ok else 'MISMATCH'}  ({time.time() - t0:.0f}s)", flush=True)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
