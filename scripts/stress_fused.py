"""Stress of the cross-CTA protocol of the fused target / loss kernel (arrival counters, bounded wait, local recount):
the step runs on two streams at once (own workspaces) while a third stream keeps the SMs busy with K1, so tiles of one
image start far apart and both the counter path and the recount path are taken; every result must equal the quiet run
bit for bit, every time.  Also hammers back-to-back launches (programmatic dependent launches chained over calls).

    python scripts/stress_fused.py [iterations]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_object_detection_b200 import _lib, ops, workloads as W

dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
lib = _lib.load()
P = W.num_points(W.COCO_LEVELS)


def make(batch, m, seed):
    gt, labels = W.gt_boxes(batch, m, W.COCO_HW, 80, seed=seed)
    g = torch.Generator(device=dev).manual_seed(seed)
    reg = [torch.exp(torch.randn(batch, 4, h, w, device=dev, generator=g) + 3) for h, w in W.COCO_LEVELS]
    cnt = [torch.randn(batch, 1, h, w, device=dev, generator=g) for h, w in W.COCO_LEVELS]
    ws = torch.zeros(int(lib.b200det_assign_loss_workspace_bytes(batch, P)), dtype=torch.uint8, device=dev)
    return dict(gt=gt.to(dev), labels=labels.to(dev), reg=reg, cnt=cnt, ws=ws, batch=batch)


def run(c):
    return ops.assign_loss_fused(c["reg"], c["cnt"], W.STRIDES, W.HISFCOS_RANGES, c["gt"], c["labels"], 1, workspace=c["ws"])


def same(a, b):
    keys = ("cls_t", "cnt_t", "reg_t", "box_loss", "cnt_loss", "num_pos", "mean")
    return all(torch.equal(a[k], b[k]) for k in keys) and \
        all(torch.equal(x, y) for x, y in zip(a["reg_grads"] + a["cnt_grads"], b["reg_grads"] + b["cnt_grads"]))


cases = [make(32, 100, 1), make(70, 60, 2), make(7, 300, 3)]
for k, c in enumerate(cases):
    c["id"] = k
quiet = [run(c) for c in cases]
torch.cuda.synchronize()
noise = [[t.to(dev) for t in part] for part in W.head_outputs(16, 80, W.COCO_LEVELS, seed=9)]
streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
bad = 0
for it in range(iters):
    outs = []
    with torch.cuda.stream(streams[0]):
        for _ in range(1 + it % 3):
            ops.score_points(noise[0], noise[1], W.STRIDES)
    for k, c in enumerate(cases[:2] if it % 2 else cases[1:]):
        with torch.cuda.stream(streams[1 + k]):
            for _ in range(1 + it % 4):                      # back-to-back calls on one stream
                o = run(c)
            outs.append((c, o))
    torch.cuda.synchronize()
    for c, o in outs:
        if not same(o, quiet[c["id"]]):
            bad += 1
            print(f"iteration {it}: batch {c['batch']} differs from the quiet run")
print(f"{iters} iterations, {bad} mismatches")
sys.exit(1 if bad else 0)
