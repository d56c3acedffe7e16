"""Run every b200det kernel a few times at BASELINE sizes (for ncu / timing).

    python scripts/profile_kernels.py [--reps N] [--time]

Config 2 inputs (COCO 832x1344, 80 classes, batch 16) for the post-process kernels, config 3
(B=32, M<=100) for assignment and the losses.  With --time, prints CUDA-event timings per entry
point (inputs rotated over 4 sets so the 127 MB batch does not sit in the 126 MB L2).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pytorch_object_detection_b200 as B  # noqa: E402
from pytorch_object_detection_b200 import ops, workloads as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--time", action="store_true")
ap.add_argument("--sets", type=int, default=4)
ap.add_argument("--only", default="")
args = ap.parse_args()
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)


def head_set(batch, ncls):
    cls, cnt, reg = [], [], []
    for h, w in W.COCO_LEVELS:
        cls.append(torch.randn(batch, ncls, h, w, device=dev, generator=g) - 4.595)
        cnt.append(torch.randn(batch, 1, h, w, device=dev, generator=g))
        reg.append(torch.exp(torch.randn(batch, 4, h, w, device=dev, generator=g) + 3.0))
    return cls, cnt, reg


sets = [head_set(16, 80) for _ in range(args.sets)]
gt, labels = W.gt_boxes(32, 100, W.COCO_HW, 80, seed=3000)
gt, labels = gt.to(dev), labels.to(dev)
tsets = [head_set(32, 80) for _ in range(2)]
head = B.FCOSHead(0.05, 0.6, 1000, W.STRIDES)
tgt = ops.assign_targets(W.COCO_LEVELS, W.STRIDES, W.HISFCOS_RANGES, gt, labels)
gl = torch.full((32,), 1.0 / 32, device=dev)
score, cls0 = ops.score_points(sets[0][0], sets[0][1], W.STRIDES)
cand = ops.select_topk(sets[0][2], W.STRIDES, score, cls0, 0.05, 1000)
npos = ops.box_loss_fwd(tsets[0][2], tgt[1], tgt[2], 1)[1]

tsets_h = [[t.half() for t in ts[0]] for ts in tsets]              # autocast: fp16 class logits
scale_state = torch.tensor([65536.0, 0.0], device=dev)             # {assumed loss scale, ticket}

crowd = [W.crowd_candidates(5000, 80, seed=400 + i) for i in range(8)]
cb = torch.stack([c[0] for c in crowd]).to(dev)
cs = torch.stack([c[1] for c in crowd]).to(dev)
cc = torch.stack([c[2] for c in crowd]).to(dev)
crowd1k = [W.crowd_candidates(1000, 80, seed=4100 + i) for i in range(16)]
kb = torch.stack([c[0] for c in crowd1k]).to(dev)
ks = torch.stack([c[1] for c in crowd1k]).to(dev)
kc = torch.stack([c[2] for c in crowd1k]).to(dev)
cases = {
    "nms_crowd1000_b16": lambda i: ops.batched_nms(kb, ks, kc, 0.05, 0.6),
    "nms_crowd5000_b8": lambda i: ops.batched_nms(cb, cs, cc, 0.05, 0.6),
    "score_points": lambda i: ops.score_points(sets[i % args.sets][0], sets[i % args.sets][1], W.STRIDES),
    "select_topk": lambda i: ops.select_topk(sets[0][2], W.STRIDES, score, cls0, 0.05, 1000),
    "batched_nms": lambda i: ops.batched_nms(cand[2], cand[0], cand[1].long(), 0.05, 0.6, cand[4]),
    "postprocess": lambda i: head.detect(sets[i % args.sets], clip_hw=W.COCO_HW),
    "assign_targets": lambda i: ops.assign_targets(W.COCO_LEVELS, W.STRIDES, W.HISFCOS_RANGES, gt, labels),
    "assign_loss_fused": lambda i: ops.assign_loss_fused(tsets[i % 2][2], None, W.STRIDES, W.HISFCOS_RANGES, gt, labels, 1),
    "assign_loss_fused_cnt": lambda i: ops.assign_loss_fused(tsets[i % 2][2], tsets[i % 2][1], W.STRIDES, W.HISFCOS_RANGES,
                                                             gt, labels, 1),
    "box_loss_fwd": lambda i: ops.box_loss_fwd(tsets[i % 2][2], tgt[1], tgt[2], 1),
    "box_loss_bwd": lambda i: ops.box_loss_bwd(tsets[i % 2][2], tgt[1], tgt[2], 1, gl, npos),
    "cnt_loss_fwd": lambda i: ops.cnt_loss_fwd(tsets[i % 2][1], tgt[1], tgt[1]),
    "cnt_loss_bwd": lambda i: ops.cnt_loss_bwd(tsets[i % 2][1], tgt[1], tgt[1], gl, npos),
    "cls_loss_fwd": lambda i: ops.cls_loss_fwd(tsets[i % 2][0], tgt[1], tgt[0]),
    "cls_loss_bwd": lambda i: ops.cls_loss_bwd(tsets[i % 2][0], tgt[0], gl, npos),
    "cls_loss_step": lambda i: ops.cls_loss_step(tsets[i % 2][0], tgt[0], num_pos=npos),
    "cls_loss_step_f16": lambda i: ops.cls_loss_step(tsets_h[i % 2], tgt[0], num_pos=npos, up_mean=scale_state),
}
only = [s for s in args.only.split(",") if s]
for name, fn in cases.items():
    if only and name not in only:
        continue
    for i in range(args.reps):
        fn(i)
    torch.cuda.synchronize()
    if args.time:
        # a CUDA graph of `per` back-to-back calls removes the Python/ctypes launch overhead
        per, n = 8, 20
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(per):
                fn(i)
        graph.replay()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in evs:
            a.record()
            graph.replay()
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) * 1e3 / per for a, b in evs)
        print(f"{name:16s} median {ts[n // 2]:8.1f} us   min {ts[0]:8.1f} us   (per call, graph of {per})")
print("done")
