"""b200det_assign_targets / b200det_assign_loss_fused at several batch sizes (multi-wave grids)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_object_detection_b200 import ops, workloads as W

dev = "cuda:0"
P = W.num_points(W.COCO_LEVELS)


def timed(fn, reps=30):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(4):
            fn(i)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / (4 * reps)


for B in (16, 32, 64, 128, 256):
    gt, labels = W.gt_boxes(B, 100, W.COCO_HW, 80, seed=3000)
    gt, labels = gt.to(dev), labels.to(dev)
    regs = [[torch.exp(torch.randn(B, 4, h, w, device=dev) + 3) for h, w in W.COCO_LEVELS] for _ in range(2)]
    ua = timed(lambda i: ops.assign_targets(W.COCO_LEVELS, W.STRIDES, W.HISFCOS_RANGES, gt, labels))
    uf = timed(lambda i: ops.assign_loss_fused(regs[i % 2], None, W.STRIDES, W.HISFCOS_RANGES, gt, labels, 1))
    print(f"B={B:4d}  assign {ua:7.2f} us = {B * P * 28 / ua / 1e3:7.1f} GB/s   fused {uf:7.2f} us = {B * P * 44 / uf / 1e3:7.1f} GB/s written")
