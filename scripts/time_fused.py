"""Time the fused target/loss step (config 3) against the separate kernels.  B200DET_FUSED_CFG=<cluster>x<threads>."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pytorch_object_detection_b200 as P
from pytorch_object_detection_b200 import ops, workloads as W

dev = "cuda:0"
B = int(os.environ.get("B", 32))
M = int(os.environ.get("M", 100))
gt, labels = W.gt_boxes(B, M, W.COCO_HW, 80, seed=3000)
gt, labels = gt.to(dev), labels.to(dev)
sets = [([torch.exp(torch.randn(B, 4, h, w, device=dev) + 3).requires_grad_(True) for h, w in W.COCO_LEVELS],
         [torch.randn(B, 1, h, w, device=dev).requires_grad_(True) for h, w in W.COCO_LEVELS]) for _ in range(4)]


def timed(fn, reps=50):
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


def kernel_only(i, with_cnt):
    reg, cnt = sets[i % 4]
    ops.assign_loss_fused(reg, cnt if with_cnt else None, W.STRIDES, W.HISFCOS_RANGES, gt, labels, 1)


step = P.FCOSTargetLoss(W.STRIDES, W.HISFCOS_RANGES, "giou")


def module_step(i, with_cnt):
    reg, cnt = sets[i % 4]
    for t in reg + cnt:
        t.grad = None
    a, c = step.box_cnt_losses(cnt if with_cnt else None, reg, gt, labels)
    (a + c if with_cnt else a).backward()


gen = P.FCOSGenTargets(W.STRIDES, W.HISFCOS_RANGES)
fake = [torch.empty(B, 1, h, w, device="meta") for h, w in W.COCO_LEVELS]


def unfused_step(i):
    reg, cnt = sets[i % 4]
    for t in reg:
        t.grad = None
    tgt = gen([[fake, fake, fake], gt, labels])
    P.compute_reg_loss(reg, tgt[2], None, "giou", _mask_src=tgt[1]).mean().backward()


P_ = W.num_points(W.COCO_LEVELS)
print(f"cfg={os.environ.get('B200DET_FUSED_CFG', 'default')} B={B} M={M}")
for with_cnt in (False, True):
    us = timed(lambda i: kernel_only(i, with_cnt))
    by = B * P_ * (28 + 16 + (4 if with_cnt else 0))
    print(f"  fused kernel only  cnt={with_cnt}: {us:7.2f} us  {by / us / 1e3:7.1f} GB/s written")
    print(f"  fused module fwd+bwd cnt={with_cnt}: {timed(lambda i: module_step(i, with_cnt)):7.2f} us")
print(f"  unfused assign + GIoU fwd + bwd: {timed(unfused_step):7.2f} us")
