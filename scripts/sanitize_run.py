"""Small end-to-end pass of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_object_detection_b200 as P
from pytorch_object_detection_b200 import ops, workloads as W

dev = "cuda:0"
x = [[t.to(dev).requires_grad_(True) for t in part] for part in W.head_outputs(2, 80, W.COCO_LEVELS, seed=5)]
head = P.FCOSHead(0.05, 0.6, 1000, W.STRIDES)
with torch.no_grad():
    s, c, b, n = head.detect(x, clip_hw=W.COCO_HW)
    s2 = P.FCOSHead(0.05, 0.6, 3000, W.STRIDES).detect(x)            # three-kernel chain, ring scan
    bb, ss, cc = W.crowd_candidates(5000, 80, seed=1)
    ops.batched_nms(bb[None].to(dev), ss[None].to(dev), cc[None].to(dev), 0.05, 0.6)
gt, labels = W.gt_boxes(2, 100, W.COCO_HW, 80, seed=6)
tgt = P.FCOSGenTargets(W.STRIDES, W.HISFCOS_RANGES)([x, gt.to(dev), labels.to(dev)])
for mode in ("giou", "iou"):
    losses = P.FCOSLoss(mode)([x, tgt])
    losses[3].backward()
torch.cuda.synchronize()
print("ok", n.tolist(), float(losses[3]))
