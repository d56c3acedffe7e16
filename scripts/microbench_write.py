"""Floor for write-only kernels: torch fill / zero_ of the assign output size, and assign with no GT."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_object_detection_b200 import ops, workloads as W

def timed(fn, per=8, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(per): fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / per)
    return sorted(ts)[n // 2]

for mb in (5.2, 20.9, 83.6, 334.0):
    buf = torch.empty(int(mb * 1e6) // 4, dtype=torch.float32, device="cuda")
    t = timed(lambda: buf.fill_(1.0))
    print(f"fill_ {mb:6.1f} MB: {t:7.2f} us  {mb*1e6/t/1e6:7.1f} GB/s")
src = torch.empty(int(20.9e6) // 4, dtype=torch.float32, device="cuda"); dst = torch.empty_like(src)
t = timed(lambda: dst.copy_(src)); print(f"copy_ 20.9 MB: {t:7.2f} us  {2*20.9e6/t/1e6:7.1f} GB/s (r+w)")
for B, M in ((32, 100), (32, 1), (128, 100)):
    gt, labels = W.gt_boxes(B, M, W.COCO_HW, 80, seed=3000)
    gt, labels = gt.cuda(), labels.cuda()
    t = timed(lambda: ops.assign_targets(W.COCO_LEVELS, W.STRIDES, W.HISFCOS_RANGES, gt, labels))
    nbytes = B * 23265 * 28
    print(f"assign B={B} M={M}: {t:7.2f} us  {nbytes/t/1e6:7.1f} GB/s")
