"""Phase timing inside the single-CTA kernels (K2 select, K3 scan) from clock64 stamps.

Needs a library built with B200DET_TRACE=1:
    B200DET_TRACE=1 python -m pytorch_object_detection_b200.build --force
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pytorch_object_detection_b200 as B  # noqa: E402
from pytorch_object_detection_b200 import _lib, workloads as W  # noqa: E402

lib = _lib.load()
dev = "cuda:0"
x = [[t.to(dev) for t in part] for part in W.head_outputs(16, 80, W.COCO_LEVELS, seed=1)]
head = B.FCOSHead(0.05, 0.6, 1000, W.STRIDES)
for _ in range(3):
    head.detect(x, clip_hw=W.COCO_HW)
torch.cuda.synchronize()
t = [0] * 64
import os as _os
fused = _os.environ.get("B200DET_NO_FUSED") != "1"
for name, lo, hi in ((("fused", 0, 14),) if fused else (("select", 0, 8), ("nms", 16, 20))):
    buf = (C.c_longlong * 64)()
    fn = getattr(lib, "b200det_debug_read_trace_" + name)
    fn.argtypes = [C.c_void_p, C.c_int]
    fn(buf, 64)
    t[lo:hi] = list(buf)[lo:hi]
names = {0: "K2 start", 1: "K2 keys loaded + min/max", 2: "K2 phase A (block passes)", 3: "K2 phase B (warp finish)",
         4: "K2 tie + compaction", 5: "K2 sort", 6: "K2 gather/decode", 7: "K2 nms boxes",
         8: "F select done", 12: "F   zero-fill + stage", 13: "F   buckets built", 9: "F   pair tests", 10: "F greedy pass", 11: "F outputs",
         16: "K3 scan start", 17: "K3 first rows staged", 18: "K3 greedy pass", 19: "K3 outputs"}
order = [0, 1, 2, 3, 4, 5, 6, 7, 8, 12, 13, 9, 10, 11]
for grp in (((0, 14),) if fused else ((0, 8), (16, 20))):
    base = t[grp[0]]
    prev = base
    for i in (order if fused else range(grp[0], grp[1])):
        print(f"{names[i]:32s} +{(t[i] - prev) / 1.965e3:8.2f} us   (t = {(t[i] - base) / 1.965e3:8.2f} us)")
        prev = t[i]

if fused:
    buf = (C.c_longlong * 64)()
    fn = getattr(lib, "b200det_debug_read_trace_fused")
    fn.argtypes = [C.c_void_p, C.c_int]
    fn(buf, 64)
    tt = list(buf)
    if tt[21]:
        labels = ["keys loaded + counted", "scan", "threshold + slot ranges", "placement", "rank in bin"]
        pts = [tt[0], tt[21], tt[22], tt[23], tt[24], tt[1]]
        print("select, histogram path:", "  ".join(f"{l} +{(b - a) / 1965:.2f}" for l, a, b in zip(labels, pts, pts[1:])))
