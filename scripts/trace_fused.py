"""Phase timing of assign_loss_fused_kernel (needs a B200DET_TRACE=1 build)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_object_detection_b200 import _lib, ops, workloads as W
lib = _lib.load()
B = int(os.environ.get("B", 32))
gt, labels = W.gt_boxes(B, 100, W.COCO_HW, 80, seed=3000)
gt, labels = gt.cuda(), labels.cuda()
reg = [torch.exp(torch.randn(B, 4, h, w, device="cuda") + 3) for h, w in W.COCO_LEVELS]
cnt = [torch.randn(B, 1, h, w, device="cuda") for h, w in W.COCO_LEVELS]
for _ in range(3):
    ops.assign_loss_fused(reg, cnt, W.STRIDES, W.HISFCOS_RANGES, gt, labels, 1)
torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
if not hasattr(lib, "b200det_debug_read_trace_train"):
    print("ran 3 fused steps (no B200DET_TRACE build: no phase stamps)")
    sys.exit(0)
fn = lib.b200det_debug_read_trace_train
fn.argtypes = [C.c_void_p, C.c_int]
fn(buf, 64)
t = list(buf)
names = ["stage", "vote", "pass A", "wait+pass B", "ticket"]
for who, o in (("level-0 tile 0", 0), ("coarsest tile", 16)):
    print(who, " ".join(f"{n} +{(t[o + i + 1] - t[o + i]) / 1965:.2f}us" for i, n in enumerate(names)))
