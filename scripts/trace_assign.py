"""Phase timing of assign_targets_kernel (needs B200DET_TRACE=1 build)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_object_detection_b200 import _lib, ops, workloads as W
lib = _lib.load()
gt, labels = W.gt_boxes(32, 100, W.COCO_HW, 80, seed=3000)
gt, labels = gt.cuda(), labels.cuda()
for _ in range(3):
    ops.assign_targets(W.COCO_LEVELS, W.STRIDES, W.HISFCOS_RANGES, gt, labels)
torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
fn = lib.b200det_debug_read_trace_assign
fn.argtypes = [C.c_void_p, C.c_int]
fn(buf, 64)
t = list(buf)
for name, o in (("level-0 tile 0", 0), ("coarsest tile", 8)):
    print(f"{name}: list built +{(t[o+1]-t[o])/1965:.2f} us, points done +{(t[o+2]-t[o+1])/1965:.2f} us, n_list={t[o+3]}")
