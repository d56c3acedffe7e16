"""Phase timing of assign_stream_kernel (needs a B200DET_TRACE=1 build: B200DET_LIB=scratch_libs/libtrace.so).

    B200DET_TRACE=1 python -c "from pytorch_object_detection_b200 import build; build.build(lib='scratch_libs/libtrace.so')"
    B200DET_LIB=scratch_libs/libtrace.so python scripts/trace_stream.py [fused|assign]
"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_object_detection_b200 import _lib, ops, workloads as W
lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "fused"
B = int(os.environ.get("B", 32))
gt, labels = W.gt_boxes(B, 100, W.COCO_HW, 80, seed=3000)
gt, labels = gt.cuda(), labels.cuda()
reg = [torch.exp(torch.randn(B, 4, h, w, device="cuda") + 3) for h, w in W.COCO_LEVELS]
cnt = [torch.randn(B, 1, h, w, device="cuda") for h, w in W.COCO_LEVELS]
name = "train" if which == "fused" else "assign"
rd = getattr(lib, f"b200det_debug_read_trace_{name}")
rd.argtypes = [C.c_void_p, C.c_int]
rs = getattr(lib, f"b200det_debug_reset_trace_{name}")


def run():
    if which == "fused":
        ops.assign_loss_fused(reg, cnt, W.STRIDES, W.HISFCOS_RANGES, gt, labels, 1)
    else:
        ops.assign_targets(W.COCO_LEVELS, W.STRIDES, W.HISFCOS_RANGES, gt, labels)


for _ in range(3):
    run()
torch.cuda.synchronize()
names = ["init+GT issue", "stage", "vote", "scan+arrive", "loss eval", "num_pos", "join wait", "patch"]
for rep in range(3):
    rs()
    run()
    torch.cuda.synchronize()
    buf = (C.c_longlong * 64)()
    rd(buf, 64)
    t = list(buf)
    print(f"rep {rep}: kernel span (first CTA start -> last CTA end) {(t[61] - t[60]) / 1e3:.2f} us")
    for who, o in (("level-0 tile 0", 0), ("coarsest tile", 16)):
        ph = []
        prev = t[o]
        for i, nm in enumerate(names):
            cur = t[o + i + 1]
            if cur and prev:
                ph.append(f"{nm} +{(cur - prev) / 1965:.2f}")
                prev = cur
        fill = f"fill warp: start +{(t[o + 10] - t[o]) / 1965:.2f} issued +{(t[o + 11] - t[o + 10]) / 1965:.2f} landed +{(t[o + 9] - t[o + 11]) / 1965:.2f}"
        print(f"  {who}: {' '.join(ph)} us | {fill} | total {(t[o + 8] - t[o]) / 1965:.2f} us | n_pos={t[o + 12]} np={t[o + 13]} n_list={t[o + 14]}")

# per-CTA start / end (ns after the first CTA's start) of the last run
import numpy as np
rc = getattr(lib, f"b200det_debug_read_cta_trace_{name}")
rc.argtypes = [C.c_void_p, C.c_int]
n_cta = min(4096, B * int(os.environ.get("TILES", 18)))
cb = (C.c_longlong * 8192)()
rc(cb, 8192)
arr = np.array(list(cb), dtype=np.int64).reshape(-1, 2)[:n_cta]
arr = arr[arr[:, 0] > 0]
t0_ = arr[:, 0].min()
st, en = (arr[:, 0] - t0_) / 1e3, (arr[:, 1] - t0_) / 1e3
print(f"{len(arr)} CTAs: start min/median/p90/max {st.min():.2f}/{np.median(st):.2f}/{np.percentile(st, 90):.2f}/{st.max():.2f} us; "
      f"end min/median/p90/max {en.min():.2f}/{np.median(en):.2f}/{np.percentile(en, 90):.2f}/{en.max():.2f} us; "
      f"duration median/max {np.median(en - st):.2f}/{(en - st).max():.2f} us")
late = np.argsort(-en)[:8]
print("latest CTAs (index, image, x, start, end):", [(int(i), int(i) // 18, int(i) % 18, round(float(st[i]), 2), round(float(en[i]), 2)) for i in late])
