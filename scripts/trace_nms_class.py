"""Phase timing of nms_class_kernel on the dense-crowd case (needs a B200DET_TRACE=1 build)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_object_detection_b200 import _lib, ops, workloads as W
lib = _lib.load()
crowd = [W.crowd_candidates(5000, 80, seed=400 + i) for i in range(8)]
cb = torch.stack([c[0] for c in crowd]).cuda(); cs = torch.stack([c[1] for c in crowd]).cuda(); cc = torch.stack([c[2] for c in crowd]).cuda()
for _ in range(3):
    ops.batched_nms(cb, cs, cc, 0.05, 0.6)
torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
fn = lib.b200det_debug_read_trace_nmsclass
fn.argtypes = [C.c_void_p, C.c_int]
fn(buf, 64)
t = list(buf)
names = ["keys", "sort", "segments", "class warps", "write"]
print(" ".join(f"{n} +{(t[i + 1] - t[i]) / 1965:.1f}us" for i, n in enumerate(names)), "segments", t[8])
if t[11]:
    print(f"first batch of CTA 0 ({t[12]} tiles of {t[13]} classes): done +{(t[11] - t[3]) / 1965:.1f}us; all batches +{(t[14] - t[3]) / 1965:.1f}us")
if t[21]:
    d = lambda a, b: (t[b] - t[a]) / 1965
    print(f"  class split: tables cleared +{d(0, 21):.1f}us, class ids loaded +{d(21, 22):.1f}us, counted +{d(22, 23):.1f}us, barriers +{d(23, 1):.1f}us; "
          f"setup: dealing order +{d(3, 24):.1f}us, batch tables + clear +{d(24, 25):.1f}us, boxes staged +{d(25, 26):.1f}us")
if t[15]:
    rel = lambda i: (t[i] - t[3]) / 1965
    print(f"  CTA 0: first unit taken at {rel(20):.1f}us, last unit taken at {rel(19):.1f}us; greedy pass of the largest class {rel(15):.1f} -> {rel(16):.1f}us, "
          f"of the smallest {rel(17):.1f} -> {rel(18):.1f}us")
fn = lib.b200det_debug_read_trace_nms
fn.argtypes = [C.c_void_p, C.c_int]
fn(buf, 64)
t = list(buf)
names = ["threshold + rank", "bin starts", "placement", "rank in bin", "gather", "nms boxes"]
print("nms_prepare_kernel, image 0: " + " ".join(f"{n} +{(t[31 + i] - t[30 + i]) / 1965:.1f}us" for i, n in enumerate(names)))

# the dense path (coordinate-trick images): 1 000 crowded candidates x 16 images
crowd = [W.crowd_candidates(1000, 80, seed=500 + i) for i in range(16)]
kb = torch.stack([c[0] for c in crowd]).cuda(); ks = torch.stack([c[1] for c in crowd]).cuda(); kc = torch.stack([c[2] for c in crowd]).cuda()
for _ in range(3):
    ops.batched_nms(kb, ks, kc, 0.05, 0.6)
torch.cuda.synchronize()
fn(buf, 64)
t = list(buf)
print("dense path, image 0: nms_prepare_kernel " + " ".join(f"{n} +{(t[31 + i] - t[30 + i]) / 1965:.1f}us" for i, n in enumerate(names)))
print("  scan kernel: " + " ".join(f"{n} +{(t[17 + i] - t[16 + i]) / 1965:.1f}us" for i, n in enumerate(["rows staged", "greedy pass", "outputs"])))
fn2 = lib.b200det_debug_read_trace_fused
fn2.argtypes = [C.c_void_p, C.c_int]
fn2(buf, 64)
t = list(buf)
if t[8]:
    seq = [(8, "set loaded"), (12, "zero-fill + stage"), (13, "buckets"), (9, "pair tests"), (10, "greedy pass"), (11, "outputs")]
    print("  bucket path (NMS half of the fused kernel), image 0: " + " ".join(f"{nm} +{(t[b] - t[a]) / 1965:.1f}us" for (a, _), (b, nm) in zip(seq, seq[1:])))
