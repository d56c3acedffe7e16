import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import fcos_oracle as O
from pytorch_object_detection_b200 import workloads as W, ops
from helpers import load_golden
g = load_golden("train_coco_b2")
batch, ncls, seed, max_gt = (int(v) for v in g["meta"][:4])
gt, labels = W.gt_boxes(batch, max_gt, W.COCO_HW, ncls, seed)
got = ops.assign_targets(W.COCO_LEVELS, W.STRIDES, g["ranges"].tolist(), gt.cuda(), labels.cuda(), want_index=True)
want = O.assign_targets(W.COCO_LEVELS, gt, labels, W.STRIDES, g["ranges"].tolist())
cnt = got[1].cpu().numpy(); wc = g["cnt_t"]
bad = np.argwhere(cnt != wc)
print("mismatches", len(bad))
for b, p, _ in bad[:10]:
    print(b, p, cnt[b, p, 0].view(np.uint32) if hasattr(cnt[b,p,0],'view') else 0, repr(cnt[b, p, 0]), repr(wc[b, p, 0]),
          "idx", int(got[3][b, p]), int(want[3][b, p]), "reg", got[2][b, p].cpu().numpy(), g["reg_t"][b, p])
    r = torch.from_numpy(g["reg_t"][b, p])
    lrmin, lrmax = torch.min(r[0], r[2]), torch.max(r[0], r[2]); tbmin, tbmax = torch.min(r[1], r[3]), torch.max(r[1], r[3])
    print("  cpu recompute", repr(float(((lrmin * tbmin) / (lrmax * tbmax + 1e-10)).sqrt())), repr(float((lrmin*tbmin)/(lrmax*tbmax+1e-10))))
