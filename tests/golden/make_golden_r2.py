"""Round-2 goldens, made by running the UNMODIFIED reference (build container only; see make_golden.py):

    CUDA_VISIBLE_DEVICES="" python tests/golden/make_golden_r2.py

* ``head_voc_saturated``  FCOSHead.forward on class logits that collapse in the fp32 sigmoid (workloads.saturate_logits):
                          pins torch.max's first-index rule over sigmoid(cls) (head.py:57-62).
* ``collate``             the datasets' ``collate_fn`` (dataset/voc.py:141-173; dataset/coco.py:135-165 is the same
                          code) compiled from the reference's source with ``ast`` (the module imports cv2 / xml readers
                          at import time, the method itself needs torch, numpy and torchvision.transforms only).
* ``coco_export``         the result loop of ``evaluate_coco`` (Test_coco.py:144-168): its statements, untouched, are
                          lifted out of the function body with ``ast`` and run on seeded detections.
"""
import ast
import os
import sys
import types

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("B200DET_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
stub = types.ModuleType("torchinfo")
stub.summary = lambda *a, **k: None
sys.modules["torchinfo"] = stub

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torchvision import transforms  # noqa: E402

from model.modules.head import FCOSHead, ClipBoxes  # noqa: E402  (reference)
from pytorch_object_detection_b200 import workloads as W  # noqa: E402

assert not torch.cuda.is_available()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


# ---- saturated logits -------------------------------------------------------------------------------------------
def run_head_saturated(name, levels, img_hw, num_classes, seed, strides, max_box=1000):
    x = W.saturate_logits(W.head_outputs(1, num_classes, levels, seed), seed + 1)
    head = FCOSHead(0.05, 0.6, max_box, strides)
    out = {"fingerprint": W.fingerprint(x[0] + x[1] + x[2]),
           "meta": np.array([1, num_classes, seed, max_box, 2, img_hw[0], img_hw[1]]),     # meta[4] = 2: saturated
           "strides": np.array(strides)}
    captured = {}
    orig = head.post_process

    def spy(preds):
        captured["topk"] = [p.clone() for p in preds]
        return orig(preds)

    head.post_process = spy
    s, c, bx = head(x)
    out["topk_score_0"] = captured["topk"][0][0].numpy()
    out["topk_class_0"] = captured["topk"][1][0].numpy().astype(np.int32)
    out["topk_box_0"] = captured["topk"][2][0].numpy()
    out["score_0"] = s[0].numpy()
    out["class_0"] = c[0].numpy().astype(np.int32)
    out["box_0"] = bx[0].numpy().copy()
    out["clipped_0"] = ClipBoxes()(torch.zeros(1, 3, *img_hw), bx)[0].numpy()
    # how many of the selected points would get another class from an argmax over the logits
    cls_flat = torch.cat([t.permute(0, 2, 3, 1).reshape(1, -1, num_classes) for t in x[0]], dim=1)[0]
    by_logit = cls_flat.argmax(dim=-1) + 1
    by_sigmoid = torch.sigmoid(cls_flat).max(dim=-1)[1] + 1
    out["points_where_logit_argmax_differs"] = int((by_logit != by_sigmoid).sum())
    print(f"  {int((by_logit != by_sigmoid).sum())} of {cls_flat.shape[0]} points: argmax(logit) != argmax(sigmoid)")
    save(name, **out)


# ---- collate_fn ---------------------------------------------------------------------------------------------------
def reference_method(path, cls_name, fn_name, env):
    tree = ast.parse(open(os.path.join(REF, path)).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls_name:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == fn_name:
                    mod = ast.Module(body=[item], type_ignores=[])
                    exec(compile(mod, os.path.join(REF, path), "exec"), env)
                    return env[fn_name]
    raise KeyError(fn_name)


def collate_cases():
    """name -> list of (img [3,h,w], boxes [n,4], classes [n]); also used (regenerated from the seed) by the tests."""
    cases = {}
    for name, seed, sizes, counts in (("ragged3", 1, [(37, 53), (64, 41), (5, 64)], [3, 0, 7]),
                                      ("same2", 2, [(32, 48), (32, 48)], [4, 4]),
                                      ("one", 3, [(1, 1)], [2]),
                                      ("coco4", 4, [(96, 128), (80, 132), (100, 100), (64, 160)], [11, 1, 0, 25])):
        cases[name] = W.collate_case(seed, sizes, counts)
    return cases


def run_collate(name):
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]            # voc.py:50-51
    env = {"np": np, "torch": torch, "transforms": transforms}
    collate_fn = reference_method("dataset/voc.py", "VOCDataset", "collate_fn", env)
    coco_fn = reference_method("dataset/coco.py", "COCODataset", "collate_fn", dict(env))
    me = types.SimpleNamespace(mean=mean, std=std)
    out = {"mean": np.array(mean), "std": np.array(std)}
    for key, data in collate_cases().items():
        imgs, boxes, classes = collate_fn(me, [(i.clone(), b.clone(), c.clone()) for i, b, c in data])
        imgs2, boxes2, classes2 = coco_fn(me, [(i.clone(), b.clone(), c.clone()) for i, b, c in data])
        assert torch.equal(imgs, imgs2) and torch.equal(boxes, boxes2) and torch.equal(classes, classes2)
        out[key + "_imgs"] = imgs.numpy()
        out[key + "_boxes"] = boxes.numpy()
        out[key + "_classes"] = classes.numpy()
    save(name, **out)


# ---- COCO result rows -----------------------------------------------------------------------------------------------
def reference_export_loop():
    """The statements of evaluate_coco's per-image loop from `scores = scores.detach()...` to the results.append
    (Test_coco.py:149-168), wrapped — untouched — into a function."""
    tree = ast.parse(open(os.path.join(REF, "Test_coco.py")).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "evaluate_coco")
    loop = next(n for n in fn.body if isinstance(n, ast.For))
    start = next(i for i, st in enumerate(loop.body)
                 if isinstance(st, ast.Assign) and getattr(st.targets[0], "id", "") == "scores")
    body = loop.body[start:]
    src = "def export_one(scores, labels, boxes, scale, threshold, generator, index, results, image_ids):\n    pass\n"
    wrapper = ast.parse(src)
    wrapper.body[0].body = body
    ast.fix_missing_locations(wrapper)
    env = {"np": np, "torch": torch}
    exec(compile(wrapper, os.path.join(REF, "Test_coco.py"), "exec"), env)
    return env["export_one"]


def run_coco_export(name):
    export_one = reference_export_loop()
    x = W.head_outputs(3, 80, W.COCO_LEVELS, seed=131)
    head = FCOSHead(0.05, 0.6, 1000, W.STRIDES)
    scales = [1.6659375, 0.8, 2.0775]
    gen = types.SimpleNamespace(ids=[11, 22, 33], id2category={k: 100 + k for k in range(1, 81)})
    results, image_ids = [], []
    det_s, det_c, det_b = np.zeros((3, 1000), np.float32), np.zeros((3, 1000), np.int64), np.zeros((3, 1000, 4), np.float32)
    det_n = np.zeros(3, np.int32)
    for i in range(3):
        xb = [[t[i:i + 1] for t in part] for part in x]
        s, c, b = head(xb)
        b = ClipBoxes()(torch.zeros(1, 3, *W.COCO_HW), b)
        n = s.shape[1]
        det_n[i] = n                                        # the stage input: the reference's own detections
        det_s[i, :n], det_c[i, :n], det_b[i, :n] = s[0].numpy(), c[0].numpy(), b[0].numpy()
        export_one(s, c, b, scales[i], 0.3, gen, i, results, image_ids)
    assert image_ids == gen.ids and len(results) > 0
    save(name, meta=np.array([3, 80, 131, 0.3]), scales=np.array(scales, dtype=np.float64), ids=np.array(gen.ids),
         det_scores=det_s, det_classes=det_c, det_boxes=det_b, det_counts=det_n,
         image_id=np.array([r["image_id"] for r in results]), category_id=np.array([r["category_id"] for r in results]),
         score=np.array([r["score"] for r in results], dtype=np.float64),
         bbox=np.array([r["bbox"] for r in results], dtype=np.float64))


if __name__ == "__main__":
    run_head_saturated("head_voc_saturated", W.VOC_LEVELS, W.VOC_HW, 20, seed=31, strides=W.STRIDES)
    run_collate("collate")
    run_coco_export("coco_export")
