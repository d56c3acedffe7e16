"""Build libb200det.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m pytorch_object_detection_b200.build [--force]

No torch headers are involved: the library is plain CUDA runtime + extern "C".
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200det.so")
SOURCES = ["api.cu", "score.cu", "select.cu", "nms.cu", "nms_class.cu", "fused.cu", "assign.cu", "loss.cu", "train_fused.cu", "collate.cu", "eval.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-prec-div=true", "-diag-suppress", "177",
    "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; b200det has no CPU fallback and cannot be built without it")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b200det.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    trace = ["-DB200DET_TRACE"] if os.environ.get("B200DET_TRACE") == "1" else []   # phase timestamps (debug)
    cmd = [_nvcc()] + NVCC_FLAGS + trace + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
