"""Build libb200det.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m pytorch_object_detection_b200.build [--force] [-v]

No torch headers are involved: the library is plain CUDA runtime + extern "C".  Every translation unit is
compiled to an object under ``build/`` (in parallel, only when it or a header changed) and the objects are
linked into the shared library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200det.so")
SOURCES = ["api.cu", "score.cu", "select.cu", "nms.cu", "nms_class.cu", "fused.cu", "assign.cu", "loss.cu",
           "train_fused.cu", "collate.cu", "eval.cu"]
COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-prec-div=true", "-diag-suppress", "177",
    "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-Wall",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; b200det has no CPU fallback and cannot be built without it")


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(HERE, "..", "include", "b200det.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b200det.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, lib: str | None = None, defines=()) -> str:
    """Compile what changed and link.  ``lib`` / ``defines`` build a variant library elsewhere (A/B timing)."""
    out = lib or LIB
    if not force and lib is None and not needs_build():
        return out
    trace = ["-DB200DET_TRACE"] if os.environ.get("B200DET_TRACE") == "1" else []   # phase timestamps (debug)
    extra = trace + [f"-D{d}" for d in defines]
    tag = ("_" + "_".join(sorted(d.replace("=", "-") for d in extra))) if extra else ""
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    newest_header = max(os.path.getmtime(h) for h in _headers())
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + tag + ".o")
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header)
        jobs.append((src, obj, stale))

    def compile_one(job):
        src, obj, stale = job
        if not stale:
            return None
        cmd = [nvcc] + COMPILE_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return res.stderr

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
        logs = list(pool.map(compile_one, jobs))
    if verbose:
        print("\n".join(l for l in logs if l))
    cmd = [nvcc] + LINK_FLAGS + [j[1] for j in jobs] + ["-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
