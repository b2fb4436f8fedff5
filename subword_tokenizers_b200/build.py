"""Builds libswt.so (hand-written sm_100a CUDA + the C ABI of include/swt.h) in-tree with nvcc.

    python -m subword_tokenizers_b200.build [--force] [-v]

Every .cu file is compiled to an object in parallel (objects under csrc/_obj/, rebuilt only when the source or a header
changed), then linked.  The shared library is written next to this file so that it travels with the repo snapshot; it
is git-ignored (built artefact).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libswt.so")
SOURCES = ["api.cu", "bpe_encode.cu", "wp_encode.cu", "bpe_train.cu", "pipeline.cu", "pretok.cu", "types.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(HERE, "..", "include", "swt.h")]


def _newer(target: str, deps) -> bool:
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or _newer(obj, [src] + hdrs):
            extra = os.environ.get("SWT_NVCC_EXTRA", "").split()          # experiments: e.g. -DSWT_COUNT_CTAS=3
            jobs.append((s, [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]))

    def run(job):
        name, cmd = job
        res = subprocess.run(cmd, capture_output=True, text=True)
        return name, res

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for name, res in ex.map(run, jobs):
                if res.returncode != 0:
                    raise RuntimeError("nvcc failed on %s:\n%s\n%s" % (name, res.stdout, res.stderr))
                if verbose:
                    print("==== %s\n%s" % (name, res.stderr))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES]
    if jobs or force or _newer(LIB, objs):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "shared"] + objs + ["-o", LIB]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
