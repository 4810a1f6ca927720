"""Build libnlps_b200.so in-tree with nvcc for sm_100a (no torch involved).

    python -m nlpscratch_b200.build          # incremental
    python -m nlpscratch_b200.build --force

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libnlps_b200.so")
SOURCES = ["util.cu", "gemm_simt.cu", "gemm_tcgen05.cu", "attention.cu", "attention_mma.cu", "attention_ts.cu", "rowwise.cu", "embed.cu", "dataset.cu",
           "optimizer.cu", "array_ops.cu", "tf32_peak.cu", "comm_nccl.cu", "model.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libnlps_b200.so cannot be built")


def _digest(paths) -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "nlps_b200.h"))
    nvcc = _nvcc()
    jobs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        stamp = obj + ".sha"
        want = _digest([sp] + headers)
        have = open(stamp).read() if os.path.exists(stamp) else ""
        if force or not os.path.exists(obj) or have != want:
            jobs.append((sp, obj, stamp, want))

    def compile_one(job):
        sp, obj, stamp, want = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", sp, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (sp, res.stdout, res.stderr))
        if verbose and res.stderr.strip():
            print(res.stderr)
        with open(stamp, "w") as fh:
            fh.write(want)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
