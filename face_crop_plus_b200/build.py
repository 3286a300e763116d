"""Builds libfcpb200.so in-tree with nvcc for sm_100a (the .so is git-ignored but travels to the GPU box)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libfcpb200.so"
SOURCES = ["runtime.cu", "graphs.cu", "conv_ffma.cu", "conv_tc.cu", "misc.cu", "det_post.cu", "align.cu", "parse.cu", "ingest.cu", "comm.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xcompiler", "-Wall", "-Xcompiler", "-ffp-contract=off", "-Xcudafe", "--diag_suppress=177",
         *os.environ.get("FCP_NVCC_FLAGS", "").split()]   # e.g. -DFCP_EXP_TIMELINE for the clock64 timeline build


def _stale(obj: Path, src: Path) -> bool:
    if not obj.exists():
        return True
    deps = [src, *CSRC.glob("*.h"), HERE.parent / "include" / "fcp_b200.h"]
    return any(d.stat().st_mtime > obj.stat().st_mtime for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    jobs = []
    for name in SOURCES:
        src, obj = CSRC / name, objdir / (name + ".o")
        if force or _stale(obj, src):
            jobs.append([NVCC, *FLAGS, "-c", str(src), "-o", str(obj)])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr, file=sys.stderr)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        list(pool.map(run, jobs))
    objs = [str(objdir / (n + ".o")) for n in SOURCES]
    if jobs or not LIB.exists():
        run([NVCC, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
