"""Build libcape_msda.so in-tree with nvcc for sm_100a.

    python category-agnostic-pose-estimation_b200/build.py [--force] [--verbose]

The shared object is written next to this file (git-ignored; it travels to the GPU box with the gpurun snapshot).
Only sm_100a SASS is embedded: no PTX fallback, no other architectures.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcape_msda.so")
SOURCES = ["cape_abi.cu", "msda_forward.cu", "msda_forward_staged.cu", "msda_backward.cu", "msda_backward_staged.cu", "msda_backward_tc.cu", "msda_variants.cu", "seq_tokens.cu", "decode_step.cu", "linear_tf32x3.cu"]
HEADERS = ["msda_common.cuh", "msda_launch.h", "async_copy.cuh", "umma_tf32.cuh", os.path.join("..", "..", "include", "cape_msda.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "") -> str:
    """variant: "" = the product library.  "nored" / "noload" = profiling-only variants of the backward kernel
    (results invalid; written to tools/, loaded only by tools/tune.py through CAPE_MSDA_LIB)."""
    global LIB
    extra = []
    if variant:
        extra = {"nored": ["-DCAPE_EXP_NO_RED"], "noload": ["-DCAPE_EXP_NO_LOAD"], "t896": ["-DCAPE_BS_THREADS=896"],
                 "t768": ["-DCAPE_BS_THREADS=768"], "nocoarsered": ["-DCAPE_EXP_NO_COARSE_RED=12"], "fwdtable": ["-DCAPE_FWD_TABLE=1"], "redcta": ["-DCAPE_EXP_RED_CTA_SCOPE"], "nored0": ["-DCAPE_EXP_NO_COARSE_RED=1"], "nored1": ["-DCAPE_EXP_NO_COARSE_RED=2"],
                 "nored2": ["-DCAPE_EXP_NO_COARSE_RED=4"], "nored3": ["-DCAPE_EXP_NO_COARSE_RED=8"], "nored01": ["-DCAPE_EXP_NO_COARSE_RED=3"], "evictlast": ["-DCAPE_EXP_EVICT_LAST"],
                 "streamstore": ["-DCAPE_EXP_STREAM_STORE"]}[variant]
        lib = os.path.join(os.path.dirname(HERE), "tools", f"libcape_msda_{variant}.so")
        import tempfile   # objects of profiling variants stay out of the tree (the tree travels to the GPU box)
        return _compile(lib, extra, verbose, os.path.join(tempfile.gettempdir(), "cape_msda_build", variant))
    if not force and not _stale():
        return LIB
    return _compile(LIB, extra, verbose, os.path.join(HERE, "build"))


def _compile(LIB: str, extra, verbose: bool, build_dir: str) -> str:
    nvcc = _nvcc()
    objs = []
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- {src}\n{out}", flush=True)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    subprocess.run(link, check=True)
    return LIB


if __name__ == "__main__":
    variant = ""
    for a in sys.argv[1:]:
        if a.startswith("--variant="):
            variant = a.split("=", 1)[1]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=variant))
