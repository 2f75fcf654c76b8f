"""Build the C restatement of the oracle.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.build_oracle        ->  oracle/_build/libmsda_oracle.so

Called by ``__graft_entry__.build()`` so the prebuilt checker travels to the GPU box with the
repo snapshot.  (The reference itself is pure Python — SURVEY.md §2.1 lists no native sources —
so there is no ``oracle/_ref`` to compile; the reference is instead imported in the build
container by ``oracle/make_golden.py`` to generate the committed fixtures.)
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "msda_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libmsda_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-std=gnu11", "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
