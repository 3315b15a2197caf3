"""Build liblsmb200.so (sm_100a only) in-tree with nvcc.  `python -m lsm_speech_classifier_b200.build`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SO = os.path.join(PKG, os.environ.get("LSM_SO_NAME", "liblsmb200.so"))
SOURCES = ["api.cu", "error_bound.cu", "frontend_gammatone.cu", "pipeline_lanes.cu", "frontend_mel.cu", "reservoir.cu", "reservoir_dense.cu", "standardize.cu", "logreg.cu", "resample.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--fmad=false",            # never contract a*b+c: bit parity with the CPU oracle
         "-Xcompiler", "-fPIC", "-cudart", "static"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "lsm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    objs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc(), *FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        if os.environ.get("LSM_PIPE_J"):      # experiment: channels per filter warp of the warp-specialised kernel (default 1)
            cmd.insert(1, "-DLSM_PIPE_J=" + os.environ["LSM_PIPE_J"])
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
        objs.append(obj)
    subprocess.check_call([nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
                           "-o", SO, *objs])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
