"""Builds libfreesplat_b200.so in-tree with nvcc for sm_100a (no torch / pybind in the library:
the boundary is the C ABI of include/freesplat_b200.h).

    python -m freesplat_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfreesplat_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math",
          "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]
# (source, extra flags).  -fmad=false: bit-exact integer decisions in the per-Gaussian stages.
SOURCES = [
    ("raster_pre.cu", ["-fmad=false"]),
    ("raster_bin.cu", ["-fmad=false"]),
    ("raster_render.cu", []),
    ("cost_volume.cu", []),
    ("ptf.cu", ["-fmad=false"]),
    ("ptf_gru_bwd.cu", []),
    ("adapter.cu", []),
    ("depth_head.cu", []),
    ("c_api.cu", []),
]


def _deps():
    out = [os.path.join(HERE, "..", "include", "freesplat_b200.h")]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps() if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    for src, extra in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [NVCC, *ARCH, *COMMON, *extra, "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed (see log above)")
    tmp = LIB + ".tmp"
    cmd = [NVCC, *ARCH, "-shared", "-o", tmp, *objs, "-cudart", "static",
           "-ccbin", COMMON[-1]]
    subprocess.check_call(cmd)
    os.replace(tmp, LIB)          # atomic: a snapshot of the tree (gpurun) never sees a half-written library
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
