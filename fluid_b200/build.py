"""Builds libpbf_b200.so (the product: CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m fluid_b200.build [--force] [--verbose]

The library has no dependency on torch or on anything under oracle/."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpbf_b200.so")
SOURCES = ["pbf_api.cu", "pbf_kernels.cu", "pbf_multi.cpp"]   # pbf_kernels.cu includes pbf_slab.inl and pbf_surface.inl; pbf_multi.cpp is host-only
HEADERS = ["pbf_internal.h", "pbf_device.cuh", "pbf_slab.inl", "pbf_surface.inl", "pbf_probe.inl", "pbf_mc_table.h", os.path.join("..", "..", "include", "pbf_b200.h"),
           os.path.join("..", "..", "include", "pbf_b200_slab.h"), os.path.join("..", "..", "include", "pbf_b200_multi.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-O3,-fno-fast-math", "-shared", "-Xptxas=-v"]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [NVCC] + FLAGS + ["-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-o", LIB] + sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libpbf_b200.so")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(LIB)
