"""Build the carmel_b200 native code in-tree.

  libcarmel_b200.so   CUDA kernels + C ABI (include/carmel_b200.h), sm_100a only
  carmel-b200         host C++ command line (carmel's -t / --train-cascade / --crp grammar)

nvcc cross-compiles without a GPU.  Outputs land in carmel_b200/_build/ (git-ignored, but they
travel to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libcarmel_b200.so")
CLI = os.path.join(OUT, "carmel-b200")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-pthread", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: carmel_b200 cannot be built (there is no CPU fallback)")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(dirpath: str) -> list[str]:
    out = []
    for base, _, files in os.walk(dirpath):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h", ".c")):
                out.append(os.path.join(base, f))
    return out


FOREST_CLI = os.path.join(OUT, "forest-em-b200")
SYNTH_LIB = os.path.join(OUT, "libcb200_synth.so")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every source to an object (only when it or a header changed), link the shared library and the
    two command lines (carmel-b200, forest-em-b200)."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OUT, exist_ok=True)
    obj_dir = os.path.join(OUT, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    all_src = _sources(CSRC) + [os.path.join(ROOT, "include", "carmel_b200.h"), os.path.abspath(__file__)]
    if not force and _newer(LIB, all_src) and _newer(CLI, all_src) and _newer(FOREST_CLI, all_src):
        return LIB
    nvcc = _nvcc()
    headers = [f for f in all_src if f.endswith((".cuh", ".hpp", ".h", ".py"))]
    host = os.path.join(CSRC, "host")
    cu = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    host_cpp = [os.path.join(host, f) for f in sorted(os.listdir(host)) if f.endswith(".cpp")]
    mains = [f for f in host_cpp if f.endswith("main.cpp")]
    lib_src = cu + [f for f in host_cpp if f not in mains]

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, os.path.basename(src) + ".o")
        if force or not _newer(obj, [src] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", "-I", os.path.join(ROOT, "include"), "-o", obj, src]
            if verbose and src.endswith(".cu"):
                cmd.insert(1, "-Xptxas=-v")
            subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(lib_src))) as ex:
        objs = list(ex.map(compile_one, lib_src))
    subprocess.run([nvcc, *NVCC_FLAGS, "-shared", "-o", LIB, *objs, "-lcudart", "-ldl"], check=True)
    # bench / test tooling (not part of the product library): the synthetic forest generator
    tool = os.path.join(CSRC, "tools", "forest_synth.c")
    if os.path.exists(tool):
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", SYNTH_LIB, tool], check=True)
    for main, exe in ((os.path.join(host, "main.cpp"), CLI), (os.path.join(host, "forest_main.cpp"), FOREST_CLI)):
        if os.path.exists(main):
            subprocess.run(["g++", "-O3", "-std=c++17", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"), "-o", exe, main,
                            "-L", OUT, "-lcarmel_b200", "-Wl,-rpath,$ORIGIN"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
