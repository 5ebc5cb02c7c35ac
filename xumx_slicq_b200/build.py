"""Build libslicq.so in-tree with nvcc for sm_100a.

    python -m xumx_slicq_b200.build [--force]

Steps: (1) regenerate the DFT codelets (csrc/gen_codelets.py -> dft_codelets.cuh, fft_sizes.inc),
(2) compile each .cu with `-gencode arch=compute_100a,code=sm_100a -lineinfo` in parallel,
(3) link xumx_slicq_b200/libslicq.so.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(HERE, "libslicq.so")
SOURCES = ["slicq_api.cu", "k_bins.cu", "k_slice.cu", "k_slice_generic.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def generate_codelets(force: bool = False) -> None:
    gen = os.path.join(CSRC, "gen_codelets.py")
    outs = [os.path.join(CSRC, "dft_codelets.cuh"), os.path.join(CSRC, "fft_sizes.inc")]
    if not force and all(_newer(o, [gen]) for o in outs):
        return
    subprocess.check_call([sys.executable, gen])


def build(force: bool = False, verbose: bool = True) -> str:
    generate_codelets(force)
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc", ".h"))]
    headers.append(os.path.join(ROOT, "include", "slicq.h"))
    nvcc = os.environ.get("NVCC", "nvcc")

    def compile_one(src: str) -> str:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or not _newer(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            with open(o + ".log", "w") as f:
                f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return o

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or not _newer(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        subprocess.check_call(cmd)
    if verbose:
        print(f"built {LIB} ({os.path.getsize(LIB) / 1e6:.1f} MB)")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
