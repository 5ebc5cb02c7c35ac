"""TEST INFRASTRUCTURE: host emulation of libslicq.

The *same* kernel sources (xumx_slicq_b200/csrc/*.cu) are compiled with g++ -DSLICQ_EMU; every
CTA runs as one emulated thread (see slicq_common.cuh).  This lets the CPU-only test tier
exercise the real kernel index math, the C-ABI and the Python wrappers without a GPU.  The
product never loads this library: tests install it by monkeypatching `xumx_slicq_b200.nsgt._BACKEND`.
"""
from __future__ import annotations

import contextlib
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "xumx_slicq_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libslicq_emu.so")
SOURCES = ["slicq_api.cu", "k_bins.cu", "k_slice.cu", "k_slice_generic.cu"]


def build_emu() -> str:
    from xumx_slicq_b200.build import generate_codelets
    generate_codelets()
    os.makedirs(OUT, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "slicq.h")]
    newest = max(os.path.getmtime(d) for d in deps)
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB

    def cc(src):
        o = os.path.join(OUT, src.replace(".cu", ".o"))
        subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-O1", "-DSLICQ_EMU", "-fPIC", "-w",
                               "-c", os.path.join(CSRC, src), "-o", o])
        return o

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(cc, SOURCES))
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs)
    return LIB


class EmuBackend:
    name = "emu"

    def __init__(self):
        from xumx_slicq_b200 import _cabi
        self._lib = _cabi.load(build_emu())

    def lib(self):
        return self._lib

    def check(self, t):
        if t.is_cuda:
            raise RuntimeError("emulation backend takes CPU tensors")

    def stream(self, device):
        return 0

    def device_guard(self, device):
        return contextlib.nullcontext()
