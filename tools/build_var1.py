"""Quick tuning variants that differ in ONE source file: compile it with extra defines and link against the base objects.
usage: python tools/build_var1.py k_slice.cu name:-DFOO=1,-DBAR=2 ...  -> build/variants/<name>/libslicq.so"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xumx_slicq_b200.build import CSRC, SOURCES, NVCC_FLAGS, OBJ
src = sys.argv[1]
def one(spec):
    name, _, defs = spec.partition(":")
    d = os.path.join(ROOT, "build", "variants", name); os.makedirs(d, exist_ok=True)
    o = os.path.join(d, src.replace(".cu", ".o"))
    r = subprocess.run(["nvcc"] + NVCC_FLAGS + [x for x in defs.split(",") if x] + ["-c", os.path.join(CSRC, src), "-o", o], capture_output=True, text=True)
    open(o + ".log", "w").write(r.stdout + r.stderr)
    if r.returncode: raise RuntimeError(r.stderr[-3000:])
    objs = [o] + [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES if s != src]
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", os.path.join(d, "libslicq.so")] + objs)
    regs = [l.strip()[-70:] for l in (r.stdout + r.stderr).splitlines() if "Used" in l or ("spill" in l and " 0 bytes spill stores" not in l)]
    return name, regs
with ThreadPoolExecutor(max_workers=8) as ex:
    for name, regs in ex.map(one, sys.argv[2:]): print(name, regs)
