"""Build tuning variants of libslicq.so into build/variants/<name>/libslicq.so (parallel nvcc).
usage: python tools/build_variants.py name:-DFOO=1,-DBAR=2 ...   (name 'base' = no extra defines)"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xumx_slicq_b200.build import CSRC, SOURCES, NVCC_FLAGS, generate_codelets
generate_codelets()
jobs = []
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    d = os.path.join(ROOT, "build", "variants", name)
    os.makedirs(d, exist_ok=True)
    defs = [x for x in defs.split(",") if x]
    jobs.append((name, d, defs))
def cc(args):
    name, d, defs, src = args
    o = os.path.join(d, src.replace(".cu", ".o"))
    r = subprocess.run(["nvcc"] + NVCC_FLAGS + defs + ["-c", os.path.join(CSRC, src), "-o", o], capture_output=True, text=True)
    open(o + ".log", "w").write(r.stdout + r.stderr)
    if r.returncode: raise RuntimeError(r.stderr[-2000:])
    regs = [l for l in (r.stdout + r.stderr).splitlines() if "Used" in l or ("spill" in l and " 0 bytes spill stores" not in l)]
    return name, src, regs
with ThreadPoolExecutor(max_workers=8) as ex:
    for name, src, regs in ex.map(cc, [(n, d, f, s) for (n, d, f) in jobs for s in SOURCES]):
        print(name, src, [r.strip()[-60:] for r in regs])
for name, d, defs in jobs:
    objs = [os.path.join(d, s.replace(".cu", ".o")) for s in SOURCES]
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", os.path.join(d, "libslicq.so")] + objs)
    print("built", os.path.join(d, "libslicq.so"))
