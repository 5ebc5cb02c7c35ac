"""Host-side overhead of the public wrappers vs the raw C-ABI calls (single 30 s stereo mixture)."""
import os, sys, io, contextlib, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xumx_slicq_b200 import NSGTBase, make_filterbanks
dev = torch.device("cuda:0")
with contextlib.redirect_stdout(io.StringIO()):
    base = NSGTBase("bark", 262, 32.9, device=dev)
nsgt, insgt = make_filterbanks(base)
T = 1323000
x = torch.rand(1, 2, T, device=dev) * 2 - 1
X = nsgt(x); Y = [torch.stack([Xb * g for g in (0.9, 0.6, 0.4, 0.2)]) for Xb in X]
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    host = (time.perf_counter() - t) / n          # host time to enqueue
    torch.cuda.synchronize(); tot = (time.perf_counter() - t) / n
    return host * 1e3, tot * 1e3
print("NSGT_SL.forward   host %.3f ms  total %.3f ms" % timeit(lambda: nsgt(x)))
print("INSGT_SL.forward  host %.3f ms  total %.3f ms" % timeit(lambda: insgt(Y, T)))
