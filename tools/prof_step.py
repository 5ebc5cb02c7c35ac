"""Run a few sliCQT steps for profiling under ncu (one forward + one 4-target inverse per step)."""
import argparse, os, sys, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xumx_slicq_b200 import NSGTBase

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--seconds", type=float, default=30.0)
a = ap.parse_args()
dev = torch.device("cuda:0")
with contextlib.redirect_stdout(io.StringIO()):
    base = NSGTBase("bark", 262, 32.9, device=dev)
nsg = base.nsgt
T = int(a.seconds * 44100)
x = torch.rand(2 * a.batch, T, device=dev) * 2 - 1
C = nsg.forward_rows(x)
Y = [torch.cat([c * g for g in (0.9, 0.6, 0.4, 0.2)], dim=0).contiguous() for c in C]
torch.cuda.synchronize()
torch.cuda.profiler.start()   # use with: ncu --profile-from-start off
for _ in range(a.steps):
    C = nsg.forward_rows(x)
    y = nsg.backward_rows(Y, T)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float((y[: 2 * a.batch] - 0.9 * x).abs().max()))
