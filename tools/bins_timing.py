"""Phase breakdown of the analysis two-pass bin kernel (needs a -DSLICQ_PHASE_TIMING build)."""
import os, sys, io, contextlib, ctypes as C, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xumx_slicq_b200 import NSGTBase, _cabi
dev = torch.device("cuda:0")
with contextlib.redirect_stdout(io.StringIO()):
    base = NSGTBase("bark", 262, 32.9, device=dev)
nsg = base.nsgt
B = int(os.environ.get("PT_BATCH", "8")); T = 1323000
x = torch.rand(2 * B, T, device=dev) * 2 - 1
nsg.forward_rows(x); torch.cuda.synchronize()
lib = _cabi.load()
lib.slicq_debug_set_bins_timing.argtypes = [C.c_void_p]
n = 8192
buf = torch.zeros(n * 8, dtype=torch.int64, device=dev)
assert lib.slicq_debug_set_bins_timing(C.c_void_p(buf.data_ptr())) == 0
nsg.forward_rows(x); torch.cuda.synchronize()
lib.slicq_debug_set_bins_timing(C.c_void_p(0))
t = buf.view(n, 8).cpu()
t = t[t[:, 7] == 1].double()
print(f"{len(t)} two-pass CTAs timed (bins_fwd)")
tot = t[:, 5].sum()
for i, name in ((1, "pass 1 (loads + DFT-A + twiddle + STS)"), (2, "barrier wait after pass 1"), (3, "pass 2 (LDS + DFT-B + stores)"), (4, "barrier wait after pass 2")):
    print(f"   {name:45s} {100 * t[:, i].sum() / tot:5.1f} %")
print(f"   mean cycles per CTA {t[:,5].mean():.0f}")
by = collections.defaultdict(list)
for r in t:
    by[int(r[6])].append(r)
print("   M : share pass1 / wait1 / pass2 / wait2, mean CTA cycles")
for M in sorted(by)[::6]:
    rr = torch.stack(by[M]); s = rr[:, 5].sum()
    print(f"   {M:3d}: " + " / ".join(f"{100 * rr[:, i].sum() / s:4.1f}" for i in (1, 2, 3, 4)) + f"   {rr[:,5].mean():8.0f}  ({len(rr)} CTAs)")

# ---- synthesis two-pass kernel: slot 0 = wait for the coefficient loads, 1 = DFT-A + twiddle + STS, 2 = barrier,
#      3 = pass 2 (LDS + DFT-B + window + stores), 4 = barrier, 5 = iteration total, 6 = M | iterations << 16
C_ = nsg.forward_rows(x)
Y = [torch.cat([c * g for g in (0.9, 0.6, 0.4, 0.2)], dim=0).contiguous() for c in C_]
nsg.backward_rows(Y, T); torch.cuda.synchronize()
buf.zero_()
assert lib.slicq_debug_set_bins_timing(C.c_void_p(buf.data_ptr())) == 0
nsg.backward_rows(Y, T); torch.cuda.synchronize()
lib.slicq_debug_set_bins_timing(C.c_void_p(0))
t = buf.view(n, 8).cpu()
t = t[t[:, 7] == 1]
Ms = (t[:, 6] & 0xffff); its = (t[:, 6] >> 16).double(); t = t.double()
sel = its > 0
t, Ms, its = t[sel], Ms[sel], its[sel]
print(f"{len(t)} two-pass CTAs timed (bins_inv; last launch wins per CTA index)")
tot = t[:, 5].sum()
for i, name in ((0, "wait for coefficient loads"), (1, "pass 1 (DFT-A + twiddle + STS)"), (2, "barrier after pass 1"), (3, "pass 2 (LDS + DFT-B + window + stores)"), (4, "barrier after pass 2")):
    print(f"   {name:45s} {100 * t[:, i].sum() / tot:5.1f} %")
print(f"   mean cycles per iteration {(t[:,5].sum() / its.sum()).item():.0f}, iterations per CTA {its.mean().item():.1f}")
print("   M : load wait / pass1 / bar / pass2 / bar (% of iteration), cycles per iteration")
for M in sorted(set(Ms.tolist()))[::5]:
    m = Ms == M
    rr = t[m]; s = rr[:, 5].sum()
    print(f"   {M:3d}: " + " / ".join(f"{100 * rr[:, i].sum() / s:4.1f}" for i in (0, 1, 2, 3, 4)) + f"   {(s / its[m].sum()).item():8.0f}  ({int(m.sum())} CTAs)")
