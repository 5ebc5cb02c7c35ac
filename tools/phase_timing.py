"""Per-phase CTA timing of the slice kernels (needs a -DSLICQ_PHASE_TIMING build, see SLICQ_B200_LIB)."""
import os, sys, io, contextlib, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xumx_slicq_b200 import NSGTBase, _cabi
dev = torch.device("cuda:0")
with contextlib.redirect_stdout(io.StringIO()):
    base = NSGTBase("bark", 262, 32.9, device=dev)
nsg = base.nsgt
B = int(os.environ.get("PT_BATCH", "8")); T = 1323000
x = torch.rand(2 * B, T, device=dev) * 2 - 1
Cc = nsg.forward_rows(x)
Y = [torch.cat([c * g for g in (0.9, 0.6, 0.4, 0.2)], dim=0).contiguous() for c in Cc]
lib = _cabi.load()
lib.slicq_debug_set_timing.argtypes = [C.c_void_p]
names = {0: "start", 1: "load/gather done", 2: "pass A done", 3: "pass B done", 5: "pass C done", 6: "end"}
def run(label, fn, n_cta):
    buf = torch.zeros(n_cta * 16, dtype=torch.int64, device=dev)
    assert lib.slicq_debug_set_timing(C.c_void_p(buf.data_ptr())) == 0
    fn(); torch.cuda.synchronize()
    lib.slicq_debug_set_timing(C.c_void_p(0))
    t = buf.view(n_cta, 16).cpu()
    ok = t[:, 6] > 0
    t = t[ok].double()
    print(f"{label}: {int(ok.sum())} CTAs timed; mean cycles per phase:")
    prev = 0
    for i in (1, 2, 3, 5, 6):
        d = (t[:, i] - t[:, prev]).mean().item()
        print(f"   {names[prev]:>18s} -> {names[i]:<18s} {d:10.0f} cycles")
        prev = i
    if (t[:, 7] > 0).all():
        print(f"   [thread 0] start -> zero-fill done {(t[:,7]-t[:,0]).mean().item():8.0f}; -> own loads done {(t[:,4]-t[:,7]).mean().item():8.0f}; -> barrier passed {(t[:,1]-t[:,4]).mean().item():8.0f}")
    print(f"   raw slots: [4]={t[:,4].mean().item():.0f}  [7]-[0]={(t[:,7]-t[:,0]).mean().item():.0f}  [1]-[7]={(t[:,1]-t[:,7]).mean().item():.0f}")
    print("   slots 8..14 (t0 done-wait, t0 issue, t32 full-wait, t32 process, t32 arrive, t352 full-wait, t352 process):", [round(t[:,i].mean().item()) for i in range(8,15)])
    print(f"   total {((t[:,6]-t[:,0]).mean().item()):10.0f} cycles; kernel span {(t[:,6].max()-t[:,0].min()).item():.0f} cycles")
S = Cc[0].shape[2]
run("slice_fft_fwd", lambda: nsg.forward_rows(x), 2 * B * S)
run("slice_fft_inv (both parities; last launch wins per CTA)", lambda: nsg.backward_rows(Y, T), 8 * B * S)
