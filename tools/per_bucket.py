"""Time bins_fwd / bins_inv for each bucket alone (SLICQ_ONLY_BUCKET) -> ns per coefficient per bucket."""
import os, sys, io, contextlib, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xumx_slicq_b200 import NSGTBase, _cabi
dev = torch.device("cuda:0")
rows = int(os.environ.get("PB_ROWS", "16")); T = 1323000
x = torch.rand(rows, T, device=dev) * 2 - 1
res = []
for b in range(-1, 70):
    if b >= 0: os.environ["SLICQ_ONLY_BUCKET"] = str(b)
    else: os.environ.pop("SLICQ_ONLY_BUCKET", None)
    with contextlib.redirect_stdout(io.StringIO()):
        base = NSGTBase("bark", 262, 32.9, device=dev)
    nsg = base.nsgt
    C = nsg.forward_rows(x); y = nsg.backward_rows(C, T); torch.cuda.synchronize()
    _cabi.profile_enable(True)
    for _ in range(3):
        C = nsg.forward_rows(x); y = nsg.backward_rows(C, T)
    torch.cuda.synchronize()
    pr = _cabi.profile_read(); _cabi.profile_enable(False)
    S = C[0].shape[2]; units = rows * S
    if b < 0:
        print("all buckets: units", units, {k: round(v[0] / 3, 4) for k, v in pr.items()})
        buckets = nsg.tables.buckets
        continue
    fb, nb, M = buckets[b]
    pts = nb * M * units
    f_ns = pr["bins_fwd"][0] / 3 * 1e6 / pts; i_ns = pr["bins_inv"][0] / 3 * 1e6 / pts
    res.append((b, nb, M, f_ns, i_ns))
    print(f"bucket {b:2d} F={nb:2d} M={M:3d}  fwd {pr['bins_fwd'][0]/3*1e3:8.1f} us ({f_ns*1e3:6.2f} ps/coef)  inv {pr['bins_inv'][0]/3*1e3:8.1f} us ({i_ns*1e3:6.2f} ps/coef)")
json.dump(res, open("gpurun_out/per_bucket.json", "w"))
tf = sum(r[1] * r[2] * r[3] for r in res); ti = sum(r[1] * r[2] * r[4] for r in res)
print("sum of alone-times per unit: fwd %.1f ns  inv %.1f ns" % (tf, ti))
