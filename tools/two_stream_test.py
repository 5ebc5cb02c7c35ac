"""Experiment: does overlapping two half-batches on two streams (optionally with L2-sized chunks) help?"""
import os, sys, io, contextlib, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xumx_slicq_b200 import NSGTBase
dev = torch.device("cuda:0")
def make(mb):
    os.environ["SLICQ_CHUNK_MB"] = str(mb)
    with contextlib.redirect_stdout(io.StringIO()):
        return NSGTBase("bark", 262, 32.9, device=dev).nsgt
B, T = 8, 1323000
x = torch.rand(2 * B, T, device=dev) * 2 - 1
nsg0 = make(2048)
C = nsg0.forward_rows(x)
Y = [torch.cat([c * g for g in (0.9, 0.6, 0.4, 0.2)], dim=0).contiguous() for c in C]   # [4*2B rows]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(nsg, nstreams, label):
    plan = nsg.plan(dev)
    S = nsg.n_slices(T)
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    parts = []
    rows = 2 * B // nstreams
    for i in range(nstreams):
        xi = x[i * rows:(i + 1) * rows]
        # targets-major rows of this part: for each target t rows [t*2B + i*rows, +rows)
        Yi = [torch.cat([y[t * 2 * B + i * rows: t * 2 * B + (i + 1) * rows] for t in range(4)], 0).contiguous() for y in Y]
        slab, Co = nsg.alloc_coefficients(rows, S, dev)
        vf = [nsg._view_of(c) for c in Co]; vi = [nsg._view_of(c) for c in Yi]
        sf, si = plan.scratch_bytes(rows, S, False), plan.scratch_bytes(4 * rows, S, True)
        scr = torch.empty(max(sf, si), dtype=torch.uint8, device=dev)
        yo = torch.empty(4 * rows, T, device=dev)
        parts.append((xi, Yi, slab, vf, vi, sf, si, scr, yo))
    def step():
        for st, (xi, Yi, slab, vf, vi, sf, si, scr, yo) in zip(streams, parts):
            st.wait_stream(torch.cuda.current_stream(dev))
            plan.forward(xi.data_ptr(), xi.shape[0], xi.stride(0), T, 0, 0, S, vf, scr.data_ptr(), sf, st.cuda_stream)
            plan.inverse(vi, 4 * xi.shape[0], S, 0, yo.data_ptr(), yo.stride(0), T, 0, 0, scr.data_ptr(), si, st.cuda_stream)
        for st in streams:
            torch.cuda.current_stream(dev).wait_stream(st)
    for _ in range(3): step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{label:40s} {sum(ts)/len(ts):7.3f} ms/step")
run(nsg0, 1, "1 stream, chunk 2048 MB")
run(nsg0, 2, "2 streams, chunk 2048 MB")
run(nsg0, 4, "4 streams, chunk 2048 MB")
for mb in (96, 48):
    n = make(mb)
    run(n, 1, f"1 stream, chunk {mb} MB")
    run(n, 2, f"2 streams, chunk {mb} MB")
    run(n, 4, f"4 streams, chunk {mb} MB")
