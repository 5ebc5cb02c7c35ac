"""GPU tier, multi-device / multi-thread behaviour of the C-ABI and the NCCL path of the slice-range sharding.
(per-device kernel attributes, thread re-entrancy of one plan, sharding over real NCCL.)"""
import os
import socket
import sys
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def base():
    import io, contextlib
    from xumx_slicq_b200 import NSGTBase
    with contextlib.redirect_stdout(io.StringIO()):
        return NSGTBase("bark", 262, 32.9, device=torch.device("cuda:0"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_devices_one_process(base):
    """One process drives cuda:0 and cuda:1 (nn.DataParallel style): plans and kernel attributes are per device."""
    nsg = base.nsgt
    T = 200000
    x0 = torch.rand(2, T, device="cuda:0") * 2 - 1
    C0 = nsg.forward_rows(x0)
    y0 = nsg.backward_rows(C0, T)
    x1 = x0.to("cuda:1")
    C1 = nsg.forward_rows(x1)
    y1 = nsg.backward_rows(C1, T)
    torch.cuda.synchronize("cuda:0"); torch.cuda.synchronize("cuda:1")
    assert all(torch.equal(a.cpu(), b.cpu()) for a, b in zip(C0, C1))
    assert torch.equal(y0.cpu(), y1.cpu())


def test_two_threads_one_plan(base):
    """Two host threads call forward + inverse on ONE plan at the same time (own streams and buffers), with calls
    large enough to take the internal two-stream row split: results equal the single-thread ones bit for bit."""
    nsg = base.nsgt
    dev = torch.device("cuda:0")
    T = 1323000
    xs = [torch.rand(4, T, device=dev) * 2 - 1 for _ in range(2)]      # 4 rows x 148 slices + inverse of 16 rows: split path
    refs = []
    for x in xs:
        C = nsg.forward_rows(x)
        Y = [torch.cat([c, 0.5 * c, 0.25 * c, 0.125 * c], 0) for c in C]
        refs.append((C, nsg.backward_rows(Y, T)))
    torch.cuda.synchronize(dev)
    outs = [None, None]
    errs = []

    def work(i):
        try:
            st = torch.cuda.Stream(dev)
            with torch.cuda.stream(st):
                for _ in range(3):
                    C = nsg.forward_rows(xs[i])
                    Y = [torch.cat([c, 0.5 * c, 0.25 * c, 0.125 * c], 0) for c in C]
                    y = nsg.backward_rows(Y, T)
                st.synchronize()
            outs[i] = (C, y)
        except Exception as e:        # surfaced by the assert below
            errs.append(e)
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for (C, y), (Cr, yr) in zip(outs, refs):
        assert all(torch.equal(a, b) for a, b in zip(C, Cr))
        assert torch.equal(y, yr)


def _nccl_worker(rank, world, port, T, out_dir):
    sys.path.insert(0, ROOT)
    import io, contextlib
    import torch.distributed as dist
    from xumx_slicq_b200 import NSGTBase
    from xumx_slicq_b200.sharding import SliceShardedSliCQT
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            base = NSGTBase("bark", 262, 32.9, device=dev)
        nsg = base.nsgt
        g = torch.Generator(device="cpu").manual_seed(3)
        x = (torch.rand(2, T, generator=g) * 2 - 1).to(dev)
        sh = SliceShardedSliCQT(nsg, T, persistent=True)
        for _ in range(2):                       # second pass runs on the cached working set
            C = sh.forward(sh.local_input(x))
            y = sh.inverse(C)
        torch.cuda.synchronize(dev)
        full = nsg.forward_rows(x)
        y_full = nsg.backward_rows(full, T)
        ok_c = all(torch.equal(pc, fc[:, :, sh.k0:sh.k1]) for pc, fc in zip(C, full))
        ok_y = torch.equal(y, y_full[:, sh.lo:sh.hi])
        flag = torch.tensor([int(ok_c and ok_y)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0 and int(flag.item()) == 1:
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_slice_sharding_over_nccl(tmp_path):
    """BASELINE.json configs[3] over real NCCL: 2 ranks, one half-slice halo per boundary and direction; every rank's
    coefficients and owned samples are bitwise equal to the unsharded transform."""
    import torch.multiprocessing as mp
    T = 20 * 9030 + 1234
    mp.spawn(_nccl_worker, args=(2, _free_port(), T, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")
