"""CPU tier: the multi-rank host logic (slice-range partition + one half-slice halo message per
boundary) over the gloo backend with world_size 2 and 3.  The per-rank transform runs on the
host-emulated kernels (test infrastructure); on the GPU box the same code runs over NCCL
(test_gpu_parity.py::test_slice_range_sharding_bitwise covers the kernel side with virtual ranks)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, T, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import io, contextlib
    from tests.emu.emu_backend import EmuBackend
    import xumx_slicq_b200.nsgt as nsgt_mod
    from xumx_slicq_b200 import NSGTBase
    from xumx_slicq_b200.sharding import SliceShardedSliCQT, slice_partition
    nsgt_mod._BACKEND = EmuBackend()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            base = NSGTBase("bark", 262, 32.9, device="cpu")
        nsg = base.nsgt
        x = torch.from_numpy(np.random.RandomState(11).rand(2, T).astype(np.float32) * 2 - 1)
        sh = SliceShardedSliCQT(nsg, T)
        assert (sh.k0, sh.k1) == slice_partition(nsg.n_slices(T), world)[rank]
        C = sh.forward(sh.local_input(x))
        y = sh.inverse(C)
        # fused mask * mixture on the shard: two targets with constant masks == scaled plain synthesis of the shard
        masks = [torch.stack([torch.full(tuple(c.shape), g, dtype=torch.float32) for g in (0.75, 0.25)]) for c in C]
        ym = sh.inverse_masked(C, masks)
        y2 = sh.inverse([torch.cat([c * 0.75, c * 0.25], dim=0) for c in C])
        assert ym.shape == y2.shape == (4, sh.hi - sh.lo)
        assert torch.equal(ym, y2), "sharded masked synthesis differs from the plain one"
        # persistent working set (buffers and bucket views built once): repeated steps give the same bits
        shp = SliceShardedSliCQT(nsg, T, persistent=True)
        for _ in range(2):
            Cp = shp.forward(sh.local_input(x))
            assert all(torch.equal(a, b) for a, b in zip(Cp, C))
            assert torch.equal(shp.inverse(Cp), y)
            assert torch.equal(shp.inverse_masked(Cp, masks), ym)
        torch.save({"k0": sh.k0, "k1": sh.k1, "lo": sh.lo, "hi": sh.hi, "C": C, "y": y},
                   os.path.join(out_dir, f"rank{rank}.pt"))
        dist.barrier()
        if rank == 0:
            full = nsg.forward_rows(x)
            y_full = nsg.backward_rows(full, T)
            ys = []
            for r in range(world):
                d = torch.load(os.path.join(out_dir, f"rank{r}.pt"))
                for pc, fc in zip(d["C"], full):
                    assert torch.equal(pc, fc[:, :, d["k0"]:d["k1"]]), f"forward shard {r} differs"
                ys.append(d["y"])
            y_cat = torch.cat(ys, dim=1)
            assert y_cat.shape == y_full.shape
            assert torch.equal(y_cat, y_full), "sharded synthesis is not bitwise equal to the unsharded one"
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,hops,extra", [(2, 4, 777), (3, 4, 777), (3, 2, 0)])
def test_slice_sharding_over_gloo(world, hops, extra, tmp_path):
    from tests.emu.emu_backend import build_emu
    build_emu()
    hop = 9030
    # 6 slices: uneven 3-way split, last hop partial; 2 hops / 3 ranks: S = 3, the last rank owns NO output samples
    # (its slice lies past the signal end) but still has to deliver its halo
    T = hops * hop + extra
    mp.spawn(_worker, args=(world, _free_port(), T, str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(tmp_path / "ok")


def test_partition_helpers():
    from xumx_slicq_b200.sharding import slice_partition, shard_tracks, owned_samples
    assert slice_partition(881, 8)[0] == (0, 110) and slice_partition(881, 8)[-1][1] == 881
    assert sum(b - a for a, b in slice_partition(881, 8)) == 881
    with pytest.raises(ValueError):
        slice_partition(3, 4)
    assert shard_tracks(10, 1, 4) == [1, 5, 9]
    assert owned_samples(2, 4, 9030, 30000) == (18060, 30000)
