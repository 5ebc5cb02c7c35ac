"""CPU tier: the real kernel sources (compiled for the host, one emulated thread per CTA) called
through the C-ABI and the Python wrappers, checked against the numpy oracle and the
reference-generated golden vectors.  The GPU tier (test_gpu_parity.py) repeats this on the B200."""
import os

import numpy as np
import pytest
import torch

import common
from oracle.slicq_oracle import SlicqOracle, snr_db


@pytest.fixture(scope="module")
def emu(request):
    from tests.emu.emu_backend import EmuBackend
    import xumx_slicq_b200.nsgt as nsgt_mod
    old = nsgt_mod._BACKEND
    nsgt_mod._BACKEND = EmuBackend()
    yield nsgt_mod._BACKEND
    nsgt_mod._BACKEND = old


@pytest.fixture(scope="module")
def base(emu):
    from xumx_slicq_b200 import NSGTBase
    return NSGTBase("bark", 262, 32.9, device="cpu")


@pytest.fixture(scope="module")
def orc():
    return SlicqOracle(**common.BARK)


def rel_err(c, r):
    return float(np.abs(c - r).max() / np.abs(r).max())


def test_forward_matches_reference_golden(base, golden_dir):
    gold = np.load(os.path.join(golden_dir, "small_fwdinv.npz"))
    buckets = [tuple(int(v) for v in b) for b in gold["buckets"]]
    x = torch.from_numpy(common.small_input())
    C = base.nsgt.forward((x,))                      # list of [S,N,F,M]
    ref = common.unpack(gold["coefs"], buckets)
    assert len(C) == 70
    worst = 0.0
    for c, r in zip(C, ref):
        assert tuple(c.shape) == r.shape
        worst = max(worst, rel_err(c.numpy(), r))
    assert worst < 1e-5, worst                       # north-star tolerance: max relative coefficient error


def test_inverse_matches_reference_golden(base, golden_dir):
    gold = np.load(os.path.join(golden_dir, "small_fwdinv.npz"))
    buckets = [tuple(int(v) for v in b) for b in gold["buckets"]]
    ref = common.unpack(gold["coefs"], buckets)
    T = common.SMALL_T
    y = base.nsgt.backward([torch.from_numpy(np.ascontiguousarray(r)) for r in ref], T).numpy()
    np.testing.assert_allclose(y, gold["y_roundtrip"], atol=5e-6)
    P = common.perturb(ref)
    yp = base.nsgt.backward([torch.from_numpy(p) for p in P], T).numpy()
    np.testing.assert_allclose(yp, gold["y_perturbed"], atol=5e-6)
    assert snr_db(gold["y_perturbed"], yp) > 120.0


def test_roundtrip_snr_vs_reference(base, orc):
    x = common.small_input()
    xt = torch.from_numpy(x)
    C = base.nsgt.forward((xt,))
    y = base.nsgt.backward(C, x.shape[-1]).numpy()
    ours = snr_db(x, y)
    o32 = SlicqOracle(**common.BARK, dtype=np.float32)
    theirs = snr_db(x, o32.backward(o32.forward(x), x.shape[-1]))
    assert ours > theirs - 0.1, (ours, theirs)       # within 0.1 dB of an fp32 reference-style path
    assert ours > 125.0


def test_wrappers_layout_and_shapes(base, orc, golden_dir):
    from xumx_slicq_b200 import make_filterbanks, ComplexNorm
    gold = np.load(os.path.join(golden_dir, "small_fwdinv.npz"))
    nsgt, insgt = make_filterbanks(base)
    x = torch.from_numpy(common.small_input()).view(1, 2, -1)
    X = nsgt(x)
    assert [list(t.shape) for t in X] == gold["wrapper_shapes"].tolist()
    assert all(t.dtype == torch.float32 and t.is_contiguous() for t in X)
    W = orc.nsgt_sl(x.numpy().astype(np.float64))
    for a, b in zip(X, W):
        assert np.abs(a.numpy() - b).max() / np.abs(b).max() < 1e-5
    mags = ComplexNorm()(X)
    assert mags[1].shape == X[1].shape[:-1]
    y = insgt(X, x.shape[-1])
    assert y.shape == x.shape
    assert snr_db(x.numpy(), y.numpy()) > 125.0
    # 7-D input (targets, batch, channels, ...) as the separator passes it (separator.py:174)
    Y4 = [torch.stack([t * s for s in (1.0, 0.5, 0.25, 0.125)]) for t in X]
    y4 = insgt(Y4, x.shape[-1])
    assert y4.shape == (4, 1, 2, x.shape[-1])
    np.testing.assert_allclose(y4[2].numpy(), 0.25 * y.numpy(), atol=2e-6)
    # inverse must not modify its input (the reference does)
    before = [t.clone() for t in X]
    insgt(X, x.shape[-1])
    assert all(torch.equal(a, b) for a, b in zip(X, before))
    # reference-style permuted (non-contiguous) inputs are accepted
    Xp = [t.permute(0, 1, 3, 2, 4, 5).contiguous().permute(0, 1, 3, 2, 4, 5) for t in X]
    assert not Xp[1].is_contiguous()
    np.testing.assert_allclose(insgt(Xp, x.shape[-1]).numpy(), y.numpy(), atol=1e-7)
    # buckets whose storage starts on an odd complex element (8- but not 16-byte aligned): the kernels
    # fall back from their 16-byte loads; same values, bit for bit
    Xo = []
    for t in X:
        flat = torch.empty(t.numel() + 2, dtype=torch.float32)
        flat[2:].copy_(t.reshape(-1))
        Xo.append(flat[2:].view(t.shape))
    assert Xo[1].data_ptr() % 16 == 8 or X[1].data_ptr() % 16 == 8
    assert torch.equal(insgt(Xo, x.shape[-1]), y)


def test_edge_lengths(base, golden_dir):
    edge = np.load(os.path.join(golden_dir, "edge_lengths.npz"))
    for T in (1, 4515, 9030, 9031, 18061):
        xe = (np.random.RandomState(T).rand(1, T).astype(np.float32) * 2 - 1)
        C = base.nsgt.forward((torch.from_numpy(xe),))
        assert C[0].shape[0] == int(edge[f"S_{T}"])
        E = np.asarray([float((c.abs() ** 2).sum()) for c in C])
        np.testing.assert_allclose(E, edge[f"E_{T}"], rtol=1e-4, atol=1e-9)
        y = base.nsgt.backward(C, T).numpy()
        np.testing.assert_allclose(y, edge[f"y_{T}"], atol=5e-6)


def test_slice_range_sharding_is_bitwise(base):
    """slices [k0,k1) computed from the local sample range, one half-slice halo per
    boundary; every output sample is a two-term sum, so the result equals the unsharded one bit for bit."""
    nsg = base.nsgt
    hop = nsg.sl_len // 2
    T = 5 * hop + 1234
    x = torch.from_numpy((np.random.RandomState(7).rand(2, T).astype(np.float32) * 2 - 1))
    S = nsg.n_slices(T)
    full = nsg.forward_rows(x)
    y_full = nsg.backward_rows(full, T)
    cuts = [0, 2, 5, S]
    ys, halos = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        lo, hi = max(0, (a - 1) * hop), min(T, b * hop)
        part = nsg.forward_rows(x[:, lo:hi].contiguous(), k0=a, n_slices=b - a, t0=lo)
        for pc, fc in zip(part, full):
            assert torch.equal(pc, fc[:, :, a:b])
        halo = torch.zeros(2, hop)
        n_out = min(T, b * hop) - a * hop
        ys.append(nsg.backward_rows(part, n_out, k0=a, t0=a * hop, halo_out=halo))
        halos.append(halo)
    for i in range(len(ys) - 1):          # the ONE message per boundary: right shard -> left shard
        ys[i][:, -hop:] += halos[i + 1]
    y = torch.cat(ys, dim=1)
    assert y.shape == y_full.shape
    assert torch.equal(y, y_full)


def test_output_alignment_paths_are_bitwise(base):
    """The same samples through the bulk-copy path (8-byte aligned rows, both 16-byte phases) and the sample-by-sample
    path (odd row stride) of the synthesis slice kernel: lengths T, T-1, T-2 change the row stride of y only."""
    nsg = base.nsgt
    T = 3 * 9030
    x = torch.from_numpy(np.random.RandomState(11).uniform(-1, 1, (3, T)).astype(np.float32))
    C = nsg.forward_rows(x)
    y = nsg.backward_rows(C, T)
    for cut in (1, 2):
        assert torch.equal(nsg.backward_rows(C, T - cut), y[:, :T - cut]), cut


def test_errors(base, emu):
    from xumx_slicq_b200 import make_filterbanks
    with pytest.raises(ValueError):
        make_filterbanks(base, sample_rate=48000.0)
    with pytest.raises(ValueError):
        base.nsgt.backward_rows([torch.zeros(1, 1, 2, 16, dtype=torch.complex64)], 100)


def test_masked_inverse_equals_plain_inverse_of_masked_coefficients(base):
    """synthesis fused with mask * mixture (realtime model)."""
    from xumx_slicq_b200 import make_filterbanks
    nsgt, insgt = make_filterbanks(base)
    x = torch.from_numpy(common.small_input()[:, :14000]).view(1, 2, -1).contiguous()
    X = nsgt(x)
    g = torch.Generator().manual_seed(5)
    masks = [torch.rand((3,) + tuple(Xb.shape[:-1]), generator=g) for Xb in X]          # [targets,B,C,F,S,M]
    y_ref = insgt([m.unsqueeze(-1) * Xb.unsqueeze(0) for m, Xb in zip(masks, X)], x.shape[-1])
    y = insgt.forward_masked(X, masks, x.shape[-1])
    assert y.shape == y_ref.shape == (3, 1, 2, x.shape[-1])
    assert torch.equal(y, y_ref)


def test_forward_with_norm_equals_complexnorm(base):
    """|X| written by the analysis epilogue == ComplexNorm()(X) of the same call."""
    from xumx_slicq_b200 import make_filterbanks, ComplexNorm
    nsgt, _ = make_filterbanks(base)
    x = torch.from_numpy(common.small_input()[:, :14000]).view(1, 2, -1).contiguous()
    X_ref = nsgt(x)
    X, Xmag = nsgt.forward_with_norm(x)
    ref = ComplexNorm()(X_ref)
    assert len(X) == len(Xmag) == len(ref)
    for a, a_ref, m, m_ref in zip(X, X_ref, Xmag, ref):
        assert torch.equal(a, a_ref)                         # the coefficients are untouched by the fusion
        assert m.shape == m_ref.shape and m.dtype == torch.float32
        scale = float(m_ref.max())
        assert float((m - m_ref).abs().max()) <= 1e-6 * scale   # hypot vs sqrt(fma): a few ulp


def test_inverse_autograd_is_the_exact_adjoint(base):
    """gradients through INSGT_SL.  The transform is linear, so the gradient of
    L = <S c, g> w.r.t. c must be S^T g: check <S c, g> == <c, S^T g> and a directional derivative."""
    from xumx_slicq_b200 import make_filterbanks
    nsgt, insgt = make_filterbanks(base)
    T = 14000
    x = torch.from_numpy(common.small_input()[:, :T]).view(1, 2, -1).contiguous()
    X = [Xb.clone().requires_grad_(True) for Xb in nsgt(x)]
    g = torch.randn(1, 2, T, generator=torch.Generator().manual_seed(3))
    y = insgt(X, T)
    assert y.requires_grad
    (y * g).sum().backward()
    lhs = float((y.detach().double() * g.double()).sum())
    rhs = float(sum((Xb.detach().double() * Xb.grad.double()).sum() for Xb in X))
    assert abs(lhs - rhs) <= 2e-5 * max(abs(lhs), 1.0), (lhs, rhs)      # <S c, g> == <c, S^T g>
    # directional derivative along a random direction d: d/de <S(c + e d), g> = <d, S^T g>
    d = [torch.randn(Xb.shape, generator=torch.Generator().manual_seed(7 + i)) for i, Xb in enumerate(X)]
    with torch.no_grad():
        y2 = insgt([Xb.detach() + 0.5 * db for Xb, db in zip(X, d)], T)
    fd = float(((y2 - y.detach()).double() * g.double()).sum()) / 0.5
    an = float(sum((db.double() * Xb.grad.double()).sum() for db, Xb in zip(d, X)))
    assert abs(fd - an) <= 2e-4 * max(abs(an), 1.0), (fd, an)


def test_forward_autograd_is_the_exact_adjoint(base):
    """SURVEY section 8(f) N3, other direction: gradients through NSGT_SL (adjoint of the analysis on the synthesis
    kernels).  <A x, d> == <x, A^T d>, a directional derivative, and the chain forward -> inverse."""
    from xumx_slicq_b200 import make_filterbanks
    nsgt, insgt = make_filterbanks(base)
    T = 14000
    x = torch.from_numpy(common.small_input()[:, :T]).view(1, 2, -1).contiguous().requires_grad_(True)
    X = nsgt(x)
    assert all(Xb.requires_grad for Xb in X)
    d = [torch.randn(Xb.shape, generator=torch.Generator().manual_seed(11 + i)) for i, Xb in enumerate(X)]
    sum((Xb * db).sum() for Xb, db in zip(X, d)).backward()
    lhs = float(sum((Xb.detach().double() * db.double()).sum() for Xb, db in zip(X, d)))
    rhs = float((x.detach().double() * x.grad.double()).sum())
    assert abs(lhs - rhs) <= 2e-5 * max(abs(lhs), 1.0), (lhs, rhs)          # <A x, d> == <x, A^T d>
    e = torch.randn(x.shape, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        X2 = nsgt(x.detach() + 0.25 * e)
    fd = float(sum(((b2 - b1.detach()).double() * db.double()).sum() for b1, b2, db in zip(X, X2, d))) / 0.25
    an = float((e.double() * x.grad.double()).sum())
    assert abs(fd - an) <= 2e-4 * max(abs(an), 1.0), (fd, an)
    # reconstruction loss through both transforms: d/dx 0.5 |insgt(nsgt(x)) - t|^2 = (perfect reconstruction) x - t
    x2 = x.detach().clone().requires_grad_(True)
    tgt = torch.randn(x.shape, generator=torch.Generator().manual_seed(6))
    (0.5 * (insgt(nsgt(x2), T) - tgt) ** 2).sum().backward()
    assert float((x2.grad - (x2.detach() - tgt)).abs().max()) < 2e-4


def test_streamed_separator_is_overlap_exact(base):
    """chunks = slice ranges of the one long transform, stitched with the half-slice halo:
    bitwise equal to the unchunked insgt(model(nsgt(x))) for a slice-local model."""
    from xumx_slicq_b200 import make_filterbanks
    from xumx_slicq_b200.pipeline import StreamedSeparator
    nsgt, insgt = make_filterbanks(base)
    hop = base.nsgt.sl_len // 2
    T = 5 * hop + 321
    x = torch.from_numpy(np.random.RandomState(9).rand(1, 2, T).astype(np.float32) * 2 - 1)
    model = lambda X: [torch.stack([0.75 * Xb, 0.25 * Xb]) for Xb in X]
    y_ref = insgt(model(nsgt(x)), T)
    for chunk in (2, 100):
        sep = StreamedSeparator(base, model, chunk_slices=chunk)
        blocks = list(sep.stream(x))
        assert blocks[0][0] == 0 and blocks[-1][1] == T and all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
        y = sep(x)
        assert y.shape == y_ref.shape == (2, 1, 2, T)
        assert torch.equal(y, y_ref), chunk


@pytest.mark.parametrize("idx", [0, 1])
def test_other_configurations_vs_reference_golden(emu, golden_dir, idx):
    """Bark(100, 50 Hz) (slice length 6884 = 4 * 1721, generic slice kernels) and tiny-mel
    (mel, 32 bins, 115.5 Hz: slice length 2016, first bin below DC -> the reference's mirrored-bin pass matters) against
    vectors generated by the unmodified reference (tests/golden/make_golden_alt.py)."""
    from xumx_slicq_b200 import NSGTBase
    gold = np.load(os.path.join(golden_dir, "alt_fwdinv.npz"))
    name = str(gold[f"name{idx}"])
    fb, fmin, sllen, T = gold[f"cfg{idx}"]
    b = NSGTBase(name, int(fb), float(fmin), device="cpu")
    assert b.sllen == int(sllen)
    x = (np.random.RandomState(100 + idx).rand(2, int(T)).astype(np.float32) * 2 - 1)
    buckets = [tuple(int(v) for v in bb) for bb in gold[f"buckets{idx}"]]
    ref = common.unpack(gold[f"coefs{idx}"], buckets)
    C = b.nsgt.forward((torch.from_numpy(x),))
    assert len(C) == len(ref)
    worst = max(rel_err(c.numpy(), r) for c, r in zip(C, ref))
    assert worst < 1e-5, worst
    y = b.nsgt.backward([torch.from_numpy(np.ascontiguousarray(r)) for r in ref], int(T)).numpy()
    np.testing.assert_allclose(y, gold[f"y_roundtrip{idx}"], atol=5e-6)
    P = common.perturb(ref)
    yp = b.nsgt.backward([torch.from_numpy(p) for p in P], int(T)).numpy()
    np.testing.assert_allclose(yp, gold[f"y_perturbed{idx}"], atol=5e-6)
