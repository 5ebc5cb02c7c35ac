"""Pin the numpy oracle (oracle/slicq_oracle.py) against vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

import common
from oracle.slicq_oracle import SlicqOracle, snr_db, BarkScale, nsgfwin


@pytest.fixture(scope="module")
def orc():
    return SlicqOracle(**common.BARK)


@pytest.fixture(scope="module")
def tab(golden_dir):
    return np.load(os.path.join(golden_dir, "tables_bark262.npz"))


def test_scalar_tables(orc, tab):
    assert orc.sl_len == int(tab["sllen"]) == 18060
    assert orc.tr_area == int(tab["trlen"]) == 4516
    assert orc.nn == int(tab["nn"])
    assert orc.nbins == int(tab["fbins_actual"]) == 263
    assert orc.ncoefs == int(tab["ncoefs"]) == 292
    np.testing.assert_array_equal(orc.frqs, tab["frqs"])
    np.testing.assert_allclose(orc.q, tab["q"], rtol=1e-6)


def test_integer_tables_exact(orc, tab):
    np.testing.assert_array_equal(orc.M.astype(np.int64), tab["M"])
    np.testing.assert_array_equal(orc.rfbas.astype(np.int64), tab["rfbas"])
    np.testing.assert_array_equal(np.asarray([w[0] for w in orc.wins]), tab["wins0"])
    assert len(orc.buckets) == 70
    assert sum(nb * M for _, nb, M in orc.buckets) == 18640


def test_windows(orc, tab):
    g = np.concatenate(orc.g[: orc.nbins])
    gd = np.concatenate(orc.gd[: orc.nbins])
    np.testing.assert_allclose(g, tab["g"], atol=1e-6)
    np.testing.assert_allclose(gd, tab["gd"], rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(orc.tukey(), tab["tukey"], atol=1e-7)
    np.testing.assert_allclose(orc.coef_factors(), tab["coef_factors"], rtol=1e-12)


def test_alt_bark_integer_tables(golden_dir):
    alt = np.load(os.path.join(golden_dir, "tables_alt.npz"))
    i = 0
    while f"cfg{i}" in alt:
        fb, fmin = alt[f"cfg{i}"]
        scl = BarkScale(float(fmin), 22050.0, int(fb))
        sllen, trlen = scl.suggested_sllen_trlen(44100.0)
        f, q = scl()
        g, rfbas, M = nsgfwin(f, q, 44100.0, sllen, min_win=16)
        assert [sllen, trlen, len(g) // 2 + 1] == list(alt[f"sl{i}"])
        np.testing.assert_array_equal(M.astype(np.int64), alt[f"M{i}"])
        np.testing.assert_array_equal(rfbas.astype(np.int64), alt[f"rfbas{i}"])
        i += 1
    assert i == 4


def test_forward_and_inverse_small(orc, golden_dir):
    gold = np.load(os.path.join(golden_dir, "small_fwdinv.npz"))
    buckets = [tuple(int(v) for v in b) for b in gold["buckets"]]
    assert buckets == [tuple(b) for b in orc.buckets]
    x = common.small_input()
    C = orc.forward(x.astype(np.float64))
    ref = common.unpack(gold["coefs"], buckets)
    worst = 0.0
    for c, r in zip(C, ref):
        assert c.shape == r.shape
        worst = max(worst, np.abs(c - r).max() / np.abs(r).max())
    assert worst < 2e-6, worst  # reference is fp32: its own rounding noise
    # inverse of the reference's coefficients (range) and of perturbed (non-range) ones
    y = orc.backward([r.astype(np.complex128) for r in ref], x.shape[-1])
    np.testing.assert_allclose(y, gold["y_roundtrip"], atol=4e-6)
    P = common.perturb(ref)
    yp = orc.backward([p.astype(np.complex128) for p in P], x.shape[-1])
    np.testing.assert_allclose(yp, gold["y_perturbed"], atol=4e-6)
    assert snr_db(gold["y_perturbed"], yp) > 120.0
    # wrapper layout (transforms.py NSGT_SL): [*lead, F, S, M, 2]
    W = orc.nsgt_sl(x.reshape(1, 2, -1))
    assert [list(w.shape) for w in W] == gold["wrapper_shapes"].tolist()
    yw = orc.insgt_sl(W, x.shape[-1])
    assert yw.shape == (1, 2, x.shape[-1])
    assert snr_db(x, yw.reshape(2, -1)) > 150.0  # float64 oracle; limited by the fp32-stored Tukey window


def test_edge_lengths(orc, golden_dir):
    edge = np.load(os.path.join(golden_dir, "edge_lengths.npz"))
    for T in (1, 4515, 9030, 9031, 13545, 18060, 18061):
        assert orc.n_slices(T) == int(edge[f"S_{T}"])
        xe = (np.random.RandomState(T).rand(1, T).astype(np.float32) * 2 - 1)
        C = orc.forward(xe.astype(np.float64))
        assert C[0].shape[0] == int(edge[f"S_{T}"])
        E = np.asarray([float((np.abs(c) ** 2).sum()) for c in C])
        np.testing.assert_allclose(E, edge[f"E_{T}"], rtol=1e-4, atol=1e-9)
        y = orc.backward(C, T)
        np.testing.assert_allclose(y, edge[f"y_{T}"], atol=4e-6)


def test_gspi_config1(orc, golden_dir):
    """BASELINE.json configs[0]: round trip of .github/gspi.wav (mono, 262144 samples)."""
    gold = np.load(os.path.join(golden_dir, "gspi.npz"))
    x = (gold["wav_int16"].astype(np.float32) / 32768.0)[None, :]
    C = orc.forward(x.astype(np.float64))
    assert C[0].shape[0] == int(gold["S"]) == 31
    # high buckets of this band-limited signal are at the reference's own fp32 noise floor
    np.testing.assert_allclose([np.abs(c).sum() for c in C], gold["bucket_abs_sum"], rtol=1e-3, atol=1e-3)
    flat = common.pack(C).reshape(-1)[::61]
    scale = float(gold["bucket_max"].max())
    assert np.abs(flat - gold["coef_sample"]).max() / scale < 2e-6
    y = orc.backward(C, x.shape[-1])
    np.testing.assert_allclose(y.reshape(-1)[::7], gold["y_sample"], atol=4e-6)
    assert float(gold["snr_db"]) > 130.0          # the reference's own fp32 round trip
    assert snr_db(x, y) > 150.0                    # float64 oracle


def test_fp32_mode_matches_reference_snr(golden_dir):
    """The float32 mode of the oracle (used as timed CPU baseline) reconstructs at the
    reference's own fp32 quality."""
    gold = np.load(os.path.join(golden_dir, "gspi.npz"))
    o32 = SlicqOracle(**common.BARK, dtype=np.float32)
    x = (gold["wav_int16"][:60000].astype(np.float32) / 32768.0)[None, :]
    y = o32.backward(o32.forward(x), x.shape[-1])
    assert snr_db(x, y) > 125.0
