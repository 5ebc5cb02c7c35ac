"""Host-side plan builder (xumx_slicq_b200/plan.py) against the reference-generated tables."""
import os

import numpy as np
import pytest

import common
from xumx_slicq_b200 import plan as P


@pytest.fixture(scope="module")
def tab(golden_dir):
    return np.load(os.path.join(golden_dir, "tables_bark262.npz"))


@pytest.fixture(scope="module")
def tables():
    scl = P.BarkScale(32.9, 22050.0, 262)
    sllen, trlen = scl.suggested_sllen_trlen(44100.0)
    return P.design(scl, 44100.0, sllen, trlen)


def test_sizes(tables, tab):
    assert (tables.sllen, tables.trlen) == (18060, 4516) == (int(tab["sllen"]), int(tab["trlen"]))
    assert tables.n_bins == 263 == int(tab["fbins_actual"])
    assert tables.ncoefs == 292 == int(tab["ncoefs"])
    assert tables.sum_M == 18640
    assert len(tables.buckets) == 70
    assert tables.buckets[:3] == [(0, 1, 28), (1, 86, 16), (87, 14, 20)]


def test_integer_tables_exact(tables, tab):
    np.testing.assert_array_equal(tables.M_all.astype(np.int64), tab["M"])
    np.testing.assert_array_equal(tables.rfbas_all.astype(np.int64), tab["rfbas"])
    np.testing.assert_array_equal(tables.frqs, tab["frqs"][: len(tables.frqs)])  # last scale bin (== Nyquist) is dropped


def test_windows(tables, tab):
    np.testing.assert_allclose(tables.win_fwd, tab["g"], atol=1e-6)
    np.testing.assert_allclose(tables.win_inv, tab["gd"], rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(tables.tukey, tab["tukey"], atol=1e-7)
    np.testing.assert_allclose(tables.coef_factors(), tab["coef_factors"], rtol=1e-12)
    # fixture constants of SURVEY.md Appendix B
    assert abs(float(tables.win_fwd.astype(np.float64).sum()) - 6739.750248) < 1e-2
    assert abs(float(tables.win_inv.astype(np.float64).sum()) - 228.180945) < 1e-3
    assert float(tables.tukey.astype(np.float64).sum()) == pytest.approx(9030.0, abs=1e-3)


def test_num_slices(tables, golden_dir):
    edge = np.load(os.path.join(golden_dir, "edge_lengths.npz"))
    for T in (1, 4515, 9030, 9031, 13545, 18060, 18061):
        assert tables.num_slices(T) == int(edge[f"S_{T}"])
    assert [tables.num_slices(t) for t in (88200, 262144, 1323000, 7938000)] == [11, 31, 148, 881]


def test_alt_bark_integer_tables(golden_dir):
    alt = np.load(os.path.join(golden_dir, "tables_alt.npz"))
    i = 0
    while f"cfg{i}" in alt:
        fb, fmin = alt[f"cfg{i}"]
        scl = P.BarkScale(float(fmin), 22050.0, int(fb))
        sllen, trlen = scl.suggested_sllen_trlen(44100.0)
        t = P.design(scl, 44100.0, sllen, trlen)
        assert [sllen, trlen, t.n_bins] == list(alt[f"sl{i}"])
        np.testing.assert_array_equal(t.M_all.astype(np.int64), alt[f"M{i}"])
        np.testing.assert_array_equal(t.rfbas_all.astype(np.int64), alt[f"rfbas{i}"])
        i += 1
    assert i == 4


def test_argument_errors():
    scl = P.BarkScale(32.9, 22050.0, 262)
    with pytest.raises(ValueError):
        P.design(scl, 44100.0, 18062, 4516)   # not a multiple of 4 (slicing.py:24-25)
    with pytest.raises(ValueError):
        P.design(scl, 44100.0, 18060, 4515)   # odd transition (slicing.py:22-23)
    with pytest.raises(NotImplementedError):
        P.make_scale("mrstft", 32.9, 22050.0, 262)


def test_other_scales_integer_tables(golden_dir):
    """mel / cqlog / vqlog / linear (fscale.py:92-188, transforms.py:31-49) and more Bark
    configurations give exactly the reference's slice length, bin lengths and bin positions."""
    tab = np.load(os.path.join(golden_dir, "tables_scales.npz"))
    i = 0
    while f"cfg{i}" in tab:
        name = str(tab[f"name{i}"])
        fb, fmin = tab[f"cfg{i}"]
        scl = P.make_scale(name, float(fmin), 22050.0, int(fb))
        sllen, trlen = scl.suggested_sllen_trlen(44100.0)
        t = P.design(scl, 44100.0, sllen, trlen)
        assert [sllen, trlen, t.n_bins] == list(tab[f"sl{i}"]), (name, fb, fmin)
        np.testing.assert_array_equal(t.M_all.astype(np.int64), tab[f"M{i}"], err_msg=f"{name} {fb} {fmin}")
        np.testing.assert_array_equal(t.rfbas_all.astype(np.int64), tab[f"rfbas{i}"], err_msg=f"{name} {fb} {fmin}")
        i += 1
    assert i == 9
