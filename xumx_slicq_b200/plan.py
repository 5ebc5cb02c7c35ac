"""Host-side filter-bank design for the sliced constant-Q transform (one-time plan setup).

Restates, in NumPy, what the reference derives when it constructs its transform object:

    xumx_slicq_v2/nsgt/fscale.py:5-89        Bark scale -> (f, q), suggested slice / transition length
    xumx_slicq_v2/nsgt/nsgfwin_sl.py:8-111   bin positions, per-bin lengths M_j, analysis windows g_j
    xumx_slicq_v2/nsgt/util.py:72-116        window ranges, canonical dual windows gd_j
    xumx_slicq_v2/nsgt/slicing.py:7-18       Tukey slicing window

The integer tables (M_j, rfbas_j, slice length) decide tensor shapes, and the reference derives
them from float32 tensors with round-half-to-even; this module follows the same float32 order of
operations so the shapes agree exactly ("integer-table fragility").
Everything here runs once per plan on the host; the hot path is in csrc/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

F32 = np.float32


# ---------------------------------------------------------------------------- frequency scales
class Scale:
    """fscale.py:5-53: a scale maps band index -> frequency F(b); Q(b) defaults to the central difference
    F dbnd / (F(b + dbnd) - F(b - dbnd)); (f, q) are float32 like the reference's tensors."""

    _DB = 1.0e-8

    def __init__(self, bnds: int):
        self.bnds = int(bnds)

    def __len__(self):
        return self.bnds

    def F(self, bnd):
        raise NotImplementedError

    def Q(self, bnd):
        return self.F(bnd) * self._DB / (self.F(bnd + self._DB) - self.F(bnd - self._DB))

    def __call__(self) -> Tuple[np.ndarray, np.ndarray]:
        idx = range(self.bnds)
        return (np.fromiter((self.F(b) for b in idx), dtype=F32, count=self.bnds),
                np.fromiter((self.Q(b) for b in idx), dtype=F32, count=self.bnds))

    def suggested_sllen_trlen(self, sr: float) -> Tuple[int, int]:
        """fscale.py:40-53."""
        f, q = self()
        need = int(np.ceil(np.max(q * F32(8.0) * F32(sr) / f)))
        sllen = need + (-need) % 4
        trlen = sllen // 4
        trlen += (-trlen) % 2
        return sllen, trlen


class BarkScale(Scale):
    """Bark frequency scale, fscale.py:56-89; Q by central difference, fscale.py:15-23."""

    def __init__(self, fmin: float, fmax: float, bnds: int, device=None):
        super().__init__(bnds)
        self.fmin, self.fmax = float(fmin), float(fmax)
        lo, hi = 6.0 * math.asinh(fmin / 600.0), 6.0 * math.asinh(fmax / 600.0)
        self._step = (hi - lo) / (bnds - 1)
        self._lo = lo

    def F(self, bnd):
        return 600.0 * math.sinh((bnd * self._step + self._lo) / 6.0)


class MelScale(Scale):
    """fscale.py:131-166 (Q by central difference: the reference's Q1 is not used)."""

    def __init__(self, fmin: float, fmax: float, bnds: int):
        super().__init__(bnds)
        self.fmin, self.fmax = float(fmin), float(fmax)
        mmin, mmax = math.log10(fmin / 700.0 + 1.0) * 2595.0, math.log10(fmax / 700.0 + 1.0) * 2595.0
        self.mbnd = (mmax - mmin) / (bnds - 1)
        self.mmin = mmin

    def F(self, bnd):
        return (math.pow(10.0, (bnd * self.mbnd + self.mmin) / 2595.0) - 1.0) * 700.0


class LogScale(Scale):
    """fscale.py:92-128: constant-Q ("cqlog", gamma = 0) and variable-Q with offset ("vqlog", gamma = fgamma)."""

    def __init__(self, fmin: float, fmax: float, bnds: int, gamma: float = 0.0):
        super().__init__(bnds)
        lfmin, lfmax = math.log2(fmin), math.log2(fmax)
        odiv = (lfmax - lfmin) / (bnds - 1)
        self.fmin, self.fmax = 2 ** lfmin, 2 ** lfmax
        self.pow2n = 2 ** odiv
        self.q = math.sqrt(self.pow2n) / (self.pow2n - 1.0) / 2.0
        self.gamma = float(gamma)

    def F(self, bnd):
        return self.fmin * self.pow2n ** bnd + self.gamma

    def Q(self, bnd):
        return self.q


class LinScale(Scale):
    """fscale.py:169-188."""

    def __init__(self, fmin: float, fmax: float, bnds: int):
        super().__init__(bnds)
        self.df = float(fmax - fmin) / (bnds - 1)
        self.fmin, self.fmax = float(fmin), float(fmax)
        if self.fmin <= 0:
            raise ValueError("Frequencies must be > 0.")

    def F(self, bnd):
        return bnd * self.df + self.fmin

    def Q(self, bnd):
        return self.F(bnd) / (self.df * 2)


def make_scale(name: str, fmin: float, fmax: float, fbins: int, fgamma: float = 15.0):
    """transforms.py:31-49."""
    if name == "bark":
        return BarkScale(fmin, fmax, fbins)
    if name == "mel":
        return MelScale(fmin, fmax, fbins)
    if name == "cqlog":
        return LogScale(fmin, fmax, fbins)
    if name == "vqlog":
        return LogScale(fmin, fmax, fbins, gamma=fgamma)
    if name == "linear":
        return LinScale(fmin, fmax, fbins)
    raise NotImplementedError(
        f"scale '{name}': bark, mel, cqlog, vqlog and linear are built; the reference's 576-band 'mrstft' scale is not")


# ---------------------------------------------------------------------------- windows
def _hann_periodic(n: int) -> np.ndarray:
    """util.py:5-11: 0.5 (1 + cos(2 pi i / n)), float64, peak at index 0."""
    return 0.5 * (1.0 + np.cos(np.arange(n, dtype=np.float64) * (2.0 * math.pi / n)))


def _blackman_harris_centered(n: int) -> np.ndarray:
    """util.py:14-46 (mod=True): 4-term window, evaluated in float32 like the reference, rotated so
    that the peak sits at index 0."""
    k = np.arange(n, dtype=F32)
    w = np.full(n, F32(0.35872), dtype=F32)
    for coef, harm in ((-0.48832, 2), (0.14128, 4), (-0.01168, 6)):
        w = w + F32(coef) * np.cos(k * F32(harm * math.pi / n))
    return np.roll(w.astype(F32), n // 2)


def tukey_window(sllen: int, trlen: int) -> np.ndarray:
    """slicing.py:7-18: zero | rising Hann half | one | falling Hann half | zero (float32)."""
    hh, htr = sllen // 4, trlen // 2
    hann = _hann_periodic(2 * trlen)
    tw = np.zeros(sllen, dtype=F32)
    tw[hh - htr: hh + htr] = hann[trlen:]
    tw[hh + htr: 3 * hh - htr] = 1.0
    tw[3 * hh - htr: 3 * hh + htr] = hann[:trlen]
    return tw


# ---------------------------------------------------------------------------- the plan
@dataclass
class SlicqTables:
    """Everything the CUDA library needs (include/slicq.h: slicq_tables) plus reference-visible
    attributes (fbins_actual, ncoefs, coef_factors)."""
    sllen: int
    trlen: int
    fs: float
    frqs: np.ndarray            # float32 scale frequencies (before DC/Nyquist are added)
    q: np.ndarray
    M_all: np.ndarray           # int32 [2*(lbas+1)]   all windows incl. the mirrored half
    rfbas_all: np.ndarray       # int32 [2*(lbas+1)]
    n_bins: int                 # J = lbas + 2 : DC, bins, Nyquist
    bin_M: np.ndarray           # int32 [J]
    bin_pos: np.ndarray         # int32 [J]
    win_fwd: np.ndarray         # float32 [sum M]  g_j[m], peak at m=0
    win_inv: np.ndarray         # float32 [sum M]  gd_j[m]
    tukey: np.ndarray           # float32 [L]
    buckets: List[Tuple[int, int, int]]  # (first_bin, n_bins, M)

    @property
    def hop(self) -> int:
        return self.sllen // 2

    @property
    def sum_M(self) -> int:
        return int(self.bin_M.sum())

    @property
    def ncoefs(self) -> int:
        return int(self.bin_M.max())

    def num_slices(self, n_samples: int) -> int:
        hh = self.sllen // 4
        return (-(-n_samples // hh) + 1) // 2 + 1

    def coef_factors(self) -> List[float]:
        return [float(m) / self.sllen for m in self.bin_M]


def design(scale, fs: float, sllen: int, trlen: int, min_win: int = 16, qvar: float = 1.0) -> SlicqTables:
    """nsgfwin (sliced=True) + calcwinrange + nsdual for a real-input transform (slicq.py:108-150)."""
    if sllen % 4 or trlen % 2:
        raise ValueError("sl_len must be a multiple of 4 and tr_area a multiple of 2")  # slicing.py:22-25
    if not (fs > 0 and sllen > 2 * trlen >= 0):
        raise ValueError("invalid slice / transition lengths")                           # slicq.py:86-90
    f, q = scale()
    f, q = f.astype(F32), q.astype(F32)
    nyq = fs / 2.0
    # the reference trims leading f<=0 and everything from the first f>=nyquist (nsgfwin_sl.py:21-30)
    first = int(np.argmax(f > 0))
    if first:
        f, q = f[first:], q[first:]
    last = int(np.argmax(f >= nyq))
    if last:
        f, q = f[:last], q[:last]
    if not (np.all(np.diff(f) > 0) and np.all(q > 0)):
        raise ValueError("scale must be increasing with positive Q")
    nb = len(f)

    # bin centres in FFT bins of one slice, positive then mirrored half (nsgfwin_sl.py:46-55)
    half = np.concatenate(([F32(0)], f, [F32(nyq)])).astype(F32)
    centres = np.concatenate((half, F32(fs) - half[-2:0:-1])).astype(F32) * F32(float(sllen) / fs)

    # per-bin length: bandwidth between neighbours, from Q at the two ends (nsgfwin_sl.py:57-82)
    width = np.zeros(len(centres), dtype=F32)
    width[0] = F32(2) * centres[1]
    width[1] = centres[1] / q[0]
    inner = np.r_[2:nb, nb + 1]
    width[inner] = centres[inner + 1] - centres[inner - 1]
    width[nb] = centres[nb] / q[nb - 1]
    width[nb + 2:] = width[nb:0:-1]
    M = np.rint(width * F32(qvar / 4.0)).astype(np.int32) * 4
    M = np.maximum(M, min_win).astype(np.int32)

    wins = [_blackman_harris_centered(int(m)) for m in M]
    # DC and Nyquist: plateau of ones with a Hann dip as wide as the neighbour (nsgfwin_sl.py:89-103)
    for j in (0, nb + 1):
        big, small = int(M[j]), int(M[j + 1])
        if big > small:
            w = np.ones(big, dtype=F32)
            lo = big // 2 - small // 2
            w[lo: lo + small] = _hann_periodic(small).astype(F32)
            wins[j] = w
    pos = np.rint(centres / F32(2)).astype(np.int32) * 2  # nsgfwin_sl.py:105

    # frame-operator diagonal over ALL windows (both halves), float64 (util.py:103-116)
    L = sllen
    if int((-pos[-1]) % L + (pos[-1] - pos[0])) != L:
        raise ValueError("window positions do not tile the slice spectrum")  # nn == Ls, nsgtf.py:44
    diag = np.zeros(L, dtype=np.float64)
    for g, m, c in zip(wins, M, pos):
        n = len(g)
        idx = (np.arange(n) + int(c)) % L                 # g[m] sits at position c + m~  (m~ = m or m-n)
        idx[n // 2:] = (np.arange(n // 2, n) - n + int(c)) % L
        diag[idx] += float(m) * g.astype(np.float64) ** 2
    J = nb + 2
    fw, iw = [], []
    for j in range(J):
        g = wins[j]
        n = len(g)
        mt = np.arange(n)
        mt[n // 2:] -= n
        d = diag[(int(pos[j]) + mt) % L]
        fw.append(g.astype(F32))
        iw.append((g.astype(np.float64) / d).astype(F32))

    buckets: List[Tuple[int, int, int]] = []
    for j in range(J):
        if buckets and buckets[-1][2] == int(M[j]):
            b = buckets[-1]
            buckets[-1] = (b[0], b[1] + 1, b[2])
        else:
            buckets.append((j, 1, int(M[j])))

    return SlicqTables(
        sllen=sllen, trlen=trlen, fs=float(fs), frqs=f, q=q, M_all=M, rfbas_all=pos, n_bins=J,
        bin_M=M[:J].copy(), bin_pos=pos[:J].copy(),
        win_fwd=np.concatenate(fw), win_inv=np.concatenate(iw),
        tukey=tukey_window(sllen, trlen), buckets=buckets)
