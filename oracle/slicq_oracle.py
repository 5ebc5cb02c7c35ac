"""CPU oracle for the sliCQT analysis/synthesis path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference's algorithm for the hot path
named by BASELINE.json:north_star.  It is *not* product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the CPU baseline.
The product package (``xumx_slicq_b200``) never imports anything from here.

Parity pinning: the reference ships no test / golden vector for this path
(SURVEY.md §4), so the oracle is pinned against outputs of the *reference
itself*, executed in the build container from ``/root/reference`` by
``tests/golden/make_golden.py`` (committed) and stored as
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every function
here against those vectors.

Every function cites the reference file:line (relative to the upstream repo
root, ``xumx_slicq_v2/...``) that it restates.  The stage structure follows the
reference on purpose (rotated slices, ``arrange`` roll, per-bin loop, mirrored
bin pass, float64 overlap-add) -- the closed forms the CUDA kernels use
 are derived from, and tested against, this restatement.

dtype: computations run in float64 by default (``dtype=np.float64``); pass
``np.float32`` to mimic the reference's storage precision (used for the timed
CPU baseline so that it moves the same number of bytes as the reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

try:  # scipy's pocketfft is multi-threaded (workers=); numpy's is the fallback
    import scipy.fft as _fft

    def _fft_fwd(a, workers=None):
        return _fft.fft(a, axis=-1, workers=workers)

    def _fft_inv(a, workers=None):
        return _fft.ifft(a, axis=-1, workers=workers)

    def _irfft(a, n, workers=None):
        return _fft.irfft(a, n=n, axis=-1, workers=workers)

except Exception:  # pragma: no cover
    def _fft_fwd(a, workers=None):
        return np.fft.fft(a, axis=-1)

    def _fft_inv(a, workers=None):
        return np.fft.ifft(a, axis=-1)

    def _irfft(a, n, workers=None):
        return np.fft.irfft(a, n=n, axis=-1)


# --------------------------------------------------------------------------
# frequency scale  (nsgt/fscale.py)
# --------------------------------------------------------------------------

class BarkScale:
    """nsgt/fscale.py:56-89 (BarkScale) + :5-53 (Scale base class)."""

    dbnd = 1.0e-8  # fscale.py:6

    def __init__(self, fmin: float, fmax: float, bnds: int):
        # fscale.py:69-80
        bmin = self.hz2bark(fmin)
        bmax = self.hz2bark(fmax)
        self.bnds = bnds
        self.fmin = float(fmin)
        self.fmax = float(fmax)
        self.bbnd = (bmax - bmin) / (bnds - 1)
        self.bmin = bmin
        self.bmax = bmax

    @staticmethod
    def hz2bark(f):  # fscale.py:58-61
        return 6 * math.asinh(f / 600)

    @staticmethod
    def bark2hz(b):  # fscale.py:63-67
        return 600 * math.sinh(b / 6)

    def F(self, bnd):  # fscale.py:86-89
        return self.bark2hz(bnd * self.bbnd + self.bmin)

    def Q(self, bnd):  # fscale.py:15-23 numerical differentiation
        return self.F(bnd) * self.dbnd / (self.F(bnd + self.dbnd) - self.F(bnd - self.dbnd))

    def __call__(self) -> Tuple[np.ndarray, np.ndarray]:
        # fscale.py:25-38: python-float values stored as float32 tensors
        f = np.asarray([self.F(b) for b in range(self.bnds)], dtype=np.float32)
        q = np.asarray([self.Q(b) for b in range(self.bnds)], dtype=np.float32)
        return f, q

    def suggested_sllen_trlen(self, sr: float) -> Tuple[int, int]:
        # fscale.py:40-53 (float32 arithmetic like the torch tensors)
        f, q = self()
        Ls = int(np.ceil(np.max((q * np.float32(8.0) * np.float32(sr)) / f)))
        Ls = Ls + -Ls % 4
        sllen = Ls
        trlen = sllen // 4
        trlen = trlen + -trlen % 2
        return sllen, trlen


# --------------------------------------------------------------------------
# windows  (nsgt/util.py)
# --------------------------------------------------------------------------

def hannwin(l: int) -> np.ndarray:
    """nsgt/util.py:5-11 (float64)."""
    r = np.arange(l, dtype=np.float64)
    r *= math.pi * 2.0 / l
    r = np.cos(r)
    r += 1.0
    r *= 0.5
    return r


def blackharr(n: int) -> np.ndarray:
    """nsgt/util.py:14-46 with mod=True, l=n.

    The reference evaluates this in float32 (int64 arange * python float ->
    float32).  We evaluate in float32 too so windows agree to the last bit or
    two; tests compare against the reference with a 1e-6 tolerance.
    """
    nn = (n // 2) * 2
    k = np.arange(n).astype(np.float32)
    bh = (
        np.float32(0.35872)
        - np.float32(0.48832) * np.cos(k * np.float32(2 * math.pi / nn))
        + np.float32(0.14128) * np.cos(k * np.float32(4 * math.pi / nn))
        - np.float32(0.01168) * np.cos(k * np.float32(6 * math.pi / nn))
    ).astype(np.float32)
    # util.py:45: peak moved to index 0
    bh = np.hstack((bh[-(n // 2):], bh[: -(n // 2)]))
    return bh


# --------------------------------------------------------------------------
# filter design  (nsgt/nsgfwin_sl.py, nsgt/util.py calcwinrange / nsdual)
# --------------------------------------------------------------------------

def nsgfwin(f: np.ndarray, q: np.ndarray, sr: float, Ls: int, min_win: int = 16, Qvar: float = 1.0):
    """nsgt/nsgfwin_sl.py:8-111, sliced=True branch, float32 arithmetic."""
    f = np.asarray(f, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    nf = sr / 2.0
    # :21-30 drop f<=0 / f>=nyquist
    lim = int(np.argmax(f > 0))
    if lim != 0:
        f = f[lim:]
        q = q[lim:]
    lim = int(np.argmax(f >= nf))
    if lim != 0:
        f = f[:lim]
        q = q[:lim]
    assert len(f) == len(q)
    assert np.all((f[1:] - f[:-1]) > 0)
    assert np.all(q > 0)

    fbas = f
    lbas = len(fbas)
    # :46-55
    frqs = np.zeros(lbas + 2, dtype=np.float32)
    frqs[0] = 0.0
    frqs[1:-1] = fbas
    frqs[-1] = nf
    fbas = np.concatenate((frqs, (np.float32(sr) - frqs[::-1][1:-1]).astype(np.float32))).astype(np.float32)
    fbas = (fbas * np.float32(float(Ls) / sr)).astype(np.float32)

    # :57-72
    M = np.zeros(fbas.shape, dtype=np.float32)
    M[0] = np.float32(2) * fbas[1]
    M[1] = fbas[1] / q[0]
    for k in list(range(2, lbas)) + [lbas + 1]:
        M[k] = fbas[k + 1] - fbas[k - 1]
    M[lbas] = fbas[lbas] / q[lbas - 1]
    M[lbas + 2: 2 * (lbas + 1)] = M[1: lbas + 1][::-1]
    M = (M * np.float32(Qvar / 4.0)).astype(np.float32)
    M = np.round(M).astype(np.int32)  # torch.round == half-to-even == np.round
    M *= 4
    # :82
    M = np.clip(M, min_win, None)

    # :84-85
    g = [blackharr(int(m)) for m in M]

    # :89-103 plateau windows for DC / Nyquist
    for kk in (1, lbas + 2):
        if M[kk - 1] > M[kk]:
            a = int(M[kk - 1])
            b = int(M[kk])
            w = np.ones(a, dtype=np.float32)
            w[a // 2 - b // 2: a // 2 + int(math.ceil(b / 2.0))] = hannwin(b).astype(np.float32)
            g[kk - 1] = w

    # :105
    rfbas = (np.round(fbas / np.float32(2.0)).astype(np.int32)) * 2
    return g, rfbas, M


def calcwinrange(g: Sequence[np.ndarray], rfbas: np.ndarray, Ls: int):
    """nsgt/util.py:72-100."""
    shift = np.concatenate((((-rfbas[-1]) % Ls,), rfbas[1:] - rfbas[:-1])).astype(np.int64)
    timepos = np.cumsum(shift)
    nn = int(timepos[-1])
    timepos = timepos - shift[0]
    wins = []
    for gii, tpii in zip(g, timepos):
        Lg = len(gii)
        win_range = np.arange(-(Lg // 2) + tpii, Lg - (Lg // 2) + tpii, dtype=np.int64)
        win_range %= nn
        wins.append(win_range)
    return wins, nn


def nsdual(g: Sequence[np.ndarray], wins: Sequence[np.ndarray], nn: int, M: np.ndarray):
    """nsgt/util.py:103-116 (float64 accumulation of the frame diagonal)."""
    x = np.zeros(nn, dtype=np.float64)
    for gi, mii, sl in zip(g, M, wins):
        xa = np.square(np.fft.fftshift(gi.astype(np.float64)))
        xa *= float(mii)
        x[sl] += xa
    gd = [gi.astype(np.float64) / np.fft.ifftshift(x[wi]) for gi, wi in zip(g, wins)]
    return gd


# --------------------------------------------------------------------------
# the sliced transform object  (nsgt/slicq.py NSGT_sliced)
# --------------------------------------------------------------------------

@dataclass
class SlicqOracle:
    """nsgt/slicq.py:70-151 with real=True, multichannel=True, reducedform=0,
    recwnd=False, min_win=16, Qvar=1 (what transforms.py:60-68 passes)."""

    scale: str = "bark"
    fbins: int = 262
    fmin: float = 32.9
    fmax: float = 22050.0
    fs: float = 44100.0
    dtype: type = np.float64
    workers: int | None = None

    frqs: np.ndarray = field(init=False)
    q: np.ndarray = field(init=False)

    def __post_init__(self):
        if self.scale != "bark":
            raise ValueError("oracle restates the Bark scale only (north_star)")
        scl = BarkScale(self.fmin, self.fmax, self.fbins)
        self.scl = scl
        self.frqs, self.q = scl()
        self.sl_len, self.tr_area = scl.suggested_sllen_trlen(self.fs)  # transforms.py:54
        assert self.sl_len % 4 == 0 and self.tr_area % 2 == 0  # slicq.py:93-94
        self.g, self.rfbas, self.M = nsgfwin(self.frqs, self.q, self.fs, self.sl_len, min_win=16)
        # slicq.py:123-131
        self.nbins = len(self.g) // 2 + 1
        self.fbins_actual = self.nbins
        # slicq.py:134-137
        self.ncoefs = max(int(math.ceil(float(len(gii)) / mii)) * mii
                          for mii, gii in zip(self.M[: self.nbins], self.g[: self.nbins]))
        self.wins, self.nn = calcwinrange(self.g, self.rfbas, self.sl_len)
        self.gd = nsdual(self.g, self.wins, self.nn, self.M)
        # bucket structure (nsgtf.py:66-78: consecutive bins of equal length)
        self.buckets: List[Tuple[int, int, int]] = []  # (first_bin, n_bins, M)
        for j in range(self.nbins):
            Lg = len(self.g[j])
            if self.buckets and self.buckets[-1][2] == Lg:
                b = self.buckets[-1]
                self.buckets[-1] = (b[0], b[1] + 1, b[2])
            else:
                self.buckets.append((j, 1, Lg))

    # -- helpers ---------------------------------------------------------
    @property
    def coef_factor(self) -> float:  # slicq.py:232-234
        return float(self.ncoefs) / self.sl_len

    def coef_factors(self) -> List[float]:  # slicq.py:236-243
        return [float(int(math.ceil(float(len(gii)) / mii)) * mii) / self.sl_len
                for mii, gii in zip(self.M[: self.nbins], self.g[: self.nbins])]

    def n_slices(self, T: int) -> int:
        """Number of slices the generator in slicing.py:21-72 emits for T samples."""
        hhop = self.sl_len // 4
        nblk = -(-T // hhop)  # reblock(fulllast=True): ceil
        return (nblk + 5 - 4) // 2 + 1

    # -- stage 1: slicing (nsgt/slicing.py) --------------------------------
    def tukey(self) -> np.ndarray:
        """nsgt/slicing.py:7-18 makewnd (float64 hann cast to float32 storage)."""
        L, tr = self.sl_len, self.tr_area
        hhop, htr = L // 4, tr // 2
        w = hannwin(2 * tr)
        tw = np.empty(L, dtype=np.float32)
        tw[: hhop - htr] = 0
        tw[hhop - htr: hhop + htr] = w[tr:]
        tw[hhop + htr: 3 * hhop - htr] = 1
        tw[3 * hhop - htr: 3 * hhop + htr] = w[:tr]
        tw[3 * hhop + htr:] = 0
        return tw

    def slicing(self, x: np.ndarray) -> np.ndarray:
        """nsgt/slicing.py:21-72 + pack of slicq.py:47-63.  x: [N,T] -> [S,N,L]
        with the reference's alternating quarter rotation."""
        x = np.asarray(x)
        N, T = x.shape
        L = self.sl_len
        hhop = L // 4
        tw = self.tukey().astype(self.dtype)
        nblk = -(-T // hhop)
        # 2 leading zero blocks, data (zero-padded to block multiple), 3 trailing
        buf = np.zeros((N, (nblk + 5) * hhop), dtype=self.dtype)
        buf[:, 2 * hhop: 2 * hhop + T] = x
        S = (nblk + 5 - 4) // 2 + 1
        out = np.empty((S, N, L), dtype=self.dtype)
        twq = [tw[o: o + hhop] for o in range(0, L, hhop)]
        for s in range(S):
            kpar = s % 2  # slicing.py:51-58: cycle of two quarter permutations
            for i in range(4):
                dst = (i + 3 - kpar * 2) % 4
                blk = buf[:, (2 * s + i) * hhop: (2 * s + i + 1) * hhop]
                out[s, :, dst * hhop: (dst + 1) * hhop] = blk * twq[i]
        return out

    # -- stage 2: forward core (nsgt/nsgtf.py) -----------------------------
    def nsgtf_sl(self, f_slices: np.ndarray) -> List[np.ndarray]:
        """nsgt/nsgtf.py:7-84.  [S,N,L] real -> list of [S,N,F_b,M_b] complex."""
        assert f_slices.shape[-1] == self.nn  # nsgtf.py:44
        cdt = np.complex128 if self.dtype == np.float64 else np.complex64
        ft = _fft_fwd(f_slices.astype(self.dtype), workers=self.workers).astype(cdt)  # :40
        ret = []
        for (j0, nb, Lg) in self.buckets:
            c = np.zeros(f_slices.shape[:2] + (nb, Lg), dtype=cdt)
            for jj in range(nb):
                j = j0 + jj
                gsh = np.fft.fftshift(self.g[j]).astype(self.dtype)
                t = ft[:, :, self.wins[j]] * gsh  # :55
                c[:, :, jj, : (Lg + 1) // 2] = t[:, :, Lg // 2:]  # :60
                c[:, :, jj, -(Lg // 2):] = t[:, :, : Lg // 2]  # :63
            ret.append(_fft_inv(c, workers=self.workers).astype(cdt))  # :69,:81
        return ret

    @staticmethod
    def arrange(cseq: List[np.ndarray], fwd: bool) -> List[np.ndarray]:
        """nsgt/slicq.py:13-33 (returns new arrays; the reference works in place)."""
        out = []
        for c in cseq:
            M = c.shape[-1]
            if fwd:
                odd_mid, even_mid = M // 4, 3 * M // 4
            else:
                odd_mid, even_mid = 3 * M // 4, M // 4
            c = c.copy()
            c[1::2] = np.concatenate((c[1::2, :, :, odd_mid:], c[1::2, :, :, :odd_mid]), axis=-1)
            c[::2] = np.concatenate((c[::2, :, :, even_mid:], c[::2, :, :, :even_mid]), axis=-1)
            out.append(c)
        return out

    def forward(self, x: np.ndarray) -> List[np.ndarray]:
        """nsgt/slicq.py:182-196.  x [N,T] -> list of [S,N,F_b,M_b] complex."""
        return self.arrange(self.nsgtf_sl(self.slicing(x)), True)

    # -- stage 3: inverse core (nsgt/nsigtf.py) ----------------------------
    def nsigtf_sl(self, cseq: List[np.ndarray]) -> np.ndarray:
        """nsgt/nsigtf.py:5-106.  list of [S,N,F_b,M_b] -> [S,N,L] real."""
        cdt = np.complex128 if self.dtype == np.float64 else np.complex64
        fc_list = [_fft_fwd(c.astype(cdt), workers=self.workers).astype(cdt) for c in cseq]  # :29-33
        S, N = cseq[0].shape[:2]
        nfreqs = sum(c.shape[2] for c in cseq)
        fr = np.zeros((S, N, self.nn), dtype=cdt)  # :35
        nwin = len(self.gd)
        fbin_ptr = 0
        mfbin_ptr = nwin
        for fc in fc_list:
            nb = fc.shape[2]
            for jj in range(nb):
                freq_idx = fbin_ptr + jj
                rr = 1 if freq_idx == 0 or freq_idx == nfreqs - 1 else 2  # :60
                for k in range(rr):
                    t = fc[:, :, jj]
                    if k == 1:  # :67-80 mirrored negative-frequency bin
                        mfbin_ptr -= 1
                        freq_idx = mfbin_ptr
                        t = np.conj(np.concatenate((t[:, :, 1:], t[:, :, 1:][:, :, ::-1]), axis=2))
                    Lg = len(self.gd[freq_idx])
                    wr = self.wins[freq_idx]
                    wr1 = wr[: Lg // 2]
                    wr2 = wr[-((Lg + 1) // 2):]
                    r = (Lg + 1) // 2
                    l = Lg // 2
                    temp = np.empty((S, N, Lg), dtype=cdt)
                    temp[:, :, :r] = t[:, :, :r]
                    temp[:, :, Lg - l: Lg] = t[:, :, Lg - l: Lg]
                    temp *= self.gd[freq_idx].astype(self.dtype)  # :91
                    temp *= Lg  # :92
                    fr[:, :, wr1] += temp[:, :, Lg - l: Lg]  # :94
                    fr[:, :, wr2] += temp[:, :, :r]  # :95
            fbin_ptr += nb
        ftr = fr[:, :, : self.nn // 2 + 1]  # :99
        return _irfft(ftr, n=self.sl_len, workers=self.workers).astype(self.dtype)  # :103

    # -- stage 4: unslicing (nsgt/unslicing.py + slicq.py:207-230) ---------
    def unslicing(self, frec: np.ndarray, length: int) -> np.ndarray:
        """nsgt/unslicing.py:6-69 (usewindow=False) followed by slicq.py:218
        (drop 2 blocks) and :221-229 (concatenate, truncate to ``length``).
        Accumulates in float64 like the reference and returns float32/64."""
        S, N, L = frec.shape
        hhop = L // 4
        fq = frec.reshape(S, N, 4, hhop)
        quads = np.empty_like(fq)
        # unslicing.py:19-28 undo the quarter rotation
        quads[::2, :, 0] = fq[::2, :, 3]
        quads[::2, :, 1] = fq[::2, :, 0]
        quads[::2, :, 2] = fq[::2, :, 1]
        quads[::2, :, 3] = fq[::2, :, 2]
        quads[1::2, :, 0] = fq[1::2, :, 1]
        quads[1::2, :, 1] = fq[1::2, :, 2]
        quads[1::2, :, 2] = fq[1::2, :, 3]
        quads[1::2, :, 3] = fq[1::2, :, 0]
        # unslicing.py:50-69 overlap-add in float64; slice s covers blocks 2s..2s+3
        acc = np.zeros((N, (2 * S + 2) * hhop), dtype=np.float64)
        for s in range(S):
            acc[:, 2 * s * hhop: (2 * s + 4) * hhop] += quads[s].reshape(N, L).astype(np.float64)
        sig = acc[:, 2 * hhop:]  # slicq.py:218
        return sig[:, :length].astype(self.dtype)  # reblock(fulllast=False) first block

    def backward(self, cseq: List[np.ndarray], length: int) -> np.ndarray:
        """nsgt/slicq.py:198-230.  list of [S,N,F_b,M_b] -> [N,length]."""
        return self.unslicing(self.nsigtf_sl(self.arrange(cseq, False)), length)

    # -- wrappers (transforms.py) -----------------------------------------
    def nsgt_sl(self, x: np.ndarray) -> List[np.ndarray]:
        """transforms.py:106-131 NSGT_SL.forward: x[*lead,T] ->
        list of [*lead, F_b, S, M_b, 2] real."""
        lead = x.shape[:-1]
        C = self.forward(x.reshape(-1, x.shape[-1]))
        out = []
        for c in C:
            c = np.moveaxis(c, 0, -2)  # [N,F,S,M]
            c = np.stack((c.real, c.imag), axis=-1)
            out.append(c.reshape(lead + c.shape[-4:]))
        return out

    def insgt_sl(self, X_list: List[np.ndarray], length: int) -> np.ndarray:
        """transforms.py:154-178 INSGT_SL.forward."""
        cs = []
        lead = None
        for X in X_list:
            Xc = X[..., 0] + 1j * X[..., 1]
            lead = Xc.shape[:-3]
            Xc = Xc.reshape((-1,) + Xc.shape[-3:])  # [N,F,S,M]
            cs.append(np.moveaxis(Xc, -2, 0))  # [S,N,F,M]
        y = self.backward(cs, length)
        return y.reshape(lead + (-1,))


def snr_db(ref: np.ndarray, est: np.ndarray) -> float:
    """Round-trip SNR used throughout tests/bench: 10 log10(sum x^2 / sum (y-x)^2)."""
    ref = np.asarray(ref, dtype=np.float64)
    est = np.asarray(est, dtype=np.float64)
    num = np.sum(ref * ref)
    den = np.sum((est - ref) ** 2)
    return float(10.0 * np.log10(num / max(den, 1e-300)))
