"""Deterministic recipes shared by make_golden.py (runs the reference) and the tests.

Nothing here touches /root/reference; it only fixes the *inputs* so that the
golden vectors can be regenerated and the tests can rebuild the same inputs.
"""
import numpy as np

BARK = dict(scale="bark", fbins=262, fmin=32.9, fmax=22050.0, fs=44100.0)

SMALL_T = 24000          # -> 4 slices
SMALL_ROWS = 2


def small_input():
    rs = np.random.RandomState(1234)
    return (rs.rand(SMALL_ROWS, SMALL_T).astype(np.float32) * 2.0 - 1.0)


def perturb(cseq, seed=4321):
    """Turn range coefficients into 'model-output-like' non-range coefficients:
    a smooth deterministic soft mask in [0.1,1) times the coefficient plus a
    small complex offset.  cseq: list of complex arrays (any leading dims)."""
    rs = np.random.RandomState(seed)
    out = []
    for c in cseq:
        mask = (0.1 + 0.9 * rs.rand(*c.shape)).astype(np.float32)
        off = (rs.randn(*c.shape) + 1j * rs.randn(*c.shape)).astype(np.complex64) * np.float32(1e-3)
        out.append((c * mask + off).astype(np.complex64))
    return out


def pack(cseq):
    """list of [S,N,F_b,M_b] -> [S,N,sum(F_b*M_b)] (bin-major, the reference's storage order)."""
    S, N = cseq[0].shape[:2]
    return np.concatenate([c.reshape(S, N, -1) for c in cseq], axis=-1)


def unpack(flat, buckets):
    """inverse of pack; buckets = list of (first_bin, n_bins, M)."""
    S, N = flat.shape[:2]
    out, o = [], 0
    for (_, nb, M) in buckets:
        out.append(flat[:, :, o:o + nb * M].reshape(S, N, nb, M))
        o += nb * M
    return out
