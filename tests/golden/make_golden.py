"""Generate golden vectors by running the UNMODIFIED reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports xumx_slicq_v2 from /root/reference (read-only), runs its torch CPU path
on fixed inputs (tests/golden/common.py) and writes small .npz fixtures next to
this file.  The GPU box has no /root/reference: tests only read the .npz files.
"""
import os
import sys
import io
import contextlib
import warnings

import numpy as np

REF = os.environ.get("SLICQ_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
warnings.filterwarnings("ignore")

import torch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import common  # noqa: E402


def ref_base(**kw):
    from xumx_slicq_v2.transforms import NSGTBase
    with contextlib.redirect_stdout(io.StringIO()):
        return NSGTBase(kw["scale"], kw["fbins"], kw["fmin"], fmax=kw.get("fmax", 22050.0),
                        fs=kw.get("fs", 44100.0), device="cpu")


def tables(base):
    n = base.nsgt
    nb = n.fbins_actual
    from xumx_slicq_v2.nsgt.slicing import makewnd
    d = dict(
        sllen=np.int64(base.sllen), trlen=np.int64(base.trlen), nn=np.int64(n.nn),
        fbins_actual=np.int64(nb), ncoefs=np.int64(int(n.ncoefs)),
        frqs=n.frqs.numpy(), q=n.q.numpy(),
        M=n.M.numpy().astype(np.int64), rfbas=n.rfbas.numpy().astype(np.int64),
        g=np.concatenate([gi.numpy() for gi in n.g[:nb]]).astype(np.float32),
        gd=np.concatenate([gi.numpy() for gi in n.gd[:nb]]).astype(np.float64),
        wins0=np.asarray([int(w[0]) for w in n.wins], dtype=np.int64),
        tukey=makewnd(base.sllen, base.trlen).numpy(),
        coef_factors=np.asarray(n.coef_factors(), dtype=np.float64),
    )
    return d


def main():
    torch.manual_seed(0)
    from xumx_slicq_v2.transforms import make_filterbanks, ComplexNorm

    base = ref_base(**common.BARK)
    np.savez_compressed(os.path.join(HERE, "tables_bark262.npz"), **tables(base))

    # integer tables for a few other Bark configs (plan-builder robustness)
    alt = {}
    for i, (fb, fmin) in enumerate([(100, 50.0), (64, 100.0), (200, 40.0), (300, 32.9)]):
        b = ref_base(scale="bark", fbins=fb, fmin=fmin)
        alt[f"cfg{i}"] = np.asarray([fb, fmin], dtype=np.float64)
        alt[f"sl{i}"] = np.asarray([b.sllen, b.trlen, b.nsgt.fbins_actual], dtype=np.int64)
        alt[f"M{i}"] = b.nsgt.M.numpy().astype(np.int64)
        alt[f"rfbas{i}"] = b.nsgt.rfbas.numpy().astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "tables_alt.npz"), **alt)

    nsgt, insgt = make_filterbanks(base)
    buckets = []
    # ---- small random input: forward, inverse of perturbed coefficients ----
    x = common.small_input()
    C = base.nsgt.forward((torch.from_numpy(x),))         # list [S,N,F,M] complex64
    Cn = [c.numpy().copy() for c in C]
    j = 0
    for c in Cn:
        buckets.append((j, c.shape[2], c.shape[3]))
        j += c.shape[2]
    y_rt = base.nsgt.backward([torch.from_numpy(c.copy()) for c in Cn], x.shape[-1]).numpy()
    P = common.perturb(Cn)
    y_p = base.nsgt.backward([torch.from_numpy(c.copy()) for c in P], x.shape[-1]).numpy()
    # wrapper-level shapes
    Xw = nsgt(torch.from_numpy(x).view(1, 2, -1))
    np.savez_compressed(
        os.path.join(HERE, "small_fwdinv.npz"),
        buckets=np.asarray(buckets, dtype=np.int64),
        coefs=common.pack(Cn).astype(np.complex64),
        y_roundtrip=y_rt.astype(np.float32),
        y_perturbed=y_p.astype(np.float32),
        wrapper_shapes=np.asarray([list(t.shape) for t in Xw], dtype=np.int64),
    )

    # ---- edge-case lengths: slice counts and outputs ----
    edge = {}
    for T in (1, 4515, 9030, 9031, 13545, 18060, 18061):
        xe = (np.random.RandomState(T).rand(1, T).astype(np.float32) * 2 - 1)
        Ce = base.nsgt.forward((torch.from_numpy(xe),))
        S = Ce[0].shape[0]
        ye = base.nsgt.backward([c.clone() for c in Ce], T).numpy()
        edge[f"S_{T}"] = np.int64(S)
        edge[f"y_{T}"] = ye.astype(np.float32)
        # per-bucket energy as a compact fingerprint of the forward
        edge[f"E_{T}"] = np.asarray([float((c.abs() ** 2).sum()) for c in Ce], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "edge_lengths.npz"), **edge)

    # ---- gspi.wav (config 1): keep the int16 samples + fingerprints ----
    from scipy.io import wavfile
    sr, w = wavfile.read(os.path.join(REF, ".github", "gspi.wav"))
    assert sr == 44100 and w.dtype == np.int16
    xg = (w.astype(np.float32) / 32768.0)[None, :]
    Cg = base.nsgt.forward((torch.from_numpy(xg),))
    Cgn = [c.numpy().copy() for c in Cg]
    flat = common.pack(Cgn)
    yg = base.nsgt.backward([c.clone() for c in Cg], xg.shape[-1]).numpy()
    num = float(np.sum(xg.astype(np.float64) ** 2))
    den = float(np.sum((yg.astype(np.float64) - xg) ** 2))
    np.savez_compressed(
        os.path.join(HERE, "gspi.npz"),
        wav_int16=w, S=np.int64(flat.shape[0]),
        bucket_abs_sum=np.asarray([float(np.abs(c).sum()) for c in Cgn], dtype=np.float64),
        bucket_max=np.asarray([float(np.abs(c).max()) for c in Cgn], dtype=np.float64),
        coef_sample=flat.reshape(-1)[::61].astype(np.complex64),
        y_sample=yg.reshape(-1)[::7].astype(np.float32),
        snr_db=np.float64(10 * np.log10(num / den)),
    )
    print("golden vectors written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
