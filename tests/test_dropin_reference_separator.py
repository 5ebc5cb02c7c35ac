"""Drop-in proof (CPU tier, build container only): the UNMODIFIED reference Separator + Unmix (seeded
random weights -- the shipped .pth files are Git-LFS pointers) run once on the reference's own torch
transforms and once on this repository's wrappers (host-emulated kernels); the separated stems must
agree.  Skipped where /root/reference does not exist (the GPU box)."""
import contextlib
import io
import os
import sys

import pytest
import torch

REF = os.environ.get("SLICQ_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "xumx_slicq_v2")),
                                reason="reference checkout not available")


@pytest.fixture(scope="module")
def emu():
    from tests.emu.emu_backend import EmuBackend
    import xumx_slicq_b200.nsgt as nsgt_mod
    old = nsgt_mod._BACKEND
    nsgt_mod._BACKEND = EmuBackend()
    yield
    nsgt_mod._BACKEND = old


def test_reference_separator_runs_on_the_dropin_transforms(emu):
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            from xumx_slicq_v2.transforms import NSGTBase as RefBase, make_filterbanks as ref_filterbanks, ComplexNorm as RefNorm
            from xumx_slicq_v2.model import Unmix
            from xumx_slicq_v2.separator import Separator
            from xumx_slicq_b200 import NSGTBase, make_filterbanks, ComplexNorm
            rbase = RefBase("bark", 262, 32.9, device="cpu")
            rn, ri = ref_filterbanks(rbase)
            base = NSGTBase("bark", 262, 32.9, device="cpu")
            n, i = make_filterbanks(base)
            torch.manual_seed(0)
            sample = RefNorm()(rbase.predict_input_size(1, 2, 2.0)[0])
            model = Unmix(sample, realtime=True)
            model.freeze()
    finally:
        sys.path.remove(REF)
    sep_ref = Separator(xumx_model=model, encoder=(rn, ri, RefNorm()), device="cpu", quiet=True)
    sep_new = Separator(xumx_model=model, encoder=(n, i, ComplexNorm()), device="cpu", quiet=True)
    torch.manual_seed(1)
    audio = torch.rand(1, 2, 30000) * 2 - 1
    with torch.no_grad():
        out_ref = sep_ref(audio)
        out_new = sep_new(audio)
    assert out_new.shape == out_ref.shape == (4, 1, 2, 30000)
    scale = float(out_ref.abs().max())
    assert float((out_new - out_ref).abs().max()) <= 2e-5 * scale
    # the same shapes come out of the shape probe the reference uses to build the model
    Xs, _ = base.predict_input_size(1, 2, 0.5)
    Xr, _ = rbase.predict_input_size(1, 2, 0.5)
    assert [tuple(a.shape) for a in Xs] == [tuple(b.shape) for b in Xr]
