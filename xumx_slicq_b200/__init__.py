"""B200-native sliced constant-Q transform (sliCQT): drop-in for the analysis/synthesis path of
sevagh/xumx-sliCQ-V2 (`xumx_slicq_v2/nsgt` + `xumx_slicq_v2/transforms.py`) on hand-written
sm_100a CUDA kernels behind a C-ABI (include/slicq.h).  See DESIGN.md."""
from .transforms import NSGTBase, NSGT_SL, INSGT_SL, ComplexNorm, make_filterbanks  # noqa: F401
from .nsgt import NSGT_sliced  # noqa: F401
from .plan import BarkScale  # noqa: F401

__all__ = ["NSGTBase", "NSGT_SL", "INSGT_SL", "ComplexNorm", "make_filterbanks", "NSGT_sliced", "BarkScale"]
