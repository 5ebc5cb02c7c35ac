"""Host-buffer streaming of the sliCQT path: overlap PCIe transfers with the kernels.

The reference's `Separator.forward` (separator.py:133-232) takes a device tensor and processes
chunks one after the other; with HOST audio in and HOST stems out, the PCIe copies (10.6 MB in,
42 MB out per 30 s stereo mixture) cost more than the B200 kernels.  `TransformStream` runs a
sequence of batches through three CUDA streams -- H2D copy, compute (NSGT_SL -> user model ->
INSGT_SL), D2H copy -- with double-buffered pinned host and device buffers, so that batch i+1 is
uploaded and batch i-1 downloaded while batch i is transformed.  Results are identical to calling
the wrappers directly; only the scheduling differs.
"""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, List, Optional

import torch

from .transforms import NSGTBase, make_filterbanks


class TransformStream:
    def __init__(self, nsgt_base: NSGTBase, model: Callable[[List[torch.Tensor]], List[torch.Tensor]],
                 device: Optional[torch.device] = None, depth: int = 2):
        """model: maps the ragged list X (per bucket [B,C,F,S,M,2]) to the list handed to INSGT_SL
        (e.g. the 4 target estimates [4,B,C,F,S,M,2] of xumx_slicq_v2.model.Unmix)."""
        self.base = nsgt_base
        self.device = torch.device(device) if device is not None else nsgt_base.nsgt.device
        self.nsgt, self.insgt = make_filterbanks(nsgt_base)
        self.model = model
        self.depth = max(2, int(depth))
        self.s_in = torch.cuda.Stream(self.device)
        self.s_cmp = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._xd: List[Optional[torch.Tensor]] = [None] * self.depth
        self._yd: List[Optional[torch.Tensor]] = [None] * self.depth
        # one more host slot than batches in flight: a result stays valid while the next `depth - 1` are produced
        self._yh: List[Optional[torch.Tensor]] = [None] * (self.depth + 1)

    def _buf(self, store, i, shape, device, pinned=False):
        t = store[i]
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=torch.float32, device=device)
            if pinned:
                t = t.pin_memory()
            store[i] = t
        return t

    def process(self, batches: Iterable[torch.Tensor]) -> Iterator[torch.Tensor]:
        """batches: pinned host tensors [B, C, T] float32.  Yields pinned host tensors [targets, B, C, T] with the inverse
        transform of the model output.  A yielded tensor stays valid until the generator has been advanced `depth` more
        times (it is one of depth + 1 rotating pinned buffers): copy it if it has to live longer."""
        nh = self.depth + 1
        ev_in = [torch.cuda.Event() for _ in range(self.depth)]
        ev_cmp = [torch.cuda.Event() for _ in range(self.depth)]
        ev_out = [torch.cuda.Event() for _ in range(nh)]
        pending = []          # (device slot, host slot) whose output copy has been queued, oldest first
        for i, xh in enumerate(batches):
            b, hb = i % self.depth, i % nh
            T = xh.shape[-1]
            # device slot b was used by batch i - depth: its D2H has to be over before the buffers are overwritten
            while pending and (pending[0][0] == b or len(pending) >= self.depth):
                pb, ph = pending.pop(0)
                ev_out[ph].synchronize()
                yield self._yh[ph]
            xd = self._buf(self._xd, b, xh.shape, self.device)
            with torch.cuda.stream(self.s_in):
                xd.copy_(xh, non_blocking=True)
                ev_in[b].record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(ev_in[b])
                X = self.nsgt(xd)
                Y = self.model(X)
                lead = tuple(Y[0].shape[:-4])
                yd = self._buf(self._yd, b, lead + (T,), self.device)
                try:
                    self.insgt(Y, T, out=yd)        # straight into the staging buffer: no extra device copy
                except ValueError:
                    yd.copy_(self.insgt(Y, T))      # model outputs that are not contiguous float32
                ev_cmp[b].record(self.s_cmp)
                del X, Y
            yh = self._buf(self._yh, hb, yd.shape, "cpu", pinned=True)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_cmp[b])
                yh.copy_(yd, non_blocking=True)
                ev_out[hb].record(self.s_out)
            pending.append((b, hb))
        for pb, ph in pending:
            ev_out[ph].synchronize()
            yield self._yh[ph]


class StreamedSeparator:
    """Chunked demix of a long signal with overlap-exact stitching.

    The reference's ``Separator.forward`` (separator.py:147-158,231) cuts the audio into independent chunks of
    2 621 440 samples, zero-pads each chunk's edges inside the transform and concatenates the results: the slices
    next to a chunk boundary see zeros instead of their neighbours' samples.  Here a chunk is a contiguous RANGE OF
    SLICES of the one long transform -- the same decomposition as the multi-GPU slice sharding (sharding.py), run
    sequentially in time on one GPU: the analysis of a chunk reads one hop of input before its first owned sample, and
    the synthesis of a chunk hands the first half of its first slice back to the previous chunk's last hop.  For a
    model that acts slice-locally (masks, phasemix) the concatenated output equals the unchunked
    ``insgt(model(nsgt(x)), T)`` bit for bit; memory is bounded by the chunk, not by the track.

        sep = StreamedSeparator(nsgt_base, model, chunk_slices=291)        # ~59 s like the reference's chunk
        for lo, hi, y in sep.stream(x):   # x [B, C, T] on the device; y [targets, B, C, hi - lo] = samples lo .. hi
            ...
        y = sep(x)                        # the concatenation [targets, B, C, T]
    """

    def __init__(self, nsgt_base: NSGTBase, model: Callable[[List[torch.Tensor]], List[torch.Tensor]],
                 chunk_slices: int = 291):
        self.base = nsgt_base
        self.nsgt, self.insgt = make_filterbanks(nsgt_base)
        self.model = model
        self.chunk = max(2, int(chunk_slices))

    def stream(self, x: torch.Tensor) -> Iterator[tuple]:
        nsg = self.base.nsgt
        hop = nsg.sl_len // 2
        lead, T = tuple(x.shape[:-1]), x.shape[-1]
        x2 = x.reshape(-1, T)
        S = nsg.n_slices(T)
        prev = None                      # (lo, hi, y) of the previous chunk, waiting for its right neighbour's halo
        for k0 in range(0, S, self.chunk):
            k1 = min(S, k0 + self.chunk)
            in_lo, own_lo, own_hi = max(0, (k0 - 1) * hop), min(T, k0 * hop), min(T, k1 * hop)
            C = nsg.forward_rows(x2[:, in_lo:own_hi].contiguous(), k0=k0, n_slices=k1 - k0, t0=in_lo,
                                 lead=lead, as_real=True)                      # wrapper layout [*lead, F, S_c, M, 2]
            Y = self.model(C)
            ylead = tuple(Y[0].shape[:-4])
            rows = 1
            for d in ylead:
                rows *= d
            views, keep = [], []
            for Yb, (_, nb, M) in zip(Y, nsg.tables.buckets):
                Yb = Yb.to(torch.float32).contiguous()
                keep.append(Yb)
                views.append((Yb.data_ptr(), nb * (k1 - k0) * M, (k1 - k0) * M, M))
            halo = torch.zeros(rows, hop, dtype=torch.float32, device=x.device) if k0 > 0 else None
            y = nsg.backward_views(views, rows, k1 - k0, x.device, own_hi - own_lo, k0=k0, t0=k0 * hop, halo_out=halo)
            del keep
            if prev is not None:
                plo, phi, py = prev
                a = (k0 - 1) * hop - plo                 # the previous chunk's last hop
                n = max(0, py.shape[1] - a)
                if n:
                    py[:, a:a + n] += halo[:, :n]
                yield plo, phi, py.view(*prev_lead, -1)
            prev, prev_lead = (own_lo, own_hi, y), ylead
        if prev is not None:
            yield prev[0], prev[1], prev[2].view(*prev_lead, -1)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return torch.cat([y for _, _, y in self.stream(x)], dim=-1)
