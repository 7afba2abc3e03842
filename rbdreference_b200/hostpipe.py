"""Host-buffer path of the fused drivers: numpy in, numpy out, copies overlapped with the kernels.

The reference takes and returns host numpy arrays (RBDReference.py:623-628, :785-806, :1345-1368).
For batched numpy arguments the engine therefore cuts the batch into chunks and sends them through
three CUDA streams, each doing  H2D copy -> kernel -> D2H copy  on its own chunk, so that the PCIe
transfers of one chunk run under the kernel and the opposite-direction copy of its neighbours.

* Page-locked arrays (`RBDReference.pinned_empty`, or anything `cudaHostRegister`-ed /
  `torch.Tensor.pin_memory()`-ed) are the source / target of the DMA directly - no host copy.
* Pageable arrays are staged chunk by chunk through persistent pinned buffers owned by the
  pipeline (one `np.copyto` per chunk and direction; that host copy then bounds the throughput).
* Device staging buffers and the streams are created once per engine and device and reused.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

__all__ = ["HostPipeline", "pinned_empty", "is_pinned"]

N_STREAMS = 3
MAX_CHUNKS = 16
MIN_CHUNK_ROWS = 4096


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array backed by page-locked host memory (torch's caching pinned allocator)."""
    tdtype = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}[np.dtype(dtype)]
    return torch.empty(tuple(shape), dtype=tdtype, pin_memory=True).numpy()


def is_pinned(a: np.ndarray) -> bool:
    try:
        return bool(torch.from_numpy(a).is_pinned())
    except Exception:
        return False


def chunk_bounds(B: int) -> np.ndarray:
    nchunk = max(1, min(MAX_CHUNKS, B // MIN_CHUNK_ROWS))
    return np.linspace(0, B, nchunk + 1).astype(np.int64)


class HostPipeline:
    """Chunked H2D -> launch -> D2H over N_STREAMS streams with persistent staging buffers."""

    def __init__(self, device: torch.device, tdtype: torch.dtype):
        self.device = device
        self.tdtype = tdtype
        self.streams = [torch.cuda.Stream(device=device) for _ in range(N_STREAMS)]
        self._dev = {}          # (slot, tag) -> flat device tensor (grow-only)
        self._pin = {}          # (slot, tag) -> flat pinned host tensor (grow-only)
        self.bytes_h2d = 0
        self.bytes_d2h = 0

    def _buf(self, pool, slot: int, tag: str, rows: int, tail: Sequence[int], pinned: bool):
        need = int(rows) * int(np.prod(tail, dtype=np.int64)) if len(tail) else int(rows)
        t = pool.get((slot, tag))
        if t is None or t.numel() < need:
            if pinned:
                t = torch.empty(need, dtype=self.tdtype, pin_memory=True)
            else:
                t = torch.empty(need, dtype=self.tdtype, device=self.device)
            pool[(slot, tag)] = t
        return t[:need].view((int(rows),) + tuple(tail))

    def run(self, B: int, ins: List[Optional[np.ndarray]], in_tails, outs: List[np.ndarray], out_tails,
            launch: Callable[[int, list, list], None]) -> None:
        """`ins[j]` is a C-contiguous (B, *in_tails[j]) array of the engine dtype or None; `outs[j]` a
        C-contiguous (B, *out_tails[j]) array that receives result j.  `launch(m, dins, douts)` enqueues
        the kernel(s) for m knot points on the current stream."""
        if B == 0:
            return
        bounds = chunk_bounds(B)
        nchunk = len(bounds) - 1
        cmax = int(np.max(np.diff(bounds)))
        in_pinned = [a is not None and is_pinned(a) for a in ins]
        out_pinned = [is_pinned(a) for a in outs]
        in_t = [torch.from_numpy(a) if (a is not None and p) else None for a, p in zip(ins, in_pinned)]
        out_t = [torch.from_numpy(a) if p else None for a, p in zip(outs, out_pinned)]
        h2d_done = [None] * N_STREAMS      # event after the last staged H2D of the slot (pinned staging reuse)
        pending = [[] for _ in range(N_STREAMS)]   # (event, j, lo, hi) staged results of the slot not yet copied out

        def drain(slot):
            for ev, j, lo, hi in pending[slot]:
                ev.synchronize()
                np.copyto(outs[j][lo:hi], self._buf(self._pin, slot, "o%d" % j, hi - lo, out_tails[j], True).numpy())
            pending[slot] = []

        with torch.cuda.device(self.device):
            for ci in range(nchunk):
                lo, hi = int(bounds[ci]), int(bounds[ci + 1])
                m = hi - lo
                k = ci % N_STREAMS
                drain(k)
                if h2d_done[k] is not None:
                    h2d_done[k].synchronize()
                with torch.cuda.stream(self.streams[k]):
                    dins = []
                    staged = False
                    for j, a in enumerate(ins):
                        if a is None:
                            dins.append(None)
                            continue
                        d = self._buf(self._dev, k, "i%d" % j, cmax, in_tails[j], False)[:m]
                        if in_pinned[j]:
                            d.copy_(in_t[j][lo:hi], non_blocking=True)
                        else:
                            s = self._buf(self._pin, k, "i%d" % j, m, in_tails[j], True)
                            np.copyto(s.numpy(), a[lo:hi])
                            d.copy_(s, non_blocking=True)
                            staged = True
                        dins.append(d)
                    if staged:
                        h2d_done[k] = torch.cuda.Event()
                        h2d_done[k].record()
                    douts = [self._buf(self._dev, k, "o%d" % j, cmax, out_tails[j], False)[:m] for j in range(len(outs))]
                    launch(m, dins, douts)
                    for j, d in enumerate(douts):
                        if out_pinned[j]:
                            out_t[j][lo:hi].copy_(d, non_blocking=True)
                        else:
                            s = self._buf(self._pin, k, "o%d" % j, m, out_tails[j], True)
                            s.copy_(d, non_blocking=True)
                            ev = torch.cuda.Event()
                            ev.record()
                            pending[k].append((ev, j, lo, hi))
            for k in range(N_STREAMS):
                drain(k)
                self.streams[k].synchronize()
        item = torch.empty(0, dtype=self.tdtype).element_size()
        self.bytes_h2d += sum(int(a.size) * item for a in ins if a is not None)
        self.bytes_d2h += sum(int(a.size) * item for a in outs)
