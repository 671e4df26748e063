"""Flat fp32 arenas over the lower problem's parameter list.

The reference keeps every K-loop vector as a Python list of per-parameter tensors and rebuilds flat
copies with ``to_vec`` three times per CG iteration (reference betty/utils.py:117-118, cg.py:42-44).
Here every vector (v, p, x, r, H.d) is ONE contiguous fp32 buffer; tensor ``i`` lives at
``offsets[i]`` (rounded up to 4 floats so every kernel can use 128-bit accesses; padding stays
zero) and the per-parameter tensors handed back to the caller are zero-copy views.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _native as N

MT_CHUNK = 16384  # == BB_MT_CHUNK


class ArenaLayout:
    def __init__(self, shapes: Sequence[torch.Size]):
        self.shapes = [tuple(s) for s in shapes]
        self.numels = [int(np.prod(s)) if len(s) else 1 for s in self.shapes]
        self.offsets: List[int] = []
        off = 0
        for n in self.numels:
            self.offsets.append(off)
            off += (n + 3) & ~3
        self.total = off  # multiple of 4
        self.n_logical = sum(self.numels)

    @classmethod
    def like(cls, tensors: Sequence[torch.Tensor]) -> "ArenaLayout":
        return cls([t.shape for t in tensors])

    def new(self, device) -> torch.Tensor:
        return torch.zeros(self.total, dtype=torch.float32, device=device)

    def views(self, flat: torch.Tensor) -> List[torch.Tensor]:
        return [flat[o:o + n].view(s) for o, n, s in zip(self.offsets, self.numels, self.shapes)]

    def chunk_table(self, tensors: Sequence[torch.Tensor], flat: torch.Tensor) -> "ChunkTable":
        """Table pairing each (contiguous fp32) tensor with its arena slice."""
        a_ptrs, b_ptrs = [], []
        for t, off in zip(tensors, self.offsets):
            a_ptrs.append(t.data_ptr())
            b_ptrs.append(flat.data_ptr() + 4 * off)
        return ChunkTable(a_ptrs, b_ptrs, self.numels, flat.device, keep=(list(tensors), flat))


class ChunkTable:
    """Device-resident ``bb_mt_chunk[]`` for the multi-tensor kernels."""

    def __init__(self, a_ptrs, b_ptrs, numels, device, keep=None):
        rows = []
        for a, b, n in zip(a_ptrs, b_ptrs, numels):
            done = 0
            while done < n:
                m = min(MT_CHUNK, n - done)
                rows.append((a + 4 * done, b + 4 * done, m, 0))
                done += m
        arr = np.array(rows, dtype=np.dtype([("a", np.uint64), ("b", np.uint64), ("n", np.int32), ("pad", np.int32)]))
        assert arr.dtype.itemsize == C.sizeof(N.MtChunk)
        self.n = len(rows)
        self.dev = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(device)
        self._keep = keep

    @property
    def ptr(self):
        return self.dev.data_ptr()


def as_f32_contig(ts: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    out = []
    for t in ts:
        t = t.detach()
        if t.dtype != torch.float32:
            t = t.float()
        out.append(t if t.is_contiguous() else t.contiguous())
    return out


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def pack(layout: ArenaLayout, tensors: Sequence[torch.Tensor], flat: torch.Tensor):
    """flat <- tensors (one multi-tensor launch)."""
    ts = as_f32_contig(tensors)
    tab = layout.chunk_table(ts, flat)
    N.call("bb_mt_copy", tab.ptr, tab.n, 0, stream_ptr())
    return tab
