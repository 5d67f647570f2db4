"""Arrow / Parquet table -> device tensor ``x[P][C][N]``.

The reference turns the long table into per-chain lists once PER PARAMETER with three
``to_pylist()`` calls, a dict of buckets and a sort (``convert.py:150-161``).  Here:

* every parameter column is copied to the device as it lies in the Arrow buffers (a column
  whose rows are ordered by (chain, draw) already is ``x[p][C][N]``);
* the (chain, draw) order is verified ONCE for the table, on the device (``k_row_order``);
* a table in another row order is permuted by one device gather shared by all parameters
  (``k_gather_rows``); the permutation itself is a host ``lexsort`` of the two int columns.
"""

from __future__ import annotations

import ctypes as C
import warnings
from collections.abc import Iterable

import numpy as np
import torch

from . import _native, batch


def _np_column(table, name: str, dtype) -> np.ndarray:
    if hasattr(table, "column"):
        col = table.column(name)
        if hasattr(col, "to_numpy"):
            return np.ascontiguousarray(col.to_numpy(zero_copy_only=False), dtype=dtype)
        return np.ascontiguousarray(np.asarray(col), dtype=dtype)
    return np.ascontiguousarray(np.asarray(table[name]), dtype=dtype)


def _host_tensor(arr: np.ndarray) -> torch.Tensor:
    """Zero-copy torch view of a (possibly read-only) Arrow-backed array; it is only ever read."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        return torch.from_numpy(arr)


_PINNED: list[torch.Tensor] = []
_PINNED_DONE: list = []


def _pinned_slabs(m: int):
    """Two pinned float64 staging slabs of at least m elements and the events of their last H2D copies (kept for
    the life of the process, grown on demand)."""
    if not _PINNED_DONE:
        _PINNED_DONE[:] = [torch.cuda.Event(), torch.cuda.Event()]
    if not _PINNED or _PINNED[0].numel() < m:
        for ev in _PINNED_DONE:
            ev.synchronize()
        _PINNED[:] = [torch.empty(max(m, 1), dtype=torch.float64).pin_memory() for _ in range(2)]
    return _PINNED, _PINNED_DONE


def chain_geometry(chain: np.ndarray) -> tuple[int, int]:
    """(n_chains, draws per chain); ragged chains are rejected (documented divergence)."""
    if chain.size == 0:
        return 0, 0
    _, counts = np.unique(chain, return_counts=True)
    if counts.min() != counts.max():
        raise ValueError("ragged chains are not supported by the b200 backend; pass equal-length chains")
    return int(counts.size), int(counts[0])


def rows_in_order(chain_dev: torch.Tensor, draw_dev: torch.Tensor) -> tuple[bool, bool]:
    """(rows are sorted by (chain, draw), an adjacent (chain, draw) pair repeats) -- device check."""
    lib = _native.load()
    flags = C.c_int(0)
    _native.check(lib.b200ref_row_order_i64(chain_dev.data_ptr(), draw_dev.data_ptr(), chain_dev.numel(),
                                            C.byref(flags), batch._stream_ptr()))
    return (flags.value & 1) == 0, (flags.value & 2) != 0


def table_to_device(table, params: Iterable[str]) -> tuple[torch.Tensor, int, int]:
    """Device tensor [P, C, N] float64 of the listed parameter columns, rows ordered by (chain, draw)."""
    if hasattr(table, "read_all"):
        table = table.read_all()
    params = list(params)
    dev = batch._device()
    chain = _np_column(table, "chain", np.int64)
    draw = _np_column(table, "draw", np.int64)
    c, n = chain_geometry(chain)
    m = c * n
    x = torch.empty((len(params), m), dtype=torch.float64, device=dev)
    # Arrow column buffers are pageable: each column is staged through one of two PINNED slabs so that the H2D
    # copies run at PCIe speed and overlap the host-side copy of the next column
    stage, done = _pinned_slabs(m)
    for i, name in enumerate(params):
        b = i & 1
        done[b].synchronize()  # the slab's previous copy (this call's or an earlier one's) has left the host
        stage[b][:m].copy_(_host_tensor(_np_column(table, name, np.float64)))
        x[i].copy_(stage[b][:m], non_blocking=True)
        done[b].record()
    if m > 1 and params:
        sorted_rows, _ = rows_in_order(_host_tensor(chain).to(dev), _host_tensor(draw).to(dev))
        if not sorted_rows:
            perm = torch.from_numpy(np.lexsort((draw, chain)).astype(np.int64)).to(dev)
            out = torch.empty_like(x)
            lib = _native.load()
            for p0 in range(0, len(params), 65535):
                pb = min(65535, len(params) - p0)
                _native.check(lib.b200ref_gather_rows_f64(x[p0:].data_ptr(), perm.data_ptr(), pb, m,
                                                          out[p0:].data_ptr(), batch._stream_ptr()))
            x = out
    return x.reshape(len(params), c, n), c, n
