"""Batched device API: one call for all P parameters (the reference loops in Python).

Inputs are ``[P, C, N]`` float64, draw index fastest -- numpy arrays (copied to the
GPU) or CUDA torch tensors (used in place).  Outputs are torch CUDA tensors unless
stated.  PyTorch supplies device memory and streams only; all arithmetic happens in
the sm_100a kernels behind the C-ABI (``_native``).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native

_DEFAULT_Q = (0.05, 0.5, 0.95)


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise _native.NativeError("no CUDA device visible: the b200 backend has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_device(x) -> torch.Tensor:
    """float64 contiguous CUDA tensor from numpy / nested lists / torch input."""
    dev = _device()
    if isinstance(x, torch.Tensor):
        t = x.to(device=dev, dtype=torch.float64)
    else:
        arr = np.asarray(x, dtype=np.float64)
        if not (arr.flags.c_contiguous and arr.flags.writeable):  # e.g. a read-only Arrow buffer view
            arr = np.array(arr, dtype=np.float64, order="C")
        t = torch.from_numpy(arr).to(dev)
    return t.contiguous()


def to_index(rows) -> torch.Tensor:
    """int64 CUDA index tensor from a host array."""
    return torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int64)).to(_device())


def _stream_ptr() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def _quantile_levels(quantiles) -> list[float]:
    """numpy.quantile (backends_numpy.py:44) raises for levels outside [0, 1]; so does this backend."""
    qs = [float(q) for q in quantiles]
    if len(qs) > _native.MAX_QUANTILES:
        raise ValueError(f"at most {_native.MAX_QUANTILES} quantile levels per call; got {len(qs)}")
    for q in qs:
        if not 0.0 <= q <= 1.0:  # NaN fails both comparisons
            raise ValueError("Quantiles must be in the range [0, 1]")
    return qs


class Workspace:
    """Caller-owned scratch the library carves its buffers from (grows on demand)."""

    def __init__(self) -> None:
        self.buf: torch.Tensor | None = None

    def ensure(self, nbytes: int) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=_device())
        return self.buf

    def for_shape(self, p: int, c: int, n: int, extra: int = 0, cap_bytes: int | None = None) -> torch.Tensor:
        lib = _native.load()
        want = int(lib.b200ref_workspace_bytes(p, c, n)) + extra
        if cap_bytes is not None:
            want = max(min(want, cap_bytes), int(lib.b200ref_workspace_min_bytes(c, n)) + extra)
        return self.ensure(want)


_WS = Workspace()


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def ranks_batch(x, folded: bool = True, workspace: Workspace | None = None) -> dict[str, torch.Tensor]:
    """Pooled tie-averaged ranks and normal scores of x (and of fold(x)).

    Mirrors ``_rank_normalize`` / ``_fold_chains`` (diagnostics.py:93-133) for all P at once.
    """
    lib = _native.load()
    xt = to_device(x)
    p, c, n = xt.shape
    ws = (workspace or _WS).for_shape(p, c, n)
    out = {"avg_rank": torch.empty_like(xt), "z": torch.empty_like(xt)}
    if folded:
        out["avg_rank_folded"] = torch.empty_like(xt)
        out["z_folded"] = torch.empty_like(xt)
    if xt.numel() == 0:
        return out
    _native.check(lib.b200ref_ranks_f64(
        xt.data_ptr(), p, c, n, out["avg_rank"].data_ptr(), out["z"].data_ptr(),
        _ptr(out.get("avg_rank_folded")), _ptr(out.get("z_folded")), ws.data_ptr(), ws.numel(), _stream_ptr()))
    return out


def _alloc_outputs(p: int, nq: int, diag: bool, dev) -> tuple[dict[str, torch.Tensor], _native.Outputs]:
    f64 = dict(dtype=torch.float64, device=dev)
    t: dict[str, torch.Tensor] = {
        "mean": torch.empty(p, **f64), "std": torch.empty(p, **f64),
        "quantiles": torch.empty((p, nq), **f64), "order_stats": torch.empty((p, nq, 2), **f64),
        "median": torch.empty(p, **f64),
    }
    if diag:
        for k in ("rhat", "rhat_bulk", "rhat_tail", "ess_bulk", "ess_tail"):
            t[k] = torch.empty(p, **f64)
        t["lag_bulk"] = torch.empty(p, dtype=torch.int64, device=dev)
        t["lag_tail"] = torch.empty(p, dtype=torch.int64, device=dev)
    o = _native.Outputs()
    for name, _ in _native.Outputs._fields_:
        setattr(o, name, t[name].data_ptr() if name in t else None)
    return t, o


def chain_stats_batch(z_bulk, z_folded=None, workspace: Workspace | None = None,
                      ess_truncation: str = "first_negative") -> dict[str, torch.Tensor]:
    """Split R-hat and ESS of already rank-normalised scores ``z_bulk`` / ``z_folded`` [P, C, N] (the second half of the
    fused pass: ``_split_chains`` + ``_rhat`` + ``_ess``, diagnostics.py:76-85, :136-193)."""
    if ess_truncation not in _native.ESS_MODES:
        raise ValueError(f"ess_truncation must be one of {sorted(_native.ESS_MODES)}; got {ess_truncation!r}")
    lib = _native.load()
    zb = to_device(z_bulk)
    zf = to_device(z_folded) if z_folded is not None else None
    p, c, n = zb.shape
    if zf is not None and zf.shape != zb.shape:
        raise ValueError("z_bulk and z_folded must have the same shape")
    t, o = _alloc_outputs(p, 0, True, zb.device)
    if p == 0:
        return t
    ws = (workspace or _WS).for_shape(p, c, n)
    _native.check(lib.b200ref_chain_stats_f64(zb.data_ptr(), _ptr(zf), p, c, n, _native.ESS_MODES[ess_truncation],
                                              C.byref(o), ws.data_ptr(), ws.numel(), _stream_ptr()))
    return {k: t[k] for k in ("rhat", "rhat_bulk", "rhat_tail", "ess_bulk", "ess_tail", "lag_bulk", "lag_tail")}


def diagnostics_rows(x_rows, n_chains: int, quantiles=_DEFAULT_Q, workspace: Workspace | None = None,
                     ess_truncation: str = "first_negative", ws_cap_bytes: int | None = None) -> dict[str, torch.Tensor]:
    """The same diagnostics for draws in the reference's numpy layout: ``x_rows[C * N, P]``, rows ordered
    (chain, draw), PARAMETER fastest -- what ``Draws.to_numpy()`` / ``reference.draws(return_="numpy")`` return
    (draws.py:20-29).  The library transposes parameter blocks on the device (tiled transposing load) and runs
    the canonical path: bit-identical to ``diagnostics_batch`` of the transposed tensor."""
    if ess_truncation not in _native.ESS_MODES:
        raise ValueError(f"ess_truncation must be one of {sorted(_native.ESS_MODES)}; got {ess_truncation!r}")
    lib = _native.load()
    xt = to_device(x_rows)
    if xt.dim() != 2:
        raise ValueError(f"x_rows must be [C * N, P]; got shape {tuple(xt.shape)}")
    rows, p = xt.shape
    c = int(n_chains)
    if c < 1 or rows % c:
        raise ValueError(f"{rows} rows do not split into {n_chains} equal chains")
    n = rows // c
    qs = _quantile_levels(quantiles)
    t, o = _alloc_outputs(p, len(qs), True, xt.device)
    if p == 0:
        return t
    want = int(lib.b200ref_rows_workspace_bytes(p, c, n))
    if ws_cap_bytes is not None:  # a smaller workspace makes the library walk the parameters in several blocks
        want = max(min(want, int(ws_cap_bytes)), int(lib.b200ref_rows_workspace_bytes(1, c, n)))
    ws = (workspace or _WS).ensure(want)
    ws_view = ws[:want]
    qarr = (C.c_double * max(len(qs), 1))(*qs)
    ws = ws_view
    _native.check(lib.b200ref_diagnostics_layout_f64(xt.data_ptr(), _native.LAYOUTS["rows"], p, c, n, qarr, len(qs),
                                                     _native.ESS_MODES[ess_truncation], C.byref(o),
                                                     ws.data_ptr(), ws.numel(), _stream_ptr()))
    return t


def diagnostics_batch(x, quantiles=_DEFAULT_Q, workspace: Workspace | None = None,
                      ws_cap_bytes: int | None = None, ess_truncation: str = "first_negative") -> dict[str, torch.Tensor]:
    """rhat / ess_bulk / ess_tail (+ mean, std, quantiles, median) for x[P, C, N].

    Batched form of ``split_rhat`` / ``ess_bulk`` / ``ess_tail`` (diagnostics.py:13-73).
    No chain-count guard here (the scalar shims in ``diagnostics`` apply it).
    ``ess_truncation="geyer"`` is an extension (Geyer's initial monotone sequence instead of the
    reference's first-negative rule, diagnostics.py:172-176); the default is the reference's rule.
    """
    if ess_truncation not in _native.ESS_MODES:
        raise ValueError(f"ess_truncation must be one of {sorted(_native.ESS_MODES)}; got {ess_truncation!r}")
    lib = _native.load()
    xt = to_device(x)
    p, c, n = xt.shape
    qs = _quantile_levels(quantiles)
    t, o = _alloc_outputs(p, len(qs), True, xt.device)
    if p == 0:
        return t
    ws = (workspace or _WS).for_shape(p, c, n, cap_bytes=ws_cap_bytes)
    qarr = (C.c_double * max(len(qs), 1))(*qs)
    _native.check(lib.b200ref_diagnostics_ex_f64(xt.data_ptr(), p, c, n, qarr, len(qs),
                                                 _native.ESS_MODES[ess_truncation], C.byref(o),
                                                 ws.data_ptr(), ws.numel(), _stream_ptr()))
    return t


def stats_batch(x, quantiles=_DEFAULT_Q, workspace: Workspace | None = None) -> dict[str, torch.Tensor]:
    """mean / std(ddof=0) / numpy-linear quantiles over the pooled draws of x[P, M].

    Batched form of the per-column body of ``NumpyBackend.stats`` (backends_numpy.py:40-47).
    """
    lib = _native.load()
    xt = to_device(x)
    if xt.dim() == 3:
        xt = xt.reshape(xt.shape[0], -1)
    p, m = xt.shape
    qs = _quantile_levels(quantiles)
    t, o = _alloc_outputs(p, len(qs), False, xt.device)
    if p == 0:
        return t
    ws = (workspace or _WS).for_shape(p, 1, m)
    qarr = (C.c_double * max(len(qs), 1))(*qs)
    _native.check(lib.b200ref_stats_f64(xt.data_ptr(), p, m, qarr, len(qs), C.byref(o),
                                        ws.data_ptr(), ws.numel(), _stream_ptr()))
    return t


def compare_batch(ref, act, tolerance: float) -> tuple[torch.Tensor, torch.Tensor, bool]:
    """rel = |act - ref| / max(|ref|, 1e-12); passed = rel <= tol (compare.py:39-43)."""
    lib = _native.load()
    r = to_device(ref).reshape(-1)
    a = to_device(act).reshape(-1)
    if r.numel() != a.numel():
        raise ValueError("ref and act must have the same number of values")
    rel = torch.empty_like(r)
    ok = torch.empty(r.numel(), dtype=torch.uint8, device=r.device)
    allp = torch.ones(1, dtype=torch.uint8, device=r.device)
    _native.check(lib.b200ref_compare_f64(r.data_ptr(), a.data_ptr(), r.numel(), float(tolerance),
                                          rel.data_ptr(), ok.data_ptr(), allp.data_ptr(), _stream_ptr()))
    return rel, ok.bool(), bool(allp.item())


def checks_batch(rhat: torch.Tensor, ess_bulk: torch.Tensor) -> tuple[bool, bool]:
    """(ess_above_400, rhat_below_1_01) of ``_checks`` (convert.py:165-166)."""
    lib = _native.load()
    v = (C.c_uint8 * 2)()
    r, e = to_device(rhat), to_device(ess_bulk)
    _native.check(lib.b200ref_checks_f64(r.data_ptr(), e.data_ptr(), r.numel(), v, _stream_ptr()))
    return bool(v[0]), bool(v[1])


SUMMARY_COLUMNS = ("rhat", "ess_bulk", "ess_tail", "mean", "std", "q5", "q50", "q95",
                   "ref_mean", "ref_std", "ref_q5", "ref_q50", "ref_q95", "rel_mean", "rel_std", "flags",
                   # extension columns (no reference counterpart): MCSE of the mean, z-score check
                   "mcse_mean", "ref_se_mean", "z_mean", "z_pass")
Z_CRIT = 4.0  # default |z| bound of the (extension) z-score check


def full_pass(cand, ref, tolerance: float = 0.15, workspace: Workspace | None = None,
              ws_cap_bytes: int | None = None, z_crit: float = Z_CRIT) -> tuple[torch.Tensor, bool]:
    """Device-resident fused pass: candidate [P, C, N] vs reference draws [P, Mref].

    Returns (summary [P, len(SUMMARY_COLUMNS)] CUDA tensor, overall compare verdict on mean & std --
    the reference's verdict, ``compare.py:38-51``; the z-score columns are an extension and never
    enter it).
    """
    lib = _native.load()
    ct, rt = to_device(cand), to_device(ref)
    p, c, n = ct.shape
    rt = rt.reshape(p, -1)
    summary = torch.empty((p, _native.SUMMARY_K), dtype=torch.float64, device=ct.device)
    # room for the reference-draw statistics to run underneath the candidate pass (aux stream)
    extra = int(lib.b200ref_full_pass_workspace_bytes(p, c, n, rt.shape[1])) - int(lib.b200ref_workspace_bytes(p, c, n))
    ws = (workspace or _WS).for_shape(p, c, n, extra=extra, cap_bytes=ws_cap_bytes)
    ok = C.c_uint8(0)
    _native.check(lib.b200ref_full_pass_f64(ct.data_ptr(), p, c, n, rt.data_ptr(), rt.shape[1], float(tolerance),
                                            float(z_crit), summary.data_ptr(), C.byref(ok), ws.data_ptr(), ws.numel(),
                                            _stream_ptr()))
    return summary, bool(ok.value)


def full_pass_host(cand_host: torch.Tensor, ref_host: torch.Tensor, tolerance: float = 0.15,
                   summary_host: torch.Tensor | None = None, workspace: Workspace | None = None,
                   ws_bytes: int | None = None, z_crit: float = Z_CRIT) -> tuple[torch.Tensor, bool]:
    """End-to-end pass from HOST buffers (pinned torch CPU tensors for full PCIe speed).

    The library stages parameter slabs through the workspace, overlapping H2D copies
    with kernels, and returns the summary rows in host memory.
    """
    lib = _native.load()
    _device()
    if cand_host.dtype != torch.float64 or ref_host.dtype != torch.float64:
        raise ValueError("host buffers must be float64")
    cand_host, ref_host = cand_host.contiguous(), ref_host.contiguous()
    p, c, n = cand_host.shape
    ref_host = ref_host.reshape(p, -1)
    mref = ref_host.shape[1]
    if summary_host is None:
        summary_host = torch.empty((p, _native.SUMMARY_K), dtype=torch.float64, pin_memory=True)
    if ws_bytes is None:
        slab = min(p, 1024)
        ws_bytes = (int(lib.b200ref_full_pass_workspace_bytes(slab, c, n, mref))
                    + 2 * slab * (c * n + mref + _native.SUMMARY_K) * 8 + 8192)
    ws = (workspace or _WS).ensure(ws_bytes)
    ok = C.c_uint8(0)
    _native.check(lib.b200ref_full_pass_host_f64(cand_host.data_ptr(), p, c, n, ref_host.data_ptr(), mref,
                                                 float(tolerance), float(z_crit), summary_host.data_ptr(), C.byref(ok),
                                                 ws.data_ptr(), ws.numel()))
    return summary_host, bool(ok.value)


def synth_fill(p: int, c: int, n: int, p0: int = 0, seed: int = 4711, kind: int = 0,
               out: torch.Tensor | None = None) -> torch.Tensor:
    """Synthetic draws of SURVEY.md 8(d), generated on the device (bench/test utility)."""
    lib = _native.load()
    if out is None:
        out = torch.empty((p, c, n), dtype=torch.float64, device=_device())
    _native.check(lib.b200ref_synth_fill_f64(out.data_ptr(), p0, p, c, n, seed, kind, _stream_ptr()))
    return out
