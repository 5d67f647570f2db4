"""Multi-GPU plumbing: parameters are independent, so the path shards by parameter column.

One process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  There is no
data-path collective: each rank runs the full pass on its contiguous block of parameters;
the only communication is ONE all_gather of the [P_local, 20] summary rows into a pre-allocated
buffer.  The global compare verdict is the AND of the gathered flags column, computed
on the device and read only when somebody asks for it as a bool -- no second collective, no host
synchronisation inside the step.
"""

from __future__ import annotations

import os
from pathlib import Path

import torch
import torch.distributed as dist


def _cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device_index: int) -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates pinned host buffers
    (first touch then places them on that node), so that several ranks of one box do not pull their host -> device
    copies across the socket interconnect.  Reads sysfs; a box that exposes one node (or none for the GPU, as VMs do)
    is left alone.  Returns what it found and did -- bench.py puts it next to the end-to-end number."""
    info: dict = {"numa_nodes": None, "gpu_numa_node": None, "cpus_bound": None}
    try:
        pr = torch.cuda.get_device_properties(device_index)
        addr = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        nodes = sorted(Path("/sys/devices/system/node").glob("node[0-9]*"))
        info["numa_nodes"] = len(nodes)
        node = int(Path(f"/sys/bus/pci/devices/{addr}/numa_node").read_text())
        info["gpu_numa_node"] = node
        if node >= 0 and len(nodes) > 1:
            cpus = _cpulist(Path(f"/sys/devices/system/node/node{node}/cpulist").read_text())
            allowed = os.sched_getaffinity(0) & cpus
            if allowed:
                os.sched_setaffinity(0, allowed)
                info["cpus_bound"] = len(allowed)
    except (OSError, ValueError, AttributeError, RuntimeError, AssertionError) as exc:  # no sysfs, no such device: nothing to bind
        info["error"] = type(exc).__name__
    return info


def shard_bounds(n_params: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of parameter columns owned by `rank` (sizes differ by at most 1)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} for world size {world}")
    base, extra = divmod(n_params, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_summaries(local: torch.Tensor, n_params: int | None = None, group=None) -> torch.Tensor:
    """All ranks receive the [P, K] summary of all parameters, in parameter order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    k = local.shape[1]
    if n_params is None:
        counts = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        all_counts = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(all_counts, counts, group=group)
        sizes = [int(c.item()) for c in all_counts]
    else:
        sizes = [hi - lo for lo, hi in (shard_bounds(n_params, world, r) for r in range(world))]
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} parameters, expected {sizes[rank]}")
    if len(set(sizes)) == 1:
        out = _gather_buffer(sum(sizes), k, local.dtype, local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad, k), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


_GATHER_BUFFERS: dict = {}


def _gather_buffer(rows: int, k: int, dtype, device) -> torch.Tensor:
    """The gathered summary of a step lives in one buffer per (shape, device), allocated once."""
    key = (rows, k, dtype, str(device))
    buf = _GATHER_BUFFERS.get(key)
    if buf is None:
        buf = _GATHER_BUFFERS[key] = torch.empty((rows, k), dtype=dtype, device=device)
    return buf


class DeviceVerdict:
    """Overall compare verdict held on the device; ``bool(v)`` is the first (and only) host read."""

    __slots__ = ("flag",)

    def __init__(self, flag: torch.Tensor) -> None:
        self.flag = flag

    def __bool__(self) -> bool:
        return bool(self.flag.item())

    def __repr__(self) -> str:
        return f"DeviceVerdict({bool(self)})"


def verdict_from_summary(summary: torch.Tensor, flags_col: int | None = None) -> DeviceVerdict:
    """AND over all parameters of (mean passed & std passed) -- compare.py:54 -- from the summary's flags
    column (bit0 mean, bit1 std; include/b200ref.h B200REF_S_FLAGS).  Stays on the device."""
    if flags_col is None:
        from . import _native

        flags_col = _native.S_FLAGS
    return DeviceVerdict((summary[:, flags_col] == 3.0).all())


def all_passed(local_ok: bool, device=None, group=None) -> bool:
    """AND of the per-rank compare verdicts."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return bool(local_ok)
    t = torch.tensor([1 if local_ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(t.item())


def gather_draw_blocks(local: torch.Tensor, n_draws: int | None = None, group=None) -> torch.Tensor:
    """All ranks receive the per-draw values (log-densities) of all draws, in draw order: rank r holds the
    contiguous block ``shard_bounds(n_draws, world, r)`` (draws are independent, the
    only communication is this gather of S doubles)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    return gather_summaries(local.reshape(-1, 1), n_draws, group).reshape(-1)


def full_pass_sharded(cand_local, ref_local, n_params: int | None = None, tolerance: float = 0.15, group=None,
                      **kw) -> tuple[torch.Tensor, bool]:
    """Full diagnostics + compare pass on this rank's parameter block, then the one small gather.  With more than
    one rank the verdict is a ``DeviceVerdict`` (truth-testable; reading it is the only host synchronisation)."""
    from . import batch

    summary, ok = batch.full_pass(cand_local, ref_local, tolerance=tolerance, **kw)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return summary, ok
    gathered = gather_summaries(summary, n_params, group)
    return gathered, verdict_from_summary(gathered)
