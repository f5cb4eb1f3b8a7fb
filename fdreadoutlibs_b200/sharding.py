"""Link sharding across GPUs and the host-side time-ordered TP merge.

Every link is an independent recurrence with private state (the reference already runs one thread per link with no
sharing: src/wibeth/WIBEthFrameProcessor.cpp:231), so links partition across GPUs as contiguous blocks with NO data-path
collective. The only cross-GPU step is host-side: per-GPU TP lists, each sorted by (time_start, link, channel), are
k-way merged — the order TriggerPrimitiveTypeAdapter::operator< imposes downstream
(include/fdreadoutlibs/TriggerPrimitiveTypeAdapter.hpp:26-29, consumed by src/TPCTPRequestHandler.cpp:99-193).
torch.distributed is used only as the transport that brings the lists to rank 0: an all_gather of the lists as byte tensors
(host memory on gloo; device-to-device on NCCL, then one copy to the host on the destination rank).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from . import frames as F
from .api import merge_sorted

LINKS_PER_APA = 40  # 2560 channels / 64 channels per WIBEth link


def shard_links(n_links: int, world_size: int, rank: int, keep_together: int = LINKS_PER_APA) -> Tuple[int, int]:
    """Contiguous block [link0, link0 + n) of rank `rank`. Blocks are aligned to `keep_together` links (one APA stays on one
    GPU) whenever there are at least as many such groups as ranks; otherwise links are split evenly."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    groups = n_links // keep_together if keep_together > 0 else 0
    if keep_together > 0 and groups >= world_size and n_links % keep_together == 0:
        g0 = groups * rank // world_size
        g1 = groups * (rank + 1) // world_size
        return g0 * keep_together, (g1 - g0) * keep_together
    l0 = n_links * rank // world_size
    l1 = n_links * (rank + 1) // world_size
    return l0, l1 - l0


def globalise(tps: np.ndarray, link0: int) -> np.ndarray:
    """Local link indices (0..n-1 inside one handle) -> global link numbers."""
    out = tps.copy()
    out["link"] = out["link"] + np.uint32(link0)
    return out


def _gather_bytes_as_tensors(local: np.ndarray, world: int, rank: int, dst: int, group) -> Optional[List[bytes]]:
    """The ranks' lists as uint8 tensors through all_gather (lengths first, then the records padded to the longest list): on an
    NCCL group the transport is device-to-device (NVLink) plus one copy to the host on `dst`, on gloo it is host memory all the
    way. 30-100x faster than gather_object's pickling for lists of 10^5 records."""
    import torch
    import torch.distributed as dist

    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    n = torch.tensor([local.nbytes], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    longest = max(max(sizes), 1)
    mine = torch.zeros(longest, dtype=torch.uint8, device=dev)
    if local.nbytes:
        mine[: local.nbytes] = torch.from_numpy(local.view(np.uint8).reshape(-1)).to(dev)
    parts = [torch.empty(longest, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    if rank != dst:
        return None
    return [parts[r][: sizes[r]].cpu().numpy().tobytes() for r in range(world)]


def gather_and_merge(local_sorted: np.ndarray, group=None, dst: int = 0, timings: Optional[dict] = None) -> Optional[np.ndarray]:
    """Rank `dst` receives every rank's sorted TP list and returns the merged list; other ranks return None. `timings`, if
    given, receives the seconds spent in the transport (`gather_s`) and in the merge itself (`merge_s`, rank `dst` only)."""
    import time

    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    t0 = time.perf_counter()
    local = np.ascontiguousarray(local_sorted, dtype=F.TP_DTYPE)
    gathered: Optional[List[bytes]] = None
    try:
        gathered = _gather_bytes_as_tensors(local, world, rank, dst, group)
    except (RuntimeError, ValueError, TypeError):  # a backend without tensor all_gather: pickled objects (slow, but always there)
        gathered = [None] * world if rank == dst else None  # type: ignore[list-item]
        dist.gather_object(local.tobytes(), gathered, dst=dst, group=group)
    t1 = time.perf_counter()
    if timings is not None:
        timings["gather_s"] = t1 - t0
    if rank != dst:
        return None
    lists = [np.frombuffer(b, dtype=F.TP_DTYPE) for b in gathered]  # type: ignore[union-attr]
    merged = merge_sorted(lists)
    if timings is not None:
        timings["merge_s"] = time.perf_counter() - t1
    return merged
