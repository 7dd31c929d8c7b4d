"""Multi-GPU partitioning of the vocoder path: one process per GPU.

The path shards naturally -- every 64-frame window is an independent fixed-shape inference
(``/root/reference/dual_npu/vocoder_server.py:91-97``) and requests are independent -- so
there is NO data-path collective.  A long request is split into contiguous window ranges;
each rank recomputes at most one neighbouring window for the crossfade at its left edge
(``voc_synthesize_range_dev``) and writes the output span its windows own.  The only
communication is the final gather of int16 PCM to rank 0 (NCCL over NVLink on GPUs, gloo in
the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def window_ranges(n_windows: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [w0, w1) ranges; ranks beyond n_windows get empty ranges."""
    base, rem = divmod(n_windows, world)
    out, w = [], 0
    for r in range(world):
        k = base + (1 if r < rem else 0)
        out.append((w, w + k))
        w += k
    return out


def num_windows(n_tokens: int, max_tokens: int = 64, overlap: int = 16) -> int:
    if n_tokens <= max_tokens:
        return 1
    step = max_tokens - overlap
    return (n_tokens + step - 1) // step


def shard_corpus(lengths: Sequence[int], world: int, max_tokens: int = 64) -> List[List[int]]:
    """Longest-processing-time greedy balance of whole utterances by window count.
    Returns, per rank, the indices of the utterances it synthesises (stable order)."""
    cost = [num_windows(int(n), max_tokens) for n in lengths]
    order = sorted(range(len(lengths)), key=lambda i: (-cost[i], i))
    load = [0] * world
    bins: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        bins[r].append(i)
        load[r] += cost[i]
    for b in bins:
        b.sort()
    return bins


def gather_pcm(local, counts: Sequence[int], rank: int, world: int, group=None):
    """Gather variable-length int16 spans to rank 0 (torch.distributed; padded to the max).

    ``local`` is a 1-D int16 torch tensor on the backend's device (cuda for nccl, cpu for
    gloo) holding this rank's span; ``counts`` are all ranks' span lengths (every rank can
    derive them from the plan).  Returns the concatenated numpy array on rank 0, else None."""
    import torch
    import torch.distributed as dist
    mx = max(max(counts), 1)
    buf = torch.zeros(mx, dtype=torch.int16, device=local.device)
    buf[: counts[rank]] = local[: counts[rank]]
    if world == 1:
        return buf[: counts[0]].cpu().numpy()
    raw = buf.view(torch.uint8)            # gloo has no int16 collectives; bytes work everywhere
    glist = [torch.empty_like(raw) for _ in range(world)] if rank == 0 else None
    dist.gather(raw, glist, dst=0, group=group)
    if rank != 0:
        return None
    return np.concatenate([glist[r].view(torch.int16)[: counts[r]].cpu().numpy() for r in range(world)])


def synthesize_sharded(voc, codes: np.ndarray, rank: int, world: int, group=None):
    """One long request across `world` ranks; rank 0 returns the full PCM16, others None.

    ``voc`` provides ``num_windows(n)`` and ``synthesize_range_pcm16(codes, w0, w1) ->
    (offset, 1-D int16 torch tensor)``; consecutive ranges tile the output exactly, so rank
    r's span starts where rank r-1's ended."""
    import torch
    import torch.distributed as dist
    n = len(codes)
    ranges = window_ranges(voc.num_windows(n), world)
    w0, w1 = ranges[rank]
    off, span = voc.synthesize_range_pcm16(codes, w0, w1)
    cnt = torch.tensor([int(span.numel())], dtype=torch.int64, device=span.device)
    if world > 1:
        all_cnt = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(all_cnt, cnt, group=group)
        counts = [int(c.item()) for c in all_cnt]
    else:
        counts = [int(cnt.item())]
    return gather_pcm(span, counts, rank, world, group)
