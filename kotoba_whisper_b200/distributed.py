"""Multi-GPU plumbing for the transcription path: one process per GPU, a full weight replica each, utterances sharded
across ranks with no data-path collective, and exactly one collective per batch — the gather of the padded token-id
matrix (the reference's `accelerator.pad_across_processes` + `gather_for_metrics`, run_pseudo_labelling.py:339-341).
Works with `nccl` (CUDA tensors, NVLink) and `gloo` (CPU tensors; used by the world_size-2 CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first n % world ranks get one extra item."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_token_ids(ids: torch.Tensor, pad_token_id: int, group=None, max_len: Optional[int] = None) -> torch.Tensor:
    """ids [B_local, L_local] (int64) on every rank -> [sum B_local, L_max] on every rank, right-padded with pad_token_id.

    With `max_len` given (fixed L_max, e.g. max_length - prompt) the length pre-exchange of pad_across_processes is
    skipped and only row counts + one all_gather of ids cross the fabric."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return ids
    world = dist.get_world_size(group)
    dev = ids.device
    shape = torch.tensor([ids.shape[0], ids.shape[1]], dtype=torch.int64, device=dev)
    shapes = [torch.zeros_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    rows = [int(s[0]) for s in shapes]
    L = max_len if max_len is not None else max(int(s[1]) for s in shapes)
    Bmax = max(rows)
    buf = torch.full((Bmax, L), pad_token_id, dtype=torch.int64, device=dev)
    buf[: ids.shape[0], : ids.shape[1]] = ids[:, :L]
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:r] for o, r in zip(out, rows)], dim=0)
