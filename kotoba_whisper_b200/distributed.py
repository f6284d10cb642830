"""Multi-GPU plumbing for the transcription path: one process per GPU, a full weight replica each, utterances sharded
across ranks with no data-path collective, and exactly one collective per batch — the gather of the padded token-id
matrix (the reference's `accelerator.pad_across_processes` + `gather_for_metrics`, run_pseudo_labelling.py:339-341).
Works with `nccl` (CUDA tensors, NVLink) and `gloo` (CPU tensors; used by the world_size-2 CPU tests).

Two forms:
  * `gather_token_ids(ids, pad)` — ragged: row counts / lengths are exchanged first (what pad_across_processes does);
    costs two collectives and a host sync per batch, so every rank runs in lock-step with the slowest one.
  * `TokenGather(rows_per_rank, max_len)` — fixed shape [rows_per_rank, max_len] int32: ONE
    `all_gather_into_tensor` per batch, no size exchange, no host sync; `submit()` enqueues it asynchronously and
    returns a handle that is read one batch later, so ranks never wait on each other inside a step."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first n % world ranks get one extra item."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def gather_token_ids(ids: torch.Tensor, pad_token_id: int, group=None, max_len: Optional[int] = None) -> torch.Tensor:
    """ids [B_local, L_local] (int64) on every rank -> [sum B_local, L_max] on every rank, right-padded with pad_token_id.

    With `max_len` given (fixed L_max, e.g. max_length - prompt) the length pre-exchange of pad_across_processes is
    skipped and only row counts + one all_gather of ids cross the fabric."""
    world = _world(group)
    if world == 1:
        return ids
    dev = ids.device
    shape = torch.tensor([ids.shape[0], ids.shape[1]], dtype=torch.int64, device=dev)
    shapes = torch.empty((world * 2,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(shapes, shape, group=group)
    shapes = shapes.view(world, 2).cpu()  # one host sync for all ranks' shapes
    rows = [int(r) for r in shapes[:, 0]]
    L = max_len if max_len is not None else int(shapes[:, 1].max())
    Bmax = max(rows)
    buf = torch.full((Bmax, L), pad_token_id, dtype=torch.int64, device=dev)
    buf[: ids.shape[0], : min(L, ids.shape[1])] = ids[:, :L]
    out = torch.empty((world * Bmax, L), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.view(world, Bmax, L)
    return torch.cat([out[r, : rows[r]] for r in range(world)], dim=0)


class PendingGather:
    """Handle of an in-flight fixed-shape gather."""

    def __init__(self, out: torch.Tensor, work, local: torch.Tensor):
        self._out, self._work, self._local = out, work, local

    def result(self) -> torch.Tensor:
        """[world * rows_per_rank, max_len] int32 on the device; the current stream waits for the collective."""
        if self._work is not None:
            self._work.wait()  # NCCL: stream-side wait, the host does not block
        return self._out


class TokenGather:
    """Fixed-shape gather of token ids: every rank contributes exactly [rows_per_rank, max_len] int32."""

    def __init__(self, rows_per_rank: int, max_len: int, pad_token_id: int, group=None):
        self.rows, self.max_len, self.pad, self.group = int(rows_per_rank), int(max_len), int(pad_token_id), group

    def submit(self, ids: torch.Tensor) -> PendingGather:
        """ids [<= rows_per_rank, <= max_len] integer tensor (device of the process group's backend)."""
        world = _world(self.group)
        local = torch.full((self.rows, self.max_len), self.pad, dtype=torch.int32, device=ids.device)
        local[: ids.shape[0], : ids.shape[1]] = ids[: self.rows, : self.max_len]
        if world == 1:
            return PendingGather(local, None, local)
        out = torch.empty((world * self.rows, self.max_len), dtype=torch.int32, device=ids.device)
        work = dist.all_gather_into_tensor(out, local, group=self.group, async_op=True)
        return PendingGather(out, work, local)
