"""world_size-2 gloo test of the only multi-GPU step on the path: shard utterances, gather padded token ids."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kotoba_whisper_b200.distributed import TokenGather, gather_token_ids, shard_range
    a, b = shard_range(n_items, rank, world)
    # "transcribe" item i -> i+1 tokens of value 100+i (ragged lengths, like per-rank generate outputs)
    L = max((i + 1 for i in range(a, b)), default=0)
    ids = torch.full((b - a, L), 50257, dtype=torch.int64)
    for r, i in enumerate(range(a, b)):
        ids[r, : i + 1] = 100 + i
    out = gather_token_ids(ids, 50257)
    fixed = gather_token_ids(ids, 50257, max_len=n_items + 3)
    # fixed-shape single-collective form, two batches in flight before the first is read (the bench's pipelining)
    tg = TokenGather(rows_per_rank=4, max_len=n_items + 1, pad_token_id=50257)
    h1 = tg.submit(ids.to(torch.int32))
    h2 = tg.submit((ids + 1).to(torch.int32))
    g1, g2 = h1.result().clone(), h2.result().clone()
    if rank == 0:
        q.put((out, fixed, g1, g2))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    world, n_items = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    out, fixed, g1, g2 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.shape == (n_items, n_items) and fixed.shape == (n_items, n_items + 3)
    for i in range(n_items):
        assert (out[i, : i + 1] == 100 + i).all() and (out[i, i + 1:] == 50257).all()
        assert (fixed[i, : i + 1] == 100 + i).all() and (fixed[i, i + 1:] == 50257).all()
    # TokenGather: rank r's rows sit at [4r, 4r + rows_r), padded to 4 rows x (n_items + 1) columns
    assert g1.shape == (8, n_items + 1) and g1.dtype == torch.int32
    from kotoba_whisper_b200.distributed import shard_range
    for r in range(world):
        a, b = shard_range(n_items, r, world)
        for k, i in enumerate(range(a, b)):
            assert (g1[4 * r + k, : i + 1] == 100 + i).all() and (g1[4 * r + k, i + 1:] == 50257).all()
            assert (g2[4 * r + k, : i + 1] == 101 + i).all()
        assert (g1[4 * r + (b - a): 4 * (r + 1)] == 50257).all()
