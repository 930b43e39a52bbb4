"""world_size-2 gloo test (CPU) of the multi-rank host logic: contiguous clip shards + one token all-gather."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_total, ret):
    import torch.distributed as dist
    from unified_audio_b200.parallel import gather_tokens, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    full = torch.arange(n_total * 3 * 4, dtype=torch.int64).reshape(n_total, 3, 4)   # "tokens" of every clip
    got = gather_tokens(full[lo:hi].clone(), n_total)
    ret[rank] = bool(torch.equal(got, full))
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from unified_audio_b200.parallel import shard_range
    for n in (0, 1, 7, 64, 256):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_token_gather_world2_gloo():
    world, n_total = 2, 5   # ragged: 3 + 2 clips
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
