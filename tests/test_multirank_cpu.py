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
    buffers = {}
    got = gather_tokens(full[lo:hi].clone(), n_total, buffers=buffers)
    ok = bool(torch.equal(got, full))
    if n_total % world == 0:        # equal shards: the preallocated buffer is reused call after call
        again = gather_tokens(full[lo:hi].clone() + 1, n_total, buffers=buffers)
        ok = ok and again.data_ptr() == got.data_ptr() and bool(torch.equal(again, full + 1)) and len(buffers) == 1
    ret[rank] = ok
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
    for n_total in (5, 8):   # ragged (3 + 2 clips: padded gather) and equal shards (one all_gather_into_tensor into a static buffer)
        world = 2
        ret = mp.Manager().dict()
        mp.spawn(_worker, args=(world, _free_port(), n_total, ret), nprocs=world, join=True)
        assert all(ret[r] for r in range(world)), n_total
