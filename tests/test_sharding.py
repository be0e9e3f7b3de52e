"""CPU tests of the multi-GPU host logic: the instance batch is split into contiguous slices, one per rank, with no
data-path collective; only the timing reduction uses torch.distributed (gloo here, nccl on the GPU box)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_ranges_cover_batch_exactly():
    from bench import shard_range
    for n in (1, 7, 64, 65536, 16384 + 3):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                a, b = shard_range(n, r, world)
                assert 0 <= a <= b <= n
                seen += list(range(a, b)) if n < 1000 else []
                if n >= 1000:
                    seen.append((a, b))
            if n < 1000:
                assert seen == list(range(n))
            else:
                assert seen[0][0] == 0 and seen[-1][1] == n and all(seen[i][1] == seen[i + 1][0] for i in range(world - 1))
                sizes = [b - a for a, b in seen]
                assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bench import reduce_max_time, shard_range
    a, b = shard_range(1001, rank, world)
    t = reduce_max_time(float(rank + 1), "cpu")            # max over ranks
    tot = torch.tensor([b - a], dtype=torch.int64)
    dist.all_reduce(tot)
    q.put((rank, t, int(tot.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_timing_reduction_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [o[1] for o in out] == [2.0, 2.0]       # both ranks see the max
    assert [o[2] for o in out] == [1001, 1001]     # slices add up to the whole batch
