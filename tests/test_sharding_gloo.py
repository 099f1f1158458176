"""world_size-2 gloo test (CPU) of the multi-GPU host logic: instance sharding, transcript gather, max-over-ranks."""
import os
import socket

import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from libiop_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_instances(5, rank, world)
    transcript = b"proof-of-rank-%d:" % rank + bytes(range(10 + 7 * rank))      # different lengths per rank
    gathered = sharding.gather_transcripts(transcript)
    mx = sharding.max_over_ranks(10.0 + rank)
    seed = sharding.instance_seed(0x9E3779B97F4A7C15, rank)
    out.put((rank, mine, gathered, mx, seed))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    want = [b"proof-of-rank-0:" + bytes(range(10)), b"proof-of-rank-1:" + bytes(range(17))]
    assert res[0][2] == want and res[1][2] == want
    assert res[0][3] == res[1][3] == 11.0
    assert res[1][4] == res[0][4] + 1
