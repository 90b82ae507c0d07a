"""CPU: the N>1 plumbing (stream sharding + max-over-ranks clock / frame totals) with two
gloo ranks.  There is no data-path collective to test: ranks own whole streams."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vp8b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert shard.rank_info() == (rank, rank, world)
    mine = shard.streams_for_rank(rank, 4, 6)
    frames, secs = shard.combine(dist, torch.device("cpu"), 100 * (rank + 1), 0.5 + rank)
    out.put((rank, mine, frames, secs))
    dist.destroy_process_group()


def test_two_ranks_shard_streams_and_take_max_time():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2, 3] and res[1][1] == [4, 5, 0, 1]      # whole streams, disjoint start
    for _, _, frames, secs in res:
        assert frames == 300.0 and secs == 1.5                           # sum of frames, MAX of clocks


def test_single_rank_passthrough():
    assert shard.streams_for_rank(0, 3, 2) == [0, 1, 0]
    assert shard.combine(None, None, 7, 0.25) == (7, 0.25)
