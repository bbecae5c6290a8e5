"""Multi-GPU host logic on CPU: env sharding + the optional {reward, done} all-gather, world_size 2, gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tds_b200.parallel import shard_range, gather_reward_done, RewardDoneExchange


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    reward = torch.arange(lo, hi, dtype=torch.float32)
    done = (torch.arange(lo, hi) % 3 == 0).float()
    r, d = gather_reward_done(reward, done, total, world)
    # the preallocated exchange the bench / rollouts use: reward and done ARE the send buffer's rows
    ns = 64
    ex = RewardDoneExchange(ns, world, "cpu", depth=2)
    sizes = [b - a for a, b in (shard_range(total, k, world) for k in range(world))]
    for step in range(3):
        ex.before_step(step)
        ex.reward(step)[:hi - lo] = reward + step
        ex.done(step)[:hi - lo] = done
        ex.gather(step)
        ex.join()
        r2, d2 = ex.full(sizes, step)
        assert torch.equal(r2, r + step) and torch.equal(d2, d)
    q.put((rank, lo, hi, r.numpy(), d.numpy()))
    dist.destroy_process_group()


def test_shard_range_partitions():
    for total in (1, 7, 4096, 32768, 100):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gather_reward_done_world2():
    total, world = 101, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, lo, hi, r, d in res:
        assert np.array_equal(r, np.arange(total, dtype=np.float32))
        assert np.array_equal(d, (np.arange(total) % 3 == 0).astype(np.float32))
