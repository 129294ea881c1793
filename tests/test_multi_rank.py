"""CPU, world_size 2 over gloo: the N>1 path = scenario sharding by index + the episode-statistics all-reduce
(there is no data-path collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metadrive_ped_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = shard.shard_scenarios(10, 6, rank)
    term = torch.tensor([1, 0, 0, 1, 0, 0], dtype=torch.uint8)
    trunc = torch.tensor([0, 0, 1, 0, 0, 0], dtype=torch.uint8)
    flags = torch.tensor([0x800, 0, 0x1000, 0x401 if rank else 0x001, 0, 0], dtype=torch.int32)
    info_f = torch.zeros((6, 8))
    info_f[:, 6] = float(rank + 1)
    info_f[:, 7] = 10.0
    st = shard.all_reduce_stats(shard.episode_stats(term, trunc, flags, info_f))
    out[rank] = (idx, st.tolist())
    dist.destroy_process_group()


def test_sharding_and_stats_allreduce():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][0] == [0, 1, 2, 3, 4, 5] and res[1][0] == [6, 7, 8, 9, 0, 1]
    # disjoint-then-wrapping coverage of the library
    assert sorted(set(res[0][0] + res[1][0])) == list(range(10))
    s0, s1 = res[0][1], res[1][1]
    assert s0 == s1  # every rank holds the global sum
    assert s0[0] == 6.0  # 3 finished episodes per rank
    assert s0[1] == 3 * 1.0 + 3 * 2.0 and s0[2] == 60.0
    assert s0[4] == 2.0 and s0[5] == 2.0 and s0[6] == 1.0 and s0[7] == 2.0
