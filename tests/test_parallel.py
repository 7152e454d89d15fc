"""Multi-rank host logic (env sharding, statistics reduction, path gather) on the gloo backend,
world_size 2, CPU only.  The GPU path uses the same functions over NCCL (bench.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cassierl_b200 import parallel as P


def test_shard_range_partitions_exactly():
    for n, w in ((16384, 8), (16385, 8), (10, 3), (5, 8)):
        ids = []
        for r in range(w):
            a, b = P.shard_range(n, r, w)
            assert 0 <= a <= b <= n and (b - a) in (n // w, n // w + 1)
            ids += list(range(a, b))
        assert ids == list(range(n))


def test_phases_independent_of_world_size():
    n = 1000
    full = P.squat_phases(torch.arange(n), n, torch.float64)
    for w in (2, 3, 8):
        parts = [P.squat_phases(P.global_env_ids(n, r, w), n, torch.float64) for r in range(w)]
        assert torch.equal(torch.cat(parts), full)


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = P.global_env_ids(n_total)
        g = torch.Generator().manual_seed(7)
        T = 5
        rew_all = torch.rand(T, n_total, generator=g, dtype=torch.float64)
        # the rollout kernel's flag: 0 running, 1 env terminated, 2 max_path_length reached
        done_all = torch.randint(0, 3, (T, n_total), generator=g, dtype=torch.uint8) * (torch.rand(T, n_total, generator=g) < 0.3).to(torch.uint8)
        obs_all = torch.rand(T, n_total, 17, generator=g)
        st = P.RolloutStats()
        for t in range(T):
            st.update(rew_all[t, ids], done_all[t, ids], obs_all[t, ids])
        red = st.reduce()
        gathered = P.gather_paths(obs_all[:, ids].contiguous(), n_total)
        ok = (abs(red["reward_sum"] - float(rew_all.sum())) < 1e-9 and red["steps"] == T * n_total
              and red["episodes"] == float((done_all != 0).sum()) and red["truncated"] == float((done_all == 2).sum())
              and red["terminated"] + red["truncated"] == red["episodes"] and torch.equal(gathered, obs_all))
        q.put((rank, bool(ok), red["mean_reward"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [64, 65])
def test_stats_reduce_and_path_gather_gloo_world2(n_total):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res)
    assert res[0][2] == res[1][2]
