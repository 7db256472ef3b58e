"""N > 1 host logic on CPU: world_size-2 gloo process group (no GPU needed).

Checks the env partition, that a shard reproduces exactly the trajectories of the same global env
ids in the unsharded batch (the Philox counter is the GLOBAL env id), and the statistics
all-gather that is the path's only collective.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wildfire_control_python_b200.sharding import gather_stats, shard_range


def test_shard_range_partitions_exactly():
    for n in (1, 7, 64, 4096, 65536):
        for world in (1, 2, 3, 8):
            if n < world:
                continue
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, steps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import wf_oracle as wo  # stands in for the GPU slice: same global-id stream contract
        base, count = shard_range(n_total, world, rank)
        cfg = dict(width=10, height=10, seed=77)
        envs = [wo.OracleEnv(cfg, env_id=base + i) for i in range(count)]
        rewards = np.zeros((steps, count))
        local = dict(env_steps=0, episodes=0, deaths=0, contained=0, burnouts=0, ticks=0)
        for e in envs:
            e.reset()
        for s in range(steps):
            for i, e in enumerate(envs):
                _, r, d, _ = e.step(e.random_action())
                rewards[s, i] = r
                local["env_steps"] += 1
                if d:
                    local["episodes"] += 1
                    local["deaths"] += int(not e.planes()["alive"])
                    e.reset()
        np.save(os.path.join(out_dir, f"rewards_{rank}.npy"), rewards)
        g = gather_stats(local)
        assert g["world_size"] == world and len(g["per_rank"]) == world
        assert g["per_rank"][rank] == {k: local.get(k, 0) for k in g["per_rank"][rank]}
        assert g["total"]["env_steps"] == n_total * steps
        if rank == 0:
            np.save(os.path.join(out_dir, "total_episodes.npy"), np.array([g["total"]["episodes"]]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_reproduce_the_unsharded_batch(tmp_path):
    from oracle import wf_oracle as wo
    n_total, steps, world = 11, 60, 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, steps, str(tmp_path)), nprocs=world, join=True)
    sharded = np.concatenate([np.load(tmp_path / f"rewards_{r}.npy") for r in range(world)], axis=1)
    cfg = dict(width=10, height=10, seed=77)
    envs = [wo.OracleEnv(cfg, env_id=i) for i in range(n_total)]
    want = np.zeros((steps, n_total))
    episodes = 0
    for e in envs:
        e.reset()
    for s in range(steps):
        for i, e in enumerate(envs):
            _, r, d, _ = e.step(e.random_action())
            want[s, i] = r
            if d:
                episodes += 1
                e.reset()
    assert np.array_equal(sharded, want)
    assert int(np.load(tmp_path / "total_episodes.npy")[0]) == episodes
