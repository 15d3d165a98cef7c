"""CPU: the N>1 host logic under world_size-2 gloo (no GPU, no cluster)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vine_robot_isaacgymenvs_b200 import distributed as vd


def test_env_shards_partition_the_global_id_range():
    for total in (1, 7, 64, 4096, 65536, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            spans = [vd.env_shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _worker(rank, world, port, q):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank)})
    r, w, _ = vd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    out = {}
    out["max"] = vd.max_over_ranks(1.0 + rank)
    # gradient + KL in one collective
    g = torch.full((10,), float(rank + 1))
    g, extra = vd.allreduce_mean_(g, torch.tensor([0.01 * (rank + 1)]))
    out["grad"], out["kl"] = g.numpy().copy(), float(extra[0])
    # running statistics of a sharded data set == statistics of the whole
    rng = np.random.default_rng(0)
    data = rng.normal(1.0, 2.0, (1000, 3)).astype(np.float32)
    start, count = vd.env_shard(len(data), rank, world)
    x = torch.from_numpy(data[start:start + count])
    mean = x.mean(0)
    m2 = ((x - mean) ** 2).sum(0)
    c, gm, gm2 = vd.merge_moments(count, mean, m2)
    out["count"], out["mean"], out["var"] = float(c), gm.numpy(), (gm2 / c).numpy()
    # 0-dim statistics (the value normaliser) keep their shape through the merge, and the RunningMeanStd of the PPO agent
    # ends up identical on every rank and equal to the single-process statistics of the whole data set
    from vine_robot_isaacgymenvs_b200.ppo.ppo import RunningMeanStd
    v = torch.from_numpy(data[start:start + count, 0].copy())
    c0, m0, s0 = vd.merge_moments(count, v.mean(), ((v - v.mean()) ** 2).sum())
    out["scalar_shape"] = (tuple(m0.shape), tuple(s0.shape))
    rms = RunningMeanStd(())
    rms.update(v)
    out["rms"] = (float(rms.running_mean), float(rms.running_var), float(rms.count))
    # the flat gradient vector (+ loss statistics) of the kernel-only PPO paths: one sum all-reduce, scaled by 1/world
    flat = torch.arange(8, dtype=torch.float32) * (rank + 1)
    dist.all_reduce(flat)
    out["flat"] = (flat / world).numpy()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_collectives_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    data = rng.normal(1.0, 2.0, (1000, 3)).astype(np.float32)
    for r in range(2):
        o = res[r]
        assert o["max"] == 2.0
        assert np.allclose(o["grad"], 1.5) and abs(o["kl"] - 0.015) < 1e-7
        assert o["count"] == 1000
        assert np.allclose(o["mean"], data.mean(0), atol=1e-5)
        assert np.allclose(o["var"], data.var(0), rtol=1e-4)
        assert o["scalar_shape"] == ((), ())
        ref = __import__("vine_robot_isaacgymenvs_b200.ppo.ppo", fromlist=["RunningMeanStd"]).RunningMeanStd(())
        ref.update(torch.from_numpy(data[:, 0].copy()))
        assert np.allclose(o["rms"], (float(ref.running_mean), float(ref.running_var), float(ref.count)), rtol=1e-5)
        assert np.allclose(o["flat"], np.arange(8) * 1.5)
