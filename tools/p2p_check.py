"""Multi-GPU check of the peer-memory gradient all-reduce (run under torchrun, one rank per GPU of one node):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/p2p_check.py
1. raw exchange: random per-rank vectors through reduce-buffer -> vine_ppo_adam's p2p_sum path vs torch.distributed.all_reduce;
2. training: PPO (MLP and reference network) for a few iterations with grad_allreduce = p2p: parameters bit-identical on all
   ranks afterwards, finite, and the same loss statistics as the NCCL baseline to tolerance;
3. time per iteration, p2p vs nccl (CUDA events, max over ranks).
Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vine_robot_isaacgymenvs_b200 as vine  # noqa: E402
from vine_robot_isaacgymenvs_b200 import config as vcfg, distributed as vd  # noqa: E402
from vine_robot_isaacgymenvs_b200.ppo.ppo import PPOAgent  # noqa: E402


def make_agent(extra, mode, rank, dev, n=4096):
    cfg = vcfg.compose(vcfg.FSTR_OVERRIDES + [f"num_envs={n}", "headless=True", f"sim_device={dev}", f"rl_device={dev}"] + extra)
    env = vine.make(cfg=cfg, global_env_offset=rank * n)
    return PPOAgent(env, cfg["train"], device=dev, seed=42 + rank, use_graphs=True, grad_allreduce=mode)


_ALL_REDUCE = dist.all_reduce


def max_over_ranks(x, dev):
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    _ALL_REDUCE(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world, local = vd.rank_world()
    dev = f"cuda:{local}"
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device(dev))
    out = {"world": world}
    for name, extra in (("mlp", ["train.params.network.rnn=null"]), ("lstm", [])):
        res = {}
        for mode in ("p2p", "nccl", "none"):
            real_all_reduce = dist.all_reduce
            if mode in ("none", "p2p_without_nccl_moments"):   # diagnosis only: "none" = no exchange at all (what unsynchronised GPUs
                torch.distributed.all_reduce = lambda *a, **k: None   # would take); the other = p2p gradients, moments NOT reduced
            agent = make_agent(extra, {"none": "nccl", "p2p_without_nccl_moments": "p2p"}.get(mode, mode), rank, dev)
            for _ in range(6):
                agent.train_epoch()
            torch.cuda.synchronize()
            flat = torch.cat([p.detach().reshape(-1) for p in agent.model.parameters()])
            gathered = [torch.empty_like(flat) for _ in range(world)]
            torch.distributed.all_reduce = real_all_reduce
            dist.all_gather(gathered, flat)
            if mode in ("none", "p2p_without_nccl_moments"):
                torch.distributed.all_reduce = lambda *a, **k: None
            identical = all(torch.equal(gathered[0], g) for g in gathered)
            st = agent.pop_stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier(); torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                agent.train_epoch()
            e1.record(); torch.cuda.synchronize()
            ms = max_over_ranks(e0.elapsed_time(e1) / 20, dev)
            e0.record()
            for _ in range(20):
                agent._g_update.replay()
            e1.record(); torch.cuda.synchronize()
            ms_upd = max_over_ranks(e0.elapsed_time(e1) / 20, dev)
            e0.record()
            for _ in range(20):
                agent._g_rollout.replay()
            e1.record(); torch.cuda.synchronize()
            torch.distributed.all_reduce = real_all_reduce
            ms_roll = max_over_ranks(e0.elapsed_time(e1) / 20, dev)
            res[mode] = {"params_bit_identical_across_ranks": bool(identical), "finite": bool(torch.isfinite(flat).all()),
                         "kl": st["kl"], "a_loss": st["a_loss"], "c_loss": st["c_loss"], "ms_per_iteration": ms, "update_ms": ms_upd, "rollout_ms": ms_roll}
            if mode.startswith("p2p"):
                res[mode]["exchange_timing"] = agent._p2p_mlp.timing()
                seq, timed_out = agent._p2p_mlp.status()
                res[mode].update({"exchanges": seq, "timed_out": timed_out})
                assert (identical or mode != "p2p") and not timed_out, res
            del agent
            torch.cuda.empty_cache()
        out[name] = res
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
