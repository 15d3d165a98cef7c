"""BASELINE.md §4-2a / SURVEY §8(d) CPU baseline (i): the reference's UNMODIFIED Python step logic
(VecTask.step + Vine5LinkMovingBase hooks, imported from /root/reference by tests/golden/ref_harness.py) on torch CPU,
with the oracle's f32 dynamics standing in for the closed PhysX `gym.simulate`.  Config C1: free space, num_envs=64,
200 random-action control steps, `torch.set_num_threads(nproc)`.  A substitute for the PhysX CPU pipeline (which cannot
be installed: closed Isaac Gym binary), labelled as such.

/root/reference exists only in the build container, so this cannot run on the GPU box: run it here,
    python tools/cpu_baseline_c1.py            # writes profiles/cpu_baseline_c1.json
and bench.py embeds the committed record (with where it was measured); if /root/reference IS present where bench.py
runs, bench.py times it live instead.
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))


def measure(num_envs=64, steps=200, warmup=10, threads=None):
    import numpy as np
    import torch
    import ref_harness as H
    from vine_robot_isaacgymenvs_b200 import config as vcfg
    threads = threads or len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)
    cfg = vcfg.task_config([f"num_envs={num_envs}", "task.env.CREATE_PIPE=False", "task.env.CREATE_SHELF=False"])
    rt = H.ReferenceTask(cfg, seed=42, use_f64=False)
    rng = np.random.default_rng(42)
    acts = rng.uniform(-1, 1, (steps + warmup, num_envs, 2)).astype(np.float32)
    for t in range(warmup):
        rt.step(acts[t])
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        rt.step(acts[t])
    dt = time.perf_counter() - t0
    return {"value": num_envs * steps / dt, "unit": "env-steps/s", "cores": threads, "kind": "reference-step-logic+port-dynamics",
            "sample": f"BASELINE configs[0] (C1): {num_envs} envs x {steps} control steps, reference's unmodified "
                      f"VecTask.step / Vine5LinkMovingBase Python (torch CPU, {threads} threads) with the oracle's f32 "
                      f"dynamics as gym.simulate; substitute for the PhysX CPU pipeline",
            "seconds": dt, "ms_per_step": 1e3 * dt / steps}


if __name__ == "__main__":
    rec = measure()
    import platform
    rec["measured_on"] = f"build container ({platform.processor() or platform.machine()}, {rec['cores']} cores), no GPU box access to /root/reference"
    out = os.path.join(REPO, "profiles", "cpu_baseline_c1.json")
    with open(out, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec))
