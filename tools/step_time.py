"""Time env.step at a given preset/size with CUDA events: python tools/step_time.py PRESET N [override ...]"""
import sys

import torch

sys.path.insert(0, ".")
import vine_robot_isaacgymenvs_b200 as vine  # noqa: E402
from vine_robot_isaacgymenvs_b200 import config as vcfg  # noqa: E402

preset = getattr(vcfg, sys.argv[1])
n = int(sys.argv[2])
env = vine.make(cfg=vcfg.compose(preset + [f"num_envs={n}", "headless=True"] + [a for a in sys.argv[3:] if a != "--graph"]))
g = torch.Generator(device="cuda").manual_seed(0)
acts = [torch.rand(n, 2, device="cuda", generator=g) * 2 - 1 for _ in range(8)]
for t in range(40):
    env.step(acts[t % 8])
torch.cuda.synchronize()
best = 1e9
if "--graph" in sys.argv:   # like bench.py: the steps captured into one CUDA graph, replayed
    extra_note = " [graph replay]"
    st = torch.cuda.Stream()
    g2 = torch.cuda.CUDAGraph()
    import ctypes as C
    with torch.cuda.stream(st):
        with torch.cuda.graph(g2, stream=st):
            for t in range(40):
                env._bind(acts[t % 8])
                env._check(env._lib.vine_step(env._h, C.c_void_p(st.cuda_stream)))
    torch.cuda.synchronize()
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(); g2.replay(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 40)
    env._bind()
for rep in range(0 if "--graph" in sys.argv else 3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(40):
        env.actions.copy_(acts[t % 8])
        env.step_device()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 40)
print(f"{sys.argv[1]} n={n} {' '.join(sys.argv[3:])}: {best:.4f} ms/step  {n / best / 1e3:.4g} env-steps/s", flush=True)
if any("CREATE_SHELF=True" in o or "CREATE_PIPE=True" in o for o in preset):
    import ctypes as C
    c = (C.c_int64 * 4)()
    env._lib.vine_route_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    env._lib.vine_route_counts(env._h, c)
    print(f"    routed: near pass {c[0]} envs ({100.0 * c[0] / n:.2f} %), far pass {c[1]}, given up and redone {c[2]} ({100.0 * c[2] / n:.3f} %)")
