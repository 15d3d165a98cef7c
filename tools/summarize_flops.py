"""Summarise tools/ncu_flops.sh captures: per kernel launch of the profiled control steps its device time and FP32
thread-instruction counts, and per config the algorithmic FLOPs per env-step -> profiles/flops_per_env_step.json.
    python tools/summarize_flops.py gpurun_out r02b"""
import csv
import glob
import json
import os
import re
import sys

src, tag = sys.argv[1], sys.argv[2]
KEYS = {"FSTR_OVERRIDES": "fstr", "SHELF_OVERRIDES": "shelf", "PIPE_DR_OVERRIDES": "pipe_dr"}
out = {}
for path in sorted(glob.glob(os.path.join(src, f"flops_*_{tag}.csv"))):
    m = re.search(r"flops_([A-Z_]+)_(\d+)_" + tag, path)
    preset, n = m.group(1), int(m.group(2))
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    launches = {}
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        lid = int(r[idx["ID"]])
        name = r[idx["Kernel Name"]].split("(")[0]
        d = launches.setdefault(lid, {"kernel": name})
        val = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        mname = r[idx["Metric Name"]]
        if mname == "gpu__time_duration.sum":
            val = val / 1e3 if unit in ("ns", "nsecond") else val * (1e3 if unit in ("ms", "msecond") else 1.0)   # -> us
        if mname.startswith("dram__bytes"):
            val *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
        d[mname] = val
    steps = 2
    L = [launches[k] for k in sorted(launches)]
    ffma = sum(x.get("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", 0) for x in L)
    fmul = sum(x.get("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", 0) for x in L)
    fadd = sum(x.get("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", 0) for x in L)
    inst = sum(x.get("smsp__thread_inst_executed.sum", 0) for x in L)
    dram = sum(x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0) for x in L)
    key = f"{KEYS[preset]}_{n}"
    out[key] = {"preset": preset, "num_envs": n, "control_steps_profiled": steps,
                "ffma_per_env_step": ffma / steps / n, "fmul_per_env_step": fmul / steps / n, "fadd_per_env_step": fadd / steps / n,
                "thread_inst_per_env_step": inst / steps / n, "flops_per_env_step": (2 * ffma + fmul + fadd) / steps / n,
                "dram_bytes_per_env_step": dram / steps / n,
                "launches_per_step": [{"kernel": x["kernel"], "us_cold_serialised": round(x.get("gpu__time_duration.sum", 0), 1)}
                                      for x in L[:len(L) // steps]],
                "source": os.path.basename(path)}
json.dump(out, open(os.path.join("profiles", "flops_per_env_step.json"), "w"), indent=1)
for k, v in out.items():
    print(k, f"{v['flops_per_env_step']:.0f} FLOP/env-step, {v['thread_inst_per_env_step']:.0f} inst, {v['dram_bytes_per_env_step']:.0f} B dram;",
          " | ".join(f"{x['kernel']} {x['us_cold_serialised']}us" for x in v["launches_per_step"]))
