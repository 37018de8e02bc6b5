#!/usr/bin/env python
"""Markdown table of the bench.py JSON lines that scripts/scale_configs.sh leaves under gpurun_out/ (<tag>_scale_<config>_n<N>.json).

usage: scale_table.py [tag] [dir]"""
import glob, json, os, re, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
root = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
rows = {}
for f in glob.glob(os.path.join(root, f"{tag}_scale_*_n*.json")):
    m = re.match(rf"{tag}_scale_(.+)_n(\d+)\.json", os.path.basename(f))
    lines = [json.loads(l) for l in open(f) if l.startswith("{")]
    if m and lines:
        rows[(m.group(1), int(m.group(2)))] = lines[-1]
names = {"c2": "C2 CartPole-v1:ppo, 65,536 envs PER GPU, 64x64 (weak)", "c3": "C3 CartPole-v1:reinforce, 262,144 envs TOTAL, 64x64 (strong)",
         "c4_acrobot": "C4 Acrobot-v1:ppo, 1,048,576 envs TOTAL, 128x128 (strong)", "c4_mcar": "C4 MountainCar-v0:ppo + StateCountBonus, 1,048,576 envs TOTAL, 256x256 (strong)",
         "c5_1m": "C5 CartPole-v1:ppo, 1,048,576 envs TOTAL, 64x64, minibatch 1 M per GPU", "c5_4m": "C5 CartPole-v1:ppo, 4,194,304 envs TOTAL, 64x64, minibatch 1 M per GPU"}
print("| config | GPUs | env-steps/s (device-timed) | e2e env-steps/s | ms per iteration | vs 1 GPU | efficiency | SM MHz (median), throttle reasons |")
print("|---|---|---|---|---|---|---|---|")
for cfg in ["c2", "c3", "c4_acrobot", "c4_mcar", "c5_1m", "c5_4m"]:
    base = rows.get((cfg, 1))
    for n in (1, 2, 4, 8):
        d = rows.get((cfg, n))
        if not d:
            continue
        sp = d["value"] / base["value"] if base else float("nan")
        print(f"| {names[cfg]} | {n} | {d['value'] / 1e6:,.1f} M | {d['e2e']['value'] / 1e6:,.1f} M | {d['ms_per_step']:.2f} | {sp:.2f}x | {sp / n:.3f} | "
              f"{d['clocks']['sm_mhz']:.0f} {', '.join(d['clocks']['reasons']) or '-'} |")
