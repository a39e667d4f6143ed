#!/usr/bin/env python
"""Markdown table of BASELINE config 4 from the committed sweep records:

    python tools/sweep_table.py r02g        # profiles/<tag>_sweep_{1,2,4,8}gpu_colav_iw.jsonl + <tag>_sweep_ncu_1gpu_colav_iw.jsonl
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(path):
    return [json.loads(l) for l in open(path) if l.strip().startswith("{")]


def main():
    tag = sys.argv[1]
    P = os.path.join(ROOT, "profiles")
    thr = {g: {(r["envs_total"], r["regime"]): r["env_steps_per_s"] for r in load(f"{P}/{tag}_sweep_{g}gpu_colav_iw.jsonl")}
           for g in (1, 2, 4, 8)}
    ncu = {(r["envs_total"], r["regime"]): r for r in load(f"{P}/{tag}_sweep_ncu_1gpu_colav_iw.jsonl")}
    print("| envs (total) | regime | 1 GPU | 2 GPUs | 4 GPUs | 8 GPUs | ncu, 1 GPU: FP64 TFLOP/s | DRAM GB/s | occupancy % | FP64 pipe % |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for key in thr[1]:
        n = ncu.get(key, {})
        cells = [f"{thr[g].get(key, float('nan')):.3g}" for g in (1, 2, 4, 8)]
        print(f"| {key[0]:.0e} | {key[1].replace('substeps ', '')} | " + " | ".join(cells) +
              f" | {n.get('ncu_fp64_tflops', float('nan')):.2f} | {n.get('ncu_dram_gbs', float('nan')):.0f} | "
              f"{n.get('ncu_achieved_occupancy_pct', float('nan')):.1f} | {n.get('ncu_pipe_fp64_active_pct', float('nan')):.1f} |")


if __name__ == "__main__":
    main()
