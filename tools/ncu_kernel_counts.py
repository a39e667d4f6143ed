#!/usr/bin/env python
"""Measured per-env-step counts of the env kernel -> profiles/kernel_counts.json (read by bench.py).

    ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,\\
smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \\
        --clock-control none --import-source on -k regex:k_env -s 11 -c 2 -o prof  python bench.py --steps 1 ...
    ncu -i prof.ncu-rep --page raw --csv > prof.raw.csv
    python tools/ncu_kernel_counts.py --key colav_iw --raw prof.raw.csv --bench bench.json --skip 11 --envs 100000

Per captured launch the executed FP64 flop (DFMA = 2, DADD = DMUL = 1; thread-level, predicated-on) and the DRAM
bytes are divided by the simulator steps that launch executed (bench.py prints them per launch index as
`launch_env_steps`; launch i of the capture is step() call (skip + i) mod 9 of an episode).  The file records the
hash of the device sources the capture belongs to: bench.py refuses to quote a roofline fraction from counts taken
on other code."""
import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def col(hdr, row, name, default=None):
    return float(row[hdr.index(name)].replace(",", "")) if name in hdr and row[hdr.index(name)] else default


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--key", required=True, help="colav_iw | rl | colav_iw+sbmpc | rl+sbmpc")
    ap.add_argument("--raw", required=True)
    ap.add_argument("--bench", required=True, help="bench.py JSON line of the same command (launch_env_steps)")
    ap.add_argument("--skip", type=int, required=True, help="the -s value of the ncu capture (k_env launches skipped)")
    ap.add_argument("--envs", type=int, required=True)
    ap.add_argument("--source", default=None, help="name of the committed ncu summary the numbers come from")
    ap.add_argument("--split", action="store_true",
                    help="every step() call is two k_env launches (SenvView::call_filter): captured launches 2j, 2j + 1 "
                         "belong to call (skip / 2 + j) mod 9 and are added up")
    a = ap.parse_args()
    csv.field_size_limit(10 ** 9)
    rows = list(csv.reader(open(a.raw)))
    hdr, data = rows[0], rows[2:]
    line = json.loads(open(a.bench).read().strip().split("\n")[-1])
    if a.key in ("colav_iw", "rl") or "workloads" not in line:
        steps_per_launch = line["launch_env_steps"][1:]
    else:
        raise SystemExit("run bench.py with --workload/--collav of the key so that launch_env_steps belongs to it")
    out = []
    for i, r in enumerate(data):
        n = steps_per_launch[((a.skip + i) // 2) % 9] if a.split else steps_per_launch[(a.skip + i) % 9]
        cyc = col(hdr, r, "sm__cycles_elapsed.max")

        def total(op):
            v = col(hdr, r, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum")
            if v is None:
                v = col(hdr, r, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") * cyc
            return v
        dfma, dadd, dmul = total("dfma"), total("dadd"), total("dmul")
        # ncu prints every byte count in the unit of ITS OWN column of the units row (the write count of a launch that
        # writes little comes in Kbyte next to a read count in Mbyte)
        def dram_bytes(name):
            if name not in hdr:
                return 0.0
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(rows[1][hdr.index(name)], 1.0)
            return col(hdr, r, name, 0.0) * scale
        dram = dram_bytes("dram__bytes_read.sum") + dram_bytes("dram__bytes_write.sum")
        out.append({"env_steps": n, "fp64_inst_per_env_step": (dfma + dadd + dmul) / n,
                    "flop_exec_per_env_step": (2 * dfma + dadd + dmul) / n, "dram_bytes": dram,
                    "pipe_fp64_pct": col(hdr, r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                    "issue_active_pct": col(hdr, r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "achieved_occupancy_pct": col(hdr, r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                    "threads_per_inst": col(hdr, r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                    "duration_us": col(hdr, r, "gpu__time_duration.sum")})
    if a.split:
        # add the two launches of a call up; percentages weighted by duration
        calls = []
        for j in range(0, len(out) - 1, 2):
            p, q = out[j], out[j + 1]
            n, t = p["env_steps"], p["duration_us"] + q["duration_us"]
            w = lambda k: (p[k] * p["duration_us"] + q[k] * q["duration_us"]) / t
            calls.append({"env_steps": n,
                          "fp64_inst_per_env_step": p["fp64_inst_per_env_step"] + q["fp64_inst_per_env_step"],
                          "flop_exec_per_env_step": p["flop_exec_per_env_step"] + q["flop_exec_per_env_step"],
                          "dram_bytes": p["dram_bytes"] + q["dram_bytes"], "pipe_fp64_pct": w("pipe_fp64_pct"),
                          "issue_active_pct": w("issue_active_pct"), "achieved_occupancy_pct": w("achieved_occupancy_pct"),
                          "threads_per_inst": w("threads_per_inst"), "duration_us": t,
                          "duration_us_launches": [p["duration_us"], q["duration_us"]]})
        out = calls
        # per env-step figures of the whole capture: weighted by the simulator steps of each call
        tot = sum(o["env_steps"] for o in out)
        mean = lambda k: (sum(o[k] * o["env_steps"] for o in out) / tot if k.endswith("per_env_step")
                          else (sum(o[k] * o["duration_us"] for o in out) / sum(o["duration_us"] for o in out)
                                if k.endswith("_pct") or k == "threads_per_inst" else sum(o[k] for o in out) / len(out)))
    else:
        mean = lambda k: sum(o[k] for o in out) / len(out)
    try:
        d = json.load(open(bench.KERNEL_COUNTS))
    except Exception:
        d = {"kernels": {}}
    h = bench.device_source_hash()
    if d.get("source_hash") != h:
        d = {"kernels": {}}                       # counts of other code are dropped, not mixed
    d["source_hash"] = h
    d["hashed_files"] = list(bench.DEVICE_SOURCES)
    d["kernels"][a.key] = {
        "flop_exec_per_env_step": round(mean("flop_exec_per_env_step"), 1),
        "fp64_inst_per_env_step": round(mean("fp64_inst_per_env_step"), 1),
        "dram_bytes_per_launch": round(mean("dram_bytes")), "envs": a.envs,
        "pipe_fp64_pct": mean("pipe_fp64_pct"), "issue_active_pct": mean("issue_active_pct"),
        "achieved_occupancy_pct": mean("achieved_occupancy_pct"), "threads_per_inst": mean("threads_per_inst"),
        "launches": out, "source": a.source or os.path.basename(a.raw)}
    if a.split and len(out) >= 9:
        # whole-episode capture: also the count of the calls in which both ships sail (what rounds 1 / 2 quoted)
        d["kernels"][a.key]["scope"] = "every launch of one episode (9 step() calls), weighted by simulator steps"
        d["kernels"][a.key]["flop_exec_per_env_step_midcall"] = round(sum(o["flop_exec_per_env_step"] for o in out[2:4]) / 2, 1)
    json.dump(d, open(bench.KERNEL_COUNTS, "w"), indent=1)
    print(json.dumps(d["kernels"][a.key], indent=1)[:800])


if __name__ == "__main__":
    main()
