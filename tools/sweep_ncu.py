#!/usr/bin/env python
"""BASELINE config 4, the ncu columns: for every (environments, regime) point of the 1-GPU sweep, one ncu pass over
the env kernel's launches of that point -> achieved FP64 FLOP/s, DRAM GB/s and achieved occupancy, measured (not
derived from constants).

    python tools/sweep_ncu.py [--workload colav_iw] [--envs 1000,...] > profiles/r02_sweep_ncu_1gpu_colav_iw.jsonl

Per point it runs `ncu --metrics <6 counters> -k regex:k_env` around `tools/scaling_sweep.py --no-warmup --repeats 1`
restricted to that point, sums the counters over the captured launches (a whole episode: the 9 step() launches after
the warm-up episode; k x _step(): 256 / k launches) and divides by the summed kernel time.  The kernel times under ncu
are serialised, cold-cache replays, so the env-steps/s of the sweep proper (tools/scaling_sweep.py) are the
throughput numbers; these rows give the three counters the metric definition asks for."""
import argparse
import csv
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
           "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
        "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}


def capture(workload, envs, regime, k):
    with tempfile.TemporaryDirectory() as td:
        log = os.path.join(td, "ncu.csv")
        cmd = ["ncu", "--metrics", ",".join(METRICS), "--clock-control", "none", "-k", "regex:k_env", "--csv", "--log-file", log]
        if regime == "episode":
            # skip the warm-up episode's launches, take the measured episode (the colav_iw env issues every step() call as
            # two k_env launches)
            n = 18 if (workload == "colav_iw" and os.environ.get("SHIPENV_SPLIT_CALLS", "1") != "0") else 9
            cmd += ["-s", str(n), "-c", str(n)]
        sweep = [sys.executable, os.path.join(ROOT, "tools", "scaling_sweep.py"), "--workload", workload, "--envs", str(envs),
                 "--repeats", "1", "--regimes", "episode" if regime == "episode" else "substeps", "--ks", str(k or 1),
                 "--substeps-total", str(max(k or 1, 32))]
        if regime != "episode":
            sweep.append("--no-warmup")
        out = subprocess.run(cmd + sweep, capture_output=True, text=True, cwd=ROOT)
        rows = [r for r in csv.reader(open(log)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    i_name, i_unit, i_val, i_id = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
    per = {}
    for r in rows[1:]:
        v = float(r[i_val].replace(",", "")) * UNIT.get(r[i_unit], 1.0)
        per.setdefault(r[i_id], {})[r[i_name]] = v
    launches = list(per.values())
    t = sum(l["gpu__time_duration.sum"] for l in launches)
    flop = sum(2 * l[METRICS[1]] + l[METRICS[2]] + l[METRICS[3]] for l in launches)
    dram = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in launches)
    occ = sum(l[METRICS[6]] * l["gpu__time_duration.sum"] for l in launches) / t
    pipe = sum(l[METRICS[7]] * l["gpu__time_duration.sum"] for l in launches) / t
    issue = sum(l[METRICS[8]] * l["gpu__time_duration.sum"] for l in launches) / t
    line = [json.loads(x) for x in out.stdout.strip().split("\n") if x.startswith("{")]
    return {"envs_total": envs, "n_gpus": 1, "workload": workload, "regime": "episode" if regime == "episode" else f"substeps k={k}",
            "ncu_launches": len(launches), "ncu_kernel_s": t, "ncu_fp64_tflops": flop / t / 1e12, "ncu_dram_gbs": dram / t / 1e9,
            "ncu_achieved_occupancy_pct": occ, "ncu_pipe_fp64_active_pct": pipe, "ncu_issue_active_pct": issue,
            "env_steps_under_ncu": line[-1]["env_steps"] if line else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="colav_iw")
    ap.add_argument("--envs", default="1000,10000,100000,1000000,10000000")
    ap.add_argument("--ks", default="1,16,128")
    a = ap.parse_args()
    for envs in [int(x) for x in a.envs.split(",")]:
        for regime, k in [("episode", None)] + [("substeps", int(k)) for k in a.ks.split(",")]:
            try:
                print(json.dumps(capture(a.workload, envs, regime, k)), flush=True)
            except Exception as exc:
                print(json.dumps({"envs_total": envs, "regime": regime, "k": k, "error": f"{type(exc).__name__}: {exc}"}), flush=True)


if __name__ == "__main__":
    main()
