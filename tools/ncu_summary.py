#!/usr/bin/env python
"""Markdown table of the metrics we quote from an `ncu -i X.ncu-rep --page raw --csv` export.

    ncu -i prof.ncu-rep --page raw --csv > prof.raw.csv
    python tools/ncu_summary.py prof.raw.csv [launch_env_steps ...]

With the env-steps each profiled launch executed (bench.py `launch_env_steps`) it also prints the executed
FP64 instructions / flop per env-step (DFMA = 2 flop) and the DRAM bytes per env-step."""
import csv
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]
STALLS = ["wait", "selected", "not_selected", "short_scoreboard", "long_scoreboard", "branch_resolving", "math_pipe_throttle",
          "no_instruction", "dispatch_stall", "barrier", "lg_throttle", "mio_throttle", "drain", "imc_miss", "sleeping", "membar",
          "tex_throttle", "misc"]


def main():
    csv.field_size_limit(10 ** 9)
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    steps = [float(x) for x in sys.argv[2:]]
    print("`" + data[0][hdr.index("Kernel Name")] + "`\n")
    print("| metric | unit | " + " | ".join(f"launch {chr(65 + i)}" for i in range(len(data))) + " |")
    print("|---|---|" + "---|" * len(data))
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print(f"| `{m}` | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    for li, r in enumerate(data):
        parts = []
        for s in STALLS:
            name = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if name in hdr and r[hdr.index(name)]:
                parts.append((float(r[hdr.index(name)]), s))
        tot = sum(p[0] for p in parts)
        if tot:
            parts.sort(reverse=True)
            print(f"\nwarp stall reasons (launch {chr(65 + li)}, share of stalled-warp cycles per issue): " +
                  ", ".join(f"{s} {100 * v / tot:.1f}%" for v, s in parts[:8]))
    if steps:
        print()
        g = lambda r, m: float(r[hdr.index(m)])
        for r, n in zip(data, steps):
            cyc = g(r, "sm__cycles_elapsed.max")
            f, a, m = (g(r, f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") for k in ("dfma", "dadd", "dmul"))
            to_b = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
            rd = g(r, "dram__bytes_read.sum") * to_b[units[hdr.index("dram__bytes_read.sum")]]
            wr = g(r, "dram__bytes_write.sum") * to_b[units[hdr.index("dram__bytes_write.sum")]]
            print(f"- {n:.0f} env-steps in the launch: {(f + a + m) * cyc / n:.0f} FP64 arithmetic instructions and "
                  f"{(2 * f + a + m) * cyc / n:.0f} FP64 flop per env-step; DRAM {rd / 1e6:.1f} MB read + {wr / 1e6:.2f} MB "
                  f"written = {(rd + wr) / n:.2f} B per env-step")


if __name__ == "__main__":
    main()
