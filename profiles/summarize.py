#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries in profiles/.

  python profiles/summarize.py launches gpurun_out/launches_X.csv profiles/X_launches.md "<cmd>"
  python profiles/summarize.py full gpurun_out/prof_X.ncu-rep profiles/X_full.md
  python profiles/summarize.py full gpurun_out/prof_X_raw.csv profiles/X_full.md

`launches`: per-kernel count / total device time / share from a
`ncu --metrics gpu__time_duration.sum --csv` log.  `full`: the handful of counters the
roofline discussion in DESIGN.md uses, from a `ncu --set full` report (read with
`ncu -i ... --page raw --csv`).
"""
import collections
import csv
import io
import subprocess
import sys

FULL_KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
    "sm__warps_active.avg.per_cycle_active",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_bytes.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "sass__inst_executed_register_spilling",
    "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def short(name):
    return name.split("(")[0].replace("void ", "")[:100]


def launches(src, dst, cmd):
    lines = [ln for ln in open(src) if not ln.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        ns = float(r["Metric Value"]) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1.0)
        a = agg.setdefault(k, [0, 0.0, r["Block Size"], r["Grid Size"]])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# launch list: `{cmd}`\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, "
                "serialised launches: compare SHARES, not absolutes).\n\n")
        f.write("| kernel | block | grid | launches | total us | mean us | share |\n")
        f.write("|---|---|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[2]} | {v[3]} | {v[0]} | {v[1] / 1e3:.1f} | "
                    f"{v[1] / 1e3 / v[0]:.1f} | {100 * v[1] / tot:.1f}% |\n")
        f.write(f"\ntotal device time in list: {tot / 1e6:.3f} ms over {len(rows)} launches\n")


def full(src, dst):
    # a .ncu-rep, or the `ncu -i X.ncu-rep --page raw --csv` dump of one made on the GPU box
    # (reports of the large ODE kernels exceed what a gpurun call may bring back)
    if src.endswith(".csv"):
        out = open(src).read()
    else:
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True,
                             text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of `{src.split('/')[-1]}`\n\n")
        for r in body:
            f.write(f"## `{short(r[col['Kernel Name']])}`  grid {r[col['Grid Size']]} block "
                    f"{r[col['Block Size']]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in FULL_KEYS:
                if k in col:
                    f.write(f"| {k} | {r[col[k]]} | {units[col[k]]} |\n")
            f.write("\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
