"""Summarise ncu reports / launch lists from gpurun_out/ into profiles/ (tracked).

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv profiles/NAME.md [iters]
    python tools/ncu_summary.py full gpurun_out/REPORT.ncu-rep profiles/NAME.md
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active"]


def launches(src, dst, iters):
    rows = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(rows):
        k = r["Kernel Name"].split("(")[0][-60:]
        a = agg.setdefault(k, [0.0, 0])
        a[0] += float(r["Metric Value"]); a[1] += 1
    if iters <= 0:   # one preprocess launch per forward
        iters = max([c for k, (v, c) in agg.items() if "preprocess_kernel" in k] + [1])
    ours = {k: v for k, v in agg.items() if ("<unnamed>::" in k or "sfgs" in k) and "at::" not in k and "std::array" not in k}   # this library's kernels live in anonymous namespaces
    tot = sum(v[0] for v in ours.values())
    other = sum(v[0] for k, v in agg.items() if k not in ours)
    with open(dst, "w") as f:
        f.write(f"# ncu launch list `{src}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write(f"{iters} iteration(s) of fwd+bwd, 1M Gaussians, 1920x1080.  Shares are over this library's kernels; "
                f"torch kernels of the harness (L2 flush fills, loss, copies) add {other / iters / 1e3:.1f} us / iteration.\n\n"
                "| kernel | launches | us / iteration | share |\n|---|---|---|---|\n")
        for k, (v, c) in sorted(ours.items(), key=lambda kv: -kv[1][0]):
            f.write(f"| `{k}` | {c} | {v / iters / 1e3:.1f} | {100 * v / tot:.1f}% |\n")
        f.write(f"\nTotal per iteration (this library's kernels): {tot / iters / 1e3:.1f} us\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none capture `{src}`\n\n")
        for r in rows[2:]:
            f.write(f"## `{r[idx['Kernel Name']][:110]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
            f.write(f"| traffic = dram read + write | {rd + wr:.3f} | {units[idx['dram__bytes_read.sum']]} |\n\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    else:
        full(sys.argv[2], sys.argv[3])
