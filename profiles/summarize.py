#!/usr/bin/env python3
"""Turn ncu output into the small text summaries kept under profiles/.

  summarize.py launches <launch-list.csv> <out.csv>     per-kernel totals of a
      `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list
  summarize.py kernel <report.ncu-rep> <out-prefix>     selected raw metrics and the
      hottest source lines of a `ncu --set full --import-source on` capture

Needs `ncu` on PATH for the second form (it only reads the report; no GPU).
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    # sector efficiency and warp divergence (north_star's evidence list)
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.ratio", "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
    "dram__cycles_active.avg", "dram__cycles_elapsed.avg",
]


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for x in csv.DictReader(io.StringIO("".join(lines))):
        name = x["Kernel Name"].split("(")[0]
        v = float(x["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(x["Metric Unit"], 1e-6)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "ms_per_launch", "share_pct"])
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            w.writerow([k, cnt[k], f"{v:.3f}", f"{v / cnt[k]:.4f}", f"{100 * v / total:.2f}"])


def kernel(rep, prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    d = {a: (b, c) for a, b, c in zip(names, units, vals)}
    with open(prefix + "_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        w.writerow(["kernel", "", d.get("Kernel Name", ("", ""))[1]])
        for m in METRICS:
            if m in d:
                w.writerow([m, d[m][0], d[m][1]])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur, lines = None, []
    for x in csv.reader(io.StringIO(src)):
        if not x:
            continue
        if x[0] == "File Path":
            cur = x[1].split("/")[-1]
        elif x[0] not in ("Function Name", "Line No", "") and len(x) >= 8 and x[2] == "-":
            try:
                lines.append((cur, int(x[0]), x[1].strip(), int(x[6]), int(x[7])))
            except ValueError:
                pass
    # a kernel compiled at run time from the sources embedded in the library carries line numbers but no readable
    # file: take the text from the source tree (the embedded copy is made from it at build time)
    import os
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pdffoam_b200", "csrc")
    cache = {}
    def text(fn, ln, have):
        if have and not have.startswith("<Unable"):
            return have
        if fn not in cache:
            try:
                cache[fn] = open(os.path.join(csrc, fn)).read().splitlines()
            except OSError:
                cache[fn] = []
        return cache[fn][ln - 1].strip() if 0 < ln <= len(cache[fn]) else have
    lines = [(l[0], l[1], text(l[0], l[1], l[2]), l[3], l[4]) for l in lines]
    ts, te = sum(l[3] for l in lines) or 1, sum(l[4] for l in lines) or 1
    with open(prefix + "_lines.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["file", "line", "stall_samples_pct", "warp_instructions_pct", "source"])
        for l in sorted(lines, key=lambda l: -l[3])[:60]:
            w.writerow([l[0], l[1], f"{100 * l[3] / ts:.2f}", f"{100 * l[4] / te:.2f}", l[2][:110]])


if __name__ == "__main__":
    if len(sys.argv) != 4 or sys.argv[1] not in ("launches", "kernel"):
        sys.exit(__doc__)
    (launches if sys.argv[1] == "launches" else kernel)(sys.argv[2], sys.argv[3])
