"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/rNN_launches.txt
    python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep profiles/rNN_kernel.txt
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        if unit in ("ns", "nsecond"):
            v /= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e3
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: %s   total %.1f us over %d launches\n" % (src, tot, sum(v[0] for v in agg.values())))
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-44s n=%5d  sum=%10.1f us  avg=%8.2f us  share=%5.1f%%\n" % (k[:44], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print(open(dst).read())


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
        "smsp__average_warp_latency_issue_stalled_not_selected.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio", "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic"]


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none, source: %s\n" % src)
        if "Kernel Name" in hdr:
            i = hdr.index("Kernel Name")
            f.write("kernels: %s\n" % [r[i].split("(")[0] for r in rows[2:]])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                f.write("%-72s %-14s %s\n" % (w, rows[1][i], [r[i] for r in rows[2:]]))
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
