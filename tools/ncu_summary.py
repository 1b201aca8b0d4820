"""Summarise ncu outputs brought back in gpurun_out/ (run here, no GPU):
  python tools/ncu_summary.py launches <launches.csv>
  python tools/ncu_summary.py kernel <file.ncu-rep>
  python tools/ncu_summary.py stalls <file.ncu-rep> [top_n]"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        agg.setdefault(row["Kernel Name"].split("(")[0], []).append(v)
    tot = sum(sum(v) / len(v) * (len(v) / max(1, min(len(x) for x in agg.values() if len(x) > 1))) for v in agg.values())
    print(f"{'kernel':60s} {'n':>4s} {'mean_us':>10s}")
    for n, v in agg.items():
        print(f"{n[:60]:60s} {len(v):4d} {sum(v) / len(v):10.1f}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w:
                print(f"  {w} = {vals[i]} {units[i]}")


def stalls(path, top=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    def col(name):
        for i, h in enumerate(hdr):
            if h.strip() == name:
                return i
        return None
    cs, csrc, csamp = col("Source"), col("#"), col("# Samples")
    if csamp is None:
        csamp = col("Warp Stall Sampling (All Samples)")
    print("columns:", [h for h in hdr][:14])
    data = []
    for r in rows[1:]:
        try:
            data.append((float(r[csamp].replace(",", "")), r[cs][:150]))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1
    for s, src in sorted(data, reverse=True)[:top]:
        print(f"{100 * s / tot:6.2f}%  {src}")


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "kernel":
        kernel(sys.argv[2])
    else:
        stalls(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
