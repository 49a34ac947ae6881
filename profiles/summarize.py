"""Summarise ncu outputs brought back in gpurun_out/ into the small text files committed under profiles/.
  python profiles/summarize.py launches gpurun_out/launches_X.csv      -> per-kernel totals / shares
  python profiles/summarize.py report gpurun_out/prof_X.ncu-rep        -> key metrics per captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_shared_mem",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_active.avg"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        ms = v / 1e6 if u.startswith("ns") else v / 1e3 if u.startswith("us") else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += ms
        tot += ms
    print(f"# {path}: total {tot:.1f} ms over {sum(n for n, _ in agg.values())} launches "
          "(ncu per-launch times are cold-cache and serialised: compare SHARES)")
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
        print(f"{ms:10.2f} ms {100 * ms / tot:5.1f}%  n={n:5d}  {k[:100]}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h in KEYS or h == "Kernel Name"]
    for r in rows[2:]:
        print("----")
        for i in idx:
            print(f"{hdr[i]} [{units[i]}] = {r[i]}")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
