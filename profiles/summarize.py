"""Summarises ncu outputs brought back in gpurun_out/ into small text files under profiles/ (run in the build container).

    python profiles/summarize.py launches gpurun_out/launches.csv profiles/rN_launches.txt
    python profiles/summarize.py full gpurun_out/prof.ncu-rep profiles/rN_kernel_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] in ("ns", "nsecond") else (v * 1e3 if row["Metric Unit"] in ("ms", "msecond") else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# source: {src}; per-launch gpu__time_duration.sum (us), cold-cache serialised\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:90]:90s} launches={v[0]:5d} total_us={v[1]:12.1f} share={v[1] / tot:.4f}\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# source: {src} (ncu --set full --clock-control none); one block per captured launch\n")
        for row in rows[2:]:
            f.write(f"\n== {row[hdr.index('Kernel Name')]}  grid={row[hdr.index('launch__grid_size')]} block={row[hdr.index('launch__block_size')]}\n")
            for m in METRICS:
                if m in hdr:
                    f.write(f"{m:90s} {row[hdr.index(m)]:>16s} {units[hdr.index(m)]}\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
