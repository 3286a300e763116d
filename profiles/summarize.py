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


def _read_multi(src):
    """{launch id: {"name": kernel, metric: value}} of an `ncu --metrics a,b,c --csv` log (one row per launch and metric)."""
    lines = [l for l in open(src) if not l.startswith("==")]
    out = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        if row["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)      # -> us
        if u in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        d = out.setdefault(int(row["ID"]), {"name": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")})
        d[row["Metric Name"]] = v
    return out


def _shapes(log):
    """[(impl, kh, kw, stride, cin, cout, M, res)] from the FCP_LOG_CONV=1 stderr of the same run, in launch order."""
    pat = re.compile(r"\[conv_tc (\d+)\] impl=(\d) k=(\d+)x(\d+) s=(\d) cin=(\d+) cout=(\d+) M=(\d+) res=(\d)")
    return [tuple(int(x) for x in m.groups()[1:]) for m in map(pat.search, open(log)) if m]


def _bn(cout):
    cp = (cout + 31) // 32 * 32
    return 128 if cp % 128 == 0 or cp == 96 else (64 if cp % 64 == 0 else 32)


def launches2(src, shapes_log, dst, traffic_json=None):
    """Per-kernel table (time share, DRAM bytes, achieved DRAM GB/s, tensor-pipe activity) + per-shape table of the tensor-core
    convolution launches (matched to their layer shapes through the FCP_LOG_CONV log) of an ncu multi-metric launch list."""
    import json
    L = _read_multi(src)
    T, R, W, TP = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in L.values():
        a = agg[d["name"]]
        a[0] += 1; a[1] += d[T]; a[2] += d.get(R, 0) + d.get(W, 0); a[3] += d.get(TP, 0) * d[T]
    tot = sum(a[1] for a in agg.values())
    lines = [f"# source: {src} (ncu --metrics time,dram bytes,tensor-pipe activity --clock-control none; per-launch times are cold-cache and",
             "# serialised: compare SHARES; GB/s = DRAM bytes / that time)",
             f"{'kernel':58s} {'n':>5s} {'total_us':>10s} {'share':>7s} {'DRAM MB':>10s} {'DRAM GB/s':>10s} {'tensor%':>8s}"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k[:58]:58s} {a[0]:5d} {a[1]:10.1f} {a[1] / tot:7.4f} {a[2] / 1e6:10.1f} {a[2] / a[1] / 1e3:10.1f} {a[3] / a[1]:8.1f}")
    # ---- tensor-core convolutions: align the captured window with the shape log through the tile-width sequence
    shapes = _shapes(shapes_log)
    conv = [d for d in L.values() if d["name"].startswith("conv_tc_kernel")]
    seq = [int(re.search(r"<(\d+),", d["name"]).group(1)) for d in conv]
    want = [_bn(sh[5]) for sh in shapes]
    off = next((o for o in range(len(want) - len(seq) + 1) if want[o:o + len(seq)] == seq), None)
    lines += ["", f"# tensor-core convolution launches of the window ({len(conv)} launches, offset {off} in the run's launch order), by layer shape;",
              "# alg MB = fp32 activations in + out (+ residual) + packed weights once; TFLOP/s = fp32-equivalent algorithmic FLOPs / cold time",
              f"{'shape':44s} {'n':>4s} {'us/launch':>10s} {'alg MB':>9s} {'DRAM MB':>9s} {'DRAM/alg':>8s} {'GB/s':>8s} {'TFLOP/s':>8s} {'tensor%':>8s}"]
    if off is not None:
        rows = collections.OrderedDict()
        tot_alg = tot_dram = 0.0
        for d, sh in zip(conv, shapes[off:]):
            impl, kh, kw, st, cin, cout, M, res = sh
            sw = 1 if kw == 1 and kh > 1 else st
            alg = 4.0 * (M * st * sw * cin + M * cout * (1 + res)) + (4 if impl == 2 else 8) * kh * kw * cin * cout
            key = f"k{kh}x{kw} s{st} cin{cin} cout{cout} M{M} res{res}"
            if (kh, kw) == (4, 1):                        # the detector stem read from the uint8 image (conv_tc stem mode)
                key = f"stem 7x7 s2 u8 RGB -> {cout} M{M}"
                alg = 3.0 * M * 4 + 4.0 * M * cout + 4 * 256 * cout
            r = rows.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0, 0.0])
            dram = d.get(R, 0) + d.get(W, 0)
            r[0] += 1; r[1] += d[T]; r[2] += alg; r[3] += dram; r[4] += 2.0 * M * cout * kh * kw * cin; r[5] += d.get(TP, 0) * d[T]
            tot_alg += alg; tot_dram += dram
        for key, r in sorted(rows.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"{key:44s} {r[0]:4d} {r[1] / r[0]:10.1f} {r[2] / r[0] / 1e6:9.1f} {r[3] / r[0] / 1e6:9.1f} {r[3] / r[2]:8.2f} "
                         f"{r[3] / r[1] / 1e3:8.0f} {r[4] / r[1] / 1e6:8.1f} {r[5] / r[1]:8.1f}")
        lines.append(f"# all {len(conv)} launches: DRAM {tot_dram / 1e9:.2f} GB vs algorithmic {tot_alg / 1e9:.2f} GB (ratio {tot_dram / tot_alg:.2f})")
        if traffic_json:
            json.dump({"kernel": "conv_tc_kernel", "dram_bytes_per_launch": tot_dram / len(conv), "algorithmic_bytes_per_launch": tot_alg / len(conv),
                       "launches": len(conv), "source": f"profiles/{dst.split('/')[-1]} (ncu dram__bytes_read.sum + dram__bytes_write.sum, average over "
                                                        f"{len(conv)} consecutive convolution launches, 64 images / 64 faces per launch)"}, open(traffic_json, "w"), indent=1)
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def full2(src, dst, shapes_log=None, skip=0):
    """`full` + the layer shape of every captured conv_tc launch (k-th captured launch = launch skip+k of the FCP_LOG_CONV log)."""
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    shapes = _shapes(shapes_log) if shapes_log else []
    col = lambda m: next((i for i, h in enumerate(hdr) if h == m or h.endswith("." + m)), None)
    with open(dst, "w") as f:
        f.write(f"# source: {src} (ncu --set full --clock-control none); one block per captured launch\n")
        for k, row in enumerate(rows[2:]):
            name = row[hdr.index("Kernel Name")].replace("void ", "").replace("unnamed>::", "")
            extra = ""
            if shapes and "conv_tc" in name and int(skip) + k < len(shapes):
                impl, kh, kw, st, cin, cout, M, res = shapes[int(skip) + k]
                extra = f"   layer: {kh}x{kw} stride {st}, {cin} -> {cout}, M = {M} pixels, residual {res}, impl {impl}"
            f.write(f"\n== {name}  grid={row[col('launch__grid_size')]} block={row[col('launch__block_size')]}{extra}\n")
            for m in METRICS:
                c = col(m)
                if c is not None:
                    f.write(f"{m:90s} {row[c]:>16s} {units[c]}\n")
    print(open(dst).read()[:5000])


def table(src_txt, dst):
    """One line per kernel of a `full2` summary: longest captured launch, its DRAM bytes and achieved DRAM GB/s."""
    blocks = open(src_txt).read().split("\n== ")[1:]
    unit_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    unit_t = {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}
    seen = collections.OrderedDict()
    for b in blocks:
        name = b.split("(")[0].split("<")[0]
        get = lambda m: next((l.split()[1:] for l in b.splitlines() if l.startswith(m)), None)
        t, r, w, pct = get("gpu__time_duration.sum"), get("dram__bytes_read.sum"), get("dram__bytes_write.sum"), get("gpu__dram_throughput")
        val = lambda v, u: float(v[0].replace(",", "")) * u[v[1]] if v else 0.0
        seen.setdefault(name, []).append((val(t, unit_t), val(r, unit_b) + val(w, unit_b), pct[0] if pct else "?"))
    lines = [f"# source: {src_txt}; longest captured launch per kernel (ncu --set full, cold cache), DRAM bytes = dram__bytes_read + write",
             f"{'kernel':26s} {'launches':>8s} {'us':>9s} {'DRAM MB':>10s} {'GB/s':>9s} {'% of ncu DRAM peak':>19s}"]
    for k, v in seen.items():
        us, by, pct = max(v)
        lines.append(f"{k:26s} {len(v):8d} {us:9.1f} {by / 1e6:10.1f} {by / us / 1e3:9.1f} {pct:>19s}")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    {"launches": launches, "full": full, "launches2": launches2, "full2": full2, "table": table}[sys.argv[1]](*sys.argv[2:])
