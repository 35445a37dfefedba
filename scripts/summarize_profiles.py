"""Turn gpurun_out/launches.csv + *.ncu-rep into small tracked summaries under profiles/."""
import collections, csv, subprocess, sys
tag = sys.argv[1]
rows = list(csv.reader(open("gpurun_out/launches.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[h], rows[h + 1:]
ki, mi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= mi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("fs2d::", "")
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[mi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
out = [f"# {tag}: ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200, python bench.py --steps 1 --warmup 1 (8192x8192 cells/GPU, 80 Jacobi sweeps/step); ns; cold-cache serialised replays: compare SHARES",
       "kernel,launches,total_ns,avg_ns,share"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k},{n},{t:.0f},{t / n:.0f},{t / tot:.4f}")
open(f"profiles/{tag}_launches_summary.csv", "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))
rep = sys.argv[2] if len(sys.argv) > 2 else None
if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hd = rr[0]
    keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    idx = [hd.index(k) for k in keep if k in hd]
    lines = [f"# {tag}: ncu --set full --clock-control none --import-source on (per launch), units row then one row per captured launch",
             ",".join(hd[i] for i in idx)]
    for r in rr[1:]:
        lines.append(",".join('"' + r[i] + '"' if "," in r[i] else r[i] for i in idx))
    open(f"profiles/{tag}_jacobi_ncu_full.csv", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[1:]))
