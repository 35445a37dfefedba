"""Turn gpurun_out/launches.csv + gpurun_out/step_full.ncu-rep into small tracked summaries under profiles/.

    python scripts/summarize_profiles.py r01_v7 [gpurun_out/step_full.ncu-rep]

profiles/<tag>_launches_summary.csv   per-kernel launch counts / times / SHARE of the step (ncu launch list of
                                      `python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-config`)
profiles/<tag>_kernels_ncu_full.csv   one row per distinct kernel from the `ncu --set full` capture of the same command
profiles/dominant_kernel_traffic.json DRAM bytes per launch of the dominant kernel (read by bench.py -> roofline.traffic)
"""
import collections, csv, json, subprocess, sys

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
out = [f"# {tag}: ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200, python bench.py --steps 1 --warmup 1 "
       "--no-e2e --no-cpu-baseline --no-extra-config (8192x8192 cells/GPU, 80 Jacobi sweeps/step); ns; cold-cache serialised replays: compare SHARES",
       "kernel,launches,total_ns,avg_ns,share"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k},{n},{t:.0f},{t / n:.0f},{t / tot:.4f}")
open(f"profiles/{tag}_launches_summary.csv", "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
rep = sys.argv[2] if len(sys.argv) > 2 else None
if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hd = rr[0]
    keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    idx = [hd.index(k) for k in keep if k in hd]
    lines = [f"# {tag}: ncu --set full --clock-control none --import-source on --kernel-id ::regex:k_:3 (3rd invocation of each kernel), "
             "python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-config; units row, then one row per kernel",
             ",".join(hd[i] for i in idx)]
    for r in rr[1:]:
        lines.append(",".join('"' + r[i] + '"' if "," in r[i] else r[i] for i in idx))
    open(f"profiles/{tag}_kernels_ncu_full.csv", "w").write("\n".join(lines) + "\n")
    ni, ti, ri, wi = hd.index("Kernel Name"), hd.index("gpu__time_duration.sum"), hd.index("dram__bytes_read.sum"), hd.index("dram__bytes_write.sum")
    for r in rr[2:]:
        print(f"{r[ni][:36]:36s} {float(r[ti]):9.1f} {rr[1][ti]}  dram rd {float(r[ri]):8.1f} wr {float(r[wi]):8.1f} {rr[1][ri]}")
    dom = [r for r in rr[2:] if "k_jacobi_fused" in r[ni] and "emit" not in r[ni]]
    if dom:
        import hashlib
        from pathlib import Path

        r = dom[0]
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rd, wr = float(r[ri]) * scale[rr[1][ri]], float(r[wi]) * scale[rr[1][wi]]
        T = int(__import__("os").environ.get("FUSED_T", "7"))  # iterations of the captured launch: the 3rd fused launch of a step; the
        # schedule of the bench's 80-iteration update is [7 x 9, 8, 8 (emitting), literal] (bench line: roofline.update_schedule)
        cells = 8192 * 8192
        # DRAM bytes of a whole step: every kernel's bytes (its one captured launch) x its launches per step (launch list;
        # one k_limit launch per step).  The fused passes of other sizes are counted with the T = 8 figure.
        per_kernel = {x[ni].split("(")[0].replace("void ", "").replace("fs2d::", ""): (float(x[ri]) * scale[rr[1][ri]] + float(x[wi]) * scale[rr[1][wi]])
                      for x in rr[2:]}
        # one step = the launches between the first and the second k_limit of the list (an eager step; later parts of the run
        # contain the bench's stand-alone timing of the fused kernel)
        names = [r[ki].split("(")[0].replace("void ", "").replace("fs2d::", "") for r in data if len(r) > mi]
        lim = [i for i, n in enumerate(names) if n.startswith("k_limit")]
        one_step = names[lim[0] + 1:lim[1] + 1] if len(lim) >= 2 else names
        steps = 1
        step_counts = collections.Counter(one_step)
        step_bytes = sum(per_kernel.get(k, 0.0) * n for k, n in step_counts.items())
        h = hashlib.sha256()
        for name in ("fs2d_fused.cu", "fs2d_common.cuh"):
            h.update((Path(__file__).resolve().parents[1] / "2d-fluid-simulator_b200" / "csrc" / name).read_bytes())
        ia = hd.index("smsp__issue_active.avg.pct_of_peak_sustained_active")
        json.dump({"kernel": f"{r[ni].split('(')[0]} (T={T} iterations per launch, 8192x8192 cells)",
                   "dram_bytes_read_per_launch": int(rd), "dram_bytes_write_per_launch": int(wr),
                   "dram_bytes_per_launch": int(rd + wr), "iterations_per_launch": T,
                   "dram_bytes_per_cell_iteration": round((rd + wr) / cells / T, 2),
                   "algorithmic_bytes_per_launch": 12 * cells * T, "launch_duration_us_under_ncu": float(r[ti]),
                   "issue_active_pct": float(r[ia]), "step_dram_bytes": int(step_bytes), "step_launches": dict(step_counts),
                   "source_sha256": h.hexdigest(),
                   "source": f"profiles/{tag}_kernels_ncu_full.csv (ncu --set full --clock-control none, one capture; bench.py refuses "
                             "this file when the kernel sources no longer hash to source_sha256)"},
                  open("profiles/dominant_kernel_traffic.json", "w"))
        print(open("profiles/dominant_kernel_traffic.json").read())
