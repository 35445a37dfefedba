"""Where does the fused Jacobi kernel spend its time?  Aggregates the warp-state samples of an `ncu --set full
--import-source on` capture of k_jacobi_fused over named regions of csrc/fs2d_fused.cu.

    ncu --set full --clock-control none --import-source on -k regex:k_jacobi_fused -s 2 -c 1 -o gpurun_out/fused_T8 python scripts/fused_prof.py 8 8 8
    python scripts/fused_regions.py gpurun_out/fused_T8.ncu-rep > profiles/r02_fused_regions.txt

Regions are ranges of SOURCE LINES of the file as embedded in the report, found by anchor strings (so the table follows
the code as it was profiled); inlined helpers keep their own lines, so "poll neighbours" is the body of flag_wait_ge /
ld_acquire_smem wherever it was inlined, "mbarrier wait" the body of mbar_wait, and so on.
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# the page has one section per source file ("File Path" row, then a header row): keep the kernel's own file
starts = [i for i, r in enumerate(rows) if r and r[0] == "File Path"]
sec = [i for i in starts if len(rows[i]) > 1 and rows[i][1].endswith("fs2d_fused.cu")][0]
nxt = min([i for i in starts if i > sec], default=len(rows))
rows = rows[sec:nxt]
h = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hd = rows[h]
li, ai = hd.index("Line No"), hd.index("Address")
ni, ii = hd.index("# Samples"), hd.index("Instructions Executed")
stall = [i for i, c in enumerate(hd) if c.startswith("stall_") and "Not Issued" not in c]
lines = {}          # line number -> (text, samples, instructions, stall counter)
for r in rows[h + 1:]:
    if len(r) != len(hd) or r[ai] != "-" or not r[li].isdigit():     # (ncu does not escape quotes: lines with string literals may split wrongly)
        continue
    st = collections.Counter({hd[c][6:]: int(r[c] or 0) for c in stall})
    lines[int(r[li])] = (r[1], int(r[ni] or 0), int(r[ii] or 0), st)
# the report's per-line rows only cover lines that own instructions; the anchors (signatures, comments) come from the file --
# which must be the one that was profiled: every line the report does carry has to match
src_path = sys.argv[2] if len(sys.argv) > 2 else str(__import__("pathlib").Path(__file__).resolve().parents[1] / "2d-fluid-simulator_b200" / "csrc" / "fs2d_fused.cu")
text = {n + 1: t for n, t in enumerate(open(src_path).read().split("\n"))}
alnum = lambda t: "".join(ch for ch in t if ch.isalnum())  # noqa: E731
stale = [n for n, (t, _, _, _) in lines.items() if alnum(text.get(n, "")) != alnum(t)]
if stale:
    raise SystemExit(f"{src_path} is not the source this report was taken from (first differing line {stale[0]}): profile again")


def find(anchor: str, after: int = 0) -> int:
    for n in sorted(text):
        if n > after and anchor in text[n]:
            return n
    raise SystemExit(f"anchor not found in the profiled source: {anchor!r}")


def fn_range(name: str) -> tuple[int, int]:
    """source lines of a __device__ helper: from its signature to the next top-level closing brace"""
    a = find(name + "(")
    for n in sorted(text):
        if n > a and text[n].startswith("}"):
            return a, n
    return a, a + 1


body0 = find("void jacobi_fused_body(")
pure = find("if (!tile_slow) {", body0)
pure_regs = find("__syncwarp();   // the slice is in registers", pure)
pure_loop = find("for (int s = 0; s < g.T; ++s) {", pure_regs)
pure_store = find("lv += g.T;", pure_loop)
slow = find("} else {", pure_store)
slow_iter = find("int cur = VOFF_P0, nxt = VOFF_SRC;", slow)
slow_end = find("lv += g.T;", slow_iter)
tail = find("parity ^= 1;", slow_end)
end = find("#undef FS2D_ISSUE", tail)
regions = [
    ("helper: mbarrier wait (TMA arrival)", *fn_range("void mbar_wait")),
    ("helper: TMA issue (expect_tx + UTMALDG)", find("void mbar_expect_tx("), find("// The iteration loop of the autonomous warps addresses")),
    ("helper: shared ld/st by address (edge rows)", find("float4 lds4_s("), find("// progress counters of the warps")),
    ("helper: release store of the counter", find("void st_release_s("), find("// the smaller of two counters")),
    ("helper: poll neighbours (peek / wait)", find("int flag_peek2("), find("// generic-proxy accesses")),
    ("helper: proxy fence", find("void fence_proxy_async("), find("void fence_proxy_async(") + 1),
    ("open tile: shuffles", *fn_range("void jacobi_rows_shuffles")),
    ("open tile: rows 1..6 (no neighbour needed)", *fn_range("void jacobi_rows_inner")),
    ("open tile: rows 0, 7 (neighbours' edge rows)", *fn_range("void jacobi_rows_outer")),
    ("slow tile: jacobi_rows (masked)", *fn_range("void jacobi_rows")),
    ("slow tile: resolve / fix-up helpers", find("uint32_t f_resolve("), find("constexpr int FS_CAP")),
    ("kernel prologue", body0, find("while (entry >= 0) {", body0)),
    ("tile head (list entries, geometry)", find("while (entry >= 0) {", body0), pure),
    ("open tile: staging -> registers", pure, pure_regs),
    ("open tile: refill issue", pure_regs, pure_loop),
    ("open tile: publish / release / edge loads", pure_loop, pure_store),
    ("open tile: store", pure_store, slow),
    ("slow tile: load, masks, list, table", slow, slow_iter),
    ("slow tile: iterations (fix-up, barriers)", slow_iter, slow_end),
    ("slow tile: refill issue + store", slow_end, tail),
    ("loop tail", tail, end),
]
tot_s = sum(v[1] for v in lines.values())
tot_i = sum(v[2] for v in lines.values())
print(f"# {rep}: warp-state samples per region of csrc/fs2d_fused.cu (k_jacobi_fused, one launch); total samples {tot_s}, warp instructions {tot_i}")
seen = set()
for name, a, b in regions:
    sel = [n for n in lines if a <= n < b and n not in seen]
    seen.update(sel)
    s = sum(lines[n][1] for n in sel)
    i = sum(lines[n][2] for n in sel)
    st = collections.Counter()
    for n in sel:
        st.update(lines[n][3])
    top = ", ".join(f"{k} {v}" for k, v in st.most_common(4) if v)
    print(f"{name:46s} lines {a:4d}-{b:4d}  samples {s:6d} {100 * s / max(tot_s, 1):5.1f} %  warp-inst {i:10d} {100 * i / max(tot_i, 1):5.1f} %  [{top}]")
rest = [n for n in lines if n not in seen]
s, i = sum(lines[n][1] for n in rest), sum(lines[n][2] for n in rest)
print(f"{'(other lines)':46s} {'':15s}  samples {s:6d} {100 * s / max(tot_s, 1):5.1f} %  warp-inst {i:10d} {100 * i / max(tot_i, 1):5.1f} %")
