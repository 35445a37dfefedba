"""Build tests/cuda_emu/_build/libfs2d_emu.so: the UNMODIFIED kernel sources of 2d-fluid-simulator_b200/csrc compiled
for the CPU against the CUDA emulation in tests/cuda_emu/include (see cuda_runtime.h there).  TEST INFRASTRUCTURE ONLY.

The sources are rewritten mechanically, nothing else:
  * `kernel<<<grid, block, smem, stream>>>(args)`  ->  `emu::launch(grid, block, smem, [=]() { kernel(args); })`
  * `extern __shared__ __align__(N) T name[];`     ->  `T *name = (T *)emu::dyn_smem();`
  * `__shared__`                                   ->  `static`   (CTAs run one after the other)
  * the inline-PTX helpers (mbarrier, cp.async.bulk.tensor, fences) -> their emulation in include/cuda.h; the packed
    f32x2 multiply is compiled out with the sources' own -DFS2D_NO_F32X2 switch (per-lane mul.rn == scalar multiply).
Compiled with -ffp-contract=off (the CUDA build uses -fmad=false): same literal fp32 operation order.
"""
from __future__ import annotations

import re
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
REPO = HERE.parents[1]
CSRC = REPO / "2d-fluid-simulator_b200" / "csrc"
OUT = HERE / "_build"
LIB = OUT / "libfs2d_emu.so"

# inline-PTX helper functions replaced wholesale (name -> emulation body)
PTX_HELPERS = {
    "smem_u32": "return 0;", "st_smem": "return 0;",
    "mbar_init": "emu::mbar_init(bar, count);", "st_mbar_init": "emu::mbar_init(bar, count);",
    "mbar_expect_tx": "emu::mbar_expect_tx(bar, bytes);", "st_mbar_expect": "emu::mbar_expect_tx(bar, bytes);",
    "mbar_wait": "emu::mbar_wait(bar, parity);", "st_mbar_wait": "emu::mbar_wait(bar, parity);",
    "tma_load_2d": "emu::tma_load_2d(dst, map, c0, c1, bar);", "st_tma_2d": "emu::tma_load_2d(dst, map, c0, c1, bar);",
    "smem_addr": "return (saddr_t)(uintptr_t)p;",
    "lds4_s": "return *reinterpret_cast<const float4 *>(a);",
    "sts4_s": "*reinterpret_cast<float4 *>(a) = make_float4(x, y, z, w);",
    "st_release_s": "emu::flag_store(reinterpret_cast<int *>(a), v);",
    "flag_peek2": "(void)after; return min(*reinterpret_cast<const int *>(a), *reinterpret_cast<const int *>(b));",
    "flag_wait2": "emu::flag_wait_ge(reinterpret_cast<const int *>(a), v); emu::flag_wait_ge(reinterpret_cast<const int *>(b), v);",
}


def _match_back(text: str, end: int, open_c: str, close_c: str) -> int:
    """index of the `open_c` matching the `close_c` at text[end]"""
    depth = 0
    for k in range(end, -1, -1):
        if text[k] == close_c:
            depth += 1
        elif text[k] == open_c:
            depth -= 1
            if depth == 0:
                return k
    raise ValueError("unbalanced")


def _match_fwd(text: str, start: int, open_c: str, close_c: str) -> int:
    depth = 0
    for k in range(start, len(text)):
        if text[k] == open_c:
            depth += 1
        elif text[k] == close_c:
            depth -= 1
            if depth == 0:
                return k
    raise ValueError("unbalanced")


def rewrite_launches(text: str) -> str:
    out, pos = [], 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            out.append(text[pos:])
            return "".join(out)
        # kernel expression: identifier with optional template arguments, right before <<<
        e = k - 1
        while text[e].isspace():
            e -= 1
        s = e
        if text[s] == ">":
            s = _match_back(text, s, "<", ">") - 1
        while s >= 0 and (text[s].isalnum() or text[s] in "_:"):
            s -= 1
        kernel = text[s + 1:e + 1]
        c_end = text.find(">>>", k)
        cfg = text[k + 3:c_end]
        a0 = text.index("(", c_end)
        a1 = _match_fwd(text, a0, "(", ")")
        args = text[a0 + 1:a1]
        parts, depth, cur = [], 0, ""
        for ch in cfg:
            if ch in "(<":
                depth += 1
            elif ch in ")>":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur.strip())
                cur = ""
            else:
                cur += ch
        parts.append(cur.strip())
        assert len(parts) == 4, (kernel, cfg)
        out.append(text[pos:s + 1])
        out.append(f"emu::launch({parts[0]}, {parts[1]}, {parts[2]}, [=]() {{ {kernel}({args}); }})")
        pos = a1 + 1


def rewrite_helpers(text: str) -> str:
    for name, body in PTX_HELPERS.items():
        m = re.search(r"__device__\s+__forceinline__\s+[\w:]+\s+" + name + r"\s*\(", text)
        if not m:
            continue
        b0 = text.index("{", m.end())
        b1 = _match_fwd(text, b0, "{", "}")
        text = text[:b0] + "{ " + body + " }" + text[b1 + 1:]
    # stand-alone fences
    text = re.sub(r'asm\s+volatile\s*\(\s*"fence[^;]*;"\s*:::\s*"memory"\s*\)\s*;', "/* fence */;", text)
    return text


def transform(text: str) -> str:
    text = rewrite_helpers(text)
    text = re.sub(r"extern\s+__shared__\s+__align__\(\d+\)\s+(\w+)\s+(\w+)\[\];", r"\1 *\2 = (\1 *)emu::dyn_smem();", text)
    text = re.sub(r"\b__shared__\b", "static", text)
    text = rewrite_launches(text)
    return text


def build(force: bool = False) -> Path:
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh"))
    deps = srcs + sorted((HERE / "include").glob("*.h")) + [Path(__file__), REPO / "include" / "fs2d.h"]
    if not force and LIB.exists() and all(d.stat().st_mtime <= LIB.stat().st_mtime for d in deps):
        return LIB
    gen = OUT / "csrc"
    gen.mkdir(parents=True, exist_ok=True)
    (OUT / "include").mkdir(exist_ok=True)
    (OUT / "include" / "fs2d.h").write_text((REPO / "include" / "fs2d.h").read_text())   # "../../include/fs2d.h" of the sources
    cpps = []
    for s in srcs:
        dst = gen / (s.name + (".cpp" if s.suffix == ".cu" else ""))
        dst.write_text(transform(s.read_text()))
        if s.suffix == ".cu":
            cpps.append(dst)
    # the sources include "../../include/fs2d.h" relative to csrc/: _build/csrc/../../include does not exist, so map it
    for f in gen.iterdir():
        f.write_text(f.read_text().replace('"../../include/fs2d.h"', '"../include/fs2d.h"'))
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fno-fast-math", "-march=x86-64-v2", "-fPIC", "-shared",
           "-DFS2D_NO_F32X2", "-DFS2D_EMU", "-Wno-unknown-pragmas", "-Wno-attributes", "-I", str(HERE / "include"), "-o", str(LIB),
           *map(str, cpps)]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
