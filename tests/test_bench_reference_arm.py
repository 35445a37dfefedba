"""bench.py --impl reference on CPU: the arm the driver times beside ours.  Small sample grid so that it runs in seconds;
checks the JSON contract of the line and the two things round 1 got wrong -- the OpenMP team under torch.distributed.run
(which exports OMP_NUM_THREADS=1 to every rank) and who prints (rank 0 alone)."""
from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
ARGS = ["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample-rows", "96", "--cols", "128", "--jacobi", "6"]


def _line(out: str) -> dict:
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, f"expected ONE JSON line, got {len(lines)}:\n{out}"
    return json.loads(lines[0])


def _check(d: dict, n_gpus: int) -> None:
    assert d["impl"] == "reference" and d["metric"] == "cell-updates/s" and d["unit"] == "cell-updates/s"
    assert d["n_gpus"] == n_gpus and d["steps"] == 1 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["dtype"] == "f32"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "taichi") and cb["value"] == d["value"] and len(cb["repeat_values"]) == 3
    assert cb["cores"] == (os.cpu_count() or 1), "the reference arm must use every host core, whatever OMP_NUM_THREADS says"
    assert d["e2e"] == {"value": d["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    s = d["config"]["reference_sample"]
    assert s["rows"] == 96 and s["cols"] == 128 and s["cells"] == 96 * 128


def test_reference_arm_single_process():
    env = dict(os.environ, OMP_NUM_THREADS="1")      # what a launcher would export: must not shrink the team
    r = subprocess.run([sys.executable, "bench.py", *ARGS], cwd=REPO, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    _check(_line(r.stdout), 1)


def test_reference_arm_under_torchrun_prints_once():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", "bench.py", "--gpus", "2", *ARGS]
    r = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    _check(_line(r.stdout), 2)
