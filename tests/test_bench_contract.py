"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the CPU port of the dense Eigen path) prints
exactly ONE JSON line on stdout with the keys the driver reads, also under torchrun's OMP_NUM_THREADS=1 environment, and ranks
other than 0 print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--landmarks", "24", "--steps", "3",
                        "--warmup", "1", *args], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line():
    out = _run(dict(OMP_NUM_THREADS="1"))
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "vision-updates/sec" and d["unit"] == "updates/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["steps"] == 3 and d["warmup"] == 1 and d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6
    assert "workload" in d["config"] and d["config"]["landmarks"] == 24
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    assert set(cb["stage_ms"]) == {"propagation", "preprocessing", "correction"}
    assert cb["single_thread"]["cores"] == 1 and cb["structured_cholesky"]["value"] > 0
    assert d["e2e"] == dict(value=d["value"], unit="updates/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_reference_arm_other_ranks_are_silent():
    out = _run(dict(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), args=("--gpus", "2"))
    assert out.strip() == ""
