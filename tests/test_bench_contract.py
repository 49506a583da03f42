"""bench.py contract checks that need no GPU: the reference arm (restated reference algorithm on host cores) prints ONE
JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["config"]["workload"].startswith("C1")


def test_non_root_rank_of_reference_arm_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "1", "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cpu_leg_runs_for_any_core_count(monkeypatch):
    """Round 1's N=1 scaling arm died because the CPU leg indexed 24 samples with cores - 1 = 31 workers: the leg
    now draws exactly one sample per worker, whatever the core count."""
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(os, "cpu_count", lambda: 64)
    r = bench.cpu_baseline_leg(bench.WORKLOADS["c1"])
    assert r["workers"] == 63 and r["units_per_s"] > 0 and "63 forked workers" in r["sample"]


def test_both_arms_share_one_config_object():
    sys.path.insert(0, ROOT)
    import bench
    w = bench.WORKLOADS["c5"]
    cfg = bench.config_of(w, 1, w["samples"])
    assert cfg["workload"].startswith("C5") and cfg["units_per_step"] == w["samples"] * 10000
    assert "model" not in cfg
