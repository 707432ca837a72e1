"""The reference arm of bench.py runs on the host cores only: its JSON line is checked here (keys the
driver reads; rank != 0 prints nothing and exits 0)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "1", "--cpu-samples", "20"], capture_output=True, text=True, env=env,
                          cwd=ROOT, timeout=600)


def test_reference_arm_line(oracle_mod):
    out = _run({"RANK": "0", "WORLD_SIZE": "1"})
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decomp-set samples/sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_are_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0, out.stderr
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]
