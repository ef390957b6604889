"""CPU: bench.py's reference arm (the oracle port of vaegan_code.py:66-135 on the host cores) prints the contract's JSON
line - same metric / unit / config as the CUDA arm, `impl`, `cpu_baseline` and a zero-copy `e2e` - and ranks other
than 0 exit without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *flags):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *flags],
                          capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_json_line():
    r = _run({}, "--gpus", "1", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        baseline = json.load(f)
    assert line["impl"] == "reference" and line["metric"] == baseline["metric"]
    assert line["unit"] == "images/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["steps"] == 1 and line["warmup"] == 0 and line["n_gpus"] == 1
    assert line["config"]["workload"] == "cfg2" and "model" not in line["config"]
    # every timed step is one GPU's FULL cfg-2 batch (256 images): the label and the measurement are the same config
    assert line["config"]["batch_per_gpu"] == 256 and line["config"]["measured_batch"] == 256
    assert line["value"] > 0 and abs(line["ms_per_step"] * 1e-3 * line["value"] - 256) < 1e-6 * 256
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "256 images" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""
