"""The reference arm of bench.py (`--impl reference`) on the host cores: runs without a GPU, so its JSON contract is
checked here — one line on stdout, the keys the driver reads, rank 0 only under a multi-rank launch."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*flags, env=None):
    result = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *flags],
                            capture_output=True, text=True, timeout=280, cwd=ROOT,
                            env=dict(os.environ, **(env or {})))
    assert result.returncode == 0, result.stderr[-2000:]
    return result.stdout


def test_reference_arm_prints_one_contract_line():
    lines = [l for l in run_bench("--steps", "1", "--warmup", "1", "--ref-seconds", "0.5",
                                  "--particles", "64", "--cells", "5").splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "ecmc_events_per_sec" and line["unit"] == "events/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    baseline = line["cpu_baseline"]
    assert baseline["kind"] in ("reference", "port") and baseline["cores"] >= 1 and baseline["value"] == line["value"]
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_without_work():
    out = run_bench("--gpus", "2", "--steps", "1", "--warmup", "1",
                    env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.strip() == ""
