"""bench.py contract checks that need no GPU: the reference arm (the C restatement of the step on the host cores) runs here and prints
the line the driver expects; both arms share one `config` object; the defaults are N = 1 with a K / W that finish in minutes."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("_bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "env_steps_per_sec" and line["unit"] == "env-steps/s"
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["warmup"] == 3 and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    bench = _bench_module()
    assert line["config"] == bench.config_of(16384, 1)                 # the GPU arm prints config_of(...) too: same_config
    assert "workload" in line["config"] and "model" not in line["config"]


def test_defaults_and_choices():
    bench = _bench_module()
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        a = bench.parse()
    finally:
        sys.argv = argv
    assert a.gpus == 1 and a.impl == "ours" and a.steps >= 20 and a.warmup >= 3 and a.envs == 16384
    assert a.metrics_collective == "peer"
    assert bench.METRIC == "env_steps_per_sec" and bench.METRICS_EVERY == 16
