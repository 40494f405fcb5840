"""bench.py's driver contract, checked without a GPU: the reference arm prints one JSON line with the agreed keys (and times
the restatement, the only thing bench.py may execute from oracle/), the CUDA arm refuses to run without a device instead
of falling back to the CPU, and the roofline's algorithmic-bytes model is the one DESIGN.md states."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reference_arm_line_has_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C4", "--steps", "1",
                          "--warmup", "0", "--ref-problems-per-step", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "registrations/sec" and line["higher_is_better"] is True
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None and line["dtype"] == "f64"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert set(cb["stage_ms_per_problem"]) >= {"tims", "scale_test", "graph", "clique", "rotation", "translation"}
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and "workload" in line["config"]


def test_cuda_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)
    assert not any(l.startswith("{") for l in out.stdout.splitlines())  # no number is printed


def test_algorithmic_bytes_model_and_config_table():
    b = _bench_module()
    # SURVEY §8d / DESIGN §3.1: 48 n (points in) + 8 n ceil(n/64) (bitset rows out) + 4 n (degrees out)
    assert b.bytes_graph(5000) == 48 * 5000 + 8 * 5000 * 79 + 4 * 5000 == 3_420_000
    assert b.bytes_graph(64) == 48 * 64 + 8 * 64 * 1 + 4 * 64
    # BASELINE.json's configs are all selectable, C2 is the default the metric is quoted on
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert {"C1", "C2", "C3", "C4", "C5"} <= set(b.CONFIGS)
    assert b.CONFIGS["C2"]["n"] == 5000 and b.CONFIGS["C3"]["n"] == 10000
    assert b.CONFIGS["C4"]["scaling"] == "strong" and b.CONFIGS["C5"]["scaling"] == "strong" and b.CONFIGS["C2"]["scaling"] == "weak"
    assert "registrations" in json.dumps(base).lower()
