"""bench.py's reference arm runs on the CPU, so its side of the driver contract can be checked here: exactly one JSON line on
stdout with the agreed keys, the reference itself behind it when oracle/_ref/ref_driver is present."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "3000", *extra], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    from oracle import refbin
    line = _run()
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "phonon_steps_per_s" and line["unit"] == "phonon-steps/s"
    assert line["vs_baseline"] is None and line["higher_is_better"] is True and line["dtype"] == "f64"
    assert line["config"]["workload"] == "C2-slab100nm-si"
    cb = line["cpu_baseline"]
    assert cb["kind"] == ("reference" if refbin.driver_available() else "port") and cb["cores"] >= 1 and cb["value"] == line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_the_other_workloads():
    for wl in ("C1-film100nm-si", "C4-tube-si"):
        line = _run("--workload", wl)
        assert line["config"]["workload"] == wl and line["value"] > 0


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
