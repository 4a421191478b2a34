"""The driver parses ONE JSON line from bench.py; the CPU reference arm runs anywhere, so its line is checked here
(the GPU arm prints the same keys plus roofline / clocks and is exercised on the B200 by the driver)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] == baseline["metric"] and d["unit"] == "env-steps/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("4096 envs/GPU batched step")
    # both arms print the SAME config object (no measured values in it): the driver compares the two lines' configs
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(4096) and d["physics_steps_per_s"] == 10 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "NOT MuJoCo itself" in cb["sample"]
    # one core AND all cores, physics steps/s next to env-steps/s; MuJoCo's own stopping rules
    assert cb["one_core"]["value"] > 0 and cb["one_core"]["physics_steps_per_s"] == 10 * cb["one_core"]["value"]
    assert cb["physics_steps_per_s_per_core"] > 0 and cb["solver"]["tolerance"] == 1e-8 and cb["solver"]["ls_tolerance"] == 0.01
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
