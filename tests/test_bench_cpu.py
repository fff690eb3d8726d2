"""bench.py's CPU-only legs: the reference arm (--impl reference) prints one contract-shaped JSON line on rank 0 and
nothing on the other ranks.  (The GPU arm is exercised on the B200 by the driver and by tools/final_r01f.sh.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--cpu-showers", "2", "--gpus", "1"], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)


def test_reference_arm_json_line():
    r = _run({"RANK": "0", "WORLD_SIZE": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "showers/sec" and d["unit"] == "showers/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("SM shower: 10 GeV photon into lead")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "showers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["particle_steps_per_sec"] > d["value"]          # hundreds of particle-steps per shower


def test_reference_arm_other_ranks_do_nothing():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
