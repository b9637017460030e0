"""CPU: the reference arm of bench.py (`--impl reference`, the oracle port on host cores) prints ONE JSON line with the
contract's keys; under a multi-rank launch only rank 0 prints.  (The B200 arm needs a GPU; its line is checked by the
driver's run and archived under profiles/.)"""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-rays", "32", "--gpus", env.get("WORLD_SIZE", "1")], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "train_rays_per_sec" and d["unit"] == "rays/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    if isinstance(base.get("metric"), str):
        assert d["metric"] in base["metric"] or base["metric"] in d["metric"] or "rays" in base["metric"]


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
