"""CPU: the reference arm of bench.py (`--impl reference`: the unmodified reference modules on host cores when a reference
tree is present -- /root/reference here, baseline/_ref on the GPU box -- else the oracle port) prints ONE JSON line with the
contract's keys, labels the ray count it actually ran; under a multi-rank launch only rank 0 prints.  (The B200 arm needs a GPU; its line is checked by the
driver's run and archived under profiles/.)"""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--rays", "64", "--samples", "16", "--hash-size", "12", "--res", "64", "--views", "4",
                        "--gpus", env.get("WORLD_SIZE", "1")], capture_output=True, text=True, env=env, timeout=300)
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
    from oracle import ref_loader
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    # the label names what ran: 64 rays per step were asked for and fit the budget
    assert d["config"]["reference_rays_per_step"] == 64 and d["config"]["same_rays_per_step_as_b200_arm"] is True
    assert "64 rays x 16 samples" in d["config"]["workload"] and "64 rays x 16 samples" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    if isinstance(base.get("metric"), str):
        assert d["metric"] in base["metric"] or base["metric"] in d["metric"] or "rays" in base["metric"]


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
