"""CPU: the reference arm of bench.py (`--impl reference`) honours the JSON contract; under a
multi-rank launch only rank 0 prints.  Sizes are shrunk through the PCU_BENCH_SCALE test hook."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra_env, *flags):
    env = dict(os.environ, PCU_BENCH_SCALE="0.01", **extra_env)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", *flags], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    return out.stdout.strip()


def test_reference_arm_json_contract():
    for wl in ("c3", "c2", "c4"):
        line = json.loads(run({}, "--workload", wl).splitlines()[-1])
        assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
        assert line["value"] > 0 and line["ms_per_step"] > 0 and line["higher_is_better"] is True
        assert line["vs_baseline"] is None and line["dtype"] == "f32" and line["data"] == "synthetic"
        assert "workload" in line["config"] and "model" not in line["config"]
        cb = line["cpu_baseline"]
        assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
        assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                               "d2h_bytes_per_step": 0}
        assert line["unit"] == ("query-points/s" if wl == "c3" else "queries/s")


def test_reference_arm_other_ranks_stay_silent():
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2") == ""
    assert json.loads(run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}, "--gpus", "2").splitlines()[-1])["n_gpus"] == 2
