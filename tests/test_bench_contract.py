"""bench.py contract checks that need no GPU: the reference arm (oracle port on the host cores) runs and prints one JSON
line with the keys the driver reads, and its `config` is the one the GPU arm prints (same workload description)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.parametrize("workload", ["c1_bgk_f64_64", "c4_dugks_f64_2048"])
def test_reference_arm_prints_the_contract_line(workload):
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "MLUPS" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 1 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "SAMPLED" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import bench
    ours = bench.config_of(workload, 1)
    for k, v in ours.items():
        assert line["config"][k] == v, k


def test_every_baseline_config_has_a_workload():
    import bench
    with open(os.path.join(ROOT, "BASELINE.json")) as fh:
        configs = json.load(fh)["configs"]
    assert len(configs) == 5
    names = set(bench.WORKLOADS)
    for want in ("c1_bgk_f64_64", "c2_trt_f64_1024", "c3_rr_f64_8192", "c3_rr_f32_8192", "c4_dugks_f64_2048", "c5_bgk_f64_slab", "c5_bgk_f64_strong"):
        assert want in names
