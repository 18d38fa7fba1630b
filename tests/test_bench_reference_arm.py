"""bench.py --impl reference on CPU (no GPU, no balm_b200 code on that path): prints ONE JSON line with the contract keys,
times a bounded sample per step and extrapolates with slope + intercept."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                          "--poses", "12", "--voxels", "5000", "--cpu-sample-small", "40", "--cpu-sample-large", "160"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ba_iterations_per_sec" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 3
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 4 and d["cpu_baseline"]["extrapolated"] is True
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    det = d["cpu_baseline"]["detail"]
    assert det["sample_voxels"] == [40, 160] and det["eval_s_per_voxel"] > 0 and det["eval_intercept_s"] >= 0
    # ms_per_step is the MEASURED time of a bounded sample step; the extrapolated iteration time is reported beside it
    assert d["ms_per_step"] > 0 and d["ms_per_iteration_extrapolated"] > 0
    assert abs(d["value"] - 1e3 / d["ms_per_iteration_extrapolated"]) <= 1e-9 * d["value"]
