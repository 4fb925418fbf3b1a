"""bench.py on the GPU: ONE JSON line carrying every key of the measurement contract (the smallest workload keeps this
a few seconds)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("workload,extra", [("c1", []), ("c2", ["--batch", "2048"]), ("c4", ["--batch", "1024"])])
def test_bench_prints_the_contract_line(workload, extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--steps", "3", "--warmup", "3"]
                         + extra, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "cpu_baseline"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0 and d["gpu_launches"] > 0
    assert d["unit"] == "transitions/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in d["e2e"], key
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in d["roofline"], key
    # the default workload (c4) reports against HBM; sub-workloads name their own binding roof
    assert d["roofline"]["bound"] in ("hbm", "tensor", "fp32_simt", "fp64_pipe")
    if workload == "c4":
        assert d["roofline"]["bound"] == "hbm" and d["config"]["mstep_status"] == 0
    for key in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert key in d["clocks"], key
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in d["cpu_baseline"], key
    assert d["cpu_baseline"]["kind"] in ("port", "reference")
