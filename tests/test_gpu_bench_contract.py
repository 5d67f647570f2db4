"""The benchmark contract (task statement): `bench.py` prints ONE JSON line with the agreed keys, for the B200 arm and
for `--impl reference`.  Run on the small self-test workload so that the check takes seconds."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _run(*flags: str) -> dict:
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--workload", "tiny", "--steps", "3", "--warmup", "3", *flags],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines  # exactly one line on stdout
    return json.loads(lines[0])


def test_b200_arm_line():
    d = _run("--no-secondary")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "pass_roofline", "cpu_baseline"):
        assert key in d, key
    assert d["metric"] == "diag+compare draw*params/s" and d["unit"] == "draw*params/s"
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]  # measured through the host-buffer call, not a copy of the device-timed number
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert "traffic" in r and r["kernel"].startswith("k_")
    c = d["cpu_baseline"]
    assert c["value"] > 0 and c["cores"] >= 1 and c["kind"] in ("port", "reference") and c["sample"]
    assert d["parity"]["ok"] is True  # the bounded CPU sample doubles as a parity gate
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_reference_arm_line():
    d = _run("--impl", "reference")
    assert d["impl"] == "reference"
    if "unavailable" in d:
        pytest.skip(d["unavailable"])
    assert d["metric"] == "diag+compare draw*params/s" and d["unit"] == "draw*params/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"]  # the same config object as the B200 arm
