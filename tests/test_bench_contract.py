"""The JSON line of bench.py carries every key the measurement contract names.  Checked on the lines committed under
profiles/ (produced on a B200 by the commands in profiles/README.md) -- no GPU here -- and on what bench.py itself
declares (metric, unit)."""
import glob
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_default_v[45].json")) +
               glob.glob(os.path.join(ROOT, "profiles", "r01_bench_n2_2x1x1_v5.json")))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "gpu_launches", "clocks", "roofline"]


def last_json_line(path):
    return json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_committed_bench_lines_follow_the_contract(path):
    d = last_json_line(path)
    for k in BASE:
        assert k in d, k
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert re.search(r'METRIC\s*=\s*"%s"' % re.escape(d["metric"]), src)
    assert d["unit"] == "integrations/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None          # BASELINE.md publishes no number for this metric on this hardware
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert "l2" in d["config"]                # how the L2 is kept cold between timed iterations
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    if d["n_gpus"] == 1 and "e2e" in d:
        e = d["e2e"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
        b = d["cpu_baseline"]
        assert {"value", "unit", "cores", "kind", "sample"} <= set(b) and b["kind"] in ("reference", "port")
    assert isinstance(baseline, dict)
