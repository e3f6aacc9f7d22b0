"""The bench line committed under profiles/ carries every key of the driver's contract (schema check only, CPU)."""
import glob
import json
import os

from tests.conftest import ROOT


def _latest():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_1gpu.json")))
    assert files, "no committed bench line"
    return json.loads(open(files[-1]).read().strip().splitlines()[-1])


def test_bench_line_has_the_contract_keys():
    d = _latest()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(c) and c["kind"] in ("port", "reference")
    e = d["e2e"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(e) and e["h2d_bytes_per_step"] > 0
    assert e["value"] < d["value"] and d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    n = d["nwd"]
    assert n["roofline"]["bound"] == "tensor" and n["e2e"]["value"] < n["value"]
