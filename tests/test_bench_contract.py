"""CPU tier: the bench line committed under profiles/ carries every key of the
bench contract (the driver and the judge parse it), with sane values."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_headline_bench_line_schema():
    d = _line("r01_bench_c4.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "clocks",
              "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["data"] == "synthetic" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["gpu_launches"] == d["steps"] > 0 and d["warmup"] >= 3
    # value = contributions / device time
    assert abs(d["value"] - 96000000 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0.0 < r["frac"] < 1.0
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert not bad & set(d["clocks"]["reasons"])


def test_scaling_lines_are_consistent():
    one = _line("r01_bench_c4.json")
    for n in (2, 4, 8):
        d = _line(f"r01_c4_n{n}.json")
        assert d["n_gpus"] == n and d["metric"] == one["metric"] and d["scaling"] == "strong"
        assert d["config"]["workload"] == one["config"]["workload"]
        assert d["value"] > one["value"]  # whole-job throughput grows with the GPUs
