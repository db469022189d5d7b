"""CPU tier: the bench lines committed under profiles/ carry every key of the
bench contract (the driver and the judge parse them), with sane values."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_headline_bench_line_schema():
    d = _line("r02_bench_c4_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "clocks",
              "e2e", "gpu_launches", "roofline", "cpu_baseline", "sustained", "residual", "jacobian",
              "c5", "small_configs"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["data"] == "synthetic" and d["vs_baseline"] is None and "workload" in d["config"]
    assert "l2" in d["config"]  # timing rule: says whether L2 is flushed or the inputs exceed it
    assert d["gpu_launches"] == d["steps"] > 0 and d["warmup"] >= 3
    # value = contributions / device time
    assert abs(d["value"] - 96000000 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert "discretize_linear" in e["what"]  # the API call, not a hand-pipelined loop
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0.0 < r["frac"] < 1.0
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]) and d["clocks"]["samples"] > 10
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert not bad & set(d["clocks"]["reasons"])
    assert d["sustained"]["launches"] * d["sustained"]["ms_per_step"] >= 1000.0  # >= 1 s


def test_reference_arm_runs_the_same_config():
    ours, ref = _line("r02_bench_c4_n1.json"), _line("r02_bench_reference.json")
    assert ref["impl"] == "reference" and ref["config"] == ours["config"]
    assert ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["e2e"] == {"value": ref["value"], "unit": ref["unit"], "h2d_bytes_per_step": 0,
                          "d2h_bytes_per_step": 0}
    full = ref["full_size_step"]
    assert full["size"] == 4001 and abs(full["value"] / ref["value"] - 1.0) < 0.2  # rate is size-independent


def test_scaling_lines_are_consistent():
    one = _line("r02_bench_c4_n1.json")
    for n in (2, 4, 8):
        d = _line(f"r02_scale_n{n}.json")
        assert d["n_gpus"] == n and d["metric"] == one["metric"] and d["scaling"] == "strong"
        assert d["config"]["workload"] == one["config"]["workload"]
        assert d["value"] > one["value"]  # whole-job throughput grows with the GPUs
        assert d["dist_parity"] is True  # N-rank results == 1-rank results, checked in the run
        assert d["c5"]["ms_per_step"] < one["c5"]["ms_per_step"]  # the leg WITH a halo exchange


def test_config5_full_size_table():
    """BASELINE config 5 (343^3 points, 200 M cells) at 1 / 2 / 4 / 8 GPUs."""
    t = [_line(f"r02_c5_343_n{n}.json") for n in (1, 2, 4, 8)]
    assert all("343^3" in d["config"]["workload"] for d in t)
    for a, b in zip(t, t[1:]):
        assert b["n_gpus"] == 2 * a["n_gpus"] and b["ms_per_step"] < 0.6 * a["ms_per_step"]
    ns = t[3]["newton_solve"]
    assert ns.get("bicgstab", ns)["newton_steps"] == 2  # (one record per Krylov solver since CG exists)
