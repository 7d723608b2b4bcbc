"""The bench.py JSON contract, checked on the committed lines of the last measured run (profiles/r2_bench_*.json): a CPU-side
guard that the keys the driver reads are there and consistent (no GPU needed; bench.py itself runs on the GPU box)."""
import json
import pathlib

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def load(name):
    p = ROOT / "profiles" / name
    if not p.exists():
        pytest.skip("%s not committed" % name)
    return json.loads(p.read_text().strip().splitlines()[-1])


def test_b200_line_has_the_contract_keys():
    d = load("r2_bench_default.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["warmup"] >= 3 and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None                       # BASELINE.md holds no published number for this metric
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 1920 * 1080 * 3 and e["d2h_bytes_per_step"] == 3840 * 2160 * 3
    assert 0 < e["value"] < d["value"]                    # host copies inside the timed region: slower than device-resident
    assert d["gpu_launches"] >= 3 * d["steps"]            # three kernels per step
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    # achieved = algorithmic flop per launch / measured kernel time
    assert abs(r["achieved"] - r["algorithmic_flop_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e12) < 1e-6 * r["achieved"]
    assert r["algorithmic_flop_per_launch"] == 3840 * 2160 * 16064
    assert r["kernel_ms"] < d["ms_per_step"]
    # value = whole-job output pixels / step time
    assert abs(d["value"] - 3840 * 2160 / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * d["value"]
    for m in d["roofline_mem"]:
        assert abs(m["frac"] - m["gbs"] / m["peak_gbs"]) < 1e-9
        assert abs(m["gbs"] - m["bytes"] / (m["ms"] * 1e-3) / 1e9) < 1e-6 * m["gbs"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["unit"] == d["unit"] and c["value"] > 0
    assert c["parity"]["ok"] is True and c["parity"]["max_abs_lsb"] <= 1 and c["parity"]["psnr_db"] >= 50.0
    assert c["parity"]["y_plane_rel_max"] is not None and c["parity"]["y_plane_rel_max"] <= 1e-5
    clk = d["clocks"]
    assert not set(clk["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_line_is_the_same_metric_and_config():
    d, r = load("r2_bench_default.json"), load("r2_bench_reference.json")
    assert r["impl"] == "reference"
    for k in ("metric", "unit", "higher_is_better"):
        assert r[k] == d[k], k
    assert r["config"]["workload"] == d["config"]["workload"]
    assert r["e2e"]["value"] == r["value"] and r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0
    assert r["cpu_baseline"]["value"] == r["value"] and "whole workload frame" in r["cpu_baseline"]["sample"]


def test_sub_results_cover_c3_c4_and_the_chain():
    d = load("r2_bench_default.json")
    sub = d.get("sub", {})
    assert {"c3", "c4", "chain_x4"} <= set(sub)
    assert sub["c4"]["config"]["frames_per_step"] == 4096         # north_star: the whole batch
    assert sub["c3"]["scaling"] == "strong" and "7680x4320" in sub["c3"]["config"]["workload"]
    assert sub["chain_x4"]["ms_per_call"] > 0
