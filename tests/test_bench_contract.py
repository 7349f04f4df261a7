"""bench.py's contract, checked without a GPU: the configurations are BASELINE.json's, the default workload is the one
the headline metric is quoted on, and the committed bench lines of the round (profiles/) carry every key the driver and
the tier's measurement rules ask for."""
import json
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_configs_are_baseline_jsons():
    import bench

    assert set(bench.CONFIGS) == {"cfg2", "cfg4", "cfg5"}
    c = bench.CONFIGS
    assert (c["cfg2"]["n"], c["cfg2"]["w"], c["cfg2"]["h"], c["cfg2"]["views"]) == (300_000, 640, 480, 9)
    assert (c["cfg4"]["n"], c["cfg4"]["w"], c["cfg4"]["h"], c["cfg4"]["views"]) == (1_000_000, 1920, 1080, 8)
    assert (c["cfg5"]["n"], c["cfg5"]["w"], c["cfg5"]["h"], c["cfg5"]["global_views"]) == (3_000_000, 3840, 2160, 32)
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert "1M-Gaussian 1920x1080" in base["configs"][3] and "3M-Gaussian 4K" in base["configs"][4]
    src = (ROOT / "bench.py").read_text()
    assert 'add_argument("--config", default="cfg4"' in src and 'add_argument("--gpus", type=int, default=1)' in src


LINE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks"}


@pytest.mark.parametrize("name,n_gpus,workload", [
    ("r02t_bench_default.json", 1, "cfg4"), ("r02y_bench_n2.json", 2, "cfg4"), ("r02r_bench_n8.json", 8, "cfg4"),
    ("r02w_bench_cfg5.json", 1, "cfg5")])
def test_committed_bench_lines_carry_the_contract_keys(name, n_gpus, workload):
    d = json.loads((ROOT / "profiles" / name).read_text())
    assert LINE_KEYS <= set(d), LINE_KEYS - set(d)
    assert d["metric"] == "dn_splatter_train_iter_per_s" and d["unit"] == "iter/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == n_gpus and d["scaling"] == "weak" and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None  # BASELINE.md publishes no number for this metric
    assert d["config"]["workload"].startswith(workload) and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["value"] > 0 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["unit"] == "iter/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] <= d["value"] * 1.02  # measured through the host-buffer path, never just a copy of `value`
    assert abs(d["value"] - d["n_gpus"] * d["config"]["views_per_iter_per_gpu"] * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    ck = d["clocks"]
    assert ck["sm_mhz"] and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(ck["reasons"]))
    if n_gpus == 1:
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
        assert r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        if workload == "cfg4":
            assert r["traffic"] and r["traffic"] > 0
