"""The bench.py output contract, checked on the committed lines of the last GPU run (profiles/): every key the driver and
the judge read is present with the right type, for our arm and for the reference arm."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads([l for l in f.read().splitlines() if l.startswith("{")][-1])


def _common(d):
    for k, t in (("metric", str), ("value", (int, float)), ("unit", str), ("n_gpus", int), ("steps", int), ("warmup", int),
                 ("ms_per_step", (int, float)), ("higher_is_better", bool), ("scaling", str), ("dtype", str), ("data", str)):
        assert isinstance(d[k], t), k
    assert d["scaling"] == "weak" and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["value"] > 0 and "h2d_bytes_per_step" in e and "d2h_bytes_per_step" in e
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == d["unit"] and c["sample"]


def test_our_line():
    for name, n in (("r1_bench_h.json", 1), ("r1_bench_n2.json", 2), ("r1_bench_n4.json", 4), ("r1_bench_n8.json", 8)):
        d = _line(name)
        assert d["n_gpus"] == n and d["warmup"] >= 3 and d["gpu_launches"] > 0 and "impl" not in d
        if n == 1:
            _common(d)
            r = d["roofline"]
            assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
            assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-4 and r["traffic"] > 0
            assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
            assert 0 < d["roofline_compute"]["frac"] <= 1
            assert d["latency"]["detect_one_message_ms"] < d["latency"]["reference_ms"]
        clk = d["clocks"]
        assert clk["sm_mhz"] > 0 and clk["sm_max_mhz"] >= clk["sm_mhz"] and isinstance(clk["reasons"], list)
        assert not set(clk["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    v = [_line(f"r1_bench_{s}.json")["value"] for s in ("h", "n2", "n4", "n8")]
    assert v[0] < v[1] < v[2] < v[3]                       # whole-job aggregate grows with the number of GPUs


def test_reference_arm_line():
    d = _line("r1_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    _common(d)
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["metric"] == _line("r1_bench_h.json")["metric"] and d["unit"] == _line("r1_bench_h.json")["unit"]
