"""The bench.py output contract, checked on lines bench.py prints NOW (not on committed artefacts): every key the driver and
the judge read is present with the right type.  The reference arm runs here on the CPU; our arm and the packing config run
under `-m gpu` on a small slice of the board."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]                 # rank 0 prints ONE JSON line
    return json.loads(lines[0])


def _common(d):
    for k, t in (("metric", str), ("value", (int, float)), ("unit", str), ("n_gpus", int), ("steps", int), ("warmup", int),
                 ("ms_per_step", (int, float)), ("higher_is_better", bool), ("scaling", str), ("dtype", str), ("data", str)):
        assert isinstance(d[k], t), k
    assert d["scaling"] in ("weak", "strong") and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["value"] > 0 and "h2d_bytes_per_step" in e and "d2h_bytes_per_step" in e
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == d["unit"] and c["sample"]


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "2")
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["steps"] == 1 and d["warmup"] == 0
    _common(d)
    assert d["scaling"] == "strong" and d["metric"] == "detected messages/sec at D=65536" and d["unit"] == "messages/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"]
    # the reference arm is quoted on OUR arm's config (the same dictionary for the same launch); its bounded per-step sample is separate
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(1, 65536, 16384, 5, 28, "cuda-core")
    assert d["sample"]["messages_per_step"] == 2 and d["sample"]["host_threads"] == d["cpu_baseline"]["cores"]


def test_traffic_lookup_reads_profiles_and_reports_absence():
    sys.path.insert(0, ROOT)
    import bench
    t, src = bench._traffic("l2_blind_rotate_kernel", 16384)
    assert t and t > 1e9 and "profiles/" in src               # a committed r2 ncu summary backs roofline.traffic
    assert bench._traffic("no_such_kernel", 1) == (None, "absent")


@pytest.mark.gpu
def test_our_line_on_a_slice():
    d = _run("--steps", "1", "--warmup", "3", "--messages-per-step", "256", "--cpu-sample", "2")
    assert "impl" not in d and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 3 and d["scaling"] == "weak"
    _common(d)
    assert d["gpu_launches"] > 0 and d["config"]["messages_per_step_per_gpu"] == 256 and d["config"]["key_switch"] == "cuda-core"
    r = d["roofline"]
    assert r["bound"] == "fp64" and r["kernel"] == "l2_blind_rotate_kernel" and abs(r["frac"] - r["achieved"] / r["peak"]) < 2e-3
    assert 0 < r["frac"] <= 1 and "traffic_source" in r and (r["traffic"] is None or r["traffic"] > 0)
    assert d["roofline_hbm"]["bound"] == "hbm" and d["roofline_hbm"]["unit"] == "GB/s"
    assert 0 < d["roofline_compute"]["frac"] <= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 256 * (512 + 7 + 612) * 2 + 32 and d["e2e"]["d2h_bytes_per_step"] == 33 * 2 * 2048 * 8
    assert d["latency"]["detect_one_message_ms"] < d["latency"]["reference_ms"]
    clk = d["clocks"]
    assert clk["sm_mhz"] > 0 and clk["sm_max_mhz"] >= clk["sm_mhz"] and isinstance(clk["reasons"], list)


@pytest.mark.gpu
def test_pack4096_config_line():
    d = _run("--config", "pack4096", "--steps", "3", "--no-cpu-baseline")
    assert d["retrieval_correct"] is True and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["config"]["D"] == 4096 and d["config"]["index_ciphertexts"] == 5 and d["config"]["payload_ciphertexts"] == 28
    assert d["roofline"]["bound"] == "hbm" and 0 < d["roofline"]["frac"] < 1 and d["gpu_launches"] > 0
    assert all(v > 0 for k, v in d["columns"].items() if "time" in k)
