"""CPU: the JSON contract of bench.py - the reference arm is run here on a tiny workload, and the committed line of the
GPU arm (profiles/bench/r2_bench_1gpu.json, written by bench.py on a B200) is checked for the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COMMON = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"]


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--res", "32", "--batch", "2",
                          "--bins", "2", "--events", "50", "--channels", "8", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in COMMON + ["impl"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True and d["value"] > 0
    from oracle import ref_shim   # the real reference when a checkout / staged copy exists, else the oracle port
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_shim.reference_available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_committed_gpu_line_has_the_contract_keys():
    d = json.loads(open(os.path.join(ROOT, "profiles", "bench", "r2_bench_1gpu.json")).read().strip().splitlines()[-1])
    for k in COMMON + ["clocks", "roofline"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["gpu_launches"] > 0 and d["data"] == "synthetic" and d["dtype"] == "f32"
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["traffic"] is not None
    assert r["frac"] == r["frac_dram"] and r["frac_algorithmic"] is not None and r["kernel"] != "dp_allreduce"
    assert d["timed"]["total_ms"] >= 500 and d["gpu_launches_per_step"] > 0
    ev = d["eval"]   # the eval half of the metric carries its own roofline / e2e / CPU reference
    assert ev["roofline"]["frac"] > 0 and ev["e2e"]["h2d_bytes_per_step"] > 0 and ev["cpu_baseline"]["kind"] == "reference"
    assert d["eval_cfg0"]["value"] > 0 and d["eval_cfg0"]["cpu_baseline"]["kind"] == "reference"
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= 1.05 * d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["sample"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
