"""bench.py's reference arm and the C header, on CPU: the arm prints exactly one JSON line with the contract's keys
(the GPU arm is exercised on the GPU box by the driver), and include/fplgpu.h is a plain C header."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "fastplong_ref")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/fastplong_ref not built")
def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, FPL_BENCH_REF_READS="300")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1000:]
    lines = [ln for ln in r.stdout.split("\n") if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gbases/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert cb["phase_s"] <= cb["wall_s"]
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert "workload" in d["config"]


@pytest.mark.parametrize("compiler,flags", [("gcc", ["-std=c99", "-x", "c"]), ("g++", ["-std=c++14", "-x", "c++"])])
def test_header_is_plain_c(compiler, flags, tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "fplgpu.h"\nint main(void) { return (int)sizeof(fpl_options) + (int)sizeof(fpl_read_result) == 0; }\n')
    r = subprocess.run([compiler, *flags, "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.parametrize("name", ["r02_bench.json", "r02u_bench_full_check_65k_reads.json"])
def test_committed_bench_lines_carry_the_contract_keys(name):
    """The lines under profiles/ are what bench.py printed on a B200: the keys the driver and the judge read are all there
    and consistent with each other (the roofline fraction follows from achieved / peak, the value from bases / time)."""
    d = json.load(open(os.path.join(ROOT, "profiles", name)))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "gpu_launches", "roofline", "kernels", "parity_checked"):
        assert k in d, k
    assert d["unit"] == "Gbases/s" and d["dtype"] == "u8" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert abs(r["achieved"] - r["alg_bytes_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e9) / r["achieved"] < 1e-3
    bases = d["config"]["bases_per_gpu"] * d["n_gpus"]
    assert abs(d["value"] - bases / (d["ms_per_step"] * 1e-3) / 1e9) / d["value"] < 1e-3
    assert d["gpu_launches"] > 0 and d["parity_checked"] is True
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d.get("e2e"):
        for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
            assert k in d["e2e"], k
        assert d["e2e"]["h2d_bytes_per_step"] >= 2 * d["config"]["bases_per_gpu"] and d["e2e"]["value"] < d["value"]
    if d.get("cpu_baseline"):
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in d["cpu_baseline"], k
    if d.get("parity_full_scale"):
        assert d["parity_full_scale"]["ok"] is True
