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
