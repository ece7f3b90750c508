"""CPU: the reference arm of bench.py (`--impl reference`: the CPU restatement of the rasterizer timed on the host cores)
prints ONE JSON line with the keys the driver reads; the product arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rendered views/sec @640x480" and d["unit"] == "views/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"].startswith("scannet_2views_640x480")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_both_arms_print_the_same_config_and_bench_does_not_import_tests():
    """The driver compares the two arms' `config` objects: both come from bench.base_config.  bench.py may execute oracle/
    (cpu_baseline / reference arm) but must not import model or helper code from tests/."""
    import importlib.util
    import re
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.base_config(1)["workload"] == b.WORKLOAD and set(b.base_config(1)) == set(b.base_config(8))
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("base_config(") >= 3                    # definition + the two arms
    assert not re.search(r"^\s*(from|import)\s+tests\b", src, re.M) and "sys.path.insert(0, os.path.join(ROOT, \"tests\"))" not in src
