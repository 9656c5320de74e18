"""CPU tier: the C-ABI library loads and exports every symbol include/mof_b200.h declares, the command-line
host behaves like the reference's parser, and nothing silently falls back to a CPU solver."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import CLI_BIN, ROOT
from meshopticalflow_b200 import api


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mof_b200.h")).read()
    declared = set(re.findall(r"\b(mof_[a-z_0-9]+)\s*\(", header))
    lib = api.load_library()
    assert declared, "no declarations found in the header"
    assert declared == set(api.EXPORTED_SYMBOLS)
    for name in sorted(declared):
        assert hasattr(lib, name), name


def test_default_params_are_the_reference_defaults():
    p = api.default_params()
    assert p.iterations == 10
    assert p.sSmooth == float(np.float32(3e-3)) and p.sMultiply == 0.25
    assert p.vfSmooth == 3e-6 and p.vMultiply == 1.0 and p.vfSThreshold == float(np.float32(1e-8))
    assert p.dogWeight == 1.0 and p.dogSmooth == float(np.float32(1e-4))
    assert p.flowTol == 1e-8
    assert p.vfMode == api.VF_WHITNEY and p.cMode == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.MofError):
        api.Aligner(0)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "meshopticalflow_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.replace("no oracle", ""), os.path.join(dirpath, f)


def _run(*args):
    return subprocess.run([CLI_BIN, *args], capture_output=True, text=True, timeout=60)


@pytest.mark.skipif(not os.path.exists(CLI_BIN), reason="OpticalFlow host not built")
def test_cli_usage_and_flag_handling():
    r = _run()
    assert r.returncode == 1 and "Usage" in r.stdout and "--sSmooth" in r.stdout  # missing --in: usage + EXIT_FAILURE (OpticalFlow.cpp:1099-1103)
    r = _run("--IN", "a.ply", "b.ply", "--bogus", "--out", "x.ply", "--vfMode", "3")
    assert "[WARNING] Invalid option: --bogus" in r.stderr  # CmdLineParser.inl:253-256, names are case-insensitive (:247)
    assert "ERROR: Unsupported vector field!" in r.stdout and r.returncode == 0  # OpticalFlow.cpp:867-868: printf + return 0 from Init
    r = _run("--in", "a.ply", "b.ply", "--out", "x.ply", "--vfMode", "2", "--cMode", "7")
    assert "Undefined Connection Mode" in r.stdout
    r = _run("--in", "a.ply", "b.ply")
    assert "pass --out" in r.stderr and r.returncode != 0
    r = _run("--in", "/nonexistent/a.ply", "/nonexistent/b.ply", "--out", "x.ply")
    assert r.returncode != 0 and "Unable to read" in r.stderr


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` times the reference's CPU implementation (oracle/_ref, or the oracle port) on a
    bounded sample and prints one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--quick"], cwd=root, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "1M-vertex pair alignments/sec" and line["unit"] == "alignments/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "alignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["vertices"] == 1048578
