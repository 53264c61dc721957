"""The C-ABI boundary without a GPU: the library builds for sm_100a, loads, exports every symbol include/clive_b200.h
declares, and refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from clive2_b200 import build as clv_build
from clive2_b200 import runtime

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(REPO, "include", "clive_b200.h")).read()
    return sorted(set(re.findall(r"CLV_API\s+[\w\s\*]+?\b(clv_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    clv_build.build()
    return runtime.load_library()


def test_header_lists_the_eight_reference_kernels():
    names = declared_functions()
    for k in ("generate_light_rays", "generate_camera_rays", "generate_paths", "reset_light_indices", "connect_paths",
              "adaptive_finalize_samples", "light_sort", "light_image_gather"):
        assert "clv_" + k in names
    assert len(names) >= 35


def test_library_exports_every_declared_symbol(lib):
    raw = C.CDLL(clv_build.LIB)
    for name in declared_functions():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"


def test_python_binding_covers_the_header(lib):
    assert sorted(runtime.SIGNATURES) == declared_functions()


def test_kernel_function_names_match_reference():
    # renderer.py:30-37 looks these up by name
    assert set(runtime.KernelLibrary.NAMES) == set(runtime._DISPATCH)


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.clv_version()
    assert isinstance(lib.clv_last_error(), bytes)


def test_no_device_is_a_loud_error(lib):
    if runtime.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(runtime.CliveError) as e:
        runtime.Device()
    assert "no CPU path" in str(e.value) or "no CUDA device" in str(e.value)


def test_missing_library_is_a_loud_error(tmp_path):
    with pytest.raises(runtime.CliveError) as e:
        runtime.load_library(str(tmp_path / "libclive_b200.so"))
    assert "no CPU execution path" in str(e.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(REPO, "clive2_b200")
    for root, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", text, re.M), f
                assert "trace_ref" not in text or f.endswith((".cuh", ".h", ".cu")), f


def test_reference_arm_json_contract():
    """`bench.py --impl reference` runs on CPU (the oracle on all host threads): check the JSON line's schema."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--config", "c1",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "bdpt_msamples_per_s" and line["unit"] == "Msamples/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["value"] > 0
    from oracle import oracle as oracle_mod
    # "reference" = oracle/_ref (the reference's own trace.metal compiled as C++) when it is built, else the restatement
    assert line["cpu_baseline"]["kind"] == ("reference" if oracle_mod.have_ref() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and line["config"]["samples_per_step"] == 256 * 256
    assert line["e2e"] == {"value": line["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]
