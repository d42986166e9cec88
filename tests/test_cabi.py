"""CPU tests: the C-ABI library builds, loads, exports every symbol include/lfba.h declares, and refuses to compute
without a GPU (there is no CPU fallback). No compute calls are made when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from lifcal_b200 import api, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lfba_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built):
    L = api.load()
    names = _declared("lfba.h")
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), f"liblfba.so does not export {n}"
    S = capi.scene_lib()
    for n in _declared("lfba_scene.h"):
        assert hasattr(S, n), f"liblfba_scene.so does not export {n}"


def test_version_and_reference_defaults(built):
    assert api.version() == 100
    o = capi.Options()
    api.load().lfba_options_init(C.byref(o))
    # src/CameraCalibration.cpp:955-962 over Ceres 2.1.0 defaults
    assert o.max_num_iterations == 200
    assert o.function_tolerance == 1e-6 and o.parameter_tolerance == 1e-8 and o.gradient_tolerance == 1e-10
    assert o.initial_trust_region_radius == 1e4 and o.min_relative_decrease == 1e-3
    assert o.min_lm_diagonal == 1e-6 and o.max_lm_diagonal == 1e32
    assert o.loss_scale == 0.5 and o.minimizer_progress_to_stdout == 1


def test_struct_layouts_match_header(built):
    # sizes the C side was compiled with (guards the ctypes mirrors in lifcal_b200/capi.py)
    assert C.sizeof(capi.Iteration) == 4 * 4 + 9 * 8
    assert C.sizeof(capi.Comm) == 8 + 128 + 8
    assert C.sizeof(capi.ReprojStats) == 6 * 8
    assert capi.Problem.obs_x.offset == 48 and capi.Problem.n_constraints.offset == 96


def test_no_gpu_means_loud_failure_not_fallback(built):
    if api.device_count() > 0:
        pytest.skip("a GPU is present; the failure path is exercised on the CPU tier")
    sc = capi.make_scene(None, n_points=10, n_frames=2, seed=1)
    with pytest.raises(capi.LfbaError) as e:
        api.solve(sc.problem, sc.camera_init, sc.views_init, sc.points_init)
    assert "status 2" in str(e.value)  # LFBA_NO_DEVICE
    with pytest.raises(capi.LfbaError):
        api.evaluate(sc.problem, sc.camera_init, sc.views_init, sc.points_init)


def test_product_package_never_imports_the_oracle():
    # the oracle is test infrastructure: nothing under lifcal_b200/ may import, link, open or execute it
    pkg = os.path.join(ROOT, "lifcal_b200")
    banned = ("import oracle", "from oracle", "oracle/", "liblfba_oracle", "oracle.binding", "_ref/")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                for b in banned:
                    assert b not in src, f"{f} references {b}"


def test_bench_reference_arm_prints_contract_line(built):
    """`bench.py --impl reference` (the CPU oracle on the host cores) prints ONE JSON line with the contract keys; it
    needs no GPU. cfg1 = BASELINE.json configs[0], the reference's own CPU-runnable case."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                        "--steps", "1", "--warmup", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lm_residual_jacobian_evals_per_s" and d["unit"] == "M evals/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "M evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
