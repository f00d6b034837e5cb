"""The g2o-facing adapter (sparse-gslam_b200/adapter/sgb_g2o_adapter.h): compile check on CPU, end-to-end run on GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "sparse-gslam_b200", "adapter")


def _build():
    from sparse_gslam_b200 import build
    build.build()
    exe = os.path.join(ADAPTER, "example_graphs")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", os.path.join(ADAPTER, "example_graphs.cpp"),
                           "-I" + os.path.join(ROOT, "include"), "-L" + os.path.join(ROOT, "sparse-gslam_b200"), "-lsgb",
                           "-Wl,-rpath," + os.path.join(ROOT, "sparse-gslam_b200"), "-o", exe])
    return exe


def test_adapter_compiles_and_fails_loudly_without_gpu():
    exe = _build()
    from sparse_gslam_b200 import capi
    if capi.load().sgb_device_count() > 0:
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0                      # no silent CPU fallback
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_adapter_runs_reference_call_sequence():
    """graphs.cpp-style set-up + drone.cpp:146-165 call sequence through the g2o plugin surface."""
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "optimize returned" in r.stdout
