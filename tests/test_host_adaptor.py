"""The C++ host adaptor (khronos_b200/host/khronos_gpu_adaptor.h) mirrors the reference's integrator /
detector interfaces over the C ABI. CPU: it compiles and links against the stub Hydra types and the
product library and reports the missing GPU cleanly. GPU: one frame through the C++ classes matches
the hand-computed known answer (same case as tests/test_oracle_kat.py::test_flat_wall_known_answer)."""
import os
import subprocess

import pytest

from harness import ROOT, has_gpu

CSRC = os.path.join(ROOT, "khronos_b200", "csrc")


@pytest.fixture(scope="module")
def adaptor_exe(tmp_path_factory, product_lib):
    exe = str(tmp_path_factory.mktemp("adaptor") / "adaptor_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "adaptor_compile_check.cpp"),
                           "-o", exe, "-L", CSRC, "-lkhronos_b200", f"-Wl,-rpath,{CSRC}"])
    return exe


@pytest.mark.skipif(has_gpu(), reason="CPU-only check")
def test_adaptor_compiles_links_and_refuses_without_gpu(adaptor_exe):
    out = subprocess.run([adaptor_exe], capture_output=True, text=True)
    assert out.returncode == 0 and "no-gpu" in out.stdout
    assert subprocess.run([adaptor_exe, "require-gpu"], capture_output=True).returncode == 1


@pytest.mark.gpu
def test_adaptor_frame_matches_known_answer(adaptor_exe):
    out = subprocess.run([adaptor_exe, "require-gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    fields = dict(kv.split("=") for kv in out.stdout.split())
    assert int(fields["blocks"]) > 20
    assert float(fields["distance"]) == pytest.approx(0.05, abs=1e-6)
    assert float(fields["weight"]) == pytest.approx(32 * 32 * 0.01 / 1.95 ** 4, rel=1e-5)
    assert int(fields["label"]) == 3
    assert int(fields["timers"]) == 3   # motion_detection/all, active_window/update_map, integration/tracking were opened
    # GpuMeshIntegrator: the wall's mesh lies on z = 2, and the second tick finds no mesh_updated block
    assert int(fields["mesh_vertices"]) > 300 and int(fields["mesh_blocks"]) > 5 and float(fields["mesh_err"]) < 2e-3
    assert int(fields["mesh_again"]) == 0
    # reconstructStaticObject: dense box allocated, 12 frames fused, some low-confidence voxels erased
    assert int(fields["object_blocks"]) > 100 and int(fields["erased"]) >= 0
