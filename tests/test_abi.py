"""CPU-side checks of the drop-in boundary: the product library loads, exports every symbol that
include/khronos_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import khronos_b200 as kb
from khronos_b200 import capi
from harness import ROOT, has_gpu


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "khronos_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|const char\*|uint64_t)\s+(kb_\w+)\s*\(", src, flags=re.M)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("kb_create", "kb_destroy", "kb_integrate_frame", "kb_update_tracking", "kb_detect_motion",
                 "kb_reset_inactive", "kb_scan_object_confidence", "kb_export_blocks", "kb_set_shard"):
        assert must in syms


def test_library_exports_every_declared_symbol(product_lib):
    for s in declared_symbols():
        assert hasattr(product_lib, s), f"libkhronos_b200.so does not export {s}"
    assert product_lib.kb_abi_version() == 8


def test_oracle_mirrors_abi(oracle_lib):
    for s in declared_symbols():
        if s == "kb_host_cluster_motion":
            continue  # product-internal host path, checked against the oracle in test_host_logic.py
        if s.startswith("kb_peer_") or s.startswith("kb_gather_"):
            continue  # CUDA IPC / NVLink plumbing: device-memory transport, nothing for a CPU oracle to restate
        assert hasattr(oracle_lib, "ko_" + s[3:]), s


def test_struct_sizes_match_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof() with the ctypes mirrors."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "khronos_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(kb_map_config),sizeof(kb_integrator_config),sizeof(kb_tracking_config),sizeof(kb_motion_config),'
                   'sizeof(kb_camera),sizeof(kb_frame),sizeof(kb_frame_stats),sizeof(kb_block_export),sizeof(kb_totals64));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(c) for c in (capi.MapConfig, capi.IntegratorConfig, capi.TrackingConfig, capi.MotionConfig,
                                       capi.Camera, capi.Frame, capi.FrameStats, capi.BlockExport, capi.Totals64)]
    assert got == want


def test_block_owner_is_balanced(product_lib):
    import numpy as np
    counts = np.zeros(8, int)
    for x in range(-10, 10):
        for y in range(-10, 10):
            for z in range(-3, 3):
                counts[product_lib.kb_block_owner(x, y, z, 8)] += 1
    assert counts.min() > 0.8 * counts.mean() and counts.max() < 1.2 * counts.mean()
    assert product_lib.kb_block_owner(1, 2, 3, 1) == 0


@pytest.mark.skipif(has_gpu(), reason="CPU-only check")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(kb.KbError) as e:
        kb.create_map(kb.default_map_config(), kb.default_integrator_config())
    assert e.value.status == capi.KB_ERR_NO_DEVICE
