"""GPU parity of kb_detect_objects (khronos::ConnectedSemantics on the device) against the CPU
oracle, which tests/test_object_detection_oracle.py pins against independent numpy / scipy restatements of the
in-tree reference code. Object images and cluster lists must be identical."""
import numpy as np
import pytest

from khronos_b200 import capi
import harness as hs
from test_object_detection_oracle import OBJECTS, scene_frame

pytestmark = pytest.mark.gpu


def check(o, g, cfg, fo, fg, what):
    io, no = o.detect_objects(cfg, fo)
    ig, ng = g.detect_objects(cfg, fg)
    assert no == ng, f"{what}: {no} vs {ng} clusters"
    np.testing.assert_array_equal(io, ig, err_msg=what)
    co, cg = o.get_object_clusters(), g.get_object_clusters()
    assert [(c["id"], c["semantic_id"], len(c["pixels"])) for c in co] == [(c["id"], c["semantic_id"], len(c["pixels"])) for c in cg]
    for a, b in zip(co, cg):
        pa, pb = a["pixels"][np.lexsort(a["pixels"].T)], b["pixels"][np.lexsort(b["pixels"].T)]
        np.testing.assert_array_equal(pa, pb, err_msg=f"{what} pixels of cluster {a['id']}")
    return no


@pytest.mark.parametrize("use_3d", [True, False])
@pytest.mark.parametrize("full,min_size,max_size,max_range", [(True, 0, -1, 0.0), (False, 0, -1, 0.0), (True, 15, 4000, 4.0)])
def test_object_detection_matches_oracle(oracle_lib, product_lib, use_3d, full, min_size, max_size, max_range):
    cam, pose, d, l = scene_frame(scale=2, noise_seed=11)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    g = hs.make_handle(product_lib, "kb_", cam=cam)
    cfg = capi.default_object_detector_config(OBJECTS, use_3d=use_3d, use_full_connectivity=full, min_cluster_size=min_size,
                                              max_cluster_size=max_size, max_range=max_range)
    n = check(o, g, cfg, o.make_frame(d, pose, 1_000_000_000, label=l), g.make_frame(d, pose, 1_000_000_000, label=l),
              f"3d={use_3d} full={full} min={min_size}")
    assert n >= 1


def test_object_detection_full_resolution_device_frames_and_vertex_map(oracle_lib, product_lib):
    """640x480 frame resident on the device; caller-supplied world-frame vertex map; repeated calls reuse the table."""
    import torch
    cam, pose, d, l = scene_frame(scale=1, noise_seed=3)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    g = hs.make_handle(product_lib, "kb_", cam=cam)
    dd, ll = torch.from_numpy(d).cuda(), torch.from_numpy(l).cuda()
    torch.cuda.synchronize()
    for use_3d in (True, False, True):
        cfg = capi.default_object_detector_config(OBJECTS, use_3d=use_3d, min_cluster_size=50)
        check(o, g, cfg, o.make_frame(d, pose, 1_000_000_000, label=l),
              g.make_frame(dd, pose, 1_000_000_000, label=ll, memory=capi.MEM_DEVICE), f"fullres 3d={use_3d}")
    # vertex map supplied by the caller (shifted by one grid cell in x: the clusters must follow it)
    H, W = l.shape
    v, u = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    T = np.asarray(pose, np.float64)
    pc = np.stack([(u - np.float32(cam.cx)) / np.float32(cam.fx) * d, (v - np.float32(cam.cy)) / np.float32(cam.fy) * d, d], -1)
    vw = (pc.astype(np.float64) @ T[:3, :3].T + T[:3, 3] + np.array([0.1, 0.0, 0.0])).astype(np.float32)
    vw = np.ascontiguousarray(vw)
    cfg = capi.default_object_detector_config(OBJECTS, use_3d=True, min_cluster_size=50)
    check(o, g, cfg, o.make_frame(d, pose, 1_000_000_000, label=l, vertex_world=vw),
          g.make_frame(d, pose, 1_000_000_000, label=l, vertex_world=vw), "vertex map")


def test_object_detection_does_not_disturb_motion_detection(oracle_lib, product_lib):
    """The detector borrows the motion detector's table memory: interleaving the two (as ActiveWindow::spinOnce does,
    active_window.cpp:127-130) must not change either result."""
    import test_sharded_pipeline as tsp
    cam = hs.small_camera(4)
    frames, poses, stamps = tsp.dynamic_scenario(cam, 26)
    mot = capi.default_motion_config(min_cluster_size=5, min_separation_distance=2.0)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam, mot_cfg=mot)
    g = hs.make_handle(product_lib, "kb_", cam=cam, mot_cfg=mot)
    cfg = capi.default_object_detector_config(OBJECTS, use_3d=True, min_cluster_size=10)
    dyn = 0
    for i, ((d, l), T, st) in enumerate(zip(frames, poses, stamps)):
        io, so, co = o.detect_motion(o.make_frame(d, T, st, label=l))
        ig, sg, cg = g.detect_motion(g.make_frame(d, T, st, label=l))
        check(o, g, cfg, o.make_frame(d, T, st, label=l), g.make_frame(d, T, st, label=l), f"frame {i}")
        assert (so, co) == (sg, cg)
        np.testing.assert_array_equal(io, ig)
        assert len(o.get_motion_clusters()) == len(g.get_motion_clusters())
        dyn += int((io > 0).sum())
        for h, img in ((o, io), (g, ig)):
            h.integrate_frame(h.make_frame(d, T, st, label=l, mask=img))
            h.update_tracking(st)
    assert dyn > 100
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="interleaved")


def test_adaptor_object_detector(tmp_path):
    """GpuConnectedSemantics through the C++ host adaptor: a wall of one object class is one cluster of all pixels."""
    import os
    import subprocess
    from harness import ROOT
    csrc = os.path.join(ROOT, "khronos_b200", "csrc")
    exe = str(tmp_path / "adaptor_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "adaptor_compile_check.cpp"),
                           "-o", exe, "-L", csrc, "-lkhronos_b200", f"-Wl,-rpath,{csrc}"])
    out = subprocess.run([exe, "require-gpu", "objects", "core"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    fields = dict(kv.split("=") for kv in out.stdout.split())
    assert int(fields["semantic_clusters"]) == 1 and int(fields["cluster_pixels"]) == 64 * 48
    # GpuActiveWindowCore: 8 frames at 10 Hz with 0.25 s output separation -> outputs at frames 0, 3, 6; finishMapping archives
    # every block (116 = the blocks of the known-answer frame) and empties the mirrored host map
    assert (int(fields["core_frames"]), int(fields["core_outputs"])) == (8, 3)
    assert int(fields["core_blocks_after_finish"]) == 0 and int(fields["core_archived"]) == int(fields["blocks"]) > 20
