"""GPU parity of kb_track_measurements / kb_get_cluster_voxels (khronos::MaxIoUTracker's voxel measurements on the device,
SURVEY.md §8f row 4) against the CPU oracle, which tests/test_track_measurements_oracle.py pins against an independent
numpy restatement of max_iou_tracker.cpp:450-459, :534-539, :551-562. Counts, index sums, intersections, the IoU floats
(bit pattern) and the voxel lists must be identical."""
import os

import numpy as np
import pytest

from khronos_b200 import capi
import harness as hs
from test_object_detection_oracle import OBJECTS, scene_frame

pytestmark = pytest.mark.gpu


def check(o, g, fo, fg, ids, max_id, voxel_size, tracks, what, ids_g=None):
    ro = o.track_measurements(fo, ids, max_id, voxel_size, tracks)
    rg = g.track_measurements(fg, ids if ids_g is None else ids_g, max_id, voxel_size, tracks)
    for k in ("voxel_counts", "voxel_sums", "intersections"):
        np.testing.assert_array_equal(ro[k], rg[k], err_msg=f"{what}: {k}")
    np.testing.assert_array_equal(ro["iou"].view(np.uint32), rg["iou"].view(np.uint32), err_msg=f"{what}: iou")
    rows = max_id if isinstance(max_id, (int, np.integer)) else len(max_id)
    lo, lg = o.get_cluster_voxels(rows), g.get_cluster_voxels(rows)
    for c, (a, b) in enumerate(zip(lo, lg)):
        np.testing.assert_array_equal(a, b, err_msg=f"{what}: voxels of cluster {c + 1}")
    return ro, lo


def make_tracks(lists, rng):
    full = [v for v in lists if len(v)]
    tracks = [full[0], full[-1] + np.array([1, 0, 0]), full[len(full) // 2][::2],
              np.unique(rng.integers(-60, 60, (300, 3)), axis=0), np.zeros((0, 3), np.int64)]
    if len(full) > 1:
        tracks.append(np.concatenate([full[0][:5], full[1][:7]]))
    return tracks


@pytest.mark.parametrize("use_3d", [True, False])
def test_track_measurements_match_oracle_on_object_clusters(oracle_lib, product_lib, use_3d):
    cam, pose, d, l = scene_frame(scale=2, noise_seed=5)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    g = hs.make_handle(product_lib, "kb_", cam=cam)
    # 2D mode keeps creation-order ids (sparse, here beyond 1022): those clusters are named by an id list
    cfg = capi.default_object_detector_config(OBJECTS, use_3d=use_3d, min_cluster_size=10 if use_3d else 2)
    ids, n = o.detect_objects(cfg, o.make_frame(d, pose, 1_000_000_000, label=l))
    ids_g, _ = g.detect_objects(cfg, g.make_frame(d, pose, 1_000_000_000, label=l))
    np.testing.assert_array_equal(ids, ids_g)
    cids = [c["id"] for c in o.get_object_clusters()]
    assert cids == [c["id"] for c in g.get_object_clusters()] and len(cids) == n >= 3
    assert (cids == list(range(1, n + 1))) == use_3d and (use_3d or cids[-1] > 1022)
    clusters = n if use_3d else cids
    fo, fg = o.make_frame(d, pose, 1_000_000_000, label=l), g.make_frame(d, pose, 1_000_000_000, label=l)
    _, lists = check(o, g, fo, fg, ids, clusters, 0.1, [], "no tracks")
    tracks = make_tracks(lists, np.random.default_rng(2))
    for vs in (0.1, 0.07, 0.25):
        r, _ = check(o, g, fo, fg, ids, clusters, vs, tracks, f"3d={use_3d} voxel_size={vs}")
    r, lists = check(o, g, fo, fg, ids, clusters, 0.1, tracks, "again")
    assert r["iou"][0, 0] == 1.0 and r["voxel_counts"].all()
    # fewer clusters than the image holds: the other pixels belong to no row
    check(o, g, fo, fg, ids, 2, 0.1, tracks, "ids 1..2")
    check(o, g, fo, fg, ids, cids[1::2], 0.07, tracks, "id sub-list")
    r1, _ = check(o, g, fo, fg, ids, cids, 0.1, tracks, "id list")
    np.testing.assert_array_equal(r1["iou"].view(np.uint32), r["iou"].view(np.uint32))


def test_track_measurements_full_resolution_device_frames_vertex_map_and_compact_depth(oracle_lib, product_lib):
    import torch
    cam, pose, d, l = scene_frame(scale=1, noise_seed=3)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    g = hs.make_handle(product_lib, "kb_", cam=cam)
    cfg = capi.default_object_detector_config(OBJECTS, use_3d=True, min_cluster_size=50)
    ids, n = o.detect_objects(cfg, o.make_frame(d, pose, 1_000_000_000, label=l))
    max_id = int(ids.max())
    rng = np.random.default_rng(7)
    fo = o.make_frame(d, pose, 1_000_000_000, label=l)
    base = o.track_measurements(fo, ids, max_id, 0.1, [])
    tracks = make_tracks(o.get_cluster_voxels(max_id), rng)
    # device-resident frame and id image
    dd, ii = torch.from_numpy(d).cuda(), torch.from_numpy(ids).cuda()
    torch.cuda.synchronize()
    check(o, g, fo, g.make_frame(dd, pose, 1_000_000_000, memory=capi.MEM_DEVICE), ids, max_id, 0.1, tracks, "device frame",
          ids_g=ii.data_ptr())
    # caller-supplied vertex map
    vw = rng.uniform(-4, 4, ids.shape + (3,)).astype(np.float32)
    check(o, g, o.make_frame(d, pose, 1, label=l, vertex_world=vw), g.make_frame(d, pose, 1, label=l, vertex_world=vw), ids,
          max_id, 0.1, tracks, "vertex map")
    # compact depth (u16 millimetres, expanded on the device like everywhere else)
    d16 = np.clip(np.round(d * 1000.0), 0, 65535).astype(np.uint16)
    check(o, g, o.make_frame(None, pose, 1, depth_u16=d16), g.make_frame(None, pose, 1, depth_u16=d16), ids, max_id, 0.1, tracks,
          "u16 depth")


@pytest.mark.parametrize("sparse", [False, True])
def test_tracking_chain_on_dynamic_clusters_interleaved_with_the_detectors(oracle_lib, product_lib, sparse):
    """The per-frame order of ActiveWindow::spinOnce (active_window.cpp:127-134): motion detection, object detection,
    tracker. Tracks are the clusters of the previous frames (Track::last_voxels); all three stages share table memory."""
    import test_sharded_pipeline as tsp
    if sparse:
        os.environ["KB_MOTION_SPARSE"] = "1"
    try:
        cam = hs.small_camera(4)
        frames, poses, stamps = tsp.dynamic_scenario(cam, 26)
        mot = capi.default_motion_config(min_cluster_size=5, min_separation_distance=2.0)
        o = hs.make_handle(oracle_lib, "ko_", cam=cam, mot_cfg=mot)
        g = hs.make_handle(product_lib, "kb_", cam=cam, mot_cfg=mot)
        cfg = capi.default_object_detector_config(OBJECTS, use_3d=True, min_cluster_size=10)
        tracks, matched = [], 0
        for i, ((d, l), T, st) in enumerate(zip(frames, poses, stamps)):
            io, so, co = o.spin_once(o.make_frame(d, T, st, label=l))
            ig, sg, cg = g.spin_once(g.make_frame(d, T, st, label=l))
            assert (so, co) == (sg, cg), f"frame {i}"
            np.testing.assert_array_equal(io, ig, err_msg=f"frame {i}")
            fo, fg = o.make_frame(d, T, st, label=l), g.make_frame(d, T, st, label=l)
            if co > 0:
                r, lists = check(o, g, fo, fg, io, co, 0.1, tracks[-4:], f"dynamic clusters, frame {i}")
                if tracks:
                    matched += int((r["iou"] > 0.0).any())
                tracks.extend(v for v in lists if len(v))
            oo, no = o.detect_objects(cfg, fo)
            og, ng = g.detect_objects(cfg, fg)
            np.testing.assert_array_equal(oo, og, err_msg=f"objects, frame {i}")
            if no > 0:
                check(o, g, fo, fg, oo, no, 0.1, tracks[-4:], f"object clusters, frame {i}")
        assert len(tracks) >= 3 and matched >= 2   # the mover overlaps its own previous observation
        hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="map after the chain")
    finally:
        os.environ.pop("KB_MOTION_SPARSE", None)


def test_track_measurement_argument_errors_and_stale_results(product_lib):
    cam, pose, d, l = scene_frame(scale=4, noise_seed=1)
    g = hs.make_handle(product_lib, "kb_", cam=cam)
    ids = np.zeros(d.shape, np.int32)
    f = g.make_frame(d, pose, 1, label=l)
    for bad in (0, 1023):
        with pytest.raises(capi.KbError):
            g.track_measurements(f, ids, bad, 0.1, [])
    with pytest.raises(capi.KbError):
        g.track_measurements(f, ids, 4, 0.0, [])
    with pytest.raises(capi.KbError):
        g.track_measurements(f, ids, [5, 3], 0.1, [])   # not ascending
    r = g.track_measurements(f, ids, 4, 0.1, [np.array([[1, 2, 3]])])   # no cluster pixel at all
    assert not r["voxel_counts"].any() and not r["intersections"].any()
    assert all(len(v) == 0 for v in g.get_cluster_voxels(4))
    ids[2:5, 3:9] = 2
    g.track_measurements(f, ids, 4, 0.1, [])
    g.detect_objects(capi.default_object_detector_config(OBJECTS), f)   # reuses the table
    with pytest.raises(capi.KbError):
        g.get_cluster_voxels(4)


def test_vertex_map_matches_oracle(oracle_lib, product_lib):
    """kb_compute_vertex_map (input conversion): bit-identical to the oracle for host, compact-depth and device frames."""
    import torch
    cam, pose, d, l = scene_frame(scale=1, noise_seed=4)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    g = hs.make_handle(product_lib, "kb_", cam=cam)
    want = o.compute_vertex_map(o.make_frame(d, pose, 1, label=l))
    got = g.compute_vertex_map(g.make_frame(d, pose, 1, label=l))
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    d16 = np.clip(np.round(d * 1000.0), 0, 65535).astype(np.uint16)
    np.testing.assert_array_equal(g.compute_vertex_map(g.make_frame(None, pose, 1, depth_u16=d16)).view(np.uint32),
                                  o.compute_vertex_map(o.make_frame(None, pose, 1, depth_u16=d16)).view(np.uint32))
    dd = torch.from_numpy(d).cuda()
    out = torch.zeros(d.shape + (3,), dtype=torch.float32).cuda()
    torch.cuda.synchronize()
    g.compute_vertex_map(g.make_frame(dd, pose, 1, memory=capi.MEM_DEVICE), out_ptr=out.data_ptr())
    np.testing.assert_array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
