"""CPU: the oracle's restatement of khronos::MaxIoUTracker's voxel measurements (track_by = voxels:
setupTrackMeasurementVoxels max_iou_tracker.cpp:450-459, computeCentroid :534-539, computeIoUVoxels :551-562) against
an independent numpy restatement and hand-computed cases. The reference has no tests for the tracker."""
import numpy as np

from khronos_b200 import capi
import harness as hs
from test_object_detection_oracle import OBJECTS, scene_frame


def numpy_measurements(cam, pose, depth, ids, clusters, voxel_size, tracks, vertex=None):
    """Python sets, as the reference's GlobalIndexSet; fp32 arithmetic in the oracle's order. clusters: n (pixel values
    1..n) or the list of the clusters' pixel values."""
    values = list(range(1, clusters + 1)) if isinstance(clusters, (int, np.integer)) else [int(c) for c in clusters]
    row = {v: i for i, v in enumerate(values)}
    max_id = len(values)
    H, W = ids.shape
    f32 = np.float32
    if vertex is None:
        T = np.asarray(pose, np.float64)
        R, t = T[:3, :3].astype(f32), T[:3, 3].astype(f32)
        v, u = np.meshgrid(np.arange(H, dtype=f32), np.arange(W, dtype=f32), indexing="ij")
        x = (u - f32(cam.cx)) / f32(cam.fx) * depth
        y = (v - f32(cam.cy)) / f32(cam.fy) * depth
        vertex = np.stack([((R[a, 0] * x + R[a, 1] * y) + R[a, 2] * depth) + t[a] for a in range(3)], -1).astype(f32)
    inv = f32(1.0) / f32(voxel_size)
    g = np.floor(vertex * inv).astype(np.int64)
    sets = [set() for _ in range(max_id)]
    for vv, uu in zip(*np.nonzero(np.isin(ids, values))):
        sets[row[int(ids[vv, uu])]].add(tuple(g[vv, uu]))
    counts = np.array([len(s) for s in sets], np.int32)
    sums = np.array([np.sum(np.array(sorted(s), np.int64).reshape(-1, 3), 0) for s in sets], np.int64)
    inter = np.zeros((max_id, len(tracks)), np.int32)
    iou = np.zeros((max_id, len(tracks)), f32)
    for j, tr in enumerate(tracks):
        ts = set(map(tuple, np.asarray(tr, np.int64).reshape(-1, 3)))
        for i, s in enumerate(sets):
            inter[i, j] = len(s & ts)
            with np.errstate(invalid="ignore", divide="ignore"):
                iou[i, j] = f32(inter[i, j]) / (f32(len(s) + len(tr)) - f32(inter[i, j]))
    lists = [np.array(sorted(s, key=lambda p: (p[2], p[1], p[0])), np.int64).reshape(-1, 3) for s in sets]
    return counts, sums, inter, iou, lists


def compare(res, lists, want):
    counts, sums, inter, iou, wl = want
    np.testing.assert_array_equal(res["voxel_counts"], counts)
    np.testing.assert_array_equal(res["voxel_sums"], sums)
    np.testing.assert_array_equal(res["intersections"], inter)
    np.testing.assert_array_equal(res["iou"].view(np.uint32), iou.view(np.uint32))
    assert len(lists) == len(wl)
    for a, b in zip(lists, wl):
        np.testing.assert_array_equal(a, b)


def object_ids(h, d, l, pose, use_3d=True, min_size=10):
    cfg = capi.default_object_detector_config(OBJECTS, use_3d=use_3d, min_cluster_size=min_size)
    img, n = h.detect_objects(cfg, h.make_frame(d, pose, 1_000_000_000, label=l))
    return img, n


def test_hand_computed_single_voxel_cases(oracle_lib):
    """Camera at the origin looking along +z (identity pose): pixel (u, v) at depth z sits at ((u-cx)/fx*z, (v-cy)/fy*z, z)."""
    from khronos_b200 import synthetic as syn
    cam = syn.make_camera(8, 6, 4.0, 4.0, max_range=10.0)  # cx = 3.5, cy = 2.5
    h = hs.make_handle(oracle_lib, "ko_", cam=cam)
    d = np.full((6, 8), 2.0, np.float32)
    ids = np.zeros((6, 8), np.int32)
    # cluster 1: pixels (u=4,v=3) and (u=5,v=3): x = 0.25, 0.75; y = 0.25; z = 2.0 -> voxels (0,0,2) and (1,0,2) at 0.5 m...
    ids[3, 4] = 1
    ids[3, 5] = 1
    # cluster 3: one pixel (u=2, v=1): x = -0.75, y = -0.75 -> voxel (-2,-2,4) at 0.5 m
    ids[1, 2] = 3
    ids[0, 0] = 7       # beyond max_id: ignored
    ids[5, 7] = -2      # negative: ignored
    f = h.make_frame(d, np.eye(4), 1, label=None)
    tracks = [np.array([[0, 0, 4], [9, 9, 9]]), np.zeros((0, 3), np.int64), np.array([[-2, -2, 4]])]
    r = h.track_measurements(f, ids, 3, 0.5, tracks)
    np.testing.assert_array_equal(r["voxel_counts"], [2, 0, 1])
    np.testing.assert_array_equal(r["voxel_sums"], [[1, 0, 8], [0, 0, 0], [-2, -2, 4]])
    np.testing.assert_array_equal(r["intersections"], [[1, 0, 0], [0, 0, 0], [0, 0, 1]])
    iou = r["iou"]
    assert iou[0, 0] == np.float32(1.0) / np.float32(3.0)   # 1 / (2 + 2 - 1)
    assert iou[0, 1] == 0.0 and iou[2, 2] == 1.0
    assert np.isnan(iou[1, 1])                               # empty cluster vs empty track: 0 / 0 as in the reference
    assert iou[1, 0] == 0.0
    vox = h.get_cluster_voxels(3)
    np.testing.assert_array_equal(vox[0], [[0, 0, 4], [1, 0, 4]])
    assert len(vox[1]) == 0
    np.testing.assert_array_equal(vox[2], [[-2, -2, 4]])
    # centroid of cluster 1 (computeCentroid, voxel mode): mean voxel centre
    c = (r["voxel_sums"][0] / r["voxel_counts"][0] + 0.5) * 0.5
    np.testing.assert_allclose(c, [0.5, 0.25, 2.25])


def test_oracle_matches_numpy_on_object_clusters(oracle_lib):
    cam, pose, d, l = scene_frame(scale=2, noise_seed=5)
    h = hs.make_handle(oracle_lib, "ko_", cam=cam)
    ids, n = object_ids(h, d, l, pose)
    assert n >= 3
    rng = np.random.default_rng(3)
    f = h.make_frame(d, pose, 1_000_000_000, label=l)
    base = h.track_measurements(f, ids, n, 0.1, [])
    lists = h.get_cluster_voxels(n)
    assert base["intersections"].shape == (n, 0)
    # tracks: a cluster's own voxels (IoU 1), a shifted copy, half of one, a mix of two clusters, random voxels, empty
    tracks = [lists[0], lists[1] + np.array([1, 0, 0]), lists[2][::2], np.concatenate([lists[0][:5], lists[1][:7]]),
              rng.integers(-50, 50, (40, 3)), np.zeros((0, 3), np.int64)]
    tracks[4] = np.unique(tracks[4], axis=0)
    for vs in (0.1, 0.07, 0.25):
        r = h.track_measurements(f, ids, n, vs, tracks)
        compare(r, h.get_cluster_voxels(n), numpy_measurements(cam, pose, d, ids, n, vs, tracks))
    r = h.track_measurements(f, ids, n, 0.1, tracks)
    assert r["iou"][0, 0] == 1.0 and r["intersections"][0, 0] == len(lists[0])
    assert r["intersections"][2, 2] == len(lists[2][::2])


def test_oracle_sparse_cluster_ids_of_the_2d_detector(oracle_lib):
    """The 2D detector keeps creation-order ids (filterClusters does not renumber), so the surviving ids are sparse and
    can be large: the caller passes them as a list and gets one row per list entry."""
    cam, pose, d, l = scene_frame(scale=2, noise_seed=9)
    h = hs.make_handle(oracle_lib, "ko_", cam=cam)
    ids, n = object_ids(h, d, l, pose, use_3d=False, min_size=2)
    cids = [c["id"] for c in h.get_object_clusters()]
    assert len(cids) == n >= 2 and cids == sorted(cids) and cids[-1] > 1022   # sparse, beyond the dense id range
    f = h.make_frame(d, pose, 1, label=l)
    h.track_measurements(f, ids, cids, 0.1, [])
    lists = h.get_cluster_voxels(n)
    tracks = [lists[0], lists[-1][::3], np.array([[0, 0, 0]])]
    r = h.track_measurements(f, ids, cids, 0.1, tracks)
    compare(r, h.get_cluster_voxels(n), numpy_measurements(cam, pose, d, ids, cids, 0.1, tracks))
    assert r["iou"][0, 0] == 1.0 and r["voxel_counts"].all()
    # a sub-list: the other clusters' pixels belong to no row
    sub = cids[1::2]
    r = h.track_measurements(f, ids, sub, 0.1, tracks)
    compare(r, h.get_cluster_voxels(len(sub)), numpy_measurements(cam, pose, d, ids, sub, 0.1, tracks))
    import pytest
    with pytest.raises(capi.KbError):
        h.track_measurements(f, ids, cids[::-1], 0.1, tracks)   # not ascending


def test_oracle_vertex_map_and_2d_ids(oracle_lib):
    cam, pose, d, l = scene_frame(scale=2, noise_seed=9)
    h = hs.make_handle(oracle_lib, "ko_", cam=cam)
    ids, n = object_ids(h, d, l, pose, use_3d=False)
    max_id = int(ids.max())
    assert max_id >= n >= 1
    rng = np.random.default_rng(1)
    vw = rng.uniform(-3, 3, ids.shape + (3,)).astype(np.float32)
    tracks = [np.unique(rng.integers(-30, 30, (500, 3)), axis=0)]
    f = h.make_frame(d, pose, 1, label=l, vertex_world=vw)
    r = h.track_measurements(f, ids, max_id, 0.1, tracks)
    compare(r, h.get_cluster_voxels(max_id), numpy_measurements(cam, pose, d, ids, max_id, 0.1, tracks, vertex=vw))


def test_oracle_vertex_map_matches_numpy(oracle_lib):
    """parseInputPacket's world-frame vertex map: p_W = R * ((u-cx)/fx*d, (v-cy)/fy*d, d) + t in fp32 (ORACLE_SPEC §8)."""
    cam, pose, d, l = scene_frame(scale=4, noise_seed=2)
    h = hs.make_handle(oracle_lib, "ko_", cam=cam)
    got = h.compute_vertex_map(h.make_frame(d, pose, 1, label=l))
    f32 = np.float32
    T = np.asarray(pose, np.float64)
    R, t = T[:3, :3].astype(f32), T[:3, 3].astype(f32)
    v, u = np.meshgrid(np.arange(cam.height, dtype=f32), np.arange(cam.width, dtype=f32), indexing="ij")
    x, y = (u - f32(cam.cx)) / f32(cam.fx) * d, (v - f32(cam.cy)) / f32(cam.fy) * d
    want = np.stack([((R[a, 0] * x + R[a, 1] * y) + R[a, 2] * d) + t[a] for a in range(3)], -1).astype(f32)
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    # the map is what the detectors compute internally: feeding it back changes nothing
    ids, n = object_ids(h, d, l, pose)
    a = h.track_measurements(h.make_frame(d, pose, 1, label=l), ids, n, 0.1, [])
    b = h.track_measurements(h.make_frame(d, pose, 1, label=l, vertex_world=got), ids, n, 0.1, [])
    np.testing.assert_array_equal(a["voxel_sums"], b["voxel_sums"])
