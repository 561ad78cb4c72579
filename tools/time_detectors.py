"""Host-side timing of the detector / tracker entry points around the fusion path (diagnostic; run on a GPU box):
kb_detect_objects (3D and 2D) and kb_track_measurements on a 640x480 frame resident on the device. These calls are
synchronous (they return counts), so perf_counter around the call is the latency a caller sees."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import khronos_b200 as kb
from khronos_b200 import capi
from test_object_detection_oracle import OBJECTS, scene_frame


def main(iters=50):
    cam, pose, d, l = scene_frame(scale=1, noise_seed=3)
    h = kb.create_map(capi.default_map_config(max_blocks=4096), capi.default_integrator_config(), capi.default_tracking_config(), None)
    h.set_camera(cam)
    dd, ll = torch.from_numpy(d).cuda(), torch.from_numpy(l).cuda()
    torch.cuda.synchronize()
    f = h.make_frame(dd.data_ptr(), pose, 1_000_000_000, label=ll.data_ptr(), memory=capi.MEM_DEVICE)
    out = {}
    for name, use_3d in (("detect_objects_3d", True), ("detect_objects_2d", False)):
        cfg = capi.default_object_detector_config(OBJECTS, use_3d=use_3d, min_cluster_size=50)
        for _ in range(5):
            ids, n = h.detect_objects(cfg, f)
        t0 = time.perf_counter()
        for _ in range(iters):
            ids, n = h.detect_objects(cfg, f)
        out[name] = {"us": round((time.perf_counter() - t0) / iters * 1e6, 1), "clusters": n}
    # clusters for the tracker step: the 3D detector's (the shipped mode, ids 1..n)
    cfg = capi.default_object_detector_config(OBJECTS, use_3d=True, min_cluster_size=50)
    ids, n = h.detect_objects(cfg, f)
    cids = [c["id"] for c in h.get_object_clusters()]
    max_id = len(cids)
    ii = torch.from_numpy(ids).cuda()
    torch.cuda.synchronize()
    h.track_measurements(f, ii.data_ptr(), cids, 0.1, [])
    lists = [v for v in h.get_cluster_voxels(max_id) if len(v)]
    tracks = lists[:16]
    for name, tr in (("track_measurements_no_tracks", []), (f"track_measurements_{len(tracks)}_tracks", tracks)):
        for _ in range(5):
            r = h.track_measurements(f, ii.data_ptr(), cids, 0.1, tr)
        t0 = time.perf_counter()
        for _ in range(iters):
            r = h.track_measurements(f, ii.data_ptr(), cids, 0.1, tr)
        out[name] = {"us": round((time.perf_counter() - t0) / iters * 1e6, 1), "clusters": max_id,
                     "cluster_voxels": int(r["voxel_counts"].sum()), "track_voxels": int(sum(len(t) for t in tr))}
    print(out)
    time_rays(scale=1 if iters >= 10 else 0.01)


def time_rays(scale=1.0):
    """Ray index (kb_rays_*): rays hashed per second and points checked per second on a synthetic mesh: vertices on the
    walls of a 40 x 30 x 4 m hall, 200 pose nodes along a sweep, ray policy kMiddle, 1 m blocks."""
    rng = np.random.default_rng(0)
    n_v, n_p, n_q = int(400_000 * scale), 200, int(200_000 * scale)
    stamps = (np.uint64(1_000_000_000) + np.arange(n_p, dtype=np.uint64) * np.uint64(100_000_000))
    poses = np.stack([np.linspace(5, 35, n_p), 15 + 8 * np.sin(np.linspace(0, 6, n_p)), np.full(n_p, 1.5)], 1).astype(np.float32)
    verts = rng.uniform([0, 0, 0], [40, 30, 4], (n_v, 3)).astype(np.float32)
    wall = rng.integers(0, 3, n_v)
    for a, hi in enumerate((40.0, 30.0, 4.0)):
        verts[wall == a, a] = np.where(rng.integers(0, 2, int((wall == a).sum())) == 1, hi, 0.0)
    first = rng.integers(1_000_000_000, 20_000_000_000, n_v).astype(np.uint64)
    last = first + rng.integers(0, 2_000_000_000, n_v).astype(np.uint64)
    r = capi.RayIndex(kb.lib(), "kb_", capi.default_ray_config())
    t0 = time.perf_counter()
    _, n_rays = r.add_vertices(capi.RAYS_MIDDLE, stamps, poses, verts, first, last)
    t_add = time.perf_counter() - t0
    pts = (verts[rng.integers(0, n_v, n_q)] + rng.normal(0, 0.05, (n_q, 3))).astype(np.float32)
    r.check(pts[:1000])                       # builds the CSR
    t0 = time.perf_counter()
    counts, _ = r.check(pts)
    t_chk = time.perf_counter() - t0
    print({"rays": n_rays, "block_entries": r.size()[1], "add_s": round(t_add, 4), "rays_per_s": round(n_rays / t_add),
           "points": n_q, "check_s": round(t_chk, 4), "points_per_s": round(n_q / t_chk),
           "verdicts": int(counts.sum()), "note": "host-side call latency incl. H2D/D2H and the Python unpacking of the stamp lists"})


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 50)
