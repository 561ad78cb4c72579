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


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 50)
