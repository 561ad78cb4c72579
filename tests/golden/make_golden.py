"""Generates the committed golden fixtures from the CPU oracle (run: python tests/golden/make_golden.py).

The reference has no tests / fixtures for this path and cannot be built or imported here (C++ with
Hydra/Eigen/OpenCV/ROS dependencies), so these vectors are produced by the oracle after it has been
pinned by the hand-computed known-answer tests (tests/test_oracle_kat.py). They freeze the oracle's
behaviour: any later change to oracle or product that alters results shows up as a golden diff.
Inputs are stored alongside outputs so the fixtures do not depend on the renderer's float behaviour.
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from khronos_b200 import capi, synthetic as syn  # noqa: E402
import harness as hs  # noqa: E402


def golden_camera():
    return syn.make_camera(80, 60, 40.0, 40.0, max_range=2.5)


def pack_blocks(b, prefix):
    ne = b.semantic_empty == 0
    out = {prefix + k: getattr(b, k) for k in ("block_index", "block_flags", "distance", "weight", "last_observed",
                                               "last_occupied", "ever_free", "active", "to_remove", "semantic_label",
                                               "semantic_empty")}
    out[prefix + "lik_values"] = b.semantic_likelihoods[ne]
    return out


def fusion_case(lib):
    cam = golden_camera()
    scene = syn.room_scene()
    poses, stamps = syn.orbit_trajectory(5, laps=0.05)
    frames = hs.render_frames(scene, cam, poses, stamps)
    h = hs.make_handle(lib, "ko_", cam=cam)
    stats = hs.run_fusion(h, frames, poses, stamps, tracking=True)
    out = {"depth": np.stack([f[0] for f in frames]), "label": np.stack([f[1] for f in frames]),
           "poses": np.stack(poses), "stamps": np.array(stamps, np.uint64),
           "stats": np.array([[s[k] for k in sorted(s)] for s in stats], np.int64)}
    out.update(pack_blocks(h.export_blocks(), "b_"))
    return out


def dynamic_case(lib):
    cam = golden_camera()
    scene = syn.room_scene()
    scene.mover = ((0.4, 0.4, 1.0), (7.6, 3.8, 0.9), (0.0, 1.2, 0.0), 1.6)
    n, dt = 18, 200_000_000
    pose = syn.look_pose((6.0, 5.0, 1.5), 0.0, np.radians(10.0))
    poses, stamps = [pose] * n, [1_000_000_000 + i * dt for i in range(n)]
    frames = hs.render_frames(scene, cam, poses, stamps)
    mot = capi.default_motion_config(min_cluster_size=4, min_separation_distance=2.0)
    h = hs.make_handle(lib, "ko_", cam=cam, mot_cfg=mot)
    dyn, seeds = [], []
    for (d, l), T, st in zip(frames, poses, stamps):
        img, ns, nc = h.detect_motion(h.make_frame(d, T, st, label=l))
        dyn.append(img)
        seeds.append((ns, nc))
        h.integrate_frame(h.make_frame(d, T, st, label=l, mask=img))
        h.update_tracking(st)
    out = {"depth": np.stack([f[0] for f in frames]), "label": np.stack([f[1] for f in frames]),
           "poses": np.stack(poses), "stamps": np.array(stamps, np.uint64),
           "dynamic_image": np.stack(dyn).astype(np.uint8), "seeds_clusters": np.array(seeds, np.int32)}
    out.update(pack_blocks(h.export_blocks(), "b_"))
    return out


def colour_case(lib):
    """K1 colour path: 6 orbit frames, the third one without a colour image."""
    cam = golden_camera()
    scene = syn.room_scene()
    poses, stamps = syn.orbit_trajectory(6, laps=0.06)
    frames = hs.render_frames(scene, cam, poses, stamps)
    cols = [syn.colorize(l, d) for d, l in frames]
    has = np.array([1, 1, 0, 1, 1, 1], np.uint8)
    h = hs.make_handle(lib, "ko_", cam=cam)
    hs.run_fusion(h, frames, poses, stamps, colors=[c if k else None for c, k in zip(cols, has)])
    out = {"depth": np.stack([f[0] for f in frames]), "label": np.stack([f[1] for f in frames]),
           "color": np.stack(cols), "has_color": has, "poses": np.stack(poses), "stamps": np.array(stamps, np.uint64)}
    b = h.export_blocks()
    out.update(pack_blocks(b, "b_"))
    out["b_color"] = b.color
    return out


def mesh_case(lib):
    """Marching cubes (kb_generate_mesh) over the map of the fusion fixture: inputs are read from fusion.npz, the fixture
    stores the mesh of all blocks and the map checksum."""
    g = np.load(os.path.join(HERE, "fusion.npz"))
    h = hs.make_handle(lib, "ko_", cam=golden_camera())
    frames = [(np.ascontiguousarray(d), np.ascontiguousarray(l)) for d, l in zip(g["depth"], g["label"])]
    hs.run_fusion(h, frames, list(g["poses"]), [int(s) for s in g["stamps"]], tracking=True)
    bi, off, pts, col, lab = h.generate_mesh(False, False)
    return {"block_index": bi, "offsets": off, "points_bits": pts.view(np.uint32), "labels": lab.astype(np.uint8),
            "checksum": np.array(h.map_checksum(), np.uint64)}


if __name__ == "__main__":
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    cases = (("fusion", fusion_case), ("dynamic", dynamic_case), ("colour", colour_case), ("mesh", mesh_case))
    for name, fn in cases:
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        data = fn(lib)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **data)
        print(name, os.path.getsize(path) // 1024, "KiB", {k: v.shape for k, v in data.items() if k in ("b_block_index", "dynamic_image")})
