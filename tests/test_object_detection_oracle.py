"""Known-answer tests of the oracle's ConnectedSemantics restatement (khronos/src/active_window/object_detection/
connected_semantics.cpp is fully in-tree, so this row is pinned by the reference's own code): independent numpy /
scipy re-derivations of both modes."""
import numpy as np
import pytest
from scipy import ndimage

from khronos_b200 import capi, synthetic as syn
import harness as hs

OBJECTS = (7, 8, 9, 10, 11, 12, 19)


def scene_frame(scale=4, noise_seed=None):
    cam = hs.small_camera(scale)
    scene = syn.room_scene()
    pose = syn.look_pose((6.0, 5.0, 1.5), 3.7, np.radians(12.0))
    d, l = syn.render(scene, cam, pose)
    d, l = d.numpy(), l.numpy().copy()
    if noise_seed is not None:  # salt the label image with small object specks and holes
        rng = np.random.default_rng(noise_seed)
        m = rng.random(l.shape) < 0.03
        l[m] = rng.choice(np.array(OBJECTS + (1, 3), np.int32), size=int(m.sum()))
    return cam, pose, d, l


def expected_2d(label, objects, full, min_size):
    """Per object label: scipy connected components; ids in the order of each component's first pixel in the
    reference's column-major scan (u outer, v inner); small clusters zeroed, ids not reused."""
    H, W = label.shape
    st = np.ones((3, 3), int) if full else ndimage.generate_binary_structure(2, 1)
    comps = []
    for lab in objects:
        cc, n = ndimage.label(label == lab, structure=st)
        for k in range(1, n + 1):
            vs, us = np.nonzero(cc == k)
            comps.append((int((us * H + vs).min()), lab, vs, us))
    comps.sort(key=lambda c: c[0])
    img = np.zeros((H, W), np.int32)
    kept = []
    for i, (_, lab, vs, us) in enumerate(comps):
        if len(vs) >= min_size:
            img[vs, us] = i + 1
            kept.append((i + 1, lab, len(vs)))
    return img, kept


@pytest.mark.parametrize("full,min_size", [(True, 0), (False, 0), (True, 12)])
def test_connected_semantics_2d_known_answer(oracle_lib, full, min_size):
    cam, pose, d, l = scene_frame(noise_seed=5)
    h = hs.make_handle(oracle_lib, "ko_", cam=cam)
    cfg = capi.default_object_detector_config(OBJECTS, use_3d=False, use_full_connectivity=full, min_cluster_size=min_size)
    img, nc = h.detect_objects(cfg, h.make_frame(d, pose, 1_000_000_000, label=l))
    want, kept = expected_2d(l, OBJECTS, full, min_size)
    np.testing.assert_array_equal(img, want)
    cl = h.get_object_clusters()
    assert nc == len(kept) == len(cl) and nc >= (1 if min_size else 6)
    assert [(c["id"], c["semantic_id"], len(c["pixels"])) for c in cl] == kept


def expected_3d(cam, pose, depth, label, objects, full, grid, min_size, max_size, max_range):
    """Independent restatement: voxel keys per pixel in fp32, BFS per semantic id; clusters of one id ordered by
    their smallest voxel (z, y, x); ids consecutive over the kept clusters."""
    F32 = np.float32
    H, W = label.shape
    v, u = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    x = ((u.astype(F32) - F32(cam.cx)) / F32(cam.fx) * depth).astype(F32)
    y = ((v.astype(F32) - F32(cam.cy)) / F32(cam.fy) * depth).astype(F32)
    T = np.asarray(pose, np.float64)
    Rw, tw = T[:3, :3].astype(F32), T[:3, 3].astype(F32)
    pw = [((Rw[r, 0] * x + Rw[r, 1] * y).astype(F32) + Rw[r, 2] * depth).astype(F32) + tw[r] for r in range(3)]
    inv = F32(1.0) / F32(grid)
    g = [np.floor((p.astype(F32) * inv).astype(F32)).astype(np.int64) for p in pw]
    offs = [(dx, dy, dz) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)
            if (dx, dy, dz) != (0, 0, 0) and (full or abs(dx) + abs(dy) + abs(dz) == 1)]
    img = np.zeros((H, W), np.int32)
    kept = []
    for lab in sorted(objects):
        sel = (label == lab)
        if max_range > 0:
            sel &= ~(depth > F32(max_range))
        vox = {}
        for vv, uu in zip(*np.nonzero(sel)):
            vox.setdefault((int(g[2][vv, uu]), int(g[1][vv, uu]), int(g[0][vv, uu])), []).append((vv, uu))
        remaining = set(vox)
        comps = []
        while remaining:
            seed = min(remaining)
            remaining.discard(seed)
            stack, members = [seed], [seed]
            while stack:
                z, yy, xx = stack.pop()
                for dx, dy, dz in offs:
                    n = (z + dz, yy + dy, xx + dx)
                    if n in remaining:
                        remaining.discard(n)
                        stack.append(n)
                        members.append(n)
            comps.append((min(members), members))
        for _, members in sorted(comps):
            px = [p for m in members for p in vox[m]]
            if len(px) < min_size or (max_size > 0 and len(px) > max_size):
                continue
            cid = len(kept) + 1
            for vv, uu in px:
                img[vv, uu] = cid
            kept.append((cid, lab, len(px)))
    return img, kept


@pytest.mark.parametrize("full,min_size,max_size,max_range", [(True, 0, -1, 0.0), (False, 0, -1, 0.0), (True, 20, 3000, 4.0)])
def test_connected_semantics_3d_known_answer(oracle_lib, full, min_size, max_size, max_range):
    cam, pose, d, l = scene_frame(noise_seed=7)
    h = hs.make_handle(oracle_lib, "ko_", cam=cam)
    cfg = capi.default_object_detector_config(OBJECTS, use_3d=True, use_full_connectivity=full, min_cluster_size=min_size,
                                              max_cluster_size=max_size, max_range=max_range, grid_size=0.1)
    img, nc = h.detect_objects(cfg, h.make_frame(d, pose, 1_000_000_000, label=l))
    want, kept = expected_3d(cam, pose, d, l, OBJECTS, full, 0.1, min_size, max_size, max_range)
    np.testing.assert_array_equal(img, want)
    cl = h.get_object_clusters()
    assert nc == len(kept) == len(cl) and nc >= (1 if min_size else 3)
    assert [(c["id"], c["semantic_id"], len(c["pixels"])) for c in cl] == kept
    for c in cl:  # the pixel lists are exactly the pixels carrying the id
        got = np.zeros_like(img, dtype=bool)
        got[c["pixels"][:, 1], c["pixels"][:, 0]] = True
        np.testing.assert_array_equal(got, img == c["id"])
