"""CPU-side tests of the product's host logic (no GPU): the M2-M4 host clustering path
(khronos_b200/csrc/kb_motion_host.cpp, through kb_host_cluster_motion) against the oracle's motion detector.
The per-pixel voxel keys + seed flags the M1 kernel would produce are re-derived in numpy (fp32, same
expression order) from the oracle's exported ever-free state."""
import ctypes as C

import numpy as np
import pytest

from khronos_b200 import capi, synthetic as syn
import harness as hs

F32 = np.float32


def m1_numpy(cam, pose, depth, blocks: capi.Blocks, vps, voxel_size):
    """FreeSpaceMotionDetector::setUpPointMapPart (free_space_motion_detector.cpp:158-203) in numpy fp32."""
    H, W = depth.shape
    T = np.asarray(pose, np.float64)
    Rw, tw = T[:3, :3].astype(F32), T[:3, 3].astype(F32)
    u = np.arange(W, dtype=F32)[None, :].repeat(H, 0)
    v = np.arange(H, dtype=F32)[:, None].repeat(W, 1)
    x = ((u - F32(cam.cx)) / F32(cam.fx) * depth).astype(F32)
    y = ((v - F32(cam.cy)) / F32(cam.fy) * depth).astype(F32)
    z = depth
    w = [(((Rw[r, 0] * x + Rw[r, 1] * y).astype(F32) + Rw[r, 2] * z).astype(F32) + tw[r]).astype(F32) for r in range(3)]
    bs = F32(F32(voxel_size) * F32(vps))
    bsi, vsi = F32(1) / bs, F32(1) / F32(voxel_size)
    b = [np.floor(w[a] * bsi).astype(np.int64) for a in range(3)]
    vx = [np.floor(((w[a] - b[a].astype(F32) * bs).astype(F32)) * vsi).astype(np.int64) for a in range(3)]
    valid = depth > 0
    for a in range(3):
        valid &= (vx[a] >= 0) & (vx[a] < vps)
    lut = {tuple(ix): k for k, ix in enumerate(blocks.block_index.tolist())}
    gidx = np.full((H, W, 3), 0, np.int32)
    gidx[..., 0] = np.iinfo(np.int32).min
    seed = np.zeros((H, W), np.uint8)
    for vv, uu in zip(*np.nonzero(valid)):
        k = lut.get((int(b[0][vv, uu]), int(b[1][vv, uu]), int(b[2][vv, uu])))
        if k is None:
            continue
        ix = [int(vx[a][vv, uu]) for a in range(3)]
        gidx[vv, uu] = [int(b[a][vv, uu]) * vps + ix[a] for a in range(3)]
        seed[vv, uu] = blocks.ever_free[k, ix[0] + vps * (ix[1] + vps * ix[2])]
    return gidx, seed


@pytest.mark.parametrize("sep", [2.0, 1.0, 0.0])
def test_host_clustering_matches_oracle(oracle_lib, product_lib, sep):
    cam = syn.make_camera(80, 60, 40.0, 40.0, max_range=2.5)
    scene = syn.room_scene()
    scene.mover = ((0.4, 0.4, 1.0), (7.6, 3.8, 0.9), (0.0, 1.2, 0.0), 1.6)
    n, dt = 16, 200_000_000
    pose = syn.look_pose((6.0, 5.0, 1.5), 0.0, np.radians(10.0))
    poses, stamps = [pose] * n, [1_000_000_000 + i * dt for i in range(n)]
    frames = hs.render_frames(scene, cam, poses, stamps)
    mot = capi.default_motion_config(min_cluster_size=4, min_separation_distance=sep)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam, mot_cfg=mot)
    fn = product_lib.kb_host_cluster_motion
    fn.restype = C.c_int
    checked = 0
    for (d, l), T, st in zip(frames, poses, stamps):
        blocks = o.export_blocks(likelihoods=False)  # ever-free state *before* this frame's detection
        img_o, ns_o, nc_o = o.detect_motion(o.make_frame(d, T, st, label=l))
        if blocks.n:
            gidx, seed = m1_numpy(cam, T, d, blocks, 16, 0.05)
            img_h = np.zeros_like(img_o)
            ns, nc = C.c_int32(0), C.c_int32(0)
            Tm = (C.c_double * 16)(*np.asarray(T, np.float64).reshape(16))
            st_ = fn(C.byref(cam), C.byref(mot), Tm, C.c_void_p(gidx.ctypes.data), C.c_void_p(seed.ctypes.data),
                     C.c_void_p(d.ctypes.data), C.c_void_p(img_h.ctypes.data), C.byref(ns), C.byref(nc))
            assert st_ == 0
            assert (ns.value, nc.value) == (ns_o, nc_o)
            np.testing.assert_array_equal(img_h, img_o)
            checked += nc_o
        o.integrate_frame(o.make_frame(d, T, st, label=l, mask=img_o), want_stats=False)
        o.update_tracking(st)
    assert checked > 3
