"""Marching cubes (docs/ORACLE_SPEC.md §13): the case table and the oracle's mesher.
The reference's mesher lives in un-vendored Hydra (parity unpinned); what can be pinned is pinned here:
  * both copies of the 256-case table (product + oracle) are identical, every row uses exactly the cube edges whose end
    points differ in sign, and meshes of random sign fields are closed, 2-manifold and consistently oriented;
  * the oracle's mesh equals an independent numpy restatement of the spec on its own exported TSDF (bit-exact vertices);
  * geometry: the mesh of a fused flat wall lies on the wall, faces the camera side, and its area matches."""
import os
import re
from collections import Counter

import numpy as np
import pytest

from khronos_b200 import capi, synthetic as syn
import harness as hs

EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
OFFS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]


def _table(path):
    src = open(os.path.join(hs.ROOT, path)).read()
    body = src[src.index("[256][16] = {"):]
    body = body[:body.index("};")]
    rows = [[int(x) for x in re.findall(r"-?\d+", ln)] for ln in body.splitlines()[1:] if "{" in ln]
    t = np.array(rows)
    assert t.shape == (256, 16)
    return t


TABLE = _table("oracle/oracle_mc_tables.hpp")


def test_table_copies_identical_and_edge_sets_exact():
    np.testing.assert_array_equal(TABLE, _table("khronos_b200/csrc/kb_mc_tables.h"))
    for c in range(256):
        active = {e for e, (a, b) in enumerate(EDGES) if ((c >> a) & 1) != ((c >> b) & 1)}
        used = [e for e in TABLE[c] if e >= 0]
        assert set(used) == active and len(used) % 3 == 0, c
        k = len(used)
        assert all(e == -1 for e in TABLE[c][k:])


def test_table_meshes_are_closed_manifold_and_oriented():
    rng = np.random.default_rng(0)
    for p in (0.5, 0.3, 0.7):
        n = 10
        neg = np.zeros((n, n, n), bool)
        neg[1:-1, 1:-1, 1:-1] = rng.random((n - 2, n - 2, n - 2)) < p
        directed = Counter()
        for x in range(n - 1):
            for y in range(n - 1):
                for z in range(n - 1):
                    c = sum(1 << i for i, (dx, dy, dz) in enumerate(OFFS) if neg[x + dx, y + dy, z + dz])
                    row = TABLE[c]
                    k = 0
                    while k < 16 and row[k] >= 0:
                        vs = []
                        for e in (row[k + 2], row[k + 1], row[k]):
                            a, b = EDGES[e]
                            pa = (x + OFFS[a][0], y + OFFS[a][1], z + OFFS[a][2])
                            pb = (x + OFFS[b][0], y + OFFS[b][1], z + OFFS[b][2])
                            vs.append((min(pa, pb), max(pa, pb)))
                        for i in range(3):
                            directed[(vs[i], vs[(i + 1) % 3])] += 1
                        k += 3
        assert directed
        for (a, b), cnt in directed.items():
            assert cnt == 1 and directed.get((b, a), 0) == 1


def numpy_mesh(blocks: capi.Blocks, voxel_size, vps, min_weight=1e-4, only=None):
    """Spec restatement (docs/ORACLE_SPEC.md §13) on an exported map: list of (block index, points (n,3) f32, labels)."""
    f32 = np.float32
    V = vps ** 3
    idx = {tuple(b): i for i, b in enumerate(blocks.block_index.reshape(-1, 3).tolist())}
    bs = f32(voxel_size) * f32(vps)
    m = vps - 1
    order = ([(x, y, z) for x in range(m) for y in range(m) for z in range(m)] + [(m, y, z) for z in range(vps) for y in range(vps)] +
             [(x, m, z) for z in range(vps) for x in range(m)] + [(x, y, m) for y in range(m) for x in range(m)])
    out = []
    for b in sorted(idx):
        if only is not None and b not in only:
            continue
        D = np.zeros((vps + 1,) * 3, f32)
        Wt = np.full((vps + 1,) * 3, -1.0, f32)
        LB = np.zeros((vps + 1,) * 3, np.uint32)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    nb = (b[0] + dx, b[1] + dy, b[2] + dz)
                    if nb not in idx:
                        continue
                    i = idx[nb]
                    d = blocks.distance[i].reshape(vps, vps, vps).transpose(2, 1, 0)  # [x, y, z]
                    w = blocks.weight[i].reshape(vps, vps, vps).transpose(2, 1, 0)
                    lb = np.where(blocks.semantic_empty[i] != 0, 0, blocks.semantic_label[i]).reshape(vps, vps, vps).transpose(2, 1, 0)
                    sx = slice(0, vps) if dx == 0 else slice(vps, vps + 1)
                    sy = slice(0, vps) if dy == 0 else slice(vps, vps + 1)
                    sz = slice(0, vps) if dz == 0 else slice(vps, vps + 1)
                    D[sx, sy, sz] = d[(slice(0, vps) if dx == 0 else slice(0, 1)), (slice(0, vps) if dy == 0 else slice(0, 1)), (slice(0, vps) if dz == 0 else slice(0, 1))]
                    Wt[sx, sy, sz] = w[(slice(0, vps) if dx == 0 else slice(0, 1)), (slice(0, vps) if dy == 0 else slice(0, 1)), (slice(0, vps) if dz == 0 else slice(0, 1))]
                    LB[sx, sy, sz] = lb[(slice(0, vps) if dx == 0 else slice(0, 1)), (slice(0, vps) if dy == 0 else slice(0, 1)), (slice(0, vps) if dz == 0 else slice(0, 1))]
        ok = np.ones((vps,) * 3, bool)
        case = np.zeros((vps,) * 3, np.int32)
        for c, (ox, oy, oz) in enumerate(OFFS):
            sl = (slice(ox, ox + vps), slice(oy, oy + vps), slice(oz, oz + vps))
            ok &= Wt[sl] >= f32(min_weight)
            case |= (D[sl] < 0).astype(np.int32) << c
        case[~ok] = 0
        case[case == 255] = 0
        pts, labs = [], []
        if case.any():
            for (x, y, z) in order:
                c = int(case[x, y, z])
                if not c:
                    continue
                pos, sdf, lab = [], [], []
                for (ox, oy, oz) in OFFS:
                    vx, vy, vz = x + ox, y + oy, z + oz
                    bx, by, bz = b[0] + (vx == vps), b[1] + (vy == vps), b[2] + (vz == vps)
                    lx, ly, lz = vx % vps, vy % vps, vz % vps
                    pos.append(np.array([f32(bx) * bs + (f32(lx) + f32(0.5)) * f32(voxel_size), f32(by) * bs + (f32(ly) + f32(0.5)) * f32(voxel_size),
                                         f32(bz) * bs + (f32(lz) + f32(0.5)) * f32(voxel_size)], f32))
                    sdf.append(D[vx, vy, vz])
                    lab.append(LB[vx, vy, vz])
                row = TABLE[c]
                k = 0
                while k < 16 and row[k] >= 0:
                    for e in (row[k + 2], row[k + 1], row[k]):
                        c0, c1 = EDGES[e]
                        diff = f32(sdf[c0] - sdf[c1])
                        if abs(diff) >= f32(1e-6):
                            t = f32(sdf[c0] / diff)
                            v = (pos[c0] + t * (pos[c1] - pos[c0])).astype(f32)
                        else:
                            t = f32(0.5)
                            v = (f32(0.5) * (pos[c0] + pos[c1])).astype(f32)
                        pts.append(v)
                        labs.append(lab[c0] if t < f32(0.5) else lab[c1])
                    k += 3
        out.append((b, np.array(pts, f32).reshape(-1, 3), np.array(labs, np.uint32)))
    return out


def _room(n=5, scale=8):
    cam = hs.small_camera(scale)
    scene = syn.room_scene()
    poses, stamps = syn.orbit_trajectory(n, laps=0.1)
    return cam, hs.render_frames(scene, cam, poses, stamps), poses, stamps


def test_oracle_mesh_equals_numpy_restatement(oracle_lib):
    cam, frames, poses, stamps = _room()
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    hs.run_fusion(o, frames, poses, stamps)
    bi, off, pts, col, lab = o.generate_mesh(only_mesh_updated=False, clear_updated_flag=False)
    ref = numpy_mesh(o.export_blocks(), 0.05, 16)
    assert len(ref) == len(bi) and off[-1] == len(pts) and len(pts) % 3 == 0 and len(pts) > 3000
    for i, (b, p, l) in enumerate(ref):
        assert tuple(bi[i]) == b
        np.testing.assert_array_equal(pts[off[i]:off[i + 1]].view(np.uint32), p.view(np.uint32), err_msg=f"block {b}")
        np.testing.assert_array_equal(lab[off[i]:off[i + 1]], l)


def test_mesh_updated_flag_semantics(oracle_lib):
    cam, frames, poses, stamps = _room(4)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    hs.run_fusion(o, frames[:2], poses[:2], stamps[:2])
    bi1, off1, *_ = o.generate_mesh(True, True)
    assert len(bi1) > 0
    bi2, off2, *_ = o.generate_mesh(True, True)
    assert len(bi2) == 0 and off2[-1] == 0          # flags cleared
    hs.run_fusion(o, frames[2:], poses[2:], stamps[2:])
    bi3, *_ = o.generate_mesh(True, False)
    bi4, *_ = o.generate_mesh(True, False)
    assert 0 < len(bi3) == len(bi4)                  # clear_updated_flag = false keeps them (extractor's call)
    flags = o.export_blocks().block_flags
    assert int(((flags & capi.FLAG_MESH_UPDATED) != 0).sum()) == len(bi3)


def test_flat_wall_mesh_geometry(oracle_lib):
    """Camera looks along +x at a wall x = 3: after fusing one frame the mesh is the plane x = 3 (vertex error well below
    a voxel), every triangle faces the camera (-x, the positive-sdf side), and the triangle areas tile the observed wall."""
    cam = hs.small_camera(4)
    T = syn.look_pose((0.0, 0.0, 1.0), 0.0, 0.0)
    d = np.zeros((cam.height, cam.width), np.float32)
    # z-depth of the plane x = 3 for a camera at the origin looking along +x is 3 everywhere
    d[:] = 3.0
    l = np.full((cam.height, cam.width), 3, np.int32)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    o.integrate_frame(o.make_frame(d, T, 1_000_000_000, label=l))
    bi, off, pts, col, lab = o.generate_mesh(False, False)
    assert len(pts) > 300
    assert np.abs(pts[:, 0] - 3.0).max() < 2e-3
    tri = pts.reshape(-1, 3, 3)
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    assert (nrm[:, 0] < 0).all()
    area = 0.5 * np.linalg.norm(nrm, axis=1).sum()
    # the frustum at depth 3 spans (W-1)/fx*3 x (H-1)/fy*3 metres; the mesh covers it up to a voxel-wide rim
    full = (cam.width - 1) / cam.fx * 3.0 * (cam.height - 1) / cam.fy * 3.0
    assert 0.85 * full < area < 1.05 * full
    assert set(np.unique(lab)) == {3}
