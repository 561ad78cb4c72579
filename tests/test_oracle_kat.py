"""Known-answer tests pinning the CPU oracle against independently written numpy restatements of the
formulas in docs/ORACLE_SPEC.md (hand-computable cases: flat wall, identity pose, single labels).
The reference ships no tests or golden vectors for this path ("parity unpinned"), so these KATs and
tests/golden/ are what pins the oracle."""
import numpy as np
import pytest

from khronos_b200 import capi, synthetic as syn
import harness as hs

F32 = np.float32


def flat_wall(cam, depth_value, label_value=3):
    d = np.full((cam.height, cam.width), depth_value, np.float32)
    l = np.full((cam.height, cam.width), label_value, np.int32)
    return d, l


def expected_flat_wall(cam, mc, ic, bidx, D, n_frames=1, bilinear_only=False):
    """numpy fp32 restatement for an identity pose (p_C == p_W) and a constant depth image."""
    vs, vps, trunc = F32(mc.voxel_size), mc.voxels_per_side, F32(mc.truncation_distance)
    bs = F32(vs * F32(vps))
    lin = np.arange(vps ** 3)
    vx, vy, vz = lin % vps, (lin // vps) % vps, lin // (vps * vps)
    px = F32(bidx[0]) * bs + (vx.astype(F32) + F32(0.5)) * vs
    py = F32(bidx[1]) * bs + (vy.astype(F32) + F32(0.5)) * vs
    pz = F32(bidx[2]) * bs + (vz.astype(F32) + F32(0.5)) * vs
    with np.errstate(divide="ignore", invalid="ignore"):
        u = F32(cam.fx) * px / pz + F32(cam.cx)
        v = F32(cam.fy) * py / pz + F32(cam.cy)
    valid = (pz > 0) & (u >= 0) & (u <= cam.width - 1) & (v >= 0) & (v <= cam.height - 1)
    if bilinear_only:  # pure bilinear needs the full 2x2 footprint inside the image
        valid &= (np.floor(u) + 1 < cam.width) & (np.floor(v) + 1 < cam.height)
    sdf = F32(D) - pz
    valid &= ~(sdf < -trunc)
    w = (F32(cam.fx) * F32(cam.fy)) * (vs * vs) / (pz * pz)
    w = w / (pz * pz)
    eps = F32(ic.weight_dropoff_epsilon) * -vs if ic.weight_dropoff_epsilon <= 0 else F32(ic.weight_dropoff_epsilon)
    drop = sdf < -eps
    w = np.where(drop, np.maximum(w * ((trunc + sdf) / (trunc - eps)), F32(0)), w).astype(F32)
    sdf_c = np.clip(sdf, -trunc, trunc).astype(F32)
    dist = np.zeros(vps ** 3, F32)
    wt = np.zeros(vps ** 3, F32)
    for _ in range(n_frames):
        with np.errstate(divide="ignore", invalid="ignore"):
            nd = ((dist * wt + sdf_c * w) / (wt + w)).astype(F32)
        nw = np.minimum(wt + w, F32(ic.max_weight)).astype(F32)
        dist = np.where(valid, nd, dist)
        wt = np.where(valid, nw, wt)
    band = valid & (np.abs(sdf) < trunc)
    return valid, band, dist, wt


@pytest.mark.parametrize("interp", [capi.INTERP_NEAREST, capi.INTERP_ADAPTIVE, capi.INTERP_BILINEAR])
def test_flat_wall_known_answer(oracle_lib, interp):
    cam = syn.make_camera(64, 48, 32.0, 32.0, max_range=3.0)
    mc = capi.default_map_config(voxel_size=0.1, vps=8, trunc=0.3)
    ic = capi.default_integrator_config(interpolation=interp, num_threads=1)
    h = hs.make_handle(oracle_lib, "ko_", map_cfg=mc, integ_cfg=ic, cam=cam)
    D = 2.0
    d, l = flat_wall(cam, D)
    for k in range(2):
        h.integrate_frame(h.make_frame(d, np.eye(4), 1_000_000_000 + k, label=l))
    b = h.export_blocks()
    checked = 0
    for i, bidx in enumerate(b.block_index):
        valid, band, dist, wt = expected_flat_wall(cam, mc, ic, bidx, D, n_frames=2,
                                                   bilinear_only=interp == capi.INTERP_BILINEAR)
        np.testing.assert_array_equal(b.last_observed[i] != 0, valid)
        if interp == capi.INTERP_NEAREST:
            np.testing.assert_array_equal(b.distance[i], dist)
            np.testing.assert_array_equal(b.weight[i], wt)
        else:  # sum(w_i * D) rounds to D within an ulp; the numpy side uses D itself
            np.testing.assert_allclose(b.distance[i], dist, rtol=0, atol=3e-7)
            np.testing.assert_allclose(b.weight[i], wt, rtol=1e-5)
        np.testing.assert_array_equal(b.semantic_empty[i] == 0, band)
        assert (b.semantic_label[i][band] == 3).all()
        checked += int(valid.sum())
    assert checked > 500
    # hand numbers: the voxel centred at (0.05, 0.05, 1.95) sees sdf = +0.05, w = 32*32*0.01/1.95^4
    i = int(np.where((b.block_index == [0, 0, 2]).all(1))[0][0])
    lin = 0 + 8 * (0 + 8 * 3)  # voxel (0,0,3): z = 1.6 + 0.35 = 1.95
    assert b.distance[i, lin] == pytest.approx(0.05, abs=1e-6)
    assert b.weight[i, lin] == pytest.approx(2 * 32 * 32 * 0.01 / 1.95 ** 4, rel=1e-5)


def test_frustum_block_selection_known_answer(oracle_lib):
    cam = syn.make_camera(64, 48, 32.0, 32.0, min_range=0.1, max_range=3.0)
    mc = capi.default_map_config(voxel_size=0.1, vps=8, trunc=0.3)
    h = hs.make_handle(oracle_lib, "ko_", map_cfg=mc, cam=cam)
    pose = syn.look_pose((0.3, -0.2, 1.1), 0.7, 0.2)
    d, l = flat_wall(cam, 0.0)  # all invalid: only allocation happens
    st = h.integrate_frame(h.make_frame(d, pose, 5, label=l))
    got = set(map(tuple, h.export_blocks().block_index.tolist()))
    # independent restatement in float32
    bs = F32(F32(0.1) * F32(8))
    Tinv = np.linalg.inv(pose)
    R, t = Tinv[:3, :3].astype(F32), Tinv[:3, 3].astype(F32)
    infl = F32(bs * F32(0.8660254))
    want = set()
    rng_ = range(-12, 13)
    xl, xr = (F32(0) - F32(cam.cx)) / F32(cam.fx), (F32(cam.width - 1) - F32(cam.cx)) / F32(cam.fx)
    yt, yb = (F32(0) - F32(cam.cy)) / F32(cam.fy), (F32(cam.height - 1) - F32(cam.cy)) / F32(cam.fy)
    il, ir = F32(1) / np.sqrt(F32(1) + xl * xl), F32(1) / np.sqrt(F32(1) + xr * xr)
    it, ib = F32(1) / np.sqrt(F32(1) + yt * yt), F32(1) / np.sqrt(F32(1) + yb * yb)
    for bx in rng_:
        for by in rng_:
            for bz in rng_:
                c = (np.array([bx, by, bz], F32) + F32(0.5)) * bs
                p = np.array([((R[r, 0] * c[0] + R[r, 1] * c[1]) + R[r, 2] * c[2]) + t[r] for r in range(3)], F32)
                if p[2] < -infl:
                    continue
                r = np.sqrt((p[0] * p[0] + p[1] * p[1]) + p[2] * p[2])
                if r < F32(cam.min_range) - infl or r > F32(cam.max_range) + infl:
                    continue
                if il * p[0] + (-xl * il) * p[2] < -infl or (-ir) * p[0] + (xr * ir) * p[2] < -infl:
                    continue
                if it * p[1] + (-yt * it) * p[2] < -infl or (-ib) * p[1] + (yb * ib) * p[2] < -infl:
                    continue
                want.add((bx, by, bz))
    assert got == want and len(got) > 20
    assert st.blocks_in_frustum == len(want) and st.voxels_updated == 0


def test_mle_semantic_known_answer(oracle_lib):
    cam = syn.make_camera(64, 48, 32.0, 32.0, max_range=3.0)
    mc = capi.default_map_config(voxel_size=0.1, vps=8, trunc=0.3)
    ic = capi.default_integrator_config(num_labels=5, interpolation=capi.INTERP_NEAREST)
    h = hs.make_handle(oracle_lib, "ko_", map_cfg=mc, integ_cfg=ic, cam=cam)
    seq = [3, 3, 1, 3, 1, 1, 1]
    for k, lab in enumerate(seq):
        d, l = flat_wall(cam, 2.0, lab)
        h.integrate_frame(h.make_frame(d, np.eye(4), 10 + k, label=l))
    b = h.export_blocks()
    a, off, init = F32(np.log(np.float64(F32(0.9)))), F32(np.log((1.0 - np.float64(F32(0.9))) / 4.0)), F32(np.log(1.0 / 5.0))
    lik = np.full(5, init, F32)
    for lab in seq:
        for k in range(5):
            lik[k] = F32(lik[k] + (a if k == lab else off))
    i = int(np.where((b.block_index == [0, 0, 2]).all(1))[0][0])
    lin = 0 + 8 * (0 + 8 * 3)
    np.testing.assert_array_equal(b.semantic_likelihoods[i, lin], lik)
    assert b.semantic_label[i, lin] == 1 and b.semantic_empty[i, lin] == 0
    # label >= N is not a valid label: TSDF integrates, semantics untouched; negative labels likewise
    for bad in (7, -1):
        h2 = hs.make_handle(oracle_lib, "ko_", map_cfg=mc, integ_cfg=ic, cam=cam)
        d, l = flat_wall(cam, 2.0, bad)
        st = h2.integrate_frame(h2.make_frame(d, np.eye(4), 10, label=l))
        assert st.voxels_in_band > 0 and st.voxels_semantic == 0
    # blocked (dynamic/invalid) label: band voxels are skipped entirely
    icb = capi.default_integrator_config(num_labels=5, blocked=(3,), interpolation=capi.INTERP_NEAREST)
    h3 = hs.make_handle(oracle_lib, "ko_", map_cfg=mc, integ_cfg=icb, cam=cam)
    d, l = flat_wall(cam, 2.0, 3)
    st = h3.integrate_frame(h3.make_frame(d, np.eye(4), 10, label=l))
    assert st.voxels_in_band == 0 and st.voxels_updated > 0


def test_tracking_transitions_known_answer(oracle_lib):
    """tracking_integrator.cpp:133-166,224-252 on a hand-made timeline (stamps in ns)."""
    cam = syn.make_camera(64, 48, 32.0, 32.0, max_range=3.0)
    mc = capi.default_map_config(voxel_size=0.1, vps=8, trunc=0.3)
    ic = capi.default_integrator_config(interpolation=capi.INTERP_NEAREST)
    h = hs.make_handle(oracle_lib, "ko_", map_cfg=mc, integ_cfg=ic, cam=cam)
    d, l = flat_wall(cam, 2.0)
    S = 1_000_000_000
    key_blk, free_lin, occ_lin = [0, 0, 1], 0 + 8 * (0 + 8 * 5), None  # z = 0.8+0.55 = 1.35: free space
    stamps = [1 * S, 1 * S + S // 2, 2 * S + S // 4, 2 * S + S // 2]
    for st in stamps:
        h.integrate_frame(h.make_frame(d, np.eye(4), st, label=l))
        h.update_tracking(st)
    b = h.export_blocks()
    i = int(np.where((b.block_index == key_blk).all(1))[0][0])
    # free-space voxel: sdf clamped to +0.3 >= thr 0.15 from its first update on, so last_occupied stays
    # 0 (it was updated by K1 before K2 ever saw it), observed every frame, active
    assert b.distance[i, free_lin] == pytest.approx(0.3)
    assert b.last_occupied[i, free_lin] == 0 and b.last_observed[i, free_lin] == stamps[-1]
    assert b.active[i, free_lin] == 1 and b.to_remove[i, free_lin] == 0
    # voxelIsFree needs toSeconds(0) < now - 1.0 (true from the 2.25 s frame on); its 18 neighbours are
    # equally free => ever_free
    assert b.ever_free[i, free_lin] == 1
    # a voxel on the surface (sdf ~ 0.05 < thr) is occupied: last_occupied == now, never ever-free
    j = int(np.where((b.block_index == [0, 0, 2]).all(1))[0][0])
    surf = 0 + 8 * (0 + 8 * 3)
    assert b.last_occupied[j, surf] == stamps[-1] and b.ever_free[j, surf] == 0
    # an unobserved voxel (behind the wall) is "occupied" (distance 0 < thr) but never observed
    behind = 0 + 8 * (0 + 8 * 7)  # z = 1.6 + 0.75 = 2.35: sdf = -0.35 < -trunc -> not integrated
    assert b.last_observed[j, behind] == 0 and b.last_occupied[j, behind] == stamps[-1]
    # active quirk: last_observed == 0 counts as active while now <= temporal_window (3 s)
    assert b.active[j, behind] == 1
    # jump 4 s ahead without observing: everything leaves the window -> to_remove, no active data
    h.update_tracking(7 * S)
    b = h.export_blocks()
    assert b.active.sum() == 0 and b.to_remove.all()
    assert (b.block_flags & capi.FLAG_HAS_ACTIVE_DATA).sum() == 0
    removed = h.reset_inactive()
    assert len(removed) == b.n and h.num_blocks() == 0


def test_everfree_needs_all_neighbours(oracle_lib):
    """Voxels at the edge of the observed volume have unobserved neighbours (or missing blocks) and
    must not become ever-free (tracking_integrator.cpp:186-215)."""
    cam = syn.make_camera(64, 48, 32.0, 32.0, max_range=3.0)
    mc = capi.default_map_config(voxel_size=0.1, vps=8, trunc=0.3)
    h = hs.make_handle(oracle_lib, "ko_", map_cfg=mc, cam=cam,
                       integ_cfg=capi.default_integrator_config(interpolation=capi.INTERP_NEAREST))
    d, l = flat_wall(cam, 2.0)
    S = 1_000_000_000
    for k in range(4):
        st = S + k * S // 2
        h.integrate_frame(h.make_frame(d, np.eye(4), st, label=l))
        h.update_tracking(st)
    b = h.export_blocks()
    ef = b.ever_free.astype(bool)
    obs = b.last_observed != 0
    assert ef.sum() > 0 and not (ef & ~obs).any()
    # brute-force re-derivation of the 18-neighbourhood rule from the exported state
    vps = 8
    lut = {tuple(ix): k for k, ix in enumerate(b.block_index.tolist())}
    now, buf = b.last_observed.max(), 1.0
    free = ((b.last_occupied.astype(np.float64) / 1e9) < now / 1e9 - buf) & obs
    offs = [(dx, dy, dz) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)
            if 0 < abs(dx) + abs(dy) + abs(dz) <= 2]
    want = np.zeros_like(ef)
    for k, ix in enumerate(b.block_index.tolist()):
        for lin in np.nonzero(free[k])[0]:
            x, y, z = lin % vps, (lin // vps) % vps, lin // (vps * vps)
            ok = True
            for dx, dy, dz in offs:
                nx, ny, nz = x + dx, y + dy, z + dz
                nb = (ix[0] + nx // vps, ix[1] + ny // vps, ix[2] + nz // vps)
                kk = lut.get(nb)
                if kk is None or not free[kk, (nx % vps) + vps * ((ny % vps) + vps * (nz % vps))]:
                    ok = False
                    break
            want[k, lin] = ok
    np.testing.assert_array_equal(ef, want)


def test_binary_confidence_scan_known_answer(oracle_lib):
    cam = syn.make_camera(64, 48, 32.0, 32.0, max_range=3.0)
    mc = capi.default_map_config(voxel_size=0.1, vps=8, trunc=0.2, with_tracking=False)
    ic = capi.default_integrator_config(semantic_mode=capi.SEM_BINARY, interpolation=capi.INTERP_NEAREST)
    h = hs.make_handle(oracle_lib, "ko_", map_cfg=mc, integ_cfg=ic, cam=cam)
    h.allocate_box((-1, -1, 1), (0, 0, 2))
    assert h.num_blocks() == 8
    d, _ = flat_wall(cam, 2.0)
    obj = np.zeros((cam.height, cam.width), np.int32)
    obj[:, cam.width // 2:] = 5  # right half of the image is object 5
    for k in range(4):
        h.integrate_frame(h.make_frame(d, np.eye(4), 10 + k, object_image=obj, target_id=5), allocate_blocks=False)
    b = h.export_blocks()
    nz = b.semantic_empty == 0
    assert nz.sum() > 0
    counts = b.semantic_likelihoods[nz]
    assert set(map(tuple, counts.tolist())) <= {(4.0, 0.0), (0.0, 4.0)}
    assert ((b.semantic_label[nz] == 1) == (counts[:, 1] == 4.0)).all()
    before = b.distance.copy()
    erased = h.scan_object_confidence(0.5, 3)
    a = h.export_blocks()
    conf = np.where(b.semantic_empty == 1, 0.0, np.where(b.semantic_likelihoods.sum(-1) < 3, -1.0,
                    b.semantic_likelihoods[..., 1] / np.maximum(b.semantic_likelihoods.sum(-1), 1e-9)))
    hit = (before <= 0) & (conf < 0.5)
    assert erased == hit.sum() and erased > 0
    np.testing.assert_array_equal(a.distance[hit], np.float32(0.2))
    np.testing.assert_array_equal(a.distance[~hit], before[~hit])


def test_motion_seed_and_cluster_known_answer(oracle_lib):
    """M1-M4 on a hand-made case: a small plate appears inside ever-free space."""
    cam = syn.make_camera(64, 48, 32.0, 32.0, max_range=3.0)
    mc = capi.default_map_config(voxel_size=0.1, vps=8, trunc=0.3)
    mot = capi.default_motion_config(min_cluster_size=3, min_separation_distance=2.0)
    h = hs.make_handle(oracle_lib, "ko_", map_cfg=mc, cam=cam, mot_cfg=mot,
                       integ_cfg=capi.default_integrator_config(interpolation=capi.INTERP_NEAREST))
    d, l = flat_wall(cam, 2.0)
    S = 1_000_000_000
    for k in range(4):
        st = S + k * S // 2
        img, ns, nc = h.detect_motion(h.make_frame(d, np.eye(4), st, label=l))
        assert ns == 0 and nc == 0 and not img.any()
        h.integrate_frame(h.make_frame(d, np.eye(4), st, label=l))
        h.update_tracking(st)
    d2 = d.copy()
    d2[20:28, 28:36] = 1.25  # plate at z = 1.25 m, deep inside observed free space
    img, ns, nc = h.detect_motion(h.make_frame(d2, np.eye(4), 4 * S, label=l))
    assert nc == 1 and ns >= 1
    assert (img[20:28, 28:36] == 1).all()
    assert (img > 0).sum() == 64  # nothing else is flagged
    cl = h.get_motion_clusters()
    assert len(cl) == 1 and len(cl[0]["pixels"]) >= 64
    assert (cl[0]["voxels"][:, 2] == 12).all()  # global z index of 1.25 m at 0.1 m voxels
    np.testing.assert_allclose(cl[0]["bbox"][[2, 5]], [1.25, 1.25])


@pytest.mark.parametrize("interp", [capi.INTERP_NEAREST, capi.INTERP_ADAPTIVE])
def test_colour_blend_known_answer(oracle_lib, interp):
    """ORACLE_SPEC §5.6: interpolateColor + Color::merge, restated independently in numpy fp32 (identity pose,
    flat wall, three frames with different textured colour images)."""
    cam = syn.make_camera(64, 48, 32.0, 32.0, max_range=3.0)
    mc = capi.default_map_config(voxel_size=0.1, vps=8, trunc=0.3)
    ic = capi.default_integrator_config(interpolation=interp, num_threads=1)
    h = hs.make_handle(oracle_lib, "ko_", map_cfg=mc, integ_cfg=ic, cam=cam)
    D = 2.0
    d, _ = flat_wall(cam, D)
    labs = [np.full((cam.height, cam.width), k, np.int32) for k in (3, 9, 4)]
    cols = [syn.colorize(l, d) for l in labs]
    for k, (l, c) in enumerate(zip(labs, cols)):
        h.integrate_frame(h.make_frame(d, np.eye(4), 1_000_000_000 + k, label=l, color=c))
    b = h.export_blocks()
    assert b.color.any()
    W, H = cam.width, cam.height
    vps, vs, trunc = 8, F32(0.1), F32(0.3)
    bs = F32(vs * F32(vps))
    checked = 0
    eps = F32(0.1)
    for i, bidx in enumerate(b.block_index):
        expect = np.zeros((vps ** 3, 3), np.uint8)
        for j in range(vps ** 3):
            vx, vy, vz = j % vps, (j // vps) % vps, j // (vps * vps)
            px = F32(F32(bidx[0]) * bs + F32(F32(F32(vx) + F32(0.5)) * vs))
            py = F32(F32(bidx[1]) * bs + F32(F32(F32(vy) + F32(0.5)) * vs))
            pz = F32(F32(bidx[2]) * bs + F32(F32(F32(vz) + F32(0.5)) * vs))
            if pz <= 0:
                continue
            u = F32(F32(F32(cam.fx) * px) / pz + F32(cam.cx))
            v = F32(F32(F32(cam.fy) * py) / pz + F32(cam.cy))
            if u < 0 or u > W - 1 or v < 0 or v > H - 1:
                continue
            u0, v0 = int(np.floor(u)), int(np.floor(v))
            bil = interp == capi.INTERP_ADAPTIVE and u0 + 1 < W and v0 + 1 < H  # flat wall: taps valid and equal
            if bil:
                du, dv = F32(u - F32(u0)), F32(v - F32(v0))
                w4 = [F32(F32(F32(1) - du) * F32(F32(1) - dv)), F32(F32(F32(1) - du) * dv), F32(du * F32(F32(1) - dv)), F32(du * dv)]
                rng = F32(F32(F32(w4[0] * F32(D)) + F32(w4[1] * F32(D))) + F32(w4[2] * F32(D)))
                rng = F32(rng + F32(w4[3] * F32(D)))
            else:
                rng = F32(D)
                # np.round is half-to-even, std::round half-away: the flat-wall coordinates avoid exact halves
                un, vn = int(np.floor(u + F32(0.5))), int(np.floor(v + F32(0.5)))
            sdf = F32(rng - pz)
            if sdf < -trunc or not abs(sdf) < trunc:
                continue
            wm = F32(F32(F32(cam.fx) * F32(cam.fy)) * F32(vs * vs) / F32(pz * pz))
            wm = F32(wm / F32(pz * pz))
            if sdf < -eps:
                wm = F32(max(F32(wm * F32(F32(trunc + sdf) / F32(trunc - eps))), F32(0)))
            w_old = F32(0)
            c = np.zeros(3, np.uint8)
            for img in cols:
                if bil:
                    taps = [img[v0, u0], img[v0 + 1, u0], img[v0, u0 + 1], img[v0 + 1, u0 + 1]]
                    cm = np.zeros(3, np.uint8)
                    for ch in range(3):
                        sc = F32(F32(F32(w4[0] * F32(taps[0][ch])) + F32(w4[1] * F32(taps[1][ch]))) + F32(w4[2] * F32(taps[2][ch])))
                        sc = F32(sc + F32(w4[3] * F32(taps[3][ch])))
                        cm[ch] = int(sc)
                else:
                    cm = img[vn, un]
                ratio = F32(wm / F32(w_old + wm))
                for ch in range(3):
                    c[ch] = int(F32(F32(F32(c[ch]) * F32(F32(1) - ratio)) + F32(F32(cm[ch]) * ratio)))
                w_old = F32(w_old + wm)
            expect[j] = c
            checked += 1
        np.testing.assert_array_equal(b.color[i], expect, err_msg=f"block {bidx}")
    assert checked > 300


def test_colour_is_left_alone_without_colour_image(oracle_lib):
    cam = syn.make_camera(64, 48, 32.0, 32.0, max_range=3.0)
    mc = capi.default_map_config(voxel_size=0.1, vps=8, trunc=0.3)
    h = hs.make_handle(oracle_lib, "ko_", map_cfg=mc, cam=cam)
    d, l = flat_wall(cam, 2.0)
    c = syn.colorize(l, d)
    h.integrate_frame(h.make_frame(d, np.eye(4), 1_000_000_000, label=l, color=c))
    before = h.export_blocks().color.copy()
    h.integrate_frame(h.make_frame(d, np.eye(4), 1_000_000_001, label=l))  # no colour image: TSDF only
    after = h.export_blocks()
    np.testing.assert_array_equal(before, after.color)
    assert before.any() and (after.weight > 0).any()
