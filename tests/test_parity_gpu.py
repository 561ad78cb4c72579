"""GPU parity tests proper: CUDA product (through the C ABI) vs the CPU oracle on the same seeded
synthetic inputs. Integer outputs (labels, stamps, flags, block sets, dynamic image) must be bit-exact;
TSDF distance / weight within 1e-4 relative (BASELINE.json north_star) — in practice they are also
bit-exact because both sides evaluate the same fp32 expression order without FMA contraction."""
import numpy as np
import pytest

import khronos_b200 as kb
from khronos_b200 import capi, synthetic as syn
import harness as hs

pytestmark = pytest.mark.gpu


def both(oracle_lib, product_lib, **kw):
    return hs.make_handle(oracle_lib, "ko_", **kw), hs.make_handle(product_lib, "kb_", **kw)


def room_frames(cam, n, laps=0.25, dt_ns=33_333_333, scene=None):
    scene = scene or syn.room_scene()
    poses, stamps = syn.orbit_trajectory(n, laps=laps, dt_ns=dt_ns)
    return hs.render_frames(scene, cam, poses, stamps), poses, stamps


@pytest.mark.parametrize("interp", [capi.INTERP_ADAPTIVE, capi.INTERP_NEAREST, capi.INTERP_BILINEAR])
def test_fusion_small_stream_bit_exact(oracle_lib, product_lib, interp):
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 10)
    cfg = dict(cam=cam, integ_cfg=capi.default_integrator_config(interpolation=interp))
    o, g = both(oracle_lib, product_lib, **cfg)
    so = hs.run_fusion(o, frames, poses, stamps)
    sg = hs.run_fusion(g, frames, poses, stamps)
    assert so == sg
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what=f"interp{interp}")


def test_fusion_full_resolution_frame(oracle_lib, product_lib):
    """BASELINE config 1/2 shape: 640x480 into a 5 cm / 16^3 map."""
    cam = syn.make_camera()
    frames, poses, stamps = room_frames(cam, 3, laps=0.02)
    o, g = both(oracle_lib, product_lib, cam=cam)
    so = hs.run_fusion(o, frames, poses, stamps)
    sg = hs.run_fusion(g, frames, poses, stamps)
    assert so == sg
    assert so[0]["voxels_updated"] > 100000
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="fullres")


def test_fusion_weight_options_and_blocked_labels(oracle_lib, product_lib):
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 6)
    ic = capi.default_integrator_config(blocked=(4, 9))
    ic.use_constant_weight = 1
    ic.use_weight_dropoff = 0
    ic.max_weight = 50.0
    o, g = both(oracle_lib, product_lib, cam=cam, integ_cfg=ic)
    assert hs.run_fusion(o, frames, poses, stamps) == hs.run_fusion(g, frames, poses, stamps)
    bo, bg = o.export_blocks(), g.export_blocks()
    hs.assert_blocks_equal(bo, bg, exact_float=True, what="options")
    assert not np.isin(bo.semantic_label[bo.semantic_empty == 0], (4, 9)).any()


def test_integration_mask(oracle_lib, product_lib):
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 4)
    rng = np.random.default_rng(3)
    masks = []
    for _ in frames:
        m = np.zeros((cam.height, cam.width), np.int32)
        m[20:70, 30:90] = rng.integers(0, 3, size=(50, 60))
        masks.append(m)
    o, g = both(oracle_lib, product_lib, cam=cam)
    assert hs.run_fusion(o, frames, poses, stamps, masks=masks) == hs.run_fusion(g, frames, poses, stamps, masks=masks)
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="mask")


def test_no_semantics_no_tracking(oracle_lib, product_lib):
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 4)
    mc = capi.default_map_config(with_semantics=False, with_tracking=False)
    ic = capi.default_integrator_config(semantic_mode=capi.SEM_NONE)
    o, g = both(oracle_lib, product_lib, cam=cam, map_cfg=mc, integ_cfg=ic)
    assert hs.run_fusion(o, frames, poses, stamps) == hs.run_fusion(g, frames, poses, stamps)
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="plain")


def dynamic_scenario(cam):
    """Static burn-in (> temporal_buffer) so free space becomes ever-free, then a cuboid moves through it."""
    scene = syn.room_scene()
    scene.mover = ((0.5, 0.5, 1.2), (8.6, 1.5, 0.9), (0.0, 2.0, 0.0), 2.0)
    n, dt = 30, 150_000_000
    pose = syn.look_pose((6.0, 5.0, 1.5), 0.0, np.radians(10.0))
    poses = [pose] * n
    stamps = [1_000_000_000 + i * dt for i in range(n)]
    # the mover waits outside the view cone, then crosses the (by then ever-free) space in front of
    # the camera from t = 2 s on.
    return scene, hs.render_frames(scene, cam, poses, stamps), poses, stamps


@pytest.mark.parametrize("sep", [2.0, 1.0, 0.0])
def test_tracking_everfree_motion(oracle_lib, product_lib, sep):
    """Per-frame pipeline detect -> integrate(mask) -> track. sep > 0 runs M2-M4 on the device (connected
    components; D = ceil(sep) = 2 merges near clusters, 1 only shared voxels), sep = 0 the host path. The
    product integrates with the device-resident dynamic image (KB_MASK_LAST_DETECTION)."""
    cam = hs.small_camera(4)
    scene, frames, poses, stamps = dynamic_scenario(cam)
    mot = capi.default_motion_config(min_cluster_size=5, min_separation_distance=sep)
    o, g = both(oracle_lib, product_lib, cam=cam, mot_cfg=mot)
    g2 = hs.make_handle(product_lib, "kb_", cam=cam, mot_cfg=mot)  # driven through the fused kb_spin_once
    total_dyn = 0
    for i, ((d, l), T, st) in enumerate(zip(frames, poses, stamps)):
        fo, fg = o.make_frame(d, T, st, label=l), g.make_frame(d, T, st, label=l)
        io, so_, co = o.detect_motion(fo)
        ig, sg_, cg = g.detect_motion(fg)
        assert (so_, co) == (sg_, cg), f"frame {i}: seeds/clusters {so_, co} vs {sg_, cg}"
        np.testing.assert_array_equal(io, ig, err_msg=f"dynamic_image frame {i}")
        i2, s2, c2 = g2.spin_once(g2.make_frame(d, T, st, label=l))
        assert (s2, c2) == (so_, co), f"spin_once frame {i}"
        np.testing.assert_array_equal(io, i2, err_msg=f"spin_once dynamic_image frame {i}")
        clo, clg = o.get_motion_clusters(), g.get_motion_clusters()
        assert len(clo) == len(clg)
        for a, b in zip(clo, clg):
            np.testing.assert_array_equal(a["voxels"], b["voxels"])
            pa = a["pixels"][np.lexsort(a["pixels"].T)]
            pb = b["pixels"][np.lexsort(b["pixels"].T)]
            np.testing.assert_array_equal(pa, pb)
            np.testing.assert_array_equal(a["bbox"], b["bbox"])
        total_dyn += int((io > 0).sum())
        fo2 = o.make_frame(d, T, st, label=l, mask=io)
        fg2 = g.make_frame(d, T, st, label=l, mask=capi.MASK_LAST_DETECTION if i % 2 else ig)
        assert o.integrate_frame(fo2).as_dict() == g.integrate_frame(fg2).as_dict()
        o.update_tracking(st)
        g.update_tracking(st)
    bo, bg = o.export_blocks(), g.export_blocks()
    hs.assert_blocks_equal(bo, bg, exact_float=True, what="dynamic")
    hs.assert_blocks_equal(bo, g2.export_blocks(), exact_float=True, what="dynamic via kb_spin_once")
    assert bo.ever_free.sum() > 1000, "scenario must produce ever-free space"
    assert total_dyn > 100, "scenario must flag dynamic pixels"


def test_reset_inactive_and_reuse(oracle_lib, product_lib):
    cam = hs.small_camera(4)
    # 0.5 s between frames, quarter orbit: early blocks leave the 3 s temporal window
    frames, poses, stamps = room_frames(cam, 14, laps=0.5, dt_ns=500_000_000)
    o, g = both(oracle_lib, product_lib, cam=cam)
    removed_total = 0
    for i, ((d, l), T, st) in enumerate(zip(frames, poses, stamps)):
        for h in (o, g):
            h.integrate_frame(h.make_frame(d, T, st, label=l))
            h.update_tracking(st)
        if i % 4 == 3:
            ro, rg = o.reset_inactive(), g.reset_inactive()
            np.testing.assert_array_equal(ro, rg)
            removed_total += len(ro)
            o.clear_updated(); g.clear_updated()
            hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what=f"after reset {i}")
    assert removed_total > 0
    hs.assert_blocks_equal(o.export_blocks(capi.EXPORT_UPDATED), g.export_blocks(capi.EXPORT_UPDATED), exact_float=True, what="updated-only")
    for h in (o, g):
        h.mark_all_inactive()
    np.testing.assert_array_equal(o.reset_inactive(), g.reset_inactive())
    assert o.num_blocks() == g.num_blocks() == 0


def test_object_extraction_path(oracle_lib, product_lib):
    """K1b + E0 + K4: private vps=8 binary-semantics map, pre-allocated box, integrate without
    allocation, low-confidence erase (mesh_object_extractor.cpp:201-264)."""
    cam = hs.small_camera(2)
    scene = syn.room_scene()
    # 12 views orbiting towards the first cuboid (label 7, [2,2,0]..[3,3.5,1.2])
    angles = np.linspace(3.45, 3.95, 12)
    poses = [syn.look_pose((6.0 + 2.5 * np.cos(a), 5.0 + 2.5 * np.sin(a), 1.5), a, np.radians(10.0)) for a in angles]
    stamps = [1_000_000_000 + i * 33_333_333 for i in range(12)]
    frames = hs.render_frames(scene, cam, poses, stamps)
    target = 7  # first cuboid's label plays the role of the object id in object_image
    voxel = 0.04
    mc = capi.default_map_config(voxel_size=voxel, vps=8, trunc=2 * voxel, with_semantics=True, with_tracking=False,
                                 max_blocks=8192)
    ic = capi.default_integrator_config(semantic_mode=capi.SEM_BINARY)
    o, g = both(oracle_lib, product_lib, cam=cam, map_cfg=mc, integ_cfg=ic)
    bs = voxel * 8
    # the first cuboid (label 7): [2,2,0]..[3,3.5,1.2]; sits in view of the first poses? choose by data
    lo = np.floor(np.array([1.5, 1.5, -0.3]) / bs).astype(int)
    hi = np.floor(np.array([3.5, 4.0, 1.7]) / bs).astype(int)
    seen = 0
    for h in (o, g):
        h.allocate_box(lo, hi)
    assert o.num_blocks() == g.num_blocks() == int(np.prod(hi - lo + 1))
    for (d, l), T, st in zip(frames, poses, stamps):
        seen += int((l == target).sum())
        so = o.integrate_frame(o.make_frame(d, T, st, object_image=l, target_id=target), allocate_blocks=False)
        sg = g.integrate_frame(g.make_frame(d, T, st, object_image=l, target_id=target), allocate_blocks=False)
        assert so.as_dict() == sg.as_dict()
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="object pre-scan")
    assert seen > 1000, "the target object must be visible"
    eo, eg = o.scan_object_confidence(0.5, 3), g.scan_object_confidence(0.5, 3)
    assert eo == eg and eo > 0
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="object post-scan")


@pytest.mark.parametrize("nshards", [2, 4])
def test_block_hash_sharding_is_exact(oracle_lib, product_lib, nshards):
    """SURVEY §4 "fake-shard" mode: S shard handles on one GPU; their union equals the unsharded map."""
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 6)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    hs.run_fusion(o, frames, poses, stamps)
    bo = o.export_blocks()
    parts = []
    for r in range(nshards):
        g = hs.make_handle(product_lib, "kb_", cam=cam)
        g.set_shard(r, nshards)
        hs.run_fusion(g, frames, poses, stamps)
        b = g.export_blocks()
        for idx in b.block_index:
            assert product_lib.kb_block_owner(int(idx[0]), int(idx[1]), int(idx[2]), nshards) == r
        parts.append(b)
    assert sum(p.n for p in parts) == bo.n
    order = np.lexsort(np.concatenate([p.block_index for p in parts])[:, ::-1].T)
    def cat(name):
        return np.concatenate([getattr(p, name) for p in parts])[order]
    np.testing.assert_array_equal(cat("block_index"), bo.block_index)
    for name in ("distance", "weight"):
        np.testing.assert_array_equal(cat(name).view(np.uint32), getattr(bo, name).view(np.uint32))
    for name in ("last_observed", "semantic_label", "semantic_empty", "block_flags"):
        np.testing.assert_array_equal(cat(name), getattr(bo, name))


def test_device_resident_frames(oracle_lib, product_lib):
    import torch
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 5)
    o, g = both(oracle_lib, product_lib, cam=cam)
    hs.run_fusion(o, frames, poses, stamps)
    for (d, l), T, st in zip(frames, poses, stamps):
        dd, ll = torch.from_numpy(d).cuda(), torch.from_numpy(l).cuda()
        torch.cuda.synchronize()
        g.integrate_frame(g.make_frame(dd, T, st, label=ll, memory=capi.MEM_DEVICE), want_stats=False)
    g.synchronize()
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="device frames")


def test_capacity_error_is_reported(product_lib):
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 1)
    g = hs.make_handle(product_lib, "kb_", cam=cam, map_cfg=capi.default_map_config(max_blocks=16))
    with pytest.raises(kb.KbError) as e:
        hs.run_fusion(g, frames, poses, stamps)
    assert e.value.status == 3


@pytest.mark.parametrize("batch", [3, 32, 37])
def test_batched_integration_equals_sequential(oracle_lib, product_lib, batch):
    """kb_integrate_frames fuses up to 32 frames per launch with voxel state in registers; the result
    must be identical to frame-by-frame integration (oracle runs strictly sequentially)."""
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 40, laps=0.4)
    o, g = both(oracle_lib, product_lib, cam=cam)
    so = hs.run_fusion(o, frames, poses, stamps)
    tot = {k: 0 for k in so[0]}
    for i in range(0, len(frames), batch):
        fr = [g.make_frame(d, T, st, label=l) for (d, l), T, st in zip(frames[i:i + batch], poses[i:i + batch], stamps[i:i + batch])]
        s = g.integrate_frames(fr).as_dict()
        want = {k: sum(x[k] for x in so[i:i + batch]) for k in s}
        want["total_blocks"] = so[min(i + batch, len(so)) - 1]["total_blocks"]
        assert s == want, (i, s, want)
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what=f"batch{batch}")


def test_culling_does_not_change_results(oracle_lib, product_lib):
    """The conservative (block, frame) depth culling may only skip work that could not have produced a
    valid measurement: maps with culling on and off are identical (hall scene: most blocks cull)."""
    cam = hs.small_camera(4)
    scene = syn.hall_scene(size=(20.0, 16.0, 6.0))
    poses, stamps = syn.sweep_trajectory(24, size=(20.0, 16.0), margin=4.0, lanes=2, yaw_turns=1.5)
    frames = hs.render_frames(scene, cam, poses, stamps)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    so = hs.run_fusion(o, frames, poses, stamps)
    bo = o.export_blocks()
    for cull in (2, 0):  # 2 = cull even single-frame calls, 0 = never
        g = hs.make_handle(product_lib, "kb_", cam=cam)
        g.set_culling(cull)
        sg = hs.run_fusion(g, frames, poses, stamps)
        assert sg == so
        hs.assert_blocks_equal(bo, g.export_blocks(), exact_float=True, what=f"cull={cull}")


def test_pinned_host_frames_overlap_path(oracle_lib, product_lib):
    """KB_MEM_HOST frames go through the double-buffered copy stream; many back-to-back calls without
    stats must still integrate every frame exactly once and in order."""
    import torch
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 70, laps=0.5)
    o, g = both(oracle_lib, product_lib, cam=cam)
    hs.run_fusion(o, frames, poses, stamps)
    hd = torch.from_numpy(np.stack([f[0] for f in frames])).pin_memory()
    hl = torch.from_numpy(np.stack([f[1] for f in frames])).pin_memory()
    fr = [g.make_frame(hd[i].data_ptr(), poses[i], stamps[i], label=hl[i].data_ptr(),
                       memory=capi.MEM_HOST if i < 40 else capi.MEM_HOST_ASYNC) for i in range(len(frames))]
    for i in range(0, 20):
        g.integrate_frame(fr[i], want_stats=False)       # borrowed per call
    g.integrate_frames(fr[20:40], want_stats=False)      # contiguous run -> coalesced copies
    g.integrate_frames(fr[40:], want_stats=False)        # caller-kept pinned buffers, copies not awaited
    g.synchronize()
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="pinned")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_lazy_tracking_random_schedules(oracle_lib, product_lib, seed):
    """The product keeps tracking state lazily (O(blocks) per pass instead of the reference's all-voxel
    rewrite). Irregular schedules — several integrations between passes, passes without integration, time
    jumps beyond the temporal window, re-observation after deactivation, block removal and re-allocation —
    must still give exactly the brute-force oracle's last_occupied / active / to_remove / ever_free /
    has_active_data."""
    rng = np.random.default_rng(100 + seed)
    cam = hs.small_camera(8)
    scene = syn.room_scene()
    n = 36
    poses, _ = syn.orbit_trajectory(n, laps=0.6)
    frames = hs.render_frames(scene, cam, poses, [0] * n)
    o, g = both(oracle_lib, product_lib, cam=cam)
    t = 500_000_000  # start below temporal_window so the "stamp 0 is active" quirk is exercised
    i = 0
    checks = 0
    while i < n:
        k = int(rng.integers(1, 4))  # integrate 1..3 frames
        batch = []
        for _ in range(k):
            if i >= n:
                break
            t += int(rng.choice([40_000_000, 150_000_000, 600_000_000]))
            batch.append((i, t))
            i += 1
        for h in (o, g):
            fr = [h.make_frame(frames[j][0], poses[j], st, label=frames[j][1]) for j, st in batch]
            if len(fr) > 1 and rng.integers(0, 2) == 0 and h is g:
                h.integrate_frames(fr, want_stats=False)
            else:
                for f in fr:
                    h.integrate_frame(f, want_stats=False)
        ev = rng.integers(0, 10)
        if ev < 7:          # normal paired pass
            for h in (o, g):
                h.update_tracking(t)
        elif ev == 7:       # pass after a long silence (> temporal_window), then another one right after
            t += 3_500_000_000
            for h in (o, g):
                h.update_tracking(t)
            t += 10_000_000
            for h in (o, g):
                h.update_tracking(t)
        # ev 8, 9: no pass this round
        if rng.integers(0, 4) == 0:
            ro, rg = o.reset_inactive(), g.reset_inactive()
            np.testing.assert_array_equal(ro, rg)
        if rng.integers(0, 3) == 0:
            hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what=f"seed{seed} frame{i}")
            checks += 1
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what=f"seed{seed} final")
    for h in (o, g):
        h.mark_all_inactive()
    bo, bg = o.export_blocks(), g.export_blocks()
    np.testing.assert_array_equal(bo.block_flags, bg.block_flags)
    np.testing.assert_array_equal(o.reset_inactive(), g.reset_inactive())
    assert checks > 2


def test_config4_shape_1280x720_2cm(oracle_lib, product_lib):
    """BASELINE config[3] shapes: 1280x720, fx=fy=640, 2 cm voxels (0.32 m blocks), trunc 0.06 — range limited
    to 2 m so the CPU oracle and the export stay small."""
    cam = syn.make_camera(1280, 720, 640.0, 640.0, max_range=2.0)
    scene = syn.room_scene()
    poses = [syn.look_pose((3.6, 3.4, 1.2), 3.9 + 0.02 * i, np.radians(12.0)) for i in range(2)]
    stamps = [1_000_000_000 + i * 33_333_333 for i in range(2)]
    frames = hs.render_frames(scene, cam, poses, stamps)
    mc = capi.default_map_config(voxel_size=0.02, vps=16, trunc=0.06, max_blocks=8192)
    o, g = both(oracle_lib, product_lib, cam=cam, map_cfg=mc)
    so = hs.run_fusion(o, frames, poses, stamps)
    fr = [g.make_frame(d, T, st, label=l) for (d, l), T, st in zip(frames, poses, stamps)]
    sg = g.integrate_frames(fr).as_dict()
    assert sg["voxels_updated"] == sum(s["voxels_updated"] for s in so) > 100000
    assert sg["blocks_in_frustum"] == sum(s["blocks_in_frustum"] for s in so)
    bo, bg = o.export_blocks(likelihoods=False), g.export_blocks(likelihoods=False)
    hs.assert_blocks_equal(bo, bg, exact_float=True, what="config4")


def test_config5_tesse_shapes_pipeline(oracle_lib, product_lib):
    """BASELINE config[4] shapes: 720x480, fx=fy=415.692, cx=360, cy=240 (khronos_eval/config/ground_truth/
    tesse_cd_office_dynamic_objects.yaml:18-21), voxel 0.1 / trunc 0.2 / 16^3 (uHumans2.yaml:45-49), motion
    detector min_cluster 500 px / separation 2 / 26-nbr (uHumans2.yaml:53-56): detect -> integrate -> track per
    frame, reset_inactive at the output cadence, then the object-extraction path (vps 8, binary) on the same frames."""
    cam = syn.make_camera(720, 480, 415.692, 415.692, cx=360.0, cy=240.0)
    scene = syn.room_scene()
    scene.mover = ((0.7, 0.7, 1.4), (8.4, 1.2, 0.9), (0.0, 2.5, 0.0), 1.6)
    n, dt = 14, 200_000_000
    pose = syn.look_pose((6.0, 5.0, 1.5), 0.0, np.radians(10.0))
    poses, stamps = [pose] * n, [1_000_000_000 + i * dt for i in range(n)]
    frames = hs.render_frames(scene, cam, poses, stamps)
    mc = capi.default_map_config(voxel_size=0.1, vps=16, trunc=0.2)
    mot = capi.default_motion_config(min_cluster_size=500, min_separation_distance=2.0)
    o, g = both(oracle_lib, product_lib, cam=cam, map_cfg=mc, mot_cfg=mot)
    flagged = 0
    for i, ((d, l), T, st) in enumerate(zip(frames, poses, stamps)):
        io, so_, co = o.detect_motion(o.make_frame(d, T, st, label=l))
        ig, sg_, cg = g.detect_motion(g.make_frame(d, T, st, label=l))
        assert (so_, co) == (sg_, cg)
        np.testing.assert_array_equal(io, ig)
        flagged += int((io > 0).sum())
        o.integrate_frame(o.make_frame(d, T, st, label=l, mask=io), want_stats=False)
        g.integrate_frame(g.make_frame(d, T, st, label=l, mask=capi.MASK_LAST_DETECTION), want_stats=False)
        o.update_tracking(st)
        g.update_tracking(st)
        if i % 2 == 1:  # min_output_separation 0.4 s (uHumans2.yaml:38)
            np.testing.assert_array_equal(o.reset_inactive(), g.reset_inactive())
            o.clear_updated(); g.clear_updated()
    assert flagged > 500
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="config5 window")
    # object extraction on the same frames: the second cuboid (label 8) as the tracked object
    target = 8
    voxel = 0.05
    emc = capi.default_map_config(voxel_size=voxel, vps=8, trunc=2 * voxel, with_tracking=False, max_blocks=16384)
    eic = capi.default_integrator_config(semantic_mode=capi.SEM_BINARY)
    eo, eg = both(oracle_lib, product_lib, cam=cam, map_cfg=emc, integ_cfg=eic)
    lo = np.floor(np.array([8.0, 1.0, -0.3]) / (voxel * 8)).astype(int)
    hi = np.floor(np.array([10.5, 3.0, 1.4]) / (voxel * 8)).astype(int)
    for h in (eo, eg):
        h.allocate_box(lo, hi)
    for (d, l), T, st in zip(frames, poses, stamps):
        eo.integrate_frame(eo.make_frame(d, T, st, object_image=l, target_id=target), allocate_blocks=False, want_stats=False)
    eg.integrate_frames([eg.make_frame(d, T, st, object_image=l, target_id=target) for (d, l), T, st in zip(frames, poses, stamps)],
                        allocate_blocks=False, want_stats=False)
    assert eo.scan_object_confidence(0.5, 10) == eg.scan_object_confidence(0.5, 10) > 0
    hs.assert_blocks_equal(eo.export_blocks(), eg.export_blocks(), exact_float=True, what="config5 object")


def test_compact_wire_format_u16_depth_u8_label(oracle_lib, product_lib):
    """kb_frame.depth_u16 / label_u8: the device expands 16-bit millimetre depth and 8-bit labels exactly like the
    host-side conversion (float(u16) * scale, int32(u8)); host, pinned-async and device-resident inputs agree with
    the oracle fed the same compact frames and with the oracle fed the pre-expanded f32 / i32 images."""
    import torch
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 12, laps=0.2)
    d16 = [np.round(d * 1000.0).astype(np.uint16) for d, _ in frames]
    l8 = [l.astype(np.uint8) for _, l in frames]
    scale = np.float32(0.001)
    o, g = both(oracle_lib, product_lib, cam=cam)
    o2 = hs.make_handle(oracle_lib, "ko_", cam=cam)
    for i in range(len(frames)):
        o.integrate_frame(o.make_frame(None, poses[i], stamps[i], depth_u16=d16[i], label_u8=l8[i]), want_stats=False)
        o2.integrate_frame(o2.make_frame((d16[i].astype(np.float32) * scale), poses[i], stamps[i], label=l8[i].astype(np.int32)), want_stats=False)
    hs.assert_blocks_equal(o2.export_blocks(), o.export_blocks(), exact_float=True, what="oracle compact vs expanded")
    # product: 4 host frames, 4 device-resident frames, 4 pinned-async frames in one batch call
    for i in range(4):
        g.integrate_frame(g.make_frame(None, poses[i], stamps[i], depth_u16=d16[i], label_u8=l8[i]), want_stats=False)
    dd = [torch.from_numpy(x).cuda() for x in d16[4:8]]
    ll = [torch.from_numpy(x).cuda() for x in l8[4:8]]
    torch.cuda.synchronize()
    g.integrate_frames([g.make_frame(None, poses[4 + j], stamps[4 + j], depth_u16=dd[j], label_u8=ll[j], memory=capi.MEM_DEVICE)
                        for j in range(4)], want_stats=False)
    pd = torch.from_numpy(np.stack(d16[8:])).pin_memory()
    pl = torch.from_numpy(np.stack(l8[8:])).pin_memory()
    g.integrate_frames([g.make_frame(None, poses[8 + j], stamps[8 + j], depth_u16=pd[j].data_ptr(), label_u8=pl[j].data_ptr(),
                                     memory=capi.MEM_HOST_ASYNC) for j in range(4)], want_stats=False)
    g.synchronize()
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="compact wire format")
    # mixed batch: compact and pre-expanded frames interleaved in one call (compact ones are expanded on the device)
    g2 = hs.make_handle(product_lib, "kb_", cam=cam)
    mixed = []
    for i in range(len(frames)):
        if i % 2:
            mixed.append(g2.make_frame(None, poses[i], stamps[i], depth_u16=d16[i], label_u8=l8[i]))
        else:
            mixed.append(g2.make_frame(d16[i].astype(np.float32) * scale, poses[i], stamps[i], label=l8[i].astype(np.int32)))
    g2.integrate_frames(mixed, want_stats=False)
    hs.assert_blocks_equal(o.export_blocks(), g2.export_blocks(), exact_float=True, what="mixed compact / f32 batch")
