"""Edge cases of the fusion path (empty / degenerate / ragged inputs, odd image sizes, far-away coordinates, sensor
garbage) — product vs oracle through the C ABI, bit-exact like the main parity suite."""
import numpy as np
import pytest

from khronos_b200 import capi, synthetic as syn
import harness as hs
from test_parity_gpu import both, room_frames

pytestmark = pytest.mark.gpu


def run_both(o, g, frames, poses, stamps, batch=None, **kw):
    so = hs.run_fusion(o, frames, poses, stamps, **kw)
    if batch:
        for i in range(0, len(frames), batch):
            g.integrate_frames([g.make_frame(d, T, st, label=l) for (d, l), T, st in
                                zip(frames[i:i + batch], poses[i:i + batch], stamps[i:i + batch])])
            if kw.get("tracking"):
                g.update_tracking(stamps[min(i + batch, len(frames)) - 1])
    else:
        assert so == hs.run_fusion(g, frames, poses, stamps, **kw)


def test_empty_and_degenerate_frames(oracle_lib, product_lib):
    """All-invalid depth (allocation only), a frame of zeros between real frames, a single valid pixel."""
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 5)
    z = np.zeros_like(frames[0][0])
    one = z.copy()
    one[60, 80] = 2.0
    frames = [(z, frames[0][1]), frames[1], (z, frames[2][1]), (one, frames[3][1]), frames[4]]
    o, g = both(oracle_lib, product_lib, cam=cam)
    run_both(o, g, frames, poses, stamps, tracking=True)
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="degenerate frames")


def test_labels_outside_the_label_space(oracle_lib, product_lib):
    """Negative labels, labels >= L and >= KB_MAX_LABELS: TSDF is fused, semantics are not (isValidLabel)."""
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 4)
    rng = np.random.default_rng(1)
    bad = np.array([-1, -7, 20, 21, 63, 64, 65, 1000, 2 ** 31 - 1, -2 ** 31], np.int32)
    frames = [(d, np.where(rng.random(l.shape) < 0.3, rng.choice(bad, size=l.shape), l).astype(np.int32)) for d, l in frames]
    o, g = both(oracle_lib, product_lib, cam=cam, integ_cfg=capi.default_integrator_config(blocked=(3,), num_threads=4))
    run_both(o, g, frames, poses, stamps)
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="bad labels")


def test_sensor_garbage_nan_inf_negative_depth(oracle_lib, product_lib):
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 4)
    rng = np.random.default_rng(2)
    out = []
    for d, l in frames:
        d = d.copy()
        r = rng.random(d.shape)
        d[r < 0.02] = np.nan
        d[(r >= 0.02) & (r < 0.04)] = np.inf
        d[(r >= 0.04) & (r < 0.06)] = -1.5
        d[(r >= 0.06) & (r < 0.07)] = 1e-30
        out.append((d, l))
    o, g = both(oracle_lib, product_lib, cam=cam)
    g.set_culling(2)
    run_both(o, g, out, poses, stamps)
    bo, bg = o.export_blocks(), g.export_blocks()
    np.testing.assert_array_equal(np.isnan(bo.distance), np.isnan(bg.distance))
    hs.assert_blocks_equal(bo, bg, exact_float=True, what="garbage depth")


@pytest.mark.parametrize("size", [(163, 117), (97, 61), (33, 17)])
def test_odd_image_sizes_with_culling(oracle_lib, product_lib, size):
    """Tile pyramid / culling tails: widths and heights that are not multiples of 8, 16, 32 or 64."""
    W, H = size
    cam = syn.make_camera(W, H, W / 2.0, W / 2.0, max_range=5.0)
    frames, poses, stamps = room_frames(cam, 12, laps=0.2)
    o, g = both(oracle_lib, product_lib, cam=cam)
    g.set_culling(2)
    run_both(o, g, frames, poses, stamps, batch=12)
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what=f"{W}x{H}")


def test_far_from_the_origin_and_negative_block_indices(oracle_lib, product_lib):
    """The same room observed from a world frame shifted by kilometres (float precision is what it is on both
    sides) and into negative coordinates: block keys far from zero, hash distribution unchanged."""
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 6)
    for shift in ((-3000.0, 5000.0, -200.0), (-12.34, -56.78, -9.0)):
        sp = []
        for T in poses:
            T = np.array(T, dtype=np.float64)
            T[:3, 3] += np.array(shift)
            sp.append(T)
        o, g = both(oracle_lib, product_lib, cam=cam)
        run_both(o, g, frames, sp, stamps, tracking=True)
        bo, bg = o.export_blocks(), g.export_blocks()
        assert (bo.block_index < 0).any()
        hs.assert_blocks_equal(bo, bg, exact_float=True, what=f"shift {shift}")


def test_vps8_mle_tracking_batched(oracle_lib, product_lib):
    """8^3 blocks with MLE semantics and tracking (the extractor uses 8^3 with binary semantics only)."""
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 20, laps=0.3)
    mc = capi.default_map_config(voxel_size=0.1, vps=8, trunc=0.3, max_blocks=16384)
    o, g = both(oracle_lib, product_lib, cam=cam, map_cfg=mc)
    g.set_culling(2)
    hs.run_fusion(o, frames, poses, stamps, tracking=False)
    for i in range(0, 20, 10):
        g.integrate_frames([g.make_frame(d, T, st, label=l) for (d, l), T, st in zip(frames[i:i + 10], poses[i:i + 10], stamps[i:i + 10])])
    o.update_tracking(stamps[-1])
    g.update_tracking(stamps[-1])
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="vps8 mle")


def test_repeated_stamps_and_identical_frames(oracle_lib, product_lib):
    """The same frame integrated 40 times in one batch call (weights saturate towards max_weight = 50)."""
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 1)
    ic = capi.default_integrator_config(num_threads=4)
    ic.max_weight = 50.0
    ic.use_constant_weight = 1
    o, g = both(oracle_lib, product_lib, cam=cam, integ_cfg=ic)
    n = 40
    st = [stamps[0] + k for k in range(n)]
    hs.run_fusion(o, frames * n, poses * n, st)
    d, l = frames[0]
    g.integrate_frames([g.make_frame(d, poses[0], s, label=l) for s in st])
    bo = o.export_blocks()
    assert (bo.weight == 50.0).any()
    hs.assert_blocks_equal(bo, g.export_blocks(), exact_float=True, what="saturation")


def test_f32_depth_with_u8_labels_in_one_frame(oracle_lib, product_lib):
    """kb_frame.depth (f32) together with kb_frame.label_u8: the lossless 5 B/pixel wire of bench.py --wire f32u8.
    Host frames one by one and device-resident frames in a batch."""
    import torch
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 12, laps=0.2)
    l8 = [l.astype(np.uint8) for _, l in frames]
    o, g = both(oracle_lib, product_lib, cam=cam)
    hs.run_fusion(o, frames, poses, stamps)
    for i in range(4):
        g.integrate_frame(g.make_frame(frames[i][0], poses[i], stamps[i], label_u8=l8[i]), want_stats=False)
    dd = [torch.from_numpy(frames[i][0]).cuda() for i in range(4, 12)]
    ll = [torch.from_numpy(l8[i]).cuda() for i in range(4, 12)]
    torch.cuda.synchronize()
    g.integrate_frames([g.make_frame(dd[j], poses[4 + j], stamps[4 + j], label_u8=ll[j], memory=capi.MEM_DEVICE) for j in range(8)],
                       want_stats=False)
    g.synchronize()
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="f32 depth + u8 labels")


def test_hash_rebuild_after_block_removal(oracle_lib, product_lib):
    """Tombstone garbage collection: with KB_REHASH_TOMBSTONES=1 every kb_reset_inactive that removes a block rebuilds
    the block hash from the live slots; lookups, re-allocation and every result must be unaffected."""
    import os
    os.environ["KB_REHASH_TOMBSTONES"] = "1"
    try:
        cam = hs.small_camera(4)
        frames, poses, stamps = room_frames(cam, 14, laps=0.5, dt_ns=500_000_000)
        frames, poses = frames + frames[:6], poses + poses[:6]   # come back: removed blocks are allocated again
        stamps = stamps + [stamps[-1] + (k + 1) * 500_000_000 for k in range(6)]
        mot = capi.default_motion_config(min_cluster_size=5, min_separation_distance=2.0, num_threads=4)
        o, g = both(oracle_lib, product_lib, cam=cam, mot_cfg=mot)
        removed = 0
        for i, ((d, l), T, st) in enumerate(zip(frames, poses, stamps)):
            io, so, co = o.spin_once(o.make_frame(d, T, st, label=l))
            ig, sg, cg = g.spin_once(g.make_frame(d, T, st, label=l))
            assert (so, co) == (sg, cg)
            np.testing.assert_array_equal(io, ig)
            if i % 3 == 2:
                ro, rg = o.reset_inactive(), g.reset_inactive()
                np.testing.assert_array_equal(ro, rg)
                removed += len(ro)
                hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what=f"after reset {i}")
        assert removed > 0
        assert int(g.get_debug_counters(32)[28]) >= 1, "the hash must have been rebuilt"
        hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="hash rebuild")
    finally:
        os.environ.pop("KB_REHASH_TOMBSTONES", None)


def test_batch_with_frames_far_apart_is_split(oracle_lib, product_lib):
    """A batch call whose frames jump hundreds of metres (the union candidate box of K0 would have ~10^8 cells) is cut
    into spatially coherent sub-batches; results equal frame-by-frame integration."""
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 12, laps=0.2)
    far = []
    for i, T in enumerate(poses):
        T = np.array(T, dtype=np.float64)
        if i % 4 >= 2:
            T[:3, 3] += np.array([300.0 * (i % 3 + 1), -450.0, 20.0])   # outside the room: only allocation happens there
        far.append(T)
    o, g = both(oracle_lib, product_lib, cam=cam)
    g.set_culling(2)
    so = hs.run_fusion(o, frames, far, stamps)
    s = g.integrate_frames([g.make_frame(d, T, st, label=l) for (d, l), T, st in zip(frames, far, stamps)]).as_dict()
    want = {k: sum(x[k] for x in so) for k in s}
    want["total_blocks"] = so[-1]["total_blocks"]
    assert s == want
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="far-apart batch")


def test_non_finite_pose_is_rejected(product_lib):
    import khronos_b200 as kb
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 2)
    g = hs.make_handle(product_lib, "kb_", cam=cam)
    T = np.array(poses[0], dtype=np.float64)
    T[0, 3] = np.nan
    with pytest.raises(kb.KbError) as e:
        g.integrate_frame(g.make_frame(frames[0][0], T, stamps[0], label=frames[0][1]))
    assert e.value.status == 1
    g.integrate_frame(g.make_frame(frames[1][0], poses[1], stamps[1], label=frames[1][1]))  # the handle stays usable
    assert g.num_blocks() > 0
