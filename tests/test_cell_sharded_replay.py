"""Spatially (cell-) sharded replay: khronos_b200/replay.py + kb_set_shard_cells / kb_frame_owners / kb_gather_*.
CPU: the schedule (stripes, pulls, slots) on numpy pools, and the protocol between oracle shards — union of the shards ==
the unsharded oracle map, checksum sums add up. GPU: product shards on one device ("virtual ranks", pools shared as
same-process peers) with all three gather transports, bit-exact against the unsharded oracle."""
import ctypes as C

import numpy as np
import pytest

from khronos_b200 import capi, synthetic as syn
from khronos_b200.replay import PeerPools, StripedSchedule, rank_grid, route_homes
import harness as hs

M64 = (1 << 64) - 1


def _stream(n=48):
    cam = hs.small_camera(4)
    scene = syn.hall_scene(size=(20.0, 16.0, 6.0))
    poses, stamps = syn.sweep_trajectory(n, size=(20.0, 16.0), margin=4.0, lanes=2, yaw_turns=2.0)
    return cam, hs.render_frames(scene, cam, poses, stamps), poses, stamps


def test_rank_grid():
    assert [rank_grid(n) for n in (1, 2, 4, 6, 8)] == [(1, 1), (2, 1), (2, 2), (3, 2), (4, 2)]


def test_cell_owner_tiling(oracle_lib):
    seen = set()
    for bx in range(-20, 20):
        for by in range(-20, 20):
            o = oracle_lib.ko_cell_owner(bx, by, 4, 4, 2, 8)
            assert 0 <= o < 8
            # all blocks of a cell share the owner; floor division for negative indices
            assert o == oracle_lib.ko_cell_owner((bx // 4) * 4, (by // 4) * 4, 4, 4, 2, 8)
            assert o == ((bx // 4) % 4) + 4 * ((by // 4) % 2)
            seen.add(o)
    assert seen == set(range(8))


def test_schedule_moves_the_right_frames():
    """Numpy pools: after executing the plan's pulls every frame a rank needs is where the plan says it is."""
    world, stripe, lap, P = 3, 4, 50, 5
    rng = np.random.default_rng(1)
    owner_mask = rng.integers(1, 1 << world, size=lap).astype(np.uint32)
    scheds = [StripedSchedule(world, r, stripe) for r in range(world)]
    pools = []
    for s in scheds:
        res = s.resident(lap)
        assert [s.local_index(g) for g in res] == list(range(len(res)))
        pools.append(np.array([[g * 10 + k for k in range(P)] for g in res]))
    assert sorted(g for s in scheds for g in s.resident(lap)) == list(range(lap))
    step = [(7 + j) % lap for j in range(lap)]  # a step that wraps around the lap
    for r, s in enumerate(scheds):
        plan = s.plan(step, owner_mask)
        rx = np.full((max(plan.n_remote, 1), P), -1)
        for (src, li, slot, cnt) in plan.ranges:
            assert src != r
            rx[slot:slot + cnt] = pools[src][li:li + cnt]
        want = [g for g in step if (owner_mask[g] >> r) & 1]
        assert [g for _, g, _ in plan.mine] == want
        assert [j for j, _, _ in plan.mine] == sorted(j for j, _, _ in plan.mine)
        for j, g, slot in plan.mine:
            row = rx[slot] if slot >= 0 else pools[r][-slot - 1]
            assert row[0] == g * 10 and step[j] == g
        assert plan.n_remote == sum(1 for g in want if s.home(g) != r)


def test_routed_homes_schedule():
    """Pose-aware placement: chunks go to a rank that needs them; the schedule still delivers every needed frame."""
    world, stripe, lap, P = 4, 4, 61, 3
    rng = np.random.default_rng(2)
    owner_mask = rng.integers(1, 1 << world, size=lap).astype(np.uint32)
    homes = route_homes(owner_mask, world, stripe)
    for c0 in range(0, lap, stripe):
        assert len(set(homes[c0:c0 + stripe].tolist())) == 1
        r = int(homes[c0])
        votes = [int(((owner_mask[c0:c0 + stripe] >> q) & 1).sum()) for q in range(world)]
        assert votes[r] == max(votes)
    scheds = [StripedSchedule(world, r, stripe, homes=homes) for r in range(world)]
    pools = []
    for s in scheds:
        res = s.resident(lap)
        assert [s.local_index(g) for g in res] == list(range(len(res)))
        pools.append(np.array([[g * 10 + k for k in range(P)] for g in res]).reshape(-1, P))
    assert sorted(g for s in scheds for g in s.resident(lap)) == list(range(lap))
    step = list(range(lap))
    local = remote = 0
    for r, s in enumerate(scheds):
        plan = s.plan(step, owner_mask)
        rx = np.full((max(plan.n_remote, 1), P), -1)
        for (src, li, slot, cnt) in plan.ranges:
            rx[slot:slot + cnt] = pools[src][li:li + cnt]
        for j, g, slot in plan.mine:
            row = rx[slot] if slot >= 0 else pools[r][-slot - 1]
            assert row[0] == g * 10
            local += slot < 0
            remote += slot >= 0
    rr = [StripedSchedule(world, r, stripe).plan(step, owner_mask).n_remote for r in range(world)]
    assert remote < sum(rr), "routing must pull fewer frames than round-robin stripes"


@pytest.mark.parametrize("world,cell", [(2, 10), (4, 8), (8, 5)])
def test_oracle_cell_shards_equal_unsharded(oracle_lib, world, cell):
    cam, frames, poses, stamps = _stream(40)
    gx, gy = rank_grid(world)
    ref = hs.make_handle(oracle_lib, "ko_", cam=cam)
    hs.run_fusion(ref, frames, poses, stamps)
    want = ref.map_checksum()
    shards = []
    for r in range(world):
        h = hs.make_handle(oracle_lib, "ko_", cam=cam)
        h.set_shard_cells(r, world, cell, gx, gy)
        shards.append(h)
    fr0 = [shards[0].make_frame(d, T, st, label=l) for (d, l), T, st in zip(frames, poses, stamps)]
    masks = shards[0].frame_owners(fr0)
    assert all(0 < int(m) < (1 << world) for m in masks)
    assert any(bin(int(m)).count("1") < world for m in masks), "cell sharding should spare some ranks some frames"
    total = [0, 0, 0, 0]
    for r, h in enumerate(shards):
        sched = StripedSchedule(world, r, stripe=8)
        plan = sched.plan(list(range(len(frames))), masks)
        for _, g, _ in plan.mine:  # each shard integrates only the frames it needs, in stream order
            d, l = frames[g]
            h.integrate_frame(h.make_frame(d, poses[g], stamps[g], label=l), want_stats=False)
        c = h.map_checksum()
        total = [(total[0] + c[0]) & M64, total[1] ^ c[1], total[2] + c[2], total[3] + c[3]]
        # a frame the mask spares a rank would have found no owned block there
        skipped = [g for g in range(len(frames)) if not (int(masks[g]) >> r) & 1]
        for g in skipped[:3]:
            d, l = frames[g]
            probe = hs.make_handle(oracle_lib, "ko_", cam=cam)
            probe.set_shard_cells(r, world, cell, gx, gy)
            assert probe.integrate_frame(probe.make_frame(d, poses[g], stamps[g], label=l)).blocks_in_frustum == 0
    assert tuple(total) == want


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_product_virtual_ranks_pull_and_fuse(oracle_lib, product_lib, mode):
    """4 product shards on one device; every shard keeps a stripe of the stream in its own pool and pulls the rest through
    a gather plan (mode 0 copy engines, 1 SM loads/stores, 2 cp.async.bulk pipeline). Union == unsharded oracle."""
    import torch
    world, cell, stripe = 4, 8, 8
    cam, frames, poses, stamps = _stream(48)
    n = len(frames)
    gx, gy = rank_grid(world)
    ref = hs.make_handle(oracle_lib, "ko_", cam=cam)
    hs.run_fusion(ref, frames, poses, stamps)
    want = ref.map_checksum()
    dev = torch.device("cuda", 0)
    H, W = cam.height, cam.width
    scheds = [StripedSchedule(world, r, stripe) for r in range(world)]
    pools = []
    for r, s in enumerate(scheds):
        res = s.resident(n)
        pool = PeerPools(product_lib, 0, len(res), H, W)
        dv, lv = pool.views(torch, dev)
        for li, g in enumerate(res):
            dv[li].copy_(torch.from_numpy(frames[g][0]))
            lv[li].copy_(torch.from_numpy(frames[g][1]))
        pools.append(pool)
    torch.cuda.synchronize()
    for r in range(world):
        for q in range(world):
            if q != r:
                pools[r].add_local_peer(q, pools[q].ptr, pools[q].n)
    handles = []
    for r in range(world):
        h = hs.make_handle(product_lib, "kb_", cam=cam)
        h.set_shard_cells(r, world, cell, gx, gy)
        handles.append(h)
    fr0 = [handles[0].make_frame(None, T, st) for T, st in zip(poses, stamps)]
    masks = handles[0].frame_owners(fr0)
    fo = [ref.make_frame(d, T, st, label=l) for (d, l), T, st in zip(frames, poses, stamps)]
    exact = hs.make_handle(oracle_lib, "ko_", cam=cam)
    exact.set_shard_cells(0, world, cell, gx, gy)
    exact_masks = exact.frame_owners(fo)
    assert all(int(a) & int(b) == int(b) for a, b in zip(masks, exact_masks)), "kb_frame_owners must be a superset of the exact selection"
    total = [0, 0, 0, 0]
    stream = torch.cuda.current_stream().cuda_stream
    for r, h in enumerate(handles):
        plan = scheds[r].plan(list(range(n)), masks)
        cap = max(plan.n_remote, 1)
        rx = PeerPools(product_lib, 0, cap, H, W)
        gp = pools[r].gather_plan(plan.ranges, rx.ptr, cap)
        assert pools[r].plan_bytes(gp) == plan.n_remote * H * W * 8
        pools[r].run(gp, mode, 8, stream)
        torch.cuda.synchronize()
        batch = []
        for _, g, slot in plan.mine:
            base, cnt, i = (rx.ptr, cap, slot) if slot >= 0 else (pools[r].ptr, pools[r].n, -slot - 1)
            batch.append(h.make_frame(pools[r].depth_ptr(base, cnt, i), poses[g], stamps[g], label=pools[r].label_ptr(base, cnt, i),
                                      memory=capi.MEM_DEVICE))
        for b0 in range(0, len(batch), 32):
            h.integrate_frames(batch[b0:b0 + 32], want_stats=False)
        c = h.map_checksum()
        total = [(total[0] + c[0]) & M64, total[1] ^ c[1], total[2] + c[2], total[3] + c[3]]
        product_lib.kb_gather_plan_destroy(gp)
        rx.close()
    assert tuple(total) == want
    # and block by block against the unsharded oracle
    bo = ref.export_blocks()
    parts = [h.export_blocks() for h in handles]
    assert sum(p.n for p in parts) == bo.n
    idx = {tuple(b): i for i, b in enumerate(bo.block_index.reshape(-1, 3).tolist())}
    for p in parts:
        for k, b in enumerate(p.block_index.reshape(-1, 3).tolist()):
            i = idx[tuple(b)]
            np.testing.assert_array_equal(p.distance[k].view(np.uint32), bo.distance[i].view(np.uint32))
            np.testing.assert_array_equal(p.weight[k].view(np.uint32), bo.weight[i].view(np.uint32))
            np.testing.assert_array_equal(p.semantic_label[k], bo.semantic_label[i])
            np.testing.assert_array_equal(p.last_observed[k], bo.last_observed[i])
    for p in pools:
        p.close()


def _cell_grid(h, frames, cell):
    """Bounding cell rectangle of a stream (generous) and the per-frame touched cells."""
    ox, oy, w, hgt = -4, -4, 16, 14
    return (ox, oy), h.frame_cells(frames, cell, (ox, oy), w, hgt)


def test_bisect_layout_properties():
    from khronos_b200.replay import bisect_layout
    rng = np.random.default_rng(5)
    F, H, W = 400, 9, 12
    t = np.zeros((F, H, W), np.uint8)
    for f in range(F):  # a 2 x 2 footprint walking along a serpentine
        cx = int(f * (W - 2) / F) if (f // 100) % 2 == 0 else int((F - f) * (W - 2) / F)
        cy = min(H - 2, 2 * (f // 100))
        t[f, cy:cy + 2, cx:cx + 2] = 1
    for world in (2, 4, 8):
        tab = bisect_layout(t, world)
        assert tab.shape == (H, W) and set(np.unique(tab)) <= set(range(world))
        loads = [int(((t != 0) & (tab[None] == r)).any(axis=(1, 2)).sum()) for r in range(world)]
        assert min(loads) > 0, loads
        assert max(loads) <= 2.2 * (sum(loads) / world), loads
        for r in range(world):  # every region is a rectangle
            ys, xs = np.nonzero(tab == r)
            assert (tab[ys.min():ys.max() + 1, xs.min():xs.max() + 1] == r).all()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_oracle_table_shards_equal_unsharded(oracle_lib, world):
    """kb_set_shard_table layout from bisect_layout on oracle shards: union == unsharded, masks consistent with the cells."""
    from khronos_b200.replay import bisect_layout
    cam, frames, poses, stamps = _stream(40)
    ref = hs.make_handle(oracle_lib, "ko_", cam=cam)
    hs.run_fusion(ref, frames, poses, stamps)
    want = ref.map_checksum()
    probe = hs.make_handle(oracle_lib, "ko_", cam=cam)
    fr0 = [probe.make_frame(d, T, st, label=l) for (d, l), T, st in zip(frames, poses, stamps)]
    cell = 4
    origin, touched = _cell_grid(probe, fr0, cell)
    assert touched.any(axis=(1, 2)).all()
    table = bisect_layout(touched, world)
    shards = []
    for r in range(world):
        h = hs.make_handle(oracle_lib, "ko_", cam=cam)
        h.set_shard_table(r, world, cell, origin, table)
        shards.append(h)
    masks = shards[0].frame_owners(fr0)
    # the owner mask of a frame = the ranks of the cells it touches
    for g in range(len(frames)):
        ranks = set(np.unique(table[touched[g] != 0]).tolist())
        assert ranks == {r for r in range(world) if (int(masks[g]) >> r) & 1}, g
    total = [0, 0, 0, 0]
    for r, h in enumerate(shards):
        for g in range(len(frames)):
            if (int(masks[g]) >> r) & 1:
                d, l = frames[g]
                h.integrate_frame(h.make_frame(d, poses[g], stamps[g], label=l), want_stats=False)
        c = h.map_checksum()
        total = [(total[0] + c[0]) & M64, total[1] ^ c[1], total[2] + c[2], total[3] + c[3]]
    assert tuple(total) == want


@pytest.mark.gpu
def test_product_table_layout_matches_oracle_masks_and_map(oracle_lib, product_lib):
    """The explicit cell table on the device: kb_frame_cells / kb_frame_owners agree with (or contain) the oracle's, and
    product shards under the table reproduce the unsharded oracle map."""
    from khronos_b200.replay import bisect_layout
    world, cell = 4, 4
    cam, frames, poses, stamps = _stream(40)
    ref = hs.make_handle(oracle_lib, "ko_", cam=cam)
    hs.run_fusion(ref, frames, poses, stamps)
    want = ref.map_checksum()
    g0 = hs.make_handle(product_lib, "kb_", cam=cam)
    fr0 = [g0.make_frame(None, T, st) for T, st in zip(poses, stamps)]
    origin, touched = _cell_grid(g0, fr0, cell)
    fo = [ref.make_frame(d, T, st, label=l) for (d, l), T, st in zip(frames, poses, stamps)]
    _, touched_o = _cell_grid(ref, fo, cell)
    assert ((touched != 0) | (touched_o == 0)).all(), "kb_frame_cells must contain the oracle's exact selection"
    table = bisect_layout(touched, world)
    total = [0, 0, 0, 0]
    for r in range(world):
        h = hs.make_handle(product_lib, "kb_", cam=cam)
        h.set_shard_table(r, world, cell, origin, table)
        masks = h.frame_owners(fr0)
        batch = [h.make_frame(d, poses[g], stamps[g], label=l) for g, (d, l) in enumerate(frames) if (int(masks[g]) >> r) & 1]
        for b0 in range(0, len(batch), 32):
            h.integrate_frames(batch[b0:b0 + 32], want_stats=False)
        c = h.map_checksum()
        total = [(total[0] + c[0]) & M64, total[1] ^ c[1], total[2] + c[2], total[3] + c[3]]
    assert tuple(total) == want


@pytest.mark.parametrize("world", [1, 3, 5, 6, 7])
def test_layouts_for_any_rank_count(world):
    """bisect_layout / route_homes / rank_grid for rank counts that are not powers of two (and the trivial one)."""
    from khronos_b200.replay import bisect_layout
    rng = np.random.default_rng(world)
    F, H, W = 300, 7, 9
    t = np.zeros((F, H, W), np.uint8)
    for f in range(F):
        cx, cy = int(rng.integers(0, W - 1)), int(rng.integers(0, H - 1))
        t[f, cy:cy + 2, cx:cx + 2] = 1
    tab = bisect_layout(t, world)
    assert set(np.unique(tab)) == set(range(world))
    masks = np.zeros(F, np.uint32)
    for f in range(F):
        for r in np.unique(tab[t[f] != 0]):
            masks[f] |= np.uint32(1 << int(r))
    homes = route_homes(masks, world, 16)
    assert homes.min() >= 0 and homes.max() < world
    gx, gy = rank_grid(world)
    assert gx * gy == world and gx >= gy
    total = 0
    for r in range(world):
        s = StripedSchedule(world, r, 16, homes=homes)
        plan = s.plan(list(range(F)), masks)
        total += len(plan.mine)
        assert all(src != r for src, _, _, _ in plan.ranges)
    assert total == int(sum(bin(int(m)).count("1") for m in masks))
