"""Block-hash sharded per-frame pipeline (SURVEY.md §8e exchange steps 1 + 2): motion detection -> integration ->
tracking with the map split over S shards. The shards live in one process here (LocalComm stands in for the
collectives); tests/test_multiproc_gloo.py and tests/test_multigpu_nccl.py run the same driver over real
process groups.

CPU: S oracle shards (the oracle implements the protocol on host buffers) == the unsharded oracle — this pins
the protocol itself (pre-pass free masks of neighbour blocks + MAX-reduced pixel flags are sufficient).
GPU: S product shards on one device == the unsharded oracle, bit for bit."""
import numpy as np
import pytest

from khronos_b200 import capi, synthetic as syn, distributed as kd
import harness as hs


_CACHE = {}


def dynamic_scenario(cam, n=30):
    """Static burn-in (> temporal_buffer) so free space becomes ever-free, then a cuboid moves through it."""
    key = ("dyn", cam.width, cam.height, n)
    if key not in _CACHE:
        _CACHE[key] = _dynamic_scenario(cam, n)
    return _CACHE[key]


def orbit_scenario(cam, n=14):
    key = ("orbit", cam.width, cam.height, n)
    if key not in _CACHE:
        poses, stamps = syn.orbit_trajectory(n, laps=0.5, dt_ns=500_000_000)
        _CACHE[key] = (hs.render_frames(syn.room_scene(), cam, poses, stamps), poses, stamps)
    return _CACHE[key]


def _dynamic_scenario(cam, n):
    scene = syn.room_scene()
    scene.mover = ((0.5, 0.5, 1.2), (8.6, 1.5, 0.9), (0.0, 2.0, 0.0), 2.0)
    dt = 150_000_000
    pose = syn.look_pose((6.0, 5.0, 1.5), 0.0, np.radians(10.0))
    poses = [pose] * n
    stamps = [1_000_000_000 + i * dt for i in range(n)]
    return hs.render_frames(scene, cam, poses, stamps), poses, stamps


def union_of(parts):
    idx = np.concatenate([p.block_index for p in parts])
    order = np.lexsort(idx[:, ::-1].T)
    return order, lambda name: np.concatenate([getattr(p, name) for p in parts])[order]


def assert_union_equals(parts, bo, what):
    assert sum(p.n for p in parts) == bo.n, f"{what}: {[p.n for p in parts]} vs {bo.n}"
    _, cat = union_of(parts)
    np.testing.assert_array_equal(cat("block_index"), bo.block_index, err_msg=what)
    for name in ("distance", "weight"):
        np.testing.assert_array_equal(cat(name).view(np.uint32), getattr(bo, name).view(np.uint32), err_msg=f"{what} {name}")
    for name in ("last_observed", "last_occupied", "ever_free", "active", "to_remove", "semantic_label", "semantic_empty",
                 "block_flags"):
        np.testing.assert_array_equal(cat(name), getattr(bo, name), err_msg=f"{what} {name}")


def run_sharded_vs_unsharded(ref_lib, lib, prefix, nshards, device, sep=2.0, caps=None, reset_every=0, scale=4, n=30, peers=False):
    cam = hs.small_camera(scale)
    frames, poses, stamps = dynamic_scenario(cam, n)
    mot = capi.default_motion_config(min_cluster_size=5, min_separation_distance=sep)
    o = hs.make_handle(ref_lib, "ko_", cam=cam, mot_cfg=mot)
    shards = []
    for r in range(nshards):
        g = hs.make_handle(lib, prefix, cam=cam, mot_cfg=mot)
        g.set_shard(r, nshards)
        if caps:
            g.set_shard_capacity(*caps)
        shards.append(g)
    win = (kd.PeerShardedActiveWindow(shards, kd.LocalPeers(nshards, device), device=device) if peers
           else kd.ShardedActiveWindow(shards, kd.LocalComm(), device=device))
    total_dyn, removed = 0, 0
    for i, ((d, l), T, st) in enumerate(zip(frames, poses, stamps)):
        img_o, ns_o, nc_o = o.spin_once(o.make_frame(d, T, st, label=l))
        res = win.spin_once([g.make_frame(d, T, st, label=l) for g in shards])
        for r, (img, ns, nc) in enumerate(res):
            assert (ns, nc) == (ns_o, nc_o), f"frame {i} shard {r}: seeds/clusters {ns, nc} vs {ns_o, nc_o}"
            np.testing.assert_array_equal(img, img_o, err_msg=f"dynamic image frame {i} shard {r}")
        total_dyn += int((img_o > 0).sum())
        if reset_every and i % reset_every == reset_every - 1:
            ro = o.reset_inactive()
            rs = np.concatenate([g.reset_inactive() for g in shards]).reshape(-1, 3)
            np.testing.assert_array_equal(ro, rs[np.lexsort(rs[:, ::-1].T)])
            removed += len(ro)
    bo = o.export_blocks()
    parts = [g.export_blocks() for g in shards]
    assert_union_equals(parts, bo, f"{nshards} shards")
    for r, p in enumerate(parts):
        own = [lib.kb_block_owner(int(b[0]), int(b[1]), int(b[2]), nshards) if prefix == "kb_" else
               lib.ko_block_owner(int(b[0]), int(b[1]), int(b[2]), nshards) for b in p.block_index]
        assert all(x == r for x in own)
    assert bo.ever_free.sum() > 1000 and total_dyn > 50, "the scenario must produce ever-free space and motion"
    return bo, removed


@pytest.mark.parametrize("nshards", [2, 3])
def test_oracle_shards_equal_unsharded_oracle(oracle_lib, nshards):
    run_sharded_vs_unsharded(oracle_lib, oracle_lib, "ko_", nshards, "cpu", scale=8, n=22)


def test_oracle_shards_with_block_removal(oracle_lib):
    """0.5 s frames on an orbit: blocks leave the temporal window and are removed shard by shard."""
    cam = hs.small_camera(8)
    frames, poses, stamps = orbit_scenario(cam)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    shards = [hs.make_handle(oracle_lib, "ko_", cam=cam) for _ in range(2)]
    for r, g in enumerate(shards):
        g.set_shard(r, 2)
    win = kd.ShardedActiveWindow(shards, kd.LocalComm(), device="cpu")
    removed = 0
    for i, ((d, l), T, st) in enumerate(zip(frames, poses, stamps)):
        o.spin_once(o.make_frame(d, T, st, label=l))
        win.spin_once([g.make_frame(d, T, st, label=l) for g in shards])
        if i % 4 == 3:
            ro = o.reset_inactive()
            rs = np.concatenate([g.reset_inactive() for g in shards]).reshape(-1, 3)
            np.testing.assert_array_equal(ro, rs[np.lexsort(rs[:, ::-1].T)])
            removed += len(ro)
    assert removed > 0
    assert_union_equals([g.export_blocks() for g in shards], o.export_blocks(), "removal")


def test_oracle_shard_exchange_overflow_is_reported(oracle_lib):
    with pytest.raises(capi.KbError):
        run_sharded_vs_unsharded(oracle_lib, oracle_lib, "ko_", 2, "cpu", caps=(4, 4), scale=8, n=22)


@pytest.mark.gpu
@pytest.mark.parametrize("nshards,sep", [(2, 2.0), (4, 1.0)])
def test_product_shards_equal_unsharded_oracle(oracle_lib, product_lib, nshards, sep):
    run_sharded_vs_unsharded(oracle_lib, product_lib, "kb_", nshards, "cuda", sep=sep)


@pytest.mark.gpu
def test_product_shards_tracking_only_with_removal(oracle_lib, product_lib):
    """Sharded K2/K3/K2r without motion detection: update_tracking over the exchange, blocks removed per shard."""
    cam = hs.small_camera(4)
    frames, poses, stamps = orbit_scenario(cam)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    shards = [hs.make_handle(product_lib, "kb_", cam=cam) for _ in range(3)]
    for r, g in enumerate(shards):
        g.set_shard(r, 3)
    win = kd.ShardedActiveWindow(shards, kd.LocalComm(), device="cuda")
    removed = 0
    for i, ((d, l), T, st) in enumerate(zip(frames, poses, stamps)):
        o.integrate_frame(o.make_frame(d, T, st, label=l))
        o.update_tracking(st)
        for g in shards:
            g.integrate_frame(g.make_frame(d, T, st, label=l), want_stats=False)
        win.update_tracking([st] * len(shards), with_motion_result=False)
        if i % 4 == 3:
            ro = o.reset_inactive()
            rs = np.concatenate([g.reset_inactive() for g in shards]).reshape(-1, 3)
            np.testing.assert_array_equal(ro, rs[np.lexsort(rs[:, ::-1].T)])
            removed += len(ro)
    assert removed > 0
    assert_union_equals([g.export_blocks() for g in shards], o.export_blocks(), "tracking-only shards")


@pytest.mark.parametrize("nshards,conn,seed", [(4, 6, 0), (8, 26, 1), (5, 18, 2)])
def test_oracle_shards_random_walk(oracle_lib, nshards, conn, seed):
    """Random camera walk through the room (irregular frame periods, so blocks enter and leave the ever-free work list
    at different passes on different shards), all three ever-free connectivities, up to 8 shards."""
    rng = np.random.default_rng(seed)
    cam = hs.small_camera(8)
    scene = syn.room_scene()
    n = 16
    poses, stamps, t = [], [], 1_000_000_000
    pos, yaw = np.array([6.0, 5.0, 1.5]), 0.0
    for _ in range(n):
        pos = np.clip(pos + rng.normal(0, 0.15, 3) * np.array([1, 1, 0.2]), [3, 3, 1.0], [9, 7, 2.0])
        yaw += rng.normal(0, 0.15)
        poses.append(syn.look_pose(tuple(pos), yaw, np.radians(10.0)))
        t += int(rng.integers(100_000_000, 400_000_000))
        stamps.append(t)
    frames = hs.render_frames(scene, cam, poses, stamps)
    trk = capi.default_tracking_config(num_threads=2)
    trk.neighbor_connectivity = conn
    mot = capi.default_motion_config(min_cluster_size=3, min_separation_distance=1.0, num_threads=2)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam, trk_cfg=trk, mot_cfg=mot)
    shards = [hs.make_handle(oracle_lib, "ko_", cam=cam, trk_cfg=trk, mot_cfg=mot) for _ in range(nshards)]
    for r, g in enumerate(shards):
        g.set_shard(r, nshards)
    win = kd.ShardedActiveWindow(shards, kd.LocalComm(), device="cpu")
    for (d, l), T, st in zip(frames, poses, stamps):
        img_o, ns_o, nc_o = o.spin_once(o.make_frame(d, T, st, label=l))
        for img, ns, nc in win.spin_once([g.make_frame(d, T, st, label=l) for g in shards]):
            assert (ns, nc) == (ns_o, nc_o)
            np.testing.assert_array_equal(img, img_o)
    bo = o.export_blocks()
    assert_union_equals([g.export_blocks() for g in shards], bo, f"random walk {nshards} shards conn {conn}")
    assert bo.ever_free.sum() > 100


def test_oracle_shards_peer_memory_exchange(oracle_lib):
    """The peer-memory variants of the exchanges (producers store into every shard's buffers) give the same maps."""
    run_sharded_vs_unsharded(oracle_lib, oracle_lib, "ko_", 3, "cpu", scale=8, n=22, peers=True)
