"""N>1 host logic on CPU: world_size-2 gloo run of the sharded-map plumbing (frame broadcast from rank 0,
per-rank ownership filter, ragged gather). Each rank fuses the broadcast frames with the CPU oracle and keeps
only the blocks kb_block_owner assigns to it (K1 is independent per block, so this equals sharded fusion);
rank 0 checks that the shards are disjoint and that their union is the unsharded map."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes
    import khronos_b200 as kb
    from khronos_b200 import synthetic as syn, distributed as kd
    import harness as hs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cam = hs.small_camera(8)
        poses, stamps = syn.orbit_trajectory(4, laps=0.1)
        n = len(poses)
        depth = torch.zeros((n, cam.height, cam.width), dtype=torch.float32)
        label = torch.zeros((n, cam.height, cam.width), dtype=torch.int32)
        if rank == 0:  # only the ingest rank has the frames
            fr = hs.render_frames(syn.room_scene(), cam, poses, stamps)
            depth = torch.from_numpy(np.stack([f[0] for f in fr]))
            label = torch.from_numpy(np.stack([f[1] for f in fr]))
        kd.broadcast_frames(depth, label, src=0)
        oracle = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        h = hs.make_handle(oracle, "ko_", cam=cam)
        frames = [(depth[i].numpy(), label[i].numpy()) for i in range(n)]
        hs.run_fusion(h, frames, poses, stamps)
        b = h.export_blocks()
        mine = kd.owner_of_blocks(kb.lib(), b.block_index, world) == rank
        parts = kd.gather_block_indices(b.block_index[mine], world)
        checksum = torch.tensor([float(b.distance[mine].astype(np.float64).sum())], dtype=torch.float64)
        dist.all_reduce(checksum)
        if rank == 0:
            union = np.concatenate(parts)
            ok = (len(union) == b.n and len({tuple(x) for x in union.tolist()}) == b.n
                  and {tuple(x) for x in union.tolist()} == {tuple(x) for x in b.block_index.tolist()}
                  and all(len(p) > 0 for p in parts)
                  and abs(checksum.item() - float(b.distance.astype(np.float64).sum())) < 1e-6)
            q.put(("ok" if ok else "mismatch", [len(p) for p in parts], b.n))
    except Exception as e:  # pragma: no cover
        q.put(("error on rank %d: %r" % (rank, e), [], 0))
        os._exit(1)
    dist.destroy_process_group()


def test_world_size_2_gloo_sharded_fusion():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    try:
        status, sizes, n = q.get(timeout=180)
    finally:
        for p in procs:
            p.join(timeout=20)
        for p in procs:
            if p.is_alive():
                p.kill()
    assert status == "ok", status
    assert sum(sizes) == n and min(sizes) > 0.25 * n


def _pipeline_worker(rank, world, port, q):
    """Sharded per-frame pipeline over a real process group (gloo): every rank drives one oracle shard through
    khronos_b200.distributed.ShardedActiveWindow with DistComm; rank 0 also runs the unsharded oracle."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(1)
    import ctypes
    from khronos_b200 import capi, distributed as kd
    import harness as hs
    import test_sharded_pipeline as tsp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cam = hs.small_camera(8)
        frames, poses, stamps = tsp.dynamic_scenario(cam, 20)
        depth = torch.from_numpy(np.stack([f[0] for f in frames])) if rank == 0 else torch.zeros((20, cam.height, cam.width))
        label = torch.from_numpy(np.stack([f[1] for f in frames])) if rank == 0 else torch.zeros((20, cam.height, cam.width), dtype=torch.int32)
        kd.broadcast_frames(depth, label, src=0)   # only the ingest rank's frames count
        oracle = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        mot = capi.default_motion_config(min_cluster_size=5, min_separation_distance=2.0, num_threads=2)
        h = hs.make_handle(oracle, "ko_", cam=cam, mot_cfg=mot)
        h.set_shard(rank, world)
        win = kd.ShardedActiveWindow([h], kd.DistComm(world), device="cpu")
        ref = hs.make_handle(oracle, "ko_", cam=cam, mot_cfg=mot) if rank == 0 else None
        ok, dyn = True, 0
        for i in range(20):
            d, l = depth[i].numpy(), label[i].numpy()
            (img, ns, nc), = win.spin_once([h.make_frame(d, poses[i], stamps[i], label=l)])
            if ref is not None:
                img_o, ns_o, nc_o = ref.spin_once(ref.make_frame(d, poses[i], stamps[i], label=l))
                ok = ok and (ns, nc) == (ns_o, nc_o) and bool((img == img_o).all())
                dyn += int((img_o > 0).sum())
        b = h.export_blocks()
        gathered = [None] * world
        dist.all_gather_object(gathered, b)
        if rank == 0:
            try:
                tsp.assert_union_equals(gathered, ref.export_blocks(), "gloo shards")
            except AssertionError as e:
                ok = False
                q.put(("mismatch: %s" % str(e)[:300], [], 0))
            if ok:
                q.put(("ok" if dyn > 30 else "no motion in the scenario", [g.n for g in gathered], ref.export_blocks().n))
            elif dyn >= 0:
                q.put(("dynamic image / counts differ", [], 0))
    except Exception as e:  # pragma: no cover
        q.put(("error on rank %d: %r" % (rank, e), [], 0))
        os._exit(1)
    dist.destroy_process_group()


def test_world_size_2_gloo_sharded_pipeline():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_pipeline_worker, args=(r, 2, port, q), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    try:
        status, sizes, n = q.get(timeout=240)
    finally:
        for p in procs:
            p.join(timeout=20)
        for p in procs:
            if p.is_alive():
                p.kill()
    assert status == "ok", status
    assert sum(sizes) == n and min(sizes) > 0


def _replay_worker(rank, world, port, q):
    """Cell-sharded striped replay (khronos_b200/replay.py) over a real process group: every rank holds only its stripe
    of the stream, computes the owner masks, pulls the frames it needs (the all-gathered pools stand in for the CUDA IPC
    mappings the GPU build reads through) and integrates its sub-sequence with an oracle shard; the all-gathered
    checksums must add up to the unsharded oracle map's."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(1)
    import ctypes
    from khronos_b200 import synthetic as syn
    from khronos_b200.replay import StripedSchedule, rank_grid
    import harness as hs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cam = hs.small_camera(8)
        n, stripe, cell = 32, 4, 10
        scene = syn.hall_scene(size=(20.0, 16.0, 6.0))
        poses, stamps = syn.sweep_trajectory(n, size=(20.0, 16.0), margin=4.0, lanes=2, yaw_turns=1.5)
        sched = StripedSchedule(world, rank, stripe)
        res = sched.resident(n)
        mine_rendered = hs.render_frames(scene, cam, [poses[g] for g in res], [stamps[g] for g in res])  # only the own stripe
        pool_d = torch.from_numpy(np.stack([f[0] for f in mine_rendered]))
        pool_l = torch.from_numpy(np.stack([f[1] for f in mine_rendered]))
        sizes = [None] * world
        dist.all_gather_object(sizes, len(res))
        cap = max(sizes)
        pad_d = torch.zeros((cap,) + pool_d.shape[1:], dtype=pool_d.dtype); pad_d[:len(res)] = pool_d
        pad_l = torch.zeros((cap,) + pool_l.shape[1:], dtype=pool_l.dtype); pad_l[:len(res)] = pool_l
        peers_d = [torch.zeros_like(pad_d) for _ in range(world)]
        peers_l = [torch.zeros_like(pad_l) for _ in range(world)]
        dist.all_gather(peers_d, pad_d)
        dist.all_gather(peers_l, pad_l)
        oracle = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        gx, gy = rank_grid(world)
        h = hs.make_handle(oracle, "ko_", cam=cam)
        h.set_shard_cells(rank, world, cell, gx, gy)
        masks = h.frame_owners([h.make_frame(None, poses[g], stamps[g]) for g in range(n)])
        plan = sched.plan(list(range(n)), masks)
        rx_d = torch.zeros((max(plan.n_remote, 1),) + pool_d.shape[1:], dtype=pool_d.dtype)
        rx_l = torch.zeros((max(plan.n_remote, 1),) + pool_l.shape[1:], dtype=pool_l.dtype)
        for (src, li, slot, cnt) in plan.ranges:
            rx_d[slot:slot + cnt] = peers_d[src][li:li + cnt]
            rx_l[slot:slot + cnt] = peers_l[src][li:li + cnt]
        for _, g, slot in plan.mine:
            d, l = (rx_d[slot], rx_l[slot]) if slot >= 0 else (pool_d[-slot - 1], pool_l[-slot - 1])
            h.integrate_frame(h.make_frame(d.numpy(), poses[g], stamps[g], label=l.numpy()), want_stats=False)
        cs = [None] * world
        dist.all_gather_object(cs, (h.map_checksum(), len(plan.mine), plan.n_remote))
        if rank == 0:
            ref = hs.make_handle(oracle, "ko_", cam=cam)
            hs.run_fusion(ref, hs.render_frames(scene, cam, poses, stamps), poses, stamps)
            want = ref.map_checksum()
            M = (1 << 64) - 1
            got = (sum(c[0][0] for c in cs) & M, cs[0][0][1] ^ cs[1][0][1], sum(c[0][2] for c in cs), sum(c[0][3] for c in cs))
            ok = got == want and all(c[1] > 0 for c in cs) and sum(c[1] for c in cs) < world * n and any(c[2] > 0 for c in cs)
            q.put(("ok" if ok else f"mismatch {got} vs {want}, frames {[c[1:] for c in cs]}", [c[1] for c in cs], n))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put(("error on rank %d: %r %s" % (rank, e, traceback.format_exc()[-400:]), [], 0))
        os._exit(1)
    dist.destroy_process_group()


def test_world_size_2_gloo_cell_sharded_replay():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_replay_worker, args=(r, 2, port, q), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    try:
        status, sizes, n = q.get(timeout=240)
    finally:
        for p in procs:
            p.join(timeout=20)
        for p in procs:
            if p.is_alive():
                p.kill()
    assert status == "ok", status
    assert max(sizes) <= n and min(sizes) > 0
