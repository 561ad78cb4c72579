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
