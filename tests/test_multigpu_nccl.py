"""Real multi-GPU run (needs >= 2 GPUs, skipped otherwise): one process per GPU, NCCL frame broadcast from
rank 0, block-hash sharded fusion with kb_set_shard, per-rank export; rank 0 checks that the union of the
shards is bit-identical to the unsharded CPU oracle."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes
    import torch.distributed as dist
    import khronos_b200 as kb
    from khronos_b200 import capi, synthetic as syn, distributed as kd
    import harness as hs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cam = hs.small_camera(4)
        poses, stamps = syn.orbit_trajectory(12, laps=0.15)
        n = len(poses)
        depth = torch.zeros((n, cam.height, cam.width), dtype=torch.float32, device=dev)
        label = torch.zeros((n, cam.height, cam.width), dtype=torch.int32, device=dev)
        if rank == 0:
            fr = hs.render_frames(syn.room_scene(), cam, poses, stamps)
            depth.copy_(torch.from_numpy(np.stack([f[0] for f in fr])))
            label.copy_(torch.from_numpy(np.stack([f[1] for f in fr])))
        kd.broadcast_frames(depth, label, src=0)   # NCCL over NVLink
        torch.cuda.synchronize()
        h = hs.make_handle(kb.lib(), "kb_", cam=cam, device=rank)  # the map lives on this rank's GPU
        h.set_shard(rank, world)
        frames = [h.make_frame(depth[i].data_ptr(), poses[i], stamps[i], label=label[i].data_ptr(), memory=capi.MEM_DEVICE)
                  for i in range(n)]
        h.integrate_frames(frames[:8], want_stats=False)
        for f in frames[8:]:
            h.integrate_frame(f, want_stats=False)
        b = h.export_blocks()
        gathered = [None] * world
        dist.all_gather_object(gathered, {k: getattr(b, k) for k in ("block_index", "distance", "weight", "last_observed",
                                                                    "semantic_label", "semantic_empty", "block_flags")})
        if rank == 0:
            oracle = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
            o = hs.make_handle(oracle, "ko_", cam=cam)
            host = [(depth[i].cpu().numpy(), label[i].cpu().numpy()) for i in range(n)]
            hs.run_fusion(o, host, poses, stamps)
            bo = o.export_blocks()
            idx = np.concatenate([g["block_index"] for g in gathered])
            order = np.lexsort(idx[:, ::-1].T)
            ok = len(idx) == bo.n and (idx[order] == bo.block_index).all()
            for k in ("distance", "weight"):
                ok = ok and (np.concatenate([g[k] for g in gathered])[order].view(np.uint32) == getattr(bo, k).view(np.uint32)).all()
            for k in ("last_observed", "semantic_label", "semantic_empty", "block_flags"):
                ok = ok and (np.concatenate([g[k] for g in gathered])[order] == getattr(bo, k)).all()
            q.put(("ok" if ok else "mismatch", [len(g["block_index"]) for g in gathered], bo.n))
    except Exception as e:  # pragma: no cover
        q.put(("error on rank %d: %r" % (rank, e), [], 0))  # any rank reports, so the parent never waits in vain
        os._exit(1)  # do not linger in a half-dead NCCL communicator
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_sharded_fusion_matches_oracle():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    try:
        status, sizes, n = q.get(timeout=240)
    finally:
        for p in procs:
            p.join(timeout=20)
        for p in procs:  # never leave a rank stuck in a collective behind
            if p.is_alive():
                p.kill()
    assert status == "ok", status
    assert sum(sizes) == n and min(sizes) > 0.25 * n


def _pipeline_worker(rank, world, port, q):
    """Sharded per-frame pipeline (motion detection -> integration -> tracking) over NCCL: the pixel-flag
    all-reduce and the two halo all-gathers run on the stream the library's kernels are ordered with."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(1)
    import ctypes
    import torch.distributed as dist
    import khronos_b200 as kb
    from khronos_b200 import capi, distributed as kd
    import harness as hs
    import test_sharded_pipeline as tsp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cam = hs.small_camera(4)
        n = 26
        frames, poses, stamps = tsp.dynamic_scenario(cam, n) if rank == 0 else (None, None, None)
        meta = [poses, stamps]
        dist.broadcast_object_list(meta, src=0)
        poses, stamps = meta
        depth = torch.zeros((n, cam.height, cam.width), dtype=torch.float32, device=dev)
        label = torch.zeros((n, cam.height, cam.width), dtype=torch.int32, device=dev)
        if rank == 0:
            depth.copy_(torch.from_numpy(np.stack([f[0] for f in frames])))
            label.copy_(torch.from_numpy(np.stack([f[1] for f in frames])))
        kd.broadcast_frames(depth, label, src=0)
        torch.cuda.synchronize()
        mot = capi.default_motion_config(min_cluster_size=5, min_separation_distance=2.0)
        h = hs.make_handle(kb.lib(), "kb_", cam=cam, mot_cfg=mot, device=rank)
        h.set_shard(rank, world)
        win = kd.ShardedActiveWindow([h], kd.DistComm(world), device=dev)
        ref = None
        if rank == 0:
            oracle = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
            ref = hs.make_handle(oracle, "ko_", cam=cam, mot_cfg=mot)
        ok, dyn = True, 0
        for i in range(n):
            f = h.make_frame(depth[i].data_ptr(), poses[i], stamps[i], label=label[i].data_ptr(), memory=capi.MEM_DEVICE)
            (img, ns, nc), = win.spin_once([f])
            if ref is not None:
                img_o, ns_o, nc_o = ref.spin_once(ref.make_frame(frames[i][0], poses[i], stamps[i], label=frames[i][1]))
                ok = ok and (ns, nc) == (ns_o, nc_o) and bool((img == img_o).all())
                dyn += int((img_o > 0).sum())
        b = h.export_blocks()
        gathered = [None] * world
        dist.all_gather_object(gathered, b)
        if rank == 0:
            msg = "ok" if (ok and dyn > 50) else "dynamic image / counts differ (or no motion)"
            try:
                tsp.assert_union_equals(gathered, ref.export_blocks(), "nccl shards")
            except AssertionError as e:
                msg = "mismatch: %s" % str(e)[:300]
            q.put((msg, [g.n for g in gathered], ref.export_blocks().n))
    except Exception as e:  # pragma: no cover
        q.put(("error on rank %d: %r" % (rank, e), [], 0))
        os._exit(1)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_sharded_pipeline_matches_oracle():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + (os.getpid() % 200)
    procs = [ctx.Process(target=_pipeline_worker, args=(r, world, port, q), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    try:
        status, sizes, n = q.get(timeout=300)
    finally:
        for p in procs:
            p.join(timeout=20)
        for p in procs:
            if p.is_alive():
                p.kill()
    assert status == "ok", status
    assert sum(sizes) == n and min(sizes) > 0
