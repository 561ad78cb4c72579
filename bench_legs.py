"""Secondary legs of bench.py (N = 1 only; each returns a dict that goes into the one JSON line):

  output_tick   what an output tick (ActiveWindow::extractOutputData, active_window.cpp:217-249, every 0.4 s) costs on the
                benchmarked map: marching cubes on the device + triangles to the host (kb_generate_mesh / kb_get_mesh)
                versus mirroring the same updated blocks back to a host VolumetricMap (kb_export_blocks UPDATED, what
                MeshIntegrator::generateMesh + cloneUpdated need on the CPU)
  dynamic       BASELINE config[2] on the hall stream: per-frame kb_spin_once (motion detection -> integration with the
                dynamic mask -> tracking pass) after a burn-in lap that builds the full map with tracking state
  next_rows     the SURVEY §8f rows around the path: object detection, tracker measurements, ray index — latency, byte
                model, CPU oracle beside it
The CPU arms call the oracle (test infrastructure) as the checker / baseline only, like bench.py's cpu_baseline leg."""
import ctypes
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
OBJECT_LABELS = (7, 8, 9, 10, 11, 12, 19)


def _oracle():
    return ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))


def leg_output_tick(h, make_batch, ticks=8, frames_per_tick=12):
    """h: the benchmarked map (all of its blocks are flagged updated after the timed laps). make_batch(t) -> (ctypes
    Frame array, n) with the 12 frames of tick t. Per tick: integrate, then time (a) kb_export_blocks(UPDATED) of every
    field mirrorBack repopulates, (b) kb_generate_mesh(true, true) + kb_get_mesh; then kb_clear_updated."""
    from khronos_b200 import capi
    import torch
    integrate_n = h._fn("integrate_frames")
    # first tick of a long run: everything is flagged; clear so that the timed ticks see 12 frames' worth of blocks
    h.generate_mesh(True, True)
    h.clear_updated()
    mesh_ms, mirror_ms, blocks, verts, mesh_bytes, mirror_bytes, gen_ms = [], [], [], [], [], [], []
    for t in range(ticks):
        arr, n = make_batch(t)
        st = integrate_n(h._h, arr, n, 1, None)
        if st != 0:
            raise RuntimeError(f"kb_integrate_frames failed: {st}")
        h.synchronize()
        t0 = time.perf_counter()
        b = h.export_blocks(capi.EXPORT_UPDATED, likelihoods=True)
        mirror_ms.append((time.perf_counter() - t0) * 1e3)
        per_block = h.V * (4 + 4 + 3 + 8 + 8 + 1 + 1 + 1 + 4 + 1 + 4 * h.L) + 13
        mirror_bytes.append(b.n * per_block)
        t0 = time.perf_counter()
        nb, nv = ctypes.c_int32(0), ctypes.c_int64(0)
        h._check(h._fn("generate_mesh")(h._h, 1, 1, ctypes.c_float(1e-4), ctypes.byref(nb), ctypes.byref(nv)))
        h.synchronize()
        t1 = time.perf_counter()
        v = nv.value
        pts, col, lab = np.empty((v, 3), np.float32), np.empty((v, 3), np.uint8), np.empty(v, np.uint32)
        bi, off = np.empty((nb.value, 3), np.int32), np.empty(nb.value + 1, np.int64)
        h._check(h._fn("get_mesh")(h._h, ctypes.c_void_p(bi.ctypes.data), ctypes.c_void_p(off.ctypes.data), ctypes.c_void_p(pts.ctypes.data),
                                   ctypes.c_void_p(col.ctypes.data), ctypes.c_void_p(lab.ctypes.data), ctypes.c_int64(v)))
        t2 = time.perf_counter()
        gen_ms.append((t1 - t0) * 1e3)
        mesh_ms.append((t2 - t0) * 1e3)
        blocks.append(nb.value)
        verts.append(v)
        mesh_bytes.append(v * (12 + 3 + 4) + nb.value * 20)
        h.clear_updated()
    torch.cuda.synchronize()
    med = lambda x: float(np.median(x))
    return {"frames_per_tick": frames_per_tick, "ticks": ticks, "updated_blocks_per_tick": med(blocks), "mesh_vertices_per_tick": med(verts),
            "mesh_tick_ms": med(mesh_ms), "mesh_generate_ms": med(gen_ms), "mesh_d2h_bytes_per_tick": med(mesh_bytes),
            "mirror_back_tick_ms": med(mirror_ms), "mirror_back_d2h_bytes_per_tick": med(mirror_bytes),
            "note": "host wall time per output tick (0.4 s of stream = 12 frames): kb_generate_mesh(true, true) + kb_get_mesh versus "
                    "kb_export_blocks(UPDATED) of the same blocks (TSDF, tracking, semantics incl. likelihoods: what mirrorBack repopulates "
                    "for a CPU MeshIntegrator); reference call sites active_window.cpp:223,229"}


def leg_dynamic_hall(args, cam, scene, poses, stamps, depth, label, dev, cpu_threads, n_timed=600, small=False):
    """Per-frame pipeline (ActiveWindow::spinOnce steps, active_window.cpp:127,209-214) on the hall stream. Burn-in: one lap
    of the static stream through kb_spin_once (builds the whole map with per-voxel tracking state and ever-free labels);
    timed: the next n_timed frames of the trajectory with a box that stays ~2.2 m in front of the camera (~20 % of the
    pixels), which crosses space the burn-in lap has labelled ever-free."""
    import torch
    import khronos_b200 as kb
    from khronos_b200 import capi, synthetic as syn
    lap = len(poses)
    L = 20
    mc = capi.default_map_config(voxel_size=0.05, vps=16, trunc=0.15, with_semantics=True, with_tracking=True,
                                 max_blocks=args.max_blocks if not small else 8192)
    ic = capi.default_integrator_config(semantic_mode=capi.SEM_MLE, num_labels=L)
    mot = capi.default_motion_config(min_cluster_size=500 if not small else 30, min_separation_distance=2.0)
    h = kb.create_map(mc, ic, capi.default_tracking_config(), mot, device=dev.index or 0)
    h.set_camera(cam)
    n_timed = min(n_timed, lap)
    dt = 33_333_333
    stamp = lambda g: 1_000_000_000 + g * dt
    t_r = time.perf_counter()
    dposes = [poses[i % lap] for i in range(n_timed)]
    extra = syn.companion_cuboids(dposes, start_frame=0, size=(1.5, 1.5, 2.2))  # ~20 % of a 640x480 / f = 320 image at 2.2 m
    dd, dl = syn.render_stream(scene, cam, dposes, [stamp(lap + i) for i in range(n_timed)], device=dev, dtype=torch.float32, extra=extra)
    torch.cuda.synchronize()
    t_r = time.perf_counter() - t_r
    spin = h._fn("spin_once")
    img_host = torch.zeros((cam.height, cam.width), dtype=torch.int32, pin_memory=True)
    img_ptr = ctypes.c_void_p(img_host.data_ptr())
    burn = [h.make_frame(depth[i].data_ptr(), poses[i], stamp(i), label=label[i].data_ptr(), memory=capi.MEM_DEVICE) for i in range(lap)]
    timed = [h.make_frame(dd[i].data_ptr(), dposes[i], stamp(lap + i), label=dl[i].data_ptr(), memory=capi.MEM_DEVICE) for i in range(n_timed)]
    ns, nc = ctypes.c_int32(0), ctypes.c_int32(0)
    t0 = time.perf_counter()
    for f in burn:
        st = spin(h._h, ctypes.byref(f), None, ctypes.byref(ns), ctypes.byref(nc))
        if st != 0:
            raise RuntimeError(f"kb_spin_once failed in the burn-in lap: {st}")
    h.synchronize()
    burn_s = time.perf_counter() - t0
    clusters, flagged = [], []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f in timed:
        st = spin(h._h, ctypes.byref(f), img_ptr, ctypes.byref(ns), ctypes.byref(nc))  # dynamic image -> pinned host, like FrameData
        if st != 0:
            raise RuntimeError(f"kb_spin_once failed: {st}")
        clusters.append(nc.value)
    h.synchronize()
    dtm = time.perf_counter() - t0
    flagged_last = float((img_host.numpy() > 0).mean())
    tot = h.get_totals64()
    out = {"workload": "hall640-dynamic (BASELINE config[2] on the config[1] stream)", "value": n_timed / dtm, "unit": "frames/s",
           "frames": n_timed, "burn_in_frames": lap, "burn_in_fps": lap / burn_s, "live_blocks": int(tot.total_blocks),
           "frames_with_clusters": int(sum(1 for c in clusters if c > 0)), "flagged_pixel_fraction_last_frame": flagged_last,
           "pipeline": "kb_spin_once per frame: M1 lookup + device clustering (M2-M4) -> K0/K1 with the dynamic image as mask -> K2 (lazy, "
                       "O(blocks)) + K3; one host round trip per frame (counts + dynamic image to pinned host memory)",
           "render_s": round(t_r, 1)}
    h.close()
    if not args.no_cpu_baseline:
        # CPU arm (oracle port): a short burn-in (the reference's tracking pass is O(all allocated blocks), tracking_integrator.cpp
        # :75,83-90, so the small map FLATTERS the CPU arm), then timed on the first dynamic frames
        n_b, n_t = (150, 24) if not small else (20, 6)
        oh = capi.MapHandle(_oracle(), "ko_", mc, capi.default_integrator_config(semantic_mode=capi.SEM_MLE, num_labels=L, num_threads=cpu_threads),
                            capi.default_tracking_config(num_threads=cpu_threads), capi.default_motion_config(min_cluster_size=mot.min_cluster_size,
                                                                                                             min_separation_distance=2.0, num_threads=cpu_threads))
        oh.set_camera(cam)
        # burn in on the frames right before the lap's end so that the timed frames (lap start) continue the trajectory
        bd, bl = depth[lap - n_b:].cpu().numpy(), label[lap - n_b:].cpu().numpy()
        for i in range(n_b):
            oh.spin_once(oh.make_frame(bd[i], poses[lap - n_b + i], stamp(lap - n_b + i), label=bl[i]), want_image=False)
        td, tl = dd[:n_t].cpu().numpy(), dl[:n_t].cpu().numpy()
        t0 = time.perf_counter()
        for i in range(n_t):
            oh.spin_once(oh.make_frame(td[i], dposes[i], stamp(lap + i), label=tl[i]))
        ct = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_t / ct, "unit": "frames/s", "cores": cpu_threads, "kind": "port",
                               "sample": f"{n_t} dynamic frames after a {n_b}-frame burn-in ({oh.get_totals().total_blocks} blocks in the map vs "
                                         f"{int(tot.total_blocks)} on the GPU arm: the reference's per-frame tracking pass scans every allocated block, "
                                         "so the short burn-in flatters the CPU arm)"}
        oh.close()
    return out


def leg_next_rows(dev, small=False, cpu=True):
    """§8f rows 2, 3, 4 at 640x480 / 1 M rays: latency of the C-ABI calls (synchronous: they return counts), algorithmic
    bytes, the oracle port on one host thread beside it."""
    import torch
    import khronos_b200 as kb
    from khronos_b200 import capi, synthetic as syn
    cam = syn.make_camera() if not small else syn.make_camera(160, 120, 80.0, 80.0)
    scene = syn.room_scene()
    pose = syn.look_pose((6.0, 5.0, 1.5), 3.7, np.radians(12.0))
    d, l = syn.render(scene, cam, pose)
    d, l = d.numpy().astype(np.float32), l.numpy().astype(np.int32).copy()
    rng = np.random.default_rng(3)
    m = rng.random(l.shape) < 0.03  # salt the label image with small object specks
    l[m] = rng.choice(np.array(OBJECT_LABELS + (1, 3), np.int32), size=int(m.sum()))
    P = cam.width * cam.height
    h = kb.create_map(capi.default_map_config(max_blocks=4096), capi.default_integrator_config(), capi.default_tracking_config(), None,
                      device=dev.index or 0)
    h.set_camera(cam)
    dd, ll = torch.from_numpy(d).to(dev), torch.from_numpy(l).to(dev)
    torch.cuda.synchronize()
    f = h.make_frame(dd.data_ptr(), pose, 1_000_000_000, label=ll.data_ptr(), memory=capi.MEM_DEVICE)
    oh = None
    if cpu:
        oh = capi.MapHandle(_oracle(), "ko_", capi.default_map_config(max_blocks=4096), capi.default_integrator_config(num_threads=1),
                            capi.default_tracking_config(num_threads=1), None)
        oh.set_camera(cam)
        fo = oh.make_frame(d, pose, 1_000_000_000, label=l)
    iters = 30 if not small else 3
    out = {"image": [cam.width, cam.height]}

    def timeit(fn, n):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(n):
            r = fn()
        return (time.perf_counter() - t0) / n * 1e6, r

    for name, use_3d in (("detect_objects_3d", True), ("detect_objects_2d", False)):
        cfg = capi.default_object_detector_config(OBJECT_LABELS, use_3d=use_3d, min_cluster_size=50)
        us, (ids, n) = timeit(lambda: h.detect_objects(cfg, f), iters)
        row = {"us": round(us, 1), "clusters": int(n), "algorithmic_bytes": P * (4 + 4 + 4),
               "reference": "connected_semantics.cpp:70-144 (3D) / :146-198 (2D)"}
        if oh is not None:
            t0 = time.perf_counter()
            ido, no = oh.detect_objects(cfg, fo)
            row["cpu_oracle_us"] = round((time.perf_counter() - t0) * 1e6, 1)
            row["matches_oracle"] = bool(no == n and np.array_equal(ido, ids))
        out[name] = row
    cfg = capi.default_object_detector_config(OBJECT_LABELS, use_3d=True, min_cluster_size=50)
    ids, n = h.detect_objects(cfg, f)
    cids = [c["id"] for c in h.get_object_clusters()]
    ii = torch.from_numpy(ids).to(dev)
    torch.cuda.synchronize()
    h.track_measurements(f, ii.data_ptr(), cids, 0.1, [])
    tracks = [v for v in h.get_cluster_voxels(len(cids)) if len(v)][:16]
    us, r = timeit(lambda: h.track_measurements(f, ii.data_ptr(), cids, 0.1, tracks), iters)
    row = {"us": round(us, 1), "clusters": len(cids), "tracks": len(tracks), "cluster_voxels": int(r["voxel_counts"].sum()),
           "algorithmic_bytes": P * 8 + int(sum(len(t) for t in tracks)) * 24 + len(cids) * len(tracks) * 8,
           "reference": "max_iou_tracker.cpp:450-459,534-539,551-562"}
    if oh is not None:
        t0 = time.perf_counter()
        ro = oh.track_measurements(fo, ids, cids, 0.1, tracks)
        row["cpu_oracle_us"] = round((time.perf_counter() - t0) * 1e6, 1)
        row["matches_oracle"] = bool(np.array_equal(ro["voxel_counts"], r["voxel_counts"]) and np.array_equal(ro["intersections"], r["intersections"]))
    out["track_measurements"] = row
    h.close()
    if oh is not None:
        oh.close()

    # ray index: ~1 M rays (vertices on the walls of a 40 x 30 x 4 m hall, policy kAll over nearby poses would explode: kMiddle x 2.5 vertices)
    scale = 1.0 if not small else 0.01
    n_v, n_p, n_q = int(1_000_000 * scale), 200, int(200_000 * scale)
    stamps = (np.uint64(1_000_000_000) + np.arange(n_p, dtype=np.uint64) * np.uint64(100_000_000))
    poses = np.stack([np.linspace(5, 35, n_p), 15 + 8 * np.sin(np.linspace(0, 6, n_p)), np.full(n_p, 1.5)], 1).astype(np.float32)
    verts = rng.uniform([0, 0, 0], [40, 30, 4], (n_v, 3)).astype(np.float32)
    wall = rng.integers(0, 3, n_v)
    for a, hi in enumerate((40.0, 30.0, 4.0)):
        verts[wall == a, a] = np.where(rng.integers(0, 2, int((wall == a).sum())) == 1, hi, 0.0)
    first = rng.integers(1_000_000_000, 20_000_000_000, n_v).astype(np.uint64)
    last = first + rng.integers(0, 2_000_000_000, n_v).astype(np.uint64)
    rr = capi.RayIndex(kb.lib(), "kb_", capi.default_ray_config())
    t0 = time.perf_counter()
    _, n_rays = rr.add_vertices(capi.RAYS_MIDDLE, stamps, poses, verts, first, last)
    t_add = time.perf_counter() - t0
    pts = (verts[rng.integers(0, n_v, n_q)] + rng.normal(0, 0.05, (n_q, 3))).astype(np.float32)
    rr.check(pts[:1000])  # builds the CSR
    t0 = time.perf_counter()
    counts, _ = rr.check(pts)
    t_chk = time.perf_counter() - t0
    entries = int(rr.size()[1])
    row = {"rays": int(n_rays), "block_entries": entries, "add_s": round(t_add, 4), "rays_per_s": round(n_rays / t_add),
           "points": n_q, "check_s": round(t_chk, 4), "points_per_s": round(n_q / t_chk), "verdicts": int(counts.sum()),
           "algorithmic_bytes_add": int(n_rays) * 32 + entries * 12, "algorithmic_bytes_check": n_q * 28 + int(counts.sum()) * 8,
           "reference": "ray_verificator.cpp:66-146,326-350", "note": "host-side call latency incl. H2D/D2H and the Python unpacking of the stamp lists"}
    if cpu:
        k_v, k_q = max(1, n_v // 50), max(1, n_q // 50)  # bounded CPU sample: 2 % of the rays / points
        ro = capi.RayIndex(_oracle(), "ko_", capi.default_ray_config())
        t0 = time.perf_counter()
        _, n_o = ro.add_vertices(capi.RAYS_MIDDLE, stamps, poses, verts[:k_v], first[:k_v], last[:k_v])
        ta = time.perf_counter() - t0
        qp = (verts[rng.integers(0, k_v, k_q)] + rng.normal(0, 0.05, (k_q, 3))).astype(np.float32)
        t0 = time.perf_counter()
        ro.check(qp)
        tc = time.perf_counter() - t0
        row["cpu_oracle"] = {"rays_per_s": round(n_o / ta), "points_per_s": round(k_q / tc), "sample": f"{n_o} rays, {k_q} points, one host thread"}
    out["ray_index"] = row
    return out
