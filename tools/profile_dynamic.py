"""Host-side timing breakdown of the per-frame pipeline (diagnostic; run on a GPU box)."""
import ctypes, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import khronos_b200 as kb
from khronos_b200 import capi, synthetic as syn

dev = torch.device("cuda", 0)
n = 400
cam = syn.make_camera()
scene = syn.room_scene(20)
poses, stamps = syn.orbit_trajectory(n, laps=n / 3000.0)
extra = syn.companion_cuboids(poses)
depth, label = syn.render_stream(scene, cam, poses, stamps, device=dev, dtype=torch.float32, extra=extra)
mc = capi.default_map_config(max_blocks=90000)
ic = capi.default_integrator_config()
mot = capi.default_motion_config(min_cluster_size=500, min_separation_distance=2.0)
h = kb.create_map(mc, ic, capi.default_tracking_config(), mot)
h.set_camera(cam)
img = torch.zeros((cam.height, cam.width), dtype=torch.int32, pin_memory=True)
frames = [h.make_frame(depth[i].data_ptr(), poses[i], stamps[i], label=label[i].data_ptr(), memory=capi.MEM_DEVICE) for i in range(n)]
fm = [h.make_frame(depth[i].data_ptr(), poses[i], stamps[i], label=label[i].data_ptr(), mask=capi.MASK_LAST_DETECTION, memory=capi.MEM_DEVICE) for i in range(n)]
det, integ, trk, spin, sync = (h._fn(x) for x in ("detect_motion", "integrate_frames", "update_tracking", "spin_once", "synchronize"))
hp = h._h
ns, nc = ctypes.c_int32(0), ctypes.c_int32(0)
ip = ctypes.c_void_p(img.data_ptr())
t = {"detect": 0.0, "integrate": 0.0, "track": 0.0, "sync": 0.0}
for i in range(200):
    a = time.perf_counter(); det(hp, ctypes.byref(frames[i]), ip, ctypes.byref(ns), ctypes.byref(nc))
    b = time.perf_counter(); integ(hp, ctypes.byref(fm[i]), 1, 1, None)
    c = time.perf_counter(); trk(hp, ctypes.c_uint64(stamps[i]))
    d = time.perf_counter(); sync(hp)
    e = time.perf_counter()
    if i >= 100:
        t["detect"] += b - a; t["integrate"] += c - b; t["track"] += d - c; t["sync"] += e - d
print("separate calls, us per frame:", {k: round(v / 100 * 1e6, 1) for k, v in t.items()})
ts = 0.0
for i in range(200, 400):
    a = time.perf_counter(); st = spin(hp, ctypes.byref(frames[i]), ip, ctypes.byref(ns), ctypes.byref(nc)); b = time.perf_counter()
    assert st == 0
    if i >= 300:
        ts += b - a
print("kb_spin_once us per frame:", round(ts / 100 * 1e6, 1), "clusters", nc.value, "seeds", ns.value)
