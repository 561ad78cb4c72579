"""Deterministic synthetic RGB-D + label streams: analytic ray casting of an
axis-aligned room with interior cuboids, pinhole camera, scripted trajectories.

Pure torch, runs on CPU (tests, oracle inputs) or CUDA (bench: frames are rendered straight into
HBM). This is input generation only — it is not part of the integrator hot path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .capi import Camera

SEED = 0x4B48524F4E4F53  # ASCII "KHRONOS"


def make_camera(width=640, height=480, fx=320.0, fy=320.0, cx=None, cy=None, min_range=0.1,
                max_range=5.0) -> Camera:
    """Jackal-like 640x480 pinhole (khronos_ros/config/vio/jackal/LeftCameraParams.yaml:22-25)."""
    cx = (width - 1) / 2.0 if cx is None else cx
    cy = (height - 1) / 2.0 if cy is None else cy
    return Camera(width, height, fx, fy, cx, cy, min_range, max_range)


@dataclass
class Scene:
    room_min: Tuple[float, float, float]
    room_max: Tuple[float, float, float]
    cuboids: np.ndarray  # (K, 6) min xyz, max xyz
    num_labels: int = 20
    # optional moving cuboid: (size xyz, start xyz, velocity xyz [m/s], t_start [s]) -> label L-1
    mover: Optional[Tuple[Sequence[float], Sequence[float], Sequence[float], float]] = None

    def cuboid_labels(self) -> np.ndarray:
        k = np.arange(len(self.cuboids))
        return (7 + k % max(self.num_labels - 8, 1)).astype(np.int32)


def room_scene(num_labels=20) -> Scene:
    """S1 "room": 12 x 10 x 3 m box (floor z=0) + 6 interior cuboids."""
    cub = np.array([
        [2.0, 2.0, 0.0, 3.0, 3.5, 1.2], [8.5, 1.5, 0.0, 10.0, 2.5, 0.9], [9.0, 6.5, 0.0, 10.5, 8.5, 1.5],
        [1.5, 7.0, 0.0, 2.5, 8.0, 2.0], [5.2, 8.6, 0.0, 6.8, 9.4, 0.8], [5.5, 0.6, 0.5, 6.5, 1.4, 1.6],
    ], dtype=np.float64)
    return Scene((0.0, 0.0, 0.0), (12.0, 10.0, 3.0), cub, num_labels)


def hall_scene(num_labels=20, size=(60.0, 40.0, 6.0), pitch=4.0) -> Scene:
    """S2 "hall": large hall with a grid of pillars/crates so every view sees surfaces within 5 m."""
    rng = np.random.default_rng(SEED & 0xFFFFFFFF)
    cubs = []
    nx, ny = int(size[0] // pitch), int(size[1] // pitch)
    for i in range(nx):
        for j in range(ny):
            cx = (i + 0.5) * pitch + rng.uniform(-0.6, 0.6)
            cy = (j + 0.5) * pitch + rng.uniform(-0.6, 0.6)
            sx, sy = rng.uniform(0.4, 1.4, size=2)
            h = rng.uniform(0.6, 3.5)
            cubs.append([cx - sx / 2, cy - sy / 2, 0.0, cx + sx / 2, cy + sy / 2, h])
    return Scene((0.0, 0.0, 0.0), size, np.array(cubs, dtype=np.float64), num_labels)


def look_pose(position, yaw, pitch_down) -> np.ndarray:
    """world_T_sensor for an optical frame (z forward, x right, y down)."""
    f = np.array([math.cos(yaw) * math.cos(pitch_down), math.sin(yaw) * math.cos(pitch_down),
                  -math.sin(pitch_down)])
    r = np.cross(f, np.array([0.0, 0.0, 1.0]))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = r, d, f, np.asarray(position, dtype=np.float64)
    return T


def orbit_trajectory(n, center=(6.0, 5.0), radius=2.5, z=1.5, laps=2.0, pitch_down_deg=10.0,
                     t0_ns=1_000_000_000, dt_ns=33_333_333):
    """"orbit": circle around the room centre looking outward, 10 deg down, 30 Hz."""
    poses, stamps = [], []
    for i in range(n):
        a = 2.0 * math.pi * laps * i / max(n, 1)
        pos = (center[0] + radius * math.cos(a), center[1] + radius * math.sin(a), z)
        poses.append(look_pose(pos, a, math.radians(pitch_down_deg)))
        stamps.append(t0_ns + i * dt_ns)
    return poses, stamps


def sweep_trajectory(n, size=(60.0, 40.0), margin=6.0, z=1.5, lanes=5, yaw_turns=24.0,
                     pitch_down_deg=10.0, t0_ns=1_000_000_000, dt_ns=33_333_333):
    """Serpentine sweep through the hall while the camera yaws continuously, so the frustum covers
    the whole floor plan (~50 k blocks at 5 cm / 16^3, config 2)."""
    poses, stamps = [], []
    ys = np.linspace(margin, size[1] - margin, lanes)
    # polyline of lane segments
    pts = []
    for k, y in enumerate(ys):
        xs = (margin, size[0] - margin) if k % 2 == 0 else (size[0] - margin, margin)
        pts.append((xs[0], y))
        pts.append((xs[1], y))
    pts = np.array(pts)
    seg = np.linalg.norm(np.diff(pts, axis=0), axis=1)
    cum = np.concatenate([[0.0], np.cumsum(seg)])
    for i in range(n):
        s = cum[-1] * i / max(n - 1, 1)
        k = min(int(np.searchsorted(cum, s, side="right")) - 1, len(seg) - 1)
        a = (s - cum[k]) / seg[k] if seg[k] > 0 else 0.0
        p = pts[k] * (1 - a) + pts[k + 1] * a
        yaw = 2.0 * math.pi * yaw_turns * i / max(n, 1)
        poses.append(look_pose((p[0], p[1], z), yaw, math.radians(pitch_down_deg)))
        stamps.append(t0_ns + i * dt_ns)
    return poses, stamps


def _slab(o, d, bmin, bmax):
    # o: (3,), d: (P,3), bmin/bmax: (K,3) -> tnear, tfar: (P,K), far axis (P,K)
    inv = 1.0 / d
    t1 = (bmin[None, :, :] - o[None, None, :]) * inv[:, None, :]
    t2 = (bmax[None, :, :] - o[None, None, :]) * inv[:, None, :]
    lo, hi = torch.minimum(t1, t2), torch.maximum(t1, t2)
    tnear = lo.max(dim=2).values
    tfar, far_axis = hi.min(dim=2)
    return tnear, tfar, far_axis, d


def render(scene: Scene, cam: Camera, pose: np.ndarray, t_s: float = 0.0, device="cpu",
           dtype=torch.float64, extra_cuboid=None):
    """Returns (depth f32 HxW, label i32 HxW). depth = z-depth in metres, 0 where out of
    [min_range, max_range] (the normalised InputData depth/range image)."""
    W, H = cam.width, cam.height
    dev = torch.device(device)
    u = torch.arange(W, device=dev, dtype=dtype)
    v = torch.arange(H, device=dev, dtype=dtype)
    dx = ((u - cam.cx) / cam.fx)[None, :].expand(H, W)
    dy = ((v - cam.cy) / cam.fy)[:, None].expand(H, W)
    dC = torch.stack([dx, dy, torch.ones_like(dx)], dim=-1).reshape(-1, 3)
    T = torch.as_tensor(np.asarray(pose), device=dev, dtype=dtype)
    d = dC @ T[:3, :3].T
    d = torch.where(d.abs() < 1e-12, torch.full_like(d, 1e-12), d)
    o = T[:3, 3]

    # room: we are inside, hit the exit face
    rmin = torch.tensor([scene.room_min], device=dev, dtype=dtype)
    rmax = torch.tensor([scene.room_max], device=dev, dtype=dtype)
    _, tfar, far_axis, _ = _slab(o, d, rmin, rmax)
    depth = tfar[:, 0]
    ax = far_axis[:, 0]
    dsel = torch.gather(d, 1, ax[:, None])[:, 0]
    # labels: floor 1, ceiling 2, walls 3..6
    wall = 3 + ax * 2 + (dsel > 0).to(ax.dtype)
    label = torch.where(ax == 2, torch.where(dsel < 0, torch.ones_like(ax), 2 * torch.ones_like(ax)), wall)

    cubs = scene.cuboids
    labels = scene.cuboid_labels()
    if scene.mover is not None:
        size, start, vel, t_start = scene.mover
        c = np.asarray(start, dtype=np.float64) + np.asarray(vel, dtype=np.float64) * max(t_s - t_start, 0.0)
        hs = np.asarray(size, dtype=np.float64) / 2
        cubs = np.concatenate([cubs, np.concatenate([c - hs, c + hs])[None, :]], axis=0)
        labels = np.concatenate([labels, np.array([scene.num_labels - 1], np.int32)])
    if extra_cuboid is not None:  # per-frame scripted object (label L-1), (6,) min xyz max xyz
        cubs = np.concatenate([cubs, np.asarray(extra_cuboid, dtype=np.float64)[None, :]], axis=0)
        labels = np.concatenate([labels, np.array([scene.num_labels - 1], np.int32)])
    if len(cubs):
        cb = torch.as_tensor(cubs, device=dev, dtype=dtype)
        tnear, tfar_c, _, _ = _slab(o, d, cb[:, :3], cb[:, 3:])
        hit = (tnear <= tfar_c) & (tnear > 0)
        s = torch.where(hit, tnear, torch.full_like(tnear, float("inf")))
        smin, k = s.min(dim=1)
        closer = smin < depth
        depth = torch.where(closer, smin, depth)
        lab_c = torch.as_tensor(labels.astype(np.int64), device=dev)[k]
        label = torch.where(closer, lab_c, label)

    valid = (depth >= cam.min_range) & (depth <= cam.max_range)
    depth = torch.where(valid, depth, torch.zeros_like(depth))
    label = torch.where(valid, label, torch.zeros_like(label))
    return (depth.reshape(H, W).to(torch.float32).contiguous(),
            label.reshape(H, W).to(torch.int32).contiguous())


def colorize(label, depth=None) -> np.ndarray:
    """Deterministic textured RGB image (H, W, 3) u8 for a label image: a per-label base colour plus a
    pixel-dependent pattern, so that bilinear colour interpolation and blending are exercised. Pixels with
    invalid depth are black. Input generation only (InputData::color_image, CV_8UC3)."""
    lab = np.asarray(label).astype(np.int64)
    H, W = lab.shape
    v, u = np.meshgrid(np.arange(H, dtype=np.int64), np.arange(W, dtype=np.int64), indexing="ij")
    rgb = np.stack([(lab * 37 + 3 * u + 5 * v) % 256, (lab * 91 + 7 * u + v) % 256, (lab * 53 + u + 11 * v) % 256], axis=-1)
    if depth is not None:
        rgb = np.where((np.asarray(depth) > 0)[..., None], rgb, 0)
    return np.ascontiguousarray(rgb.astype(np.uint8))


def companion_cuboids(poses, start_frame=60, distance=2.2, size=(1.05, 1.05, 1.6), sway=0.8, period=120):
    """Config 3 dynamic object: an axis-aligned box that stays `distance` metres in front of the camera
    (swaying sideways), appearing after `start_frame` frames of burn-in. At 2.2 m a 1.05 x 1.6 m face covers
    roughly 20 % of a 640x480 / f=320 image. Returns a list of (6,) cuboids or None per frame."""
    out = []
    for i, T in enumerate(poses):
        if i < start_frame:
            out.append(None)
            continue
        T = np.asarray(T)
        lateral = sway * math.sin(2.0 * math.pi * (i - start_frame) / period)
        c = T[:3, 3] + T[:3, 2] * distance + T[:3, 0] * lateral
        c[2] = size[2] / 2 + 0.05
        h = np.asarray(size) / 2
        out.append(np.concatenate([c - h, c + h]))
    return out


def render_stream(scene, cam, poses, stamps, device="cpu", dtype=torch.float64, extra=None):
    """Renders all frames; returns depth (N,H,W) f32 and label (N,H,W) i32 tensors on `device`."""
    n = len(poses)
    depth = torch.empty((n, cam.height, cam.width), dtype=torch.float32, device=device)
    label = torch.empty((n, cam.height, cam.width), dtype=torch.int32, device=device)
    t0 = stamps[0]
    for i, (T, st) in enumerate(zip(poses, stamps)):
        d, l = render(scene, cam, T, (st - t0) * 1e-9, device=device, dtype=dtype,
                      extra_cuboid=None if extra is None else extra[i])
        depth[i], label[i] = d, l
    return depth, label
