"""Shared scenario drivers for parity tests: the same calls are issued to the CPU oracle (ko_) and the
CUDA product (kb_) through the identical C ABI, and the exported maps are compared."""
import ctypes
import os

import numpy as np

from khronos_b200 import capi, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# Oracle worker threads for handles built with default configs: small test inputs gain nothing from one thread per
# core (the oracle spawns its workers per call like the reference), and CI containers oversubscribe badly.
TEST_THREADS = int(os.environ.get("KB_TEST_THREADS", "4"))


def make_handle(lib, prefix, map_cfg=None, integ_cfg=None, trk_cfg="default", mot_cfg="default", cam=None, device=0):
    map_cfg = map_cfg or capi.default_map_config()
    integ_cfg = integ_cfg or capi.default_integrator_config(num_threads=TEST_THREADS)
    trk = capi.default_tracking_config(num_threads=TEST_THREADS) if trk_cfg == "default" else trk_cfg
    mot = capi.default_motion_config(num_threads=TEST_THREADS) if mot_cfg == "default" else mot_cfg
    if not map_cfg.with_tracking:
        trk, mot = None, None
    h = capi.MapHandle(lib, prefix, map_cfg, integ_cfg, trk, mot, device)
    h.set_camera(cam or syn.make_camera())
    return h


def small_camera(scale=4, max_range=5.0):
    """640x480 jackal-like camera scaled down for fast CPU runs."""
    return syn.make_camera(640 // scale, 480 // scale, 320.0 / scale, 320.0 / scale, max_range=max_range)


def render_frames(scene, cam, poses, stamps):
    out = []
    t0 = stamps[0]
    for T, st in zip(poses, stamps):
        d, l = syn.render(scene, cam, T, (st - t0) * 1e-9)
        out.append((d.numpy(), l.numpy()))
    return out


def run_fusion(h, frames, poses, stamps, tracking=False, masks=None, colors=None):
    stats = []
    for i, ((d, l), T, st) in enumerate(zip(frames, poses, stamps)):
        m = None if masks is None else masks[i]
        c = None if colors is None else colors[i]
        f = h.make_frame(d, T, st, label=l, mask=m, color=c)
        stats.append(h.integrate_frame(f).as_dict())
        if tracking:
            h.update_tracking(st)
    return stats


def assert_blocks_equal(a: capi.Blocks, b: capi.Blocks, rtol=1e-4, exact_float=False, what=""):
    """a = oracle, b = product. Integer fields bit-exact; TSDF within rtol (north_star: 1e-4 rel)."""
    assert a.n == b.n, f"{what}: block count {a.n} vs {b.n}"
    np.testing.assert_array_equal(a.block_index, b.block_index, err_msg=f"{what} block_index")
    np.testing.assert_array_equal(a.block_flags, b.block_flags, err_msg=f"{what} block_flags")
    for name in ("last_observed", "last_occupied", "ever_free", "active", "to_remove",
                 "semantic_label", "semantic_empty", "color"):
        np.testing.assert_array_equal(getattr(a, name), getattr(b, name), err_msg=f"{what} {name}")
    if exact_float:
        np.testing.assert_array_equal(a.distance.view(np.uint32), b.distance.view(np.uint32), err_msg=f"{what} distance bits")
        np.testing.assert_array_equal(a.weight.view(np.uint32), b.weight.view(np.uint32), err_msg=f"{what} weight bits")
    np.testing.assert_allclose(b.distance, a.distance, rtol=rtol, atol=1e-7, equal_nan=True, err_msg=f"{what} distance")
    np.testing.assert_allclose(b.weight, a.weight, rtol=rtol, atol=0, equal_nan=True, err_msg=f"{what} weight")
    if a.semantic_likelihoods is not None and b.semantic_likelihoods is not None:
        np.testing.assert_allclose(b.semantic_likelihoods, a.semantic_likelihoods, rtol=rtol, atol=1e-6,
                                   err_msg=f"{what} likelihoods")


_M1, _M2 = np.uint64(0xff51afd7ed558ccd), np.uint64(0xc4ceb9fe1a85ec53)


def _mix64(k):
    k = k.astype(np.uint64, copy=True)
    k ^= k >> np.uint64(33)
    k *= _M1
    k ^= k >> np.uint64(33)
    k *= _M2
    k ^= k >> np.uint64(33)
    return k


def map_checksum(b: capi.Blocks):
    """kb_map_checksum (include/khronos_b200.h) restated in numpy over an exported map (product or oracle):
    (sum mod 2^64, xor, blocks, voxels observed at least once)."""
    if b.n == 0:
        return (0, 0, 0, 0)
    V = b.distance.shape[1] if b.distance.ndim == 2 else b.distance.size // b.n
    bi = b.block_index.reshape(b.n, 3).astype(np.int64)
    o, m = np.int64(1 << 20), np.uint64((1 << 21) - 1)
    key = (((bi[:, 0] + o).astype(np.uint64) & m) | (((bi[:, 1] + o).astype(np.uint64) & m) << np.uint64(21)) |
           (((bi[:, 2] + o).astype(np.uint64) & m) << np.uint64(42)))
    lin = _mix64(np.arange(V, dtype=np.uint64) + np.uint64(1))
    with np.errstate(over="ignore"):
        v = _mix64(key[:, None] ^ lin[None, :])
        d = np.ascontiguousarray(b.distance, np.float32).reshape(b.n, V).view(np.uint32).astype(np.uint64)
        w = np.ascontiguousarray(b.weight, np.float32).reshape(b.n, V).view(np.uint32).astype(np.uint64)
        v = _mix64(v ^ (d | (w << np.uint64(32))))
        label = np.where(b.semantic_empty.reshape(b.n, V) != 0, np.uint64(0xFFFFFFFF), b.semantic_label.reshape(b.n, V).astype(np.uint64))
        v = _mix64(v ^ label)
        stamp = b.last_observed.reshape(b.n, V).astype(np.uint64)
        v = _mix64(v ^ stamp)
        total = int(np.add.reduce(v.ravel(), dtype=np.uint64))
    xr = int(np.bitwise_xor.reduce(v.ravel()))
    seen = int(((stamp != 0) | (b.weight.reshape(b.n, V) > 0)).sum())
    return (total, xr, int(b.n), seen)


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
