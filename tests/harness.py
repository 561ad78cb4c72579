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


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
