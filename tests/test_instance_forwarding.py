"""khronos::InstanceForwarding (object_detection/instance_forwarding.cpp:80-149): the oracle against a numpy restatement of
the in-tree reference code (CPU), the product against the oracle (GPU)."""
import numpy as np
import pytest

from khronos_b200 import capi, synthetic as syn
import harness as hs


def instance_frame(scale=4):
    cam = hs.small_camera(scale)
    scene = syn.room_scene()
    pose = syn.look_pose((6.0, 5.0, 1.5), 3.7, np.radians(12.0))
    d, l = syn.render(scene, cam, pose)
    d, l = d.numpy().astype(np.float32), l.numpy().astype(np.int32).copy()
    # instance ids: the surface labels shifted into a sparse id space, some pixels unlabelled
    ids = np.where(l > 0, l * 37 % 300 + 1, 0).astype(np.int32)
    ids[::7, ::5] = 0
    return cam, pose, d, ids


def numpy_forward(cam, pose, d, ids, max_range, min_size, max_size, min_vol, max_vol, background):
    """Restatement of instance_forwarding.cpp:80-149 on arrays (vertex map = kb_compute_vertex_map's formula in fp32)."""
    f32 = np.float32
    H, W = ids.shape
    T = np.asarray(pose, np.float64).reshape(4, 4)
    R, t = T[:3, :3].astype(f32), T[:3, 3].astype(f32)
    u = np.arange(W, dtype=f32)[None, :].repeat(H, 0)
    v = np.arange(H, dtype=f32)[:, None].repeat(W, 1)
    pc = np.stack([(u - f32(cam.cx)) / f32(cam.fx) * d, (v - f32(cam.cy)) / f32(cam.fy) * d, d], -1).astype(f32)
    vm = np.empty_like(pc)
    for a in range(3):
        vm[..., a] = ((R[a, 0] * pc[..., 0] + R[a, 1] * pc[..., 1]).astype(f32) + R[a, 2] * pc[..., 2]).astype(f32) + t[a]
    out = []
    for i in sorted(set(np.unique(ids)) - {0}):
        if background is not None and i < len(background) and background[i]:
            continue
        m = ids == i
        if max_range > 0:
            m &= ~(d > f32(max_range))
        n = int(m.sum())
        if n == 0 or n < min_size or (max_size > 0 and n > max_size):
            continue
        vs, us = np.nonzero(m.T)  # transposed: row index of m.T is u -> column-major scan (u outer, v inner)
        pix = np.stack([vs, us], 1)  # (u, v)
        pts = vm[pix[:, 1], pix[:, 0]]
        bbox = np.concatenate([pts.min(0), pts.max(0)]).astype(f32)
        if min_vol > 0 or max_vol > 0:
            vol = f32(bbox[3] - bbox[0]) * f32(bbox[4] - bbox[1]) * f32(bbox[5] - bbox[2])
            if vol < min_vol or (max_vol > 0 and vol > max_vol):
                continue
        out.append({"id": int(i), "pixels": pix.astype(np.int32), "bbox": bbox})
    return out


CASES = [dict(), dict(max_range=3.5, min_cluster_size=40), dict(min_cluster_size=10, max_cluster_size=2500),
         dict(min_object_volume=1e-9, max_object_volume=5.0), dict(background=True, min_cluster_size=5)]


def _run(h, cam, pose, d, ids, kw):
    kw = dict(kw)
    bg = None
    if kw.pop("background", False):
        bg = np.zeros(400, np.uint8)
        bg[[int(x) for x in np.unique(ids)[1::3]]] = 1
    f = h.make_frame(d, pose, 1_000_000_000, label=ids)
    img, cl = h.forward_instances(f, background=bg, **kw)
    return img, cl, bg


def _assert_same(a, b):
    assert [c["id"] for c in a] == [c["id"] for c in b]
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x["pixels"], y["pixels"])
        np.testing.assert_array_equal(np.asarray(x["bbox"], np.float32).view(np.uint32), np.asarray(y["bbox"], np.float32).view(np.uint32))


@pytest.mark.parametrize("kw", CASES)
def test_oracle_matches_reference_restatement(oracle_lib, kw):
    cam, pose, d, ids = instance_frame()
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    img, cl, bg = _run(o, cam, pose, d, ids, kw)
    np.testing.assert_array_equal(img, ids)  # object_image shares the label image's buffer in the reference
    want = numpy_forward(cam, pose, d, ids, kw.get("max_range", 0.0), kw.get("min_cluster_size", 0), kw.get("max_cluster_size", -1),
                         kw.get("min_object_volume", 0.0), kw.get("max_object_volume", -1.0), bg)
    assert len(want) >= 1
    _assert_same(want, cl)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
def test_product_matches_oracle(oracle_lib, product_lib, kw):
    cam, pose, d, ids = instance_frame()
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    g = hs.make_handle(product_lib, "kb_", cam=cam)
    io, co, _ = _run(o, cam, pose, d, ids, kw)
    ig, cg, _ = _run(g, cam, pose, d, ids, kw)
    np.testing.assert_array_equal(io, ig)
    assert len(co) >= 1
    _assert_same(co, cg)


@pytest.mark.gpu
def test_product_rejects_out_of_range_ids(product_lib):
    cam, pose, d, ids = instance_frame()
    ids = ids.copy()
    ids[3, 3] = 5000
    g = hs.make_handle(product_lib, "kb_", cam=cam)
    with pytest.raises(capi.KbError):
        g.forward_instances(g.make_frame(d, pose, 1_000_000_000, label=ids))
