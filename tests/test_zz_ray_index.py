"""GPU parity of the ray index (kb_rays_*: khronos::RayVerificator on the device) against the CPU
oracle, which tests/test_ray_index_oracle.py pins against a numpy restatement of ray_verificator.cpp. Observed blocks,
block-entry counts, per-point absent / present counts and the stamp lists must be identical (the classification is a
chain of fp32 comparisons: any arithmetic difference would flip verdicts)."""
import numpy as np
import pytest

from khronos_b200 import capi
from test_ray_index_oracle import compare_checks, query_points, random_rays

pytestmark = pytest.mark.gpu
f32 = np.float32


def both(oracle_lib, product_lib, cfg):
    return capi.RayIndex(oracle_lib, "ko_", cfg), capi.RayIndex(product_lib, "kb_", cfg)


def check_same(o, g, pts, lo=0, hi=2**64 - 1):
    co, lo_ = o.check(pts, lo, hi)
    cg, lg = g.check(pts, lo, hi)
    np.testing.assert_array_equal(co, cg)
    compare_checks((cg, lg), lo_)
    return co


@pytest.mark.parametrize("block_size,radial,depth_tol", [(1.0, 0.1, 0.1), (0.5, 0.05, 0.2), (2.0, 0.3, 0.05)])
def test_ray_index_matches_oracle(oracle_lib, product_lib, block_size, radial, depth_tol):
    rng = np.random.default_rng(8)
    o, g = both(oracle_lib, product_lib, capi.default_ray_config(block_size, radial, depth_tol))
    src, tgt, ts = random_rays(rng, 3000, n_poses=12)
    for a, b in ((0, 1), (1, 700), (700, 3000)):   # incremental updates (updateDsg), buffers grow
        np.testing.assert_array_equal(o.add(src[a:b], tgt[a:b], ts[a:b]), g.add(src[a:b], tgt[a:b], ts[a:b]))
        assert o.size() == g.size()
        pts = query_points(rng, src[:b], tgt[:b], 500)
        check_same(o, g, pts)
    pts = query_points(rng, src, tgt, 4000)
    c = check_same(o, g, pts)
    assert c[:, 0].sum() > 200 and c[:, 1].sum() > 200
    lo = rng.integers(1_000_000_000, 5_000_000_000, len(pts)).astype(np.uint64)
    check_same(o, g, pts, lo, lo + np.uint64(2_000_000_000))
    # deformation (loop closure): endpoints move, the hash stays; then recomputeHash
    src2 = src + rng.normal(0, 0.05, src.shape).astype(f32)
    tgt2 = tgt + rng.normal(0, 0.15, tgt.shape).astype(f32)
    o.set_endpoints(src2, tgt2); g.set_endpoints(src2, tgt2)
    check_same(o, g, pts)
    o.rehash(); g.rehash()
    assert o.size() == g.size()
    check_same(o, g, pts)
    o.clear(); g.clear()
    assert g.size() == (0, 0)
    assert not check_same(o, g, pts[:50]).any()


def test_ray_index_degenerate_inputs(oracle_lib, product_lib):
    o, g = both(oracle_lib, product_lib, capi.default_ray_config())
    pts = np.array([[1, 1, 1], [2, 2, 1], [3, 2, 1], [-4.5, -0.2, 0.3]], f32)
    assert not check_same(o, g, pts).any()                       # no rays yet
    s = np.array([[1, 1, 1], [2, 2, 1], [-0.5, -0.5, 0.5], [0, 0, 0]], f32)
    t = np.array([[1, 1, 1], [5, 2, 1], [-7.5, 0.1, 0.2], [0, 0, 40.0]], f32)   # zero-length ray, negative blocks, long ray
    np.testing.assert_array_equal(o.add(s, t, [5, 6, 7, 8]), g.add(s, t, [5, 6, 7, 8]))
    assert o.size() == g.size()
    check_same(o, g, pts)                                        # includes points at ray sources (depth 0 -> NaN tests)
    check_same(o, g, np.zeros((0, 3), f32))
    many = np.repeat(pts, 300, 0)                                # more points than warps in flight
    check_same(o, g, many)
    # many rays through one block: lists longer than a warp
    rng = np.random.default_rng(0)
    s = np.tile(np.array([[0.5, 0.5, 0.5]], f32), (200, 1))
    t = s + rng.normal(0, 1, (200, 3)).astype(f32) * f32(3)
    ts = rng.integers(1, 50, 200).astype(np.uint64)
    np.testing.assert_array_equal(o.add(s, t, ts), g.add(s, t, ts))
    q = (s + f32(0.8) * (t - s)).astype(f32)
    c = check_same(o, g, q)
    assert c.sum() > 100
    with pytest.raises(capi.KbError):
        g.add([[0, 0, np.nan]], [[1, 1, 1]], [1])
    with pytest.raises(capi.KbError):
        g.set_endpoints(s[:3], t[:3])
    with pytest.raises(capi.KbError):
        capi.RayIndex(product_lib, "kb_", capi.default_ray_config(radial_tolerance=0.0))


@pytest.mark.parametrize("policy", [capi.RAYS_FIRST_AND_LAST, capi.RAYS_MIDDLE, capi.RAYS_ALL])
def test_ray_index_add_vertices_matches_oracle(oracle_lib, product_lib, policy):
    """addVertices: rays chosen per mesh vertex from the pose stamps (computeVertexSources), then hashed."""
    from test_ray_index_oracle import mesh_scenario
    rng = np.random.default_rng(20 + policy)
    stamps, poses, verts, first, last = mesh_scenario(rng, n_poses=14, n_vertices=1500)
    o, g = both(oracle_lib, product_lib, capi.default_ray_config())
    for a, b, npz, aw in ((0, 600, 9, 0.0), (600, 1500, 14, 0.5)):
        ro = o.add_vertices(policy, stamps[:npz], poses[:npz], verts[a:b], first[a:b], last[a:b], a, aw)
        rg = g.add_vertices(policy, stamps[:npz], poses[:npz], verts[a:b], first[a:b], last[a:b], a, aw)
        np.testing.assert_array_equal(ro[0], rg[0])
        assert ro[1] == rg[1] and o.size() == g.size()
    for x, y in zip(o.ray_ids(), g.ray_ids()):
        np.testing.assert_array_equal(x, y)
    pose, vert, _ = g.ray_ids()
    pts = (verts[rng.integers(0, len(verts), 2000)] + rng.normal(0, 0.03, (2000, 3))).astype(f32)
    c = check_same(o, g, pts)
    assert c.sum() > 300
    # deformation through the ids: every pose / vertex moves, endpoints are re-gathered by the caller
    poses2 = poses + rng.normal(0, 0.05, poses.shape).astype(f32)
    verts2 = verts + rng.normal(0, 0.1, verts.shape).astype(f32)
    o.set_endpoints(poses2[pose], verts2[vert]); g.set_endpoints(poses2[pose], verts2[vert])
    check_same(o, g, pts)
    g.rehash(); o.rehash()
    for x, y in zip(o.ray_ids(), g.ray_ids()):
        np.testing.assert_array_equal(x, y)
    check_same(o, g, pts)


def test_adaptor_track_measurements_and_ray_verificator(tmp_path):
    """measureTracks and GpuRayVerificator through the C++ host adaptor (known answers on the flat-wall frame)."""
    import os
    import subprocess
    from harness import ROOT
    csrc = os.path.join(ROOT, "khronos_b200", "csrc")
    exe = str(tmp_path / "adaptor_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "adaptor_compile_check.cpp"),
                           "-o", exe, "-L", csrc, "-lkhronos_b200", f"-Wl,-rpath,{csrc}"])
    out = subprocess.run([exe, "require-gpu", "objects", "core", "rays"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    fields = dict(kv.split("=") for kv in out.stdout.split())
    # wall at z = 2 m, 64 x 48 px at f = 32: 4 m x 3 m -> 40 x 30 voxels of 0.1 m, half of them per cluster
    a, b = (int(v) for v in fields["track_counts"].split(","))
    assert a == b == 20 * 30
    assert fields["track_iou"] == "1.000,0.000"
    # ray (0,0,0) -> (0,0,2), 1 m blocks, step 0.25: samples z = 0.25 .. 2.25 -> blocks z = 0, 1, 2
    assert int(fields["ray_blocks"]) == 3
    assert fields["ray_verdicts"] == "01,10,00"    # hit: present; half way: absent; 0.5 m behind: occluded


def test_ray_index_table_growth(oracle_lib, product_lib):
    """10 cm blocks: tens of thousands of distinct blocks, so the block table (sized by distinct blocks, starting at 2^14
    slots) has to grow and the count pass is repeated."""
    rng = np.random.default_rng(5)
    o, g = both(oracle_lib, product_lib, capi.default_ray_config(0.1, 0.05, 0.05))
    src, tgt, ts = random_rays(rng, 1500, n_poses=8)
    np.testing.assert_array_equal(o.add(src, tgt, ts), g.add(src, tgt, ts))
    assert o.size() == g.size() and g.size()[1] > 40_000
    pts = query_points(rng, src, tgt, 1500)
    c = check_same(o, g, pts)
    assert c.sum() > 50
    src2, tgt2, ts2 = random_rays(rng, 1500, n_poses=8)     # second update: grows again or reuses the grown table
    np.testing.assert_array_equal(o.add(src2, tgt2, ts2), g.add(src2, tgt2, ts2))
    check_same(o, g, pts)
