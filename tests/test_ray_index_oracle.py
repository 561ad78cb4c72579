"""CPU: the oracle's restatement of khronos::RayVerificator (backend/change_detection/ray_verificator.cpp: addRayToHash
:326-350, check :66-146) against an independent numpy restatement and hand-computed cases. The reference has no tests
for it. tests/test_zz_ray_index.py compares the product with this oracle."""
import numpy as np
import pytest

from khronos_b200 import capi

f32 = np.float32


def np_norm(v):
    return np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])


def np_normalized(v):
    n = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]
    return v / np.sqrt(n) if n > 0 else v


def np_march(src, dst, block_size):
    """Blocks of one ray, in visiting order without repeats (float32 throughout)."""
    src, dst = np.asarray(src, f32), np.asarray(dst, f32)
    d = dst - src
    direction, max_depth = np_normalized(d), np_norm(d)
    step, inv = f32(block_size) / f32(4), f32(1) / f32(block_size)
    dist, out = f32(0), []
    while dist <= max_depth:
        dist = f32(dist + step)
        p = src + dist * direction
        b = tuple(int(v) for v in np.floor(p * inv))
        if b not in out:
            out.append(b)
    return out


class NumpyRays:
    def __init__(self, cfg):
        self.cfg, self.src, self.dst, self.ts, self.blocks = cfg, [], [], [], {}

    def add(self, s, t, ts):
        obs = set()
        for a, b, c in zip(np.asarray(s, f32).reshape(-1, 3), np.asarray(t, f32).reshape(-1, 3), ts):
            i = len(self.src)
            self.src.append(a); self.dst.append(b); self.ts.append(int(c))
            for blk in np_march(a, b, self.cfg.block_size):
                self.blocks.setdefault(blk, set()).add(i)
                obs.add(blk)
        return np.array(sorted(obs, key=lambda p: (p[2], p[1], p[0])), np.int32).reshape(-1, 3)

    def check(self, points, earliest, latest):
        res = []
        inv = f32(1) / f32(self.cfg.block_size)
        rt, dt = f32(self.cfg.radial_tolerance), f32(self.cfg.depth_tolerance)
        pts = np.asarray(points, f32).reshape(-1, 3)
        lo = np.broadcast_to(np.asarray(earliest, np.uint64), (len(pts),))
        hi = np.broadcast_to(np.asarray(latest, np.uint64), (len(pts),))
        for p, e, l in zip(pts, lo, hi):
            absent, present = [], []
            for i in self.blocks.get(tuple(int(v) for v in np.floor(p * inv)), ()):
                if self.ts[i] < e or self.ts[i] > l:
                    continue
                ps = p - self.src[i]
                with np.errstate(invalid="ignore", divide="ignore"):
                    direction, depth = np_normalized(ps), np_norm(ps)
                    radial = np_norm(np.cross(ps, self.src[i] - self.dst[i]).astype(f32)) / depth
                if radial > rt:
                    continue
                vs = self.dst[i] - self.src[i]
                dd = (vs[0] * direction[0] + vs[1] * direction[1]) + vs[2] * direction[2]
                if depth - dd > dt:
                    continue
                (absent if dd - depth > dt else present).append(self.ts[i])
            res.append((np.array(sorted(absent), np.uint64), np.array(sorted(present), np.uint64)))
        return res


def random_rays(rng, n, n_poses=6):
    """Sensor positions along a short path, targets on the walls of a 10 x 8 x 3 m room around it."""
    poses = np.stack([np.linspace(2, 8, n_poses), np.linspace(3, 5, n_poses), np.full(n_poses, 1.2)], 1).astype(f32)
    k = rng.integers(0, n_poses, n)
    tgt = rng.uniform([0, 0, 0], [10, 8, 3], (n, 3)).astype(f32)
    wall = rng.integers(0, 3, n)
    side = rng.integers(0, 2, n)
    for a, hi in enumerate((10.0, 8.0, 3.0)):
        tgt[wall == a, a] = np.where(side[wall == a] == 1, hi, 0.0)
    stamps = (1_000_000_000 + k.astype(np.uint64) * np.uint64(500_000_000)).astype(np.uint64)
    return poses[k], tgt, stamps


def query_points(rng, src, tgt, n):
    """Points on rays (present), in front of the hit (absent), behind it (occluded), beside the ray and random."""
    i = rng.integers(0, len(src), n)
    u = rng.uniform(0.05, 1.3, n).astype(f32)[:, None]
    pts = src[i] + u * (tgt[i] - src[i])
    pts[::3] += rng.normal(0, 0.05, pts[::3].shape).astype(f32)
    pts[::7] = rng.uniform([0, 0, 0], [10, 8, 3], pts[::7].shape)
    pts[::5] = tgt[i[::5]]
    return pts.astype(f32)


def compare_checks(got, want):
    counts, lists = got
    assert len(lists) == len(want)
    for k, ((a, p), (wa, wp)) in enumerate(zip(lists, want)):
        np.testing.assert_array_equal(a, wa, err_msg=f"absent stamps of point {k}")
        np.testing.assert_array_equal(p, wp, err_msg=f"present stamps of point {k}")
        assert (counts[k, 0], counts[k, 1]) == (len(wa), len(wp))


def test_hand_computed_ray(oracle_lib):
    """One ray along +x from (0.5, 0.5, 0.5) to (3.6, 0.5, 0.5), 1 m blocks, step 0.25: samples at x = 0.75 ... 3.75
    (the loop steps once past max_depth = 3.1 -> 3.75 is the last sample), so blocks x = 0..3."""
    r = capi.RayIndex(oracle_lib, "ko_", capi.default_ray_config())
    obs = r.add([[0.5, 0.5, 0.5]], [[3.6, 0.5, 0.5]], [100])
    np.testing.assert_array_equal(obs, [[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0]])
    assert r.size() == (1, 4)
    pts = [[3.6, 0.5, 0.5],     # the hit itself: present
           [2.0, 0.53, 0.5],    # 3 cm beside the ray, 1.6 m before the hit: the hit is 3.1 * sin(atan(.03/1.5)) = 6 cm from
                                #   the line source -> point (that is what :101 measures): seen through -> absent
           [2.0, 0.7, 0.5],     # 20 cm beside the ray: no overlap
           [3.65, 0.5, 0.5],    # 5 cm behind the hit: within the depth tolerance -> present
           [3.9, 0.5, 0.5],     # 30 cm behind the hit: occluded
           [2.0, 0.5, 1.5]]     # another block: unobserved
    counts, lists = r.check(pts)
    np.testing.assert_array_equal(counts, [[0, 1], [1, 0], [0, 0], [0, 1], [0, 0], [0, 0]])
    assert lists[0][1][0] == 100 and lists[1][0][0] == 100
    counts, _ = r.check(pts, earliest=101)          # outside the time window
    assert not counts.any()
    counts, _ = r.check(pts, earliest=[0, 0, 0, 0, 0, 0], latest=[99, 100, 100, 100, 100, 100])
    np.testing.assert_array_equal(counts, [[0, 0], [1, 0], [0, 0], [0, 1], [0, 0], [0, 0]])
    # deformation: the hit moves 1 m closer; the hash is unchanged, the verdicts follow the new endpoint
    r.set_endpoints([[0.5, 0.5, 0.5]], [[2.6, 0.5, 0.5]])
    counts, _ = r.check(pts)
    np.testing.assert_array_equal(counts, [[0, 0], [1, 0], [0, 0], [0, 0], [0, 0], [0, 0]])
    assert r.size() == (1, 4)
    r.rehash()
    assert r.size() == (1, 3)   # samples up to x = 2.75
    with pytest.raises(capi.KbError):
        capi.RayIndex(oracle_lib, "ko_", capi.default_ray_config(block_size=0.0))


@pytest.mark.parametrize("block_size,radial,depth_tol", [(1.0, 0.1, 0.1), (0.5, 0.05, 0.2), (2.0, 0.3, 0.05)])
def test_oracle_matches_numpy(oracle_lib, block_size, radial, depth_tol):
    rng = np.random.default_rng(4)
    cfg = capi.default_ray_config(block_size, radial, depth_tol)
    r, ref = capi.RayIndex(oracle_lib, "ko_", cfg), NumpyRays(cfg)
    src, tgt, ts = random_rays(rng, 300)
    for a, b in ((0, 120), (120, 300)):     # incremental, as updateDsg
        np.testing.assert_array_equal(r.add(src[a:b], tgt[a:b], ts[a:b]), ref.add(src[a:b], tgt[a:b], ts[a:b]))
    assert r.size() == (300, sum(len(v) for v in ref.blocks.values()))
    pts = query_points(rng, src, tgt, 400)
    got = r.check(pts)
    compare_checks(got, ref.check(pts, 0, 2**64 - 1))
    assert got[0][:, 0].sum() > 20 and got[0][:, 1].sum() > 20
    lo = rng.integers(1_000_000_000, 2_500_000_000, len(pts)).astype(np.uint64)
    hi = lo + np.uint64(1_200_000_000)
    compare_checks(r.check(pts, lo, hi), ref.check(pts, lo, hi))
    # degenerate rays and points: zero-length ray, point at a source
    r2, ref2 = capi.RayIndex(oracle_lib, "ko_", cfg), NumpyRays(cfg)
    s2, t2 = np.array([[1, 1, 1], [2, 2, 1]], f32), np.array([[1, 1, 1], [5, 2, 1]], f32)
    np.testing.assert_array_equal(r2.add(s2, t2, [5, 6]), ref2.add(s2, t2, [5, 6]))
    p2 = np.array([[1, 1, 1], [2, 2, 1], [3, 2, 1]], f32)
    compare_checks(r2.check(p2), ref2.check(p2, 0, 2**64 - 1))


def np_vertex_sources(policy, stamps, first, last):
    """computeVertexSources (ray_verificator.cpp:278-330) with numpy's searchsorted (upper_bound = side 'right',
    lower_bound = side 'left'); the resulting set ascending."""
    n, out = len(stamps), set()
    def take(i):
        if i < n:
            out.add(int(i))
    if policy in (capi.RAYS_FIRST, capi.RAYS_FIRST_AND_LAST):
        take(np.searchsorted(stamps, first, "right"))
    if policy in (capi.RAYS_LAST, capi.RAYS_FIRST_AND_LAST):
        take(np.searchsorted(stamps, last, "left"))
    if policy == capi.RAYS_MIDDLE:
        take(np.searchsorted(stamps, np.uint64((int(last) + int(first)) % 2**64 // 2), "left"))
    if policy == capi.RAYS_ALL:
        out.update(range(int(np.searchsorted(stamps, first, "right")), int(np.searchsorted(stamps, last, "left"))))
    return sorted(out)


def mesh_scenario(rng, n_poses=10, n_vertices=200):
    stamps = (np.uint64(1_000_000_000) + np.arange(n_poses, dtype=np.uint64) * np.uint64(400_000_000))
    poses = np.stack([np.linspace(2, 8, n_poses), np.linspace(3, 5, n_poses), np.full(n_poses, 1.2)], 1).astype(f32)
    _, verts, _ = random_rays(rng, n_vertices)
    first = rng.integers(500_000_000, 4_500_000_000, n_vertices).astype(np.uint64)
    last = first + rng.integers(0, 2_000_000_000, n_vertices).astype(np.uint64)
    first[::11] = stamps[rng.integers(0, n_poses, len(first[::11]))]     # exactly on a pose stamp: upper vs lower bound
    last[::13] = stamps[rng.integers(0, n_poses, len(last[::13]))]
    return stamps, poses, verts, first, last


@pytest.mark.parametrize("policy", [capi.RAYS_FIRST, capi.RAYS_LAST, capi.RAYS_FIRST_AND_LAST, capi.RAYS_MIDDLE, capi.RAYS_ALL])
@pytest.mark.parametrize("aw", [0.0, 0.75])
def test_oracle_add_vertices_matches_numpy(oracle_lib, policy, aw):
    rng = np.random.default_rng(policy)
    stamps, poses, verts, first, last = mesh_scenario(rng)
    cfg = capi.default_ray_config()
    r, ref = capi.RayIndex(oracle_lib, "ko_", cfg), NumpyRays(cfg)
    # first half of the mesh, then the rest (updateDsg), with more poses known by then
    obs_all = []
    for a, b, npz in ((0, 90, 7), (90, 200, 10)):
        offset = np.uint64(int(np.float32(aw) * 1e9)) if aw > 0 else np.uint64(0)
        want_s, want_t, want_ts, want_pose, want_vert = [], [], [], [], []
        for v in range(a, b):
            with np.errstate(over="ignore"):
                shifted = np.uint64(last[v] - offset)
            for k in np_vertex_sources(policy, stamps[:npz], first[v], shifted):
                want_s.append(poses[k]); want_t.append(verts[v]); want_ts.append(stamps[k]); want_pose.append(k); want_vert.append(v)
        obs, nadd = r.add_vertices(policy, stamps[:npz], poses[:npz], verts[a:b], first[a:b], last[a:b], vertex_index_base=a,
                                   active_window_duration=aw)
        assert nadd == len(want_ts)
        want_obs = ref.add(np.array(want_s, f32).reshape(-1, 3), np.array(want_t, f32).reshape(-1, 3), want_ts)
        np.testing.assert_array_equal(obs, want_obs)
        obs_all.append((want_pose, want_vert, want_ts))
    pose, vert, ts = r.ray_ids()
    np.testing.assert_array_equal(pose, np.concatenate([np.array(o[0], np.int32) for o in obs_all]))
    np.testing.assert_array_equal(vert, np.concatenate([np.array(o[1], np.int32) for o in obs_all]))
    np.testing.assert_array_equal(ts, np.concatenate([np.array(o[2], np.uint64) for o in obs_all]))
    assert len(ts) > 60
    pts = query_points(rng, np.array(ref.src), np.array(ref.dst), 200)
    compare_checks(r.check(pts), ref.check(pts, 0, 2**64 - 1))
    # plain rays carry no ids; rehash keeps them
    r.add([[0, 0, 0]], [[1, 1, 1]], [7])
    r.rehash()
    pose2, vert2, _ = r.ray_ids()
    assert pose2[-1] == -1 and vert2[-1] == -1
    np.testing.assert_array_equal(pose2[:-1], pose)
    with pytest.raises(capi.KbError):
        r.add_vertices(5, stamps, poses, verts, first, last)             # random policies are not offered
    with pytest.raises(capi.KbError):
        r.add_vertices(policy, stamps[::-1].copy(), poses, verts, first, last)
