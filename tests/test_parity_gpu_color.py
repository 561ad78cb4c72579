"""GPU parity of the colour path of K1 (TsdfVoxel::color; ORACLE_SPEC §5.6): interpolateColor + Color::merge in the
fuse kernel's COLOR instantiations vs the CPU oracle, through the C ABI. u8 colours must be bit-exact."""
import numpy as np
import pytest

from khronos_b200 import capi, synthetic as syn
import harness as hs
from test_parity_gpu import both, room_frames

pytestmark = pytest.mark.gpu


def colours_of(frames):
    return [syn.colorize(l, d) for d, l in frames]


@pytest.mark.parametrize("interp", [capi.INTERP_ADAPTIVE, capi.INTERP_NEAREST, capi.INTERP_BILINEAR])
def test_colour_fusion_per_frame(oracle_lib, product_lib, interp):
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 8)
    cols = colours_of(frames)
    o, g = both(oracle_lib, product_lib, cam=cam, integ_cfg=capi.default_integrator_config(interpolation=interp))
    so = hs.run_fusion(o, frames, poses, stamps, tracking=True, colors=cols)
    sg = hs.run_fusion(g, frames, poses, stamps, tracking=True, colors=cols)
    assert so == sg
    bo, bg = o.export_blocks(), g.export_blocks()
    assert bo.color.any()
    hs.assert_blocks_equal(bo, bg, exact_float=True, what=f"colour interp{interp}")


@pytest.mark.parametrize("batch", [5, 32])
def test_colour_batched_with_gaps_and_culling(oracle_lib, product_lib, batch):
    """kb_integrate_frames with colour on some frames only (every third frame has no colour image), culling on."""
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 36, laps=0.4)
    cols = [None if i % 3 == 1 else c for i, c in enumerate(colours_of(frames))]
    o, g = both(oracle_lib, product_lib, cam=cam)
    g.set_culling(2)
    hs.run_fusion(o, frames, poses, stamps, colors=cols)
    for i in range(0, len(frames), batch):
        fr = [g.make_frame(d, T, st, label=l, color=c) for (d, l), T, st, c in
              zip(frames[i:i + batch], poses[i:i + batch], stamps[i:i + batch], cols[i:i + batch])]
        g.integrate_frames(fr)
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what=f"colour batch{batch}")


def test_colour_appears_late_and_device_frames(oracle_lib, product_lib):
    """The colour layer is allocated with the first coloured frame: earlier colour-less frames leave black voxels;
    device-resident RGB images are read in place."""
    import torch
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 9)
    cols = [None] * 4 + colours_of(frames)[4:]
    o, g = both(oracle_lib, product_lib, cam=cam)
    hs.run_fusion(o, frames, poses, stamps, colors=cols)
    keep = []
    for (d, l), T, st, c in zip(frames, poses, stamps, cols):
        dd, ll = torch.from_numpy(d).cuda(), torch.from_numpy(l).cuda()
        cc = None if c is None else torch.from_numpy(c).cuda()
        keep.append((dd, ll, cc))
        torch.cuda.synchronize()
        g.integrate_frame(g.make_frame(dd, T, st, label=ll, color=cc, memory=capi.MEM_DEVICE), want_stats=False)
    g.synchronize()
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="late colour")


def test_colour_on_extraction_map_and_block_reuse(oracle_lib, product_lib):
    """vps = 8 binary map without tracking (the extractor's private map) with colour; and removal + re-allocation
    of a coloured block starts from black again."""
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 6)
    cols = colours_of(frames)
    mc = capi.default_map_config(voxel_size=0.1, vps=8, trunc=0.2, with_tracking=False, max_blocks=16384)
    ic = capi.default_integrator_config(semantic_mode=capi.SEM_BINARY)
    o, g = both(oracle_lib, product_lib, cam=cam, map_cfg=mc, integ_cfg=ic)
    for (d, l), T, st, c in zip(frames, poses, stamps, cols):
        for h in (o, g):
            h.integrate_frame(h.make_frame(d, T, st, object_image=l, target_id=7, color=c))
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="colour vps8")
    # tracking map: blocks leave the temporal window, get removed, and are re-observed later
    frames, poses, stamps = room_frames(cam, 14, laps=0.5, dt_ns=500_000_000)
    frames, poses = frames + frames[:4], poses + poses[:4]
    stamps = stamps + [stamps[-1] + (k + 1) * 500_000_000 for k in range(4)]
    cols = colours_of(frames)
    o, g = both(oracle_lib, product_lib, cam=cam)
    removed = 0
    for i, ((d, l), T, st, c) in enumerate(zip(frames, poses, stamps, cols)):
        for h in (o, g):
            h.integrate_frame(h.make_frame(d, T, st, label=l, color=c))
            h.update_tracking(st)
        if i % 4 == 3:
            ro, rg = o.reset_inactive(), g.reset_inactive()
            np.testing.assert_array_equal(ro, rg)
            removed += len(ro)
    assert removed > 0
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="colour reuse")


def test_product_matches_golden_colour(product_lib):
    from test_golden import run_colour_case
    run_colour_case(product_lib, "kb_")
