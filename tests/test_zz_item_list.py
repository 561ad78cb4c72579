"""KB_FUSE_ITEM_LIST=1 (experiment, off by default): the fuse kernel takes its items from compacted heaviest-first
lists instead of the dense box range. Only the processing order changes, so every result must stay bit-identical."""
import os

import numpy as np
import pytest

from khronos_b200 import capi, synthetic as syn
import harness as hs

pytestmark = pytest.mark.gpu


@pytest.fixture
def item_list_env():
    os.environ["KB_FUSE_ITEM_LIST"] = "1"   # read by kb_create
    yield
    os.environ.pop("KB_FUSE_ITEM_LIST", None)


@pytest.mark.parametrize("vps,batch", [(16, 32), (16, 11), (8, 32)])
def test_item_list_variant_is_bit_identical(oracle_lib, product_lib, item_list_env, vps, batch):
    cam = hs.small_camera(4)
    scene = syn.hall_scene(size=(20.0, 16.0, 6.0))
    poses, stamps = syn.sweep_trajectory(40, size=(20.0, 16.0), margin=4.0, lanes=2, yaw_turns=1.5)
    frames = hs.render_frames(scene, cam, poses, stamps)
    mc = capi.default_map_config(voxel_size=0.05 if vps == 16 else 0.1, vps=vps, trunc=0.15 if vps == 16 else 0.3, max_blocks=16384)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam, map_cfg=mc)
    g = hs.make_handle(product_lib, "kb_", cam=cam, map_cfg=mc)
    so = hs.run_fusion(o, frames, poses, stamps)
    tot = {k: 0 for k in so[0]}
    for i in range(0, len(frames), batch):
        fr = [g.make_frame(d, T, st, label=l) for (d, l), T, st in zip(frames[i:i + batch], poses[i:i + batch], stamps[i:i + batch])]
        s = g.integrate_frames(fr).as_dict()
        want = {k: sum(x[k] for x in so[i:i + batch]) for k in s}
        want["total_blocks"] = so[min(i + batch, len(so)) - 1]["total_blocks"]
        assert s == want, (i, s, want)
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what=f"item list vps{vps} batch{batch}")
