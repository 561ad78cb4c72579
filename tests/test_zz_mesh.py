"""kb_generate_mesh (marching cubes on the device) against the oracle's mesher: same blocks, same
vertex order, bit-identical vertex positions, colours and labels; flag semantics of (only_mesh_updated, clear_updated_flag)
as used at active_window.cpp:223 (true, true) and mesh_object_extractor.cpp:267 (true, false)."""
import numpy as np
import pytest

from khronos_b200 import capi, synthetic as syn
import harness as hs
from test_parity_gpu import room_frames

pytestmark = pytest.mark.gpu


def assert_mesh_equal(mo, mg, what=""):
    bo, oo, po, co, lo = mo
    bg, og, pg, cg, lg = mg
    np.testing.assert_array_equal(bo, bg, err_msg=f"{what} mesh blocks")
    np.testing.assert_array_equal(oo, og, err_msg=f"{what} vertex offsets")
    np.testing.assert_array_equal(po.view(np.uint32), pg.view(np.uint32), err_msg=f"{what} vertex positions (bits)")
    np.testing.assert_array_equal(co, cg, err_msg=f"{what} colours")
    np.testing.assert_array_equal(lo, lg, err_msg=f"{what} labels")


def test_mesh_room_stream_with_colour_bit_exact(oracle_lib, product_lib):
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 8)
    colors = [syn.colorize(l, d) for d, l in frames]
    o = hs.make_handle(oracle_lib, "ko_", cam=cam)
    g = hs.make_handle(product_lib, "kb_", cam=cam)
    hs.run_fusion(o, frames[:5], poses[:5], stamps[:5], colors=colors[:5])
    hs.run_fusion(g, frames[:5], poses[:5], stamps[:5], colors=colors[:5])
    mo, mg = o.generate_mesh(True, True), g.generate_mesh(True, True)
    assert len(mo[2]) > 20000 and mo[3].any() and mo[4].any()
    assert_mesh_equal(mo, mg, "first tick")
    # second tick: nothing updated -> empty; then more frames -> only the blocks they touch
    assert len(g.generate_mesh(True, True)[0]) == 0
    hs.run_fusion(o, frames[5:], poses[5:], stamps[5:], colors=colors[5:])
    hs.run_fusion(g, frames[5:], poses[5:], stamps[5:], colors=colors[5:])
    mo, mg = o.generate_mesh(True, False), g.generate_mesh(True, False)
    assert_mesh_equal(mo, mg, "second tick")
    assert_mesh_equal(o.generate_mesh(True, True), g.generate_mesh(True, True), "flags kept by clear_updated_flag = false")
    assert_mesh_equal(o.generate_mesh(False, False), g.generate_mesh(False, False), "all blocks")
    hs.assert_blocks_equal(o.export_blocks(), g.export_blocks(), exact_float=True, what="map after meshing")


def test_mesh_hall_batched_full_resolution(oracle_lib, product_lib):
    """The output tick of the benchmarked configuration: 640x480 hall frames fused in one 32-frame call, then meshed."""
    import torch
    cam = syn.make_camera()
    scene = syn.hall_scene(20)
    poses, stamps = syn.sweep_trajectory(5000)
    poses, stamps = poses[800:832], stamps[800:832]
    d, l = syn.render_stream(scene, cam, poses, stamps, device="cuda", dtype=torch.float32)
    d, l = d.cpu().numpy(), l.cpu().numpy()
    mc = capi.default_map_config(max_blocks=8192)
    ic = capi.default_integrator_config(num_threads=-1)
    o = capi.MapHandle(oracle_lib, "ko_", mc, ic, capi.default_tracking_config(), None)
    g = capi.MapHandle(product_lib, "kb_", mc, ic, capi.default_tracking_config(), None)
    o.set_camera(cam)
    g.set_camera(cam)
    o.integrate_frames([o.make_frame(d[i], poses[i], stamps[i], label=l[i]) for i in range(32)])
    g.integrate_frames([g.make_frame(d[i], poses[i], stamps[i], label=l[i]) for i in range(32)])
    mo, mg = o.generate_mesh(True, True), g.generate_mesh(True, True)
    assert len(mo[0]) > 300 and len(mo[2]) > 50000
    assert_mesh_equal(mo, mg, "hall640 tick")


def test_mesh_object_extraction_map(oracle_lib, product_lib):
    """MeshObjectExtractor's private map (vps 8, binary semantics, pre-allocated box, confidence scan) then
    generateMesh(map, true, false) as at mesh_object_extractor.cpp:267."""
    cam = hs.small_camera(4)
    frames, poses, stamps = room_frames(cam, 6)
    mc = capi.default_map_config(voxel_size=0.1, vps=8, trunc=0.2, with_tracking=False, max_blocks=4096)
    ic = capi.default_integrator_config(semantic_mode=capi.SEM_BINARY)
    o = hs.make_handle(oracle_lib, "ko_", cam=cam, map_cfg=mc, integ_cfg=ic)
    g = hs.make_handle(product_lib, "kb_", cam=cam, map_cfg=mc, integ_cfg=ic)
    for h in (o, g):
        h.allocate_box((0, 0, 0), (14, 12, 3))
        for (d, l), T, st in zip(frames, poses, stamps):
            obj = (l % 5).astype(np.int32)
            h.integrate_frame(h.make_frame(d, T, st, object_image=obj, target_id=2), allocate_blocks=False)
        h.scan_object_confidence(0.5, 2)
    mo, mg = o.generate_mesh(True, False), g.generate_mesh(True, False)
    assert len(mo[2]) > 1000
    assert_mesh_equal(mo, mg, "object map")
    assert set(np.unique(mo[4])) <= {0, 1}
