"""kb_frame_owners_host / kb_frame_cells_host — the product's host-side scheduling arithmetic, which needs no GPU — against
the oracle's exact frustum selection on the benchmark trajectory (hall640 lap): the product's answers must contain the
oracle's (they are evaluated with a 1 mm larger inflation) and may exceed them only marginally. These are the functions the
8-GPU replay derives its shard layout, frame placement and pull plans from (bench.py --gpus N)."""
import numpy as np
import pytest

from khronos_b200 import capi, synthetic as syn
from khronos_b200.replay import bisect_layout, rank_grid


@pytest.fixture(scope="module")
def lap():
    cam = syn.make_camera()
    poses, stamps = syn.sweep_trajectory(5000)
    sel = list(range(0, 5000, 5))  # every 5th frame of the lap
    frames = [capi.MapHandle.make_frame(None, poses[i], stamps[i]) for i in sel]
    return cam, frames, [poses[i] for i in sel]


def _grid(cam, poses, cell):
    bsz = 0.05 * 16 * cell
    reach = cam.max_range + 2 * 0.05 * 16
    px = np.array([np.asarray(T, np.float64).reshape(4, 4)[0, 3] for T in poses])
    py = np.array([np.asarray(T, np.float64).reshape(4, 4)[1, 3] for T in poses])
    ox, oy = int(np.floor((px.min() - reach) / bsz)), int(np.floor((py.min() - reach) / bsz))
    return (ox, oy), int(np.floor((px.max() + reach) / bsz)) - ox + 1, int(np.floor((py.max() + reach) / bsz)) - oy + 1


def test_cells_contain_the_oracle_selection(oracle_lib, product_lib, lap):
    cam, frames, poses = lap
    origin, w, h = _grid(cam, poses, 4)
    tp = capi.frame_cells_host(product_lib, "kb_", cam, 0.05, 16, frames, 4, origin, w, h)
    to = capi.frame_cells_host(oracle_lib, "ko_", cam, 0.05, 16, frames, 4, origin, w, h)
    assert to.any(axis=(1, 2)).all()
    assert ((tp != 0) | (to == 0)).all(), "a cell the oracle selects is missing"
    extra = int(((tp != 0) & (to == 0)).sum())
    assert extra <= 0.002 * int((to != 0).sum()), extra  # the 1 mm margin adds next to nothing


@pytest.mark.parametrize("world", [2, 4, 8])
def test_owner_masks_for_all_layouts(oracle_lib, product_lib, lap, world):
    cam, frames, poses = lap
    gx, gy = rank_grid(world)
    origin, w, h = _grid(cam, poses, 4)
    table = bisect_layout(capi.frame_cells_host(product_lib, "kb_", cam, 0.05, 16, frames, 4, origin, w, h), world)
    layouts = [dict(cell_blocks=0), dict(cell_blocks=16, grid=(gx, gy)), dict(cell_blocks=24, grid=(gx, gy)),
               dict(cell_blocks=4, grid=(gx, gy), origin=origin, table=table)]
    for lay in layouts:
        mp = capi.frame_owners_host(product_lib, "kb_", cam, 0.05, 16, frames, world, **lay)
        mo = capi.frame_owners_host(oracle_lib, "ko_", cam, 0.05, 16, frames, world, **lay)
        assert ((mp & mo) == mo).all(), lay          # superset of the exact selection
        assert (mp > 0).all() and (mp < (1 << world)).all()
        differ = int((mp != mo).sum())
        assert differ <= 0.01 * len(frames), (lay.get("cell_blocks"), differ)
    # the table layout puts fewer frames on the busiest rank than the tilings (what bench.py selects by)
    busiest = lambda m: max(int(((m >> r) & 1).sum()) for r in range(world))
    m_tab = capi.frame_owners_host(product_lib, "kb_", cam, 0.05, 16, frames, world, **layouts[3])
    m_t16 = capi.frame_owners_host(product_lib, "kb_", cam, 0.05, 16, frames, world, **layouts[1])
    assert busiest(m_tab) < busiest(m_t16)
