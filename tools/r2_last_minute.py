"""Last GPU minute of round 2: torch-free check (numpy + ctypes only, so that it starts in seconds) of the two things added
after the final validation calls — the golden mesh fixture on the product, and the handle-based kb_frame_owners / kb_frame_cells
after their refactoring onto the shared host implementation (must equal the handle-free variants)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from khronos_b200 import capi  # noqa: E402

lib = capi.load_product_library()
G = os.path.join(ROOT, "tests", "golden")
g, m = np.load(os.path.join(G, "fusion.npz")), np.load(os.path.join(G, "mesh.npz"))
cam = capi.Camera(80, 60, 40.0, 40.0, 39.5, 29.5, 0.1, 2.5)
h = capi.MapHandle(lib, "kb_", capi.default_map_config(), capi.default_integrator_config(), capi.default_tracking_config(), capi.default_motion_config())
h.set_camera(cam)
for i in range(len(g["stamps"])):
    d, l = np.ascontiguousarray(g["depth"][i]), np.ascontiguousarray(g["label"][i])
    h.integrate_frame(h.make_frame(d, g["poses"][i], int(g["stamps"][i]), label=l))
    h.update_tracking(int(g["stamps"][i]))
assert tuple(int(x) for x in m["checksum"]) == h.map_checksum(), "map checksum differs from the golden fixture"
bi, off, pts, col, lab = h.generate_mesh(False, False)
assert np.array_equal(m["block_index"], bi) and np.array_equal(m["offsets"], off)
assert np.array_equal(m["points_bits"], pts.view(np.uint32)) and np.array_equal(m["labels"], lab.astype(np.uint8))
print("golden mesh ok:", len(pts), "vertices")
frames = [h.make_frame(None, g["poses"][i], int(g["stamps"][i])) for i in range(len(g["stamps"]))]
table = (np.arange(6 * 8, dtype=np.uint8).reshape(6, 8) % 4)
h.set_shard_table(0, 4, 2, (-2, -2), table)
a = h.frame_owners(frames)
b = capi.frame_owners_host(lib, "kb_", cam, 0.05, 16, frames, 4, cell_blocks=2, grid=(2, 2), origin=(-2, -2), table=table)
assert np.array_equal(a, b), (a, b)
c = h.frame_cells(frames, 2, (-2, -2), 8, 6)
d2 = capi.frame_cells_host(lib, "kb_", cam, 0.05, 16, frames, 2, (-2, -2), 8, 6)
assert np.array_equal(c, d2)
h.set_shard_cells(1, 4, 3, 2, 2)
assert np.array_equal(h.frame_owners(frames), capi.frame_owners_host(lib, "kb_", cam, 0.05, 16, frames, 4, cell_blocks=3, grid=(2, 2)))
print("frame owners / cells: handle-based == handle-free")
