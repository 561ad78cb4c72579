"""The C++ replay scheduler (khronos_b200/host/khronos_gpu_replay.h) is the same function as khronos_b200/replay.py:
layout table, chunk homes and step plans agree on random scenarios (CPU only; the scheduler is pure host arithmetic)."""
import os
import subprocess

import numpy as np
import pytest

from khronos_b200.replay import StripedSchedule, bisect_layout, rank_grid, route_homes
from harness import ROOT


@pytest.fixture(scope="module")
def replay_exe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("replay") / "replay_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "replay_check.cpp"), "-o", exe])
    return exe


@pytest.mark.parametrize("world,seed", [(2, 0), (4, 1), (8, 2), (3, 3), (6, 4)])
def test_cpp_scheduler_matches_python(replay_exe, world, seed):
    rng = np.random.default_rng(seed)
    F, H, W, stripe = 160, 6, 9, 8
    touched = np.zeros((F, H, W), np.uint8)
    x, y = 0.0, 0.0
    for f in range(F):  # a footprint wandering over the cell grid
        x = (x + rng.uniform(0.0, 0.3)) % (W - 2)
        y = (y + rng.uniform(-0.2, 0.25)) % (H - 2)
        touched[f, int(y):int(y) + 2, int(x):int(x) + 2] = 1
    table = bisect_layout(touched, world)
    masks = np.zeros(F, np.uint32)
    for f in range(F):
        for r in np.unique(table[touched[f] != 0]):
            masks[f] |= np.uint32(1 << int(r))
    step = [(11 + j) % F for j in range(F)]
    text = f"{F} {H} {W} {world} {stripe} {len(step)}\n" + " ".join(map(str, touched.ravel().tolist())) + "\n" + \
           " ".join(map(str, masks.tolist())) + "\n" + " ".join(map(str, step)) + "\n"
    out = subprocess.run([replay_exe], input=text, capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0] == "grid %d %d" % rank_grid(world)
    assert [int(v) for v in out[1].split()[1:]] == table.ravel().tolist()
    homes = route_homes(masks, world, stripe)
    assert [int(v) for v in out[2].split()[1:]] == homes.tolist()
    lines = out[3:]
    k = 0
    for routed in (0, 1):
        for r in range(world):
            s = StripedSchedule(world, r, stripe, homes=homes if routed else None)
            p = s.plan(step, masks)
            head, rest = lines[k].split(" mine")
            mine_txt, rest = rest.split(" ranges")
            ranges_txt, res_txt = rest.split(" resident")
            assert head == f"plan {routed} {r} {p.n_remote}"
            assert [tuple(int(v) for v in m.split(",")) for m in mine_txt.split()] == [tuple(m) for m in p.mine]
            assert [tuple(int(v) for v in m.split(",")) for m in ranges_txt.split()] == [tuple(g) for g in p.ranges]
            assert int(res_txt) == len(s.resident(F))
            k += 1
