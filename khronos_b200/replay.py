"""Spatially sharded replay of a frame stream over the GPUs of one box (no counterpart in the reference).

Layout. The map is sharded by CELLS (kb_set_shard_cells): square groups of cell x cell blocks, tiled periodically over
a gx x gy grid of ranks, so a frame's frustum touches 1-4 ranks and each rank needs only the frames that touch its cells
(kb_frame_owners). The stream is resident STRIPED over the ranks' frame pools (stripe = 32 consecutive frames per rank,
round robin: what a host feeding each GPU over its own PCIe link produces). For every step each rank integrates, in
stream order, the sub-sequence of frames it needs: frames of its own stripe are read in place, the others are PULLED out
of the peers' pools over NVLink (kb_gather_*: CUDA IPC mappings + copy engines / a cp.async.bulk kernel) into a double
buffered receive area while the previous step is being fused. Pools are read-only, so there is no collective and no
cross-rank synchronisation inside a step; ranks run ahead of each other freely. Per voxel the sequence of updates is the
stream order restricted to the frames that see its block, exactly as on one GPU, so the union of the shards is
bit-identical to the unsharded map (kb_map_checksum sums add up).

`StripedSchedule` is pure Python/numpy (tested on CPU); `PeerPools` wraps the kb_peer_* C ABI."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np


def rank_grid(world: int) -> Tuple[int, int]:
    """gx x gy = world with gx >= gy as square as possible (8 -> 4 x 2, 4 -> 2 x 2, 2 -> 2 x 1, 6 -> 3 x 2)."""
    gy = int(np.floor(np.sqrt(world)))
    while world % gy:
        gy -= 1
    return world // gy, gy


@dataclass
class StepPlan:
    """One rank's share of a step: `mine` = (position in the step, global frame, slot) in stream order, slot >= 0 indexes
    the receive buffer, slot < 0 encodes a frame of the rank's own pool as -(local index) - 1; `ranges` = contiguous
    pulls (src rank, first local index on src, first receive slot, frame count)."""
    mine: List[Tuple[int, int, int]]
    ranges: List[Tuple[int, int, int, int]]
    n_remote: int


def route_homes(owner_mask: np.ndarray, world: int, stripe: int = 32) -> np.ndarray:
    """Pose-aware placement of the stream: the ingest host knows the poses, so it can hand every chunk of `stripe`
    consecutive frames to a rank that will integrate it (the rank most of the chunk's frames touch; ties go to the rank
    holding the fewest frames so far) instead of dealing chunks round robin. One delivery per frame is then local and only
    the frames' OTHER owners pull it over NVLink. Deterministic: every rank computes the same table from the same masks."""
    n = len(owner_mask)
    homes = np.zeros(n, np.int32)
    held = np.zeros(world, np.int64)
    for c0 in range(0, n, stripe):
        chunk = owner_mask[c0:c0 + stripe]
        votes = np.array([int(((chunk >> r) & 1).sum()) for r in range(world)])
        best = votes.max()
        cand = [r for r in range(world) if votes[r] == best]
        r = min(cand, key=lambda q: (held[q], q))
        homes[c0:c0 + stripe] = r
        held[r] += len(chunk)
    return homes


def bisect_layout(touched: np.ndarray, world: int) -> np.ndarray:
    """Trajectory-aware cell -> rank table (kb_set_shard_table). touched[f, cy, cx] != 0 iff frame f's frustum touches cell
    (cx, cy) (kb_frame_cells). Recursive bisection of the cell rectangle: a rectangle that gets k ranks is cut, along x or
    y, where the busier side's frames-per-rank is smallest — the load of a side is the number of frames that touch ANY of
    its cells (a frame on the cut counts on both sides, so cuts avoid busy places) — until every rectangle has one rank.
    Regions are contiguous, so a frustum touches few of them, and equally busy, which is what bounds a sharded replay
    (the per-batch cost of a rank hardly depends on how many of its blocks a frame holds)."""
    t = np.asarray(touched) != 0
    F, H, W = t.shape
    table = np.zeros((H, W), np.uint8)

    def load(x0, x1, y0, y1):
        return int(t[:, y0:y1, x0:x1].any(axis=(1, 2)).sum())

    def split(x0, x1, y0, y1, base, k):
        if k == 1 or (x1 - x0 <= 1 and y1 - y0 <= 1):
            table[y0:y1, x0:x1] = base  # (more ranks than cells: the surplus ranks stay idle)
            return
        k1 = k // 2
        k2 = k - k1
        best = None
        for axis, lo, hi in (("x", x0, x1), ("y", y0, y1)):
            for c in range(lo + 1, hi):
                a = load(x0, c, y0, y1) if axis == "x" else load(x0, x1, y0, c)
                b = load(c, x1, y0, y1) if axis == "x" else load(x0, x1, c, y1)
                cost = max(a / k1, b / k2)
                if best is None or cost < best[0]:
                    best = (cost, axis, c)
        _, axis, c = best
        if axis == "x":
            split(x0, c, y0, y1, base, k1)
            split(c, x1, y0, y1, base + k1, k2)
        else:
            split(x0, x1, y0, c, base, k1)
            split(x0, x1, c, y1, base + k1, k2)

    split(0, W, 0, H, 0, int(world))
    return table


class StripedSchedule:
    """Where the frames of a lap live (`homes[g]`, default: chunks of `stripe` frames dealt round robin) and what each rank
    pulls per step."""

    def __init__(self, world: int, rank: int, stripe: int = 32, homes: np.ndarray = None):
        self.world, self.rank, self.stripe = int(world), int(rank), int(stripe)
        self.homes = None if homes is None else np.asarray(homes, np.int32)
        self._local = None
        if self.homes is not None:  # local index = position among the frames of the same home, ascending frame number
            self._local = np.zeros(len(self.homes), np.int64)
            cnt = np.zeros(self.world, np.int64)
            for g, r in enumerate(self.homes):
                self._local[g] = cnt[r]
                cnt[r] += 1

    def home(self, g: int) -> int:
        if self.homes is not None:
            return int(self.homes[g])
        return (g // self.stripe) % self.world

    def local_index(self, g: int) -> int:
        if self._local is not None:
            return int(self._local[g])
        return (g // (self.stripe * self.world)) * self.stripe + g % self.stripe

    def resident(self, lap: int) -> List[int]:
        """Global frame numbers of a lap that live in this rank's pool, in local-index order."""
        return [g for g in range(lap) if self.home(g) == self.rank]

    def plan(self, step_frames: Sequence[int], owner_mask: np.ndarray) -> StepPlan:
        """step_frames[j] = global frame of position j of the step; owner_mask[g] = kb_frame_owners bit mask."""
        mine, ranges, slot = [], [], 0
        for j, g in enumerate(step_frames):
            if not (int(owner_mask[g]) >> self.rank) & 1:
                continue
            src = self.home(g)
            li = self.local_index(g)
            if src == self.rank:
                mine.append((j, g, -li - 1))
                continue
            if ranges and ranges[-1][0] == src and ranges[-1][1] + ranges[-1][3] == li and ranges[-1][2] + ranges[-1][3] == slot:
                r = ranges[-1]
                ranges[-1] = (r[0], r[1], r[2], r[3] + 1)
            else:
                ranges.append((src, li, slot, 1))
            mine.append((j, g, slot))
            slot += 1
        return StepPlan(mine, ranges, slot)


class CudaArray:
    """Raw device memory as a `__cuda_array_interface__` object: torch.as_tensor(CudaArray(...), device=...) is a
    zero-copy view (the frame pools are cudaMalloc allocations of the library so that they can be shared by CUDA IPC)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerPools:
    """This rank's frame pool (depth f32 [n, H, W] followed by label i32 [n, H, W] in one allocation) and read-only
    mappings of the other ranks' pools. `exchange` is a callable that all-gathers a 64-byte handle + pool size over the
    ranks (torch.distributed in bench.py; a local list in single-process tests)."""

    def __init__(self, lib, device: int, n_local: int, height: int, width: int):
        self.lib, self.device = lib, int(device)
        self.n, self.H, self.W = int(n_local), int(height), int(width)
        self.P = self.H * self.W
        self.bytes = max(self.n, 1) * self.P * 8
        p = C.c_void_p()
        self._check(lib.kb_peer_alloc(self.device, C.c_size_t(self.bytes), C.byref(p)), "kb_peer_alloc")
        self.ptr = int(p.value)
        self.peer_ptr = {}
        self.peer_n = {}
        self._opened = []

    def _check(self, st, what):
        if st != 0:
            self.lib.kb_peer_last_error.restype = C.c_char_p
            raise RuntimeError(f"{what} failed ({st}): {self.lib.kb_peer_last_error().decode()}")

    def depth_ptr(self, base: int, n: int, i: int) -> int:
        return base + i * self.P * 4

    def label_ptr(self, base: int, n: int, i: int) -> int:
        return base + max(n, 1) * self.P * 4 + i * self.P * 4

    def views(self, torch, dev):
        """(depth f32 [n, H, W], label i32 [n, H, W]) torch views of the local pool."""
        n = max(self.n, 1)
        d = torch.as_tensor(CudaArray(self.ptr, (n, self.H, self.W), "<f4"), device=dev)
        l = torch.as_tensor(CudaArray(self.ptr + n * self.P * 4, (n, self.H, self.W), "<i4"), device=dev)
        return d, l

    def export_handle(self) -> bytes:
        h = (C.c_uint8 * 64)()
        self._check(self.lib.kb_peer_export(self.device, C.c_void_p(self.ptr), h), "kb_peer_export")
        return bytes(h)

    def open_peer(self, rank: int, handle: bytes, n_frames: int):
        p = C.c_void_p()
        hb = (C.c_uint8 * 64).from_buffer_copy(handle)
        self._check(self.lib.kb_peer_open(self.device, hb, C.byref(p)), f"kb_peer_open(rank {rank})")
        self.peer_ptr[rank], self.peer_n[rank] = int(p.value), int(n_frames)
        self._opened.append(int(p.value))

    def add_local_peer(self, rank: int, ptr: int, n_frames: int):
        """Same-process peer (tests: several 'ranks' on one or more devices of one process)."""
        self.peer_ptr[rank], self.peer_n[rank] = int(ptr), int(n_frames)

    def gather_plan(self, ranges, rx_ptr: int, rx_capacity: int):
        """kb_gather_plan for a StepPlan's ranges: per range one copy of the depth images and one of the label images
        into the receive buffer (same layout as a pool: depth [cap, H, W] then label [cap, H, W])."""
        src, dst, nbytes = [], [], []
        for (r, li, slot, cnt) in ranges:
            base, n = self.peer_ptr[r], self.peer_n[r]
            assert li + cnt <= max(n, 1) and slot + cnt <= rx_capacity
            src += [self.depth_ptr(base, n, li), self.label_ptr(base, n, li)]
            dst += [self.depth_ptr(rx_ptr, rx_capacity, slot), self.label_ptr(rx_ptr, rx_capacity, slot)]
            nbytes += [cnt * self.P * 4, cnt * self.P * 4]
        k = len(src)
        plan = C.c_void_p()
        a_src = (C.c_void_p * max(k, 1))(*src)
        a_dst = (C.c_void_p * max(k, 1))(*dst)
        a_b = (C.c_uint64 * max(k, 1))(*nbytes)
        self._check(self.lib.kb_gather_plan_create(self.device, k, a_src, a_dst, a_b, C.byref(plan)), "kb_gather_plan_create")
        return plan

    def run(self, plan, mode: int, max_ctas: int, stream: int):
        self._check(self.lib.kb_gather_run(plan, int(mode), int(max_ctas), C.c_void_p(stream)), "kb_gather_run")

    def plan_bytes(self, plan) -> int:
        self.lib.kb_gather_plan_bytes.restype = C.c_uint64
        return int(self.lib.kb_gather_plan_bytes(plan))

    def close(self):
        for p in self._opened:
            self.lib.kb_peer_close(self.device, C.c_void_p(p))
        self._opened = []
        if self.ptr:
            self.lib.kb_peer_free(self.device, C.c_void_p(self.ptr))
            self.ptr = 0
