"""Multi-GPU plumbing for the block-hash sharded map: one process per GPU,
torch.distributed for the frame broadcast. The data path has no other collective for K0/K1: every rank sees
the whole frame and integrates only the blocks it owns (kb_set_shard / kb_block_owner)."""
from __future__ import annotations

import numpy as np


def broadcast_frames(depth, label, src=0):
    """In-place broadcast of a batch of frames (depth f32 [F,H,W], label i32 [F,H,W]) from `src`."""
    import torch.distributed as dist
    dist.broadcast(depth, src)
    dist.broadcast(label, src)
    return depth, label


def owner_of_blocks(lib, block_index: np.ndarray, nranks: int) -> np.ndarray:
    """Shard owner of each block index (n,3) via the product library's kb_block_owner (pure function,
    needs no GPU)."""
    return np.array([lib.kb_block_owner(int(b[0]), int(b[1]), int(b[2]), int(nranks)) for b in block_index],
                    dtype=np.int32)


def gather_block_indices(local_index: np.ndarray, world: int):
    """All-gather of per-rank (n_i, 3) int32 block-index lists (ragged) -> list of arrays on every rank."""
    import torch
    import torch.distributed as dist
    n = torch.tensor([local_index.shape[0]], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, n)
    mx = int(max(int(s.item()) for s in sizes))
    buf = torch.zeros((max(mx, 1), 3), dtype=torch.int32)
    if local_index.shape[0]:
        buf[: local_index.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_index))
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return [o[: int(s.item())].numpy() for o, s in zip(out, sizes)]


# ---- sharded per-frame pipeline (SURVEY.md §8e exchange steps 1 and 2) ------------------------------------------

class DistComm:
    """Collectives of one rank of a torch.distributed job (NCCL on the GPU box, gloo in the CPU tests). Every method
    takes / returns a list with this rank's single tensor so that LocalComm can stand in for it."""

    def __init__(self, world: int, group=None):
        self.world, self.group = world, group

    def all_gather(self, bufs):
        import torch
        import torch.distributed as dist
        (buf,) = bufs
        out = torch.empty((self.world * buf.numel(),), dtype=buf.dtype, device=buf.device)
        try:
            dist.all_gather_into_tensor(out, buf, group=self.group)
        except (RuntimeError, NotImplementedError):  # backends without the flat variant
            parts = [torch.empty_like(buf) for _ in range(self.world)]
            dist.all_gather(parts, buf, group=self.group)
            out = torch.cat(parts)
        return [out]

    def all_reduce_max(self, bufs):
        import torch.distributed as dist
        dist.all_reduce(bufs[0], op=dist.ReduceOp.MAX, group=self.group)
        return bufs


class LocalComm:
    """All shards live in this process (tests: S handles on one device): collectives become tensor ops."""

    def all_gather(self, bufs):
        import torch
        cat = torch.cat([b.reshape(-1) for b in bufs]).contiguous()
        return [cat] * len(bufs)

    def all_reduce_max(self, bufs):
        import torch
        m = torch.stack(bufs).max(dim=0).values
        for b in bufs:
            b.copy_(m)
        return bufs


class ShardedActiveWindow:
    """The per-frame active-window loop (ActiveWindow::spinOnce, khronos/src/active_window/active_window.cpp:127,
    209-214: motion detection -> integration with the dynamic image as mask -> tracking update) over a block-hash
    sharded map. `handles` are this process' shard handles — one per rank in a torch.distributed job (comm =
    DistComm), or all of them for an in-process emulation (comm = LocalComm). Every handle must already be
    configured (camera, kb_set_shard). Results: the union of the shards equals the unsharded map, and every rank
    gets the same dynamic image (tests/test_sharded_pipeline.py, tests/test_multiproc_gloo.py).

    Per frame: 1 all-reduce (H*W flag bytes) + 2 all-gathers (pending block list, free masks); no host round trip
    before the final kb_motion_result."""

    def __init__(self, handles, comm, device="cpu"):
        import torch
        self.handles, self.comm = list(handles), comm
        self.device = torch.device(device)
        pb, hb, fb = self.handles[0].shard_buffer_sizes()
        mk = lambda nbytes, dt: [torch.zeros(nbytes // torch.empty((), dtype=dt).element_size(), dtype=dt, device=self.device)
                                 for _ in self.handles]
        self.pending, self.halo = mk(pb, torch.int32), mk(hb, torch.int32)
        self.flags = mk(fb, torch.uint8)
        self._keep = None
        if self.device.type == "cuda":  # order the library's kernels with the collectives on torch's current stream
            for h in self.handles:
                h.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def spin_once(self, frames, want_image=True):
        """frames: one kb_frame per local handle (the same images on every rank). Returns per handle
        (dynamic image or None, n_seeds, n_clusters)."""
        from . import capi
        hs = self.handles
        for h, f, fl in zip(hs, frames, self.flags):
            h.motion_lookup_local(f, fl)
        self.comm.all_reduce_max(self.flags)
        for h, f, fl in zip(hs, frames, self.flags):
            h.motion_cluster_global(fl)
            g = capi.Frame.from_buffer_copy(f)
            g.mask = capi.MASK_LAST_DETECTION
            h.integrate_frame(g, want_stats=False)
        return self.update_tracking([int(f.stamp_ns) for f in frames], want_image=want_image)

    def update_tracking(self, stamps, want_image=False, with_motion_result=True):
        """Sharded TrackingIntegrator::updateBlocks (tracking_integrator.cpp:71-104)."""
        hs = self.handles
        for h, st, pb in zip(hs, stamps, self.pending):
            h.tracking_begin(st, pb)
        all_pending = self.comm.all_gather(self.pending)
        for h, ap, hb in zip(hs, all_pending, self.halo):
            h.tracking_pack_halo(ap, hb)
        all_halo = self.comm.all_gather(self.halo)
        for h, ap, ah in zip(hs, all_pending, all_halo):
            h.tracking_finish(ap, ah)
        self._keep = (all_pending, all_halo)  # the ever-free kernel reads them asynchronously
        if not with_motion_result:
            return None
        return [h.motion_result(want_image) for h in hs]


# ---- peer-memory exchange (no collectives: producers store straight into every rank's buffers) -----------------------

class LocalPeers:
    """All shards live in this process: every shard's buffers are directly addressable, the barrier is the program order."""

    def __init__(self, n_shards: int, device="cpu"):
        import torch
        self.n, self.device = n_shards, torch.device(device)

    def alloc(self, numel: int, dtype):
        """One buffer per local shard + for each local shard the list of all ranks' buffers (as seen from that shard)."""
        import torch
        bufs = [torch.zeros(numel, dtype=dtype, device=self.device) for _ in range(self.n)]
        return bufs, [bufs for _ in range(self.n)]

    def barrier(self):
        pass


class SymmMemPeers:
    """One rank of a torch.distributed job: buffers come from torch.distributed._symmetric_memory (peer-mapped over
    NVLink), the barrier is the symmetric-memory barrier on the current stream. UNVERIFIED: written against the PyTorch
    2.11 API without a multi-GPU box at hand; tests cover the kernels through LocalPeers only."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem, self.group = symm_mem, group if group is not None else dist.group.WORLD
        self.device, self.handles = device, []

    def alloc(self, numel: int, dtype):
        t = self.symm_mem.empty(numel, dtype=dtype, device=self.device)
        t.zero_()
        hdl = self.symm_mem.rendezvous(t, self.group)
        self.handles.append(hdl)
        # buffer_ptrs are allocation bases; a tensor carved out of a pool sits `offset` bytes into it. Resolve which
        # convention this PyTorch build uses against the one pointer we know (our own) instead of assuming.
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        off = int(getattr(hdl, "offset", 0) or 0)
        own = ptrs[int(hdl.rank)]
        if own + off == t.data_ptr():
            ptrs = [p + off for p in ptrs]
        elif own != t.data_ptr():
            raise RuntimeError("symmetric memory: cannot relate buffer_ptrs to the tensor's data_ptr")
        return [t], [ptrs]

    def barrier(self):
        self.handles[0].barrier()


class PeerShardedActiveWindow:
    """ShardedActiveWindow with the peer-memory variants of the three exchanges: kb_motion_lookup_peers,
    kb_tracking_begin_peers and kb_tracking_pack_halo_peers store this rank's flag bytes / pending list / free masks
    straight into every rank's buffers; the host only places a barrier between producer and consumer."""

    def __init__(self, handles, peers, device="cpu"):
        import torch
        self.handles, self.peers = list(handles), peers
        self.device = torch.device(device)
        pb, hb, fb = self.handles[0].shard_buffer_sizes()
        world = None
        self.flags, self.flags_peers = peers.alloc(fb, torch.uint8)
        world = len(self.flags_peers[0])
        self.all_pending, self.pending_peers = peers.alloc(world * (pb // 4), torch.int32)
        self.all_halo, self.halo_peers = peers.alloc(world * (hb // 4), torch.int32)
        if self.device.type == "cuda":
            for h in self.handles:
                h.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def spin_once(self, frames, want_image=True):
        from . import capi
        hs = self.handles
        for h, f, fp in zip(hs, frames, self.flags_peers):
            h.motion_lookup_peers(f, fp)
        self.peers.barrier()
        for h, f, fl in zip(hs, frames, self.flags):
            h.motion_cluster_global(fl)
            fl.zero_()  # ready for the peers' stores of the next frame (ordered by the barriers of the tracking exchange)
            g = capi.Frame.from_buffer_copy(f)
            g.mask = capi.MASK_LAST_DETECTION
            h.integrate_frame(g, want_stats=False)
        return self.update_tracking([int(f.stamp_ns) for f in frames], want_image=want_image)

    def update_tracking(self, stamps, want_image=False, with_motion_result=True):
        hs = self.handles
        for h, st, pp in zip(hs, stamps, self.pending_peers):
            h.tracking_begin_peers(st, pp)
        self.peers.barrier()
        for h, ap, hp in zip(hs, self.all_pending, self.halo_peers):
            h.tracking_pack_halo_peers(ap, hp)
        self.peers.barrier()
        for h, ap, ah in zip(hs, self.all_pending, self.all_halo):
            h.tracking_finish(ap, ah)
        # tracking_finish(k) reads this rank's all_pending / all_halo (ghost table, overflow words); a faster rank's
        # tracking_begin_peers(k + 1) stores into them. Close the pass with a barrier so that back-to-back passes are safe
        # without relying on spin_once's motion barrier for the ordering.
        self.peers.barrier()
        if not with_motion_result:
            return None
        return [h.motion_result(want_image) for h in hs]
