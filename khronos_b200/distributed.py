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
