"""Multi-GPU plumbing (one process per GPU, torch.distributed): clip sharding for extraction (no collective) and
the single gather of per-GPU top-k lists for database search.  No kernels here."""
import numpy as np


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items for `rank`; shards differ in size by at most one and cover everything."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_topk(scores: np.ndarray, indices: np.ndarray, device=None):
    """all_gather of one rank's [queries][k] top-k lists -> ([world][queries][k] scores, indices) on every rank.
    Uses the default process group (NCCL on GPUs: one collective over NVLink; gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return scores[None].copy(), indices[None].copy()
    world = dist.get_world_size()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    # one payload: scores and indices travel together as raw 32-bit words
    stacked = np.stack([np.ascontiguousarray(scores).view(np.int32), np.ascontiguousarray(indices).view(np.int32)])
    payload = torch.from_numpy(stacked.reshape(-1)).to(dev)
    out = torch.empty(world * payload.numel(), dtype=payload.dtype, device=dev)
    dist.all_gather_into_tensor(out, payload)
    out = out.cpu().numpy().reshape((world,) + stacked.shape)
    return np.ascontiguousarray(out[:, 0]).view(np.float32), np.ascontiguousarray(out[:, 1]).view(np.uint32)


def merge_reference(scores: np.ndarray, indices: np.ndarray):
    """numpy statement of the merge order (score desc, clip index asc) — used by tests to check the device merge."""
    n_lists, n_q, k = scores.shape
    s = scores.transpose(1, 0, 2).reshape(n_q, -1); i = indices.transpose(1, 0, 2).reshape(n_q, -1)
    valid_last = np.where(i == 0xFFFFFFFF, 1, 0)
    o = np.lexsort((i, -s.astype(np.float64), valid_last), axis=1)[:, :k]
    return np.take_along_axis(s, o, 1), np.take_along_axis(i, o, 1)
