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


def gather_and_merge_topk_device(d_scores, d_idx, stream=None):
    """Device-resident form: d_scores (float32) / d_idx (int32) are CUDA tensors [queries][k] holding this rank's top-k with global
    clip ids.  ONE NCCL all_gather (scores and indices travel as one int32 payload) and the device merge kernel; returns CUDA tensors
    [queries][k] identical on every rank.  No host round trip."""
    import torch
    import torch.distributed as dist
    from . import api
    n_q, k = d_scores.shape
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return d_scores, d_idx
    payload = torch.stack([d_scores.view(torch.int32), d_idx.view(torch.int32)]).reshape(-1)
    out = torch.empty(world * payload.numel(), dtype=torch.int32, device=payload.device)
    dist.all_gather_into_tensor(out, payload)
    out = out.view(world, 2, n_q, k)
    g_sc = out[:, 0].contiguous().view(torch.float32); g_id = out[:, 1].contiguous()
    m_sc = torch.empty((n_q, k), dtype=torch.float32, device=payload.device); m_id = torch.empty((n_q, k), dtype=torch.int32, device=payload.device)
    api.merge_topk_device(g_sc.data_ptr(), g_id.data_ptr(), world, n_q, k, m_sc.data_ptr(), m_id.data_ptr(),
                          stream if stream is not None else torch.cuda.current_stream().cuda_stream)
    return m_sc, m_id

