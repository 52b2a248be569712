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


class ShardedTopK:
    """Device-resident top-k exchange of a database sharded over the ranks (one process per GPU): the search writes this rank's
    [queries][k] scores and clip indices into ONE payload buffer, ONE NCCL all_gather moves the payloads of all ranks, and the
    library's merge kernel reads the gather buffer in place (LBAudioDetectiveDatabaseMergeTopKDeviceStrided) — no stacking or
    repacking copies in between, nothing visits the host.  Buffers are allocated once and reused by every search."""

    def __init__(self, n_q: int, k: int, device=None):
        import torch
        import torch.distributed as dist
        self.n_q, self.k = n_q, k
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.payload = torch.empty((2, n_q, k), dtype=torch.int32, device=dev)              # [0] = scores (float32 bits), [1] = clip indices
        self.gathered = torch.empty((self.world, 2, n_q, k), dtype=torch.int32, device=dev) if self.world > 1 else None
        self.merged = torch.empty((2, n_q, k), dtype=torch.int32, device=dev) if self.world > 1 else self.payload

    @property
    def scores_ptr(self):
        return self.payload[0].data_ptr()

    @property
    def indices_ptr(self):
        return self.payload[1].data_ptr()

    def gather_and_merge(self, stream=None):
        """After the search has been enqueued (on torch's current stream): returns (scores float32 [q][k], indices int32 [q][k]) CUDA
        tensors, identical on every rank and equal to a single-GPU search over the union of the shards."""
        import torch
        import torch.distributed as dist
        from . import api
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.payload)
            n = self.n_q * self.k
            api.merge_topk_device_strided(self.gathered.data_ptr(), self.gathered.data_ptr() + 4 * n, self.world, 2 * n, self.n_q, self.k,
                                          self.merged[0].data_ptr(), self.merged[1].data_ptr(),
                                          stream if stream is not None else torch.cuda.current_stream().cuda_stream)
        return self.merged[0].view(torch.float32), self.merged[1]


def gather_and_merge_topk_device(d_scores, d_idx, stream=None):
    """Convenience form over separate tensors (one extra copy into the payload); see ShardedTopK for the copy-free path."""
    n_q, k = d_scores.shape
    ex = ShardedTopK(n_q, k, d_scores.device)
    import torch
    ex.payload[0].copy_(d_scores.view(torch.int32)); ex.payload[1].copy_(d_idx.view(torch.int32))
    return ex.gather_and_merge(stream)
