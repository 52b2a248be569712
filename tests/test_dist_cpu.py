"""world_size-2 gloo tests of the multi-GPU host logic: clip sharding and the top-k gather (no GPU needed)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_everything():
    from lbaudiodetective_b200.dist import shard_range
    for n in (0, 1, 7, 8, 9, 1000, 120000, 1000003):
        for world in (1, 2, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def merge_reference(scores, indices):
    """numpy statement of the merge order (score desc, clip index asc, empty slots last) — the checker of this CPU test only; the
    product merges on the device (LBAudioDetectiveDatabaseMergeTopK[Device])."""
    n_lists, n_q, k = scores.shape
    s = scores.transpose(1, 0, 2).reshape(n_q, -1); i = indices.transpose(1, 0, 2).reshape(n_q, -1)
    valid_last = np.where(i == 0xFFFFFFFF, 1, 0)
    o = np.lexsort((i, -s.astype(np.float64), valid_last), axis=1)[:, :k]
    return np.take_along_axis(s, o, 1), np.take_along_axis(i, o, 1)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from lbaudiodetective_b200.dist import shard_range, gather_topk
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.default_rng(5)                     # same "database scores" on every rank
    n_clips, n_q, k = 1001, 13, 10
    full = np.round(rng.random((n_q, n_clips)), 2).astype(np.float32)
    lo, hi = shard_range(n_clips, rank, world)
    local = full[:, lo:hi]
    order = np.lexsort((np.arange(lo, hi)[None, :].repeat(n_q, 0), -local.astype(np.float64)), axis=1)[:, :k]
    sc = np.take_along_axis(local, order, 1); idx = (order + lo).astype(np.uint32)
    gs, gi = gather_topk(sc, idx)
    assert gs.shape == (world, n_q, k) and np.array_equal(gs[rank], sc) and np.array_equal(gi[rank], idx)
    ms, mi = merge_reference(gs, gi)
    want = np.lexsort((np.arange(n_clips)[None, :].repeat(n_q, 0), -full.astype(np.float64)), axis=1)[:, :k]
    assert np.array_equal(mi, want.astype(np.uint32)) and np.array_equal(ms, np.take_along_axis(full, want, 1))
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


def test_sharded_topk_gather_equals_single_shard(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
