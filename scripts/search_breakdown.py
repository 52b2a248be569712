"""Where a sharded search step's time goes on ONE shard: the search kernel (the library's own timer) against the whole search_device
call (kernel + partial-list merge), for the shard sizes of 1, 2, 4 and 8 GPUs and for the config-5 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lbaudiodetective_b200 as lb

s = torch.cuda.Stream(); torch.cuda.set_stream(s); st = s.cuda_stream
def run(n_db, c_db, n_q, c_q, reps=10, silent_every=0):
    """silent_every = m: every m-th clip (beyond the query sources) loses the sign bits of one subfingerprint — an empty-rank code as
    digital silence produces — so the database as a whole is no longer 'regular' and only the tiles without such a clip keep the short form"""
    db = lb.Database(200)
    codes = torch.empty((n_db, c_db, 8), dtype=torch.int32, device="cuda"); lb.random_codes_device(codes.data_ptr(), n_db * c_db, 200, seed=5, stream=st)
    if silent_every:
        with torch.cuda.stream(s):
            codes[n_q + 1::silent_every, c_db // 2, :] = 0
    db.add_packed_device(codes.data_ptr(), n_db, c_db, producer_stream=st)
    q = codes[:n_q, :c_q].contiguous(); sc = torch.empty((n_q, 10), dtype=torch.float32, device="cuda"); ix = torch.empty((n_q, 10), dtype=torch.int32, device="cuda")
    for _ in range(3):
        db.search_device(q.data_ptr(), n_q, c_q, 10, sc.data_ptr(), ix.data_ptr(), stream=st)
    torch.cuda.synchronize(); db.kernel_timing(enable=True, reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        db.search_device(q.data_ptr(), n_q, c_q, 10, sc.data_ptr(), ix.data_ptr(), stream=st)
    e1.record(); torch.cuda.synchronize()
    n, k_ms = db.kernel_timing(enable=False, reset=True)
    total = e0.elapsed_time(e1) / reps; kern = k_ms / n
    cmp_ = n_q * n_db * (c_db - c_q + 1) * c_q
    print(("silent 1/%d " % silent_every if silent_every else "") + "db %8d x %2d, %4d queries x %d: search kernel %.3f ms (%.3e compares/s), whole call %.3f ms, merge + gaps %.3f ms" % (n_db, c_db, n_q, c_q, kern, cmp_ / kern * 1e3, total, total - kern), flush=True)
    assert (ix[:, 0].cpu() == torch.arange(n_q, dtype=torch.int32)).all()

if os.environ.get("BREAKDOWN_SIZES"):                      # BREAKDOWN_SIZES=65536,131072,...: the fixed cost of a search (intercept over the database size)
    for n in os.environ["BREAKDOWN_SIZES"].split(","):
        run(int(n), 19, 1000, 6)
elif os.environ.get("BREAKDOWN_MIXED"):
    run(250000, 19, 1000, 6); run(250000, 19, 1000, 6, silent_every=1000); run(250000, 19, 1000, 6, silent_every=100); run(250000, 19, 1000, 6, silent_every=1)
elif os.environ.get("BREAKDOWN_SHORT"):
    run(250000, 19, 1000, 6); run(125000, 19, 1000, 6)
else:
    for n in (1000000, 500000, 250000, 125000):
        run(n, 19, 1000, 6)
    run(100000, 5, 1000, 1)
    run(100000, 5, 1000, 2)
    run(1000000, 19, 32, 6)
    run(1000000, 19, 1, 6)
