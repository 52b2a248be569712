import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import lbaudiodetective_b200 as lb
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for n in (100000, 1000000):
    db = lb.Database(200)
    codes = torch.empty((n, 19, 8), dtype=torch.int32, device="cuda"); lb.random_codes_device(codes.data_ptr(), n * 19, 200, seed=5, stream=s.cuda_stream); torch.cuda.synchronize()
    db.add_packed_device(codes.data_ptr(), n, 19)
    for nq in (1, 4, 32, 128, 1000):
        q = codes[:nq, 3:9].contiguous(); sc = torch.empty((nq, 10), dtype=torch.float32, device="cuda"); ix = torch.empty((nq, 10), dtype=torch.int32, device="cuda")
        for _ in range(3): db.search_device(q.data_ptr(), nq, 6, 10, sc.data_ptr(), ix.data_ptr(), stream=s.cuda_stream)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): db.search_device(q.data_ptr(), nq, 6, 10, sc.data_ptr(), ix.data_ptr(), stream=s.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        assert (ix[:, 0].cpu().numpy() == np.arange(nq)).all()
        print("db %8d clips, %4d queries: %8.3f ms per search" % (n, nq, e0.elapsed_time(e1) / 10))
    del db, codes
