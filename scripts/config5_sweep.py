"""BASELINE.json configs[4]: bird-song-style 3 s noisy queries against a 100k-clip database, sweeping the window size and the
number of top-t wavelets (subfingerprint length L = 2 t Booleans).

For every (window, L): the database is 100,000 synthetic 9 s clips (5 subfingerprints each) extracted on the GPU; a query is a 3 s
excerpt (1 subfingerprint) of a database clip that starts on a frame boundary, plus uniform noise of 1.58 % / 3.16 % of full scale
(the essay's levels, p.34-35); the search is LBAudioDetectiveDatabaseSearch (time-offset search, top-10).  Reported per case: recall@1,
recall@10, the mean score of the true clip and of the best wrong clip, and the device times of extraction and search.  One JSON line
per case on stdout; `--out` also writes the list to a file (profiles/).  Window 4096 is not in the sweep: the reference reads out of
bounds there."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lbaudiodetective_b200 as lb

SR = 5512.0
ap = argparse.ArgumentParser()
ap.add_argument("--db-clips", type=int, default=100000)
ap.add_argument("--queries", type=int, default=1000)
ap.add_argument("--windows", default="512,1024,2048")
ap.add_argument("--sublens", default="100,200,400")
ap.add_argument("--noise", default="0.0158,0.0316")
ap.add_argument("--out", default=None)
a = ap.parse_args()

DB_LEN = 49608          # 9 s
Q_LEN = 16536           # 3 s
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
n = a.db_clips
pcm = torch.empty((n, DB_LEN), dtype=torch.float32, device="cuda")
lb.synthesize_device(pcm.data_ptr(), n, DB_LEN, DB_LEN, first_clip_id=0, stream=s.cuda_stream)
g = torch.Generator(device="cpu"); g.manual_seed(11)
q_src = torch.randint(0, n, (a.queries,), generator=g)
q_off = torch.randint(0, 4, (a.queries,), generator=g)                    # frame of the database clip the excerpt starts at
gn = torch.Generator(device="cuda"); gn.manual_seed(12)
results = []
for window in [int(x) for x in a.windows.split(",")]:
    for L in [int(x) for x in a.sublens.split(",")]:
        d = lb.Detective(); d.set_window_size(window); d.set_subfingerprint_length(L)
        assert d.check_configuration() == 0
        c_db = d.subfingerprints_for_length(DB_LEN); c_q = d.subfingerprints_for_length(Q_LEN)
        W2 = 2 * lb.words_per_plane(L)
        words = torch.zeros((n, c_db, W2), dtype=torch.int32, device="cuda")
        e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        d.process_batch_device(pcm.data_ptr(), n, DB_LEN, DB_LEN, words.data_ptr(), s.cuda_stream)      # warm-up (plan, scratch)
        e0.record()
        d.process_batch_device(pcm.data_ptr(), n, DB_LEN, DB_LEN, words.data_ptr(), s.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        db = lb.Database(L); db.add_packed_device(words.data_ptr(), n, c_db)
        for noise in [float(x) for x in a.noise.split(",")]:
            # excerpt starting at frame q_off of the source clip: the first window of the query is the first window of that frame
            idx = (q_off * 8192).unsqueeze(1) + torch.arange(Q_LEN).unsqueeze(0)
            q = pcm[q_src.cuda().unsqueeze(1), idx.cuda()].contiguous()
            q = (q + (torch.rand(q.shape, device="cuda", generator=gn) * 2.0 - 1.0) * noise).clamp_(-1.0, 1.0).contiguous()
            qw = torch.zeros((a.queries, c_q, W2), dtype=torch.int32, device="cuda")
            d.process_batch_device(q.data_ptr(), a.queries, Q_LEN, Q_LEN, qw.data_ptr(), s.cuda_stream)
            sc = torch.empty((a.queries, 10), dtype=torch.float32, device="cuda"); ix = torch.empty((a.queries, 10), dtype=torch.int32, device="cuda")
            db.search_device(qw.data_ptr(), a.queries, c_q, 10, sc.data_ptr(), ix.data_ptr(), stream=s.cuda_stream)   # warm-up
            e2.record()
            db.search_device(qw.data_ptr(), a.queries, c_q, 10, sc.data_ptr(), ix.data_ptr(), stream=s.cuda_stream)
            e3.record(); torch.cuda.synchronize()
            ixc = ix.cpu().numpy().view(np.uint32); scc = sc.cpu().numpy(); src = q_src.numpy().astype(np.uint32)
            hit1 = ixc[:, 0] == src; hit10 = (ixc == src[:, None]).any(axis=1)
            true_score = np.where(hit10, (scc * (ixc == src[:, None])).max(axis=1), np.nan)
            wrong_best = np.where(hit1, scc[:, 1], scc[:, 0])
            compares = a.queries * n * (c_db - c_q + 1) * c_q
            r = {"window": window, "subfingerprint_length": L, "top_t": L // 2, "noise": noise, "db_clips": n, "db_subfingerprints_per_clip": int(c_db),
                 "queries": a.queries, "query_subfingerprints": int(c_q), "recall_at_1": float(hit1.mean()), "recall_at_10": float(hit10.mean()),
                 "mean_true_score": float(np.nanmean(true_score)), "mean_best_wrong_score": float(wrong_best.mean()),
                 "extract_ms": e0.elapsed_time(e1), "extract_audio_hours_per_s": n * DB_LEN / SR / 3600.0 / (e0.elapsed_time(e1) * 1e-3),
                 "search_ms": e2.elapsed_time(e3), "search_compares_per_s": compares / (e2.elapsed_time(e3) * 1e-3)}
            results.append(r); print(json.dumps(r), flush=True)
        del db, words
        d.dispose()
if a.out:
    json.dump(results, open(a.out, "w"), indent=1)
