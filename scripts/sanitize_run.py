"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lbaudiodetective_b200 as lb
from oracle.oracle import Port, Cfg

p = Port()
pcm = np.stack([p.synth_clip(i, 55120) for i in range(3)])
d = lb.Detective()
w = d.process_batch(pcm)                                   # fast path (TMA staging)
w2 = d.process_batch(np.ascontiguousarray(pcm[:, :55001]))  # plain-load staging
img, haar, bits = d.process_stages(pcm[0], fused=True)
img2, haar2, bits2 = d.process_stages(pcm[0], fused=False) # generic kernels
d2 = lb.Detective(); d2.set_window_size(512); d2.set_subfingerprint_length(100)
b3 = d2.process_pcm(pcm[1]).booleans()
db = lb.Database(200); db.add_packed(w)
sc, idx, full = db.search_packed(w[:, 1:5], k=2, all_scores=True)      # generic (cq = 4 -> fast), masked below
sc2, idx2 = db.search_packed(w[:, :6], k=3, rng=77)
sc3, idx3 = db.search_packed(np.concatenate([w[:1], w[:1]], axis=1)[:, :9], k=2)   # query longer than the clips -> generic kernel
f0 = lb.Fingerprint(200); f0.add_packed(w[0]); f1 = lb.Fingerprint(200); f1.add_packed(w[1])
print("ok", w.shape, (bits != bits2).sum(), b3.shape, sc[:, 0], f0.compare(f1, 200), lb.merge_topk(np.stack([sc, sc]), np.stack([idx, idx + 10]))[1][0])
