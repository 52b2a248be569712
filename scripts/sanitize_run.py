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
# Haar/select register kernel incl. its checked-division and tie paths; recording-rate front end (D = 4 fast path and a generic rate)
rng = np.random.default_rng(3)
imgs = np.concatenate([(rng.random((2, 128, 32)) ** 4 * 50).astype(np.float32), (rng.random((1, 128, 32)) * 1e-36).astype(np.float32),
                       rng.integers(0, 3, size=(1, 128, 32)).astype(np.float32)])
th, tb = d.transform_images(imgs)
hi = p.synth_clip(9, 3 * 44100 + 1234, 44100.0)          # not a whole number of tiles or pairs
r1 = d.resample(hi); fr = d.process_recorded_pcm(hi)
d3 = lb.Detective(); d3.set_recording_rate(16000.0); r2 = d3.resample(p.synth_clip(9, 2 * 16000, 16000.0))
# compare-audio as one device pipeline (equal and unequal lengths); few-query search (lane = clip) with the two-level merge of its chunk lists
m1 = d.compare_pcm(pcm[0], pcm[1], 0); m2 = d.compare_pcm(pcm[0], pcm[1][:30000], 0)
big = lb.Database(200); big.add_packed(np.tile(w, (1500, 1, 1)))            # 4,500 clips -> 141 chunk lists
sc4, idx4 = big.search_packed(w[:1, :6], k=3); sc5, idx5 = big.search_packed(w[:, 1:4], k=10)
assert idx4[0, 0] == 0 and sc4[0, 0] == 1.0 and m1 > 0.0
# 100-rank form of the search (regular queries in every warp: tiles rewritten in place) against a regular and a mixed database
rs = np.random.default_rng(8); sign = rs.integers(0, 2, size=(700, 19, 100)); cbits = np.zeros((700, 19, 200), np.uint8); cbits[..., 0::2] = sign == 0; cbits[..., 1::2] = sign == 1
qpk = lb.pack_booleans(cbits[:40, 2:8])                                                # regular queries, packed before the database is disturbed
reg = lb.Database(200); reg.add_packed(lb.pack_booleans(cbits)); sc6, idx6 = reg.search_packed(qpk, k=3)
cbits[45, 3, 10:12] = 0; cbits[600, :, :] = 0
mix = lb.Database(200); mix.add_packed(lb.pack_booleans(cbits)); sc7, idx7 = mix.search_packed(qpk, k=3)
assert (idx6[:, 0] == np.arange(40)).all() and (idx7[:, 0] == np.arange(40)).all()
# the Frame API's kernels on a frame that is not a power of two in either direction
frm = lb.Frame.from_array(rs.standard_normal((37, 19)).astype(np.float32)); frm.decompose(); fb = frm.extract_fingerprint(50)
f0 = lb.Fingerprint(200); f0.add_packed(w[0]); f1 = lb.Fingerprint(200); f1.add_packed(w[1])
print("ok", int(fb.sum()), tb.shape, r1.shape, r2.shape, fr.count, w.shape, (bits != bits2).sum(), b3.shape, sc[:, 0], f0.compare(f1, 200), lb.merge_topk(np.stack([sc, sc]), np.stack([idx, idx + 10]))[1][0])
