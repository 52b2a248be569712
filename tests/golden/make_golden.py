"""Generates tests/golden/*.npz|json by running the COMPILED REFERENCE (oracle/_ref, i.e. the reference's own .m files
behind the disclosed shim) on seeded inputs.  Run in the container that has /root/reference:

    python tests/golden/make_golden.py

The fixtures are committed; the GPU box and CI only read them.
"""
import json
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.oracle import Ref, Cfg  # noqa: E402


def main():
    r = Ref()
    cfg = Cfg.default()
    # ---- config 1 of BASELINE.json: two 10 s clips through the compare-audio path at reference defaults ----
    pcm = np.stack([r.synth_clip(0, 55120), r.synth_clip(1, 55120)])
    bits, images, haar = [], [], []
    for c in range(2):
        b, i, h = r.process(cfg, pcm[c], stages=True)
        assert (b == r.process(cfg, pcm[c], direct=False)).all()      # ProcessAudioURL as written gives the same Booleans
        bits.append(b); images.append(i); haar.append(h)
    np.savez_compressed(os.path.join(HERE, "config1.npz"), pcm=pcm, bits=np.stack(bits), images=np.stack(images), haar=np.stack(haar),
                        score_01=np.float32(r.compare_pcm(cfg, pcm[0], pcm[1], 0)), score_10=np.float32(r.compare_pcm(cfg, pcm[1], pcm[0], 0)),
                        score_00=np.float32(r.compare_pcm(cfg, pcm[0], pcm[0], 0)), score_01_r100=np.float32(r.compare_pcm(cfg, pcm[0], pcm[1], 100)))
    # ---- known-answer vectors ----
    kat = {}
    h = np.array([[538, 940, 1940, 1794], [1840, 213, 1320, 913], [192, 591, 492, 1921]], np.float32)     # LBAudioDetectiveTests.m:160-162
    kat["haar_3x4_in"] = h.tolist(); kat["haar_3x4_out"] = r.haar(h).tolist()
    a = [1, 0, 0, 1, 0, 0, 1, 0]; b = [1, 0, 1, 0, 0, 1, 1, 0]
    kat["compare_toy"] = {"a": a, "b": b, "ab8": r.compare_sub(np.array(a), np.array(b), 8, 8), "ba8": r.compare_sub(np.array(b), np.array(a), 8, 8),
                          "ab4": r.compare_sub(np.array(a), np.array(b), 8, 4)}
    kat["band_tables"] = {}
    for n in (256, 512, 1024, 2048):
        idx, lo, hi = r.band_table(Cfg.default(window=n))
        kat["band_tables"][str(n)] = {"indices": idx.tolist(), "klow": lo.tolist(), "khigh": hi.tolist()}
    kat["set_window_size_status"] = {str(n): r.set_window_size_status(n) for n in (256, 512, 1000, 1024, 2048, 3000)}
    kat["subfp_counts"] = {"55120": 6, "165360": 19, "16536": 1}
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)
    # ---- matcher: random rank-sign codes, all count/range shapes the search kernel distinguishes ----
    rng = np.random.default_rng(20131017)
    cases = []
    for (c1, c2, L, rg) in [(19, 6, 200, 200), (6, 19, 200, 200), (6, 6, 200, 200), (19, 6, 200, 77), (5, 1, 200, 200), (1, 5, 100, 100), (5, 1, 400, 400),
                            (3, 0, 200, 200), (0, 0, 200, 200), (7, 7, 200, 13), (12, 3, 400, 250), (9, 9, 100, 1000)]:
        def codes(n):
            sign = rng.integers(0, 3, size=(n, L // 2))        # 0: none, 1: positive, 2: negative
            out = np.zeros((n, L), np.uint8); out[:, 0::2] = sign == 1; out[:, 1::2] = sign == 2
            return out
        f1, f2 = codes(c1), codes(c2)
        if c2 and c1 >= c2 and rng.random() < 0.5:
            f2[:] = f1[:c2]; flip = rng.random(f2.shape) < 0.05; f2 = np.where(flip, 1 - f2, f2).astype(np.uint8)   # a noisy excerpt, incl. illegal '11' codes
        cases.append({"c1": c1, "c2": c2, "L": L, "range": rg, "fp1": f1, "fp2": f2, "score": np.float32(r.compare_fp(f1, f2, rg))})
    np.savez_compressed(os.path.join(HERE, "compare_cases.npz"), n=len(cases),
                        **{"%s_%d" % (k, i): np.asarray(c[k]) for i, c in enumerate(cases) for k in c})
    # ---- config 5 sweep: 3 s noisy query (1 subfp) and a 9 s clip across window sizes and subfingerprint lengths ----
    sweep = {}
    base = r.synth_clip(7, 49608)            # 9 s
    query = r.add_noise(base[8192:8192 + 16536], 99, 0.0316)
    for n in (512, 1024, 2048):
        for L in (100, 200, 400):
            c = Cfg.default(window=n, sublen=L)
            sweep["clip_%d_%d" % (n, L)] = r.process(c, base, direct=True)
            sweep["query_%d_%d" % (n, L)] = r.process(c, query, direct=True)
    np.savez_compressed(os.path.join(HERE, "sweep.npz"), base=base, query=query, **sweep)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
