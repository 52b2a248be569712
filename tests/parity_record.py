"""Parity at the full sizes of BASELINE.json's configs, recorded: the CUDA path (through the C-ABI) against the compiled
reference (oracle/_ref, f64 FFT definition, all host threads) on identical PCM, with the figures the tolerances of
tests/test_gpu_parity.py are judged by written to a JSON file.

    python tests/parity_record.py --out gpurun_out/r02_parity.json            # full sizes (about 3 minutes on the GPU box)
    python tests/parity_record.py --clips 400 --db-clips 2000 ...             # reduced (what tests/test_parity_record.py runs)

TEST INFRASTRUCTURE: this is the one place outside the pytest files where the oracle is executed next to the product, and only as
the checker.  Per config and per window size it records
  * band energies: the largest pure per-element relative error, how many elements exceed 1e-4, how small those elements are
    relative to their image's largest energy, and the norm-wise error;
  * Haar coefficients: the largest error relative to the image's largest |coefficient|;
  * Booleans: mismatch count and rate, subfingerprints touched;
  * scores: largest |difference| of compare / search scores on the two sides' own bits, and on identical bits (must be 0).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 5512.0


class StageStats:
    """Accumulates band / Haar / bit statistics over chunks of (GPU, reference) stage dumps."""

    def __init__(self):
        self.elements = 0; self.above_1e4 = 0; self.above_5e4 = 0; self.above_1e3 = 0; self.beyond_floor = 0
        self.max_rel = 0.0; self.max_rel_beyond_floor = 0.0
        self.largest_violator_over_image_max = 0.0            # among elements whose pure relative error exceeds 1e-4: the largest |ref| / image max
        self.err2 = 0.0; self.ref2 = 0.0
        self.haar_max = 0.0; self.frames = 0
        self.bits = 0; self.bit_mismatches = 0; self.subfps = 0; self.subfps_touched = 0
        self.zero_ref = 0; self.zero_ref_mismatch = 0

    def add_images(self, got, want):
        got = got.reshape(-1, got.shape[-2] * got.shape[-1]).astype(np.float64); want = want.reshape(got.shape).astype(np.float64)
        imax = np.abs(want).max(axis=1, keepdims=True)
        err = np.abs(got - want)
        nz = want != 0
        self.zero_ref += int((~nz).sum()); self.zero_ref_mismatch += int((err[~nz] != 0).sum())
        rel = np.zeros_like(err); np.divide(err, np.abs(want), out=rel, where=nz)
        self.elements += int(nz.sum())
        viol = rel > 1e-4
        self.above_1e4 += int(viol.sum()); self.above_5e4 += int((rel > 5e-4).sum()); self.above_1e3 += int((rel > 1e-3).sum())
        self.max_rel = max(self.max_rel, float(rel.max()) if rel.size else 0.0)
        if viol.any():
            self.largest_violator_over_image_max = max(self.largest_violator_over_image_max, float((np.abs(want) / np.maximum(imax, 1e-300))[viol].max()))
        beyond = err > 1e-4 * np.abs(want) + 1e-9 * imax      # the tolerance of tests/test_gpu_parity.py
        self.beyond_floor += int(beyond.sum())
        excess = np.zeros_like(err); np.divide(np.maximum(err - 1e-9 * imax, 0), np.abs(want), out=excess, where=nz)
        self.max_rel_beyond_floor = max(self.max_rel_beyond_floor, float(excess.max()) if excess.size else 0.0)
        self.err2 += float((err ** 2).sum()); self.ref2 += float((want ** 2).sum())
        self.frames += got.shape[0]

    def add_haar(self, got, want):
        got = got.reshape(-1, got.shape[-2] * got.shape[-1]).astype(np.float64); want = want.reshape(got.shape).astype(np.float64)
        cmax = np.abs(want).max(axis=1)
        e = np.abs(got - want).max(axis=1)
        ok = cmax > 0
        if ok.any():
            self.haar_max = max(self.haar_max, float((e[ok] / cmax[ok]).max()))

    def add_bits(self, got, want):
        assert got.shape == want.shape, (got.shape, want.shape)
        d = got != want
        self.bits += d.size; self.bit_mismatches += int(d.sum())
        per = d.reshape(-1, d.shape[-1]).any(axis=1)
        self.subfps += per.size; self.subfps_touched += int(per.sum())

    def report(self):
        out = {}
        if self.elements:
            out["band_energies"] = {"elements": self.elements, "max_pure_relative_error": self.max_rel, "elements_above_1e-4": self.above_1e4,
                                    "elements_above_5e-4": self.above_5e4, "elements_above_1e-3": self.above_1e3,
                                    "fraction_above_1e-4": self.above_1e4 / self.elements,
                                    "largest_violator_relative_to_image_max": self.largest_violator_over_image_max,
                                    "elements_beyond_test_tolerance (1e-4 rel + 1e-9 of image max)": self.beyond_floor,
                                    "max_relative_error_beyond_the_floor": self.max_rel_beyond_floor,
                                    "normwise_error": (self.err2 / self.ref2) ** 0.5 if self.ref2 else 0.0,
                                    "reference_zero_elements": self.zero_ref, "reference_zero_elements_not_zero_on_gpu": self.zero_ref_mismatch, "frames": self.frames}
            out["haar"] = {"max_error_relative_to_image_max_coefficient": self.haar_max, "frames": self.frames}
        if self.bits:
            out["booleans"] = {"compared": self.bits, "mismatches": self.bit_mismatches, "mismatch_rate": self.bit_mismatches / self.bits,
                               "subfingerprints": self.subfps, "subfingerprints_touched": self.subfps_touched,
                               "subfingerprints_touched_fraction": self.subfps_touched / max(self.subfps, 1)}
        return out


def detective_for(lb, cfg):
    d = lb.Detective()
    d.set_window_size(cfg.window); d.set_analysis_stride(cfg.stride); d.set_pitch_steps(cfg.bands); d.set_subfingerprint_length(cfg.sublen); d.set_sample_rate(cfg.sample_rate)
    assert d.check_configuration() == 0
    return d


def compare_batch(lb, ref, cfg, pcm, threads, chunk, stats, fused=True, stages=True, log=None, cpu_f32_stats=None, cpu_f32_chunks=1):
    """pcm [clips][samples] host float32: GPU batch words + stage dumps against the reference's, chunk by chunk.  Returns (gpu bits, ref bits).
    cpu_f32_stats: for the first cpu_f32_chunks chunks, the SAME reference code run on a single-precision CPU FFT is held against the f64
    definition as well — the yardstick for what float32 arithmetic (a real vDSP included) does to these figures."""
    d = detective_for(lb, cfg)
    n = pcm.shape[0]
    words_all = d.process_batch(pcm)                                      # LBAudioDetectiveProcessPCMBatch: the e2e entry point, whole batch in one call
    got_bits_all = []; want_bits_all = []
    ref_secs = 0.0
    for c0 in range(0, n, chunk):
        part = pcm[c0:c0 + chunk]
        wbits, wimg, whaar, secs = ref.extract_batch_stages(cfg, part, threads=threads, images=stages, haar=stages)
        ref_secs += secs
        if stages:
            words, img, haar = d.process_batch_stages(part, fused=fused)
            assert np.array_equal(words, words_all[c0:c0 + chunk]), "stage-dump call and batch call disagree"
            stats.add_images(img, wimg); stats.add_haar(haar, whaar)
        if cpu_f32_stats is not None and c0 < cpu_f32_chunks * chunk and hasattr(ref, "extract_batch_stages"):
            fbits, fimg, fhaar, _ = ref.extract_batch_stages(cfg, part, threads=threads, images=stages, haar=stages, fft_mode="fast")
            if stages:
                cpu_f32_stats.add_images(fimg, wimg); cpu_f32_stats.add_haar(fhaar, whaar)
            cpu_f32_stats.add_bits(fbits, wbits)
        gbits = lb.unpack_words(words_all[c0:c0 + chunk], cfg.sublen)
        stats.add_bits(gbits, wbits)
        got_bits_all.append(gbits); want_bits_all.append(wbits)
        if log:
            log("  clips %d..%d of %d: reference %.1f s so far, %d mismatching Booleans so far" % (c0, c0 + part.shape[0], n, ref_secs, stats.bit_mismatches))
    return np.concatenate(got_bits_all), np.concatenate(want_bits_all), ref_secs


def device_synth(lb, n_clips, clip_len, first_clip_id=0, sample_rate=SR):
    """The bench's synthetic clips (device generator, global clip ids), brought to the host so both sides see the same floats."""
    import torch
    out = np.empty((n_clips, clip_len), np.float32)
    step = 2000
    for c0 in range(0, n_clips, step):
        nc = min(step, n_clips - c0)
        x = torch.empty((nc, clip_len), dtype=torch.float32, device="cuda")
        lb.synthesize_device(x.data_ptr(), nc, clip_len, clip_len, first_clip_id=first_clip_id + c0, sample_rate=sample_rate)
        torch.cuda.synchronize()
        out[c0:c0 + nc] = x.cpu().numpy()
        del x
    return out


def run(clips=10000, chunk=250, sweep_clips=2000, sweep_queries=500, db_clips=20000, db_queries=200, threads=None, log=print):
    import lbaudiodetective_b200 as lb
    from oracle import oracle as o
    lb.load_library(build_if_missing=False)
    if not lb.device_available():
        raise RuntimeError("parity_record needs a CUDA device")
    ref = o.best()
    threads = threads or (os.cpu_count() or 1)
    record = {"checker": ref.kind, "checker_fft": "f64 (exact DFT x2 rounded once to float32: the parity definition, oracle/shim/shim.c mode 0)",
              "host_threads": threads, "tolerances": {"band": "1e-4 relative (+ 1e-9 of the image maximum in the tests)", "haar": "1e-4 of the image's largest |coefficient|",
                                                      "booleans": "<= 1e-3 mismatching", "scores": "bit-exact on identical bits"}}
    t_all = time.time()

    # ---- config 1: two 10 s clips, stages + the compare-audio score ----
    cfg = o.Cfg.default()
    a, b = ref.synth_clip(0, 55120), ref.synth_clip(1, 55120)
    st = StageStats(); f32 = StageStats()
    gb, wb, _ = compare_batch(lb, ref, cfg, np.stack([a, b]), threads, 2, st, cpu_f32_stats=f32)
    d = lb.Detective()
    got = np.float32(d.compare_pcm(a, b, 0)); want = np.float32(ref.compare_pcm(cfg, a, b, 0))
    on_ref_bits = np.float32(lb.Fingerprint.from_booleans(wb[0]).compare(lb.Fingerprint.from_booleans(wb[1]), 200))
    rep = st.report()
    rep["yardstick: the reference on a float32 CPU FFT against the same f64 definition"] = f32.report()
    rep["score"] = {"LBAudioDetectiveComparePCM": float(got), "reference CompareAudioURLs": float(want), "abs_diff": float(abs(got - want)),
                    "gpu_compare_on_reference_bits": float(on_ref_bits), "abs_diff_on_identical_bits": float(abs(on_ref_bits - want))}
    record["config1 (two 10 s clips, compare-audio path)"] = rep
    log("config 1: %s" % json.dumps(rep))

    # ---- config 2: the bench workload, clips x 30 s, every Boolean and every stage element ----
    log("config 2: %d x 30 s clips (device-synthesised, the clips bench.py times), reference on %d threads" % (clips, threads))
    pcm = device_synth(lb, clips, 165360)
    st = StageStats(); f32 = StageStats()
    t0 = time.time()
    gb, wb, ref_secs = compare_batch(lb, ref, cfg, pcm, threads, chunk, st, log=log, cpu_f32_stats=f32, cpu_f32_chunks=8)
    rep = st.report()
    rep["yardstick: the reference on a float32 CPU FFT against the same f64 definition (first %d clips)" % min(clips, 8 * chunk)] = f32.report()
    rep["clips"] = clips; rep["reference_seconds"] = ref_secs; rep["wall_seconds"] = time.time() - t0
    rep["reference_audio_hours_per_s (f64 FFT + stage dumps, not the timing baseline)"] = clips * 30.0 / 3600.0 / ref_secs
    record["config2 (batch extraction of %d x 30 s clips)" % clips] = rep
    log("config 2: %s" % json.dumps(rep))

    # ---- config 4 shape: search scores on identical bits (the reference's bits of the clips above as the database) ----
    n_db = min(db_clips, clips)
    dbb = wb[:n_db]
    rng = np.random.default_rng(4)
    src = rng.integers(0, n_db, db_queries); off = rng.integers(0, 19 - 6 + 1, db_queries)
    qb = np.stack([dbb[c, o_:o_ + 6] for c, o_ in zip(src, off)]).copy()
    flip = rng.random((db_queries, 6, 100)) < 0.02                                 # noisy excerpts: sign flips
    qb[..., 0::2] ^= flip.astype(np.uint8); qb[..., 1::2] ^= flip.astype(np.uint8)
    db = lb.Database(200); db.add_packed(lb.pack_booleans(dbb))
    sc, idx, full = db.search_packed(lb.pack_booleans(qb), 10, all_scores=True)
    want_full, _ = ref.search(dbb, qb, 200, threads=threads)
    order = np.lexsort((np.broadcast_to(np.arange(n_db), want_full.shape), -want_full.astype(np.float64)), axis=1)[:, :10]
    record["config4 shape (%d six-subfingerprint queries x %d clips of 19, identical bits on both sides)" % (db_queries, n_db)] = {
        "scores_compared": int(full.size), "max_abs_score_diff": float(np.abs(full - want_full).max()), "scores_bit_identical": bool(np.array_equal(full, want_full)),
        "top10_indices_identical": bool(np.array_equal(idx, order.astype(np.uint32))), "top10_scores_identical": bool(np.array_equal(sc, np.take_along_axis(want_full, order, axis=1)))}
    # the same queries against the GPU's own bits of the same clips: what a user of the CUDA path sees (bits differ within the budget)
    db2 = lb.Database(200); db2.add_packed(lb.pack_booleans(gb[:n_db]))
    _, idx2, full2 = db2.search_packed(lb.pack_booleans(qb), 10, all_scores=True)
    record["config4 shape, GPU-extracted database against reference-extracted database"] = {
        "max_abs_score_diff": float(np.abs(full2 - want_full).max()), "top1_agree_fraction": float((idx2[:, 0] == order[:, 0]).mean())}
    log("config 4 shape: %s" % json.dumps(record["config4 shape (%d six-subfingerprint queries x %d clips of 19, identical bits on both sides)" % (db_queries, n_db)]))
    del pcm

    # ---- config 5: 3 s noisy queries vs 9 s clips, window x subfingerprint-length sweep ----
    sweep = {}
    base = device_synth(lb, sweep_clips, 49608, first_clip_id=500000)              # 9 s clips
    qsrc = np.arange(sweep_queries) % sweep_clips
    queries = np.stack([ref.add_noise(base[c, 8192:8192 + 16536], 7000 + i, 0.0158) for i, c in enumerate(qsrc)])      # 3 s excerpts + 1.58 % noise (essay p.34)
    for window in (512, 1024, 2048):
        wst = StageStats(); wf32 = StageStats()
        for sublen in (100, 200, 400):
            cfg5 = o.Cfg.default(window=window, sublen=sublen)
            st = StageStats()
            want_stage = sublen == 200
            gdb, wdb, _ = compare_batch(lb, ref, cfg5, base, threads, chunk * 4, st if not want_stage else wst, stages=want_stage,
                                        cpu_f32_stats=wf32 if want_stage else None, cpu_f32_chunks=1000)
            if want_stage:
                st.add_bits(gdb, wdb)
            gq, wq, _ = compare_batch(lb, ref, cfg5, queries, threads, chunk * 4, st, stages=False)
            # scores: the GPU search on its own bits against the reference's linear scan on its own bits, and on identical bits
            dbg = lb.Database(sublen); dbg.add_packed(lb.pack_booleans(gdb))
            _, idx_g, full_g = dbg.search_packed(lb.pack_booleans(gq), 1, all_scores=True)
            nq = min(64, sweep_queries)
            want_s, _ = ref.search(wdb, wq[:nq], sublen, threads=threads)
            dbr = lb.Database(sublen); dbr.add_packed(lb.pack_booleans(wdb))
            _, _, full_r = dbr.search_packed(lb.pack_booleans(wq[:nq]), 1, all_scores=True)
            r = st.report()["booleans"]
            r["score_max_abs_diff_own_bits"] = float(np.abs(full_g[:nq] - want_s).max())
            r["score_max_abs_diff_identical_bits"] = float(np.abs(full_r - want_s).max())
            r["recall_at_1_gpu"] = float((idx_g[:, 0] == qsrc.astype(np.uint32)).mean())
            r["recall_at_1_reference"] = float((want_s.argmax(axis=1) == qsrc[:nq]).mean())
            sweep["window %d, subfingerprint length %d" % (window, sublen)] = r
            log("config 5 window %d L %d: %s" % (window, sublen, json.dumps(r)))
        sweep["window %d, stages (subfingerprint length 200)" % window] = {k: v for k, v in wst.report().items() if k != "booleans"}
        sweep["window %d, stages, yardstick: the reference on a float32 CPU FFT against the same f64 definition" % window] = wf32.report()
        log("config 5 window %d stages: %s" % (window, json.dumps(sweep["window %d, stages (subfingerprint length 200)" % window])))
    record["config5 (%d x 9 s clips, %d noisy 3 s queries; window x subfingerprint-length sweep)" % (sweep_clips, sweep_queries)] = sweep
    record["total_seconds"] = time.time() - t_all
    return record


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_parity.json"))
    ap.add_argument("--clips", type=int, default=10000)
    ap.add_argument("--chunk", type=int, default=250)
    ap.add_argument("--sweep-clips", type=int, default=2000)
    ap.add_argument("--sweep-queries", type=int, default=500)
    ap.add_argument("--db-clips", type=int, default=20000)
    ap.add_argument("--db-queries", type=int, default=200)
    args = ap.parse_args()
    rec = run(args.clips, args.chunk, args.sweep_clips, args.sweep_queries, args.db_clips, args.db_queries)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rec, open(args.out, "w"), indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
