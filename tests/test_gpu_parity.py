"""Parity tests proper: the CUDA path, called through the C-ABI, against the oracle on identical inputs.

Tolerances (BASELINE.json north_star):
  * band energies: 1e-4 relative per element, with an absolute floor of 1e-9 x the image's largest energy: the
    reference's Q4 quirk (only positive parts are scaled down by 512) makes an energy ill-conditioned when its
    negative component is tiny, and a float32 FFT cannot resolve a band 9 orders of magnitude below the strongest
    one (a float32 CPU FFT standing in for vDSP misses the pure 1e-4 on 4x more elements than the kernels do).
    The pure per-element relative maximum is printed and bounded by 5e-4, the norm-wise error by 1e-6;
  * Haar coefficients: 1e-4 relative to the image's largest |coefficient| (small ones are cancellation results);
  * sign bits: <= 0.1 % mismatching Booleans (rank swaps where magnitudes tie within tolerance), reported;
  * everything downstream of identical inputs — Haar given images, bits given coefficients, scores and top-k given
    bits — bit-exact.
"""
import os

import ctypes as C

import numpy as np
import pytest
from oracle.oracle import Cfg

pytestmark = pytest.mark.gpu

BAND_RTOL = 1e-4
BAND_FLOOR_OF_MAX = 1e-9
BAND_RTOL_PURE_MAX = 5e-4
BAND_NORMWISE = 1e-6
HAAR_RTOL_OF_MAX = 1e-4
BIT_MISMATCH_BUDGET = 1e-3


def mismatch_rate(a, b):
    assert a.shape == b.shape
    return float((a != b).mean())


# ------------------------------------------------------------------ extraction ----

@pytest.mark.parametrize("fused", [True, False], ids=["fused", "generic"])
def test_config1_stages_against_golden(lb, config1, fused, figure):
    d = lb.Detective()
    worst_band = worst_haar = 0.0; mism = 0; total = 0
    for c in range(2):
        img, haar, bits = d.process_stages(config1["pcm"][c], fused=fused)
        gi, gh, gb = config1["images"][c], config1["haar"][c], config1["bits"][c]
        assert img.shape == gi.shape == (6, 128, 32)
        worst_band = max(worst_band, float((np.abs(img - gi) / np.abs(gi)).max()))
        floor = BAND_FLOOR_OF_MAX * np.abs(gi).reshape(6, -1).max(axis=1)[:, None, None]
        assert (np.abs(img - gi) <= BAND_RTOL * np.abs(gi) + floor).all()
        assert np.linalg.norm(img - gi) / np.linalg.norm(gi) < BAND_NORMWISE
        worst_haar = max(worst_haar, float((np.abs(haar - gh).reshape(6, -1).max(axis=1) / np.abs(gh).reshape(6, -1).max(axis=1)).max()))
        mism += int((bits != gb).sum()); total += bits.size
    figure("band rel err %.3g, haar err/max %.3g, bit mismatches %d/%d" % (worst_band, worst_haar, mism, total))
    assert worst_band < BAND_RTOL_PURE_MAX
    assert worst_haar < HAAR_RTOL_OF_MAX
    assert mism / total <= BIT_MISMATCH_BUDGET


@pytest.mark.parametrize("window", [1024, 512, 256])
def test_register_fft_kernel_other_windows(lb, checker, window, figure):
    """Windows of 1024 / 512 / 256 samples run the same register-FFT kernel with 2 / 4 / 8 windows per warp: stages against the oracle."""
    cfg = Cfg.default(window=window); d = lb.Detective(); d.set_window_size(window)
    pcm = checker.synth_clip(33, 40000)
    want_bits, want_img, want_haar = checker.process(cfg, pcm, stages=True) if checker.kind == "port" else checker.process(cfg, pcm, stages=True, direct=True)
    for fused in (True, False):
        img, haar, bits = d.process_stages(pcm, fused=fused)
        assert img.shape == want_img.shape
        floor = BAND_FLOOR_OF_MAX * np.abs(want_img).reshape(want_img.shape[0], -1).max(axis=1)[:, None, None]
        empty = want_img == 0                                     # short windows have empty bands (klow == khigh): energy exactly 0
        assert np.array_equal(img[empty], want_img[empty])
        excess = np.abs(img - want_img) - floor
        rel = float((np.maximum(excess, 0)[~empty] / np.abs(want_img[~empty])).max())
        figure("window %d %s: band rel err beyond the floor %.3g, %d of %d elements above 1e-4" % (window, "register" if fused else "generic", rel, int((excess[~empty] > 1e-4 * np.abs(want_img[~empty])).sum()), img.size))
        assert rel < BAND_RTOL_PURE_MAX * 2, (window, fused)      # short windows: 1-3 bins per band, no averaging of the ill-conditioned Q4 terms
        assert np.linalg.norm(img - want_img) / np.linalg.norm(want_img) < BAND_NORMWISE
        assert float((np.abs(haar - want_haar).reshape(haar.shape[0], -1).max(axis=1) / np.abs(want_haar).reshape(haar.shape[0], -1).max(axis=1)).max()) < HAAR_RTOL_OF_MAX
        assert mismatch_rate(bits, want_bits) <= 2e-3, (window, fused, int((bits != want_bits).sum()))


def test_fused_equals_generic_bits(lb, port):
    """Two independent kernels (register radix-32 FFT vs shared-memory radix-2 FFT) agree within the bit budget."""
    d = lb.Detective(); pcm = port.synth_clip(31, 165360)
    _, _, bf = d.process_stages(pcm, fused=True)
    _, _, bg = d.process_stages(pcm, fused=False)
    assert bf.shape == (19, 200) and mismatch_rate(bf, bg) <= BIT_MISMATCH_BUDGET


def test_process_pcm_against_oracle(lb, checker, figure):
    cfg = Cfg.default(); d = lb.Detective(); mism = 0; total = 0
    for clip, n in ((40, 55120), (41, 165360), (42, 16536), (43, 10240), (44, 10240 + 8191), (45, 10240 + 8192)):
        pcm = checker.synth_clip(clip, n)
        fp = d.process_pcm(pcm); want = checker.process(cfg, pcm)
        assert fp.count == want.shape[0] and fp.subfingerprint_length == 200
        got = fp.booleans(); mism += int((got != want).sum()); total += want.size
        assert np.array_equal(lb.unpack_words(fp.packed(), 200), got)
    figure("bit mismatches %d/%d vs %s oracle" % (mism, total, checker.kind))
    assert mism / total <= BIT_MISMATCH_BUDGET


def test_bit_mismatch_rate_on_a_large_sample(lb, checker, figure):
    """The 0.1 % budget measured where it means something: 120 x 30 s clips (2,280 subfingerprints, 456,000 Booleans) plus quiet,
    loud and tonal variants, against the oracle run on all host threads."""
    cfg = Cfg.default(); d = lb.Detective()
    base = np.stack([checker.synth_clip(500 + i, 165360) for i in range(120)])
    t = np.arange(165360) / 5512.0
    variants = {"chirp+tone+noise": base,
                "x 1e-3 (quiet)": (base[:24] * np.float32(1e-3)).astype(np.float32),
                "x 30 (clipped range exceeded)": (base[:24] * np.float32(30.0)).astype(np.float32),
                "two pure tones, no noise": np.stack([(0.5 * np.sin(2 * np.pi * (440.0 + 7 * i) * t) + 0.3 * np.sin(2 * np.pi * (1234.5 + 3 * i) * t)).astype(np.float32) for i in range(12)])}
    worst = 0.0
    for name, pcm in variants.items():
        want, _ = checker.extract_batch(cfg, pcm, threads=os.cpu_count() or 1)
        got = lb.unpack_words(d.process_batch(pcm), 200)
        assert got.shape == want.shape
        rate = mismatch_rate(got, want)
        per_sub = (got != want).reshape(-1, 200).any(axis=1).mean()
        figure("%-32s %8d Booleans, mismatch rate %.2e (%.2f %% of subfingerprints touched)" % (name, want.size, rate, 100 * per_sub))
        if "pure tones" not in name:
            worst = max(worst, rate)
    assert worst <= BIT_MISMATCH_BUDGET


@pytest.mark.parametrize("kernel", ["register", "generic"])
def test_transform_images_bit_exact(lb, checker, config1, kernel, monkeypatch):
    """Given identical spectral images, Haar (IEEE divides, same order) and the ordered top-t are bit-exact — in the register kernel
    the batch path runs for 32 bands (haar_select32_kernel) and in the any-geometry one; the images include magnitudes outside the
    range of the short constant division (the register kernel then redoes the image with checked divisions), ties and silence."""
    if kernel == "generic":
        monkeypatch.setenv("LBAD_TRANSFORM", "generic")
    d = lb.Detective()
    rng = np.random.default_rng(7)
    images = np.concatenate([config1["images"][0], (rng.random((4, 128, 32)) ** 4 * 50).astype(np.float32)])
    ties = np.zeros((1, 128, 32), np.float32); ties[0, 5, 3] = 2.0; ties[0, 6, 3] = -2.0; ties[0, 100, 31] = 2.0; ties[0, 64:, :8] = 1.0
    tiny = (rng.random((2, 128, 32)).astype(np.float32) * np.float32(1e-30)); tiny[1] *= np.float32(1e-8)          # below 2^-100, down to denormals
    huge = (rng.random((1, 128, 32)).astype(np.float32) * np.float32(3e37))
    mixed = (rng.random((1, 128, 32)) ** 4 * 50).astype(np.float32); mixed[0, 17, 5] = np.float32(1e-36); mixed[0, 90, 30] = np.float32(2e37)
    images = np.concatenate([images, ties, tiny, huge, mixed, np.zeros((1, 128, 32), np.float32)])
    haar, bits = d.transform_images(images)
    for i in range(images.shape[0]):
        want_h = checker.haar(images[i])
        assert np.array_equal(haar[i], want_h), i
        assert np.array_equal(bits[i], checker.extract_bits(want_h, 200)[:200]), i
    assert bits[-1].sum() == 0                                            # silence: every coefficient is zero -> no bit set


@pytest.mark.parametrize("kernel", ["register", "generic"])
def test_bits_from_coefficients_with_exact_ties(lb, checker, kernel, monkeypatch):
    """Top-t on coefficient sets full of exact magnitude ties: the stable order (lower flat index first) must hold."""
    if kernel == "generic":
        monkeypatch.setenv("LBAD_TRANSFORM", "generic")
    d = lb.Detective(); rng = np.random.default_rng(8)
    # images whose Haar transform is easy to control: a constant image has a single non-zero coefficient; use small integers instead
    images = rng.integers(0, 3, size=(6, 128, 32)).astype(np.float32)
    haar, bits = d.transform_images(images)
    for i in range(6):
        want_h = checker.haar(images[i])
        assert np.array_equal(haar[i], want_h)
        assert np.array_equal(bits[i], checker.extract_bits(want_h, 200)[:200])


def test_config5_sweep_against_golden(lb, sweep, figure):
    """Window-size x subfingerprint-length sweep (BASELINE config 5) through the generic path; budget over the sweep."""
    mism = total = 0
    for window in (512, 1024, 2048):
        for sublen in (100, 200, 400):
            d = lb.Detective(); d.set_window_size(window); d.set_subfingerprint_length(sublen)
            assert d.check_configuration() == 0
            for name in ("clip", "query"):
                want = sweep["%s_%d_%d" % (name, window, sublen)]
                fp = d.process_pcm(sweep["base" if name == "clip" else "query"])
                assert fp.count == want.shape[0] and fp.subfingerprint_length == sublen
                m = int((fp.booleans() != want).sum()); mism += m; total += want.size
                if m:
                    figure("window %d sublen %d %s: %d mismatching Booleans" % (window, sublen, name, m))
    figure("sweep: %d/%d mismatching Booleans" % (mism, total))
    assert mism / total <= BIT_MISMATCH_BUDGET


def test_other_geometry_against_oracle(lb, checker, figure):
    cases = [dict(window=256, stride=64), dict(stride=128, sample_rate=8000.0), dict(bands=16, sublen=64), dict(stride=50), dict(bands=64, sublen=512),
             dict(stride=2), dict(window=1024, stride=1000),
             # window 2048 at hop 64 with other band tables: the carried-transform kernel with run-time band rows (rows above 23, rows below 2,
             # rows needed only as mirror images)
             dict(sample_rate=6000.0), dict(sample_rate=4100.0), dict(sample_rate=4800.0), dict(sample_rate=8000.0), dict(sample_rate=22050.0)]
    mism = total = 0
    for kw in cases:
        cfg = Cfg.default(**kw); d = lb.Detective()
        d.set_window_size(cfg.window); d.set_analysis_stride(cfg.stride); d.set_pitch_steps(cfg.bands); d.set_subfingerprint_length(cfg.sublen); d.set_sample_rate(cfg.sample_rate)
        assert d.check_configuration() == 0, kw
        n = 40000 if cfg.stride <= 128 else 128 * cfg.stride * 2 + cfg.window + 7
        pcm = checker.synth_clip(50, n, cfg.sample_rate)
        want = checker.process(cfg, pcm) if checker.kind == "port" else checker.process(cfg, pcm, direct=True)
        got = d.process_pcm(pcm).booleans()
        assert got.shape == want.shape and want.shape[0] >= 1, kw
        m = int((got != want).sum()); mism += m; total += want.size
        if m:
            figure("%s: %d mismatching Booleans of %d" % (kw, m, want.size))
    assert mism / total <= BIT_MISMATCH_BUDGET


def test_invariants_of_the_reference_tests(lb, port):
    pcm = port.synth_clip(60, 82680)
    a = lb.Detective().process_pcm(pcm); b = lb.Detective().process_pcm(pcm)
    assert a.equal(b)                                                     # testFingerprintVersatility (Tests.m:119-139)
    assert a.copy().equal(a)                                              # testFingerprintComparison (Tests.m:141-155)
    assert a.compare(a, 200) == 1.0


def test_edge_cases(lb, port):
    d = lb.Detective()
    st, fp = d.process_pcm(np.zeros(2047, np.float32), check=False)
    assert st == lb.ARGUMENT_INVALID and fp.count == 0                    # shorter than one window (upstream underflows, m:250)
    fp = d.process_pcm(np.zeros(10239, np.float32)); assert fp.count == 0 # not enough windows for one frame
    fp = d.process_pcm(np.zeros(10240, np.float32)); assert fp.count == 1 and fp.booleans().sum() == 0
    big = np.full(10240, 1e30, np.float32)                                # energies overflow to inf and are skipped (m:398-401)
    fp = d.process_pcm(big); assert fp.count == 1
    d.set_window_size(4096); st, fp = d.process_pcm(np.zeros(20000, np.float32), check=False)
    assert st == lb.ARGUMENT_INVALID


def test_non_finite_samples_against_oracle(lb, checker):
    """A NaN or Inf sample poisons every window that contains it: all of that window's bins are non-finite, are skipped (m:398-401,
    SURVEY.md Q6) and its image row is zero; the other windows are untouched.  Same bits as the oracle."""
    cfg = Cfg.default(); d = lb.Detective()
    pcm = checker.synth_clip(70, 55120).copy()
    pcm[12345] = np.nan; pcm[30000] = np.inf; pcm[30001] = -np.inf
    want = checker.process(cfg, pcm)
    got = d.process_pcm(pcm).booleans()
    assert got.shape == want.shape
    assert (got != want).mean() <= BIT_MISMATCH_BUDGET
    img, _, _ = d.process_stages(pcm, fused=True)
    first = (12345 - 2048) // 64 + 1; last = 12345 // 64                     # windows containing sample 12345
    rows = img.reshape(-1, 32)
    assert (rows[first:last + 1] == 0).all() and (rows[first - 1] != 0).any() and (rows[last + 1] != 0).any()


def test_batch_equals_single(lb, port):
    d = lb.Detective()
    pcm = np.stack([port.synth_clip(70 + i, 55120) for i in range(5)])
    words = d.process_batch(pcm)
    assert words.shape == (5, 6, 8)
    for i in range(5):
        assert np.array_equal(words[i], d.process_pcm(pcm[i]).packed())
    # ragged stride / unaligned clip length take the non-TMA staging path and must agree
    odd = np.ascontiguousarray(pcm[:, :55001])
    w2 = d.process_batch(odd)
    for i in range(5):
        assert np.array_equal(w2[i], d.process_pcm(odd[i]).packed())


def test_slabs_and_staging_variants_agree(lb, port, monkeypatch):
    """The image scratch is processed in slabs of frames and the samples are staged by TMA or by plain loads: same words either way."""
    pcm = np.stack([port.synth_clip(120 + i, 82680) for i in range(6)])      # 9 frames per clip, 54 frames
    want = lb.Detective().process_batch(pcm)
    monkeypatch.setenv("LBAD_SLAB_FRAMES", "7")                              # slabs that cut through clips
    assert np.array_equal(lb.Detective().process_batch(pcm), want)
    monkeypatch.delenv("LBAD_SLAB_FRAMES")
    monkeypatch.setenv("LBAD_STAGE", "ldg")
    assert np.array_equal(lb.Detective().process_batch(pcm), want)


def test_frames_split_between_ctas_agree(lb, port, monkeypatch):
    """Calls that bring few frames (a single clip) spread each frame over eight CTAs of 16 windows, big batches give a CTA a whole
    frame: the same spectral images, bit for bit — both ways forced on the same input, by TMA and by plain loads, and as it happens by
    itself (one clip alone against the clip inside a batch of 200 frames)."""
    x = port.synth_clip(77, 165360)
    out = {}
    for subs in ("1", "8"):
        monkeypatch.setenv("LBAD_SUBFRAMES", subs)
        img, haar, bits = lb.Detective().process_stages(x, fused=True)
        monkeypatch.setenv("LBAD_STAGE", "ldg")
        img2, _, bits2 = lb.Detective().process_stages(x[:165001], fused=True)       # odd length: plain-load staging, 19 frames all the same
        monkeypatch.delenv("LBAD_STAGE")
        out[subs] = (img, bits, img2, bits2)
    monkeypatch.delenv("LBAD_SUBFRAMES")
    assert out["1"][0].shape == (19, 128, 32)
    for a, b in zip(out["1"], out["8"]):
        assert np.array_equal(a, b)
    assert np.array_equal(out["1"][0], out["1"][2])
    # the same with a band table whose rows are only known at run time (another processing rate)
    y = port.synth_clip(78, 60000, 6000.0); alt = {}
    for subs in ("1", "8"):
        monkeypatch.setenv("LBAD_SUBFRAMES", subs)
        d6 = lb.Detective(); d6.set_sample_rate(6000.0)
        alt[subs] = d6.process_stages(y, fused=True)
    monkeypatch.delenv("LBAD_SUBFRAMES")
    assert alt["1"][0].shape[0] == 7 and np.array_equal(alt["1"][0], alt["8"][0]) and np.array_equal(alt["1"][2], alt["8"][2])
    pcm = np.stack([port.synth_clip(500 + i, 82680) for i in range(23)])     # 207 frames: a frame per CTA
    d = lb.Detective(); batch = d.process_batch(pcm)
    for i in (0, 11, 22):
        assert np.array_equal(d.process_pcm(pcm[i]).packed(), batch[i])      # 9 frames: eight CTAs per frame


def test_device_resident_batch(lb, port):
    import torch
    d = lb.Detective(); n, clip_len = 64, 165360
    x = torch.empty((n, clip_len), dtype=torch.float32, device="cuda")
    lb.synthesize_device(x.data_ptr(), n, clip_len, clip_len, first_clip_id=1000)
    out = torch.zeros((n, 19, 8), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    d.process_batch_device(x.data_ptr(), n, clip_len, clip_len, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host = d.process_batch(x.cpu().numpy())
    assert np.array_equal(out.cpu().numpy().view(np.uint32), host)
    bits = lb.unpack_words(host, 200)
    assert bits.reshape(-1, 100, 2).sum(axis=2).max() <= 1                # never both sign bits of a rank
    assert bits.sum() == bits.shape[0] * bits.shape[1] * 100              # noisy audio: every selected wavelet is non-zero
    assert d.kernel_launches >= 2


def test_int16_batch_equals_float_batch(lb, port):
    """Signed 16-bit PCM uploaded as 2 bytes per sample and converted on the device (x / 32768, exact) gives the same words."""
    d = lb.Detective()
    f = np.stack([port.synth_clip(80 + i, 55120) for i in range(4)])
    i16 = np.clip(np.round(f * 32767.0), -32768, 32767).astype(np.int16)
    assert np.array_equal(d.process_batch_int16(i16), d.process_batch(i16.astype(np.float32) / np.float32(32768.0)))
    odd = np.ascontiguousarray(i16[:, :55003])
    assert np.array_equal(d.process_batch_int16(odd), d.process_batch(odd.astype(np.float32) / np.float32(32768.0)))


def test_streaming_equals_one_shot(lb, checker, figure):
    """Any chunking of the PCM gives, at every prefix, what the reference's whole-clip loop (m:250-262) gives for that prefix: the
    stream's Booleans are held against the oracle run on pcm[:pos] (within the bit budget) and against the one-shot GPU call (equal)."""
    cfg = Cfg.default()
    d = lb.Detective(); pcm = checker.synth_clip(90, 100000); rng = np.random.default_rng(15)
    s = lb.Stream(d); pos = 0; mism = total = 0
    while pos < len(pcm):
        n = int(rng.choice([1, 63, 64, 1000, 8192, 10175, 10176, 20000])); n = min(n, len(pcm) - pos)
        s.append(pcm[pos:pos + n]); pos += n
        want = d.subfingerprints_for_length(pos)
        fp = s.fingerprint()
        assert fp.count == want and s.pending == pos - want * 8192
        if n >= 8192 and want:
            assert fp.equal(d.process_pcm(pcm[:pos]))
            ob = checker.process(cfg, pcm[:pos])                             # the oracle on the same prefix
            gb = fp.booleans()
            assert ob.shape == gb.shape
            mism += int((ob != gb).sum()); total += ob.size
    whole = checker.process(cfg, pcm)
    final = s.fingerprint().booleans()
    assert final.shape == whole.shape == (11, 200)
    mism += int((final != whole).sum()); total += whole.size
    figure("stream prefixes vs %s oracle: %d mismatching Booleans of %d" % (checker.kind, mism, total))
    assert mism / total <= BIT_MISMATCH_BUDGET
    assert s.fingerprint().equal(d.process_pcm(pcm)) and s.fingerprint().count == 11
    s2 = lb.Stream(d); s2.append(pcm[:10239]); assert s2.fingerprint().count == 0
    s2.append(pcm[10239:10240]); assert s2.fingerprint().count == 1          # exactly the reference's threshold: 128*64 + 2048 samples


# --------------------------------------------------------------------- matching ----

def test_compare_toy_vectors(lb, kat):
    t = kat["compare_toy"]; a = np.array(t["a"], np.uint8); b = np.array(t["b"], np.uint8)
    fp = lb.Fingerprint(8)
    assert fp.compare_subfingerprints(a, b, 8) == t["ab8"]
    assert fp.compare_subfingerprints(b, a, 8) == t["ba8"]
    assert fp.compare_subfingerprints(a, b, 4) == t["ab4"]
    assert fp.compare_subfingerprints(np.zeros(8, np.uint8), b, 8) == 0.0  # possible == 0 (FP.m:171-173)


def test_compare_cases_bit_exact(lb, compare_cases):
    for c in compare_cases:
        f1 = lb.Fingerprint.from_booleans(c["fp1"].reshape(int(c["c1"]), int(c["L"]))) if int(c["c1"]) else lb.Fingerprint(int(c["L"]))
        f2 = lb.Fingerprint.from_booleans(c["fp2"].reshape(int(c["c2"]), int(c["L"]))) if int(c["c2"]) else lb.Fingerprint(int(c["L"]))
        got = np.float32(f1.compare(f2, int(c["range"])))
        assert got == c["score"], (int(c["c1"]), int(c["c2"]), int(c["L"]), int(c["range"]), got, c["score"])


def test_compare_random_bit_exact(lb, checker):
    rng = np.random.default_rng(11)
    for _ in range(60):
        L = int(rng.choice([100, 200, 400])); c1 = int(rng.integers(0, 9)); c2 = int(rng.integers(0, 9)); rg = int(rng.integers(1, L + 40))
        if c1 + c2 == 0:
            continue
        b1 = (rng.random((c1, L)) < 0.4).astype(np.uint8); b2 = (rng.random((c2, L)) < 0.4).astype(np.uint8)
        f1 = lb.Fingerprint.from_booleans(b1) if c1 else lb.Fingerprint(L); f2 = lb.Fingerprint.from_booleans(b2) if c2 else lb.Fingerprint(L)
        assert np.float32(f1.compare(f2, rg)) == np.float32(checker.compare_fp(b1, b2, rg)), (L, c1, c2, rg)


def test_compare_pcm_config1(lb, config1):
    d = lb.Detective(); p = config1["pcm"]
    got01 = np.float32(d.compare_pcm(p[0], p[1], 0)); got00 = np.float32(d.compare_pcm(p[0], p[0], 0))
    assert got00 == 1.0
    # bits may differ within the 0.1 % budget, which moves the score by at most a few 1/possible steps
    assert abs(got01 - config1["score_01"]) <= 0.01
    f0 = lb.Fingerprint.from_booleans(config1["bits"][0]); f1 = lb.Fingerprint.from_booleans(config1["bits"][1])
    assert np.float32(f0.compare(f1, 200)) == config1["score_01"]          # on identical bits: bit-exact
    assert np.float32(f1.compare(f0, 200)) == config1["score_10"]
    assert np.float32(f0.compare(f1, 100)) == config1["score_01_r100"]


def test_compare_pcm_equals_its_three_steps(lb, port):
    """LBAudioDetectiveComparePCM runs upload, fingerprinting and comparison as one device pipeline (one two-clip batch when the lengths
    agree): the match must be the one of ProcessPCM twice + CompareToFingerprint (LBAudioDetective.m:442-464), bit for bit, for either
    argument order, any range, growing and shrinking clips, and with the batch entry point used in between (they share device buffers)."""
    d = lb.Detective()
    clips = {n: port.synth_clip(300 + i, n) for i, n in enumerate((55120, 55121, 165360, 16536, 12288, 400000))}
    fps = {n: d.process_pcm(x) for n, x in clips.items()}
    for a, b in ((55120, 55120), (55120, 55121), (55120, 165360), (165360, 55120), (16536, 165360), (12288, 16536), (400000, 55120), (55120, 16536)):
        for rng in (0, 200, 100, 7, 1000):
            want = np.float32(fps[a].compare(fps[b], rng if rng else 200))
            assert np.float32(d.compare_pcm(clips[a], clips[b], rng)) == want, (a, b, rng)
        if a == 165360:
            w = d.process_batch(np.stack([clips[55120], clips[55121][:55120]]))
            assert np.array_equal(w[0], fps[55120].packed())
    assert np.float32(d.compare_pcm(clips[55120], clips[55120], 0)) == 1.0
    # a clip too short for one subfingerprint takes the step-by-step route: compared with an empty fingerprint the match is 0 (FP.m:133)
    assert d.compare_pcm(clips[55120], clips[55120][:9000], 0) == 0.0
    d2 = lb.Detective(); d2.set_window_size(1024); d2.set_subfingerprint_length(100)
    f = [d2.process_pcm(clips[n]) for n in (55120, 165360)]
    assert np.float32(d2.compare_pcm(clips[55120], clips[165360], 0)) == np.float32(f[0].compare(f[1], 100))


def test_sharded_batch_equals_one_detective(lb, port, monkeypatch):
    """LBAudioDetectiveProcessPCMBatchSharded: the clips of one batch spread over several detectives (one per GPU where there are
    several; here they share the devices round-robin), each on its own host thread taking chunks from a shared cursor — same words as
    one detective, with one clip, two clips or the whole batch per chunk (LBAD_CHUNK_CLIPS pins the chunk size, read when a plan is
    built), for fewer clips than detectives, and with a non-default geometry."""
    pcm = np.stack([port.synth_clip(40 + i, 55120) for i in range(7)])
    n_dev = lb.device_count()
    for window, sublen, chunk in ((2048, 200, "1"), (2048, 200, "2"), (2048, 200, None), (1024, 100, "1")):
        if chunk: monkeypatch.setenv("LBAD_CHUNK_CLIPS", chunk)
        else: monkeypatch.delenv("LBAD_CHUNK_CLIPS", raising=False)
        dets = []
        for i in range(3):
            d = lb.Detective(); d.set_window_size(window); d.set_subfingerprint_length(sublen)
            assert d.set_device(i % n_dev) == 0 and d.device == i % n_dev
            dets.append(d)
        one = lb.Detective(); one.set_window_size(window); one.set_subfingerprint_length(sublen)
        want = one.process_batch(pcm)
        assert np.array_equal(lb.Detective.process_batch_sharded(dets, pcm), want)
        assert np.array_equal(lb.Detective.process_batch_sharded(dets, pcm[:2]), want[:2])
        assert np.array_equal(lb.Detective.process_batch_sharded(dets[:1], pcm), want)
    monkeypatch.delenv("LBAD_CHUNK_CLIPS", raising=False)
    assert dets[0].set_device(n_dev) == lb.ARGUMENT_INVALID
    dets[1].set_subfingerprint_length(200)
    with pytest.raises(lb.LBADError):
        lb.Detective.process_batch_sharded(dets, pcm)


def test_frame_api_on_the_gpu(lb, ref, kat):
    """The reference's Frame API (include/LBAudioDetectiveFrame.h): Decompose and ExtractFingerprint run on the GPU for any shape and
    must equal the compiled reference's own functions bit for bit — first the 3 x 4 frame of the reference's Haar test
    (LBAudioDetectiveTests.m:158-172; expected values in tests/golden/kat.json came from the compiled reference), then shapes that are
    not powers of two (integer halving leaves their tails alone, Frame.m:144-152), ties, zeros and every wavelet count up to the size."""
    f = lb.Frame.from_array(np.array(kat["haar_3x4_in"], np.float32))
    f.decompose()
    assert np.array_equal(f.array(), np.array(kat["haar_3x4_out"], np.float32).reshape(3, 4))
    rng = np.random.default_rng(77)
    for rows, cols in ((3, 4), (1, 1), (1, 9), (5, 7), (128, 32), (17, 64), (100, 33), (2, 300)):
        img = (rng.standard_normal((rows, cols)) * 10.0 ** rng.integers(-3, 4)).astype(np.float32)
        if rows * cols > 8:
            img.flat[3] = 0.0; img.flat[5] = -img.flat[4]                      # a zero and a pair of equal magnitudes survive as ties in the input
        fr = lb.Frame.from_array(img)
        assert fr.decompose(status=True) == 0
        want = ref.haar(img)
        assert np.array_equal(fr.array(), want), (rows, cols)
        coef = want.copy()
        coef.flat[::3] = np.round(coef.flat[::3], 1)                           # plenty of exactly equal magnitudes, zeros included
        fc = lb.Frame.from_array(coef)
        n = rows * cols
        for t in sorted({1, min(n, 2), n // 2, n - 1, n} - {0}):
            assert np.array_equal(fc.extract_fingerprint(t), ref.extract_bits(coef, t)), (rows, cols, t)
    # like upstream, ExtractFingerprint only ever sets TRUE: what the caller's buffer held stays
    out = np.ones(8, np.uint8); fr = lb.Frame.from_array(np.array([[1.0, -2.0], [0.0, 3.0]], np.float32))
    lb.lib().LBAudioDetectiveFrameExtractFingerprint(fr.ref, 4, out.ctypes.data_as(C.c_void_p))
    assert out.all()


def rank_sign_codes(rng, n, count, L):
    sign = rng.integers(0, 2, size=(n, count, L // 2))
    out = np.zeros((n, count, L), np.uint8); out[..., 0::2] = sign == 0; out[..., 1::2] = sign == 1
    return out


@pytest.mark.parametrize("L,q_count,db_count,rng_len", [(200, 6, 19, 0), (200, 1, 5, 0), (200, 6, 19, 77), (100, 1, 5, 0), (400, 1, 5, 0),
                                                         (200, 3, 19, 0), (200, 6, 4, 0), (200, 6, 6, 0), (400, 6, 19, 300)])
def test_search_scores_and_topk_bit_exact(lb, checker, L, q_count, db_count, rng_len):
    rng = np.random.default_rng(100 + L + q_count + db_count + rng_len)
    n_db, n_q, k = 300, 37, 10
    dbb = rank_sign_codes(rng, n_db, db_count, L)
    qb = rank_sign_codes(rng, n_q, q_count, L)
    for q in range(0, n_q, 3):                                             # plant noisy excerpts so the top of the list is interesting
        c = int(rng.integers(0, n_db)); n = min(q_count, db_count); o = int(rng.integers(0, db_count - n + 1))
        qb[q, :n] = dbb[c, o:o + n]; flip = rng.random(qb[q].shape) < 0.03; qb[q] = np.where(flip, 1 - qb[q], qb[q])
    db = lb.Database(L); db.add_packed(lb.pack_booleans(dbb))
    assert db.clips == n_db and db.subfingerprints == n_db * db_count
    sc, idx, full = db.search_packed(lb.pack_booleans(qb), k, rng=rng_len, all_scores=True)
    want, _ = checker.search(dbb, qb, rng_len if rng_len else L)
    assert np.array_equal(full, want)
    order = np.lexsort((np.arange(n_db)[None, :].repeat(n_q, 0), -want.astype(np.float64)), axis=1)[:, :k]     # score desc, clip asc
    assert np.array_equal(idx, order.astype(np.uint32))
    assert np.array_equal(sc, np.take_along_axis(want, order, axis=1))
    assert db.compares_per_query(q_count) == n_db * (abs(db_count - q_count) + 1) * min(db_count, q_count)


@pytest.mark.parametrize("L,q_count,db_count,rng_len,n_q", [(200, 6, 19, 0, 1), (200, 6, 19, 77, 5), (200, 1, 5, 0, 8), (100, 3, 7, 0, 2), (100, 1, 5, 31, 7),
                                                             (400, 6, 19, 300, 3), (400, 2, 9, 0, 8), (200, 6, 6, 0, 4), (200, 4, 40, 0, 2)])
def test_few_query_kernel_bit_exact(lb, checker, L, q_count, db_count, rng_len, n_q):
    """Up to 8 queries run with lane = clip (search_few_kernel): the whole score matrix and the top-k must equal the oracle's, for
    every word count (L = 100 / 200 / 400), shortened ranges, irregular codes ('00' and '11' ranks on either side) and ragged clips."""
    rng = np.random.default_rng(700 + L + q_count + db_count + rng_len + n_q)
    n_db, k = 777, 10
    dbb = rank_sign_codes(rng, n_db, db_count, L)
    qb = rank_sign_codes(rng, n_q, q_count, L)
    for q in range(0, n_q, 2):
        c = int(rng.integers(0, n_db)); o = int(rng.integers(0, db_count - q_count + 1))
        qb[q] = dbb[c, o:o + q_count]; flip = rng.random(qb[q].shape) < 0.02; qb[q] = np.where(flip, 1 - qb[q], qb[q])
    dbb[5, 1, 0:2] = 0; dbb[6, 0, 2:4] = 1; dbb[7, :, :] = 0                  # '00', '11', an all-zero clip (possible = 0 everywhere)
    qb[-1, 0, 4:6] = 0
    db = lb.Database(L); db.add_packed(lb.pack_booleans(dbb))
    sc, idx, full = db.search_packed(lb.pack_booleans(qb), k, rng=rng_len, all_scores=True)
    want, _ = checker.search(dbb, qb, rng_len if rng_len else L)
    assert np.array_equal(full, want)
    order = np.lexsort((np.arange(n_db)[None, :].repeat(n_q, 0), -want.astype(np.float64)), axis=1)[:, :k]
    assert np.array_equal(idx, order.astype(np.uint32)) and np.array_equal(sc, np.take_along_axis(want, order, axis=1))
    # ragged: clips of different lengths (all at least as long as the query) through the fingerprint API
    if q_count <= 4:
        db2 = lb.Database(L); counts = rng.integers(q_count, q_count + 9, 50)
        fps = [rank_sign_codes(rng, 1, int(c), L)[0] for c in counts]
        for f in fps: db2.add_fingerprint(lb.Fingerprint.from_booleans(f))
        qf = lb.Fingerprint.from_booleans(fps[17][:q_count])
        sc2, idx2 = db2.search([qf], k=5)
        want2 = np.array([np.float32(checker.compare_fp(f, fps[17][:q_count], L)) for f in fps])
        o2 = np.lexsort((np.arange(50), -want2.astype(np.float64)))[:5]
        assert np.array_equal(idx2[0], o2.astype(np.uint32)) and np.array_equal(sc2[0], want2[o2]) and idx2[0, 0] == 17


@pytest.mark.parametrize("L,q_count,db_count,rng_len,n_q", [(200, 6, 19, 0, 1), (200, 6, 19, 77, 4), (200, 1, 5, 0, 8), (100, 3, 7, 0, 3), (400, 2, 9, 300, 5)])
def test_few_query_kernel_regular_short_form(lb, checker, L, q_count, db_count, rng_len, n_q):
    """The lane-per-clip kernel on a database of regular codes: queries that are regular too take the short form (P planes only, possible =
    the range), decided per query — the last query here carries a '00' rank and runs the general form in the same launch.  Whole score
    matrix and top-k against the oracle, also with a shortened comparison range."""
    rng = np.random.default_rng(1300 + L + q_count + db_count + rng_len + n_q)
    n_db, k = 900, 10
    dbb = rank_sign_codes(rng, n_db, db_count, L)
    qb = rank_sign_codes(rng, n_q, q_count, L)
    for q in range(n_q):
        c = int(rng.integers(0, n_db)); o = int(rng.integers(0, db_count - q_count + 1))
        qb[q] = dbb[c, o:o + q_count]; flip = rng.random((q_count, L // 2)) < 0.04
        qb[q, :, 0::2] ^= flip.astype(np.uint8); qb[q, :, 1::2] ^= flip.astype(np.uint8)     # sign flips keep the codes regular
    if n_q > 1:
        qb[-1, 0, 6] = qb[-1, 0, 7] = 0
    db = lb.Database(L); db.add_packed(lb.pack_booleans(dbb))
    sc, idx, full = db.search_packed(lb.pack_booleans(qb), k, rng=rng_len, all_scores=True)
    want, _ = checker.search(dbb, qb, rng_len if rng_len else L)
    assert np.array_equal(full, want)
    order = np.lexsort((np.arange(n_db)[None, :].repeat(n_q, 0), -want.astype(np.float64)), axis=1)[:, :k]
    assert np.array_equal(idx, order.astype(np.uint32)) and np.array_equal(sc, np.take_along_axis(want, order, axis=1))


@pytest.mark.parametrize("L,q_count,db_count", [(200, 6, 19), (200, 1, 5), (100, 3, 7), (400, 2, 9)])
def test_search_regular_codes_short_form(lb, checker, L, q_count, db_count):
    """Databases in which every rank carries exactly one sign bit take the kernel's short form (one LOP3 per word, no M plane);
    warps whose queries do not all qualify fall back to the general form inside the same launch.  Both must be bit-exact."""
    rng = np.random.default_rng(300 + L + q_count)
    n_db, n_q, k = 500, 70, 7
    dbb = rank_sign_codes(rng, n_db, db_count, L)                          # regular: exactly one of (P, M) per rank
    qb = rank_sign_codes(rng, n_q, q_count, L)
    for q in range(0, n_q, 2):
        c = int(rng.integers(0, n_db)); o = int(rng.integers(0, db_count - q_count + 1))
        qb[q] = dbb[c, o:o + q_count]; flip = rng.random((q_count, L // 2)) < 0.05
        qb[q, :, 0::2] ^= flip.astype(np.uint8); qb[q, :, 1::2] ^= flip.astype(np.uint8)     # sign flips keep the codes regular
    qb[40:, 0, 6] = 0; qb[40:, 0, 7] = 0                                    # queries 40.. get a '00' rank: their warps take the general form
    qb[69, 0, 10] = 1; qb[69, 0, 11] = 1                                    # and one illegal '11'
    db = lb.Database(L); db.add_packed(lb.pack_booleans(dbb))
    sc, idx, full = db.search_packed(lb.pack_booleans(qb), k, all_scores=True)
    want, _ = checker.search(dbb, qb, L)
    assert np.array_equal(full, want)
    order = np.lexsort((np.arange(n_db)[None, :].repeat(n_q, 0), -want.astype(np.float64)), axis=1)[:, :k]
    assert np.array_equal(idx, order.astype(np.uint32)) and np.array_equal(sc, np.take_along_axis(want, order, axis=1))
    # an irregular database (one '00' rank somewhere) must give the same answers through the general form
    dbb2 = dbb.copy(); dbb2[123, 2, 0] = 0; dbb2[123, 2, 1] = 0
    db2 = lb.Database(L); db2.add_packed(lb.pack_booleans(dbb2))
    _, _, full2 = db2.search_packed(lb.pack_booleans(qb), k, all_scores=True)
    want2, _ = checker.search(dbb2, qb, L)
    assert np.array_equal(full2, want2)


@pytest.mark.parametrize("q_count,db_count,n_q", [(6, 19, 96), (4, 9, 40), (5, 5, 20), (6, 19, 33)])
def test_search_100_rank_form_and_mixed_databases(lb, checker, q_count, db_count, n_q):
    """L = 200 with regular queries in every warp of the CTA: the two-POPC form (tile words rewritten in place, leftover bits through the
    window words).  Then the same queries against a database in which a few clips carry empty or doubled ranks: regularity is decided
    per landed tile, so those tiles take the general form and all others keep the short one — every score must still equal the oracle's."""
    rng = np.random.default_rng(900 + q_count + db_count + n_q); L = 200      # (40 / 20 queries: the CTA's spare warps share each tile's clips)
    n_db, k = 1500, 10
    dbb = rank_sign_codes(rng, n_db, db_count, L)
    qb = rank_sign_codes(rng, n_q, q_count, L)
    for q in range(0, n_q, 2):
        c = int(rng.integers(0, n_db)); o = int(rng.integers(0, db_count - q_count + 1))
        qb[q] = dbb[c, o:o + q_count]; flip = rng.random((q_count, L // 2)) < 0.05
        qb[q, :, 0::2] ^= flip.astype(np.uint8); qb[q, :, 1::2] ^= flip.astype(np.uint8)     # sign flips keep the codes regular
    def check(bits):
        db = lb.Database(L); db.add_packed(lb.pack_booleans(bits))
        sc, idx, full = db.search_packed(lb.pack_booleans(qb), k, all_scores=True)
        want, _ = checker.search(bits, qb, L)
        assert np.array_equal(full, want)
        order = np.lexsort((np.arange(n_db)[None, :].repeat(n_q, 0), -want.astype(np.float64)), axis=1)[:, :k]
        assert np.array_equal(idx, order.astype(np.uint32)) and np.array_equal(sc, np.take_along_axis(want, order, axis=1))
    check(dbb)
    mixed = dbb.copy()
    for c in rng.choice(n_db, 12, replace=False):
        j = int(rng.integers(0, db_count)); r = int(rng.integers(0, L // 2))
        mixed[c, j, 2 * r] = mixed[c, j, 2 * r + 1] = int(rng.integers(0, 2))          # a '00' or a '11' rank (the last four ranks included)
    mixed[7, :, :] = 0                                                                  # digital silence: no rank carries a bit
    mixed[8, 0, 198] = mixed[8, 0, 199] = 0
    check(mixed)


def test_search_100_rank_form_on_a_ragged_database(lb, checker):
    """The two-POPC form with clips of different lengths (tile bounds and clip positions come from the offsets array instead of a uniform
    count), regular queries in every warp, clips of 6 to 30 subfingerprints — and the same database with a silent clip in it."""
    rng = np.random.default_rng(77); L = 200
    counts = rng.integers(6, 31, size=400)
    fps_bits = [rank_sign_codes(rng, 1, int(c), L)[0] for c in counts]
    qs = []
    for q in range(40):
        c = int(rng.integers(0, len(counts))); o = int(rng.integers(0, counts[c] - 6 + 1))
        b = fps_bits[c][o:o + 6].copy(); flip = rng.random((6, L // 2)) < 0.05
        b[:, 0::2] ^= flip.astype(np.uint8); b[:, 1::2] ^= flip.astype(np.uint8)
        qs.append(b)
    for silent in (False, True):
        if silent:
            fps_bits[123] = fps_bits[123].copy(); fps_bits[123][2, :] = 0
        db = lb.Database(L)
        for b in fps_bits:
            db.add_fingerprint(lb.Fingerprint.from_booleans(b))
        sc, idx = db.search([lb.Fingerprint.from_booleans(q) for q in qs], k=6)
        for qi, q in enumerate(qs):
            want = np.array([np.float32(checker.compare_fp(b, q, L)) for b in fps_bits])
            order = np.lexsort((np.arange(len(want)), -want.astype(np.float64)))[:6]
            assert np.array_equal(idx[qi], order.astype(np.uint32)) and np.array_equal(sc[qi], want[order]), (silent, qi)


def test_search_ragged_database_and_fingerprint_api(lb, checker):
    rng = np.random.default_rng(12); L = 200
    counts = rng.integers(0, 25, size=120)
    fps_bits = [rank_sign_codes(rng, 1, int(c), L)[0] for c in counts]
    db = lb.Database(L); db.set_clip_index_base(5000)
    for b in fps_bits:
        db.add_fingerprint(lb.Fingerprint.from_booleans(b) if len(b) else lb.Fingerprint(L))
    qs = [rank_sign_codes(rng, 1, 6, L)[0] for _ in range(9)]
    qs[0] = fps_bits[int(np.argmax(counts))][:6].copy()
    sc, idx = db.search([lb.Fingerprint.from_booleans(q) for q in qs], k=5)
    for qi, q in enumerate(qs):
        want = np.array([np.float32(checker.compare_fp(b, q, L)) if len(b) else np.float32(0) for b in fps_bits])
        order = np.lexsort((np.arange(len(want)), -want.astype(np.float64)))[:5]
        assert np.array_equal(idx[qi], (order + 5000).astype(np.uint32))
        assert np.array_equal(sc[qi], want[order])
    assert sc[0, 0] == 1.0 and idx[0, 0] == 5000 + int(np.argmax(counts))


def test_database_save_load_round_trip(lb, tmp_path):
    rng = np.random.default_rng(16); L = 200
    counts = rng.integers(0, 25, size=60)
    words = lb.pack_booleans(rank_sign_codes(rng, 1, int(counts.sum()), L)[0])
    db = lb.Database(L); db.add_packed(words, counts=counts)
    path = tmp_path / "birds.lbaddb"; db.save(path)
    assert os.path.getsize(path) == 8 + 16 + 8 + 4 * 60 + 32 * int(counts.sum())
    db2 = lb.Database.load(path, L)
    assert db2.clips == 60 and db2.subfingerprints == int(counts.sum())
    q = lb.pack_booleans(rank_sign_codes(rng, 5, 6, L))
    a = db.search_packed(q, k=4, all_scores=True); b = db2.search_packed(q, k=4, all_scores=True)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    open(path, "r+b").write(b"XXXX")
    assert lb.Database.load(path, L) is None and lb.Database.load(tmp_path / "missing", L) is None


def test_topk_with_fewer_clips_than_k(lb):
    rng = np.random.default_rng(13); db = lb.Database(200)
    db.add_packed(lb.pack_booleans(rank_sign_codes(rng, 3, 6, 200)))
    sc, idx = db.search_packed(lb.pack_booleans(rank_sign_codes(rng, 2, 6, 200)), k=8)
    assert (idx[:, 3:] == 0xFFFFFFFF).all() and (sc[:, 3:] == -1).all() and (idx[:, :3] < 3).all()


def test_few_queries_many_chunks_two_level_merge(lb):
    """One to three queries cut the database into thousands of chunks (to fill the device) whose lists are merged in two levels:
    the top-k must be the first k of the full score matrix ordered (score desc, clip asc), ties included."""
    rng = np.random.default_rng(21); n_db = 20000
    words = lb.pack_booleans(rank_sign_codes(rng, n_db, 7, 200))
    db = lb.Database(200); db.add_packed(words)
    for n_q in (1, 3):
        q = words[rng.integers(0, n_db, n_q), 1:5]
        sc, idx, full = db.search_packed(q, k=10, all_scores=True)
        o = np.lexsort((np.broadcast_to(np.arange(n_db), full.shape), -full.astype(np.float64)), axis=1)[:, :10]
        assert np.array_equal(idx, o.astype(np.uint32)) and np.array_equal(sc, np.take_along_axis(full, o, 1))
        assert (sc[:, 0] == 1.0).all()


def test_merge_topk_equals_single_list(lb):
    rng = np.random.default_rng(14); n_lists, n_q, k = 8, 50, 10
    sc = np.round(rng.random((n_lists, n_q, k)), 2).astype(np.float32)     # rounded: plenty of ties across lists
    idx = rng.permutation(n_lists * n_q * k).reshape(n_lists, n_q, k).astype(np.uint32)
    o = np.lexsort((idx, -sc.astype(np.float64)), axis=2)
    sc = np.take_along_axis(sc, o, 2); idx = np.take_along_axis(idx, o, 2)
    ms, mi = lb.merge_topk(sc, idx)
    alls = sc.transpose(1, 0, 2).reshape(n_q, -1); alli = idx.transpose(1, 0, 2).reshape(n_q, -1)
    o = np.lexsort((alli, -alls.astype(np.float64)), axis=1)[:, :k]
    assert np.array_equal(mi, np.take_along_axis(alli, o, 1)) and np.array_equal(ms, np.take_along_axis(alls, o, 1))


def test_extract_then_search_end_to_end(lb, port):
    """Queries cut from database clips (config 4 shape, small): the right clip comes first; noise keeps it first."""
    d = lb.Detective(); n_db = 40
    pcm = np.stack([port.synth_clip(200 + i, 165360) for i in range(n_db)])
    words = d.process_batch(pcm)
    db = lb.Database(200); db.add_packed(words)
    q_pcm = np.stack([pcm[i, 8192 * 3: 8192 * 3 + 55120] for i in range(0, n_db, 4)])
    noisy = np.stack([port.add_noise(q, 900 + j, 0.0158) for j, q in enumerate(q_pcm)])
    for q, floor in ((q_pcm, 1.0), (noisy, 0.6)):
        sc, idx = db.search_packed(d.process_batch(q), k=3)
        assert np.array_equal(idx[:, 0], np.arange(0, n_db, 4).astype(np.uint32))
        assert (sc[:, 0] >= floor).all() and (sc[:, 1] < sc[:, 0]).all()


def test_full_size_properties(lb):
    """BASELINE config 2 at full size — 10,000 x 30 s clips resident on the device — through size-independent properties:
    the same PCM gives the same words wherever it is scheduled, the pass is idempotent, every subfingerprint has the structure
    the reference's construction implies, and excerpts of database clips find their clip."""
    import torch
    d = lb.Detective(); n, clip_len = 10000, 165360
    x = torch.empty((n, clip_len), dtype=torch.float32, device="cuda")
    lb.synthesize_device(x.data_ptr(), n // 2, clip_len, clip_len, first_clip_id=0)
    x[n // 2:] = x[:n // 2]                                                 # second half duplicates the first
    out = torch.zeros((n, 19, 8), dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    d.process_batch_device(x.data_ptr(), n, clip_len, clip_len, out.data_ptr(), s)
    torch.cuda.synchronize()
    w = out.cpu().numpy().view(np.uint32)
    assert np.array_equal(w[:n // 2], w[n // 2:])                           # same PCM -> same words wherever it is scheduled
    out2 = torch.zeros_like(out); d.process_batch_device(x.data_ptr(), n, clip_len, clip_len, out2.data_ptr(), s); torch.cuda.synchronize()
    assert torch.equal(out, out2)                                           # idempotent
    assert ((w[..., :4] & w[..., 4:]) == 0).all()                           # P and M planes never overlap
    pop = np.unpackbits((w[..., :4] | w[..., 4:]).view(np.uint8), axis=-1).sum(axis=-1)
    assert (pop == 100).all()                                               # exactly t_eff = 100 ranks carry a sign
    assert (w[..., 3] >> 4 == 0).all() and (w[..., 7] >> 4 == 0).all()      # nothing beyond pair 99
    db = lb.Database(200); db.add_packed_device(out.data_ptr(), n // 2, 19)
    sc, idx = db.search_packed(w[5:n // 2:97, 4:10], k=1)                   # 6-subfp excerpts of database clips
    assert np.array_equal(idx[:, 0], np.arange(5, n // 2, 97).astype(np.uint32)) and (sc == 1.0).all()


def test_full_size_search_sharded_equals_whole(lb):
    """BASELINE config 4 at full size — 1,000 six-subfingerprint queries against 1,000,000 clips of 19 — through properties that do
    not need the oracle: every query finds the clip it was cut from with score exactly 1, and the database split into eight shards
    (the 8-GPU layout: per-shard top-k with global clip ids, then the merge kernel) returns what the whole database returns, bit for bit."""
    import torch
    n_db, n_q, c, k = 1000000, 1000, 19, 10
    codes = torch.empty((n_db, c, 8), dtype=torch.int32, device="cuda")
    lb.random_codes_device(codes.data_ptr(), n_db * c, 200, seed=4242); torch.cuda.synchronize()
    g = torch.Generator().manual_seed(9)
    src = torch.randint(0, n_db, (n_q,), generator=g); off = torch.randint(0, c - 6 + 1, (n_q,), generator=g)
    q = torch.stack([codes[int(a), int(o):int(o) + 6] for a, o in zip(src, off)]).contiguous()
    whole = lb.Database(200); whole.add_packed_device(codes.data_ptr(), n_db, c)
    sc = torch.empty((n_q, k), dtype=torch.float32, device="cuda"); ix = torch.empty((n_q, k), dtype=torch.int32, device="cuda")
    whole.search_device(q.data_ptr(), n_q, 6, k, sc.data_ptr(), ix.data_ptr()); torch.cuda.synchronize()
    sc_w, ix_w = sc.cpu().numpy(), ix.cpu().numpy().view(np.uint32)
    assert np.array_equal(ix_w[:, 0], src.numpy().astype(np.uint32)) and (sc_w[:, 0] == 1.0).all()
    assert (np.diff(sc_w, axis=1) <= 0).all() and (sc_w[:, 1] < 1.0).all()
    del whole
    shards = 8; per = n_db // shards
    s_sc = torch.empty((shards, n_q, k), dtype=torch.float32, device="cuda"); s_ix = torch.empty((shards, n_q, k), dtype=torch.int32, device="cuda")
    for r in range(shards):
        db = lb.Database(200); db.set_clip_index_base(r * per)
        db.add_packed_device(codes[r * per:(r + 1) * per].data_ptr(), per, c)
        db.search_device(q.data_ptr(), n_q, 6, k, s_sc[r].data_ptr(), s_ix[r].data_ptr()); torch.cuda.synchronize()
        del db
    m_sc = torch.empty((n_q, k), dtype=torch.float32, device="cuda"); m_ix = torch.empty((n_q, k), dtype=torch.int32, device="cuda")
    lb.merge_topk_device(s_sc.data_ptr(), s_ix.data_ptr(), shards, n_q, k, m_sc.data_ptr(), m_ix.data_ptr()); torch.cuda.synchronize()
    assert np.array_equal(m_ix.cpu().numpy().view(np.uint32), ix_w) and np.array_equal(m_sc.cpu().numpy(), sc_w)


# ------------------------------------------------------- multi-GPU group, stream ordering ----

@pytest.mark.parametrize("n_shards", [1, 3, 8])
def test_database_group_equals_single_database(lb, checker, n_shards):
    """LBAudioDetectiveDatabaseGroup: a database sharded over the GPUs of one process returns what ONE database returns, bit for bit —
    ragged clips, several appends (so shards hold several runs of global indices), clips added one by one, few and many queries,
    a shortened range; scores are held against the oracle as well."""
    rng = np.random.default_rng(40 + n_shards); L = 200
    n_dev = lb.device_count()
    group = lb.DatabaseGroup(L, [i % n_dev for i in range(n_shards)])
    whole = lb.Database(L)
    all_bits = []
    for n_clips in (700, 1, 333):
        counts = rng.integers(6, 25, size=n_clips)
        bits = [rank_sign_codes(rng, 1, int(c), L)[0] for c in counts]
        words = np.concatenate([lb.pack_booleans(b) for b in bits])
        first = group.add_packed(words, counts=counts); whole.add_packed(words, counts=counts)
        assert first == len(all_bits)
        all_bits += bits
    for _ in range(5):                                                       # one by one, through fingerprint objects
        b = rank_sign_codes(rng, 1, 9, L)[0]
        assert group.add_fingerprint(lb.Fingerprint.from_booleans(b)) == len(all_bits)
        whole.add_fingerprint(lb.Fingerprint.from_booleans(b)); all_bits.append(b)
    assert group.clips == whole.clips == len(all_bits) and group.shards == n_shards
    assert sum(group.shard_clips(i) for i in range(n_shards)) == len(all_bits)
    for n_q, rg in ((40, 0), (3, 0), (1, 77)):
        src = rng.integers(0, len(all_bits), n_q)
        qb = np.stack([all_bits[c][0:6] for c in src]).copy()
        qb[::2, 0, 0:2] ^= 1                                                 # every other query: one rank flipped
        q = lb.pack_booleans(qb)
        g_sc, g_id = group.search_packed(q, 10, rng=rg)
        w_sc, w_id = whole.search_packed(q, 10, rng=rg)
        assert np.array_equal(g_sc, w_sc) and np.array_equal(g_id, w_id)
        for qi in range(min(n_q, 3)):                                        # and against the oracle
            want = np.array([np.float32(checker.compare_fp(b, qb[qi], rg if rg else L)) for b in all_bits])
            order = np.lexsort((np.arange(len(want)), -want.astype(np.float64)))[:10]
            assert np.array_equal(g_id[qi], order.astype(np.uint32)) and np.array_equal(g_sc[qi], want[order])
    assert group.kernel_launches > 0 and group.last_search_ms > 0


def test_database_group_device_appends_with_global_ids(lb):
    """Shards filled on their devices with explicit global clip indices (what bench.py does at full size): same top-k as one database."""
    import torch
    n_dev = lb.device_count(); shards = 4; per = 5000; c = 19
    group = lb.DatabaseGroup(200, [i % n_dev for i in range(shards)])
    whole = lb.Database(200)
    for s in range(shards):
        with torch.cuda.device(group.shard_device(s)):
            codes = torch.empty((per, c, 8), dtype=torch.int32, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            lb.random_codes_device(codes.data_ptr(), per * c, 200, seed=99, stream=st, first_subfp=s * per * c)
            group.add_packed_device_to_shard(s, codes.data_ptr(), per, c, s * per, producer_stream=st)     # no synchronisation in between: ordered by the library
            host = codes.cpu().numpy().view(np.uint32)
        whole.add_packed(host)
    allc = torch.empty((shards * per, c, 8), dtype=torch.int32, device="cuda")
    lb.random_codes_device(allc.data_ptr(), shards * per * c, 200, seed=99); torch.cuda.synchronize()
    q = allc[::997, 3:9].contiguous().cpu().numpy().view(np.uint32)
    g = group.search_packed(q, 10); w = whole.search_packed(q, 10)
    assert np.array_equal(g[0], w[0]) and np.array_equal(g[1], w[1])
    assert np.array_equal(g[1][:, 0], np.arange(0, shards * per, 997).astype(np.uint32)) and (g[0][:, 0] == 1.0).all()


def test_calls_on_different_streams_are_ordered(lb, port):
    """One detective (one database) used from several CUDA streams without synchronising in between: the calls share device scratch
    (spectral images, partial top-k lists), so the library orders them; every result must equal the single-stream one."""
    import torch
    d = lb.Detective(); n, clip_len = 96, 165360
    xs = [torch.empty((n, clip_len), dtype=torch.float32, device="cuda") for _ in range(3)]
    for i, x in enumerate(xs):
        lb.synthesize_device(x.data_ptr(), n, clip_len, clip_len, first_clip_id=3000 + 1000 * i)
    torch.cuda.synchronize()
    want = []
    for x in xs:
        o = torch.zeros((n, 19, 8), dtype=torch.int32, device="cuda")
        d.process_batch_device(x.data_ptr(), n, clip_len, clip_len, o.data_ptr(), torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
        want.append(o.cpu().numpy())
    streams = [torch.cuda.Stream() for _ in range(3)]
    for rep in range(3):
        outs = [torch.zeros((n, 19, 8), dtype=torch.int32, device="cuda") for _ in range(3)]
        torch.cuda.synchronize()
        for x, o, s in zip(xs, outs, streams):
            d.process_batch_device(x.data_ptr(), n, clip_len, clip_len, o.data_ptr(), s.cuda_stream)
        torch.cuda.synchronize()
        for o, w in zip(outs, want):
            assert np.array_equal(o.cpu().numpy(), w)
    # extraction on one stream, database append ordered after it, searches from three streams
    db = lb.Database(200)
    o = torch.zeros((n, 19, 8), dtype=torch.int32, device="cuda"); torch.cuda.synchronize()
    d.process_batch_device(xs[0].data_ptr(), n, clip_len, clip_len, o.data_ptr(), streams[0].cuda_stream)
    db.add_packed_device(o.data_ptr(), n, 19, producer_stream=streams[0].cuda_stream)
    q = torch.from_numpy(want[0][::5, 2:8].copy()).cuda()
    res = [(torch.empty((q.shape[0], 5), dtype=torch.float32, device="cuda"), torch.empty((q.shape[0], 5), dtype=torch.int32, device="cuda")) for _ in range(3)]
    torch.cuda.synchronize()
    for (sc, ix), s in zip(res, streams):
        db.search_device(q.data_ptr(), q.shape[0], 6, 5, sc.data_ptr(), ix.data_ptr(), stream=s.cuda_stream)
    torch.cuda.synchronize()
    for sc, ix in res:
        assert np.array_equal(ix[:, 0].cpu().numpy(), np.arange(0, n, 5).astype(np.int32)) and (sc[:, 0] == 1.0).all()
        assert torch.equal(sc, res[0][0]) and torch.equal(ix, res[0][1])


@pytest.mark.parametrize("L,q_count,db_count,n_q,silent", [(200, 1, 5, 100, 0), (100, 2, 6, 40, 0), (200, 6, 19, 70, 0), (200, 6, 19, 70, 37), (200, 4, 5, 33, 501)])
def test_threshold_pass_changes_nothing(lb, checker, monkeypatch, L, q_count, db_count, n_q, silent):
    """(silent > 0: the same with empty-rank subfingerprints sprinkled over the database.)
    Databases of 65,536 clips and more are searched in two passes — the top k of a sample first, whose k-th score then keeps clips
    below it out of the per-chunk lists.  The result must be the one of the single pass, ties included (short codes and noisy excerpts
    make plenty of equal scores), and equal the oracle's on the queries checked."""
    rng = np.random.default_rng(900 + L + q_count)
    n_db, k = 70000, 10
    dbb = rank_sign_codes(rng, n_db, db_count, L)
    dbb[1000:1200] = dbb[0:200]                                             # exact duplicates: equal scores, the lower clip index must win
    if silent:                                                              # every silent-th clip loses the bits of one subfingerprint: a mixed database,
        dbb[2000::silent, db_count // 2, :] = 0                             # regular tiles keep the short compare forms, the others take the general one
    src = rng.integers(0, n_db, n_q)
    qb = np.stack([dbb[c, 1:1 + q_count] for c in src]).copy()
    flip = rng.random(qb.shape[:2] + (L // 2,)) < 0.1
    qb[..., 0::2] ^= flip.astype(np.uint8); qb[..., 1::2] ^= flip.astype(np.uint8)
    qb[0] = dbb[50, 1:1 + q_count]                                          # clip 50 and its duplicate 1050 both score 1
    words = lb.pack_booleans(dbb); q = lb.pack_booleans(qb)
    db = lb.Database(L); db.add_packed(words)
    two = db.search_packed(q, k)
    monkeypatch.setenv("LBAD_SEARCH_NO_FLOOR", "1")
    one = db.search_packed(q, k)
    monkeypatch.delenv("LBAD_SEARCH_NO_FLOOR")
    assert np.array_equal(two[0], one[0]) and np.array_equal(two[1], one[1])
    assert two[1][0, 0] == 50 and two[1][0, 1] == 1050 and two[0][0, 0] == 1.0 and two[0][0, 1] == 1.0
    for qi in (0, 1, n_q - 1):
        want = np.array([np.float32(checker.compare_fp(dbb[c], qb[qi], L)) for c in range(0, n_db, 1)][:4000] , np.float32)      # the oracle on the first 4,000 clips
        sub = lb.Database(L); sub.add_packed(words[:4000])
        sc, ix = sub.search_packed(q[qi:qi + 1], k)
        order = np.lexsort((np.arange(4000), -want.astype(np.float64)))[:k]
        assert np.array_equal(ix[0], order.astype(np.uint32)) and np.array_equal(sc[0], want[order])
