"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/*.h declares, and its host logic
(containers, codecs, configuration rules, error behaviour) matches the reference.  No kernels are launched."""
import glob
import os
import re
import subprocess
import ctypes as C

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = open(h).read()
        names += re.findall(r"LBAD_API\s+(?:extern\s+)?[\w\s\*]+?\b(k?LBAudioDetective\w+)\s*(?:\(|;)", txt)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(lb):
    syms = declared_symbols()
    assert len(syms) > 60
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "lbaudiodetective_b200", "libLBAudioDetectiveCUDA.so")],
                         stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = set(re.findall(r"\s[TDRB]\s+(\w+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, missing
    # and nothing from the oracle is linked into the product
    assert not [s for s in exported if "oracle" in s.lower() or s.startswith("lbad_ref")]


def test_python_mirror_binds_every_function(lb):
    bound = set(lb.lib()._lbad_signatures)
    funcs = [s for s in declared_symbols() if not s.startswith("k")]
    missing = [f for f in funcs if f not in bound and f != "LBAudioDetectiveDefaultProcessingFormat"]
    assert not missing, missing


def test_constants(lb):
    L = lb.lib()
    get = lambda n, t: t.in_dll(L, n).value
    assert get("kLBAudioDetectiveArgumentInvalid", C.c_int32) == 1
    assert get("kLBAudioDetectiveDefaultWindowSize", C.c_uint32) == 2048
    assert get("kLBAudioDetectiveDefaultAnalysisStride", C.c_uint32) == 64
    assert get("kLBAudioDetectiveDefaultNumberOfPitchSteps", C.c_uint32) == 32
    assert get("kLBAudioDetectiveDefaultNumberOfRowsPerFrame", C.c_uint32) == 128
    assert get("kLBAudioDetectiveDefaultSubfingerprintLength", C.c_uint32) == 200
    assert get("kLBAudioDetectiveDefaultFingerprintComparisonRange", C.c_uint32) == 0


def test_default_processing_format(lb):
    class ASBD(C.Structure):
        _fields_ = [("mSampleRate", C.c_double), ("mFormatID", C.c_uint32), ("mFormatFlags", C.c_uint32), ("mBytesPerPacket", C.c_uint32),
                    ("mFramesPerPacket", C.c_uint32), ("mBytesPerFrame", C.c_uint32), ("mChannelsPerFrame", C.c_uint32),
                    ("mBitsPerChannel", C.c_uint32), ("mReserved", C.c_uint32)]
    f = lb.lib().LBAudioDetectiveDefaultProcessingFormat; f.restype = ASBD; f.argtypes = []
    a = f()
    assert C.sizeof(ASBD) == 40
    assert (a.mSampleRate, a.mFormatID, a.mFormatFlags, a.mBytesPerPacket, a.mFramesPerPacket, a.mBytesPerFrame, a.mChannelsPerFrame, a.mBitsPerChannel) == \
        (5512.0, 0x6c70636d, 1 | 8, 4, 1, 4, 1, 32)                       # LBAudioDetective.m:116-131


def test_detective_defaults_and_setters(lb, kat):
    d = lb.Detective()
    assert (d.window_size, d.analysis_stride, d.pitch_steps, d.subfingerprint_length, d.sample_rate) == (2048, 64, 32, 200, 5512.0)
    for n, st in kat["set_window_size_status"].items():                   # Q13: inverted power-of-two check, value still applied
        assert d.set_window_size(int(n)) == st and d.window_size == int(n)
    assert d.set_analysis_stride(32) == 0 and d.analysis_stride == 32
    assert d.set_pitch_steps(16) == 0 and d.pitch_steps == 16
    assert d.set_subfingerprint_length(100) == 0 and d.subfingerprint_length == 100
    assert d.set_sample_rate(8000.0) == 0 and d.sample_rate == 8000.0
    assert lb.lib().LBAudioDetectiveSetRecordingSampleRate(d.ref, 44100.0) == 0
    assert d.dispose() == 0
    assert lb.lib().LBAudioDetectiveDispose(None) == lb.ARGUMENT_INVALID   # LBAudioDetective.m:93-95


def test_band_table_matches_reference(lb, kat, port):
    from oracle.oracle import Cfg
    d = lb.Detective()
    for n, t in kat["band_tables"].items():
        d.set_window_size(int(n))
        idx, lo, hi = d.band_table()
        assert idx.tolist() == t["indices"] and lo.tolist() == t["klow"] and hi.tolist() == t["khigh"]
    d.set_window_size(2048); d.set_sample_rate(8000.0); d.set_pitch_steps(16)
    idx, lo, hi = d.band_table()
    pi, pl, ph = port.band_table(Cfg.default(sample_rate=8000.0, bands=16))
    assert np.array_equal(idx, pi) and np.array_equal(lo, pl) and np.array_equal(hi, ph)


def test_configuration_rules(lb):
    d = lb.Detective()
    assert d.check_configuration() == 0
    for bad in (4096, 1000, 128, 0):                                       # Q15: 4096 reads out of bounds upstream
        d.set_window_size(bad); assert d.check_configuration() == lb.ARGUMENT_INVALID
    d.set_window_size(2048)
    for bad in (0, 3, 201, 514):
        d.set_subfingerprint_length(bad); assert d.check_configuration() == lb.ARGUMENT_INVALID
    d.set_subfingerprint_length(200)
    d.set_analysis_stride(0); assert d.check_configuration() == lb.ARGUMENT_INVALID
    d.set_analysis_stride(64)
    d.set_pitch_steps(33); assert d.check_configuration() == lb.ARGUMENT_INVALID
    d.set_pitch_steps(32)
    d.set_sample_rate(500.0); assert d.check_configuration() == lb.ARGUMENT_INVALID     # Nyquist below the 318 Hz floor
    d.set_sample_rate(5512.0); assert d.check_configuration() == 0


def test_subfingerprint_counts(lb, kat):
    d = lb.Detective()
    for n, c in kat["subfp_counts"].items():
        assert d.subfingerprints_for_length(int(n)) == c
    assert d.subfingerprints_for_length(2047) == 0 and d.subfingerprints_for_length(10239) == 0 and d.subfingerprints_for_length(10240) == 1
    assert d.subfingerprints_for_length(19843200) == 2422                 # one hour


def test_fingerprint_container(lb):
    rng = np.random.default_rng(0)
    bits = (rng.random((5, 200)) < 0.5).astype(np.uint8)
    fp = lb.Fingerprint(0)
    assert fp.set_subfingerprint_length(200) == (True, 200)
    for row in bits:
        fp.add_subfingerprint(row)
    assert fp.set_subfingerprint_length(100) == (False, 200)              # frozen after the first add (FP.m:81-89)
    assert fp.count == 5 and fp.subfingerprint_length == 200
    assert np.array_equal(fp.booleans(), bits)
    cp = fp.copy()
    assert cp.equal(fp) and fp.equal(cp)                                   # testFingerprintComparison (Tests.m:141-155)
    other = lb.Fingerprint.from_booleans(np.vstack([bits[:4], 1 - bits[4:]]))
    assert not other.equal(fp)
    assert not lb.Fingerprint.from_booleans(bits[:4]).equal(fp)
    lb.lib().LBAudioDetectiveFingerprintDispose(None)                      # no-op (FP.m:29-31)


def test_frame_container_matches_reference(lb, ref):
    """The host side of the Frame API (include/LBAudioDetectiveFrame.h) against the compiled reference's own functions
    (LBAudioDetectiveFrame.m:22-105, :155-161, :193-210), call by call."""
    R = ref.ref
    for name, res in (("LBAudioDetectiveFrameNew", C.c_void_p), ("LBAudioDetectiveFrameCopy", C.c_void_p), ("LBAudioDetectiveFrameGetNumberOfRows", C.c_uint32),
                      ("LBAudioDetectiveFrameGetValue", C.c_float), ("LBAudioDetectiveFrameFull", C.c_ubyte), ("LBAudioDetectiveFrameSetRow", C.c_ubyte),
                      ("LBAudioDetectiveFrameFingerprintSize", C.c_size_t), ("LBAudioDetectiveFrameFingerprintLength", C.c_uint32),
                      ("LBAudioDetectiveFrameEqualToFrame", C.c_ubyte), ("LBAudioDetectiveFrameDispose", None)):
        getattr(R, name).restype = res
    vp = C.c_void_p
    rng = np.random.default_rng(5)
    rows = [rng.standard_normal(n).astype(np.float32) for n in (7, 5, 9, 5)]
    mine = lb.Frame(3); theirs = vp(R.LBAudioDetectiveFrameNew(C.c_uint32(3)))
    for i, row in enumerate(rows):                                            # the fourth SetRow meets a full frame on both sides
        want = R.LBAudioDetectiveFrameSetRow(theirs, row.ctypes.data_as(vp), C.c_uint32(min(i, 2)), C.c_uint32(len(row))) if i < 3 else \
            R.LBAudioDetectiveFrameSetRow(theirs, row.ctypes.data_as(vp), C.c_uint32(2), C.c_uint32(len(row)))
        got = mine.set_row(row, min(i, 2))
        assert bool(want) == got
        assert mine.rows == R.LBAudioDetectiveFrameGetNumberOfRows(theirs)
        assert mine.full() == bool(R.LBAudioDetectiveFrameFull(theirs))
        assert mine.fingerprint_length == R.LBAudioDetectiveFrameFingerprintLength(theirs)
        assert mine.fingerprint_size == R.LBAudioDetectiveFrameFingerprintSize(theirs)
    assert mine.rows == 3 and mine.row_length == 5                             # the shortest row sets the row length (Frame.m:96-101)
    for r in range(3):
        for c in range(5):
            assert mine.value(r, c) == R.LBAudioDetectiveFrameGetValue(theirs, C.c_uint32(r), C.c_uint32(c))
    cp = mine.copy(); tcp = vp(R.LBAudioDetectiveFrameCopy(theirs))
    assert cp.equal(mine) and mine.equal(cp) and bool(R.LBAudioDetectiveFrameEqualToFrame(theirs, tcp))
    other = lb.Frame.from_array(np.stack([r[:5] for r in rows[:3]]) + np.float32(1))
    assert not other.equal(mine) and not lb.Frame(3).equal(mine)
    assert np.array_equal(cp.array(), np.stack([r[:5] for r in rows[:3]]))
    R.LBAudioDetectiveFrameDispose(theirs); R.LBAudioDetectiveFrameDispose(tcp)
    lb.lib().LBAudioDetectiveFrameDispose(None)                                # no-op, as upstream (Frame.m:34-36)
    empty = lb.Frame(4)
    assert empty.rows == 0 and not empty.full() and empty.fingerprint_length == 0 and lb.Frame(0).full()


def test_packed_layout_round_trip(lb):
    rng = np.random.default_rng(1)
    for L, W in ((100, 2), (128, 2), (130, 4), (200, 4), (256, 4), (400, 8), (512, 8)):
        assert lb.words_per_plane(L) == W
        bits = (rng.random((3, L)) < 0.5).astype(np.uint8)
        fp = lb.Fingerprint.from_booleans(bits)
        words = fp.packed()
        assert np.array_equal(words, lb.pack_booleans(bits))
        assert np.array_equal(lb.unpack_words(words, L), bits)
        f2 = lb.Fingerprint(L); f2.add_packed(words)
        assert f2.equal(fp)
        # P plane bit b of word w = Boolean[2*(32w+b)], M plane = Boolean[2*(32w+b)+1]
        assert ((words[0, 0] >> 5) & 1) == bits[0, 10] and ((words[0, W] >> 5) & 1) == bits[0, 11]
    assert lb.words_per_plane(0) == 0 and lb.words_per_plane(514) == 0


def test_string_codec(lb):
    rng = np.random.default_rng(2)
    bits = (rng.random((3, 8)) < 0.5).astype(np.uint8)
    fp = lb.Fingerprint.from_booleans(bits)
    s = fp.to_string()
    assert s == "+".join("".join(str(int(b)) for b in row) for row in bits)      # Tests.m:22-37
    assert lb.Fingerprint.from_string(s).equal(fp)
    assert lb.Fingerprint.from_string("0101+01") is None and lb.Fingerprint.from_string("01x1") is None
    assert lb.Fingerprint.from_string("").count == 0


def test_markstein_ratio_is_exact():
    """hits/possible via multiply + 2 FMAs against RN(1/possible) equals the IEEE quotient for every case the kernel can see."""
    p = np.arange(1, 257, dtype=np.float32)[:, None]; h = np.arange(0, 257, dtype=np.float32)[None, :]
    r = (np.float32(1.0) / p).astype(np.float32)
    q0 = (h * r).astype(np.float32)
    rem = (h.astype(np.float64) - q0.astype(np.float64) * p.astype(np.float64)).astype(np.float32)       # exact in f64 = fma
    q = (q0.astype(np.float64) + rem.astype(np.float64) * r.astype(np.float64)).astype(np.float32)
    ok = h <= p
    assert np.array_equal(q[ok], (h / p).astype(np.float32)[np.broadcast_to(ok, q.shape)])


def test_no_gpu_means_loud_failure(lb):
    if lb.device_available():
        pytest.skip("a CUDA device is present")
    d = lb.Detective()
    st, fp = d.process_pcm(np.zeros(55120, np.float32), check=False)
    assert st == lb.DEVICE_UNAVAILABLE and fp is None
    with pytest.raises(lb.LBADError):
        lb.Database(200)
    with pytest.raises(lb.LBADError):
        lb.microbench()
    with pytest.raises(lb.LBADError) as err:                                  # the compare-audio pipeline has no CPU route either
        d.compare_pcm(np.zeros(55120, np.float32), np.zeros(55120, np.float32), 0)
    assert "-7001" in str(err.value)
    with pytest.raises(lb.LBADError):
        d.resample(np.zeros(44100, np.float32))
    # the value-returning compare functions have no status: NaN (never a score computed some other way), the process lives on
    a = lb.Fingerprint.from_booleans(np.ones((2, 200), np.uint8)); b = lb.Fingerprint.from_booleans(np.ones((2, 200), np.uint8))
    assert np.isnan(a.compare(b, 200))
    assert np.isnan(a.compare_subfingerprints(np.ones(200, np.uint8), np.ones(200, np.uint8), 200))
    st, _ = a.compare_status(b, 200)
    assert st == lb.DEVICE_UNAVAILABLE
    with pytest.raises(lb.LBADError):
        lb.DatabaseGroup(200, [0, 0])
    # device selection and the sharded batch: no device to choose, nothing computed
    assert d.device == -1 and d.set_device(-1) == 0 and d.set_device(0) == lb.DEVICE_UNAVAILABLE and d.set_device(-2) == lb.ARGUMENT_INVALID
    with pytest.raises(lb.LBADError) as err2:
        lb.Detective.process_batch_sharded([lb.Detective(), lb.Detective()], np.zeros((3, 55120), np.float32))
    assert "-7001" in str(err2.value)
    # the Frame API's two computing functions: status -7001, the frame and the output stay as they were
    img = np.arange(12, dtype=np.float32).reshape(3, 4); f = lb.Frame.from_array(img)
    assert f.decompose(status=True) == lb.DEVICE_UNAVAILABLE and np.array_equal(f.array(), img)
    f.decompose(); assert np.array_equal(f.array(), img)
    st, bits = f.extract_fingerprint(3, status=True)
    assert st == lb.DEVICE_UNAVAILABLE and not bits.any() and not f.extract_fingerprint(3).any()
