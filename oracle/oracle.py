"""TEST INFRASTRUCTURE — ctypes bindings for the two CPU checkers. Not part of the product.

* ``Port``  — oracle/liblbad_oracle.so, the plain-C restatement (oracle/lbad_oracle.c), always buildable.
* ``Ref``   — oracle/_ref/libLBAudioDetectiveRef.so, the reference's own .m files compiled behind the disclosed
              shim (oracle/shim/); only buildable where /root/reference exists, but the built .so travels.

Both expose the same Python surface so tests can run either against the CUDA path.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liblbad_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libLBAudioDetectiveRef.so")
REFERENCE_SRC = "/root/reference/LBAudioDetective"
ROWS_PER_FRAME = 128
BASE_SEED = 0x1BAD5EED


def build(ref: bool = True) -> None:
    """Compile the checkers (gcc). The compiled reference is only rebuilt when its sources are present."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir(REFERENCE_SRC):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


class Cfg(C.Structure):
    _fields_ = [("window", C.c_uint32), ("stride", C.c_uint32), ("bands", C.c_uint32), ("sublen", C.c_uint32),
                ("sample_rate", C.c_double)]

    @staticmethod
    def default(**kw):
        c = Cfg(2048, 64, 32, 200, 5512.0)
        for k, v in kw.items():
            setattr(c, k, v)
        return c


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def subfp_count(cfg: Cfg, n: int) -> int:
    if n < cfg.window:
        return 0
    return ((n - cfg.window) // cfg.stride) // ROWS_PER_FRAME


class Port:
    """The C restatement."""
    kind = "port"

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        L = self.lib = C.CDLL(PORT_SO)
        L.lbad_oracle_compare_sub.restype = C.c_float
        L.lbad_oracle_compare_fp.restype = C.c_float
        L.lbad_oracle_extract_batch.restype = C.c_double
        L.lbad_oracle_search.restype = C.c_double
        L.lbad_oracle_subfp_count.restype = C.c_uint64
        L.lbad_oracle_resample.restype = C.c_uint64
        L.lbad_oracle_resampled_length.restype = C.c_uint64

    def resample(self, x, in_rate, out_rate=5512.0):
        """include/LBAudioDetectiveResample.h, scalar restatement (the reference leaves this step to ExtAudioFile)."""
        x = _f32(x); n = int(self.lib.lbad_oracle_resampled_length(C.c_double(in_rate), C.c_double(out_rate), C.c_uint64(len(x))))
        out = np.zeros(n, np.float32)
        got = self.lib.lbad_oracle_resample(C.c_double(in_rate), C.c_double(out_rate), _p(x, C.c_float), C.c_int64(len(x)), _p(out, C.c_float))
        assert got == n
        return out

    def band_table(self, cfg, nframes=None):
        B = cfg.bands
        idx = np.zeros(B + 1, np.uint32); lo = np.zeros(B, np.uint32); hi = np.zeros(B, np.uint32)
        self.lib.lbad_oracle_band_table(C.byref(cfg), C.c_uint32(nframes or cfg.window), _p(idx, C.c_uint32), _p(lo, C.c_uint32), _p(hi, C.c_uint32))
        return idx, lo, hi

    def fft2x(self, x):
        x = _f32(x); out = np.zeros(len(x), np.float32)
        self.lib.lbad_oracle_fft2x(_p(x, C.c_float), C.c_uint32(len(x)), _p(out, C.c_float))
        return out

    def band_energies(self, cfg, pcm, n_windows):
        pcm = _f32(pcm); out = np.zeros((n_windows, cfg.bands), np.float32)
        for w in range(n_windows):
            win = np.ascontiguousarray(pcm[cfg.stride * w: cfg.stride * w + cfg.window])
            self.lib.lbad_oracle_window_bands(C.byref(cfg), _p(win, C.c_float), _p(out[w], C.c_float))
        return out

    def haar(self, image):
        img = _f32(image).copy()
        self.lib.lbad_oracle_haar(_p(img, C.c_float), C.c_uint32(img.shape[0]), C.c_uint32(img.shape[1]))
        return img

    def extract_bits(self, coeffs, t):
        c = _f32(coeffs).reshape(-1); out = np.zeros(2 * t, np.uint8)
        self.lib.lbad_oracle_extract_bits(_p(c, C.c_float), C.c_uint32(c.size), C.c_uint32(t), _p(out, C.c_uint8))
        return out

    def process(self, cfg, pcm, stages=False):
        pcm = _f32(pcm); n = subfp_count(cfg, len(pcm))
        bits = np.zeros((max(n, 1), cfg.sublen), np.uint8); cnt = C.c_uint32(0)
        img = np.zeros((max(n, 1), ROWS_PER_FRAME, cfg.bands), np.float32) if stages else None
        haar = np.zeros_like(img) if stages else None
        self.lib.lbad_oracle_process(C.byref(cfg), _p(pcm, C.c_float), C.c_int64(len(pcm)), _p(bits, C.c_uint8), C.c_uint32(n), C.byref(cnt),
                                     _p(img, C.c_float) if stages else None, _p(haar, C.c_float) if stages else None)
        assert cnt.value == n
        return (bits[:n], img[:n], haar[:n]) if stages else bits[:n]

    def compare_sub(self, s1, s2, length, rng):
        s1 = _u8(np.concatenate([s1, [0, 0]])); s2 = _u8(np.concatenate([s2, [0, 0]]))
        return float(self.lib.lbad_oracle_compare_sub(_p(s1, C.c_uint8), _p(s2, C.c_uint8), C.c_uint32(length), C.c_uint32(rng)))

    def compare_fp(self, b1, b2, rng):
        b1 = _u8(b1); b2 = _u8(b2)
        l1 = b1.shape[1] if b1.ndim == 2 else 0; l2 = b2.shape[1] if b2.ndim == 2 else 0
        return float(self.lib.lbad_oracle_compare_fp(_p(b1, C.c_uint8), C.c_uint32(b1.shape[0]), C.c_uint32(l1),
                                                     _p(b2, C.c_uint8), C.c_uint32(b2.shape[0]), C.c_uint32(l2), C.c_uint32(rng)))

    def compare_pcm(self, cfg, p1, p2, rng=0):
        p1 = _f32(p1); p2 = _f32(p2); out = C.c_float(-1)
        self.lib.lbad_oracle_compare_pcm(C.byref(cfg), _p(p1, C.c_float), C.c_int64(len(p1)), _p(p2, C.c_float), C.c_int64(len(p2)), C.c_uint32(rng), C.byref(out))
        return float(out.value)

    def extract_batch(self, cfg, pcm2d, threads=1, want_bits=True, fft_f32=False):
        pcm2d = _f32(pcm2d); n_clips, clip_len = pcm2d.shape; n = subfp_count(cfg, clip_len)
        bits = np.zeros((n_clips, max(n, 1), cfg.sublen), np.uint8) if want_bits else None
        secs = self._extract_batch(cfg, pcm2d, n_clips, clip_len, threads, bits, n, fft_f32)
        return (bits[:, :n] if want_bits else None), float(secs)

    def _extract_batch(self, cfg, pcm2d, n_clips, clip_len, threads, bits, n, fft_f32):
        return self.lib.lbad_oracle_extract_batch(C.byref(cfg), _p(pcm2d, C.c_float), C.c_uint32(n_clips), C.c_int64(clip_len), C.c_uint32(threads),
                                                  _p(bits, C.c_uint8) if bits is not None else None, C.c_uint32(max(n, 1)), None, C.c_int(1 if fft_f32 else 0))

    def search(self, db_bits, q_bits, rng, threads=1):
        """scores[q][c] = CompareToFingerprint(db[c], query[q], range); returns (scores, seconds)."""
        db = _u8(db_bits); q = _u8(q_bits)
        n_db, dbc, L = db.shape; n_q, qc, L2 = q.shape
        assert L == L2
        scores = np.zeros((n_q, n_db), np.float32)
        secs = self.lib.lbad_oracle_search(_p(db, C.c_uint8), C.c_uint32(n_db), C.c_uint32(dbc), _p(q, C.c_uint8), C.c_uint32(n_q), C.c_uint32(qc),
                                           C.c_uint32(L), C.c_uint32(rng), C.c_uint32(threads), _p(scores, C.c_float))
        return scores, float(secs)

    # -- synthetic PCM (defined once, in the port library, used by every leg) --
    def synth_clip(self, clip_id, n, sample_rate=5512.0, base_seed=BASE_SEED):
        out = np.zeros(n, np.float32)
        self.lib.lbad_synth_clip(C.c_uint64(base_seed), C.c_uint64(clip_id), C.c_int64(n), C.c_double(sample_rate), _p(out, C.c_float))
        return out

    def add_noise(self, pcm, seed, amplitude):
        out = _f32(pcm).copy()
        self.lib.lbad_synth_add_noise(C.c_uint64(seed), C.c_int64(len(out)), C.c_double(amplitude), _p(out, C.c_float))
        return out


class Ref(Port):
    """The compiled reference (its own .m files + the disclosed shim). Same surface as ``Port``."""
    kind = "reference"

    def __init__(self, fft_mode="f64"):
        Port.__init__(self)          # synth + helpers come from the port library
        if not os.path.exists(REF_SO):
            if not os.path.isdir(REFERENCE_SRC):
                raise FileNotFoundError("oracle/_ref not built and /root/reference not present")
            build(ref=True)
        R = self.ref = C.CDLL(REF_SO)
        R.lbad_ref_compare_fp.restype = C.c_float
        R.lbad_ref_compare_sub.restype = C.c_float
        R.lbad_ref_extract_batch.restype = C.c_double
        R.lbad_ref_search.restype = C.c_double
        self.set_fft_mode(fft_mode)

    @staticmethod
    def available():
        return os.path.exists(REF_SO) or os.path.isdir(REFERENCE_SRC)

    def set_fft_mode(self, mode):
        """"f64": the parity definition (exact DFT rounded once); "fast": the tuned single-precision FFT of the timing baseline;
        "f32": round 1's radix-2 Stockham stand-in (kept for comparison)."""
        self.ref.lbad_shim_set_fft_mode(C.c_int({"f64": 0, "f32": 1, "fast": 2}[mode]))

    def fft_us_per_window(self, n=2048, mode="fast", reps=20000):
        """Cost of one vDSP_ctoz + vDSP_fft_zrip + vDSP_ztoc (m:353-355) through the shim, in microseconds."""
        self.ref.lbad_ref_time_fft.restype = C.c_double
        return float(self.ref.lbad_ref_time_fft(C.c_uint32(n), C.c_int({"f64": 0, "f32": 1, "fast": 2}[mode]), C.c_uint32(reps))) * 1e6

    def band_energies(self, cfg, pcm, n_windows):
        pcm = _f32(pcm); out = np.zeros((n_windows, cfg.bands), np.float32)
        e = self.ref.lbad_ref_band_energies(C.byref(cfg), _p(pcm, C.c_float), C.c_int64(len(pcm)), C.c_uint32(n_windows), _p(out, C.c_float))
        assert e == 0
        return out

    def haar(self, image):
        img = _f32(image).copy()
        self.ref.lbad_ref_haar(_p(img, C.c_float), C.c_uint32(img.shape[0]), C.c_uint32(img.shape[1]))
        return img

    def extract_bits(self, coeffs, t):
        c = _f32(coeffs); assert c.ndim == 2
        out = np.zeros(2 * t, np.uint8)
        self.ref.lbad_ref_extract_bits(_p(c, C.c_float), C.c_uint32(c.shape[0]), C.c_uint32(c.shape[1]), C.c_uint32(t), _p(out, C.c_uint8))
        return out

    def process(self, cfg, pcm, stages=False, direct=None):
        """direct=False: LBAudioDetectiveProcessAudioURL as written; True: the Q14-safe route through the exported internals."""
        pcm = _f32(pcm); n = subfp_count(cfg, len(pcm))
        if direct is None:
            direct = stages or cfg.window not in (2048, 1024)
        bits = np.zeros((max(n, 1), cfg.sublen), np.uint8); cnt = C.c_uint32(0); ln = C.c_uint32(0)
        if not direct:
            self.ref.lbad_ref_process_pcm(C.byref(cfg), _p(pcm, C.c_float), C.c_int64(len(pcm)), _p(bits, C.c_uint8), C.c_uint32(n), C.byref(cnt), C.byref(ln))
            assert cnt.value == n and (n == 0 or ln.value == cfg.sublen)
            return bits[:n]
        img = np.zeros((max(n, 1), ROWS_PER_FRAME, cfg.bands), np.float32); haar = np.zeros_like(img)
        self.ref.lbad_ref_process_pcm_direct(C.byref(cfg), _p(pcm, C.c_float), C.c_int64(len(pcm)), _p(bits, C.c_uint8), C.c_uint32(n), C.byref(cnt), C.byref(ln),
                                             _p(img, C.c_float), _p(haar, C.c_float))
        assert cnt.value == n
        return (bits[:n], img[:n], haar[:n]) if stages else bits[:n]

    def compare_sub(self, s1, s2, length, rng):
        s1 = _u8(np.concatenate([s1, [0, 0]])); s2 = _u8(np.concatenate([s2, [0, 0]]))
        return float(self.ref.lbad_ref_compare_sub(_p(s1, C.c_uint8), _p(s2, C.c_uint8), C.c_uint32(length), C.c_uint32(rng)))

    def compare_fp(self, b1, b2, rng):
        b1 = _u8(b1); b2 = _u8(b2)
        L = b1.shape[1] if b1.ndim == 2 and b1.shape[0] else b2.shape[1]
        return float(self.ref.lbad_ref_compare_fp(_p(b1, C.c_uint8), C.c_uint32(b1.shape[0]), _p(b2, C.c_uint8), C.c_uint32(b2.shape[0]), C.c_uint32(L), C.c_uint32(rng)))

    def compare_pcm(self, cfg, p1, p2, rng=0):
        p1 = _f32(p1); p2 = _f32(p2); out = C.c_float(-1)
        self.ref.lbad_ref_compare_pcm(C.byref(cfg), _p(p1, C.c_float), C.c_int64(len(p1)), _p(p2, C.c_float), C.c_int64(len(p2)), C.c_uint32(rng), C.byref(out))
        return float(out.value)

    def _extract_batch(self, cfg, pcm2d, n_clips, clip_len, threads, bits, n, fft_f32):
        self.set_fft_mode(fft_f32 if isinstance(fft_f32, str) else ("fast" if fft_f32 else "f64"))
        try:
            return self.ref.lbad_ref_extract_batch(C.byref(cfg), _p(pcm2d, C.c_float), C.c_uint32(n_clips), C.c_int64(clip_len), C.c_uint32(threads),
                                                   _p(bits, C.c_uint8) if bits is not None else None, C.c_uint32(max(n, 1)), None)
        finally:
            self.set_fft_mode("f64")

    def extract_batch_stages(self, cfg, pcm2d, threads=1, images=True, haar=True, fft_mode="f64"):
        """(bits, images, haar, seconds) of a whole batch through the reference's own internals, threaded.  fft_mode "f64" is the parity
        definition; "fast" / "f32" run the same reference code on a single-precision FFT (what a float32 vDSP would deliver)."""
        pcm2d = _f32(pcm2d); n_clips, clip_len = pcm2d.shape; n = subfp_count(cfg, clip_len)
        bits = np.zeros((n_clips, max(n, 1), cfg.sublen), np.uint8)
        img = np.zeros((n_clips, max(n, 1), ROWS_PER_FRAME, cfg.bands), np.float32) if images else None
        hr = np.zeros((n_clips, max(n, 1), ROWS_PER_FRAME, cfg.bands), np.float32) if haar else None
        self.ref.lbad_ref_extract_batch_stages.restype = C.c_double
        self.set_fft_mode(fft_mode)
        try:
            secs = self.ref.lbad_ref_extract_batch_stages(C.byref(cfg), _p(pcm2d, C.c_float), C.c_uint32(n_clips), C.c_int64(clip_len), C.c_uint32(threads), C.c_uint32(max(n, 1)),
                                                          _p(img, C.c_float) if images else None, _p(hr, C.c_float) if haar else None, _p(bits, C.c_uint8))
        finally:
            self.set_fft_mode("f64")
        return bits[:, :n], (img[:, :n] if images else None), (hr[:, :n] if haar else None), float(secs)

    def search(self, db_bits, q_bits, rng, threads=1):
        db = _u8(db_bits); q = _u8(q_bits)
        n_db, dbc, L = db.shape; n_q, qc, L2 = q.shape
        assert L == L2
        scores = np.zeros((n_q, n_db), np.float32)
        secs = self.ref.lbad_ref_search(_p(db, C.c_uint8), C.c_uint32(n_db), C.c_uint32(dbc), _p(q, C.c_uint8), C.c_uint32(n_q), C.c_uint32(qc),
                                        C.c_uint32(L), C.c_uint32(rng), C.c_uint32(threads), _p(scores, C.c_float))
        return scores, float(secs)

    def set_window_size_status(self, n):
        return int(self.ref.lbad_ref_set_window_size_status(C.c_uint32(n)))


def best():
    """The strongest checker available: the compiled reference if it exists, else the port."""
    return Ref() if Ref.available() else Port()
