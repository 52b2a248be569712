"""ctypes mirror of include/LBAudioDetective*.h — same names, argument meaning and error behaviour as the C API.

Nothing here computes: every method marshals numpy buffers (or raw device pointers) into libLBAudioDetectiveCUDA.so.
"""
import ctypes as C
import os
import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libLBAudioDetectiveCUDA.so")
ROWS_PER_FRAME = 128
ARGUMENT_INVALID = 1
DEVICE_UNAVAILABLE = -7001
DEVICE_ERROR = -7002

_lib = None


class LBADError(RuntimeError):
    def __init__(self, status, what):
        self.status = status
        RuntimeError.__init__(self, "%s failed with OSStatus %d%s" % (what, status, _last_error()))


def _last_error():
    try:
        s = _lib.LBAudioDetectiveSupportLastError()
        return (": " + s.decode()) if s else ""
    except Exception:
        return ""


def load_library(build_if_missing: bool = True):
    """Loads the C-ABI library; fails loudly if it is missing (there is no fallback implementation)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LBAD_LIBRARY") or LIB_PATH      # LBAD_LIBRARY: another build of the same CUDA library (kernel A/B experiments, build.py --variant)
    if not os.path.exists(path):
        if not build_if_missing or path != LIB_PATH:
            raise FileNotFoundError(path + " is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        from . import build as _b
        _b.build()
    L = C.CDLL(path)
    vp, u32, u64, f32, f64, u8 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_double, C.c_ubyte
    P = C.POINTER
    sig = {
        "LBAudioDetectiveNew": (vp, []),
        "LBAudioDetectiveDispose": (C.c_int32, [vp]),
        "LBAudioDetectiveGetProcessingSampleRate": (f64, [vp]),
        "LBAudioDetectiveGetNumberOfPitchSteps": (u32, [vp]),
        "LBAudioDetectiveGetSubfingerprintLength": (u32, [vp]),
        "LBAudioDetectiveGetWindowSize": (u32, [vp]),
        "LBAudioDetectiveGetAnalysisStride": (u32, [vp]),
        "LBAudioDetectiveSetRecordingSampleRate": (C.c_int32, [vp, f64]),
        "LBAudioDetectiveSetProcessingSampleRate": (C.c_int32, [vp, f64]),
        "LBAudioDetectiveSetNumberOfPitchSteps": (C.c_int32, [vp, u32]),
        "LBAudioDetectiveSetSubfingerprintLength": (C.c_int32, [vp, u32]),
        "LBAudioDetectiveSetWindowSize": (C.c_int32, [vp, u32]),
        "LBAudioDetectiveSetAnalysisStride": (C.c_int32, [vp, u32]),
        "LBAudioDetectiveProcessPCM": (C.c_int32, [vp, vp, u64, P(vp)]),
        "LBAudioDetectiveComparePCM": (C.c_int32, [vp, vp, u64, vp, u64, u32, P(f32)]),
        "LBAudioDetectiveCheckConfiguration": (C.c_int32, [vp]),
        "LBAudioDetectiveGetRecordingSampleRate": (f64, [vp]),
        "LBAudioDetectiveGetResampledLength": (u64, [vp, u64]),
        "LBAudioDetectiveResamplePCM": (C.c_int32, [vp, vp, u64, vp]),
        "LBAudioDetectiveProcessRecordedPCM": (C.c_int32, [vp, vp, u64, P(vp)]),
        "LBAudioDetectiveProcessRecordedPCMBatchDevice": (C.c_int32, [vp, vp, u32, u64, u64, vp, vp]),
        "LBAudioDetectiveGetNumberOfSubfingerprintsForLength": (u64, [vp, u64]),
        "LBAudioDetectiveGetBandTable": (C.c_int32, [vp, vp, vp, vp]),
        "LBAudioDetectiveProcessPCMBatch": (C.c_int32, [vp, vp, u32, u64, u64, vp]),
        "LBAudioDetectiveProcessPCMBatchInt16": (C.c_int32, [vp, vp, u32, u64, u64, vp]),
        "LBAudioDetectiveSetDevice": (C.c_int32, [vp, C.c_int]),
        "LBAudioDetectiveGetDevice": (C.c_int, [vp]),
        "LBAudioDetectiveProcessPCMBatchSharded": (C.c_int32, [P(vp), u32, vp, u32, u64, u64, vp]),
        "LBAudioDetectiveProcessPCMBatchDevice": (C.c_int32, [vp, vp, u32, u64, u64, vp, vp]),
        "LBAudioDetectiveProcessPCMStages": (C.c_int32, [vp, vp, u64, vp, vp, vp, u8]),
        "LBAudioDetectiveProcessPCMBatchStages": (C.c_int32, [vp, vp, u32, u64, u64, vp, vp, vp, u8]),
        "LBAudioDetectiveTransformImages": (C.c_int32, [vp, vp, u32, vp, vp]),
        "LBAudioDetectiveStreamNew": (vp, [vp]),
        "LBAudioDetectiveStreamDispose": (C.c_int32, [vp]),
        "LBAudioDetectiveStreamAppend": (C.c_int32, [vp, vp, u64]),
        "LBAudioDetectiveStreamGetFingerprint": (vp, [vp]),
        "LBAudioDetectiveStreamGetNumberOfPendingFrames": (u64, [vp]),
        "LBAudioDetectiveGetKernelLaunchCount": (u64, [vp]),
        "LBAudioDetectiveGetKernelTiming": (u32, [vp, u8, u8, P(f64)]),
        "LBAudioDetectiveGetTransformKernelTiming": (u32, [vp, u8, u8, P(f64)]),
        "LBAudioDetectiveFingerprintNew": (vp, [u32]),
        "LBAudioDetectiveFingerprintDispose": (None, [vp]),
        "LBAudioDetectiveFingerprintCopy": (vp, [vp]),
        "LBAudioDetectiveFingerprintGetSubfingerprintLength": (u32, [vp]),
        "LBAudioDetectiveFingerprintGetNumberOfSubfingerprints": (u32, [vp]),
        "LBAudioDetectiveFingerprintGetSubfingerprintAtIndex": (u32, [vp, u32, vp]),
        "LBAudioDetectiveFingerprintSetSubfingerprintLength": (u8, [vp, P(u32)]),
        "LBAudioDetectiveFingerprintAddSubfingerprint": (None, [vp, vp]),
        "LBAudioDetectiveFingerprintEqualToFingerprint": (u8, [vp, vp]),
        "LBAudioDetectiveFingerprintCompareToFingerprint": (f32, [vp, vp, u32]),
        "LBAudioDetectiveFingerprintCompareSubfingerprints": (f32, [vp, vp, vp, u32]),
        "LBAudioDetectiveFingerprintCompareToFingerprintStatus": (C.c_int32, [vp, vp, u32, P(f32)]),
        "LBAudioDetectiveFingerprintPackedWordsPerPlane": (u32, [u32]),
        "LBAudioDetectiveFingerprintGetPackedSubfingerprintAtIndex": (u32, [vp, u32, vp]),
        "LBAudioDetectiveFingerprintAddPackedSubfingerprints": (C.c_int32, [vp, vp, u32]),
        "LBAudioDetectiveFingerprintToString": (C.c_size_t, [vp, C.c_char_p, C.c_size_t]),
        "LBAudioDetectiveFingerprintFromString": (vp, [C.c_char_p]),
        "LBAudioDetectiveFrameNew": (vp, [u32]),
        "LBAudioDetectiveFrameDispose": (None, [vp]),
        "LBAudioDetectiveFrameCopy": (vp, [vp]),
        "LBAudioDetectiveFrameGetNumberOfRows": (u32, [vp]),
        "LBAudioDetectiveFrameGetRow": (P(f32), [vp, u32]),
        "LBAudioDetectiveFrameGetValue": (f32, [vp, u32, u32]),
        "LBAudioDetectiveFrameFull": (u8, [vp]),
        "LBAudioDetectiveFrameSetRow": (u8, [vp, vp, u32, u32]),
        "LBAudioDetectiveFrameDecompose": (None, [vp]),
        "LBAudioDetectiveFrameFingerprintSize": (C.c_size_t, [vp]),
        "LBAudioDetectiveFrameFingerprintLength": (u32, [vp]),
        "LBAudioDetectiveFrameExtractFingerprint": (None, [vp, u32, vp]),
        "LBAudioDetectiveFrameEqualToFrame": (u8, [vp, vp]),
        "LBAudioDetectiveFrameDecomposeStatus": (C.c_int32, [vp]),
        "LBAudioDetectiveFrameExtractFingerprintStatus": (C.c_int32, [vp, u32, vp]),
        "LBAudioDetectiveDatabaseNew": (vp, [u32]),
        "LBAudioDetectiveDatabaseDispose": (C.c_int32, [vp]),
        "LBAudioDetectiveDatabaseGetNumberOfClips": (u32, [vp]),
        "LBAudioDetectiveDatabaseGetNumberOfSubfingerprints": (u64, [vp]),
        "LBAudioDetectiveDatabaseSetClipIndexBase": (C.c_int32, [vp, u32]),
        "LBAudioDetectiveDatabaseAddFingerprint": (C.c_int32, [vp, vp, P(u32)]),
        "LBAudioDetectiveDatabaseAddPacked": (C.c_int32, [vp, vp, u32, vp, u32]),
        "LBAudioDetectiveDatabaseAddPackedDevice": (C.c_int32, [vp, vp, u32, u32, vp]),
        "LBAudioDetectiveDatabaseSearchPacked": (C.c_int32, [vp, vp, u32, u32, u32, u32, vp, vp, vp]),
        "LBAudioDetectiveDatabaseSearch": (C.c_int32, [vp, P(vp), u32, u32, u32, vp, vp]),
        "LBAudioDetectiveDatabaseSearchDevice": (C.c_int32, [vp, vp, u32, u32, u32, u32, vp, vp, vp]),
        "LBAudioDetectiveDatabaseMergeTopK": (C.c_int32, [vp, vp, u32, u32, u32, vp, vp]),
        "LBAudioDetectiveDatabaseMergeTopKDevice": (C.c_int32, [vp, vp, u32, u32, u32, vp, vp, vp]),
        "LBAudioDetectiveDatabaseMergeTopKDeviceStrided": (C.c_int32, [vp, vp, u32, u64, u32, u32, vp, vp, vp]),
        "LBAudioDetectiveDatabaseGroupNew": (vp, [u32, P(C.c_int), u32]),
        "LBAudioDetectiveDatabaseGroupDispose": (C.c_int32, [vp]),
        "LBAudioDetectiveDatabaseGroupGetNumberOfShards": (u32, [vp]),
        "LBAudioDetectiveDatabaseGroupGetNumberOfClips": (u64, [vp]),
        "LBAudioDetectiveDatabaseGroupGetShardDevice": (C.c_int, [vp, u32]),
        "LBAudioDetectiveDatabaseGroupGetShardNumberOfClips": (u32, [vp, u32]),
        "LBAudioDetectiveDatabaseGroupAddPacked": (C.c_int32, [vp, vp, u32, vp, u32, P(u64)]),
        "LBAudioDetectiveDatabaseGroupAddFingerprint": (C.c_int32, [vp, vp, P(u64)]),
        "LBAudioDetectiveDatabaseGroupAddPackedDeviceToShard": (C.c_int32, [vp, u32, vp, u32, u32, u64, vp]),
        "LBAudioDetectiveDatabaseGroupSearchPacked": (C.c_int32, [vp, vp, u32, u32, u32, u32, vp, vp]),
        "LBAudioDetectiveDatabaseGroupSearch": (C.c_int32, [vp, P(vp), u32, u32, u32, vp, vp]),
        "LBAudioDetectiveDatabaseGroupGetKernelLaunchCount": (u64, [vp]),
        "LBAudioDetectiveDatabaseGroupGetLastSearchMilliseconds": (f64, [vp]),
        "LBAudioDetectiveDatabaseSave": (C.c_int32, [vp, C.c_char_p]),
        "LBAudioDetectiveDatabaseLoad": (vp, [C.c_char_p]),
        "LBAudioDetectiveDatabaseComparesPerQuery": (u64, [vp, u32]),
        "LBAudioDetectiveDatabaseGetKernelLaunchCount": (u64, [vp]),
        "LBAudioDetectiveDatabaseGetKernelTiming": (u32, [vp, u8, u8, P(f64)]),
        "LBAudioDetectiveSupportSynthesizeDevice": (C.c_int32, [vp, u32, u64, u64, u64, u64, f64, vp]),
        "LBAudioDetectiveSupportRandomCodesDevice": (C.c_int32, [vp, u64, u32, u64, vp]),
        "LBAudioDetectiveSupportRandomCodesDeviceAt": (C.c_int32, [vp, u64, u32, u64, u64, vp]),
        "LBAudioDetectiveSupportDeviceCount": (C.c_int, []),
        "LBAudioDetectiveSupportMicrobench": (C.c_int32, [P(f64), P(f64), P(f64)]),
        "LBAudioDetectiveSupportLastError": (C.c_char_p, []),
        "LBAudioDetectiveSupportDeviceAvailable": (u8, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError here = the library does not export what include/*.h declares
        fn.restype = res
        fn.argtypes = args
    L._lbad_signatures = sig
    _lib = L
    return L


def lib():
    return load_library()


def device_available() -> bool:
    return bool(lib().LBAudioDetectiveSupportDeviceAvailable())


def _check(status, what):
    if status != 0:
        raise LBADError(status, what)


def _ptr(a):
    return a.ctypes.data if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def words_per_plane(length: int) -> int:
    return int(lib().LBAudioDetectiveFingerprintPackedWordsPerPlane(length))


def pack_booleans(bits: np.ndarray) -> np.ndarray:
    """[..., L] Booleans -> [..., 2*W] packed words (P plane then M plane). Pure data-format conversion."""
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    L = bits.shape[-1]; W = words_per_plane(L)
    flat = bits.reshape(-1, L)
    out = np.zeros((flat.shape[0], 2 * W), np.uint32)
    for plane in (0, 1):
        b = flat[:, plane::2].astype(np.uint64)
        for w in range(W):
            chunk = b[:, 32 * w:32 * (w + 1)]
            if chunk.shape[1]:
                out[:, plane * W + w] = (chunk << np.arange(chunk.shape[1], dtype=np.uint64)).sum(axis=1).astype(np.uint32)
    return out.reshape(bits.shape[:-1] + (2 * W,))


def unpack_words(words: np.ndarray, length: int) -> np.ndarray:
    """Inverse of pack_booleans."""
    words = np.ascontiguousarray(words, dtype=np.uint32)
    W = words.shape[-1] // 2
    flat = words.reshape(-1, 2 * W)
    out = np.zeros((flat.shape[0], length), np.uint8)
    for i in range(length):
        pair = i >> 1
        out[:, i] = (flat[:, (W if i & 1 else 0) + (pair >> 5)] >> np.uint32(pair & 31)) & 1
    return out.reshape(words.shape[:-1] + (length,))


class Fingerprint:
    """LBAudioDetectiveFingerprintRef (include/LBAudioDetectiveFingerprint.h)."""

    def __init__(self, subfingerprint_length=0, _ref=None):
        self._L = lib()
        self.ref = _ref if _ref is not None else self._L.LBAudioDetectiveFingerprintNew(subfingerprint_length)

    def dispose(self):
        if self.ref:
            self._L.LBAudioDetectiveFingerprintDispose(self.ref); self.ref = None

    def __del__(self):
        try:
            self.dispose()
        except Exception:
            pass

    def copy(self):
        return Fingerprint(_ref=self._L.LBAudioDetectiveFingerprintCopy(self.ref))

    @property
    def subfingerprint_length(self):
        return int(self._L.LBAudioDetectiveFingerprintGetSubfingerprintLength(self.ref))

    @property
    def count(self):
        return int(self._L.LBAudioDetectiveFingerprintGetNumberOfSubfingerprints(self.ref))

    def set_subfingerprint_length(self, n):
        io = C.c_uint32(n)
        ok = bool(self._L.LBAudioDetectiveFingerprintSetSubfingerprintLength(self.ref, C.byref(io)))
        return ok, int(io.value)

    def subfingerprint(self, i):
        out = np.zeros(self.subfingerprint_length, np.uint8)
        self._L.LBAudioDetectiveFingerprintGetSubfingerprintAtIndex(self.ref, i, _ptr(out))
        return out

    def booleans(self):
        n, L = self.count, self.subfingerprint_length
        return np.stack([self.subfingerprint(i) for i in range(n)]) if n else np.zeros((0, L), np.uint8)

    def packed(self):
        n, W = self.count, words_per_plane(self.subfingerprint_length)
        out = np.zeros((n, 2 * W), np.uint32)
        for i in range(n):
            self._L.LBAudioDetectiveFingerprintGetPackedSubfingerprintAtIndex(self.ref, i, _ptr(out[i]))
        return out

    def add_subfingerprint(self, bits):
        b = np.ascontiguousarray(bits, dtype=np.uint8)
        assert b.size >= self.subfingerprint_length
        self._L.LBAudioDetectiveFingerprintAddSubfingerprint(self.ref, _ptr(b))

    def add_packed(self, words):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        _check(self._L.LBAudioDetectiveFingerprintAddPackedSubfingerprints(self.ref, _ptr(w), w.reshape(-1, w.shape[-1]).shape[0]), "AddPackedSubfingerprints")

    @staticmethod
    def from_booleans(bits):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        fp = Fingerprint(bits.shape[1] if bits.ndim == 2 else 0)
        for row in bits:
            fp.add_subfingerprint(row)
        return fp

    def equal(self, other):
        return bool(self._L.LBAudioDetectiveFingerprintEqualToFingerprint(self.ref, other.ref))

    def compare(self, other, rng):
        """LBAudioDetectiveFingerprintCompareToFingerprint(self, other, rng) — runs on the GPU."""
        return float(self._L.LBAudioDetectiveFingerprintCompareToFingerprint(self.ref, other.ref, rng))

    def compare_status(self, other, rng):
        """(OSStatus, match) of LBAudioDetectiveFingerprintCompareToFingerprintStatus."""
        out = C.c_float(-1.0)
        st = self._L.LBAudioDetectiveFingerprintCompareToFingerprintStatus(self.ref, other.ref, rng, C.byref(out))
        return int(st), float(out.value)

    def compare_subfingerprints(self, s1, s2, rng):
        a = np.ascontiguousarray(np.concatenate([s1, [0, 0]]), dtype=np.uint8); b = np.ascontiguousarray(np.concatenate([s2, [0, 0]]), dtype=np.uint8)
        return float(self._L.LBAudioDetectiveFingerprintCompareSubfingerprints(self.ref, _ptr(a), _ptr(b), rng))

    def to_string(self):
        need = self._L.LBAudioDetectiveFingerprintToString(self.ref, None, 0)
        buf = C.create_string_buffer(need)
        self._L.LBAudioDetectiveFingerprintToString(self.ref, buf, need)
        return buf.value.decode()

    @staticmethod
    def from_string(s):
        ref = lib().LBAudioDetectiveFingerprintFromString(s.encode())
        return Fingerprint(_ref=ref) if ref else None


class Frame:
    """ctypes mirror of LBAudioDetectiveFrameRef (include/LBAudioDetectiveFrame.h; LBAudioDetectiveFrame.h:27-162 upstream)."""

    def __init__(self, max_rows, _ref=None):
        self._L = lib()
        self.ref = _ref if _ref is not None else self._L.LBAudioDetectiveFrameNew(max_rows)

    def dispose(self):
        if getattr(self, "ref", None):
            self._L.LBAudioDetectiveFrameDispose(self.ref); self.ref = None

    def __del__(self):
        try:
            self.dispose()
        except Exception:
            pass

    @staticmethod
    def from_array(a):
        a = np.asarray(a, np.float32); f = Frame(a.shape[0])
        for r in range(a.shape[0]):
            assert f.set_row(a[r], r)
        return f

    def copy(self): return Frame(0, _ref=self._L.LBAudioDetectiveFrameCopy(self.ref))
    @property
    def rows(self): return int(self._L.LBAudioDetectiveFrameGetNumberOfRows(self.ref))
    @property
    def fingerprint_length(self): return int(self._L.LBAudioDetectiveFrameFingerprintLength(self.ref))
    @property
    def fingerprint_size(self): return int(self._L.LBAudioDetectiveFrameFingerprintSize(self.ref))
    @property
    def row_length(self): return self.fingerprint_length // (2 * self.rows) if self.rows else 0
    def full(self): return bool(self._L.LBAudioDetectiveFrameFull(self.ref))
    def value(self, r, c): return float(self._L.LBAudioDetectiveFrameGetValue(self.ref, r, c))
    def equal(self, other): return bool(self._L.LBAudioDetectiveFrameEqualToFrame(self.ref, other.ref))

    def set_row(self, row, index, count=None):
        row = _f32(row)
        return bool(self._L.LBAudioDetectiveFrameSetRow(self.ref, _ptr(row), index, len(row) if count is None else count))

    def row(self, r):
        p = self._L.LBAudioDetectiveFrameGetRow(self.ref, r)
        return np.ctypeslib.as_array(p, shape=(self.row_length,)).copy()

    def array(self):
        return np.stack([self.row(r) for r in range(self.rows)]) if self.rows else np.zeros((0, 0), np.float32)

    def decompose(self, status=False):
        if status:
            return int(self._L.LBAudioDetectiveFrameDecomposeStatus(self.ref))
        self._L.LBAudioDetectiveFrameDecompose(self.ref)

    def extract_fingerprint(self, wavelets, status=False):
        out = np.zeros(2 * wavelets, np.uint8)
        if status:
            return int(self._L.LBAudioDetectiveFrameExtractFingerprintStatus(self.ref, wavelets, _ptr(out))), out
        self._L.LBAudioDetectiveFrameExtractFingerprint(self.ref, wavelets, _ptr(out))
        return out


class Detective:
    """LBAudioDetectiveRef (include/LBAudioDetective.h)."""

    def __init__(self):
        self._L = lib()
        self.ref = self._L.LBAudioDetectiveNew()

    def dispose(self):
        if self.ref:
            st = self._L.LBAudioDetectiveDispose(self.ref); self.ref = None
            return st
        return ARGUMENT_INVALID

    def __del__(self):
        try:
            self.dispose()
        except Exception:
            pass

    # getters / setters, same names as the C API minus the prefix
    sample_rate = property(lambda s: float(s._L.LBAudioDetectiveGetProcessingSampleRate(s.ref)))
    pitch_steps = property(lambda s: int(s._L.LBAudioDetectiveGetNumberOfPitchSteps(s.ref)))
    subfingerprint_length = property(lambda s: int(s._L.LBAudioDetectiveGetSubfingerprintLength(s.ref)))
    window_size = property(lambda s: int(s._L.LBAudioDetectiveGetWindowSize(s.ref)))
    analysis_stride = property(lambda s: int(s._L.LBAudioDetectiveGetAnalysisStride(s.ref)))

    def set_sample_rate(self, v): return int(self._L.LBAudioDetectiveSetProcessingSampleRate(self.ref, v))
    def set_pitch_steps(self, v): return int(self._L.LBAudioDetectiveSetNumberOfPitchSteps(self.ref, v))
    def set_subfingerprint_length(self, v): return int(self._L.LBAudioDetectiveSetSubfingerprintLength(self.ref, v))
    def set_window_size(self, v): return int(self._L.LBAudioDetectiveSetWindowSize(self.ref, v))
    def set_analysis_stride(self, v): return int(self._L.LBAudioDetectiveSetAnalysisStride(self.ref, v))
    def check_configuration(self): return int(self._L.LBAudioDetectiveCheckConfiguration(self.ref))

    # ---- recording-rate front end (include/LBAudioDetectiveResample.h) ----
    def set_recording_rate(self, v): return int(self._L.LBAudioDetectiveSetRecordingSampleRate(self.ref, v))

    @property
    def recording_rate(self): return float(self._L.LBAudioDetectiveGetRecordingSampleRate(self.ref))

    def resampled_length(self, n): return int(self._L.LBAudioDetectiveGetResampledLength(self.ref, n))

    def resample(self, pcm):
        pcm = _f32(pcm); out = np.zeros(self.resampled_length(len(pcm)), np.float32)
        _check(self._L.LBAudioDetectiveResamplePCM(self.ref, _ptr(pcm), len(pcm), _ptr(out)), "LBAudioDetectiveResamplePCM")
        return out

    def process_recorded_pcm(self, pcm):
        pcm = _f32(pcm); ref = C.c_void_p()
        st = self._L.LBAudioDetectiveProcessRecordedPCM(self.ref, _ptr(pcm), len(pcm), C.byref(ref))
        fp = Fingerprint(_ref=ref.value) if ref.value else None
        _check(st, "LBAudioDetectiveProcessRecordedPCM")
        return fp

    def process_recorded_batch_device(self, d_pcm_ptr, n_clips, clip_len, clip_stride, d_words_ptr, stream=None):
        _check(self._L.LBAudioDetectiveProcessRecordedPCMBatchDevice(self.ref, d_pcm_ptr, n_clips, clip_len, clip_stride, d_words_ptr, stream), "LBAudioDetectiveProcessRecordedPCMBatchDevice")

    def subfingerprints_for_length(self, n):
        return int(self._L.LBAudioDetectiveGetNumberOfSubfingerprintsForLength(self.ref, n))

    def band_table(self):
        B = self.pitch_steps
        idx = np.zeros(B + 1, np.uint32); lo = np.zeros(B, np.uint32); hi = np.zeros(B, np.uint32)
        _check(self._L.LBAudioDetectiveGetBandTable(self.ref, _ptr(idx), _ptr(lo), _ptr(hi)), "GetBandTable")
        return idx, lo, hi

    def process_pcm(self, pcm, check=True):
        """LBAudioDetectiveProcessPCM -> Fingerprint (raises LBADError unless check=False, then returns (status, fp))."""
        pcm = _f32(pcm); out = C.c_void_p()
        st = self._L.LBAudioDetectiveProcessPCM(self.ref, _ptr(pcm), pcm.size, C.byref(out))
        fp = Fingerprint(_ref=out.value) if out.value else None
        if not check:
            return int(st), fp
        _check(st, "LBAudioDetectiveProcessPCM")
        return fp

    def compare_pcm(self, pcm1, pcm2, rng=0):
        a = _f32(pcm1); b = _f32(pcm2); out = C.c_float(-1.0)
        _check(self._L.LBAudioDetectiveComparePCM(self.ref, _ptr(a), a.size, _ptr(b), b.size, rng, C.byref(out)), "LBAudioDetectiveComparePCM")
        return float(out.value)

    def process_batch(self, pcm2d, out_words=None):
        """Host [clips][samples] float32 -> [clips][subfps][2W] packed words."""
        pcm2d = _f32(pcm2d); n_clips, clip_len = pcm2d.shape
        n = self.subfingerprints_for_length(clip_len); W = words_per_plane(self.subfingerprint_length)
        if out_words is None:
            out_words = np.zeros((n_clips, n, 2 * W), np.uint32)
        _check(self._L.LBAudioDetectiveProcessPCMBatch(self.ref, _ptr(pcm2d), n_clips, clip_len, clip_len, _ptr(out_words)), "LBAudioDetectiveProcessPCMBatch")
        return out_words

    def set_device(self, device): return int(self._L.LBAudioDetectiveSetDevice(self.ref, device))
    @property
    def device(self): return int(self._L.LBAudioDetectiveGetDevice(self.ref))

    @staticmethod
    def process_batch_sharded(detectives, pcm2d, out_words=None, host_ptr=None, n_clips=None, clip_len=None, clip_stride=None):
        """LBAudioDetectiveProcessPCMBatchSharded: one batch over several detectives (one per GPU), each on its own host thread."""
        d0 = detectives[0]
        refs = (C.c_void_p * len(detectives))(*[d.ref for d in detectives])
        if host_ptr is None:
            pcm2d = np.ascontiguousarray(pcm2d, np.float32); n_clips, clip_len = pcm2d.shape; clip_stride = clip_len; host_ptr = pcm2d.ctypes.data
        count = d0.subfingerprints_for_length(clip_len); W = words_per_plane(d0.subfingerprint_length)
        if out_words is None:
            out_words = np.zeros((n_clips, count, 2 * W), np.uint32)
        out_ptr = out_words if isinstance(out_words, int) else out_words.ctypes.data
        _check(d0._L.LBAudioDetectiveProcessPCMBatchSharded(refs, len(detectives), host_ptr, n_clips, clip_len, clip_stride, out_ptr), "LBAudioDetectiveProcessPCMBatchSharded")
        return out_words

    def process_batch_int16(self, pcm2d):
        """Host [clips][samples] int16 -> packed words (converted on the device as x / 32768)."""
        pcm2d = np.ascontiguousarray(pcm2d, dtype=np.int16); n_clips, clip_len = pcm2d.shape
        n = self.subfingerprints_for_length(clip_len); W = words_per_plane(self.subfingerprint_length)
        out = np.zeros((n_clips, n, 2 * W), np.uint32)
        _check(self._L.LBAudioDetectiveProcessPCMBatchInt16(self.ref, _ptr(pcm2d), n_clips, clip_len, clip_len, _ptr(out)), "LBAudioDetectiveProcessPCMBatchInt16")
        return out

    def process_batch_int16_ptr(self, host_ptr, n_clips, clip_len, clip_stride, out_ptr):
        _check(self._L.LBAudioDetectiveProcessPCMBatchInt16(self.ref, host_ptr, n_clips, clip_len, clip_stride, out_ptr), "LBAudioDetectiveProcessPCMBatchInt16")

    def process_batch_ptr(self, host_ptr, n_clips, clip_len, clip_stride, out_ptr):
        _check(self._L.LBAudioDetectiveProcessPCMBatch(self.ref, host_ptr, n_clips, clip_len, clip_stride, out_ptr), "LBAudioDetectiveProcessPCMBatch")

    def process_batch_device(self, d_pcm_ptr, n_clips, clip_len, clip_stride, d_words_ptr, stream=None):
        _check(self._L.LBAudioDetectiveProcessPCMBatchDevice(self.ref, d_pcm_ptr, n_clips, clip_len, clip_stride, d_words_ptr, stream), "LBAudioDetectiveProcessPCMBatchDevice")

    def process_stages(self, pcm, fused):
        """(images, haar, booleans) of one clip; fused selects the fused kernel or the generic two-kernel path."""
        pcm = _f32(pcm); n = self.subfingerprints_for_length(pcm.size); B = self.pitch_steps; L = self.subfingerprint_length
        img = np.zeros((n, ROWS_PER_FRAME, B), np.float32); haar = np.zeros_like(img); bits = np.zeros((n, L), np.uint8)
        _check(self._L.LBAudioDetectiveProcessPCMStages(self.ref, _ptr(pcm), pcm.size, _ptr(img), _ptr(haar), _ptr(bits), 1 if fused else 0), "LBAudioDetectiveProcessPCMStages")
        return img, haar, bits

    def process_batch_stages(self, pcm2d, fused=True, images=True, haar=True):
        """Host [clips][samples] -> (words [clips][subfps][2W], images, haar [clips][subfps][128][B]) through the batch kernels."""
        pcm2d = _f32(pcm2d); n_clips, clip_len = pcm2d.shape
        n = self.subfingerprints_for_length(clip_len); W = words_per_plane(self.subfingerprint_length); B = self.pitch_steps
        words = np.zeros((n_clips, n, 2 * W), np.uint32)
        img = np.zeros((n_clips, n, ROWS_PER_FRAME, B), np.float32) if images else None
        hr = np.zeros((n_clips, n, ROWS_PER_FRAME, B), np.float32) if haar else None
        _check(self._L.LBAudioDetectiveProcessPCMBatchStages(self.ref, _ptr(pcm2d), n_clips, clip_len, clip_len, _ptr(words), _ptr(img), _ptr(hr), 1 if fused else 0), "LBAudioDetectiveProcessPCMBatchStages")
        return words, img, hr

    def transform_images(self, images):
        images = _f32(images); n = images.shape[0]
        haar = np.zeros_like(images); bits = np.zeros((n, self.subfingerprint_length), np.uint8)
        _check(self._L.LBAudioDetectiveTransformImages(self.ref, _ptr(images), n, _ptr(haar), _ptr(bits)), "LBAudioDetectiveTransformImages")
        return haar, bits

    @property
    def kernel_launches(self):
        return int(self._L.LBAudioDetectiveGetKernelLaunchCount(self.ref))

    def kernel_timing(self, enable=True, reset=True, transform=False):
        """(launches, total ms) of the FFT + band-energy kernel, or of the Haar/select/pack kernel with transform=True."""
        ms = C.c_double(0.0)
        fn = self._L.LBAudioDetectiveGetTransformKernelTiming if transform else self._L.LBAudioDetectiveGetKernelTiming
        n = fn(self.ref, 1 if enable else 0, 1 if reset else 0, C.byref(ms))
        return int(n), float(ms.value)


class Stream:
    """LBAudioDetectiveStreamRef: append PCM, subfingerprints appear as frames complete."""

    def __init__(self, detective):
        self._L = lib(); self.detective = detective
        self.ref = self._L.LBAudioDetectiveStreamNew(detective.ref)
        if not self.ref:
            raise LBADError(ARGUMENT_INVALID, "LBAudioDetectiveStreamNew")

    def dispose(self):
        if self.ref:
            self._L.LBAudioDetectiveStreamDispose(self.ref); self.ref = None

    def __del__(self):
        try:
            self.dispose()
        except Exception:
            pass

    def append(self, pcm):
        pcm = _f32(pcm)
        _check(self._L.LBAudioDetectiveStreamAppend(self.ref, _ptr(pcm), pcm.size), "LBAudioDetectiveStreamAppend")

    def fingerprint(self):
        """A copy of the stream's fingerprint so far."""
        return Fingerprint(_ref=self._L.LBAudioDetectiveFingerprintCopy(self._L.LBAudioDetectiveStreamGetFingerprint(self.ref)))

    @property
    def pending(self):
        return int(self._L.LBAudioDetectiveStreamGetNumberOfPendingFrames(self.ref))


class Database:
    """LBAudioDetectiveDatabaseRef (include/LBAudioDetectiveDatabase.h)."""

    def __init__(self, subfingerprint_length=200, _ref=None):
        self._L = lib()
        self.L = subfingerprint_length
        self.ref = _ref if _ref is not None else self._L.LBAudioDetectiveDatabaseNew(subfingerprint_length)
        if not self.ref:
            raise LBADError(DEVICE_UNAVAILABLE, "LBAudioDetectiveDatabaseNew")

    def save(self, path):
        _check(self._L.LBAudioDetectiveDatabaseSave(self.ref, os.fsencode(path)), "DatabaseSave")

    @staticmethod
    def load(path, subfingerprint_length=200):
        ref = lib().LBAudioDetectiveDatabaseLoad(os.fsencode(path))
        return Database(subfingerprint_length, _ref=ref) if ref else None

    def dispose(self):
        if self.ref:
            self._L.LBAudioDetectiveDatabaseDispose(self.ref); self.ref = None

    def __del__(self):
        try:
            self.dispose()
        except Exception:
            pass

    clips = property(lambda s: int(s._L.LBAudioDetectiveDatabaseGetNumberOfClips(s.ref)))
    subfingerprints = property(lambda s: int(s._L.LBAudioDetectiveDatabaseGetNumberOfSubfingerprints(s.ref)))

    def set_clip_index_base(self, base):
        _check(self._L.LBAudioDetectiveDatabaseSetClipIndexBase(self.ref, base), "SetClipIndexBase")

    def add_fingerprint(self, fp):
        idx = C.c_uint32(0)
        _check(self._L.LBAudioDetectiveDatabaseAddFingerprint(self.ref, fp.ref, C.byref(idx)), "DatabaseAddFingerprint")
        return int(idx.value)

    def add_packed(self, words, counts=None):
        """words: [clips][count][2W] (uniform) or flat [total][2W] with counts[clips]."""
        w = np.ascontiguousarray(words, dtype=np.uint32)
        if counts is None:
            n_clips, uniform = w.shape[0], w.shape[1]
            _check(self._L.LBAudioDetectiveDatabaseAddPacked(self.ref, _ptr(w), n_clips, None, uniform), "DatabaseAddPacked")
        else:
            c = np.ascontiguousarray(counts, dtype=np.uint32)
            _check(self._L.LBAudioDetectiveDatabaseAddPacked(self.ref, _ptr(w), c.size, _ptr(c), 0), "DatabaseAddPacked")

    def add_packed_device(self, d_words_ptr, n_clips, uniform_count, producer_stream=None):
        """producer_stream: the CUDA stream the words are being produced on (the append is ordered after it); None if they are complete."""
        _check(self._L.LBAudioDetectiveDatabaseAddPackedDevice(self.ref, d_words_ptr, n_clips, uniform_count, producer_stream), "DatabaseAddPackedDevice")

    def search_packed(self, qwords, k, rng=0, all_scores=False):
        """qwords [queries][count][2W] -> (scores [q][k], clip indices [q][k][, full score matrix])."""
        q = np.ascontiguousarray(qwords, dtype=np.uint32); n_q, cq = q.shape[0], q.shape[1]
        sc = np.zeros((n_q, k), np.float32); idx = np.zeros((n_q, k), np.uint32)
        full = np.zeros((n_q, self.clips), np.float32) if all_scores else None
        _check(self._L.LBAudioDetectiveDatabaseSearchPacked(self.ref, _ptr(q), n_q, cq, rng, k, _ptr(sc), _ptr(idx), _ptr(full)), "DatabaseSearchPacked")
        return (sc, idx, full) if all_scores else (sc, idx)

    def search(self, fingerprints, k, rng=0):
        n = len(fingerprints); arr = (C.c_void_p * n)(*[f.ref for f in fingerprints])
        sc = np.zeros((n, k), np.float32); idx = np.zeros((n, k), np.uint32)
        _check(self._L.LBAudioDetectiveDatabaseSearch(self.ref, arr, n, rng, k, _ptr(sc), _ptr(idx)), "DatabaseSearch")
        return sc, idx

    def search_device(self, d_q_ptr, n_q, q_count, k, d_scores_ptr, d_idx_ptr, rng=0, stream=None):
        _check(self._L.LBAudioDetectiveDatabaseSearchDevice(self.ref, d_q_ptr, n_q, q_count, rng, k, d_scores_ptr, d_idx_ptr, stream), "DatabaseSearchDevice")

    def compares_per_query(self, q_count):
        return int(self._L.LBAudioDetectiveDatabaseComparesPerQuery(self.ref, q_count))

    @property
    def kernel_launches(self):
        return int(self._L.LBAudioDetectiveDatabaseGetKernelLaunchCount(self.ref))

    def kernel_timing(self, enable=True, reset=True):
        ms = C.c_double(0.0)
        n = self._L.LBAudioDetectiveDatabaseGetKernelTiming(self.ref, 1 if enable else 0, 1 if reset else 0, C.byref(ms))
        return int(n), float(ms.value)


class DatabaseGroup:
    """LBAudioDetectiveDatabaseGroupRef: a database sharded over several GPUs of this process (include/LBAudioDetectiveDatabase.h)."""

    def __init__(self, subfingerprint_length=200, devices=(0,)):
        self._L = lib(); self.L = subfingerprint_length
        arr = (C.c_int * len(devices))(*devices)
        self.ref = self._L.LBAudioDetectiveDatabaseGroupNew(subfingerprint_length, arr, len(devices))
        if not self.ref:
            raise LBADError(DEVICE_UNAVAILABLE, "LBAudioDetectiveDatabaseGroupNew")

    def dispose(self):
        if self.ref:
            self._L.LBAudioDetectiveDatabaseGroupDispose(self.ref); self.ref = None

    def __del__(self):
        try:
            self.dispose()
        except Exception:
            pass

    shards = property(lambda s: int(s._L.LBAudioDetectiveDatabaseGroupGetNumberOfShards(s.ref)))
    clips = property(lambda s: int(s._L.LBAudioDetectiveDatabaseGroupGetNumberOfClips(s.ref)))

    def shard_device(self, i): return int(self._L.LBAudioDetectiveDatabaseGroupGetShardDevice(self.ref, i))
    def shard_clips(self, i): return int(self._L.LBAudioDetectiveDatabaseGroupGetShardNumberOfClips(self.ref, i))

    def add_packed(self, words, counts=None):
        """words: [clips][count][2W] (uniform) or flat [total][2W] with counts[clips]; returns the first global clip index."""
        w = np.ascontiguousarray(words, dtype=np.uint32); first = C.c_uint64(0)
        if counts is None:
            _check(self._L.LBAudioDetectiveDatabaseGroupAddPacked(self.ref, _ptr(w), w.shape[0], None, w.shape[1], C.byref(first)), "DatabaseGroupAddPacked")
        else:
            c = np.ascontiguousarray(counts, dtype=np.uint32)
            _check(self._L.LBAudioDetectiveDatabaseGroupAddPacked(self.ref, _ptr(w), c.size, _ptr(c), 0, C.byref(first)), "DatabaseGroupAddPacked")
        return int(first.value)

    def add_fingerprint(self, fp):
        idx = C.c_uint64(0)
        _check(self._L.LBAudioDetectiveDatabaseGroupAddFingerprint(self.ref, fp.ref, C.byref(idx)), "DatabaseGroupAddFingerprint")
        return int(idx.value)

    def add_packed_device_to_shard(self, shard, d_words_ptr, n_clips, uniform_count, first_clip_index, producer_stream=None):
        _check(self._L.LBAudioDetectiveDatabaseGroupAddPackedDeviceToShard(self.ref, shard, d_words_ptr, n_clips, uniform_count, first_clip_index, producer_stream), "DatabaseGroupAddPackedDeviceToShard")

    def search_packed(self, qwords, k, rng=0):
        q = np.ascontiguousarray(qwords, dtype=np.uint32); n_q, cq = q.shape[0], q.shape[1]
        sc = np.zeros((n_q, k), np.float32); idx = np.zeros((n_q, k), np.uint32)
        _check(self._L.LBAudioDetectiveDatabaseGroupSearchPacked(self.ref, _ptr(q), n_q, cq, rng, k, _ptr(sc), _ptr(idx)), "DatabaseGroupSearchPacked")
        return sc, idx

    def search(self, fingerprints, k, rng=0):
        n = len(fingerprints); arr = (C.c_void_p * n)(*[f.ref for f in fingerprints])
        sc = np.zeros((n, k), np.float32); idx = np.zeros((n, k), np.uint32)
        _check(self._L.LBAudioDetectiveDatabaseGroupSearch(self.ref, arr, n, rng, k, _ptr(sc), _ptr(idx)), "DatabaseGroupSearch")
        return sc, idx

    @property
    def kernel_launches(self): return int(self._L.LBAudioDetectiveDatabaseGroupGetKernelLaunchCount(self.ref))

    @property
    def last_search_ms(self): return float(self._L.LBAudioDetectiveDatabaseGroupGetLastSearchMilliseconds(self.ref))


def merge_topk(scores, indices):
    """[lists][queries][k] -> merged [queries][k], ordered (score desc, index asc); runs the device merge kernel."""
    s = np.ascontiguousarray(scores, dtype=np.float32); i = np.ascontiguousarray(indices, dtype=np.uint32)
    n_lists, n_q, k = s.shape
    os_ = np.zeros((n_q, k), np.float32); oi = np.zeros((n_q, k), np.uint32)
    _check(lib().LBAudioDetectiveDatabaseMergeTopK(_ptr(s), _ptr(i), n_lists, n_q, k, _ptr(os_), _ptr(oi)), "DatabaseMergeTopK")
    return os_, oi


def merge_topk_device(d_scores_ptr, d_idx_ptr, n_lists, n_q, k, d_out_scores_ptr, d_out_idx_ptr, stream=None):
    _check(lib().LBAudioDetectiveDatabaseMergeTopKDevice(d_scores_ptr, d_idx_ptr, n_lists, n_q, k, d_out_scores_ptr, d_out_idx_ptr, stream), "DatabaseMergeTopKDevice")


def merge_topk_device_strided(d_scores_ptr, d_idx_ptr, n_lists, list_stride, n_q, k, d_out_scores_ptr, d_out_idx_ptr, stream=None):
    """Lists list_stride elements apart (e.g. an all-gather buffer of [scores | indices] payloads, merged in place)."""
    _check(lib().LBAudioDetectiveDatabaseMergeTopKDeviceStrided(d_scores_ptr, d_idx_ptr, n_lists, list_stride, n_q, k, d_out_scores_ptr, d_out_idx_ptr, stream), "DatabaseMergeTopKDeviceStrided")


def synthesize_device(d_out_ptr, n_clips, clip_len, clip_stride, first_clip_id=0, base_seed=0x1BAD5EED, sample_rate=5512.0, stream=None):
    _check(lib().LBAudioDetectiveSupportSynthesizeDevice(d_out_ptr, n_clips, clip_len, clip_stride, first_clip_id, base_seed, sample_rate, stream), "SupportSynthesizeDevice")


def random_codes_device(d_words_ptr, n_subfps, subfingerprint_length=200, seed=1, stream=None, first_subfp=0):
    """Random rank-sign codes; subfingerprint i of the call is subfingerprint first_subfp + i of the whole set (a function of seed and global index)."""
    _check(lib().LBAudioDetectiveSupportRandomCodesDeviceAt(d_words_ptr, n_subfps, subfingerprint_length, seed, first_subfp, stream), "SupportRandomCodesDeviceAt")


def device_count() -> int:
    return int(lib().LBAudioDetectiveSupportDeviceCount())


def microbench():
    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    _check(lib().LBAudioDetectiveSupportMicrobench(C.byref(a), C.byref(b), C.byref(c)), "SupportMicrobench")
    return {"fp32_tflops": a.value, "popc_gops": b.value, "lop3_gops": c.value}
