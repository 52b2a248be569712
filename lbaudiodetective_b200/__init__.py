"""B200-native fingerprint path of LBAudioDetective behind the reference's C API.

The product is ``libLBAudioDetectiveCUDA.so`` (C-ABI, see ``include/``); this package is the thin ctypes mirror of
that API used by the tests and ``bench.py``.  There is no CPU fallback: every call that computes runs CUDA kernels.
"""
from .api import (Detective, Fingerprint, Frame, Database, DatabaseGroup, Stream, lib, load_library, device_available, merge_topk, merge_topk_device, merge_topk_device_strided, microbench,  # noqa: F401
                  synthesize_device, random_codes_device, device_count, pack_booleans, unpack_words, words_per_plane,
                  ARGUMENT_INVALID, DEVICE_UNAVAILABLE, DEVICE_ERROR, LBADError)
