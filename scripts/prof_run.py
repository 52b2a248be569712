"""Small driver for ncu captures: a few launches of the extraction kernel and of the search kernel on synthetic data."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lbaudiodetective_b200 as lb

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=2000)
ap.add_argument("--db-clips", type=int, default=200000)
ap.add_argument("--queries", type=int, default=1000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--what", default="both")
a = ap.parse_args()
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
if a.what in ("both", "extract"):
    d = lb.Detective(); n, L = a.clips, 165360
    x = torch.empty((n, L), dtype=torch.float32, device="cuda"); lb.synthesize_device(x.data_ptr(), n, L, L, stream=s.cuda_stream)
    w = torch.zeros((n, 19, 8), dtype=torch.int32, device="cuda")
    d.process_batch_device(x.data_ptr(), n, L, L, w.data_ptr(), s.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        d.process_batch_device(x.data_ptr(), n, L, L, w.data_ptr(), s.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    print("extract: %d clips, %.3f ms per pass, %.1f audio-hours/s" % (n, e0.elapsed_time(e1) / a.reps, n * L / 5512.0 / 3600.0 / (e0.elapsed_time(e1) / a.reps * 1e-3)))
if a.what in ("both", "search"):
    db = lb.Database(200); n = a.db_clips
    codes = torch.empty((n, 19, 8), dtype=torch.int32, device="cuda"); lb.random_codes_device(codes.data_ptr(), n * 19, 200, seed=5, stream=s.cuda_stream); torch.cuda.synchronize()
    db.add_packed_device(codes.data_ptr(), n, 19)
    q = codes[:a.queries, 3:9].contiguous(); sc = torch.empty((a.queries, 10), dtype=torch.float32, device="cuda"); ix = torch.empty((a.queries, 10), dtype=torch.int32, device="cuda")
    for _ in range(a.reps):
        db.search_device(q.data_ptr(), a.queries, 6, 10, sc.data_ptr(), ix.data_ptr(), stream=s.cuda_stream)
    torch.cuda.synchronize()
    assert (ix[:, 0].cpu().numpy() == np.arange(a.queries)).all()

if a.what == "resample":
    # recording-rate front end: 30 s clips at 44.1 kHz, resampled and fingerprinted on the device
    d = lb.Detective(); n, L = a.clips, 30 * 44100
    x = torch.empty((n, L), dtype=torch.float32, device="cuda"); lb.synthesize_device(x.data_ptr(), n, L, L, sample_rate=44100.0, stream=s.cuda_stream)
    w = torch.zeros((n, 19, 8), dtype=torch.int32, device="cuda")
    d.process_recorded_batch_device(x.data_ptr(), n, L, L, w.data_ptr(), s.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        d.process_recorded_batch_device(x.data_ptr(), n, L, L, w.data_ptr(), s.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    print("resample + extract: %d clips of 30 s at 44.1 kHz, %.3f ms per pass, %.1f audio-hours/s, %.0f GB/s of recorded PCM" % (n, ms, n * 30.0 / 3600.0 / (ms * 1e-3), n * L * 4 / (ms * 1e-3) / 1e9))
if a.what == "latency":
    import time
    from oracle.oracle import Port
    p = Port(); x = p.synth_clip(0, 55120); y = p.synth_clip(1, 55120)
    d = lb.Detective(); fa = d.process_pcm(x); fb = d.process_pcm(y)
    for name, fn in (("ProcessPCM(10 s clip)", lambda: d.process_pcm(x)), ("CompareToFingerprint(6 x 6)", lambda: fa.compare(fb, 200)),
                     ("ComparePCM(2 x 10 s)", lambda: d.compare_pcm(x, y, 0))):
        fn(); t0 = time.perf_counter()
        for _ in range(200): fn()
        print("%-30s %.1f us per call" % (name, (time.perf_counter() - t0) / 200 * 1e6))
    # what the two kernels themselves take for one 10 s clip (6 frames: the GPU is all but empty, so this is latency, not throughput)
    d.kernel_timing(enable=True, reset=True); d.kernel_timing(enable=True, reset=True, transform=True)
    for _ in range(50): d.process_pcm(x)
    n1, ms1 = d.kernel_timing(enable=False, reset=True); n2, ms2 = d.kernel_timing(enable=False, reset=True, transform=True)
    print("kernels for one 10 s clip: FFT + bands %.1f us, Haar/select %.1f us" % (1e3 * ms1 / max(n1, 1), 1e3 * ms2 / max(n2, 1)))
print("prof_run done")
