#!/usr/bin/env python
"""bench.py — headline benchmark of the fingerprint path (BASELINE.json: audio-hours/s fingerprinted, compares/s).

    python bench.py --gpus N --steps K --warmup W                 # the CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's own CPU code on the host cores

A step = one pass of fingerprint extraction over one batch of synthetic PCM: BASELINE config 2, 10,000 x 30 s clips
(83.3 audio-hours, 6.6 GB of float32) per GPU, resident in HBM when the timed region starts.  Prints ONE JSON line; besides the
headline it carries the other configs of BASELINE.json as legs of their own (config 3: 1,000 hours sharded over the GPUs; config 4:
the sharded database search; config 5: one-subfingerprint queries; config 1: one compare-audio call) and result hashes that must not
change with the number of GPUs (`extract.words_sha256`, `search.topk_sha256`).
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 5512.0
CLIP_LEN = 165360            # 30 s
SUBFPS = 19
ALGO_BYTES_PER_CLIP = CLIP_LEN * 4 + SUBFPS * 25        # 4 B per sample in + 25 B per subfingerprint out
FLOP_PER_WINDOW = 60300.0                               # 2.5 N log2 N + band stage
WINDOWS_PER_CLIP = SUBFPS * 128
HASH_BLOCKS = 8              # extraction hash: the first HASH_CLIPS clips of the 10,000-clip block of each of 8 ranks (global clip ids)
HASH_CLIPS = 8
DB_SEED = 1234               # the search database is a function of (DB_SEED, global subfingerprint index), whatever the sharding


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._t = [], set(), None, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml; self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8: "hw_slowdown",
                 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True); self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


CPU_FFT = "fast"             # oracle/shim/shim.c mode 2: the tuned single-precision FFT standing in for vDSP in every CPU timing


def cpu_extract_baseline(n_clips, threads, fft=CPU_FFT):
    """The reference's own code (oracle/_ref) — or the C port when it was never built — on a bounded sample of the workload."""
    from oracle import oracle as o
    chk = o.best()
    cfg = o.Cfg.default()
    pcm = np.stack([chk.synth_clip(c, CLIP_LEN) for c in range(min(n_clips, 64))])
    if n_clips > pcm.shape[0]:
        pcm = np.concatenate([pcm] * ((n_clips + pcm.shape[0] - 1) // pcm.shape[0]))[:n_clips]      # timing only: content repeats
    _, secs = chk.extract_batch(cfg, pcm, threads=threads, want_bits=False, fft_f32=fft if chk.kind == "reference" else True)
    hours = n_clips * CLIP_LEN / SR / 3600.0
    return chk, hours / secs, secs


def cpu_fft_note(chk):
    """What stands in for vDSP in the CPU timing, and what it costs per window (so that the reader can discount)."""
    if chk.kind != "reference":
        return {"fft": "float32 radix-2 Stockham of the C port"}
    return {"fft": "oracle/shim/shim.c mode 'fast': single-precision four-step FFT (32 x 32 points, SIMD clones AVX-512 / AVX2) behind vDSP_fft_zrip",
            "fft_us_per_2048_window (ctoz + fft_zrip + ztoc, one core)": round(min(chk.fft_us_per_window(2048, CPU_FFT, 20000) for _ in range(3)), 2),
            "round1_fft_us_per_2048_window (radix-2 Stockham stand-in, for comparison)": round(min(chk.fft_us_per_window(2048, "f32", 5000) for _ in range(2)), 2)}


def cpu_search_baseline(chk, threads, n_db=20000, n_q=4):
    rng = np.random.default_rng(1)
    def codes(n, c):
        s = rng.integers(0, 2, size=(n, c, 100)); out = np.zeros((n, c, 200), np.uint8); out[..., 0::2] = s == 0; out[..., 1::2] = s == 1
        return out
    _, secs = chk.search(codes(n_db, SUBFPS), codes(n_q, 6), 200, threads=threads)
    return n_db * n_q * 84 / secs


_REAL_STDOUT = None


def guard_stdout():
    """stdout carries ONE JSON line: anything a library prints there (NCCL's version banner, for one) goes to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the same path, same metric, all host threads.
    Each step fingerprints a bounded sample (args.ref_clips x 30 s) of the config-2 workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    times = []
    chk = None
    for i in range(args.warmup + args.steps):
        chk, _, secs = cpu_extract_baseline(args.ref_clips, threads)
        if i >= args.warmup:
            times.append(secs)
    total = float(np.sum(times)); hours = args.ref_clips * CLIP_LEN / SR / 3600.0
    value = hours * args.steps / total
    sample = "%d x 30 s clips (%.2f audio-hours) of the 10,000-clip workload per step; %d threads in one process" % (args.ref_clips, hours, threads)
    cpu = {"value": value, "unit": "audio-hours/s", "cores": threads, "kind": chk.kind, "sample": sample}
    cpu.update(cpu_fft_note(chk))
    line = {"impl": "reference", "metric": "audio-hours/s fingerprinted", "value": value, "unit": "audio-hours/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "configs[1] sample: batch fingerprint extraction of synthetic 30 s clips at reference defaults", "clips_per_step": args.ref_clips,
                                            "clip_seconds": 30, "window": 2048, "stride": 64, "bands": 32, "subfingerprint_length": 200},
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "audio-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--clips", type=int, default=10000, help="30 s clips per GPU per step (config 2: 10,000)")
    ap.add_argument("--ref-clips", type=int, default=384, help="clips per step of the CPU reference arm / cpu_baseline sample")
    ap.add_argument("--db-clips", type=int, default=1000000, help="database clips (whole job) for the search leg (config 4: 1M)")
    ap.add_argument("--queries", type=int, default=1000)
    ap.add_argument("--config3-clips", type=int, default=120000, help="30 s clips of the whole job in the config-3 leg (1,000 hours = 120,000), sharded over the GPUs")
    ap.add_argument("--no-search", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-config3", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--no-group", action="store_true", help="skip the one-process sharded search (LBAudioDetectiveDatabaseGroup) leg that rank 0 runs when N > 1")
    ap.add_argument("--microbench", action="store_true", help="(default; kept for old command lines) measure the FP32 / POPC / LOP3 pipe rates the rooflines are quoted against")
    ap.add_argument("--no-microbench", action="store_true", help="skip the pipe-rate measurement (a fraction of a second)")
    args = ap.parse_args()
    guard_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import lbaudiodetective_b200 as lb
    from lbaudiodetective_b200.dist import shard_range, ShardedTopK

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lb.load_library(build_if_missing=False)          # the bench must run the in-tree CUDA library, never a fallback
    if not lb.device_available():
        raise RuntimeError("no CUDA device: the fingerprint path has no CPU fallback")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_clips = args.clips
    det = lb.Detective()
    tstream = torch.cuda.Stream()                      # a real (non-NULL) stream: NULL would mean "the detective's own stream" to the C API
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    pcm = torch.empty((n_clips, CLIP_LEN), dtype=torch.float32, device="cuda")
    lb.synthesize_device(pcm.data_ptr(), n_clips, CLIP_LEN, CLIP_LEN, first_clip_id=rank * n_clips, stream=stream)      # every rank: its own clips, numbered globally
    words = torch.zeros((n_clips, SUBFPS, 8), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()

    def step():
        det.process_batch_device(pcm.data_ptr(), n_clips, CLIP_LEN, CLIP_LEN, words.data_ptr(), stream)

    for _ in range(args.warmup):
        step()
    barrier()
    det.kernel_timing(enable=True, reset=True); det.kernel_timing(enable=True, reset=True, transform=True)
    launches0 = det.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    n_timed, kernel_ms_total = det.kernel_timing(enable=False, reset=True)
    n_timed2, kernel2_ms_total = det.kernel_timing(enable=False, reset=True, transform=True)
    gpu_launches = det.kernel_launches - launches0
    hours_per_gpu_step = n_clips * CLIP_LEN / SR / 3600.0
    value = hours_per_gpu_step * world * args.steps / (ms_total * 1e-3)
    # per STEP, not per launch: a step of more than 2^18 frames runs the two kernels once per slab (config 3 at full size: 9 slabs)
    kernel_ms = kernel_ms_total / args.steps
    kernel2_ms = kernel2_ms_total / args.steps

    # ---- sanity on the timed output: structure of the words (never the thing measured) ----
    w = words[:: max(1, n_clips // 64)].cpu().numpy().view(np.uint32)
    assert ((w[..., :4] & w[..., 4:]) == 0).all() and (np.unpackbits((w[..., :4] | w[..., 4:]).view(np.uint8), axis=-1).sum(-1) == 100).all()

    # ---- result hash that must not depend on the number of GPUs: the packed words of global clips r*clips .. r*clips+7, r = 0..7.
    # Block r comes out of rank r's TIMED output when that rank exists; the blocks of ranks that do not exist in this run are
    # extracted by rank 0 from the same synthetic clips (outside the timed region), so the list is the same at N = 1, 2, 4, 8. ----
    my_block = sha(words[:HASH_CLIPS].cpu().numpy()) if n_clips >= HASH_CLIPS else None
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, my_block)
    else:
        gathered = [my_block]
    extract_hashes = None
    if rank == 0 and n_clips >= HASH_CLIPS:
        blocks = list(gathered[:HASH_BLOCKS])
        small = torch.empty((HASH_CLIPS, CLIP_LEN), dtype=torch.float32, device="cuda"); small_w = torch.zeros((HASH_CLIPS, SUBFPS, 8), dtype=torch.int32, device="cuda")
        for r in range(len(blocks), HASH_BLOCKS):
            lb.synthesize_device(small.data_ptr(), HASH_CLIPS, CLIP_LEN, CLIP_LEN, first_clip_id=r * n_clips, stream=stream)
            det.process_batch_device(small.data_ptr(), HASH_CLIPS, CLIP_LEN, CLIP_LEN, small_w.data_ptr(), stream); torch.cuda.synchronize()
            blocks.append(sha(small_w.cpu().numpy()))
        extract_hashes = {"words_sha256": hashlib.sha256("".join(blocks).encode()).hexdigest(), "blocks_from_timed_output": min(world, HASH_BLOCKS),
                          "what": "sha256 over the packed words of global clips r*%d .. r*%d+%d, r = 0..%d" % (n_clips, n_clips, HASH_CLIPS - 1, HASH_BLOCKS - 1)}
        del small, small_w

    # ---- roofline of the dominant kernel (the fused extraction kernel): FP32-pipe bound; HBM figures beside it ----
    hbm_peak, peak_src = measured_peaks()
    algo_bytes = n_clips * ALGO_BYTES_PER_CLIP
    hbm_achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    flops = n_clips * WINDOWS_PER_CLIP * FLOP_PER_WINDOW
    fp32_achieved = flops / (kernel_ms * 1e-3) / 1e12
    fp32_nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    traffic = None; traffic_src = None
    for name in ("r02_traffic.json", "r01_traffic.json"):      # DRAM bytes of the dominant kernel from the committed ncu capture, scaled to this launch's clip count
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            traffic = tj["traffic_bytes_per_launch"] * n_clips / tj["clips_per_launch"] / 1e9; traffic_src = "profiles/" + name + " (ncu --set full capture, not measured in this run)"
            break
        except Exception:
            pass
    mb = None
    if not args.no_microbench:
        mb = lb.microbench()          # every rank runs it (keeps the ranks in step); rank 0's numbers are reported
    fp32_peak = mb["fp32_tflops"] if mb else fp32_nominal
    roofline = {"bound": "fp32", "achieved": fp32_achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": fp32_achieved / fp32_peak,
                "peak_source": "measured in this run: dependent-FMA microbenchmark, LBAudioDetectiveSupportMicrobench" if mb else "nominal 148 SM x 128 lanes x 2 x 1.965 GHz",
                "traffic": traffic, "traffic_unit": "GB per launch (ncu dram read+write)", "traffic_source": traffic_src,
                "kernel": "bands_fused_kernel (framing + real FFT + band energies)", "kernel_ms": kernel_ms, "second_kernel": "haar_select32_kernel", "second_kernel_ms": kernel2_ms,
                "launches_per_step": n_timed // max(args.steps, 1), "algorithmic_flop_per_launch": flops, "algorithmic_bytes_per_launch": algo_bytes,
                "fp32_peak_tflops_nominal": fp32_nominal,
                "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak, "peak_source": peak_src,
                        "note": "not the binding roofline: 235 flop per new PCM byte against a ridge of about 11 flop/B (SURVEY.md §8d)"}}
    if mb:
        roofline["popc_gops_measured"] = mb["popc_gops"]; roofline["lop3_gops_measured"] = mb["lop3_gops"]

    # ---- e2e: the same pass through the host-buffer C-ABI call (H2D of the PCM and D2H of the words inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        host_pcm = torch.empty((n_clips, CLIP_LEN), dtype=torch.float32, pin_memory=True); host_pcm.copy_(pcm); torch.cuda.synchronize()
        host_words = torch.zeros((n_clips, SUBFPS, 8), dtype=torch.int32, pin_memory=True)
        # the ceiling first: the same pinned bytes through plain cudaMemcpyAsync on every rank at once, nothing else running
        scratch = torch.empty_like(pcm)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        scratch.copy_(host_pcm, non_blocking=True); barrier()
        c0.record()
        for _ in range(3):
            scratch.copy_(host_pcm, non_blocking=True)
        c1.record(); barrier()
        ceil_ms = max_over_ranks(c0.elapsed_time(c1)) / 3
        del scratch
        def e2e_step():
            det.process_batch_ptr(host_pcm.data_ptr(), n_clips, CLIP_LEN, CLIP_LEN, host_words.data_ptr())      # LBAudioDetectiveProcessPCMBatch: returns when the words are on the host
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        assert np.array_equal(host_words.numpy()[:: max(1, n_clips // 64)].view(np.uint32), w)
        bytes_in = n_clips * CLIP_LEN * 4
        e2e = {"value": hours_per_gpu_step * world * args.steps / dt, "unit": "audio-hours/s", "h2d_bytes_per_step": bytes_in,
               "d2h_bytes_per_step": n_clips * SUBFPS * 32, "ms_per_step": 1e3 * dt / args.steps, "api": "LBAudioDetectiveProcessPCMBatch (pinned host buffers)",
               "pcie_gbs": bytes_in / dt * args.steps / 1e9,
               "h2d_ceiling_gbs_per_gpu": bytes_in / (ceil_ms * 1e-3) / 1e9, "h2d_ceiling_gbs_aggregate": world * bytes_in / (ceil_ms * 1e-3) / 1e9,
               "h2d_ceiling_note": "the same pinned bytes copied by plain cudaMemcpyAsync on all %d rank(s) at once (max over ranks), nothing else running" % world,
               "frac_of_h2d_ceiling": (bytes_in / dt * args.steps) / (bytes_in / (ceil_ms * 1e-3))}
        # extra: the same clips as signed 16-bit PCM through LBAudioDetectiveProcessPCMBatchInt16 (half the PCIe bytes; not the headline)
        host_i16 = torch.empty((n_clips, CLIP_LEN), dtype=torch.int16, pin_memory=True)
        host_i16.copy_((pcm * 32767.0).round().clamp_(-32768, 32767).to(torch.int16)); torch.cuda.synchronize()
        det.process_batch_int16_ptr(host_i16.data_ptr(), n_clips, CLIP_LEN, CLIP_LEN, host_words.data_ptr())
        barrier(); t0 = time.perf_counter()
        for _ in range(args.steps):
            det.process_batch_int16_ptr(host_i16.data_ptr(), n_clips, CLIP_LEN, CLIP_LEN, host_words.data_ptr())
        torch.cuda.synchronize(); dt16 = max_over_ranks(time.perf_counter() - t0)
        e2e["int16_pcm"] = {"value": hours_per_gpu_step * world * args.steps / dt16, "unit": "audio-hours/s", "h2d_bytes_per_step": n_clips * CLIP_LEN * 2,
                            "ms_per_step": 1e3 * dt16 / args.steps, "api": "LBAudioDetectiveProcessPCMBatchInt16 (extension; pinned host buffers)"}
        del host_pcm, host_words, host_i16
    del pcm, words
    torch.cuda.empty_cache()

    # ---- config 3: 1,000 hours (120,000 x 30 s clips) sharded by clip over the GPUs, device-resident, no collective on the data path ----
    config3 = None
    if not args.no_config3:
        try:
            lo3, hi3 = shard_range(args.config3_clips, rank, world); n3 = hi3 - lo3
            pcm3 = torch.empty((n3, CLIP_LEN), dtype=torch.float32, device="cuda"); words3 = torch.zeros((n3, SUBFPS, 8), dtype=torch.int32, device="cuda")
            lb.synthesize_device(pcm3.data_ptr(), n3, CLIP_LEN, CLIP_LEN, first_clip_id=lo3, stream=stream)
            det.process_batch_device(pcm3.data_ptr(), n3, CLIP_LEN, CLIP_LEN, words3.data_ptr(), stream)       # warm-up (and the scratch slab)
            barrier()
            l0 = det.kernel_launches
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps3 = 2
            a0.record()
            for _ in range(steps3):
                det.process_batch_device(pcm3.data_ptr(), n3, CLIP_LEN, CLIP_LEN, words3.data_ptr(), stream)
            a1.record(); barrier()
            ms3 = max_over_ranks(a0.elapsed_time(a1)) / steps3
            hours3 = args.config3_clips * CLIP_LEN / SR / 3600.0
            # the first 8 global clips are the first 8 clips of the config-2 workload: their words must be the same here
            first_block = sha(words3[:HASH_CLIPS].cpu().numpy()) if rank == 0 else None
            config3 = {"workload": "configs[2]: %d x 30 s clips = %.0f audio-hours resident in HBM, sharded by clip over %d GPU(s) (%d clips = %.1f GB per GPU), no collective" %
                                   (args.config3_clips, hours3, world, n3, n3 * CLIP_LEN * 4 / 1e9),
                       "value": hours3 / (ms3 * 1e-3), "unit": "audio-hours/s", "ms_per_step": ms3, "steps": steps3, "scaling": "strong",
                       "gpu_launches_per_step": int((det.kernel_launches - l0) // steps3), "first_block_matches_config2": (first_block == gathered[0]) if rank == 0 else None}
            del pcm3, words3
        except Exception as ex:                                              # an allocation that does not fit must not take the headline with it
            config3 = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
        torch.cuda.empty_cache()

    # ---- search leg (config 4): args.queries 10 s queries vs a args.db_clips-clip database sharded over the ranks, top-10, one gather ----
    search = None
    if not args.no_search:
        lo, hi = shard_range(args.db_clips, rank, world)
        n_local = hi - lo
        db = lb.Database(200); db.set_clip_index_base(lo)
        codes = torch.empty((n_local, SUBFPS, 8), dtype=torch.int32, device="cuda")
        # the database is a function of the GLOBAL subfingerprint index: the same 1M clips however many ranks hold them
        lb.random_codes_device(codes.data_ptr(), n_local * SUBFPS, 200, seed=DB_SEED, stream=stream, first_subfp=lo * SUBFPS)
        db.add_packed_device(codes.data_ptr(), n_local, SUBFPS, producer_stream=stream)
        del codes
        g = torch.Generator(device="cpu"); g.manual_seed(7)
        # queries: 6-subfingerprint excerpts of database clips anywhere in the database, at a random offset — built on every rank from the generator
        q_src = torch.randint(0, args.db_clips, (args.queries,), generator=g)
        q_off = torch.randint(0, SUBFPS - 6 + 1, (args.queries,), generator=g)
        qw = torch.empty((args.queries, 6, 8), dtype=torch.int32, device="cuda")
        for i, (c, o_) in enumerate(zip(q_src.tolist(), q_off.tolist())):
            lb.random_codes_device(qw[i].data_ptr(), 6, 200, seed=DB_SEED, stream=stream, first_subfp=c * SUBFPS + o_)
        torch.cuda.synchronize()
        ex = ShardedTopK(args.queries, 10)
        host_res = torch.empty((2, args.queries, 10), dtype=torch.int32, pin_memory=True)
        def search_step():
            # per-GPU top-k on the shard written straight into the exchange payload, ONE NCCL all-gather, the merge kernel on the gather
            # buffer in place, then ONE copy of the merged result to the host
            db.search_device(qw.data_ptr(), args.queries, 6, 10, ex.scores_ptr, ex.indices_ptr, stream=stream)
            ex.gather_and_merge(stream)
            host_res.copy_(ex.merged, non_blocking=True)                       # scores and indices are one buffer: one copy
            torch.cuda.current_stream().synchronize()
            return host_res[0].numpy().view(np.float32), host_res[1].numpy().view(np.uint32)
        for _ in range(2):
            sc, idx = search_step()
        db.kernel_timing(enable=True, reset=True)
        barrier(); t0 = time.perf_counter()
        for _ in range(args.steps):
            sc, idx = search_step()
        barrier(); dt = max_over_ranks(time.perf_counter() - t0)
        n_k, k_ms = db.kernel_timing(enable=False, reset=True)
        assert np.array_equal(idx[:, 0], q_src.numpy().astype(np.uint32)) and (sc[:, 0] == 1.0).all()       # every query finds the clip it was cut from
        compares = args.queries * args.db_clips * 84
        search = {"metric": "Hamming compares/s", "value": compares * args.steps / dt, "unit": "compares/s", "ms_per_step": 1e3 * dt / args.steps,
                  "kernel_ms": k_ms / max(n_k, 1), "kernel_compares_per_s_per_gpu": (args.queries * n_local * 84) / (k_ms / max(n_k, 1) * 1e-3),
                  "exchange_ms (all-gather + merge + result to host)": 1e3 * dt / args.steps - k_ms / max(n_k, 1),
                  "queries": args.queries, "db_clips": args.db_clips, "k": 10, "scaling": "strong", "workload": "configs[3]: 1,000 x 6-subfp queries vs 1M x 19-subfp clips, 14 offsets",
                  "gpu_launches": db.kernel_launches,
                  "topk_sha256": sha(sc.copy(), idx.copy()), "topk_sha256_what": "sha256 over the merged [query][k] scores and global clip indices; the database and the queries are functions of global indices, so it must not change with the number of GPUs"}
        topk_ref = (sc.copy(), idx.copy())

    # ---- the same sharded search through the C-ABI alone: ONE process (rank 0) drives all the GPUs of the job (LBAudioDetectiveDatabaseGroup) ----
    if search is not None and world > 1 and not args.no_group:
        barrier()
        store = dist.distributed_c10d._get_default_store()      # the other ranks wait on the host (a NCCL barrier would spin on their GPUs, which rank 0 is about to use)
        if rank != 0:
            store.wait(["lbad_group_leg_done"])
        if rank == 0:
            try:
                group = lb.DatabaseGroup(200, list(range(world)))
                for s in range(world):
                    slo, shi = shard_range(args.db_clips, s, world)
                    with torch.cuda.device(s):
                        buf = torch.empty((shi - slo, SUBFPS, 8), dtype=torch.int32, device="cuda")
                        st = torch.cuda.current_stream().cuda_stream
                        lb.random_codes_device(buf.data_ptr(), (shi - slo) * SUBFPS, 200, seed=DB_SEED, stream=st, first_subfp=slo * SUBFPS)
                        group.add_packed_device_to_shard(s, buf.data_ptr(), shi - slo, SUBFPS, slo, producer_stream=st)
                        del buf
                q_host = qw.cpu().numpy().view(np.uint32)
                for _ in range(2):
                    g_sc, g_id = group.search_packed(q_host, 10)
                t0 = time.perf_counter(); dev_ms = []
                for _ in range(args.steps):
                    g_sc, g_id = group.search_packed(q_host, 10); dev_ms.append(group.last_search_ms)
                dtg = time.perf_counter() - t0
                search["group_api"] = {"api": "LBAudioDetectiveDatabaseGroupSearchPacked: one process, %d shards on %d GPUs, peer copies + device merge, no NCCL, host buffers in and out" % (world, world),
                                       "value": compares * args.steps / dtg, "unit": "compares/s", "ms_per_step": 1e3 * dtg / args.steps, "device_ms_per_step": float(np.mean(dev_ms)),
                                       "equals_torchrun_result": bool(np.array_equal(g_sc, topk_ref[0]) and np.array_equal(g_id, topk_ref[1])), "topk_sha256": sha(g_sc, g_id),
                                       "gpu_launches": group.kernel_launches}
                del group
            except Exception as ex_:
                search["group_api"] = {"error": "%s: %s" % (type(ex_).__name__, str(ex_)[:200])}
            store.set("lbad_group_leg_done", "1")
        barrier()

    # ---- extraction over all the GPUs of the job through the C-ABI alone: ONE process (rank 0), one detective per device
    #      (LBAudioDetectiveSetDevice), one call (LBAudioDetectiveProcessPCMBatchSharded), pinned host buffers in, words out ----
    sharded_api = None
    if world > 1 and not args.no_group and not args.no_e2e:
        barrier()
        store = dist.distributed_c10d._get_default_store()      # the other ranks wait on the host: rank 0 is about to use their GPUs
        if rank != 0:
            store.wait(["lbad_sharded_extract_done"])
        if rank == 0:
            try:
                per = min(5000, args.clips); n_all = per * world          # enough chunks (about 300 clips each) for the shares to follow the links' speeds
                gen = torch.empty((n_all, CLIP_LEN), dtype=torch.float32, device="cuda")
                lb.synthesize_device(gen.data_ptr(), n_all, CLIP_LEN, CLIP_LEN, first_clip_id=0, stream=stream)
                h_pcm = torch.empty((n_all, CLIP_LEN), dtype=torch.float32, pin_memory=True); h_pcm.copy_(gen)
                ref_words = torch.zeros((per, SUBFPS, 8), dtype=torch.int32, device="cuda")
                det.process_batch_device(gen.data_ptr(), per, CLIP_LEN, CLIP_LEN, ref_words.data_ptr(), stream); torch.cuda.synchronize()
                del gen; torch.cuda.empty_cache()
                h_words = torch.zeros((n_all, SUBFPS, 8), dtype=torch.int32, pin_memory=True)
                dets = []
                for dev in range(world):
                    dd = lb.Detective(); assert dd.set_device(dev) == 0; dets.append(dd)
                call = lambda: lb.Detective.process_batch_sharded(dets, None, out_words=h_words.data_ptr(), host_ptr=h_pcm.data_ptr(), n_clips=n_all, clip_len=CLIP_LEN, clip_stride=CLIP_LEN)
                call()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    call()
                dts = time.perf_counter() - t0
                same = bool(np.array_equal(h_words.numpy()[:per], ref_words.cpu().numpy()))
                sharded_api = {"api": "LBAudioDetectiveProcessPCMBatchSharded: one process, one detective per GPU (%d), one call; chunks taken from a shared cursor as each GPU's buffers drain; pinned host buffers in, words out" % world,
                               "clips": n_all, "value": n_all * CLIP_LEN / SR / 3600.0 * args.steps / dts, "unit": "audio-hours/s", "ms_per_step": 1e3 * dts / args.steps,
                               "pcie_gbs_aggregate": n_all * CLIP_LEN * 4 * args.steps / dts / 1e9, "first_shard_equals_one_detective": same,
                               "words_sha256": hashlib.sha256(h_words.numpy().tobytes()).hexdigest()}
                del dets, h_pcm, h_words, ref_words
            except Exception as ex_:
                sharded_api = {"error": "%s: %s" % (type(ex_).__name__, str(ex_)[:200])}
            store.set("lbad_sharded_extract_done", "1")
        barrier()

    if search is not None and world == 1:
        # the server-style call: ONE query against the whole database (lane-per-clip kernel + two-level merge), device-timed
        d_sc = torch.empty((args.queries, 10), dtype=torch.float32, device="cuda"); d_idx = torch.empty((args.queries, 10), dtype=torch.int32, device="cuda")
        for _ in range(2):
            db.search_device(qw.data_ptr(), 1, 6, 10, d_sc.data_ptr(), d_idx.data_ptr(), stream=stream)
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(10):
            db.search_device(qw.data_ptr(), 1, 6, 10, d_sc.data_ptr(), d_idx.data_ptr(), stream=stream)
        q1.record(); torch.cuda.synchronize()
        assert int(d_idx[0, 0].item()) == int(q_src[0]) and float(d_sc[0, 0].item()) == 1.0
        search["single_query_ms"] = q0.elapsed_time(q1) / 10
    if search is not None and rank == 0 and mb:
        # the search is POPC-bound — 4 POPC per compare (one per 32-pair word) against the measured lane-POPC rate;
        # HBM only sees the database once per 128 queries
        per_gpu = search["kernel_compares_per_s_per_gpu"]
        bound = mb["popc_gops"] * 1e9 / 4.0
        db_bytes = (args.db_clips / world) * SUBFPS * (32 + 8)
        search["roofline"] = {"bound": "int-pipe (POPC)", "achieved": per_gpu, "peak": bound, "unit": "compares/s per GPU", "frac": per_gpu / bound,
                              "note": "peak = measured lane-POPC rate / 4 POPC per compare (SURVEY.md §8d); the kernel issues 2 per compare for the reference's 100-rank codes (carry-save over three words, leftover bits in integer arithmetic), so frac exceeds 1 — frac_of_2_popc_bound is the fraction of what its own instruction mix allows",
                              "frac_of_3_popc_bound": per_gpu / (mb["popc_gops"] * 1e9 / 3.0), "frac_of_2_popc_bound": per_gpu / (mb["popc_gops"] * 1e9 / 2.0),
                              "hbm_gbs": db_bytes * ((args.queries + 127) // 128) / (search["kernel_ms"] * 1e-3) / 1e9, "hbm_frac": db_bytes * ((args.queries + 127) // 128) / (search["kernel_ms"] * 1e-3) / 1e9 / hbm_peak}
    if search is not None:
        del db
    if search is not None and world == 1 and rank == 0 and not args.no_config5:
        # codes with empty ranks (what digital silence produces) cannot take the short compare forms.  Regularity is decided per landed tile:
        # a database with one such clip in 1,000 stays near the full rate, one made of nothing else runs the general form throughout.
        irregular = {"workload": "1,000 x 6-subfp queries vs 250,000 x 19-subfp clips; 'silent' clips lose the sign bits of one subfingerprint"}
        n_i = 250000
        for label, every in (("regular", 0), ("one_silent_clip_in_1000", 1000), ("every_clip_silent (general form)", 1)):
            dbi = lb.Database(200)
            ci = torch.empty((n_i, SUBFPS, 8), dtype=torch.int32, device="cuda")
            lb.random_codes_device(ci.data_ptr(), n_i * SUBFPS, 200, seed=DB_SEED, stream=stream)
            qi = ci[:args.queries, 3:9].contiguous()
            if every:
                ci[args.queries + 1::every, SUBFPS // 2, :] = 0
            dbi.add_packed_device(ci.data_ptr(), n_i, SUBFPS, producer_stream=stream)
            si = torch.empty((args.queries, 10), dtype=torch.float32, device="cuda"); ii = torch.empty((args.queries, 10), dtype=torch.int32, device="cuda")
            for _ in range(2):
                dbi.search_device(qi.data_ptr(), args.queries, 6, 10, si.data_ptr(), ii.data_ptr(), stream=stream)
            dbi.kernel_timing(enable=True, reset=True)
            for _ in range(5):
                dbi.search_device(qi.data_ptr(), args.queries, 6, 10, si.data_ptr(), ii.data_ptr(), stream=stream)
            torch.cuda.synchronize()
            nki, msi = dbi.kernel_timing(enable=False, reset=True)
            assert (ii[:, 0].cpu() == torch.arange(args.queries, dtype=torch.int32)).all()
            rate = args.queries * n_i * 84 / (msi / max(nki, 1) * 1e-3)
            irregular[label] = {"kernel_ms": msi / max(nki, 1), "compares_per_s": rate, "frac_of_4_popc_bound": rate / (mb["popc_gops"] * 1e9 / 4.0) if mb else None}
            del dbi, ci, qi
        search["irregular_codes"] = irregular

    # ---- config 5 shape (rank 0, one GPU): 1,000 one-subfingerprint (3 s) queries against 100,000 clips of 5 subfingerprints (9 s), 5 offsets each ----
    config5 = None
    if rank == 0 and not args.no_config5:
        n5, c5, q5 = 100000, 5, 1000
        db5 = lb.Database(200)
        codes5 = torch.empty((n5, c5, 8), dtype=torch.int32, device="cuda")
        lb.random_codes_device(codes5.data_ptr(), n5 * c5, 200, seed=55, stream=stream)
        db5.add_packed_device(codes5.data_ptr(), n5, c5, producer_stream=stream)
        g5 = torch.Generator(device="cpu"); g5.manual_seed(5)
        src5 = torch.randint(0, n5, (q5,), generator=g5); off5 = torch.randint(0, c5, (q5,), generator=g5)
        qw5 = codes5[src5.cuda(), off5.cuda()].reshape(q5, 1, 8).contiguous()
        s5 = torch.empty((q5, 10), dtype=torch.float32, device="cuda"); i5 = torch.empty((q5, 10), dtype=torch.int32, device="cuda")
        for _ in range(3):
            db5.search_device(qw5.data_ptr(), q5, 1, 10, s5.data_ptr(), i5.data_ptr(), stream=stream)
        db5.kernel_timing(enable=True, reset=True)
        for _ in range(10):
            db5.search_device(qw5.data_ptr(), q5, 1, 10, s5.data_ptr(), i5.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        nk5, ms5 = db5.kernel_timing(enable=False, reset=True)
        assert (s5[:, 0] == 1.0).all()
        cmp5 = q5 * n5 * c5
        config5 = {"workload": "configs[4] shape: 1,000 one-subfingerprint queries (3 s) vs 100,000 clips of 5 subfingerprints (9 s), 5 offsets per pair, top-10, one GPU",
                   "kernel_ms": ms5 / max(nk5, 1), "value": cmp5 / (ms5 / max(nk5, 1) * 1e-3), "unit": "compares/s"}
        if mb:
            config5["roofline"] = {"bound": "int-pipe (POPC)", "achieved": config5["value"], "peak": mb["popc_gops"] * 1e9 / 4.0, "unit": "compares/s per GPU", "frac": config5["value"] / (mb["popc_gops"] * 1e9 / 4.0)}
        del db5, codes5

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        chk, v, secs = cpu_extract_baseline(args.ref_clips, threads)
        cpu = {"value": v, "unit": "audio-hours/s", "cores": threads, "kind": chk.kind,
               "sample": "%d x 30 s clips of the same synthetic workload (%.1f s of wall time), %d threads in one process" % (args.ref_clips, secs, threads),
               "search_compares_per_s": cpu_search_baseline(chk, threads)}
        cpu.update(cpu_fft_note(chk))
    # ---- configs[0] (the reference's own headline call): compare two 10 s clips through the compare-audio path, one call at a time ----
    # (part of the cpu_baseline leg, skipped with --no-cpu: the oracle generates the two clips and is the timed CPU baseline, nothing else)
    config1 = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle as o
        chk = o.best(); cfg = o.Cfg.default()
        a, b = chk.synth_clip(0, 55120), chk.synth_clip(1, 55120)
        det.compare_pcm(a, b, 0)
        t0 = time.perf_counter()
        for _ in range(200):
            got = det.compare_pcm(a, b, 0)
        gpu_us = (time.perf_counter() - t0) / 200 * 1e6
        t0 = time.perf_counter()
        for _ in range(3):
            want = chk.compare_pcm(cfg, a, b, 0)
        cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
        config1 = {"workload": "configs[0]: compare two synthetic 10 s clips through the compare-audio path (LBAudioDetectiveComparePCM), one call at a time, host buffers",
                   "us_per_call": gpu_us, "match": got, "reference_ms_per_call": cpu_ms, "reference_match": want, "reference_kind": chk.kind, "reference_cores": 1,
                   "reference_fft": "f64 definition (parity mode)"}
    if sharded_api is not None and e2e is not None:
        e2e["sharded_api"] = sharded_api
    line = {"metric": "audio-hours/s fingerprinted", "value": value, "unit": "audio-hours/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: batch fingerprint extraction of %d synthetic 30 s clips per GPU (FFT+band-energy kernel, then Haar+top-t+pack kernel)" % n_clips,
                       "clips_per_gpu": n_clips, "clip_seconds": 30, "window": 2048, "stride": 64, "bands": 32, "subfingerprint_length": 200,
                       "l2_policy": "inputs (%.1f GB per GPU) larger than L2" % (n_clips * CLIP_LEN * 4 / 1e9)},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu, "extract": extract_hashes,
            "config3": config3, "search": search, "config5": config5, "config1": config1}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
