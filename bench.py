#!/usr/bin/env python
"""bench.py — headline benchmark of the fingerprint path (BASELINE.json: audio-hours/s fingerprinted, compares/s).

    python bench.py --gpus N --steps K --warmup W                 # the CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's own CPU code on the host cores

A step = one pass of fingerprint extraction over one batch of synthetic PCM: BASELINE config 2, 10,000 x 30 s clips
(83.3 audio-hours, 6.6 GB of float32) per GPU, resident in HBM when the timed region starts.  Prints ONE JSON line.
"""
import argparse
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


def cpu_extract_baseline(n_clips, threads, fft_f32=True):
    """The reference's own code (oracle/_ref) — or the C port when it was never built — on a bounded sample of the workload."""
    from oracle import oracle as o
    chk = o.best()
    cfg = o.Cfg.default()
    pcm = np.stack([chk.synth_clip(c, CLIP_LEN) for c in range(min(n_clips, 64))])
    if n_clips > pcm.shape[0]:
        pcm = np.concatenate([pcm] * ((n_clips + pcm.shape[0] - 1) // pcm.shape[0]))[:n_clips]      # timing only: content repeats
    _, secs = chk.extract_batch(cfg, pcm, threads=threads, want_bits=False, fft_f32=fft_f32)
    hours = n_clips * CLIP_LEN / SR / 3600.0
    return chk, hours / secs, secs


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
    sample = "%d x 30 s clips (%.2f audio-hours) of the 10,000-clip workload per step; float32 FFT stands in for vDSP" % (args.ref_clips, hours)
    line = {"impl": "reference", "metric": "audio-hours/s fingerprinted", "value": value, "unit": "audio-hours/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "configs[1] sample: batch fingerprint extraction of synthetic 30 s clips at reference defaults", "clips_per_step": args.ref_clips,
                                            "clip_seconds": 30, "window": 2048, "stride": 64, "bands": 32, "subfingerprint_length": 200},
            "cpu_baseline": {"value": value, "unit": "audio-hours/s", "cores": threads, "kind": chk.kind, "sample": sample},
            "e2e": {"value": value, "unit": "audio-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


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
    ap.add_argument("--no-search", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--microbench", action="store_true", help="(default; kept for old command lines) measure the FP32 / POPC / LOP3 pipe rates the rooflines are quoted against")
    ap.add_argument("--no-microbench", action="store_true", help="skip the pipe-rate measurement (a fraction of a second)")
    args = ap.parse_args()
    guard_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import lbaudiodetective_b200 as lb
    from lbaudiodetective_b200.dist import shard_range, gather_and_merge_topk_device

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    try:        # keep the pinned staging buffers of this rank on the NUMA node of its GPU (matters once 8 ranks upload at once)
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1}
        if cpus:
            os.sched_setaffinity(0, cpus & set(os.sched_getaffinity(0)) or cpus)
    except Exception:
        pass
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
    lb.synthesize_device(pcm.data_ptr(), n_clips, CLIP_LEN, CLIP_LEN, first_clip_id=rank * n_clips, stream=stream)      # every rank: its own clips
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

    # ---- roofline of the dominant kernel (the fused extraction kernel) ----
    hbm_peak, peak_src = measured_peaks()
    algo_bytes = n_clips * ALGO_BYTES_PER_CLIP
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    flops = n_clips * WINDOWS_PER_CLIP * FLOP_PER_WINDOW
    traffic = None
    try:        # DRAM bytes of the dominant kernel from the committed ncu capture, scaled to this launch's clip count
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        traffic = tj["traffic_bytes_per_launch"] * n_clips / tj["clips_per_launch"] / 1e9
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "traffic_unit": "GB per launch (ncu dram read+write)",
                "peak_source": peak_src, "kernel": "bands_fused_kernel (FFT + band energies)", "kernel_ms": kernel_ms, "second_kernel": "haar_select32_kernel", "second_kernel_ms": kernel2_ms, "launches_per_step": n_timed // max(args.steps, 1), "algorithmic_bytes_per_launch": algo_bytes,
                "note": "FP32-issue bound, not HBM bound (235 flop per new PCM byte, SURVEY.md §8d); fp32 figures alongside",
                "fp32_tflops_algorithmic": flops / (kernel_ms * 1e-3) / 1e12, "fp32_peak_tflops_nominal": 148 * 128 * 2 * 1.965e9 / 1e12}
    if not args.no_microbench and rank == 0:
        mb = lb.microbench(); roofline["fp32_peak_tflops_measured"] = mb["fp32_tflops"]; roofline["fp32_frac_of_measured"] = roofline["fp32_tflops_algorithmic"] / mb["fp32_tflops"]
        roofline["popc_gops_measured"] = mb["popc_gops"]; roofline["lop3_gops_measured"] = mb["lop3_gops"]

    # ---- e2e: the same pass through the host-buffer C-ABI call (H2D of the PCM and D2H of the words inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        host_pcm = torch.empty((n_clips, CLIP_LEN), dtype=torch.float32, pin_memory=True); host_pcm.copy_(pcm); torch.cuda.synchronize()
        host_words = torch.zeros((n_clips, SUBFPS, 8), dtype=torch.int32, pin_memory=True)
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
        e2e = {"value": hours_per_gpu_step * world * args.steps / dt, "unit": "audio-hours/s", "h2d_bytes_per_step": n_clips * CLIP_LEN * 4,
               "d2h_bytes_per_step": n_clips * SUBFPS * 32, "ms_per_step": 1e3 * dt / args.steps, "api": "LBAudioDetectiveProcessPCMBatch (pinned host buffers)"}
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
        e2e["pcie_gbs"] = n_clips * CLIP_LEN * 4 / dt * args.steps / 1e9
        del host_pcm, host_words, host_i16
    del pcm
    torch.cuda.empty_cache()

    # ---- search leg (config 4): args.queries 10 s queries vs a args.db_clips-clip database sharded over the ranks, top-10, one gather ----
    search = None
    if not args.no_search:
        lo, hi = shard_range(args.db_clips, rank, world)
        n_local = hi - lo
        db = lb.Database(200); db.set_clip_index_base(lo)
        codes = torch.empty((n_local, SUBFPS, 8), dtype=torch.int32, device="cuda")
        lb.random_codes_device(codes.data_ptr(), n_local * SUBFPS, 200, seed=1234 + rank, stream=stream); torch.cuda.synchronize()
        db.add_packed_device(codes.data_ptr(), n_local, SUBFPS)
        g = torch.Generator(device="cpu"); g.manual_seed(7)
        # queries: 6-subfingerprint excerpts of database clips of rank 0's shard (so every rank can build the same batch), at a random offset
        q_src = torch.randint(0, max(1, args.db_clips // world), (args.queries,), generator=g)
        q_off = torch.randint(0, SUBFPS - 6 + 1, (args.queries,), generator=g)
        if rank == 0:
            qw = torch.stack([codes[int(c), int(o):int(o) + 6] for c, o in zip(q_src, q_off)]).contiguous()
        else:
            qw = torch.empty((args.queries, 6, 8), dtype=torch.int32, device="cuda")
        if world > 1:
            dist.broadcast(qw, src=0)
        del codes
        d_sc = torch.empty((args.queries, 10), dtype=torch.float32, device="cuda"); d_idx = torch.empty((args.queries, 10), dtype=torch.int32, device="cuda")
        def search_step():
            # per-GPU top-k on the shard, ONE NCCL all-gather of the [query][k] lists, device merge, then the result to the host
            db.search_device(qw.data_ptr(), args.queries, 6, 10, d_sc.data_ptr(), d_idx.data_ptr(), stream=stream)
            m_sc, m_id = gather_and_merge_topk_device(d_sc, d_idx, stream)
            return m_sc.cpu().numpy(), m_id.cpu().numpy().view(np.uint32)
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
                  "queries": args.queries, "db_clips": args.db_clips, "k": 10, "scaling": "strong", "workload": "configs[3]: 1,000 x 6-subfp queries vs 1M x 19-subfp clips, 14 offsets",
                  "gpu_launches": db.kernel_launches}

    if search is not None and world == 1:
        # the server-style call: ONE query against the whole database (lane-per-clip kernel + two-level merge), device-timed
        for _ in range(2):
            db.search_device(qw.data_ptr(), 1, 6, 10, d_sc.data_ptr(), d_idx.data_ptr(), stream=stream)
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(10):
            db.search_device(qw.data_ptr(), 1, 6, 10, d_sc.data_ptr(), d_idx.data_ptr(), stream=stream)
        q1.record(); torch.cuda.synchronize()
        assert int(d_idx[0, 0].item()) == int(q_src[0]) and float(d_sc[0, 0].item()) == 1.0
        search["single_query_ms"] = q0.elapsed_time(q1) / 10
    if search is not None and rank == 0 and "popc_gops_measured" in roofline:
        # the search is POPC-bound — 4 POPC per compare (one per 32-pair word) against the measured lane-POPC rate;
        # HBM only sees the database once per 128 queries
        per_gpu = search["kernel_compares_per_s_per_gpu"]
        bound = roofline["popc_gops_measured"] * 1e9 / 4.0
        db_bytes = (args.db_clips / world) * SUBFPS * (32 + 8)
        search["roofline"] = {"bound": "int-pipe (POPC)", "achieved": per_gpu, "peak": bound, "unit": "compares/s per GPU", "frac": per_gpu / bound,
                              "note": "peak = measured lane-POPC rate / 4 POPC per compare; the kernel needs 3 (carry-save adder), so frac can exceed 1",
                              "hbm_gbs": db_bytes * ((args.queries + 127) // 128) / (search["kernel_ms"] * 1e-3) / 1e9, "hbm_frac": db_bytes * ((args.queries + 127) // 128) / (search["kernel_ms"] * 1e-3) / 1e9 / hbm_peak}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        chk, v, secs = cpu_extract_baseline(args.ref_clips, threads)
        cpu = {"value": v, "unit": "audio-hours/s", "cores": threads, "kind": chk.kind,
               "sample": "%d x 30 s clips of the same synthetic workload (%.1f s of wall time); float32 FFT stands in for vDSP" % (args.ref_clips, secs),
               "search_compares_per_s": cpu_search_baseline(chk, threads)}
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
                   "us_per_call": gpu_us, "match": got, "reference_ms_per_call": cpu_ms, "reference_match": want, "reference_kind": chk.kind, "reference_cores": 1}
    line = {"metric": "audio-hours/s fingerprinted", "value": value, "unit": "audio-hours/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: batch fingerprint extraction of %d synthetic 30 s clips per GPU (FFT+band-energy kernel, then Haar+top-t+pack kernel)" % n_clips,
                       "clips_per_gpu": n_clips, "clip_seconds": 30, "window": 2048, "stride": 64, "bands": 32, "subfingerprint_length": 200,
                       "l2_policy": "inputs (%.1f GB per GPU) larger than L2" % (n_clips * CLIP_LEN * 4 / 1e9)},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu, "search": search, "config1": config1}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
