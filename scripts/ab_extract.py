"""A/B timing of extraction-kernel variants on one GPU: every variant is a set of environment variables read when a detective's plan is
built (LBAD_TRANSPOSE, LBAD_SUBFRAMES, ...) or another build of the library (LBAD_LIBRARY, one process per library).  Prints the device
time of the two kernels per 10,000-clip pass and checks that every variant produces the same words as the first one.

    python scripts/ab_extract.py "name1:VAR=a,VAR2=b" "name2:" ...
"""
import hashlib
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lbaudiodetective_b200 as lb

variants = sys.argv[1:] or ["base:"]
n, L = int(os.environ.get("AB_CLIPS", "10000")), 165360
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
x = torch.empty((n, L), dtype=torch.float32, device="cuda"); lb.synthesize_device(x.data_ptr(), n, L, L, stream=s.cuda_stream)
ref_hash = None
for v in variants:
    name, _, envs = v.partition(":")
    keys = []
    for kv in filter(None, envs.split(",")):
        k, _, val = kv.partition("="); os.environ[k] = val; keys.append(k)
    d = lb.Detective()
    w = torch.zeros((n, 19, 8), dtype=torch.int32, device="cuda")
    for _ in range(3):
        d.process_batch_device(x.data_ptr(), n, L, L, w.data_ptr(), s.cuda_stream)
    torch.cuda.synchronize()
    d.kernel_timing(enable=True, reset=True); d.kernel_timing(enable=True, reset=True, transform=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        d.process_batch_device(x.data_ptr(), n, L, L, w.data_ptr(), s.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    _, k1 = d.kernel_timing(enable=False, reset=True); _, k2 = d.kernel_timing(enable=False, reset=True, transform=True)
    h = hashlib.sha256(w.cpu().numpy().tobytes()).hexdigest()[:16]
    if ref_hash is None:
        ref_hash = h
    print("%-28s bands %.3f ms, haar/select %.3f ms, pass %.3f ms, words %s %s" % (name, k1 / reps, k2 / reps, e0.elapsed_time(e1) / reps, h, "(same)" if h == ref_hash else "(DIFFERENT)"), flush=True)
    for k in keys:
        del os.environ[k]
    del d, w
