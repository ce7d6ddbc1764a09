"""Where the time of ONE small pool pass goes (BASELINE config 1: 64 images @512x1024, C=19 -- a single scoring launch).

Host clock around every API call of a pass (pool_begin / pool_score_batch / pool_select), CUDA events around the scoring
launch and around the whole pass, and a tiny-pool pass (8 images of 16x16) that is pure call overhead.  Usage (GPU box):
    python profiles/pass_breakdown.py [bf16]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from semanticsegmentationactivelearning_b200 import Scorer  # noqa: E402

dtype = "bfloat16" if "bf16" in sys.argv else "float32"
sc = Scorer(0)


def one(N, H, W, C, reps=200):
    x = sc.synth_logits(1, 0, N, H, W, C, dtype=dtype)
    idx = np.arange(N, dtype=np.int64)
    unl = np.arange(N, dtype=np.int64)
    t = {"begin": 0.0, "score": 0.0, "select": 0.0}
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for _ in range(10):
        sc.pool_begin(N); sc.pool_score_batch(x, idx, "entropy"); sc.pool_select(unl, 50)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for r in range(reps):
        a = time.perf_counter()
        sc.pool_begin(N)
        b = time.perf_counter()
        ev[r][0].record()
        sc.pool_score_batch(x, idx, "entropy")
        ev[r][1].record()
        c = time.perf_counter()
        sc.pool_select(unl, 50)
        d = time.perf_counter()
        t["begin"] += b - a; t["score"] += c - b; t["select"] += d - c
    w1 = time.perf_counter()
    torch.cuda.synchronize()
    k = sum(e0.elapsed_time(e1) for e0, e1 in ev) / reps * 1e3
    per = (w1 - w0) / reps * 1e6
    print("%s %dx%dx%dx%d: pass %.1f us wall | scoring launch (events) %.1f us | host: pool_begin %.1f, pool_score_batch %.1f "
          "(issue only), pool_select %.1f (issue + wait for the GPU + unpack) | pass - scoring launch = %.1f us"
          % (dtype, N, H, W, C, per, k, t["begin"] / reps * 1e6, t["score"] / reps * 1e6, t["select"] / reps * 1e6, per - k))


one(8, 16, 16, 19)        # nothing to score: the floor of a pass (launch latencies, one synchronisation, Python)
one(64, 512, 1024, 19)    # BASELINE config 1
one(64, 480, 640, 6)
