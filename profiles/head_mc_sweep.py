"""Fused-head kernel time per accumulator tile over class count C and sample count T (bring-up aid).

    python profiles/head_mc_sweep.py [N h w]

Prints, per (C, T): ms per launch, G(sample-pixels)/s and SM clocks per (tile, sample) at the current SM clock --
which role bounds the kernel shows in how the time moves with C (epilogue) or not (loaders / MMA)."""
import os
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semanticsegmentationactivelearning_b200 import Scorer  # noqa: E402

N, h, w = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (48, 256, 512)
cases = [(c, t) for c in (2, 6, 12, 19, 24) for t in (1, 2, 8)]
if os.environ.get("SWEEP_CASES"):
    cases = [tuple(int(v) for v in s.split("x")) for s in os.environ["SWEEP_CASES"].split(",")]
sc = Scorer(0)
gen = torch.Generator(device="cuda").manual_seed(1)
base = torch.randn((8, N, h, w, 16), generator=gen, device="cuda")
for C, T in cases:
    kern = (0.4 * np.random.default_rng(C).standard_normal((3, 3, C, 16))).astype(np.float32)
    sc.prepare_head(kern)
    f = base[0] if T == 1 else base[:T].contiguous()
    out = torch.empty(N, dtype=torch.float64, device="cuda")
    measure = "entropy" if T == 1 else "variance"
    for _ in range(3):
        sc.score_features(f, measure, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        sc.score_features(f, measure, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-i", "0"],
                               capture_output=True, text=True).stdout.strip() or 0)
    tiles = N * h * ((w + 127) // 128) * T               # accumulators in the launch
    per_sm = tiles / 148.0
    print("C=%2d T=%d: %.3f ms  %.1f Gsample-pix/s  %.1f Gpix/s  ~%.0f clk per (tile,sample) at %.0f MHz"
          % (C, T, ms, 4 * N * h * w * T / ms / 1e6, 4 * N * h * w / ms / 1e6, ms * 1e-3 * mhz * 1e6 / per_sm, mhz), flush=True)
