"""Which per-pixel output costs what on a batch-of-8 call (training-path call site, /root/reference/active_learning.py:229-275)?"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from semanticsegmentationactivelearning_b200 import Scorer
sc = Scorer(0)
B, H, W, C = 8, 512, 1024, 19
bufs = [sc.synth_logits(1, 8 * i, B, H, W, C) for i in range(16)]     # 5.1 GB: every call reads from HBM
P = B * H * W
for name, kw in (("scores only", None), ("conf", dict(want_label=False, want_mask=False)), ("conf+mask", dict(want_label=False, want_mask=True)),
                 ("conf+label", dict(want_label=True, want_mask=False)), ("conf+label+mask", dict(want_label=True, want_mask=True))):
    outs = [None] * len(bufs)
    def call(i):
        if kw is None:
            return sc.score(bufs[i], "entropy")
        outs[i] = sc.pseudo_annotation(bufs[i], "entropy", 0.9, out=outs[i], **kw)
    for i in range(len(bufs)): call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        for i in range(len(bufs)): call(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * len(bufs))
    wb = 0 if kw is None else 4 + (1 if kw["want_label"] else 0) + (1 if kw["want_mask"] else 0)
    print("%-16s %.1f us per call of 8 images  (%.0f GB/s incl. %d B/pixel written)" % (name, us, P * (C * 4 + wb) / us / 1e3, wb))
