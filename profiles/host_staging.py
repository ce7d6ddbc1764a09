"""How fast do HOST logits reach the scorer?  Pinned torch tensors (what bench.py's e2e feeds) against plain pageable
NumPy arrays (what the reference's sess.run hands out, /root/reference/active_learning.py:697-698)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from semanticsegmentationactivelearning_b200 import Scorer
sc = Scorer(0)
B, H, W, C = 8, 512, 1024, 19
x = sc.synth_logits(1, 0, B, H, W, C)
pinned = torch.empty((B, H, W, C), dtype=torch.float32).pin_memory(); pinned.copy_(x)
pageable = np.ascontiguousarray(pinned.numpy().copy())
nbytes = pageable.nbytes
idx = np.arange(B)
for name, src in (("pinned torch tensor", pinned), ("pageable numpy array", pageable)):
    sc.pool_begin(B)
    for _ in range(3):
        sc.pool_score_batch(src, idx, "entropy")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        sc.pool_score_batch(src, idx, "entropy")
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print("%-22s %.2f ms per batch of 8 (%.1f MB) = %.1f GB/s host->scorer" % (name, dt * 1e3, nbytes / 1e6, nbytes / dt / 1e9))
