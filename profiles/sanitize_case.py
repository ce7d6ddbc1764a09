"""Small pass over every kernel family for compute-sanitizer (one tool per gpurun call, see B200_PROFILING.md).
Usage: compute-sanitizer --tool memcheck python profiles/sanitize_case.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentationactivelearning_b200 import Scorer, rank_confidence

with Scorer(0) as sc:
    for (T, N, H, W, C, dtype, measure) in [
        (1, 5, 33, 47, 19, "float32", "entropy"),    # odd sizes: ragged last tile, image straddles
        (1, 3, 64, 64, 6, "float32", "margin"),      # several pixels per thread
        (1, 3, 40, 36, 66, "bfloat16", "confidence"),  # 2 lanes per pixel
        (3, 4, 32, 32, 19, "float32", "variance"),   # multi-sample ring
        (2, 2, 16, 24, 23, "float32", "entropy"),    # generic kernel (C not specialised)
        (1, 2, 16, 16, 150, "float32", "entropy"),   # 8 lanes per pixel
    ]:
        x = sc.synth_logits(T, 0, N, H, W, C, dtype=dtype)
        out = sc.pseudo_annotation(x, measure, 0.5)
        ids, u = rank_confidence(x, np.arange(N), 2, measure, scorer=sc)
        host = x.cpu()
        hs = sc.score(host.numpy() if dtype == "float32" else host.view(torch.uint16).numpy(), measure,
                      dtype=dtype)
        torch.cuda.synchronize()
        assert np.allclose(hs, out["pseudo_mean_confidence"].cpu().numpy(), rtol=0, atol=0), (hs, out)
print("sanitize case ok")
