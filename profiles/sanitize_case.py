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
    # round 2: streamed Monte-Carlo accumulation (state in HBM), one-call pool pass, selection with a ragged unlabelled
    # set and k > 1024 (second ranking kernel), fused classifier head on an odd-sized feature map, world-1 NCCL exchange
    x = sc.synth_logits(3, 0, 4, 33, 31, 19)
    sc.mc_begin((4, 33, 31, 19))
    for t in range(3):
        sc.mc_add_sample(x[t].contiguous())
    s = sc.mc_finish("variance", batch_indices=np.arange(4))
    torch.cuda.synchronize()
    ids, u = sc.rank_pool(x, np.array([3, 0, 2]), 2, "variance")
    sc.pool_begin(3000)
    ids, u = sc.pool_select(np.arange(0, 3000, 2), 1400)
    assert len(ids) == 1400
    rng = np.random.default_rng(0)
    feat = torch.from_numpy(rng.standard_normal((2, 9, 131, 16)).astype(np.float32)).cuda()
    sc.prepare_head((0.4 * rng.standard_normal((3, 3, 19, 16))).astype(np.float32))
    sc.score_features(feat, "entropy")
    sc.score_features(torch.stack([feat, 0.9 * feat]), "variance")
    Scorer.comm_init_all([sc])
    sc.pool_begin(8)
    ids, u = sc.pool_select_global(np.arange(8), 3, (0, 8))
    torch.cuda.synchronize()
print("sanitize case ok")
