"""Bring-up tool: per-tile timeline of one CTA of the fused-head kernel (needs a trace build:
   NVCC_EXTRA=-DALS_HEAD_TRACE python -m semanticsegmentationactivelearning_b200.build --force)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentationactivelearning_b200 import Scorer, _lib

Cc = int(sys.argv[1]) if len(sys.argv) > 1 else 19
with Scorer(0) as sc:
    sc.prepare_head((0.4 * np.random.default_rng(0).standard_normal((3, 3, Cc, 16))).astype(np.float32))
    f = torch.randn((64, 256, 512, 16), device="cuda")
    for _ in range(2):
        sc.score_features(f, "entropy")
    torch.cuda.synchronize()
    lib = _lib.load()
    n = 1024
    buf = np.zeros((n, 8), np.int64)
    got = lib.als_debug_head_trace(C.c_void_p(buf.ctypes.data), n)
    assert got == n, got
t0 = buf[0, 7]
b = buf - t0
names = ["split0", "split1", "mma0", "mma1", "acc_full", "acc_rel", "epi_done", "copy"]
print("tile " + " ".join("%9s" % x for x in names))
for i in list(range(0, 6)) + list(range(200, 212)):
    print("%4d " % i + " ".join("%9d" % v for v in b[i]))
d = np.diff(buf[100:900], axis=0)
print("steady-state period per tile (clks), median by column:", dict(zip(names, np.median(d, axis=0).astype(int).tolist())))
dur = {"split": np.median(buf[100:900, 1] - buf[100:900, 0]), "mma_issue": np.median(buf[100:900, 3] - buf[100:900, 2]),
       "mma_exec(issue_start->acc_full)": np.median(buf[100:900, 4] - buf[100:900, 2]),
       "epi_until_release": np.median(buf[100:900, 5] - buf[100:900, 4]), "epi_total": np.median(buf[100:900, 6] - buf[100:900, 4]),
       "copy->split_start": np.median(buf[100:900, 0] - buf[100:900, 7]), "split_end->mma_start": np.median(buf[100:900, 2] - buf[100:900, 1])}
print({k: int(v) for k, v in dur.items()})
