"""Bring-up tool: per-tile timeline of one CTA of the fused-head kernel (needs a trace build:
   NVCC_EXTRA=-DALS_HEAD_TRACE python -m semanticsegmentationactivelearning_b200.build --force)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semanticsegmentationactivelearning_b200 import Scorer, _lib

Cc = int(sys.argv[1]) if len(sys.argv) > 1 else 19
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1
with Scorer(0) as sc:
    sc.prepare_head((0.4 * np.random.default_rng(0).standard_normal((3, 3, Cc, 16))).astype(np.float32))
    f = torch.randn((64, 256, 512, 16) if T == 1 else (T, 64 // T, 256, 512, 16), device="cuda")
    for _ in range(2):
        sc.score_features(f, "entropy" if T == 1 else "variance")
    torch.cuda.synchronize()
    lib = _lib.load()
    n = 1024
    buf = np.zeros((n, 16), np.int64)
    got = lib.als_debug_head_trace(C.c_void_p(buf.ctypes.data), n)
    assert got == n, got
t0 = buf[0, 7]
b = buf - t0
names = ["split0", "split1", "mma0", "mma1", "acc_full", "acc_rel", "epi_done", "copy", "psplit0", "psplit1", "pslot", "cslot", "w_prev", "w_cur", "w_acc", "w_begin"]
print("tile " + " ".join("%9s" % x for x in names))
for i in list(range(0, 6)) + list(range(200, 212)):
    print("%4d " % i + " ".join("%9d" % v for v in b[i]))
last = int(np.max(np.nonzero(buf[:, 2])[0])) if buf[:, 2].any() else 0
if last < 900:   # fewer tiles traced than the default window: use what there is
    buf = buf[:last + 1]
    print("(%d tiles traced)" % (last + 1))
d = np.diff(buf[100:900], axis=0)
print("steady-state period per tile (clks), median by column:", dict(zip(names, np.median(d, axis=0).astype(int).tolist())))
dur = {"split": np.median(buf[100:900, 1] - buf[100:900, 0]), "mma_issue": np.median(buf[100:900, 3] - buf[100:900, 2]),
       "mma_exec(issue_start->acc_full)": np.median(buf[100:900, 4] - buf[100:900, 2]),
       "epi_until_release": np.median(buf[100:900, 5] - buf[100:900, 4]), "epi_total": np.median(buf[100:900, 6] - buf[100:900, 4]),
       "copy->split_start": np.median(buf[100:900, 0] - buf[100:900, 7]), "split_end->mma_start": np.median(buf[100:900, 2] - buf[100:900, 1])}
if T > 1:
    w = buf[100:900]
    dur.update({"prev_split": np.median(w[:, 9] - w[:, 8]), "prev: start->slot free": np.median(w[:, 10] - w[:, 8]),
                "cur: start->slot free": np.median(w[:, 11] - w[:, 0]), "cur: slot free->done": np.median(w[:, 1] - w[:, 11]),
                "mma_start - max(split ends)": np.median(w[:, 2] - np.maximum(w[:, 1], w[:, 9])),
                "mma_start - prev epi release (2 back)": np.median(w[2:, 2] - w[:-2, 5]),
                "mma_start - prev epi release (3 back)": np.median(w[3:, 2] - w[:-3, 5])})
print({k: int(v) for k, v in dur.items()})
