"""Thread-parallel driver for the NumPy oracle (test infrastructure).

oracle/reference_np.py is single-threaded NumPy; at BASELINE geometry (512x1024 and 1024x2048 images, up to
T = 16 samples of 66 classes) one image is 10^7..10^9 elements.  The per-pixel graph is independent per pixel, so the
map is computed in row blocks on a thread pool (NumPy releases the GIL inside ufuncs) -- the same functions, the same
arithmetic per pixel, only faster wall-clock."""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def _workers() -> int:
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(1, min(32, n))


def pixel_confidence(x: np.ndarray, measure: str, rows_per_block: int = 32) -> np.ndarray:
    """oracle.reference_np.pixel_confidence over [N,H,W,C] / [T,N,H,W,C], row blocks in parallel -> f32 [N,H,W]."""
    from oracle import reference_np as R
    x = np.asarray(x)
    five = x.ndim == 5
    N, H, W = x.shape[-4], x.shape[-3], x.shape[-2]
    out = np.empty((N, H, W), np.float32)
    jobs = [(n, r) for n in range(N) for r in range(0, H, rows_per_block)]

    def run(job):
        n, r = job
        blk = x[:, n:n + 1, r:r + rows_per_block] if five else x[n:n + 1, r:r + rows_per_block]
        out[n, r:r + rows_per_block] = R.pixel_confidence(np.ascontiguousarray(blk), measure)[0]

    with ThreadPoolExecutor(_workers()) as ex:
        list(ex.map(run, jobs))
    return out


def pseudo_label(x: np.ndarray) -> np.ndarray:
    from oracle import reference_np as R
    return R.pseudo_label(x)
