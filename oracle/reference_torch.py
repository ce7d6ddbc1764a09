"""Multi-threaded CPU restatement of the reference graph -- TEST / BASELINE INFRASTRUCTURE.

Same op list as oracle/reference_np.py (/root/reference/active_learning.py:239-263), executed
op by op with every intermediate materialised -- the way TensorFlow's executor runs the
reference's un-fused graph -- on torch CPU kernels with all host threads (the stand-in for
TF 1.13's multi-threaded Eigen kernels; TF itself is not installable here).  Used only by
bench.py's ``cpu_baseline`` / ``--impl reference`` legs and validated against the NumPy oracle
in tests/test_oracle_torch.py.  Never imported by the product package.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

EPSILON = float(np.finfo(np.float32).tiny)   # active_learning.py:40


def set_threads(n: int | None = None) -> int:
    n = n or len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (n or os.cpu_count())
    torch.set_num_threads(int(n))
    return torch.get_num_threads()


def pixel_confidence(logits: torch.Tensor, measure: str) -> torch.Tensor:
    """[N,H,W,C] or [T,N,H,W,C] float32 CPU tensor -> confidence map [N,H,W] (float32)."""
    x = logits
    C = x.shape[-1]
    if x.dim() == 5 and x.shape[0] == 1:
        x = x[0]
    if x.dim() == 4:
        prob = torch.softmax(x, dim=-1)                                  # :239
        return _measure(prob, measure, C)
    T = x.shape[0]
    mu = torch.zeros_like(x[0])
    m2 = torch.zeros(x.shape[1:-1], dtype=torch.float32)
    for t in range(T):                                                  # Welford (repo spec)
        p = torch.softmax(x[t], dim=-1)
        delta = p - mu
        mu = mu + delta / float(t + 1)
        m2 = m2 + (delta * (p - mu)).sum(dim=-1)
    if measure == "variance":
        return 1.0 - m2 / float(T)
    return _measure(mu, measure, C)


def _measure(prob: torch.Tensor, measure: str, C: int) -> torch.Tensor:
    if measure == "entropy":                                            # :243-251
        ent = -prob * torch.log(prob + EPSILON)
        ent = ent.sum(dim=-1)
        ent = ent / math.log(np.float32(C))
        return 1.0 - ent
    if measure == "margin":                                             # :254-255
        values, _ = torch.topk(prob, k=2, dim=-1)
        return values[..., 0] - values[..., 1]
    if measure == "confidence":                                         # :258
        return prob.max(dim=-1).values
    raise NotImplementedError("Uncertainty function not implemented.")  # :259-260


def final_head(features: torch.Tensor, kernel: torch.Tensor) -> torch.Tensor:
    """`Final.call` (models/enet/enet_modules.py:1359-1381): tf.nn.conv2d_transpose(features [N,h,w,16],
    kernel [3,3,C,16], output [N,2h,2w,C], strides 2, padding SAME).  torch's conv_transpose2d without padding
    yields the full (2h+1) x (2w+1) result; TF's SAME padding drops the last row and column.
    [T,N,h,w,16] (T Monte-Carlo forward passes, repo extension) runs the layer once per sample."""
    if features.dim() == 5:
        return torch.stack([final_head(features[t], kernel) for t in range(features.shape[0])])
    n, h, w, _ = features.shape
    wt = kernel.permute(3, 2, 0, 1).contiguous()                 # [in, out, kh, kw]
    full = torch.nn.functional.conv_transpose2d(features.permute(0, 3, 1, 2), wt, stride=2)
    return full[:, :, :2 * h, :2 * w].permute(0, 2, 3, 1).contiguous()


def score_pool(logits: torch.Tensor, measure: str) -> torch.Tensor:
    """Per-image f64 mean of the f32 map (:261-263)."""
    conf = pixel_confidence(logits, measure)
    return conf.to(torch.float64).mean(dim=(1, 2))


def rank_confidence(logits: torch.Tensor, unlabelled: np.ndarray, selection_size: int, measure: str,
                    batch_size: int = 8, head_kernel: torch.Tensor | None = None):
    """:682-715 with the TF graph replaced by the torch restatement; selection is verbatim NumPy.
    With ``head_kernel`` the first argument is the `Final` layer's input and the layer runs first (enet.py:367)."""
    n = logits.shape[-4]
    confidence = np.zeros(n, dtype=np.float32)
    for i in range(0, n, batch_size):
        xb = logits[i:i + batch_size] if logits.dim() == 4 else logits[:, i:i + batch_size]
        if head_kernel is not None:
            xb = final_head(xb, head_kernel)
        confidence[i:i + batch_size] = score_pool(xb, measure).numpy()
    unlabelled_confidence = confidence[unlabelled]
    selection_size = np.minimum(len(unlabelled), selection_size)
    if selection_size >= len(unlabelled):
        example_indices = np.arange(len(unlabelled))
    else:
        example_indices = np.argpartition(unlabelled_confidence, selection_size)[:selection_size]
    return unlabelled[example_indices], unlabelled_confidence
