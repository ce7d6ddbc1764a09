"""CPU oracle for the pool-scoring hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
(``semanticsegmentationactivelearning_b200``) never does; it fails loudly when
its CUDA library is missing.

What this file restates (all ``file:line`` relative to /root/reference):

* ``active_learning.py:234-236``  pseudo_label  = u8(argmax_c logits)
* ``active_learning.py:239``      pseudo_prob   = tf.nn.softmax(logits, axis=-1)
* ``active_learning.py:240-251``  entropy confidence
* ``active_learning.py:252-255``  margin confidence (top_k, k=2)
* ``active_learning.py:256-258``  max-probability confidence
* ``active_learning.py:259-260``  NotImplementedError for unknown measures
* ``active_learning.py:261-263``  per-image score = f64 mean over (H, W)
* ``active_learning.py:265-269``  pseudo_mask = conf < threshold ? 0 : 1
* ``active_learning.py:682-715``  rank_confidence(): scatter scores by example
  index into an f32 array, filter to ``unlabelled``, np.argpartition k smallest

The arithmetic of those lines lives in TensorFlow 1.13.2 (``requirements.txt:1``),
which is not vendored and cannot be installed here.  Every op is therefore
restated with NumPy in float32, one materialised intermediate per TF op, exactly
in graph order.  PARITY PINNING: the *composition* (op order, EPSILON, log base,
f64 cast, axes, selection code) is pinned against the reference's own source
lines, executed verbatim over a NumPy stand-in for the ``tf`` namespace by
``tests/golden/make_golden.py`` (fixtures in ``tests/golden/*.npz``).  The
*primitive kernels* (TF's exp/log/softmax implementations) are unpinned -- no TF
binary is available -- so agreement is asserted to the north-star tolerance
(1e-5 relative in fp32), not bit-for-bit at the per-pixel level.

The MC-dropout variance measure does not exist in the reference
(SURVEY.md section 8(a) row a13); its definition below is this repo's frozen spec.
"""
from __future__ import annotations

import numpy as np

# active_learning.py:40
EPSILON = np.finfo(np.float32).tiny

MEASURES = ("entropy", "margin", "confidence", "variance")


# --------------------------------------------------------------------------- #
# Primitive ops (TensorFlow semantics, fp32)                                  #
# --------------------------------------------------------------------------- #
def softmax(logits: np.ndarray, flavour: str = "gpu") -> np.ndarray:
    """tf.nn.softmax(logits, axis=-1) in float32 -- active_learning.py:239.

    flavour "gpu": e / sum   (TF 1.13 CUDA kernel divides)
    flavour "cpu": e * (1/sum) (TF 1.13 Eigen kernel multiplies by the inverse)
    """
    x = np.asarray(logits, dtype=np.float32)
    m = np.max(x, axis=-1, keepdims=True)
    with np.errstate(invalid="ignore"):
        e = np.exp(x - m, dtype=np.float32)
    s = np.sum(e, axis=-1, keepdims=True, dtype=np.float32)
    if flavour == "gpu":
        return (e / s).astype(np.float32)
    if flavour == "cpu":
        return (e * (np.float32(1.0) / s)).astype(np.float32)
    raise ValueError(flavour)


def confidence_from_prob(prob: np.ndarray, measure: str, num_classes: int | None = None) -> np.ndarray:
    """Per-pixel confidence from a probability map -- active_learning.py:240-260."""
    p = np.asarray(prob, dtype=np.float32)
    if measure == "entropy":
        # :243  entropy = - pseudo_prob * tf.math.log(pseudo_prob + EPSILON)
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (p + EPSILON).astype(np.float32)
            l = np.log(t, dtype=np.float32)
            q = (-p) * l
        # :244  reduce_sum over classes
        h = np.sum(q, axis=-1, dtype=np.float32)
        # :248  log_base = tf.math.log(np.float32(dataset.num_classes))
        c = p.shape[-1] if num_classes is None else num_classes
        log_base = np.log(np.float32(c), dtype=np.float32)
        # :249  entropy / log_base ;  :251  1.0 - entropy
        h = (h / log_base).astype(np.float32)
        return (np.float32(1.0) - h).astype(np.float32)
    if measure == "margin":
        # :254-255  values, _ = tf.math.top_k(p, k=2); values[...,0] - values[...,1]
        if p.shape[-1] < 2:
            raise ValueError("margin needs at least two classes")
        part = np.partition(p, p.shape[-1] - 2, axis=-1)
        v0 = part[..., -1]
        v1 = part[..., -2]
        # NaN-propagating like a sorted top_k of a NaN row would be irrelevant:
        # NaN rows are handled by the caller (score becomes NaN).
        return (v0 - v1).astype(np.float32)
    if measure == "confidence":
        # :258  tf.math.reduce_max(pseudo_prob, axis=-1)
        return np.max(p, axis=-1).astype(np.float32)
    # :259-260
    raise NotImplementedError("Uncertainty function not implemented.")


def welford_mean_m2(probs: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """fp32 Welford recurrence over the leading (sample) axis.

    probs [T, ..., C] -> (mean [..., C], m2sum [...]) where m2sum = sum_c M2_c.
    Repo spec (no reference counterpart, SURVEY.md section 8(c) "variance extension").
    """
    p = np.asarray(probs, dtype=np.float32)
    mu = np.zeros(p.shape[1:], dtype=np.float32)
    m2 = np.zeros(p.shape[1:-1], dtype=np.float32)
    for t in range(p.shape[0]):
        delta = (p[t] - mu).astype(np.float32)
        mu = (mu + delta / np.float32(t + 1)).astype(np.float32)
        m2 = (m2 + np.sum(delta * (p[t] - mu), axis=-1, dtype=np.float32)).astype(np.float32)
    return mu, m2


def pixel_confidence(logits: np.ndarray, measure: str, flavour: str = "gpu") -> np.ndarray:
    """logits [N,H,W,C] or [T,N,H,W,C] -> confidence map f32 [N,H,W].

    T == 1 / 4-D input: exactly the reference graph (active_learning.py:239-258).
    5-D input with T > 1 (extension): softmax per sample, Welford mean over T,
    then the reference measure applied to the predictive mean; "variance" gives
    1 - sum_c var_c with population variance (ddof = 0).
    """
    x = np.asarray(logits, dtype=np.float32)
    if measure not in MEASURES:
        raise NotImplementedError("Uncertainty function not implemented.")
    if x.ndim == 4:
        if measure == "variance":
            raise ValueError("variance needs T >= 2 Monte-Carlo samples")
        return confidence_from_prob(softmax(x, flavour), measure)
    if x.ndim != 5:
        raise ValueError("logits must be [N,H,W,C] or [T,N,H,W,C]")
    T = x.shape[0]
    if T == 1:
        return pixel_confidence(x[0], measure, flavour)
    probs = np.stack([softmax(x[t], flavour) for t in range(T)], axis=0)
    mu, m2 = welford_mean_m2(probs)
    if measure == "variance":
        v = (m2 / np.float32(T)).astype(np.float32)
        return (np.float32(1.0) - v).astype(np.float32)
    return confidence_from_prob(mu, measure, num_classes=x.shape[-1])


def image_scores(conf_map: np.ndarray) -> np.ndarray:
    """f64 mean over (H, W) of the f32 map -- active_learning.py:261-263."""
    return np.mean(np.asarray(conf_map, dtype=np.float32).astype(np.float64), axis=(1, 2))


def pseudo_label(logits: np.ndarray) -> np.ndarray:
    """u8(argmax_c logits), first maximum wins -- active_learning.py:234-236.
    For T > 1 the label is taken from sample 0 (repo spec)."""
    x = np.asarray(logits, dtype=np.float32)
    if x.ndim == 5:
        x = x[0]
    return np.argmax(x, axis=-1).astype(np.uint8)


def pseudo_mask(conf_map: np.ndarray, threshold: float) -> np.ndarray:
    """conf < threshold ? 0 : 1 -- active_learning.py:265-269 (threshold is cast
    to the map's dtype, float32, by TF's python-scalar conversion)."""
    return np.where(np.asarray(conf_map, np.float32) < np.float32(threshold), 0, 1).astype(np.uint8)


def score_pool(logits: np.ndarray, measure: str, flavour: str = "gpu") -> np.ndarray:
    """logits -> per-image f64 scores [N]."""
    return image_scores(pixel_confidence(logits, measure, flavour))


# --------------------------------------------------------------------------- #
# fp64 truth model (error budgets only; never the parity target)              #
# --------------------------------------------------------------------------- #
def pixel_confidence_f64(logits: np.ndarray, measure: str) -> np.ndarray:
    x = np.asarray(logits, dtype=np.float64)
    if x.ndim == 4:
        x = x[None]
    T, C = x.shape[0], x.shape[-1]
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    p = e / e.sum(axis=-1, keepdims=True)
    mu = p.mean(axis=0)
    if measure == "variance":
        return 1.0 - ((p - mu) ** 2).mean(axis=0).sum(axis=-1)
    if measure == "entropy":
        with np.errstate(divide="ignore", invalid="ignore"):
            q = np.where(mu > 0, -mu * np.log(mu), 0.0)
        return 1.0 - q.sum(axis=-1) / np.log(float(C))
    if measure == "margin":
        part = np.partition(mu, C - 2, axis=-1)
        return part[..., -1] - part[..., -2]
    if measure == "confidence":
        return mu.max(axis=-1)
    raise NotImplementedError("Uncertainty function not implemented.")


# --------------------------------------------------------------------------- #
# rank_confidence (active_learning.py:682-715)                                #
# --------------------------------------------------------------------------- #
def scatter_scores(num_examples: int, batches) -> np.ndarray:
    """:684-700 -- confidence = zeros(f32); confidence[idx] = f64 batch scores."""
    confidence = np.zeros(num_examples, dtype=np.float32)
    for batch_confidence, batch_indices in batches:
        confidence[np.asarray(batch_indices)] = batch_confidence
    return confidence


def select_lowest(confidence: np.ndarray, unlabelled: np.ndarray, selection_size: int):
    """:705-715 verbatim semantics, including np.argpartition's ValueError when
    selection_size >= len(unlabelled).  Returns (low_conf_examples, unlabelled_confidence)."""
    unlabelled = np.asarray(unlabelled)
    unlabelled_confidence = confidence[unlabelled]
    selection_size = np.minimum(len(unlabelled), selection_size)
    example_indices = np.argpartition(unlabelled_confidence, selection_size)
    example_indices = example_indices[:selection_size]
    low_conf_examples = unlabelled[example_indices]
    return low_conf_examples, unlabelled_confidence


def select_lowest_total_order(confidence: np.ndarray, unlabelled: np.ndarray, selection_size: int):
    """The deterministic completion of :705-715 that the product implements:
    k = min(len, selection_size) smallest under the total order
    (score with -0.0 == +0.0 and NaN last, then lower example id); k >= len returns all.
    Returned ids are ordered by that total order."""
    unlabelled = np.asarray(unlabelled, dtype=np.int64)
    u = np.asarray(confidence, dtype=np.float32)[unlabelled]
    k = int(min(len(unlabelled), max(int(selection_size), 0)))
    key = np.where(np.isnan(u), np.float32(np.inf), u + np.float32(0.0))
    nan_rank = np.isnan(u).astype(np.int8)
    order = np.lexsort((unlabelled, nan_rank, key))
    return unlabelled[order[:k]], u


def rank_confidence(logits: np.ndarray, unlabelled, selection_size: int, measure: str,
                    batch_size: int = 8, flavour: str = "gpu", num_examples: int | None = None,
                    example_index=None):
    """Whole closure: batches of ``batch_size`` images are scored, scattered by
    example index into an f32 vector, then the k lowest unlabelled are picked."""
    x = np.asarray(logits)
    n = x.shape[-4]
    if example_index is None:
        example_index = np.arange(n)
    example_index = np.asarray(example_index)
    if num_examples is None:
        num_examples = int(example_index.max()) + 1 if n else 0

    def _batches():
        for i in range(0, n, batch_size):
            sl = slice(i, i + batch_size)
            xb = x[sl] if x.ndim == 4 else x[:, sl]
            yield score_pool(xb, measure, flavour), example_index[sl]

    confidence = scatter_scores(num_examples, _batches())
    return select_lowest(confidence, np.asarray(unlabelled), selection_size)


# --------------------------------------------------------------------------- #
# Logits producer: ENet `Final` (models/enet/enet_modules.py:1294-1381)       #
# --------------------------------------------------------------------------- #
def conv2d_transpose_same(value: np.ndarray, kernel: np.ndarray, strides=(2, 2), dtype=np.float32) -> np.ndarray:
    """tf.nn.conv2d_transpose(value, kernel, output_shape=[B, s*h, s*w, C], strides=[1,s,s,1], padding="SAME")
    -- the only op of ``Final.call`` (models/enet/enet_modules.py:1376-1380).

    value  [B, h, w, Cin]; kernel [kh, kw, Cout, Cin] (TF's conv2d_transpose filter layout, :1341).
    Defined as the gradient of conv2d w.r.t. its input: the forward conv with SAME padding reads
    X[s*i + ky - pad_top, s*j + kx - pad_left]; for an even output size 2h, stride 2 and a 3x3 kernel
    the total padding is 1 and TF puts it at the bottom/right (pad_top = pad_left = 0)
    [TF-upstream: GetWindowedOutputSize, pad_before = pad_total // 2].  Hence
        out[s*i + ky - pad_top, s*j + kx - pad_left, c] += sum_o value[i, j, o] * kernel[ky, kx, c, o].
    One fp32 matmul per tap, accumulated in tap order (TF's own summation order is unspecified)."""
    v = np.asarray(value, dtype=dtype)
    k = np.asarray(kernel, dtype=dtype)
    B, h, w, cin = v.shape
    kh, kw, cout, cin2 = k.shape
    assert cin == cin2, "kernel must be [kh, kw, out_channels, in_channels]"
    sy, sx = strides
    H, W = sy * h, sx * w
    pad_y = max((h - 1) * sy + kh - H, 0)
    pad_x = max((w - 1) * sx + kw - W, 0)
    top, left = pad_y // 2, pad_x // 2
    full = np.zeros((B, (h - 1) * sy + kh, (w - 1) * sx + kw, cout), dtype=dtype)
    for ky in range(kh):
        for kx in range(kw):
            contrib = (v.reshape(-1, cin) @ k[ky, kx].T).reshape(B, h, w, cout).astype(dtype)
            full[:, ky:ky + (h - 1) * sy + 1:sy, kx:kx + (w - 1) * sx + 1:sx, :] += contrib
    return np.ascontiguousarray(full[:, top:top + H, left:left + W, :])


def final_head(features: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """``Final.call`` (models/enet/enet_modules.py:1359-1381): [B,h,w,16] -> logits [B,2h,2w,C], fp32."""
    return conv2d_transpose_same(features, kernel, (2, 2), np.float32)


def score_pool_from_features(features: np.ndarray, kernel: np.ndarray, measure: str, flavour: str = "gpu") -> np.ndarray:
    """Final head + scoring: features [N,h,w,16] or [T,N,h,w,16] -> per-image f64 scores [N]."""
    f = np.asarray(features, np.float32)
    if f.ndim == 5:
        logits = np.stack([final_head(f[t], kernel) for t in range(f.shape[0])], axis=0)
    else:
        logits = final_head(f, kernel)
    return score_pool(logits, measure, flavour)
