"""The host-side remainder of one active-learning iteration around the scorer.

Mirrors, with the reference's own names, what /root/reference/active_learning.py does between the
call of ``rank_confidence()`` and the ``state.json`` dump (steps 3 and 4 of the major loop):

* :779-784  ``selection_size > 0``: rank the pool, feed ``unlabelled_conf`` to the
            ``ConfidenceDistribution`` histogram summary (:425-427)
* :785-793  ``selection_size <= 0``: the random-acquisition baseline
            (conf/enet_cityscapes_active_learning_baseline.json:58)
* :838-851  labelled += low_conf_examples, unlabelled -= low_conf_examples, iteration += 1,
            checkpoint path; the dict layout of ``state.json`` (:111-130) is kept so an existing
            log directory resumes unchanged.

Nothing here is arithmetic on the pool: ranking runs on the GPU (``acquisition.rank_confidence``);
this module is bookkeeping on index lists of at most a few thousand entries.
"""
from __future__ import annotations

import json
import os
from typing import Callable, Optional, Sequence, Tuple

import numpy as np

__all__ = ["histogram_bucket_limits", "confidence_distribution", "draw_random", "select_examples",
           "update_state", "dump_state", "acquisition_step"]


# ---- ConfidenceDistribution (:425-427, :781-784) -------------------------------------------------
def histogram_bucket_limits() -> np.ndarray:
    """Bucket upper limits of a TF-1.x histogram summary [TF-upstream: core/lib/histogram
    InitDefaultBucketsInner]: +-1e-12 * 1.1^i up to 1e20, mirrored around 0, DBL_MAX last."""
    pos = []
    v = 1e-12
    while v < 1e20:
        pos.append(v)
        v *= 1.1
    pos.append(float(np.finfo(np.float64).max))
    return np.asarray([-x for x in reversed(pos)] + [0.0] + pos, dtype=np.float64)


_LIMITS = None


def confidence_distribution(unlabelled_conf) -> dict:
    """What ``sess.run(conf_summary, {conf_summary_ph: unlabelled_conf})`` (:781-783) puts into the
    event file: the HistogramProto fields (min, max, num, sum, sum_squares, bucket_limit, bucket)
    of the values cast to float64 (placeholder dtype, :425).  Runs of empty buckets are collapsed
    into one entry like Histogram::EncodeToProto does [TF-upstream]."""
    global _LIMITS
    if _LIMITS is None:
        _LIMITS = histogram_bucket_limits()
    v = np.asarray(unlabelled_conf, dtype=np.float64).ravel()
    counts = np.zeros(_LIMITS.size, dtype=np.float64)
    if v.size:
        # bucket = first limit strictly greater than the value (std::upper_bound); NaN goes last
        b = np.searchsorted(_LIMITS, v, side="right")
        b = np.minimum(b, _LIMITS.size - 1)
        np.add.at(counts, b, 1.0)
    limit, bucket = [], []
    i = 0
    n = counts.size
    while i < n:
        end, count = _LIMITS[i], counts[i]
        i += 1
        if count <= 0.0:
            while i < n and counts[i] <= 0.0:
                end, count = _LIMITS[i], counts[i]
                i += 1
        limit.append(float(end))
        bucket.append(float(count))
    return {
        "min": float(v.min()) if v.size else float(np.finfo(np.float64).max),
        "max": float(v.max()) if v.size else -float(np.finfo(np.float64).max),
        "num": float(v.size),
        "sum": float(v.sum()),
        "sum_squares": float((v * v).sum()),
        "bucket_limit": limit,
        "bucket": bucket,
    }


# ---- example selection (:779-793) ------------------------------------------------------------------
def draw_random(unlabelled, selection_size: int, rng=None, replace: bool = True):
    """The random baseline (:785-793): ``np.random.choice(unlabelled, abs(selection_size))``.

    The reference draws WITH replacement (NumPy's default) and more examples than are left if
    ``abs(selection_size) > len(unlabelled)``; that behaviour is kept by default so a seeded run
    reproduces it.  ``replace=False`` draws min(|selection_size|, len(unlabelled)) distinct examples.
    ``rng`` is anything with ``.choice`` (``np.random`` itself by default, like the reference)."""
    rng = np.random if rng is None else rng
    unlabelled = np.asarray(unlabelled)
    # :787-788  np.minimum(selection_size (<= 0), len(unlabelled)) is 0 only for selection_size == 0
    if int(np.minimum(selection_size, len(unlabelled))) == 0:
        return []
    k = int(np.abs(selection_size))
    if not replace:
        k = min(k, len(unlabelled))
        return rng.choice(unlabelled, k, replace=False)
    return rng.choice(unlabelled, k)


def select_examples(alparams: dict, unlabelled, rank_fn: Callable[[], Tuple[np.ndarray, np.ndarray]],
                    rng=None, replace: bool = True):
    """Step 3 (:776-793).  ``rank_fn`` is the ``rank_confidence`` closure.  Returns
    ``(low_conf_examples, unlabelled_conf or None, histogram or None)``."""
    if alparams["selection_size"] > 0:
        low_conf_examples, unlabelled_conf = rank_fn()
        return low_conf_examples, unlabelled_conf, confidence_distribution(unlabelled_conf)
    return draw_random(unlabelled, alparams["selection_size"], rng, replace), None, None


# ---- state update (:838-851) -----------------------------------------------------------------------
def update_state(state: dict, labelled, unlabelled, low_conf_examples, checkpoint_path=None,
                 train_examples: Optional[Sequence[str]] = None):
    """Step 4: returns the new ``(labelled, unlabelled)`` arrays and updates ``state`` in place."""
    labelled = np.asarray(labelled)
    unlabelled = np.asarray(unlabelled)
    low = np.asarray(low_conf_examples, dtype=unlabelled.dtype if unlabelled.size else np.int64)
    labelled = np.append(labelled, low)                                           # :839
    unlabelled = unlabelled[np.isin(unlabelled, low, assume_unique=True, invert=True)]  # :840-841
    train = state["dataset"]["train"]
    if train_examples is not None:
        train["filenames"] = np.asarray(train_examples).tolist()                  # :842
    train["labelled"] = labelled.tolist()                                         # :843
    train["unlabelled"] = unlabelled.tolist()                                     # :844
    state["iteration"] += 1                                                       # :845
    state["checkpoint"] = checkpoint_path                                         # :846
    return labelled, unlabelled


def dump_state(state: dict, state_filename: str) -> None:
    """:848-851 -- ``json.dump(state, f, indent=2)`` (written to a temporary file first, then renamed,
    so an interrupted dump never leaves a truncated state.json)."""
    tmp = state_filename + ".tmp"
    with open(tmp, "w") as f:
        json.dump(state, f, indent=2)
    os.replace(tmp, state_filename)


def acquisition_step(state: dict, alparams: dict, rank_fn, *, checkpoint_path=None, rng=None,
                     replace: bool = True, state_filename: Optional[str] = None) -> dict:
    """Steps 3 + 4 of one major iteration on a ``state`` dict laid out like state.json (:111-130)."""
    train = state["dataset"]["train"]
    labelled = np.asarray(train["labelled"], dtype=np.int64)
    unlabelled = np.asarray(train["unlabelled"], dtype=np.int64)
    low, conf, hist = select_examples(alparams, unlabelled, rank_fn, rng, replace)
    labelled, unlabelled = update_state(state, labelled, unlabelled, low, checkpoint_path)
    if state_filename:
        dump_state(state, state_filename)
    return {"low_conf_examples": np.asarray(low, dtype=np.int64), "unlabelled_conf": conf,
            "ConfidenceDistribution": hist, "labelled": labelled, "unlabelled": unlabelled}
