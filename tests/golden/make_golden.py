#!/usr/bin/env python
"""Generate tests/golden/reference_path.npz by EXECUTING THE REFERENCE'S OWN SOURCE.

Runs only in the build container (needs /root/reference, read-only); the GPU box
and the test-suite use the committed .npz.  Nothing is copied from the reference
into this repo: the statements are located with ``ast`` and compiled in memory.

What is executed verbatim:
  * ``EPSILON``                         active_learning.py:40
  * the PseudoAnnotation graph          active_learning.py:234-269
    (pseudo_label, pseudo_prob, the measure branch, pseudo_mean_confidence,
    pseudo_mask), once per ``alparams["measure"]``
  * the ``rank_confidence`` closure     active_learning.py:682-715

TensorFlow 1.13.2 (requirements.txt:1) is not installable here, so ``tf`` is a
NumPy stand-in that implements each op the graph calls with TF's documented
semantics in float32 (softmax = exp(x-max)/sum, top_k sorted descending,
reduce_* over the given axes, cast, where, less).  This pins the oracle's
*composition* of ops to the reference source; the primitive kernels stay
unpinned (see oracle/reference_np.py header).
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("ALS_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "active_learning.py")

from oracle import synth  # noqa: E402


# ----------------------------- NumPy stand-in for tf ------------------------ #
def _softmax(logits, axis=-1, name=None):
    x = np.asarray(logits, np.float32)
    m = np.max(x, axis=axis, keepdims=True)
    with np.errstate(invalid="ignore"):
        e = np.exp(x - m, dtype=np.float32)
    return (e / np.sum(e, axis=axis, keepdims=True, dtype=np.float32)).astype(np.float32)


def _top_k(x, k=1, sorted=True, name=None):
    x = np.asarray(x)
    idx = np.argsort(-x, axis=-1, kind="stable")[..., :k]
    return np.take_along_axis(x, idx, axis=-1), idx.astype(np.int32)


def _log(x, name=None):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(x, dtype=np.asarray(x).dtype)


class _OutOfRangeError(Exception):
    pass


class _NameScope:
    def __init__(self, *_a, **_k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def make_tf():
    tf = types.SimpleNamespace()
    tf.float64, tf.float32, tf.uint8 = np.float64, np.float32, np.uint8
    tf.name_scope = _NameScope
    tf.stop_gradient = lambda x: x
    tf.cast = lambda x, dtype, name=None: np.asarray(x).astype(dtype)
    tf.reduce_mean = lambda x, axis=None, name=None: np.mean(x, axis=axis, dtype=np.asarray(x).dtype)
    tf.where = lambda c, a, b, name=None: np.where(c, a, b)
    tf.zeros_like = lambda x, dtype=None: np.zeros_like(x, dtype=dtype)
    tf.ones_like = lambda x, dtype=None: np.ones_like(x, dtype=dtype)
    tf.nn = types.SimpleNamespace(softmax=_softmax)
    tf.math = types.SimpleNamespace(
        argmax=lambda x, axis=None, name=None: np.argmax(x, axis=axis).astype(np.int64),
        log=_log,
        reduce_sum=lambda x, axis=None, name=None: np.sum(x, axis=axis, dtype=np.asarray(x).dtype),
        reduce_max=lambda x, axis=None, name=None: np.max(x, axis=axis),
        top_k=_top_k,
        less=lambda a, b, name=None: np.less(a, b),
    )
    tf.errors = types.SimpleNamespace(OutOfRangeError=_OutOfRangeError)
    return tf


# ----------------------------- locate reference code ------------------------ #
def _load():
    with open(SRC) as f:
        src = f.read()
    return src, ast.parse(src)


def _find_pseudo_annotation(tree):
    for node in ast.walk(tree):
        if isinstance(node, ast.With):
            ce = node.items[0].context_expr
            if (isinstance(ce, ast.Call) and getattr(ce.func, "attr", "") == "name_scope"
                    and ce.args and isinstance(ce.args[0], ast.Constant)
                    and ce.args[0].value == "PseudoAnnotation"):
                return node
    raise RuntimeError("PseudoAnnotation scope not found")


def _targets(stmt):
    if isinstance(stmt, ast.Assign):
        out = []
        for t in stmt.targets:
            out += [e.id for e in ast.walk(t) if isinstance(e, ast.Name)]
        return out
    return []


def graph_code(tree):
    """Statements of the scope from `pseudo_label = argmax` through `pseudo_mask = ...`
    (skips the network call :231 and the tf.where merges with the ground truth :272-275)."""
    scope = _find_pseudo_annotation(tree)
    keep = []
    for st in scope.body:
        tg = _targets(st)
        if isinstance(st, ast.If):       # the measure branch :240-260
            keep.append(st)
        elif tg and tg[0] in ("pseudo_label", "pseudo_prob", "pseudo_mean_confidence", "pseudo_mask"):
            keep.append(st)
    lines = (min(s.lineno for s in keep), max(s.end_lineno for s in keep))
    mod = ast.Module(body=keep, type_ignores=[])
    return compile(mod, SRC, "exec"), lines


def epsilon_code(tree):
    for st in tree.body:
        if "EPSILON" in _targets(st):
            return compile(ast.Module(body=[st], type_ignores=[]), SRC, "exec"), st.lineno
    raise RuntimeError("EPSILON not found")


def rank_confidence_code(tree):
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "rank_confidence":
            return (compile(ast.Module(body=[node], type_ignores=[]), SRC, "exec"),
                    (node.lineno, node.end_lineno))
    raise RuntimeError("rank_confidence not found")


# ----------------------------- run the graph -------------------------------- #
def run_graph(code, eps_code, logits, measure, threshold, num_classes):
    tf = make_tf()
    g = {"tf": tf, "np": np}
    exec(eps_code, g)
    g.update(
        pseudo_logits=np.asarray(logits, np.float32),
        alparams={"measure": measure, "threshold": threshold},
        dataset=types.SimpleNamespace(num_classes=num_classes),
        train_label=np.zeros(logits.shape[:-1], np.uint8),
    )
    exec(code, g)
    return (g["pseudo_confidence"], g["pseudo_mean_confidence"], g["pseudo_label"], g["pseudo_mask"])


def run_rank_confidence(code, scores64, example_index, num_examples, unlabelled, selection_size,
                        batch_size, truncate_after=None):
    """Drive the reference closure with a fake session that hands out the given
    per-image f64 scores batch by batch (as sess.run([pseudo_mean_confidence, train_index]))."""
    tf = make_tf()
    batches = [(scores64[i:i + batch_size], example_index[i:i + batch_size])
               for i in range(0, len(scores64), batch_size)]
    if truncate_after is not None:
        batches = batches[:truncate_after]
    it = iter(batches)

    class Sess:
        def run(self, fetches):
            try:
                return next(it)
            except StopIteration:
                raise tf.errors.OutOfRangeError()

    g = {
        "tf": tf, "np": np,
        "state": {"dataset": {"train": {"filenames": [None] * num_examples}}},
        "train_input_stage": types.SimpleNamespace(init_iterator=lambda *a, **k: None),
        "train_input": types.SimpleNamespace(size=len(scores64) + (batch_size if truncate_after else 0),
                                             feed_dict={}),
        "sess": Sess(), "params": {"batch_size": batch_size}, "show_progress": False,
        "labelled": [], "pseudo_mean_confidence": "pmc", "train_index": "ti",
        "unlabelled": np.asarray(unlabelled), "alparams": {"selection_size": selection_size},
    }
    exec(code, g)
    return g["rank_confidence"]()


def special_logits(C):
    """Analytic known-answer pixels as one [1, 1, K, C] image."""
    rows = []
    rows.append(np.zeros(C))                                  # uniform
    rows.append(np.full(C, 3.25))                             # uniform, shifted
    d = np.zeros(C); d[0] = 100.0; rows.append(d)             # one dominant logit (p_rest underflows)
    t = np.full(C, -80.0); t[1] = t[C - 1] = 2.0; rows.append(t)   # exact two-way tie
    r = np.linspace(-4, 4, C); rows.append(r)                 # ramp
    rows.append(r[::-1].copy())                               # permuted ramp
    rows.append(r + 17.0)                                     # shifted ramp
    m = np.zeros(C); m[2] = -np.inf; rows.append(m)           # a masked (-inf) class
    return np.asarray(rows, np.float32)[None, None]


def main():
    src, tree = _load()
    gcode, glines = graph_code(tree)
    ecode, eline = epsilon_code(tree)
    rcode, rlines = rank_confidence_code(tree)
    out = {"meta_reference_lines": np.asarray([eline, *glines, *rlines], np.int64)}

    cases = []
    for ci, (N, H, W, C) in enumerate([(3, 6, 10, 19), (2, 5, 7, 6), (2, 4, 4, 66), (2, 3, 5, 2)]):
        cases.append((f"rand{ci}", synth.synth_logits(1, 5 * ci, N, H, W, C, seed=1234 + ci)))
    for C in (19, 6, 3):
        cases.append((f"special_c{C}", special_logits(C)))
    names = []
    for name, logits in cases:
        out[f"{name}.logits"] = logits
        names.append(name)
        for measure in ("entropy", "margin", "confidence"):
            conf, mean, label, mask = run_graph(gcode, ecode, logits, measure, 0.9, logits.shape[-1])
            assert conf.dtype == np.float32 and mean.dtype == np.float64
            out[f"{name}.{measure}.conf"] = conf
            out[f"{name}.{measure}.mean"] = mean
            out[f"{name}.{measure}.mask"] = mask.astype(np.uint8)
        out[f"{name}.label"] = label.astype(np.uint8)
    out["graph_cases"] = np.asarray(names)

    # unknown measure -> NotImplementedError (:259-260)
    try:
        run_graph(gcode, ecode, cases[0][1], "bald", 0.9, 19)
        raise AssertionError("expected NotImplementedError")
    except NotImplementedError as e:
        out["unknown_measure_message"] = np.asarray(str(e))

    # rank_confidence cases
    rng = np.random.default_rng(7)
    rnames = []

    def add_rank(name, scores64, num_examples, unlabelled, k, batch_size=8, shuffle=True, truncate_after=None):
        idx = np.arange(len(scores64))
        if shuffle:
            idx = rng.permutation(len(scores64))       # NumpyCapsule shuffles (tensortools/input.py:350-359)
        ids, uconf = run_rank_confidence(rcode, scores64[idx], idx, num_examples, unlabelled, k, batch_size,
                                         truncate_after)
        out[f"{name}.scores64"] = scores64
        out[f"{name}.order"] = idx
        out[f"{name}.num_examples"] = np.int64(num_examples)
        out[f"{name}.unlabelled"] = np.asarray(unlabelled, np.int64)
        out[f"{name}.k"] = np.int64(k)
        out[f"{name}.batch_size"] = np.int64(batch_size)
        out[f"{name}.truncate_after"] = np.int64(-1 if truncate_after is None else truncate_after)
        out[f"{name}.ids"] = np.asarray(ids, np.int64)
        out[f"{name}.uconf"] = np.asarray(uconf)
        assert np.asarray(uconf).dtype == np.float32
        rnames.append(name)

    s = rng.random(200)
    add_rank("rank_basic", s, 200, np.sort(rng.choice(200, 150, replace=False)), 50)
    add_rank("rank_small_k", s, 200, np.arange(0, 200, 3), 1)
    add_rank("rank_k_m_minus_1", s[:40], 40, np.arange(40), 39, batch_size=7)
    s2 = np.round(rng.random(120), 1)                       # many exact duplicates
    add_rank("rank_ties", s2, 120, np.arange(120), 20, batch_size=5)
    s3 = rng.random(64); s3[[3, 17, 40]] = np.nan            # NaN sorts last
    add_rank("rank_nan", s3, 64, np.arange(64), 10)
    # f64 -> f32 rounding on scatter (:700): scores closer than an f32 ulp collapse
    s4 = 0.5 + np.arange(32) * 1e-9
    add_rank("rank_f32_round", s4, 32, np.arange(32), 4, shuffle=False)
    # short pass (:701-702): unvisited examples keep 0.0 and are selected first
    add_rank("rank_short_pass", 0.25 + 0.5 * rng.random(48), 48, np.arange(48), 6, batch_size=8,
             shuffle=False, truncate_after=4)
    out["rank_cases"] = np.asarray(rnames)

    # k == len(unlabelled): the reference raises ValueError (kth out of bounds)
    try:
        run_rank_confidence(rcode, s[:10], np.arange(10), 10, np.arange(10), 10, 8)
        out["rank_k_eq_m_raises"] = np.asarray("no")
    except ValueError:
        out["rank_k_eq_m_raises"] = np.asarray("ValueError")

    path = os.path.join(HERE, "reference_path.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "(%d arrays, %d bytes)" % (len(out), os.path.getsize(path)))
    print("reference lines used: EPSILON :%d, graph :%d-%d, rank_confidence :%d-%d" % (eline, *glines, *rlines))


if __name__ == "__main__":
    main()
