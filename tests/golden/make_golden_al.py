#!/usr/bin/env python
"""Generate tests/golden/al_loop.npz by EXECUTING THE REFERENCE'S OWN STATEMENTS for steps 3/4 of
the active-learning iteration (build container only; needs /root/reference).

Executed verbatim (located with ``ast``, compiled in memory, nothing copied):
  * the ``if alparams["selection_size"] > 0: ... else: ...`` statement   active_learning.py:779-793
    (rank branch with a stub ``rank_confidence``/``sess``; random-baseline branch with seeded np.random)
  * the state update from ``labelled = np.append(...)`` to ``state["checkpoint"] = ...``  :839-846
"""
from __future__ import annotations

import ast
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ALS_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "active_learning.py")


def _is_selection_if(node):
    if not isinstance(node, ast.If) or not isinstance(node.test, ast.Compare):
        return False
    left = node.test.left
    return (isinstance(left, ast.Subscript) and getattr(left.value, "id", "") == "alparams"
            and isinstance(left.slice, ast.Constant) and left.slice.value == "selection_size"
            and isinstance(node.test.ops[0], ast.Gt) and node.orelse)


def locate(tree):
    sel = [n for n in ast.walk(tree) if _is_selection_if(n)]
    assert len(sel) == 1, "selection branch not found"
    upd = None
    for node in ast.walk(tree):
        body = getattr(node, "body", None)
        if not isinstance(body, list):
            continue
        for i, st in enumerate(body):
            if (isinstance(st, ast.Assign) and getattr(st.targets[0], "id", "") == "labelled"
                    and isinstance(st.value, ast.Call) and getattr(st.value.func, "attr", "") == "append"):
                j = i
                while not (isinstance(body[j], ast.Assign) and isinstance(body[j].targets[0], ast.Subscript)
                           and getattr(body[j].targets[0].slice, "value", "") == "checkpoint"):
                    j += 1
                upd = body[i:j + 1]
    assert upd, "state update not found"
    return sel[0], upd


def run_selection(code, selection_size, unlabelled, seed, rank_result=None):
    calls = {}

    class Sess:
        def run(self, fetch, feed):
            calls["hist_input"] = np.asarray(list(feed.values())[0])
            return "summary"

    g = {"np": np, "alparams": {"selection_size": selection_size}, "unlabelled": np.asarray(unlabelled),
         "rank_confidence": lambda: rank_result, "sess": Sess(), "conf_summary": "cs", "conf_summary_ph": "ph",
         "test_writer": types.SimpleNamespace(add_summary=lambda s, it: calls.setdefault("written", (s, it))),
         "state": {"iteration": 3}}
    np.random.seed(seed)
    exec(code, g)
    return g["low_conf_examples"], calls


def run_update(code, labelled, unlabelled, low, filenames, iteration, ckpt):
    state = {"checkpoint": None, "iteration": iteration,
             "dataset": {"train": {"filenames": list(filenames), "labelled": [], "unlabelled": [], "no_label": []}}}
    g = {"np": np, "labelled": np.asarray(labelled), "unlabelled": np.asarray(unlabelled),
         "low_conf_examples": low, "train_examples": np.asarray(filenames), "state": state, "checkpoint_path": ckpt}
    exec(code, g)
    return g["labelled"], g["unlabelled"], state


def main():
    with open(SRC) as f:
        tree = ast.parse(f.read())
    sel, upd = locate(tree)
    sel_code = compile(ast.Module(body=[sel], type_ignores=[]), SRC, "exec")
    upd_code = compile(ast.Module(body=upd, type_ignores=[]), SRC, "exec")
    out = {"meta_reference_lines": np.asarray([sel.lineno, sel.end_lineno, upd[0].lineno, upd[-1].end_lineno], np.int64)}

    # random baseline: selection_size <= 0
    unl = np.arange(100, 160)
    names = []
    for name, size, u, seed in [("rand_50", -50, unl, 11), ("rand_more_than_left", -80, unl[:30], 12),
                                ("rand_zero", 0, unl, 13), ("rand_one", -1, unl[:1], 14)]:
        low, _ = run_selection(sel_code, size, u, seed)
        out[name + ".selection_size"] = np.int64(size)
        out[name + ".unlabelled"] = np.asarray(u, np.int64)
        out[name + ".seed"] = np.int64(seed)
        out[name + ".low"] = np.asarray(low, np.int64)
        names.append(name)
    out["random_cases"] = np.asarray(names)

    # rank branch: passes rank_confidence()'s tuple through and feeds the histogram placeholder
    conf = np.random.default_rng(5).random(60).astype(np.float32)
    low, calls = run_selection(sel_code, 7, unl, 0, rank_result=(unl[:7], conf))
    assert np.array_equal(low, unl[:7]) and np.array_equal(calls["hist_input"], conf) and calls["written"] == ("summary", 3)
    out["rank_branch.hist_input_is_unlabelled_conf"] = np.asarray(True)

    # state update
    rng = np.random.default_rng(9)
    files = ["ex_%03d.tfrecord" % i for i in range(40)]
    lab = np.arange(0, 10); unl2 = np.arange(10, 40)
    names = []
    for name, low in [("upd_basic", rng.choice(unl2, 6, replace=False)), ("upd_dupes", np.asarray([12, 12, 30])),
                      ("upd_empty_list", [])]:
        l2, u2, st = run_update(upd_code, lab, unl2, low, files, 4, "ckpt/model-5")
        out[name + ".low"] = np.asarray(low, np.int64)
        out[name + ".labelled_in"] = lab.astype(np.int64)
        out[name + ".unlabelled_in"] = unl2.astype(np.int64)
        out[name + ".labelled"] = np.asarray(l2, np.float64)     # np.append of [] promotes to float64: keep the dtype evidence
        out[name + ".labelled_dtype"] = np.asarray(str(np.asarray(l2).dtype))
        out[name + ".unlabelled"] = np.asarray(u2, np.int64)
        out[name + ".state_json"] = np.asarray(json.dumps(st, sort_keys=True))
        names.append(name)
    out["update_cases"] = np.asarray(names)

    path = os.path.join(HERE, "al_loop.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "lines", out["meta_reference_lines"])


if __name__ == "__main__":
    main()
