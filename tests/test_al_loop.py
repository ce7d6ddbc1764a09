"""Steps 3/4 of the active-learning iteration (al_loop.py) against the golden fixture produced by
executing /root/reference/active_learning.py:779-793 and :839-846 verbatim (tests/golden/make_golden_al.py)."""
import json
import os

import numpy as np
import pytest

from semanticsegmentationactivelearning_b200 import al_loop

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(ROOT, "tests", "golden", "al_loop.npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def test_random_baseline_matches_reference_draws(gold):
    for name in gold["random_cases"]:
        size, unl, seed = int(gold[name + ".selection_size"]), gold[name + ".unlabelled"], int(gold[name + ".seed"])
        np.random.seed(seed)                                   # the reference uses the global NumPy RNG (:790)
        low = al_loop.draw_random(unl, size)
        assert np.array_equal(np.asarray(low, np.int64), gold[name + ".low"]), name
        low2 = al_loop.draw_random(unl, size, rng=np.random.RandomState(seed))
        assert np.array_equal(np.asarray(low2, np.int64), gold[name + ".low"]), name


def test_random_baseline_without_replacement_is_a_subset():
    unl = np.arange(50, 80)
    low = al_loop.draw_random(unl, -100, rng=np.random.RandomState(0), replace=False)
    assert len(low) == 30 and len(set(low.tolist())) == 30 and set(low.tolist()) <= set(unl.tolist())


def test_state_update_matches_reference(gold):
    files = ["ex_%03d.tfrecord" % i for i in range(40)]
    for name in gold["update_cases"]:
        state = {"checkpoint": None, "iteration": 4,
                 "dataset": {"train": {"filenames": list(files), "labelled": [], "unlabelled": [], "no_label": []}}}
        lab, unl = al_loop.update_state(state, gold[name + ".labelled_in"], gold[name + ".unlabelled_in"],
                                        gold[name + ".low"], "ckpt/model-5", train_examples=files)
        assert np.array_equal(unl, gold[name + ".unlabelled"]), name
        assert np.array_equal(np.asarray(lab, np.float64), gold[name + ".labelled"]), name
        want = json.loads(str(gold[name + ".state_json"]))
        got = json.loads(json.dumps(state, sort_keys=True))
        # np.append(int array, []) yields float64 in the reference (ids print as 3.0); the ids themselves agree
        assert [int(v) for v in got["dataset"]["train"]["labelled"]] == [int(v) for v in want["dataset"]["train"]["labelled"]]
        got["dataset"]["train"]["labelled"] = want["dataset"]["train"]["labelled"] = None
        assert got == want, name


def test_select_examples_branches(gold):
    unl = np.arange(100, 160)
    conf = np.linspace(0.1, 0.9, 60).astype(np.float32)
    low, c, hist = al_loop.select_examples({"selection_size": 7}, unl, lambda: (unl[:7], conf))
    assert np.array_equal(low, unl[:7]) and c is conf
    assert hist["num"] == 60 and abs(hist["sum"] - float(conf.astype(np.float64).sum())) < 1e-12
    assert bool(gold["rank_branch.hist_input_is_unlabelled_conf"])
    called = []
    low, c, hist = al_loop.select_examples({"selection_size": -5}, unl, lambda: called.append(1),
                                           rng=np.random.RandomState(1))
    assert len(low) == 5 and c is None and hist is None and not called


def test_confidence_distribution_buckets():
    lim = al_loop.histogram_bucket_limits()
    assert lim[0] == -np.finfo(np.float64).max and lim[-1] == np.finfo(np.float64).max
    assert np.all(np.diff(lim) > 0) and 0.0 in lim
    pos = lim[lim > 0]
    assert pos[0] == 1e-12 and np.allclose(pos[1:-1] / pos[:-2], 1.1)
    v = np.asarray([0.5, 0.5, 0.25, 0.0, -0.1, 1.0], np.float32)
    h = al_loop.confidence_distribution(v)
    assert h["num"] == 6 and h["min"] == pytest.approx(-0.1) and h["max"] == 1.0
    assert sum(h["bucket"]) == 6 and len(h["bucket"]) == len(h["bucket_limit"])
    # every value sits in the first bucket whose limit is strictly greater
    for x in v.astype(np.float64):
        i = int(np.searchsorted(np.asarray(h["bucket_limit"]), x, side="right"))
        assert h["bucket"][i] >= 1
    # runs of empty buckets are collapsed: no two consecutive empty entries
    b = h["bucket"]
    assert all(not (b[i] == 0 and b[i + 1] == 0) for i in range(len(b) - 1))
    e = al_loop.confidence_distribution([])
    assert e["num"] == 0 and sum(e["bucket"]) == 0


def test_acquisition_step_roundtrip(tmp_path):
    state = {"checkpoint": None, "iteration": 0,
             "dataset": {"train": {"filenames": ["f%d" % i for i in range(20)], "labelled": [0, 1, 2],
                                   "unlabelled": list(range(3, 20)), "no_label": []},
                         "val": {"filenames": []}, "test": {"filenames": []}}}
    conf = np.linspace(0, 1, 17).astype(np.float32)
    path = str(tmp_path / "state.json")
    r = al_loop.acquisition_step(state, {"selection_size": 4}, lambda: (np.asarray([3, 4, 5, 6]), conf),
                                 checkpoint_path="ck-1", state_filename=path)
    assert r["labelled"].tolist() == [0, 1, 2, 3, 4, 5, 6] and r["unlabelled"].tolist() == list(range(7, 20))
    with open(path) as f:
        on_disk = json.load(f)
    assert on_disk == state and on_disk["iteration"] == 1 and on_disk["checkpoint"] == "ck-1"
    assert set(on_disk["dataset"]["train"]) == {"filenames", "labelled", "unlabelled", "no_label"}
