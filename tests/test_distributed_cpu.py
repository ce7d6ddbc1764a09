"""world_size-2 gloo test of the sharded selection's host logic (candidate all-gather + merge).

The per-rank scoring kernel is CUDA-only, so the test feeds per-rank score vectors computed by the
oracle and injects the oracle's total-order selection as the `select` callable; what is under
test is the sharding, the fixed-size candidate exchange, padding, and the merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_select(keys, ids, k):
    keys = np.asarray(keys, np.float32)
    ids = np.asarray(ids, np.int64)
    kk = np.where(np.isnan(keys), np.float32(np.inf), keys + np.float32(0.0))
    order = np.lexsort((ids, np.isnan(keys).astype(np.int8), kk))[:max(0, min(k, len(ids)))]
    return keys[order], ids[order]


def _worker(rank, world, port, cases, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from semanticsegmentationactivelearning_b200 import rank_confidence_sharded, shard_bounds
        for name, (scores, unlabelled, k) in cases.items():
            lo, hi = shard_bounds(len(scores), rank, world)
            ids, uconf = rank_confidence_sharded(scores[lo:hi], np.arange(lo, hi), unlabelled, k,
                                                 select=_oracle_select)
            # the single-collective form (upper bound on a rank's unlabelled count known up front) must agree
            ids1, uconf1 = rank_confidence_sharded(scores[lo:hi], np.arange(lo, hi), unlabelled, k, select=_oracle_select,
                                                   max_unlabelled_per_rank=-(-len(scores) // world) + 3)
            assert np.array_equal(ids, ids1) and np.array_equal(uconf, uconf1, equal_nan=True), name
            np.savez(os.path.join(out_dir, "%s_r%d.npz" % (name, rank)), ids=ids, uconf=uconf)
    finally:
        dist.destroy_process_group()


def _cases():
    rng = np.random.default_rng(11)
    n = 101
    cases = {}
    for name in ("spread", "ties", "one_sided", "k_ge_m", "nan", "empty_unlabelled"):
        scores = rng.random(n).astype(np.float32)
        unlabelled = np.sort(rng.choice(n, 70, replace=False))
        k = 12
        if name == "ties":
            scores = np.round(scores, 1)
        elif name == "one_sided":                     # every winner lives on rank 1; rank 0 pads
            scores[: n // 2] += 10
        elif name == "k_ge_m":
            k = 200
        elif name == "nan":
            scores[unlabelled[:5]] = np.nan
        elif name == "empty_unlabelled":
            unlabelled = np.zeros(0, np.int64)
        cases[name] = (scores, unlabelled, k)
    return cases


def test_sharded_selection_matches_single_process(tmp_path):
    from oracle import reference_np as R
    world = 2
    cases = _cases()
    mp.spawn(_worker, args=(world, _free_port(), cases, str(tmp_path)), nprocs=world, join=True)
    for name, (scores, unlabelled, k) in cases.items():
        want_ids, want_u = R.select_lowest_total_order(scores, unlabelled, k)
        for r in range(world):
            got = np.load(os.path.join(tmp_path, "%s_r%d.npz" % (name, r)))
            assert np.array_equal(got["ids"], want_ids), name
            assert np.array_equal(got["uconf"], want_u, equal_nan=True), name
