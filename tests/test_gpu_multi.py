"""The sharded pool pass behind the C ABI (csrc/comm.cu): per-rank candidates + score slices, ONE ncclAllGather, merge on
the device -- merged ids and the full unlabelled_confidence vector against the oracle's selection
(/root/reference/active_learning.py:705-715 restated in oracle/reference_np.py).

* single process, every visible GPU (als_comm_init_all / als_pool_select_global_all -- the reference's own process
  model, active_learning.py:221,277); with one GPU this still runs the NCCL all-gather and the merge kernel at world 1;
* one process per GPU under torchrun (als_comm_init_rank / als_pool_select_global) when >= 2 GPUs are visible:
  run with `gpurun --gpus 2` / `--gpus 8`; tests/multi_rank_worker.py is the per-rank program."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-5


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return t


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cases(n_pool):
    rng = np.random.default_rng(7)
    unl_some = np.sort(rng.choice(n_pool, max(1, n_pool - n_pool // 5), replace=False))
    return [
        ("all_unlabelled", np.arange(n_pool), 5),
        ("subset", unl_some, 7),
        ("k_ge_m", unl_some[:4], 50),
        ("shuffled", rng.permutation(unl_some), 3),
        ("k_zero", unl_some, 0),
        ("empty", np.zeros(0, np.int64), 4),
    ]


def test_single_process_all_gpus(torch):
    from oracle import reference_np as R, synth
    from semanticsegmentationactivelearning_b200 import Scorer, shard_bounds
    world = torch.cuda.device_count()
    N, H, W, C = 8 * world + 3, 16, 24, 19
    x = synth.synth_logits(1, 0, N, H, W, C)
    conf = R.scatter_scores(N, [(R.score_pool(x, "entropy"), np.arange(N))])
    scorers = [Scorer(d) for d in range(world)]
    try:
        Scorer.comm_init_all(scorers)
        shards = [shard_bounds(N, r, world) for r in range(world)]
        for r, sc in enumerate(scorers):
            lo, hi = shards[r]
            with torch.cuda.device(r):
                sc.pool_begin(N)
                # two batches per rank, shuffled inside the shard, like sess.run would hand them out (:697-700)
                order = np.random.default_rng(r).permutation(np.arange(lo, hi))
                for part in np.array_split(order, 2):
                    if part.size:
                        sc.pool_score_batch(torch.from_numpy(x[part]).cuda(r), part, "entropy")
        for name, unl, k in _cases(N):
            ids, u = Scorer.pool_select_global_all(scorers, unl, k, shards)
            want_ids, want_u = R.select_lowest_total_order(conf, unl, k)
            np.testing.assert_allclose(u, want_u, rtol=RTOL, err_msg=name)
            assert sorted(ids.tolist()) == sorted(want_ids.tolist()), name
            # ... and every rank's pool vector is complete afterwards
        for r, sc in enumerate(scorers):
            with torch.cuda.device(r):
                np.testing.assert_allclose(sc.pool_scores(N), conf, rtol=RTOL)
        # a selection larger than the one-launch limit (1024 survivors): second ranking kernel in both phases
        big = [Scorer(d) for d in range(world)]
        try:
            Scorer.comm_init_all(big)
            NB = 2600 + world
            bshards = [shard_bounds(NB, r, world) for r in range(world)]
            for r, sc in enumerate(big):
                with torch.cuda.device(r):
                    sc.pool_begin(NB)                     # nothing scored: every confidence is 0.0, ties go to the lower id
            ids, u = Scorer.pool_select_global_all(big, np.arange(NB)[::-1].copy(), 1500, bshards)
            assert ids.tolist() == list(range(1500)) and np.all(u == 0)
        finally:
            for sc in big:
                sc.close()
        # shards that do not tile the pool are an error, not a silently short selection
        if world > 1:
            bad = list(shards)
            bad[0] = (shards[0][0], shards[0][1] - 1)
            with pytest.raises(ValueError):
                Scorer.pool_select_global_all(scorers, np.arange(N), 3, bad)
    finally:
        for sc in scorers:
            sc.close()


def test_unvisited_examples_in_a_shard_are_selected_first(torch):
    """:685 / :701-702 -- a rank that did not visit some of the examples it owns leaves them at 0.0."""
    from semanticsegmentationactivelearning_b200 import Scorer
    world = torch.cuda.device_count()
    scorers = [Scorer(d) for d in range(world)]
    try:
        Scorer.comm_init_all(scorers)
        N = 4 * world
        shards = [(4 * r, 4 * r + 4) for r in range(world)]
        for r, sc in enumerate(scorers):
            with torch.cuda.device(r):
                sc.pool_begin(N)
                x = torch.full((3, 8, 8, 19), 0.0, device="cuda:%d" % r)
                x[..., 0] = 3.0 + r                                   # confident pixels: score well above 0
                sc.pool_score_batch(x, np.arange(4 * r, 4 * r + 3), "confidence")   # example 4r+3 never visited
        ids, u = Scorer.pool_select_global_all(scorers, np.arange(N), world, shards)
        assert sorted(ids.tolist()) == [4 * r + 3 for r in range(world)]
        assert np.all(u[ids] == 0) and np.all(np.delete(u, ids) > 0.5)
    finally:
        for sc in scorers:
            sc.close()


@pytest.mark.parametrize("world", [2, 8])
def test_torchrun_ranks_match_oracle(torch, world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (run with gpurun --gpus %d)" % (world, world))
    out = tmp_path / "result.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_rank_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(out.read_text())
    assert res["world"] == world and res["checked_cases"] >= 5 and res["ok"], res
