"""Per-rank program of tests/test_gpu_multi.py::test_torchrun_ranks_match_oracle (launched by torchrun, one rank per GPU).

Every rank scores its shard of a synthetic pool on its GPU, then als_pool_select_global (csrc/comm.cu) merges the
candidates over NCCL.  Rank 0 compares the merged ids / the full confidence vector with the oracle
(/root/reference/active_learning.py:705-715 restated in oracle/reference_np.py) and writes the verdict as JSON."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_path):
    import torch
    import torch.distributed as dist
    from oracle import reference_np as R, synth
    from semanticsegmentationactivelearning_b200 import Scorer, comm_init_torch, rank_confidence_sharded_device, shard_bounds

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")            # only carries the 128-byte NCCL id; the exchange itself is the library's
    sc = Scorer(local)
    comm_init_torch(sc)
    ok, checked, errors = True, 0, []
    for measure, T, N, H, W, C in (("entropy", 1, 37, 16, 24, 19), ("variance", 4, 8 * world + 1, 8, 16, 19),
                                   ("margin", 1, 3, 8, 8, 6)):          # 3 examples: ranks >= 3 own nothing
        lo, hi = shard_bounds(N, rank, world)
        sc.pool_begin(N)
        if hi > lo:
            x = sc.synth_logits(T, lo, hi - lo, H, W, C, squeeze_t=(T == 1))
            sc.pool_score_batch(x, np.arange(lo, hi), measure)
        rng = np.random.default_rng(3)
        for unl, k in ((np.arange(N), 5), (np.sort(rng.choice(N, max(1, N // 2), replace=False)), 4), (np.arange(N)[::-1].copy(), N + 2)):
            ids, u = rank_confidence_sharded_device(sc, unl, k, (lo, hi))
            full = sc.pool_scores(N)
            if rank == 0:
                ref_x = synth.synth_logits(T, 0, N, H, W, C)
                conf = R.scatter_scores(N, [(R.score_pool(ref_x, measure), np.arange(N))])
                want_ids, want_u = R.select_lowest_total_order(conf, unl, k)
                good = (np.allclose(u, want_u, rtol=1e-5, atol=0) and sorted(ids.tolist()) == sorted(want_ids.tolist())
                        and np.allclose(full, conf, rtol=1e-5, atol=0))
                if not good:
                    errors.append((measure, N, int(k), ids.tolist(), want_ids.tolist()))
                ok = ok and good
                checked += 1
            # every rank must hold the same answer
            box = [None] * world
            dist.all_gather_object(box, (ids.tolist(), u.tobytes()))
            if any(b != box[0] for b in box):
                ok = False
                errors.append(("ranks disagree", measure, N, int(k)))
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump({"world": world, "ok": bool(ok), "checked_cases": checked, "errors": errors[:5]}, f)
    sc.close()
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main(sys.argv[1])
