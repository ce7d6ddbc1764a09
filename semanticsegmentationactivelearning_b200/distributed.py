"""Pool sharded by image across GPUs: every rank scores its own images, only the per-rank
top-k candidates and the small score vector cross NVLink (ONE all-gather).

The reference runs the whole pool on GPU:0, eight images per sess.run
(/root/reference/active_learning.py:689-700); images are independent, so the pool shards with
no data-path collective.

Two forms of the same exchange:

* ``comm_init_torch`` + ``rank_confidence_sharded_device``: the product path.  The exchange lives behind the C ABI
  (csrc/comm.cu: ``als_comm_init_rank`` / ``als_pool_select_global``): candidates are selected, all-gathered with
  NCCL and merged on the device, one device->host copy.  torch.distributed is only used to hand the NCCL unique id
  around.  (Single-process multi-GPU: ``Scorer.comm_init_all`` / ``Scorer.pool_select_global_all``.)
* ``rank_confidence_sharded``: the same record exchange written against ``torch.distributed`` with host arrays, so
  that the sharding / padding / merge logic is testable with gloo on machines without a GPU.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

_PAD_ID = 1 << 62


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced blocks: the first n % world ranks own one image more."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def comm_init_torch(scorer, group=None) -> None:
    """Create the library's own NCCL communicator for ``scorer`` over the ranks of a torch.distributed group:
    rank 0 draws the NCCL unique id, torch broadcasts the 128 bytes, every rank joins (collective)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [scorer.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    scorer.comm_init(rank, world, box[0])


def rank_confidence_sharded_device(scorer, unlabelled, selection_size: int, shard: Tuple[int, int], max_shard: int = 0):
    """:705-715 for a pool sharded by image, on the device (csrc/comm.cu).  ``scorer`` holds a full-size pool vector
    (``pool_begin(num_examples)``) in which this rank has scored the example ids [shard[0], shard[1]); the shards of
    all ranks tile [0, num_examples).  Returns (low_conf_examples, unlabelled_confidence), identical on all ranks."""
    return scorer.pool_select_global(unlabelled, selection_size, shard, max_shard)


def _device_select(scorer):
    def select(keys: np.ndarray, ids: np.ndarray, k: int):
        import torch
        dev = torch.device("cuda", scorer.device)
        ok, oi = scorer.select_smallest(torch.from_numpy(np.ascontiguousarray(keys, np.float32)).to(dev),
                                        torch.from_numpy(np.ascontiguousarray(ids, np.int64)).to(dev), k)
        return ok.cpu().numpy(), oi.cpu().numpy()
    return select


def merge_candidates(keys: np.ndarray, ids: np.ndarray, k: int, select: Callable):
    """Final k of the gathered per-rank candidates; padding entries (id >= 2^62) are dropped."""
    mk, mi = select(keys, ids, k)
    keep = mi < _PAD_ID
    return mk[keep], mi[keep]


def rank_confidence_sharded(local_scores: np.ndarray, local_ids: np.ndarray, unlabelled, selection_size: int, *,
                            group=None, select: Optional[Callable] = None, scorer=None,
                            max_unlabelled_per_rank: Optional[int] = None):
    """Global selection from per-rank results.

    local_scores   f32[n_local]  this rank's per-image confidences (rounded to f32 as :700 does)
    local_ids      i64[n_local]  their global example ids
    unlabelled     the global unlabelled index array (same on every rank)
    Returns the reference tuple (low_conf_examples, unlabelled_confidence) (:715), identical on all ranks.
    `select(keys, ids, k) -> (keys, ids)` defaults to the CUDA block-radix select of `scorer`.

    Exchange: every rank contributes one byte record [count | k candidate keys | k candidate ids | scores | ids].
    With ``max_unlabelled_per_rank`` (an upper bound on any rank's number of unlabelled images, e.g. the largest shard
    size; the SAME value on every rank)
    that is ONE all-gather; without it the record is sent in two parts (counts + candidates first, then the scores
    padded to the largest count).  The payload is a few hundred KB at most: latency, not bandwidth.
    """
    import torch
    import torch.distributed as dist

    if select is None:
        if scorer is None:
            from .acquisition import default_scorer
            scorer = default_scorer()
        select = _device_select(scorer)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")

    unlabelled = np.asarray(unlabelled, dtype=np.int64)
    local_ids = np.asarray(local_ids, dtype=np.int64)
    local_scores = np.asarray(local_scores, dtype=np.float32)
    k = int(max(0, min(int(selection_size), unlabelled.size)))

    # candidates: this rank's k lowest among its unlabelled images
    if local_ids.size:
        order = np.argsort(local_ids, kind="stable")
        sorted_ids = local_ids[order]
        pos = np.clip(np.searchsorted(sorted_ids, unlabelled), 0, local_ids.size - 1)
        mine = sorted_ids[pos] == unlabelled
        my_unl = unlabelled[mine]
        my_conf = local_scores[order][pos[mine]]
    else:
        my_unl = np.zeros(0, np.int64)
        my_conf = np.zeros(0, np.float32)
    ck, ci = select(my_conf, my_unl, k) if (k > 0 and my_unl.size) else (np.zeros(0, np.float32), np.zeros(0, np.int64))
    pad = k - ck.size
    cand_k = np.concatenate([ck, np.full(pad, np.nan, np.float32)]).astype(np.float32)
    cand_i = np.concatenate([ci, _PAD_ID + rank * max(k, 1) + np.arange(pad, dtype=np.int64)]).astype(np.int64)

    def gather(record: np.ndarray) -> np.ndarray:
        """all-gather equal-sized byte records -> [world, len(record)] uint8"""
        out = torch.empty(world * record.size, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, torch.from_numpy(record).to(dev), group=group)
        return out.cpu().numpy().reshape(world, record.size)

    def score_part(width: int) -> np.ndarray:
        sc = np.zeros(width, np.float32); sc[:my_unl.size] = my_conf
        sid = np.full(width, -1, np.int64); sid[:my_unl.size] = my_unl
        return np.concatenate([sc.view(np.uint8), sid.view(np.uint8)])

    head = np.concatenate([np.asarray([my_unl.size], np.int64).view(np.uint8), cand_k.view(np.uint8), cand_i.view(np.uint8)])
    if max_unlabelled_per_rank is not None:
        width = int(max_unlabelled_per_rank)
        if my_unl.size > width:
            raise ValueError("max_unlabelled_per_rank=%d but this rank holds %d unlabelled images" % (width, my_unl.size))
        both = gather(np.concatenate([head, score_part(width)]))
        heads, parts = both[:, :head.size], both[:, head.size:]
    else:
        heads = gather(head)
        width = int(heads[:, :8].copy().view(np.int64).max())
        parts = gather(score_part(width)) if width > 0 else np.zeros((world, 0), np.uint8)
    gk = heads[:, 8:8 + 4 * k].copy().view(np.float32).reshape(-1)
    gi = heads[:, 8 + 4 * k:].copy().view(np.int64).reshape(-1)
    all_sc = parts[:, :4 * width].copy().view(np.float32).reshape(-1)
    all_id = parts[:, 4 * width:].copy().view(np.int64).reshape(-1)

    valid = all_id >= 0
    vid, vsc = all_id[valid], all_sc[valid]
    # unlabelled_confidence = confidence[unlabelled] (:705); unvisited examples keep 0.0 (:685)
    unlabelled_confidence = np.zeros(unlabelled.size, np.float32)
    if vid.size:
        o = np.argsort(vid, kind="stable")
        vid, vsc = vid[o], vsc[o]
        at = np.clip(np.searchsorted(vid, unlabelled), 0, vid.size - 1)
        hit = vid[at] == unlabelled
        unlabelled_confidence[hit] = vsc[at[hit]]

    if k == 0:
        return np.zeros(0, np.int64), unlabelled_confidence
    _, ids = merge_candidates(gk, gi, k, select)
    return ids, unlabelled_confidence
