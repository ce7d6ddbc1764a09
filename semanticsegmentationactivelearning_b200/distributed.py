"""Pool sharded by image across GPUs: every rank scores its own images, only the per-rank
top-k candidates and the small score vector cross NVLink (one all-gather each).

The reference runs the whole pool on GPU:0, eight images per sess.run
(/root/reference/active_learning.py:689-700); images are independent, so the pool shards with
no data-path collective.  One process per GPU (torchrun); NCCL on GPUs, gloo in CPU tests.
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
                            group=None, select: Optional[Callable] = None, scorer=None):
    """Global selection from per-rank results.

    local_scores   f32[n_local]  this rank's per-image confidences (rounded to f32 as :700 does)
    local_ids      i64[n_local]  their global example ids
    unlabelled     the global unlabelled index array (same on every rank)
    Returns the reference tuple (low_conf_examples, unlabelled_confidence) (:715), identical on all ranks.
    `select(keys, ids, k) -> (keys, ids)` defaults to the CUDA block-radix select of `scorer`.
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
    order = np.argsort(local_ids, kind="stable")
    pos = np.searchsorted(local_ids[order], unlabelled)
    pos = np.clip(pos, 0, max(local_ids.size - 1, 0))
    mine = (local_ids.size > 0) & (local_ids[order][pos] == unlabelled) if local_ids.size else np.zeros(unlabelled.size, bool)
    my_unl = unlabelled[mine]
    my_conf = local_scores[order][pos[mine]] if local_ids.size else np.zeros(0, np.float32)
    ck, ci = select(my_conf, my_unl, k) if (k > 0 and my_unl.size) else (np.zeros(0, np.float32), np.zeros(0, np.int64))
    pad = k - ck.size
    cand_k = np.concatenate([ck, np.full(pad, np.nan, np.float32)]).astype(np.float32)
    cand_i = np.concatenate([ci, _PAD_ID + rank * max(k, 1) + np.arange(pad, dtype=np.int64)]).astype(np.int64)

    # exchange 1: fixed-size candidate buffers (k * 12 bytes per rank)
    gk = torch.empty(world * k, dtype=torch.float32, device=dev)
    gi = torch.empty(world * k, dtype=torch.int64, device=dev)
    if k > 0:
        dist.all_gather_into_tensor(gk, torch.from_numpy(cand_k).to(dev), group=group)
        dist.all_gather_into_tensor(gi, torch.from_numpy(cand_i).to(dev), group=group)
    # exchange 2: every rank's (id, score) pairs for the histogram consumer (:781-784), padded to the max shard
    sizes = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, torch.tensor([my_unl.size], dtype=torch.int64, device=dev), group=group)
    mmax = int(sizes.max().item())
    sid = np.full(mmax, -1, np.int64); sid[:my_unl.size] = my_unl
    ssc = np.zeros(mmax, np.float32); ssc[:my_unl.size] = my_conf
    all_id = torch.empty(world * mmax, dtype=torch.int64, device=dev)
    all_sc = torch.empty(world * mmax, dtype=torch.float32, device=dev)
    if mmax > 0:
        dist.all_gather_into_tensor(all_id, torch.from_numpy(sid).to(dev), group=group)
        dist.all_gather_into_tensor(all_sc, torch.from_numpy(ssc).to(dev), group=group)
    all_id = all_id.cpu().numpy(); all_sc = all_sc.cpu().numpy()
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
    _, ids = merge_candidates(gk.cpu().numpy(), gi.cpu().numpy(), k, select)
    return ids, unlabelled_confidence
