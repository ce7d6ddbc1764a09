"""Host->device ceiling of the box: what a plain pinned cudaMemcpyAsync reaches when N ranks copy at the same time.

The end-to-end number of bench.py (rank_confidence() fed pinned host batches, like sess.run hands them out,
/root/reference/active_learning.py:697-700) is bound by this, not by the scoring kernels: compare the library's
`e2e.h2d_GBps_per_gpu` with the per-rank rate printed here for the same N.

    python profiles/h2d_ceiling.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 profiles/h2d_ceiling.py
Prints one JSON line on rank 0."""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 1 << 30
    res = {}
    for chunk_mb in (8, 64, 320, 1024):                    # 320 MB ~ one batch of 8 f32 logits images @512x1024x19
        n = chunk_mb << 20
        host = torch.empty(n, dtype=torch.uint8).pin_memory()
        devb = torch.empty(n, dtype=torch.uint8, device=dev)
        reps = max(4, nbytes * 4 // n)
        for _ in range(3):
            devb.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            devb.copy_(host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        gbs = reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
        t = torch.tensor([gbs], device=dev)
        if world > 1:
            lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            res["%dMB" % chunk_mb] = {"min_rank_GBps": float(lo.item()), "sum_GBps": float(sm.item())}
        else:
            res["%dMB" % chunk_mb] = {"min_rank_GBps": gbs, "sum_GBps": gbs}
        del host, devb
    if rank == 0:
        print(json.dumps({"what": "pinned host->device copy, all ranks at once", "n_gpus": world, "by_copy_size": res,
                          "cpus_allowed": len(os.sched_getaffinity(0))}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
