#!/usr/bin/env python
"""Pool-scoring benchmark (BASELINE.json metric: pool pixels scored / s, + HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]

One "step" = one pass of the hot path over the whole synthetic pool of the workload:
als_pool_begin -> als_pool_score_batch per chunk -> als_pool_select; when N > 1 the selection is
als_pool_select_global: per-GPU candidates on the device, ONE ncclAllGather, merge on the device.  Default workload = BASELINE.json configs[1] (ENet MC-dropout T=8
variance, 2975 images @512x1024, C=19, fp32).  Multi-GPU is weak scaling: every rank scores its
own 2975-image shard of an N x 2975 pool; the only exchange is the candidate/score all-gather.

The pool does not fit HBM (948 GB of logits per shard), so storage is aliased: `resident`
distinct images live in HBM and pool image n reads slot n % resident.  Every byte scored comes
from HBM (the resident set is hundreds of times the 126 MB L2).

Prints ONE JSON line (see README / DESIGN.md section "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: per-GPU pool N, T, H, W, C, measure, resident images, chunk images, description
    "cfg1": dict(N=64, T=1, H=512, W=1024, C=19, measure="entropy", resident=64, chunk=64,
                 desc="ENet Cityscapes 19-class entropy, pool 64 @512x1024"),
    "cfg2": dict(N=2975, T=8, H=512, W=1024, C=19, measure="variance", resident=175, chunk=175,
                 desc="ENet MC-dropout T=8 variance, pool 2975 @512x1024, C=19"),
    "cfg3": dict(N=372, T=1, H=1024, W=2048, C=19, measure="margin", resident=372, chunk=124,
                 desc="full-res 1024x2048 margin, 2975-image pool sharded 8 ways (372 per GPU)"),
    "cfg4": dict(N=4096, T=1, H=480, W=640, C=6, measure="entropy", resident=4096, chunk=512,
                 desc="Freiburg Forest 6-class entropy @480x640, pool 4096"),
    "cfg5": dict(N=2250, T=16, H=512, W=1024, C=66, measure="variance", resident=50, chunk=50,
                 desc="Mapillary Vistas 66-class MC-dropout T=16, 18000-image pool sharded 8 ways (2250 per GPU)"),
    # streamed Monte-Carlo accumulation (als_mc_*): the T samples of a chunk are handed over one at a time, the Welford
    # state (C+1 floats per pixel) stays in HBM -- [T,...] never has to exist at once (here it does, to have samples to feed)
    "cfg2s": dict(N=2975, T=8, H=512, W=1024, C=19, measure="variance", resident=175, chunk=175, stream=True,
                  desc="ENet MC-dropout T=8 variance, samples STREAMED one at a time (als_mc_*), pool 2975 @512x1024, C=19"),
    # training-path call site (:229-275): the same pass also writes pseudo_confidence f32, pseudo_label u8 and
    # pseudo_mask u8 (+6 B/pixel of writes); batches of 8 like params["batch_size"], and of 64
    "train8": dict(N=512, T=1, H=512, W=1024, C=19, measure="entropy", resident=512, chunk=8, maps=True,
                   desc="PseudoAnnotation scope with per-pixel outputs (conf+label+mask), batches of 8 @512x1024, C=19"),
    "train64": dict(N=512, T=1, H=512, W=1024, C=19, measure="entropy", resident=512, chunk=64, maps=True,
                    desc="PseudoAnnotation scope with per-pixel outputs (conf+label+mask), batches of 64 @512x1024, C=19"),
    # fused classifier head: the scorer is fed the `Final` layer's INPUT [N,h,w,16]
    # (16 B per output pixel instead of 4*C) and runs the 16->C transposed convolution on the tensor cores itself
    "cfg1h": dict(N=2975, T=1, H=512, W=1024, C=19, measure="entropy", resident=2975, chunk=425, head=True,
                  desc="fused Final head + entropy, pool 2975 @512x1024 (features 256x512x16), C=19"),
    "cfg2h": dict(N=2975, T=8, H=512, W=1024, C=19, measure="variance", resident=340, chunk=85, head=True,
                  desc="fused Final head + MC-dropout T=8 variance, pool 2975 @512x1024 (features 8x256x512x16), C=19"),
    "cfg3h": dict(N=372, T=1, H=1024, W=2048, C=19, measure="margin", resident=372, chunk=124, head=True,
                  desc="fused Final head + margin @1024x2048 (features 512x1024x16), 372 images per GPU"),
    "cfg4h": dict(N=4096, T=1, H=480, W=640, C=6, measure="entropy", resident=4096, chunk=512, head=True,
                  desc="fused Final head + entropy @480x640 (features 240x320x16), C=6, pool 4096"),
}
K_SELECT = 50          # conf/enet_cityscapes_active_learning.json:59
SEED = 20191013


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--pool", type=int, default=0, help="override the per-GPU pool size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget for the CPU baseline sample")
    return ap.parse_args()


def load_traffic(workload: str, dtype: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (or None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)[workload]
        if t.get("dtype") != dtype:
            return None
        return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(w, seconds: float, dtype: str):
    """The CPU restatement of the reference graph (oracle/reference_torch.py, all host threads) on a
    bounded sample of the same workload.  Returns (pixels/s, cores, sample description, ids check)."""
    import torch
    from oracle import reference_torch as RT, synth
    cores = RT.set_threads()
    T, H, W, C = w["T"], w["H"], w["W"], w["C"]
    n_img = 2
    if w.get("head"):
        x, kern = _cpu_head_inputs(w, n_img)
        run = lambda xx: RT.score_pool(RT.final_head(xx, kern), w["measure"])
    else:
        x = torch.from_numpy(synth.synth_logits(T, 0, n_img, H, W, C, seed=SEED, squeeze_t=False))
        if T == 1:
            x = x[0]
        run = lambda xx: RT.score_pool(xx, w["measure"])
    run(x[:1] if T == 1 else x[:, :1])               # warm-up
    t0 = time.perf_counter()
    done = 0
    while True:
        run(x)
        done += n_img
        el = time.perf_counter() - t0
        if el >= seconds or done >= 64:
            break
    rate = done * H * W / el
    return rate, cores, "%d images x T=%d @%dx%dx%d, %s, torch CPU op-by-op, %.1f s" % (done, T, H, W, C, w["measure"], el)


def _cpu_head_inputs(w, n_img):
    """Random-init `Final` input and kernel for the CPU legs of the fused-head workloads."""
    import torch
    g = torch.Generator().manual_seed(SEED)
    shape = (n_img, w["H"] // 2, w["W"] // 2, 16)
    x = torch.randn(shape if w["T"] == 1 else (w["T"],) + shape, generator=g)
    kern = torch.from_numpy((0.4 * np.random.default_rng(SEED).standard_normal((3, 3, w["C"], 16))).astype(np.float32))
    return x, kern


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (restated op by op; TF 1.13 is
    not installable), all host threads, bounded sample per step.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import reference_torch as RT, synth
    w = dict(WORKLOADS[args.workload])
    cores = RT.set_threads()
    T, H, W, C = w["T"], w["H"], w["W"], w["C"]
    n_img = 2                                   # images per step (bounded sample of the pool)
    kern = None
    if w.get("head"):
        x, kern = _cpu_head_inputs(w, n_img)
    else:
        x = torch.from_numpy(synth.synth_logits(T, 0, n_img, H, W, C, seed=SEED, squeeze_t=False))
        if T == 1:
            x = x[0]
    unl = np.arange(n_img)
    for _ in range(max(args.warmup, 1)):
        RT.rank_confidence(x, unl, 1, w["measure"], batch_size=8, head_kernel=kern)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        RT.rank_confidence(x, unl, 1, w["measure"], batch_size=8, head_kernel=kern)
    el = time.perf_counter() - t0
    rate = args.steps * n_img * H * W / el / 1e9
    sample = "%d images x T=%d @%dx%dx%d per step (bounded sample of the %d-image pool), %s" % (
        n_img, T, H, W, C, w["N"], w["measure"])
    line = {
        "impl": "reference", "metric": "pool pixels scored/s", "value": rate, "unit": "Gpix/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "k": K_SELECT, "sample": sample},
        "cpu_baseline": {"value": rate, "unit": "Gpix/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "TensorFlow 1.13.2 cannot be installed here; this is the op-by-op torch-CPU restatement of "
                "active_learning.py:239-263,682-715 (oracle/reference_torch.py) on all host threads",
    }
    print(json.dumps(line), flush=True)


def pin_to_gpu_cpus(local_rank: int) -> dict:
    """Bind this rank to the CPUs next to its GPU's PCIe root (pinned staging buffers are then allocated on that
    node).  Best effort: containers often expose one NUMA node and the same CPU list for every GPU."""
    info = {"numa_node": None, "cpus": None, "pinned": False}
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=10).stdout.strip()       # 00000000:1B:00.0
        dom, rest = out.split(":", 1)
        bus = dom[-4:] + ":" + rest                                                                # 0000:1B:00.0
        dev = "/sys/bus/pci/devices/" + str(bus).lower()
        with open(dev + "/numa_node") as f:
            info["numa_node"] = int(f.read().strip())
        with open(dev + "/local_cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                if "-" in part:
                    a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = sorted(cpus & allowed)
        info["cpus"] = "%d local of %d allowed" % (len(use), len(allowed))
        if use and len(use) < len(allowed):
            os.sched_setaffinity(0, use)
            info["pinned"] = True
    except Exception as e:   # pragma: no cover
        info["error"] = str(e)[:80]
    return info


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from semanticsegmentationactivelearning_b200 import Scorer, comm_init_torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the pool-scoring path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    affinity = pin_to_gpu_cpus(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    w = dict(WORKLOADS[args.workload])
    if args.pool:
        w["N"] = args.pool
    T, H, W, C, N = w["T"], w["H"], w["W"], w["C"], w["N"]
    measure = w["measure"]
    dtype = "float32" if args.dtype == "f32" else "bfloat16"
    es = 4 if args.dtype == "f32" else 2
    resident = min(w["resident"], N)
    chunk = min(w["chunk"], resident)
    P = H * W
    id0 = rank * N                                    # this rank's global example ids: [id0, id0 + N)

    sc = Scorer(local_rank)
    if world > 1:
        comm_init_torch(sc)                           # the library's own NCCL communicator (csrc/comm.cu)
    head = bool(w.get("head"))
    n_chunk_bufs = resident // chunk
    if head:
        # resident `Final`-layer inputs [chunk, H/2, W/2, 16] (random-init features and kernel, seeded)
        if args.dtype != "f32":
            raise SystemExit("the fused-head workloads are fp32")
        gen = torch.Generator(device=dev); gen.manual_seed(SEED + rank)
        head_kernel = (0.4 * np.random.default_rng(SEED).standard_normal((3, 3, C, 16))).astype(np.float32)
        sc.prepare_head(head_kernel)
        bufs = []
        for i in range(n_chunk_bufs):
            f = torch.randn((chunk, H // 2, W // 2, 16), generator=gen, device=dev, dtype=torch.float32)
            f *= 0.3 + torch.rand((chunk, 1, 1, 1), generator=gen, device=dev)
            if T > 1:
                # T dropout forward passes: channel-wise keep masks (spatial_dropout, extra_ops.py:137-151) + noise
                ft = torch.empty((T,) + tuple(f.shape), device=dev, dtype=torch.float32)
                for t in range(T):
                    keep = (torch.rand((chunk, 1, 1, 16), generator=gen, device=dev) > 0.1).float() / 0.9
                    ft[t] = f * keep
                    ft[t] += 0.05 * torch.randn(f.shape, generator=gen, device=dev)
                f = ft
            bufs.append(f)
    else:
        # resident logits [T, resident, H, W, C]; chunks are dense [T, chunk, ...] tensors of their own
        bufs = [sc.synth_logits(T, id0 + i * chunk, chunk, H, W, C, dtype=dtype, seed=SEED, squeeze_t=False)
                for i in range(n_chunk_bufs)]
    torch.cuda.synchronize()
    # pool chunk list: (buffer, first local id, count)
    chunks = []
    n0 = 0
    while n0 < N:
        nb = min(chunk, N - n0)
        buf = bufs[(n0 // chunk) % n_chunk_bufs]
        if nb < chunk and head:
            buf = buf[:nb] if T == 1 else buf[:, :nb].contiguous()
        elif nb < chunk:   # ragged tail: a dense [T, nb] tensor of its own
            buf = sc.synth_logits(T, id0 + n0, nb, H, W, C, dtype=dtype, seed=SEED, squeeze_t=False)
        chunks.append((buf, n0, nb))
        n0 += nb
    unl_global = np.arange(world * N, dtype=np.int64)   # every pool image is unlabelled

    ev_pairs = []
    streamed = bool(w.get("stream"))
    maps = bool(w.get("maps"))
    out_bytes_pix = 6 if maps else 0
    map_outs = {}

    sparse_events = maps and chunk <= 16
    one_call = (world == 1 and len(chunks) == 1 and not head and not maps and not streamed)
    sel_pairs = []
    lib_ms = []
    if one_call:
        sc.enable_timing(True)
        one_in = chunks[0][0] if T > 1 else chunks[0][0][0]

    def step(record: bool):
        if one_call:
            # a pool that is ONE resident tensor: the whole closure (:682-715) is one library call (als_rank_pool);
            # the scoring launch is timed from the host-visible events around the call minus nothing -- see `full` below
            out = sc.rank_pool(one_in, unl_global, K_SELECT, measure, num_examples=N)
            if record:   # the library's own CUDA events around the scoring launch (als_ctx_enable_timing)
                lib_ms.append((sc.last_scoring_ms(), chunks[0][2]))
            return out
        # every rank keeps the full-size confidence vector and scores the ids it owns: [id0, id0 + N)
        sc.pool_begin(world * N)
        for ci, (buf, first, nb) in enumerate(chunks):
            idx = np.arange(id0 + first, id0 + first + nb, dtype=np.int64)
            # launches of a few tens of microseconds (batches of 8): an event pair around EVERY launch puts stream
            # operations between consecutive kernels and breaks their programmatic-dependent-launch overlap (61 us per
            # call instead of 58), so these workloads take the launch duration from the CUDA events around the whole
            # timed region divided by the number of launches (see `full` below)
            timed = record and not sparse_events
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if head:
                sc.pool_score_features_batch(buf, idx, measure)
            elif streamed:
                sc.mc_begin((nb, H, W, C), dtype)
                for t in range(T):
                    sc.mc_add_sample(buf[t])
                sc.mc_finish(measure, batch_indices=idx)
            elif maps:
                map_outs[nb] = sc.pseudo_annotation(buf if T > 1 else buf[0], measure, 0.9, out=map_outs.get(nb))
            else:
                sc.pool_score_batch(buf, idx, measure)
            if timed:
                e1.record()
                ev_pairs.append((e0, e1, nb))
        if maps:
            return None, None                  # the training-path call site has no selection (:229-275)
        # :705-715; N > 1: local candidates -> one ncclAllGather -> merge, all on the device, one D2H
        if record:
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
        out = sc.pool_select_global(unl_global, K_SELECT, (id0, id0 + N), N)
        if record:
            s1.record()
            sel_pairs.append((s0, s1))
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: at least W (>= 3) untimed steps, and for workloads whose step is a fraction of a millisecond at least 50 ms
    # of them, so that clocks and caches are where a real pool pass would find them (identical count on every rank)
    warm_steps = max(args.warmup, 3)
    t_w = time.perf_counter()
    for _ in range(warm_steps):
        step(False)
    torch.cuda.synchronize()
    per = (time.perf_counter() - t_w) / warm_steps
    extra = 0 if world > 1 else max(0, min(200, int(0.05 / max(per, 1e-6)) - warm_steps))
    for _ in range(extra):
        step(False)
    warm_steps += extra
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sc.launch_count
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for _ in range(args.steps):
        ids, conf = step(True)
    t_end.record()
    barrier()
    ms_total = t_start.elapsed_time(t_end)
    launches = sc.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * N * P / (ms_step * 1e-3) / 1e9

    # ---- outside the timed region: the merged ids against the oracle's selection on the same score vector ----
    ids_check = None
    if not maps:
        full_scores = sc.pool_scores(world * N)       # after the exchange every rank holds the whole vector
        if rank == 0:
            # /root/reference/active_learning.py:705-714 verbatim (NumPy only; the oracle package stays out of this path)
            want_u = full_scores[unl_global]
            k_sel = np.minimum(len(unl_global), K_SELECT)
            want_ids = unl_global[np.argpartition(want_u, k_sel)[:k_sel]]
            # the synthetic pool aliases `resident` distinct images, so scores repeat and the k-th boundary falls inside
            # a group of EQUAL scores, where np.argpartition's choice is arbitrary: compare the selected score multiset
            # and require every id strictly below the k-th score on both sides
            kth = np.sort(full_scores[want_ids])[-1] if len(want_ids) else np.float32(0)
            below = lambda sel: sorted(int(i) for i in sel if full_scores[i] < kth)
            same = (sorted(full_scores[ids].tolist()) == sorted(full_scores[want_ids].tolist()) and below(ids) == below(want_ids))
            ids_check = {"ids_match_oracle": bool(same),
                         "unlabelled_confidence_match": bool(np.array_equal(conf, want_u)),
                         "scores_finite": bool(np.all(np.isfinite(full_scores))),
                         "distinct_scores": int(len(np.unique(full_scores))),
                         "how": "np.argpartition exactly as active_learning.py:705-714 on the GPU path's own float32 "
                                "score vector of the last timed pass; equal ids below the k-th score, equal score multiset "
                                "(ties at the k-th score excused: the synthetic pool aliases its resident images)"}

    # dominant kernel: score_tiles_kernel, one launch per chunk (the 2 us finalize launch rides along)
    full = [(a.elapsed_time(b), nb) for a, b, nb in ev_pairs if nb == chunk] or [(a.elapsed_time(b), nb) for a, b, nb in ev_pairs]
    if one_call:
        full = lib_ms
    if sparse_events:   # back-to-back launches: timed region / launches (an upper bound of the kernel's own duration)
        full = [(ms_total / (args.steps * len(chunks)), chunk)]
    avg_ms = sum(m for m, _ in full) / len(full)
    bytes_launch = full[0][1] * P * (T * C * es + out_bytes_pix)
    if streamed:   # per chunk: T sample launches + finish: logits read + state written T times, read T-1 times + once by finish
        bytes_launch = full[0][1] * P * (T * C * es + 2 * T * (C + 1) * 4)
    if head:
        bytes_launch = full[0][1] * T * (P // 4) * 64      # 16 fp32 channels per INPUT pixel = 16 B per output pixel (and sample)
    peak, peak_src = load_peaks()
    achieved = bytes_launch / (avg_ms * 1e-3) / 1e9
    kernel_share = (sum(m for m, _ in lib_ms) if one_call else sum(a.elapsed_time(b) for a, b, _ in ev_pairs)) / ms_total
    if sparse_events:
        kernel_share = 1.0

    # ---- per-rank view (N > 1): a step ends with a collective, so it lasts as long as the SLOWEST rank's scoring plus the
    # exchange; rank 0's kernel share therefore also contains its wait for slower GPUs (they differ by 1-2 % under the
    # power cap).  Report every rank's own scoring time and its time in the selection call (wait + exchange + merge).
    per_rank = None
    if world > 1:
        mine = torch.tensor([sum(a.elapsed_time(b) for a, b, _ in ev_pairs) / args.steps,
                             (sum(a.elapsed_time(b) for a, b in sel_pairs) / max(len(sel_pairs), 1))], device=dev)
        allr = torch.empty(world * 2, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.view(world, 2).cpu().numpy()
        per_rank = {"scoring_ms_per_step": [round(float(v), 3) for v in allr[:, 0]],
                    "select_call_ms": [round(float(v), 3) for v in allr[:, 1]],
                    "slowest_rank_scoring_share_of_step": float(allr[:, 0].max() / ms_step),
                    "exchange_ms_on_slowest_rank": float(allr[int(allr[:, 0].argmax()), 1]),
                    "note": "select_call_ms on a fast rank includes waiting for the slowest rank at the all-gather; on the "
                            "slowest rank it is the cost of the exchange itself (candidates + one ncclAllGather + merge + results)"}

    # ---- diagnostic: the scoring kernel alone, launched back to back on one chunk (host latency hidden) ----
    burst_buf = chunks[0][0]
    burst_in = burst_buf if (T > 1 or head) else burst_buf[0]     # (fused-head buffers are [N,..] at T = 1 already)
    burst_out = torch.empty(chunks[0][2], dtype=torch.float64, device=dev)
    n_burst = max(4, min(40, int(0.25 / max(avg_ms * 1e-3, 1e-5))))
    burst_fn = (lambda: sc.score_features(burst_in, measure, out=burst_out)) if head else \
               (lambda: sc.score(burst_in, measure, out=burst_out))
    if streamed:
        def burst_fn():
            sc.mc_begin((chunks[0][2], H, W, C), dtype)
            for t in range(T):
                sc.mc_add_sample(burst_buf[t])
            sc.mc_finish(measure)
    for _ in range(3):
        burst_fn()
    torch.cuda.synchronize()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    for _ in range(n_burst):
        burst_fn()
    b1.record()
    torch.cuda.synchronize()
    burst_ms = b0.elapsed_time(b1) / n_burst
    burst = {"launches": n_burst, "avg_ms": burst_ms, "GBps": bytes_launch / (burst_ms * 1e-3) / 1e9,
             "Gpix_s": chunks[0][2] * P / (burst_ms * 1e-3) / 1e9,
             "outputs": "scores only",
             "note": "score+finalize launched back to back on one resident chunk, outside the timed region"}

    # ---- e2e: the public API with HOST (pinned) inputs, H2D inside the timed region ----------------
    from semanticsegmentationactivelearning_b200 import rank_confidence

    def measure_e2e(kind: str):
        """kind: "native" = what the workload feeds (f32/bf16 logits, or features for the *h workloads) in PINNED memory;
        "pageable" = the same logits as a plain NumPy array (the library stages it with its parallel bounce pipeline);
        "bf16" = the same logits rounded to bfloat16; "features" = the `Final` layer's input + fused head."""
        use_head = head or kind == "features"
        e_es = 2 if (kind == "bf16" or args.dtype == "bf16") else 4
        per_img = T * (P // 4) * 64 if use_head else T * P * C * e_es
        bsz = max(1, min(8, int((2 << 30) // per_img)))              # images per sess.run-like batch (<= 8, :689)
        n_e2e = max(bsz, min(N, int((8 << 30) // per_img) // bsz * bsz))   # ~8 GB of input per step
        if use_head:
            if head:
                src = bufs[0][:bsz] if T == 1 else bufs[0][:, :bsz]
                kern = head_kernel
            else:
                g2 = torch.Generator(device=dev); g2.manual_seed(SEED + 17 + rank)
                src = torch.randn(((bsz, H // 2, W // 2, 16) if T == 1 else (T, bsz, H // 2, W // 2, 16)), generator=g2,
                                  device=dev, dtype=torch.float32)
                kern = (0.4 * np.random.default_rng(SEED).standard_normal((3, 3, C, 16))).astype(np.float32)
            host = torch.empty(tuple(src.shape), dtype=torch.float32).pin_memory()
            host.copy_(src)
        else:
            kern = None
            tdt = torch.bfloat16 if e_es == 2 else torch.float32
            host = torch.empty((T, bsz, H, W, C), dtype=tdt).pin_memory()
            host.copy_(bufs[0][:, :bsz].to(tdt))
        torch.cuda.synchronize()
        if kind == "pageable":     # what sess.run hands out (:697-698): a plain NumPy array, not page-locked
            host = host.numpy().copy()
        one = host if (T > 1 or use_head) else host[0]

        def e2e_step():
            batches = ((one, np.arange(i, i + bsz, dtype=np.int64)) for i in range(0, n_e2e, bsz))
            return rank_confidence(batches, np.arange(n_e2e, dtype=np.int64), K_SELECT, measure, num_examples=n_e2e,
                                   scorer=sc, head_kernel=kern)

        e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k_e2e = max(2, min(args.steps, 5))
        for _ in range(k_e2e):
            e2e_step()
        e1.record()
        barrier()
        ms_e = e0.elapsed_time(e1) / k_e2e
        if world > 1:
            t = torch.tensor([ms_e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = float(t.item())
        h2d = int(n_e2e * per_img + n_e2e * 8 + n_e2e * 8)
        return {"value": world * n_e2e * P / (ms_e * 1e-3) / 1e9, "unit": "Gpix/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(min(K_SELECT, n_e2e) * 8 + n_e2e * 4),
                "pool_images_per_step": n_e2e, "batch_images": bsz, "ms_per_step": ms_e,
                "h2d_GBps_per_gpu": h2d / (ms_e * 1e-3) / 1e9,
                "input": ("Final-layer features f32 [%sB,%d,%d,16] + fused head" % ("T," if T > 1 else "", H // 2, W // 2)) if use_head
                         else ("%s logits [%sB,%d,%d,%d]%s" % ("bf16" if e_es == 2 else "f32", "T," if T > 1 else "", H, W, C,
                                                            " in pageable memory (NumPy)" if kind == "pageable" else "")),
                "note": "rank_confidence() fed pinned host batches like sess.run (:697-700); PCIe-bound"}

    e2e = None
    e2e_alt = None
    if not args.no_e2e and not maps and not streamed:
        e2e = measure_e2e("native")
        # the two byte-reducing inputs the library accepts for the same pool pass (same metric, fewer bytes over PCIe)
        e2e_alt = {}
        if not head and args.dtype == "f32":
            e2e_alt["pageable_numpy_logits"] = measure_e2e("pageable")
            e2e_alt["bf16_logits"] = measure_e2e("bf16")
        if not head and Scorer.head_supported(C, measure, T):
            e2e_alt["fused_head_features"] = measure_e2e("features")
            if world == 1:
                sc.pool_begin(N)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            rate, cores, sample = cpu_reference_rate(w, args.cpu_seconds, dtype)
            cpu = {"value": rate / 1e9, "unit": "Gpix/s", "cores": cores, "kind": "port", "sample": sample}
        if head:
            desc = sc.describe_head_launch(T, measure)
        elif streamed:
            desc = sc.describe_launch(dtype, T, chunk, H, W, C, measure)
            desc["kernel"] = "mc_update_kernel x T + mc_finish_kernel (streamed samples; same tile plan as " + desc["kernel"] + ")"
        else:
            desc = sc.describe_launch(dtype, T, chunk, H, W, C, measure)
        line = {
            "metric": "pool pixels scored/s", "value": value, "unit": "Gpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm_steps, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": args.workload + ": " + w["desc"], "pool_images_per_gpu": N, "T": T, "H": H, "W": W,
                       "C": C, "measure": measure, "k": K_SELECT,
                       "outputs": "scores + pseudo_confidence f32 + pseudo_label u8 + pseudo_mask u8" if maps else "scores",
                       "resident_images": resident, "chunk_images": chunk,
                       "l2": "inputs larger than L2 (resident set %.1f GB, aliased over the pool)" % (
                           resident * T * ((P // 4) * 64 if head else P * C * es) / 1e9),
                       "kernel": desc},
            "clocks": clocks, "e2e": e2e, "e2e_alt": e2e_alt, "gpu_launches": int(launches * world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(args.workload, args.dtype) if not args.pool else None,
                         "traffic_source": "profiles/ncu_traffic.json (dram__bytes_read+write of one ncu --set full launch)",
                         "peak_source": peak_src, "kernel": desc["kernel"],
                         "bytes_per_launch": bytes_launch, "avg_launch_ms": avg_ms, "launches_timed": len(full),
                         "kernel_share_of_step": kernel_share, "per_rank": per_rank, "kernel_burst": burst},
            "cpu_baseline": cpu,
            "selected_ids_head": [int(i) for i in ids[:5]] if ids is not None else None,
            "ids_check": ids_check, "cpu_affinity": affinity,
        }
        if head:
            # the fused kernel is not HBM bound: say what limits it and how busy the tensor pipe is
            from semanticsegmentationactivelearning_b200.acquisition import head_mma_flops_per_pixel
            fl = head_mma_flops_per_pixel(C)
            line["roofline"]["limiter"] = ("SM issue slots + MUFU (19 ex2 + lg2 + rcp per pixel); ncu: XU pipe 71 %, tensor pipe 63 %, "
                                           "issue 54 %, DRAM 25 % (profiles/r01_ncu_full_cfg1h.txt)")
            line["roofline"]["tensor"] = {"kind": "tf32, 3-product split (hi*lo + lo*hi + hi*hi)", "flops_per_pixel": fl,
                                          "achieved_tflops": fl * T * chunks[0][2] * P / (avg_ms * 1e-3) / 1e12,
                                          "nominal_peak_tflops": 1100.0}
            line["config"]["input"] = "Final-layer input [%sN,%d,%d,16] fp32 + kernel [3,3,%d,16]; logits never materialised" % (
                "" if T == 1 else "T=%d," % T, H // 2, W // 2, C)
            if T > 1:
                line["roofline"]["limiter"] = ("MUFU + issue slots in the Welford epilogue (T x (19 ex2 + rcp) per pixel) and the "
                                               "loader warps (two feature rows per accumulator); not HBM bound")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
