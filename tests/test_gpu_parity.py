"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.

Tolerances (BASELINE.json north_star): per-pixel and per-image values within 1e-5 relative in
fp32.  Per-pixel confidences live in [0, 1] and are differences of O(1) quantities, so the
reference's own two softmax flavours (TF-CPU "x 1/S" vs TF-GPU "/ S") already disagree by
~2e-7 absolute; the per-pixel check therefore carries an absolute floor of
ATOL_PIX = 1e-6 on top of RTOL = 1e-5.  Per-image scores are checked with the pure relative
tolerance.  Labels and ids are bit-exact (ids modulo score gaps below ID_TIE_TOL).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5
ATOL_PIX = 1e-6
ID_TIE_TOL = 2e-7
MEASURES = ("entropy", "margin", "confidence")


@pytest.fixture(scope="module")
def torch():
    import torch as t
    if not t.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return t


@pytest.fixture(scope="module")
def scorer(torch):
    from semanticsegmentationactivelearning_b200 import Scorer
    with Scorer(0) as sc:
        yield sc


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def assert_pix(got, want, what=""):
    got, want = _np(got), np.asarray(want)
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan), what + ": NaN pattern differs"
    err = np.abs(got[~nan] - want[~nan])
    tol = RTOL * np.abs(want[~nan]) + ATOL_PIX
    assert np.all(err <= tol), "%s: max err %.3e (tol %.1e rel + %.1e abs) at %d of %d pixels" % (
        what, err.max(), RTOL, ATOL_PIX, int((err > tol).sum()), err.size)


def assert_scores(got, want, what=""):
    got, want = _np(got).astype(np.float64), np.asarray(want, np.float64)
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan), what + ": NaN pattern differs"
    np.testing.assert_allclose(got[~nan], want[~nan], rtol=RTOL, atol=0, err_msg=what)


def assert_ids(got_ids, want_scores32, unlabelled, k, what=""):
    """Same id set as the oracle's selection on the same f32 scores, excusing swaps between
    examples whose scores differ by less than ID_TIE_TOL (north_star)."""
    from oracle import reference_np as R
    want_ids, _ = R.select_lowest_total_order(want_scores32, unlabelled, k)
    g, w = set(np.asarray(got_ids).tolist()), set(want_ids.tolist())
    assert len(g) == len(w) == len(got_ids), what
    if g == w:
        return
    s = {int(i): float(want_scores32[i]) for i in unlabelled}
    kth = max(s[i] for i in w)
    for i in g ^ w:
        assert abs(s[i] - kth) <= ID_TIE_TOL, "%s: id %d (score %.9g) differs from the oracle beyond the tie tolerance (k-th %.9g)" % (what, i, s[i], kth)


# --------------------------------------------------------------------------------------------
def test_library_is_the_cuda_one():
    import semanticsegmentationactivelearning_b200 as A
    from semanticsegmentationactivelearning_b200 import _lib
    assert _lib.load().als_version() == 110
    assert A.LIB_PATH.endswith("libalscore.so")


def test_synth_generator_bit_exact(torch, scorer):
    from oracle import synth
    for (T, n0, n, H, W, C, dt) in [(1, 0, 3, 8, 16, 19, "float32"), (4, 5, 2, 6, 10, 6, "float32"),
                                    (1, 2, 2, 4, 8, 66, "bfloat16"), (3, 0, 2, 4, 4, 19, "bfloat16")]:
        dev = scorer.synth_logits(T, n0, n, H, W, C, dtype=dt)
        ref = synth.synth_logits(T, n0, n, H, W, C, dtype=dt)
        if dt == "bfloat16":
            got = dev.view(torch.int16).cpu().numpy().view(np.uint16)
        else:
            got = dev.cpu().numpy()
        assert np.array_equal(got, ref), (T, n0, n, H, W, C, dt)


@pytest.mark.parametrize("measure", MEASURES)
def test_golden_graph_cases(torch, scorer, golden, measure):
    for name in golden["graph_cases"].tolist():
        logits = golden[f"{name}.logits"]
        out = scorer.pseudo_annotation(torch.from_numpy(logits).cuda(), measure, threshold=0.9)
        want = golden[f"{name}.{measure}.conf"]
        assert_pix(out["pseudo_confidence"], want, f"{name}/{measure} conf")
        assert_scores(out["pseudo_mean_confidence"], golden[f"{name}.{measure}.mean"], f"{name}/{measure} mean")
        assert np.array_equal(_np(out["pseudo_label"]), golden[f"{name}.label"]), name
        mask, wmask = _np(out["pseudo_mask"]), golden[f"{name}.{measure}.mask"]
        near = np.abs(want - np.float32(0.9)) <= 2 * ATOL_PIX
        assert np.array_equal(mask[~near], wmask[~near]), name
        # host (staged) path gives the same numbers as the zero-copy path
        out_h = scorer.pseudo_annotation(logits, measure, threshold=0.9)
        assert np.array_equal(out_h["pseudo_confidence"], _np(out["pseudo_confidence"]))
        assert np.array_equal(out_h["pseudo_mean_confidence"], _np(out["pseudo_mean_confidence"]))
        assert np.array_equal(out_h["pseudo_label"], _np(out["pseudo_label"]))


SHAPES = [
    # (T, N, H, W, C)    C=19 Cityscapes, 6 Freiburg, 66 Vistas, 2/4/8/12/33/150: other lane layouts, 23: generic kernel
    (1, 3, 32, 64, 19), (1, 2, 24, 40, 6), (1, 2, 16, 24, 66), (1, 2, 8, 24, 2), (1, 2, 8, 24, 4), (1, 2, 8, 24, 8),
    (1, 2, 8, 24, 12), (1, 2, 8, 24, 33), (1, 1, 8, 24, 150), (1, 2, 9, 13, 23), (1, 3, 7, 9, 19), (1, 5, 3, 3, 19),
    (4, 3, 16, 32, 19), (3, 2, 12, 20, 6), (4, 2, 8, 16, 66), (2, 2, 8, 8, 150), (3, 2, 7, 9, 23), (8, 2, 16, 16, 19),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "T%d_N%d_%dx%d_C%d" % s)
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_scores_and_maps_vs_oracle(torch, scorer, shape, dtype):
    from oracle import reference_np as R, synth
    T, N, H, W, C = shape
    x = synth.synth_logits(T, 3, N, H, W, C, dtype=dtype, squeeze_t=False)
    xf = synth.bf16_bits_to_f32(x) if dtype == "bfloat16" else x
    xt = torch.from_numpy(x.view(np.int16) if dtype == "bfloat16" else x).cuda()
    if dtype == "bfloat16":
        xt = xt.view(torch.bfloat16)
    if T == 1:
        xt, xf = xt[0], xf[0]
    for measure in MEASURES + (("variance",) if T > 1 else ()):
        out = scorer.pseudo_annotation(xt, measure, threshold=0.5)
        want = R.pixel_confidence(xf, measure)
        assert_pix(out["pseudo_confidence"], want, f"{shape}/{dtype}/{measure}")
        assert_scores(out["pseudo_mean_confidence"], R.image_scores(want), f"{shape}/{dtype}/{measure}")
        assert np.array_equal(_np(out["pseudo_label"]), R.pseudo_label(xf)), f"{shape}/{dtype}/{measure} label"
        near = np.abs(want - np.float32(0.5)) <= 2 * ATOL_PIX
        assert np.array_equal(_np(out["pseudo_mask"])[~near], R.pseudo_mask(want, 0.5)[~near])
        # ranked path (no per-pixel outputs) must give bit-identical scores to the mapped path
        s2 = scorer.score(xt, measure)
        assert np.array_equal(_np(s2), _np(out["pseudo_mean_confidence"]))


def test_variance_needs_samples(torch, scorer):
    x = torch.zeros(2, 4, 4, 19, device="cuda")
    with pytest.raises(ValueError):
        scorer.score(x, "variance")
    with pytest.raises(NotImplementedError) as ei:
        scorer.score(x, "bald")
    assert str(ei.value) == "Uncertainty function not implemented."


def test_input_validation(torch, scorer):
    with pytest.raises(ValueError):
        scorer.score(torch.zeros(2, 4, 4, 19, device="cuda", dtype=torch.float16), "entropy")
    with pytest.raises(ValueError):
        scorer.score(torch.zeros(4, 4, 19, device="cuda"), "entropy")
    with pytest.raises(ValueError):
        scorer.score(torch.zeros(2, 4, 4, 38, device="cuda")[..., ::2], "entropy")     # strided
    with pytest.raises(ValueError):
        scorer.score(torch.zeros(2, 4, 4, 1, device="cuda"), "entropy")                 # one class
    assert scorer.score(torch.zeros(0, 4, 4, 19, device="cuda"), "entropy").numel() == 0


def test_misaligned_and_odd_shapes(torch, scorer):
    """No alignment requirement on the logits pointer (the reference accepts any shape): a device view that starts
    4 bytes into an allocation, and batches of an odd-sized pool sliced by rank_confidence(batch_size=...), where
    H*W*C*elemsize is not a multiple of 16, score like the aligned copy (the generic kernel takes over)."""
    from oracle import reference_np as R, synth
    from semanticsegmentationactivelearning_b200 import rank_confidence
    x = synth.synth_logits(1, 0, 5, 7, 9, 19)
    flat = torch.zeros(x.size + 1, device="cuda")
    flat[1:] = torch.from_numpy(x).cuda().reshape(-1)
    mis = flat[1:].view(5, 7, 9, 19)
    assert mis.data_ptr() % 16 != 0
    for measure in MEASURES:
        want = R.score_pool(x, measure)
        assert_scores(scorer.score(mis, measure), want, measure)
        assert_scores(scorer.score(x, measure), want, measure)                     # host array (staged)
    xt = synth.synth_logits(3, 0, 5, 7, 9, 19)
    flat = torch.zeros(xt.size + 1, device="cuda")
    flat[1:] = torch.from_numpy(xt).cuda().reshape(-1)
    assert_scores(scorer.score(flat[1:].view(3, 5, 7, 9, 19), "variance"), R.score_pool(xt, "variance"))
    # odd H*W, batch_size < N: every slice after the first starts at an address that is not a multiple of 16
    unl = np.arange(5)
    for src in (torch.from_numpy(x).cuda(), x, torch.from_numpy(x).cuda().to(torch.bfloat16)):
        ids, u = rank_confidence(src, unl, 2, "entropy", batch_size=1, scorer=scorer)
        ref = x if not hasattr(src, "dtype") or src.dtype != torch.bfloat16 else src.float().cpu().numpy()
        want = R.score_pool(ref, "entropy").astype(np.float32)
        np.testing.assert_allclose(u, want, rtol=RTOL)
        assert sorted(ids.tolist()) == sorted(np.argsort(want, kind="stable")[:2].tolist())


def test_large_pageable_host_logits(torch, scorer):
    """Plain NumPy arrays (pageable memory -- what the reference's sess.run hands out, :697-698) above 8 MB go through the
    parallel bounce-buffer staging (csrc/stage.cu): same bits as the device path, for a whole pool in one call, for a batch
    that is a slice (T > 1: one copy per sample plane) and with the source buffer reused right after the call returns."""
    x = scorer.synth_logits(1, 0, 6, 384, 512, 19)                 # 6 x 14.9 MB
    want = _np(scorer.score(x, "entropy"))
    host = x.cpu().numpy().copy()
    assert np.array_equal(scorer.score(host, "entropy"), want)
    scorer.pool_begin(6)
    buf = np.empty_like(host[:2])
    for i in (0, 2, 4):                                            # reuse ONE host buffer: "returns once staged"
        buf[...] = host[i:i + 2]
        scorer.pool_score_batch(buf, np.arange(i, i + 2), "entropy")
        buf[...] = 0
    assert np.array_equal(scorer.pool_scores(6), want.astype(np.float32))
    xt = scorer.synth_logits(3, 0, 4, 256, 512, 19)                # T = 3
    want_t = _np(scorer.score(xt, "variance"))
    host_t = xt.cpu().numpy().copy()
    assert np.array_equal(scorer.score(host_t, "variance"), want_t)
    out = scorer.pseudo_annotation(host, "margin", 0.7)
    dev = scorer.pseudo_annotation(x, "margin", 0.7)
    for key in ("pseudo_confidence", "pseudo_label", "pseudo_mask"):
        assert np.array_equal(out[key], _np(dev[key])), key


def test_non_default_stream(torch, scorer):
    """Scoring under `with torch.cuda.stream(side)` is ordered after the producer on that stream, pool_* entries
    follow torch's current stream too, and alternating streams on one context never share the accumulators in flight."""
    from oracle import reference_np as R, synth
    x = synth.synth_logits(1, 0, 6, 64, 96, 19)
    want = R.score_pool(x, "entropy")
    side = torch.cuda.Stream()
    big = torch.empty(64 << 20, device="cuda")
    for rep in range(3):
        with torch.cuda.stream(side):
            big.normal_()                                         # keep `side` busy so a missing dependency would show
            xt = torch.from_numpy(x).pin_memory().cuda(non_blocking=True)   # produced on `side`
            xt = xt + 0.0
            got = scorer.score(xt, "entropy")
            scorer.pool_begin(6)
            scorer.pool_score_batch(xt, np.arange(6), "entropy")
            ids, u = scorer.pool_select(np.arange(6), 3)
        side.synchronize()
        assert_scores(got, want)
        np.testing.assert_allclose(u, want.astype(np.float32), rtol=RTOL)
        # straight back on the default stream: must wait for the side-stream launch sequence (shared accumulators)
        y = scorer.synth_logits(1, 7, 6, 64, 96, 19)
        a = scorer.score(y, "margin")
        with torch.cuda.stream(side):
            side.wait_stream(torch.cuda.current_stream())
            b = scorer.score(y, "margin")
        torch.cuda.synchronize()
        assert np.array_equal(_np(a), _np(b))
        assert_scores(a, R.score_pool(y.cpu().numpy(), "margin"))


def test_t16_c66_small(torch, scorer):
    """BASELINE config 5's class / sample count (C=66, T=16) on small images, f32 and bf16, all four measures."""
    from oracle import reference_np as R, synth
    for dtype in ("float32", "bfloat16"):
        x = synth.synth_logits(16, 2, 2, 12, 16, 66, dtype=dtype)
        xf = synth.bf16_bits_to_f32(x) if dtype == "bfloat16" else x
        xt = torch.from_numpy(x.view(np.int16) if dtype == "bfloat16" else x).cuda()
        if dtype == "bfloat16":
            xt = xt.view(torch.bfloat16)
        for measure in MEASURES + ("variance",):
            out = scorer.pseudo_annotation(xt, measure, threshold=0.5)
            want = R.pixel_confidence(xf, measure)
            assert_pix(out["pseudo_confidence"], want, f"T16 C66 {dtype} {measure}")
            assert_scores(out["pseudo_mean_confidence"], R.image_scores(want), f"T16 C66 {dtype} {measure}")


def test_outputs_stay_inside_their_buffers(torch, scorer):
    """compute-sanitizer is closed on this pool (profiles/r02_sanitize_memcheck_refusal.txt), so the write side of memory
    safety is checked by hand: every output lives in the middle of a sentinel-filled allocation and the guard bands must
    be untouched after the kernels ran -- odd shapes (ragged last tile, images straddling tiles), every lane layout, the
    generic kernel, T > 1, bf16, the streamed path and the selection block."""
    from oracle import synth
    G = 4096

    def guarded(n, dtype, fill):
        buf = torch.full((n + 2 * G,), fill, dtype=dtype, device="cuda")
        return buf, buf[G:G + n]

    def intact(buf, n, fill):
        return bool((buf[:G] == fill).all() and (buf[G + n:] == fill).all())

    for (T, N, H, W, C, dt) in [(1, 5, 33, 47, 19, "float32"), (1, 3, 17, 23, 6, "bfloat16"), (3, 2, 9, 13, 66, "float32"),
                                (2, 3, 7, 9, 23, "float32"), (1, 2, 5, 5, 150, "bfloat16"), (4, 3, 31, 33, 19, "bfloat16")]:
        x = synth.synth_logits(T, 0, N, H, W, C, dtype=dt)
        xt = torch.from_numpy(x.view(np.int16) if dt == "bfloat16" else x).cuda()
        if dt == "bfloat16":
            xt = xt.view(torch.bfloat16)
        P = N * H * W
        cb, conf = guarded(P, torch.float32, -7.0)
        lb, label = guarded(P, torch.uint8, 0xAB)
        mb, mask = guarded(P, torch.uint8, 0xCD)
        sb, scores = guarded(N, torch.float64, -3.0)
        out = {"pseudo_confidence": conf.view(N, H, W), "pseudo_mean_confidence": scores, "pseudo_label": label.view(N, H, W),
               "pseudo_mask": mask.view(N, H, W)}
        for measure in MEASURES + (("variance",) if T > 1 else ()):
            scorer.pseudo_annotation(xt, measure, 0.5, out=out)
            torch.cuda.synchronize()
            assert intact(cb, P, -7.0) and intact(lb, P, 0xAB) and intact(mb, P, 0xCD) and intact(sb, N, -3.0), (T, N, H, W, C, dt, measure)
            assert bool((conf > -1.0).all()) and bool((scores > -1.0).all())          # ... and every element was written
        if T > 1:
            scorer.mc_begin((N, H, W, C), dt, label=label.view(N, H, W))
            for t in range(T):
                scorer.mc_add_sample(xt[t].contiguous())
            scorer.mc_finish("variance")
            torch.cuda.synchronize()
            assert intact(lb, P, 0xAB)
    kb, keys = guarded(777, torch.float32, -7.0)
    ib, ids = guarded(777, torch.int64, -5)
    keys.copy_(torch.rand(777, device="cuda"))
    ids.copy_(torch.randperm(777, device="cuda"))
    ok, oi = scorer.select_smallest(keys, ids, 100)
    torch.cuda.synchronize()
    assert intact(kb, 777, -7.0) and intact(ib, 777, -5) and len(oi) == 100


def test_pool_select_rejects_duplicates_and_bad_ids(torch, scorer):
    scorer.pool_begin(8)
    with pytest.raises(ValueError):
        scorer.pool_select(np.array([1, 2, 2, 3]), 2)
    with pytest.raises(ValueError):
        scorer.pool_select(np.array([1, 8]), 1)
    ids, u = scorer.pool_select(np.array([5, 1, 7]), 5)          # k >= M returns all; unvisited examples are 0.0
    assert ids.tolist() == [1, 5, 7] and np.all(u == 0)
    # device primitive with duplicated pairs: no out-of-bounds write, the distinct smallest pairs are still found
    keys = torch.tensor([0.5, 0.25, 0.25, 0.75], device="cuda")
    idt = torch.tensor([4, 9, 9, 1], device="cuda")
    ok, oi = scorer.select_smallest(keys, idt, 2)
    assert len(oi) == 2


def test_nan_and_inf_logits(torch, scorer):
    from oracle import reference_np as R, synth
    x = synth.synth_logits(1, 0, 4, 8, 8, 19)
    x[1, 2, 3, 4] = np.nan
    x[2, 0, 0, :] = -np.inf
    x[2, 0, 0, 5] = 1.0          # all but one class masked: p = one-hot
    x[3, 1, 1, 7] = np.inf
    for measure in MEASURES:
        got = _np(scorer.score(torch.from_numpy(x).cuda(), measure))
        want = R.score_pool(x, measure)
        assert np.isnan(got[1]) and np.isnan(want[1]) and np.isnan(got[3]) and np.isnan(want[3])
        assert_scores(got[[0, 2]], want[[0, 2]], measure)


def test_run_to_run_determinism(torch, scorer):
    x = scorer.synth_logits(1, 0, 6, 64, 96, 19)
    a = _np(scorer.score(x, "entropy"))
    for _ in range(5):
        assert np.array_equal(_np(scorer.score(x, "entropy")), a)


def test_properties_shift_and_permutation(torch, scorer):
    x = scorer.synth_logits(1, 0, 3, 32, 32, 19)
    perm = torch.randperm(19, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    for measure in MEASURES:
        base = _np(scorer.score(x, measure))
        np.testing.assert_allclose(_np(scorer.score((x + 8.0).contiguous(), measure)), base, rtol=RTOL)
        np.testing.assert_allclose(_np(scorer.score(x[..., perm].contiguous(), measure)), base, rtol=RTOL)
    const = torch.full((2, 16, 16, 19), 1.5, device="cuda")
    np.testing.assert_allclose(_np(scorer.score(const, "confidence")), 1.0 / 19, rtol=RTOL)
    assert np.all(np.abs(_np(scorer.score(const, "entropy"))) <= 3e-7)
    assert np.all(_np(scorer.score(const, "margin")) == 0)


def test_golden_rank_cases(torch, scorer, golden):
    """als_pool_* against the reference closure's outputs (executed source, make_golden.py)."""
    for name in golden["rank_cases"].tolist():
        scores64 = golden[f"{name}.scores64"]
        unl, k = golden[f"{name}.unlabelled"], int(golden[f"{name}.k"])
        ids_ref, uconf_ref = golden[f"{name}.ids"], golden[f"{name}.uconf"]
        n = int(golden[f"{name}.num_examples"])
        # feed the pool vector exactly as :700 would (f64 -> f32), via the device select primitive
        conf32 = np.zeros(n, np.float32)
        order, bs, trunc = golden[f"{name}.order"], int(golden[f"{name}.batch_size"]), int(golden[f"{name}.truncate_after"])
        visited = order if trunc < 0 else order[:trunc * bs]
        conf32[visited] = scores64[visited]
        keys = torch.from_numpy(conf32[unl]).cuda()
        ok, oi = scorer.select_smallest(keys, torch.from_numpy(unl).cuda(), k)
        assert np.array_equal(conf32[unl], uconf_ref, equal_nan=True)
        got = _np(oi)
        assert len(got) == len(ids_ref)
        # identical set unless np.argpartition broke an exact tie differently
        u = np.nan_to_num(conf32[unl].astype(np.float64), nan=np.inf)
        kth = np.sort(u)[k - 1]
        strictly = set(unl[u < kth].tolist())
        ties = set(unl[u == kth].tolist())
        assert strictly <= set(got.tolist()) and strictly <= set(ids_ref.tolist()), name
        assert set(got.tolist()) - strictly <= ties, name
        # our completion: ascending (score, id)
        gk = np.nan_to_num(_np(ok).astype(np.float64), nan=np.inf)
        assert np.all(np.diff(gk) >= 0), name
        same = np.diff(gk) == 0
        assert np.all(np.diff(got)[same] > 0), name


def test_select_edge_cases(torch, scorer):
    dev = "cuda"
    keys = torch.tensor([0.5, -0.0, 0.0, float("nan"), float("inf"), -1.0, 0.5, float("-inf")], device=dev)
    ids = torch.tensor([7, 6, 5, 4, 3, 2, 1, 0], device=dev)
    ok, oi = scorer.select_smallest(keys, ids, 8)
    assert _np(oi).tolist() == [0, 2, 5, 6, 1, 7, 3, 4]          # -inf, -1, (+-0 by id), 0.5 (by id), inf, NaN
    ok, oi = scorer.select_smallest(keys, ids, 3)
    assert _np(oi).tolist() == [0, 2, 5]
    ok, oi = scorer.select_smallest(keys, ids, 100)                # k >= M returns all (reference raises)
    assert len(oi) == 8
    ok, oi = scorer.select_smallest(keys, ids, 0)
    assert len(oi) == 0
    ok, oi = scorer.select_smallest(keys[:0], ids[:0], 5)
    assert len(oi) == 0
    # large, duplicated keys, 64-bit ids
    g = torch.Generator(device=dev).manual_seed(3)
    M = 50000
    keys = torch.randint(0, 64, (M,), device=dev, generator=g).float() / 64
    ids = torch.randperm(M, device=dev, generator=g) + (1 << 40)
    for k in (1, 50, 999, 20000):
        ok, oi = scorer.select_smallest(keys, ids, k)
        order = np.lexsort((_np(ids), _np(keys)))
        assert np.array_equal(_np(oi), _np(ids)[order[:k]])
        assert np.array_equal(_np(ok), _np(keys)[order[:k]])


def test_select_random_against_total_order(torch, scorer):
    """The radix select stops early when a digit bin holds exactly the elements still wanted: exercise boundaries that
    fall inside, at the end of and between groups of equal keys, with NaN / inf / signed zeros and negative ids."""
    from oracle import reference_np as R
    rng = np.random.default_rng(20191013)
    for trial in range(60):
        M = int(rng.choice([1, 2, 3, 17, 255, 256, 257, 1000, 2975, 18000]))
        distinct = int(rng.choice([1, 2, 5, 50, 10 ** 6]))
        vals = rng.standard_normal(distinct).astype(np.float32)
        if trial % 3 == 0:
            vals[: min(4, distinct)] = np.array([np.nan, np.inf, -0.0, 0.0], np.float32)[: min(4, distinct)]
        keys = vals[rng.integers(0, distinct, M)]
        ids = rng.permutation(4 * M)[:M].astype(np.int64) - (M if trial % 2 else 0)     # unique, some negative
        for k in {0, 1, M // 2, max(M - 1, 0), M, M + 3, int(rng.integers(0, M + 1))}:
            ok, oi = scorer.select_smallest(torch.from_numpy(keys).cuda(), torch.from_numpy(ids).cuda(), k)
            # oracle: total order over (key with -0 == +0 and NaN last, id); confidence indexed by position
            order = np.lexsort((ids, np.isnan(keys), np.where(np.isnan(keys), np.inf, keys + np.float32(0.0))))
            want = order[: min(k, M)]
            assert np.array_equal(_np(oi), ids[want]), (trial, M, distinct, k)
            assert np.array_equal(_np(ok), keys[want], equal_nan=True), (trial, M, distinct, k)


@pytest.mark.parametrize("measure,T", [("entropy", 1), ("margin", 1), ("confidence", 1), ("variance", 4)])
def test_rank_confidence_end_to_end(torch, scorer, measure, T):
    """Whole closure: device and host (batched like sess.run, shuffled) against the oracle."""
    from oracle import reference_np as R, synth
    from semanticsegmentationactivelearning_b200 import rank_confidence
    N, H, W, C, k = 24, 16, 32, 19, 7
    x = synth.synth_logits(T, 0, N, H, W, C)
    rng = np.random.default_rng(5)
    unl = np.sort(rng.choice(N, 18, replace=False))
    conf = R.scatter_scores(N, [(R.score_pool(x, measure), np.arange(N))])
    want_ids, want_u = R.select_lowest(conf, unl, k)
    # (1) whole pool resident on the device
    ids, u = rank_confidence(torch.from_numpy(x).cuda(), unl, k, measure, scorer=scorer)
    np.testing.assert_allclose(u, want_u, rtol=RTOL)
    assert_ids(ids, conf, unl, k, measure)
    assert sorted(ids.tolist()) == sorted(want_ids.tolist())
    # (2) host batches of 8 in shuffled order with example indices (:697-700)
    perm = rng.permutation(N)
    xs = x[perm] if T == 1 else x[:, perm]
    ids2, u2 = rank_confidence(xs, unl, k, measure, batch_size=8, example_index=perm, num_examples=N, scorer=scorer)
    assert np.array_equal(u2, u) and np.array_equal(ids2, ids)
    # (3) short pass: unvisited examples keep 0.0 and are selected first (:685, :701-702)
    batches = [(xs[i:i + 8] if T == 1 else np.ascontiguousarray(xs[:, i:i + 8]), perm[i:i + 8]) for i in (0, 8)]
    ids3, u3 = rank_confidence(batches, unl, k, measure, num_examples=N, scorer=scorer)
    unseen = [i for i in unl if i not in set(perm[:16].tolist())]
    assert set(unseen[:k]) <= set(ids3.tolist()) or len(unseen) >= k
    assert np.all(u3[np.isin(unl, unseen)] == 0)


def test_rank_pool_one_call_equals_three_calls(torch, scorer):
    """als_rank_pool (:682-715 in one call) == pool_begin + pool_score_batch + pool_select, for device and host logits,
    shuffled example indices, a pool larger than the batch (unvisited examples stay 0.0), k = 0, an empty and a large
    unlabelled set (late upload path), and the library's timing hook around the scoring launch."""
    from oracle import synth
    N, H, W, C = 12, 16, 24, 19
    x = synth.synth_logits(1, 0, N, H, W, C)
    xt = torch.from_numpy(x).cuda()
    rng = np.random.default_rng(9)
    for src in (xt, x):
        for idx, num in ((None, N), (rng.permutation(40)[:N], 40)):
            unl = np.sort(rng.choice(num, num - 3, replace=False))
            for k in (5, 0, 100):
                ids, u = scorer.rank_pool(src, unl, k, "margin", example_index=idx, num_examples=num)
                scorer.pool_begin(num)
                scorer.pool_score_batch(src, np.arange(N) if idx is None else idx, "margin")
                ids2, u2 = scorer.pool_select(unl, k)
                assert np.array_equal(ids, ids2) and np.array_equal(u, u2), (type(src).__name__, idx is None, k)
    ids, u = scorer.rank_pool(xt, np.zeros(0, np.int64), 3, "entropy")
    assert len(ids) == 0 and len(u) == 0
    big = np.arange(6000)                                            # > 4096 ids: uploaded while the GPU scores
    ids, u = scorer.rank_pool(xt, big, 7, "entropy", num_examples=6000)
    assert ids.tolist() == list(range(N, N + 7)) and np.all(u[N:] == 0) and np.all(u[:N] != 0)
    scorer.enable_timing(True)
    scorer.rank_pool(xt, np.arange(N), 3, "entropy")
    assert 0.0 < scorer.last_scoring_ms() < 50.0
    scorer.enable_timing(False)
    with pytest.raises(RuntimeError):
        scorer.last_scoring_ms()


def test_dlpack_zero_copy_and_host(torch, scorer):
    from oracle import reference_np as R, synth
    x = synth.synth_logits(1, 0, 3, 8, 16, 19)
    want = R.score_pool(x, "entropy")
    xt = torch.from_numpy(x).cuda()
    assert_scores(scorer.score_dlpack(xt, "entropy"), want)               # kDLCUDA
    assert_scores(scorer.score_dlpack(x, "entropy"), want)                # kDLCPU (NumPy)
    assert_scores(scorer.score_dlpack(xt.to(torch.bfloat16), "entropy"),
                  R.score_pool(xt.to(torch.bfloat16).float().cpu().numpy(), "entropy"))
    with pytest.raises(ValueError):
        scorer.score_dlpack(xt.double(), "entropy")
    with pytest.raises(ValueError):
        scorer.score_dlpack(xt.permute(0, 3, 1, 2), "entropy")            # not dense NHWC
    x5 = synth.synth_logits(3, 0, 2, 8, 8, 6)
    assert_scores(scorer.score_dlpack(torch.from_numpy(x5).cuda(), "variance"), R.score_pool(x5, "variance"))


def test_cfg1_shape_full_resolution_subset(torch, scorer):
    """BASELINE config 1 geometry (512x1024, C=19, entropy) on 2 images, per-pixel + scores."""
    from oracle import reference_np as R, synth
    x = synth.synth_logits(1, 0, 2, 512, 1024, 19)
    out = scorer.pseudo_annotation(torch.from_numpy(x).cuda(), "entropy")
    want = R.pixel_confidence(x, "entropy")
    assert_pix(out["pseudo_confidence"], want)
    assert_scores(out["pseudo_mean_confidence"], R.image_scores(want))


def test_full_size_properties(torch, scorer):
    """Size-independent checks at full pool geometry (64 x 512x1024 x 19, BASELINE config 1):
    duplicated images score identically, a constant image scores its analytic value, and the
    selection equals the oracle's selection on the GPU's own score vector."""
    from oracle import reference_np as R
    N = 64
    x = scorer.synth_logits(1, 0, N, 512, 1024, 19)
    x[7] = x[3]
    x[11] = 2.0
    s = _np(scorer.score(x, "entropy"))
    assert s[7] == s[3]
    assert abs(s[11]) <= 3e-7
    assert np.all(np.isfinite(s)) and len(set(np.delete(s, [7]).tolist())) == N - 1
    s32 = s.astype(np.float32)
    unl = np.arange(N)
    from semanticsegmentationactivelearning_b200 import rank_confidence
    ids, u = rank_confidence(x, unl, 50, "entropy", scorer=scorer)
    assert np.array_equal(u, s32)
    want, _ = R.select_lowest(s32, unl, 50)
    assert sorted(ids.tolist()) == sorted(want.tolist())
