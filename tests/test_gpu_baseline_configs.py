"""GPU parity at the BASELINE.json geometries (SURVEY.md section 8(d) "parity coverage"):

* cfg1 in full -- all 64 images @512x1024, C=19, entropy: per-pixel maps, labels, per-image scores and the
  selected ids against oracle/reference_np.py (restating /root/reference/active_learning.py:239-263, :705-714);
* cfg2 (T=8 variance @512x1024, C=19), cfg3 (margin @1024x2048: P = 2^21, the smallest fixed-point shift),
  cfg4 (C=6 entropy @480x640), cfg5 (C=66, T=16 variance @512x1024) on a deterministic subset, f32 and bf16 logits;
* the selection step on a full 2975-image pass: the GPU's score vector goes through the oracle's
  np.argpartition (:705-714) and the sorted id sets must agree.

Inputs come from the device generator, which test_gpu_parity.py::test_synth_generator_bit_exact pins bit for bit to
oracle/synth.py; the oracle runs on the host copy of the same bytes.  Tolerances as in test_gpu_parity.py for f32
logits.  bf16 logits: the kernels compute in fp32 on the bf16 values, so they are held to the same fp32 tolerance
against the oracle evaluated on the bf16-rounded inputs (north_star allows 1e-2)."""
import numpy as np
import pytest

from _oracle_par import pixel_confidence as oracle_conf

pytestmark = pytest.mark.gpu

RTOL = 1e-5
ATOL_PIX = 1e-6
ID_TIE_TOL = 2e-7
K = 50  # conf/enet_cityscapes_active_learning.json:59


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


def _host_f32(torch, xt):
    """Device logits -> the float32 values the kernels see (bf16 widened exactly)."""
    return xt.float().cpu().numpy()


def _check_maps(torch, scorer, xt, measure, what):
    """Per-pixel map, label, per-image scores of one device batch against the oracle.  Returns the oracle's scores."""
    from oracle import reference_np as R
    xf = _host_f32(torch, xt)
    out = scorer.pseudo_annotation(xt, measure, threshold=0.9)
    want = oracle_conf(xf, measure)
    got = out["pseudo_confidence"].cpu().numpy()
    err = np.abs(got - want)
    tol = RTOL * np.abs(want) + ATOL_PIX
    assert np.all(err <= tol), "%s: max per-pixel err %.3e at %d pixels" % (what, err.max(), int((err > tol).sum()))
    want_scores = R.image_scores(want)
    np.testing.assert_allclose(out["pseudo_mean_confidence"].cpu().numpy(), want_scores, rtol=RTOL, atol=0, err_msg=what)
    assert np.array_equal(out["pseudo_label"].cpu().numpy(), R.pseudo_label(xf)), what + ": labels"
    near = np.abs(want - np.float32(0.9)) <= 2 * ATOL_PIX
    assert np.array_equal(out["pseudo_mask"].cpu().numpy()[~near], R.pseudo_mask(want, 0.9)[~near]), what + ": mask"
    # the ranked path (no per-pixel outputs) gives bit-identical scores
    assert np.array_equal(scorer.score(xt, measure).cpu().numpy(), out["pseudo_mean_confidence"].cpu().numpy()), what
    return want_scores


def _assert_same_ids(got_ids, want_ids, scores32, what):
    g, w = set(np.asarray(got_ids).tolist()), set(np.asarray(want_ids).tolist())
    assert len(g) == len(w) == len(got_ids), what
    if g == w:
        return
    kth = max(float(scores32[i]) for i in w)
    for i in g ^ w:   # north_star: swaps excused only where the score gap is below tolerance
        assert abs(float(scores32[i]) - kth) <= ID_TIE_TOL, "%s: id %d differs from the oracle beyond the tie tolerance" % (what, i)


def test_cfg1_full_pool_maps_scores_ids(torch, scorer):
    """BASELINE config 1 in full: 64 images @512x1024, C=19, entropy, k=50."""
    from oracle import reference_np as R
    from semanticsegmentationactivelearning_b200 import rank_confidence
    N, H, W, C = 64, 512, 1024, 19
    x = scorer.synth_logits(1, 0, N, H, W, C)
    want_scores = np.empty(N, np.float64)
    for n0 in range(0, N, 8):                              # batches of 8 like params["batch_size"] (:689)
        want_scores[n0:n0 + 8] = _check_maps(torch, scorer, x[n0:n0 + 8], "entropy", "cfg1 images %d..%d" % (n0, n0 + 7))
    unl = np.arange(N)
    ids, u = rank_confidence(x, unl, K, "entropy", batch_size=8, scorer=scorer)
    conf32 = R.scatter_scores(N, [(want_scores, np.arange(N))])              # :685, :700
    want_ids, want_u = R.select_lowest(conf32, unl, K)                          # :705-714 verbatim
    np.testing.assert_allclose(u, want_u, rtol=RTOL, atol=0)
    _assert_same_ids(ids, want_ids, conf32, "cfg1")


CASES = [
    # name, T, N, H, W, C, measure
    ("cfg2", 8, 2, 512, 1024, 19, "variance"),
    ("cfg3", 1, 2, 1024, 2048, 19, "margin"),
    ("cfg4", 1, 4, 480, 640, 6, "entropy"),
    ("cfg5", 16, 2, 512, 1024, 66, "variance"),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_config_geometry_subset(torch, scorer, case, dtype):
    """BASELINE configs 2-5 at their stated geometry / class count / sample count on a deterministic subset."""
    name, T, N, H, W, C, measure = case
    for n in range(N):                                       # one image at a time bounds the oracle's host memory
        x = scorer.synth_logits(T, 1000 + n, 1, H, W, C, dtype=dtype, squeeze_t=True)
        _check_maps(torch, scorer, x, measure, "%s/%s image %d" % (name, dtype, n))
        del x
    torch.cuda.empty_cache()


def test_cfg2_other_measures_on_mc_mean(torch, scorer):
    """T=8 @512x1024, C=19: entropy / margin / max-prob of the predictive mean (one image)."""
    x = scorer.synth_logits(8, 7, 1, 512, 1024, 19)
    for measure in ("entropy", "margin", "confidence"):
        _check_maps(torch, scorer, x, measure, "cfg2 mean/" + measure)


def test_full_pool_selection_2975(torch, scorer):
    """One full pool pass (2975 images @512x1024, C=19, entropy, chunks of 175 like bench.py): the GPU's score vector
    fed to the oracle's selection (:705-714) gives the same sorted id set; scores are finite and (nearly all) distinct."""
    from oracle import reference_np as R
    N, H, W, C, chunk = 2975, 512, 1024, 19, 175
    rng = np.random.default_rng(20191013)
    unl = np.sort(rng.choice(N, N - 270, replace=False))        # 270 labelled images, as in the shipped AL config
    buf = None
    scorer.pool_begin(N)
    for n0 in range(0, N, chunk):
        nb = min(chunk, N - n0)
        if buf is None or buf.shape[1] != nb:
            buf = None
            torch.cuda.empty_cache()
        buf = scorer.synth_logits(1, n0, nb, H, W, C, squeeze_t=False, out=buf)
        scorer.pool_score_batch(buf[0], np.arange(n0, n0 + nb), "entropy")
        torch.cuda.synchronize()
    ids, u = scorer.pool_select(unl, K)
    scores32 = scorer.pool_scores(N)
    assert np.all(np.isfinite(scores32)) and len(np.unique(scores32)) >= N - 16      # float32 scores: a few coincide
    assert np.array_equal(u, scores32[unl])
    want_ids, want_u = R.select_lowest(scores32, unl, K)
    assert np.array_equal(want_u, u)
    _assert_same_ids(ids, want_ids, scores32, "2975-image pass")
    # and the deterministic completion: ascending (score, id)
    tot_ids, _ = R.select_lowest_total_order(scores32, unl, K)
    assert np.array_equal(ids, tot_ids)
    # spot-check three images of the pass against the oracle's per-image score (first, middle, last chunk)
    for n in (0, 1500, N - 1):
        x = scorer.synth_logits(1, n, 1, H, W, C)
        want = R.image_scores(oracle_conf(x.cpu().numpy(), "entropy"))[0]
        assert abs(float(scores32[n]) - want) <= RTOL * abs(want) + 6e-8 * abs(want)
