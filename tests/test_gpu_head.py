"""GPU parity of the fused classifier-head path (csrc/head.cu, tcgen05 split-TF32) against the CPU oracle:
oracle/reference_np.final_head (Final.call, models/enet/enet_modules.py:1359-1381) followed by the scoring
graph (active_learning.py:234-269), and against the golden fixture made by executing Final.call itself.

Tolerances.  Per-image scores: 1e-5 relative against the fp32 oracle (north star).  Per pixel the logits are
themselves sums of ~36 fp32 products of magnitude |x| -- any fp32 evaluation (TF's included) carries a few
ulp(|x|) of rounding noise, which the softmax passes on as an ABSOLUTE confidence error of up to ~4x that.  The
per-pixel bar is therefore  |gpu - truth64| <= 1e-5 |truth64| + 4 eps32 (1 + max_c |logit|)  against a float64
evaluation, and the GPU must be no worse than ~4x the fp32 oracle's own worst error against the same truth: the
split-TF32 tensor-core contraction has to be as accurate as an fp32 one."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5
EPS32 = float(np.finfo(np.float32).eps)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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


def _case(C, N, h, w, seed):
    rng = np.random.default_rng(seed)
    feat = rng.standard_normal((N, h, w, 16)).astype(np.float32)
    feat *= rng.uniform(0.3, 1.3, size=(N, 1, 1, 1)).astype(np.float32)     # images differ in confidence
    kern = (0.4 * rng.standard_normal((3, 3, C, 16))).astype(np.float32)
    return feat, kern


def _check_maps(got, feat, kern, measure, threshold=0.6, logits32=None):
    from oracle import reference_np as R
    logits = R.final_head(feat, kern) if logits32 is None else logits32
    logits64 = R.conv2d_transpose_same(feat, kern, dtype=np.float64)
    truth = R.pixel_confidence_f64(logits64, measure)
    want = R.pixel_confidence(logits, measure)                       # the fp32 oracle
    conf = got["pseudo_confidence"].cpu().numpy()
    err_gpu = np.abs(conf - truth)
    err_ref = np.abs(want - truth)
    tol = RTOL * np.abs(truth) + 4 * EPS32 * (1.0 + np.abs(logits64).max(axis=-1))
    assert np.all(err_gpu <= tol), "%s: max err %.3e vs truth at %d of %d pixels (fp32 oracle: %.3e)" % (
        measure, err_gpu.max(), int((err_gpu > tol).sum()), err_gpu.size, err_ref.max())
    assert err_gpu.max() <= 4 * err_ref.max() + 5e-7, "%s: GPU %.3e vs fp32 oracle %.3e" % (measure, err_gpu.max(), err_ref.max())
    np.testing.assert_allclose(got["pseudo_mean_confidence"].cpu().numpy(), R.image_scores(want), rtol=RTOL, atol=0)
    # labels: identical wherever the top-2 logits are separated by more than the logits' own rounding noise
    lab = got["pseudo_label"].cpu().numpy()
    srt = np.sort(logits, axis=-1)
    clear = (srt[..., -1] - srt[..., -2]) > 1e-4
    assert np.array_equal(lab[clear], R.pseudo_label(logits)[clear])
    # mask: consistent with the returned confidence map, and with the oracle away from the threshold
    mask = got["pseudo_mask"].cpu().numpy()
    assert np.array_equal(mask, R.pseudo_mask(conf, threshold))
    far = np.abs(want - np.float32(threshold)) > 1e-4
    assert np.array_equal(mask[far], R.pseudo_mask(want, threshold)[far])
    return float(err_gpu.max()), float(err_ref.max())


@pytest.mark.parametrize("C,shape", [(c, s) for c in (19, 6) for s in [(3, 37, 150), (2, 8, 128), (1, 1, 1), (2, 5, 129), (1, 70, 300)]]
                         + [(c, (2, 9, 140)) for c in (2, 3, 4, 5, 8, 12, 13, 16, 20, 21, 24, 27, 32)],
                         ids=lambda v: str(v) if isinstance(v, int) else "N%d_%dx%d" % v)
def test_fused_head_vs_oracle(torch, scorer, C, shape):
    from oracle import reference_np as R
    N, h, w = shape
    feat, kern = _case(C, N, h, w, seed=100 * C + h + w)
    scorer.prepare_head(kern)
    f = torch.from_numpy(feat).cuda()
    for measure in MEASURES:
        got = scorer.pseudo_annotation_features(f, measure, 0.6)
        eg, er = _check_maps(got, feat, kern, measure)
        print("C=%d %s %s: max |conf - truth64| gpu %.2e, fp32 oracle %.2e" % (C, shape, measure, eg, er))
        s2 = scorer.score_features(f, measure)
        assert torch.equal(s2, got["pseudo_mean_confidence"]), "scores must not depend on the optional outputs"
        s3 = scorer.score_features(f, measure)
        assert torch.equal(s2, s3), "run-to-run determinism"


def _mc_case(C, T, N, h, w, seed):
    """T Monte-Carlo-dropout samples of the `Final` input: a shared feature map with channel-wise dropout noise per
    sample (spatial_dropout, models/util/extra_ops.py:137-151: noise shape [B,1,1,C]) plus a small perturbation."""
    rng = np.random.default_rng(seed)
    base, kern = _case(C, N, h, w, seed)
    keep = (rng.uniform(size=(T, N, 1, 1, 16)) > 0.1).astype(np.float32) / np.float32(0.9)
    feat = base[None] * keep + (0.05 * rng.standard_normal((T, N, h, w, 16))).astype(np.float32)
    return np.ascontiguousarray(feat, np.float32), kern


@pytest.mark.parametrize("C,T,shape", [(19, 8, (3, 21, 150)), (19, 2, (2, 5, 129)), (19, 3, (1, 1, 1)), (6, 4, (2, 9, 140)),
                                       (6, 8, (1, 40, 260)), (2, 2, (2, 9, 140)), (5, 3, (2, 9, 140)), (12, 5, (2, 9, 140)),
                                       (16, 2, (2, 9, 140)), (20, 4, (2, 9, 140)), (21, 3, (2, 9, 140)), (24, 2, (2, 9, 140))],
                         ids=lambda v: str(v) if isinstance(v, int) else "N%d_%dx%d" % v)
def test_fused_head_mc_samples_vs_oracle(torch, scorer, C, T, shape):
    """T > 1: per-sample logits = Final(features[t]); Welford mean / variance over t as in the logits path
    (oracle.pixel_confidence on the stacked per-sample logits)."""
    from oracle import reference_np as R
    N, h, w = shape
    feat, kern = _mc_case(C, T, N, h, w, seed=1000 * T + 10 * C + h)
    scorer.prepare_head(kern)
    f = torch.from_numpy(feat).cuda()
    logits = np.stack([R.final_head(feat[t], kern) for t in range(T)])
    logits64 = np.stack([R.conv2d_transpose_same(feat[t], kern, dtype=np.float64) for t in range(T)])
    for measure in MEASURES + ("variance",):
        got = scorer.pseudo_annotation_features(f, measure, 0.6)
        want = R.pixel_confidence(logits, measure)
        truth = R.pixel_confidence_f64(logits64, measure)
        conf = got["pseudo_confidence"].cpu().numpy()
        err_gpu, err_ref = np.abs(conf - truth), np.abs(want - truth)
        tol = RTOL * np.abs(truth) + 4 * EPS32 * (1.0 + np.abs(logits64).max(axis=(0, -1)))
        assert np.all(err_gpu <= tol), "%s: max err %.3e at %d of %d pixels (fp32 oracle %.3e)" % (
            measure, err_gpu.max(), int((err_gpu > tol).sum()), err_gpu.size, err_ref.max())
        assert err_gpu.max() <= 4 * err_ref.max() + 5e-7, "%s: GPU %.3e vs fp32 oracle %.3e" % (measure, err_gpu.max(), err_ref.max())
        np.testing.assert_allclose(got["pseudo_mean_confidence"].cpu().numpy(), R.image_scores(want), rtol=RTOL, atol=0)
        srt = np.sort(logits[0], axis=-1)
        clear = (srt[..., -1] - srt[..., -2]) > 1e-4
        assert np.array_equal(got["pseudo_label"].cpu().numpy()[clear], R.pseudo_label(logits)[clear])   # sample 0
        assert np.array_equal(got["pseudo_mask"].cpu().numpy(), R.pseudo_mask(conf, 0.6))
        s2 = scorer.score_features(f, measure)
        assert torch.equal(s2, got["pseudo_mean_confidence"]), "scores must not depend on the optional outputs"
        assert torch.equal(s2, scorer.score_features(f, measure)), "run-to-run determinism"
        # the logits path on the materialised per-sample logits gives the same scores
        s_logits = scorer.score(torch.from_numpy(logits).cuda(), measure)
        np.testing.assert_allclose(s2.cpu().numpy(), s_logits.cpu().numpy(), rtol=RTOL)


def test_fused_head_mc_rank_confidence(torch, scorer):
    """rank_confidence() from MC feature samples: device pool, device batches and host batches agree with the logits path."""
    from oracle import reference_np as R
    from semanticsegmentationactivelearning_b200 import rank_confidence
    feat, kern = _mc_case(19, 4, 20, 12, 130, seed=11)
    logits = torch.from_numpy(np.stack([R.final_head(feat[t], kern) for t in range(4)])).cuda()
    unl = np.arange(1, 20)
    ids_a, conf_a = rank_confidence(logits, unl, 5, "variance", scorer=scorer, batch_size=8)
    ids_b, conf_b = rank_confidence(torch.from_numpy(feat).cuda(), unl, 5, "variance", scorer=scorer, batch_size=8, head_kernel=kern)
    ids_c, conf_c = rank_confidence(feat, unl, 5, "variance", scorer=scorer, batch_size=8, head_kernel=kern)   # host batches
    np.testing.assert_allclose(conf_b, conf_a, rtol=RTOL)
    assert np.array_equal(conf_b, conf_c)
    assert sorted(ids_a.tolist()) == sorted(ids_b.tolist()) == sorted(ids_c.tolist())


def test_fused_head_golden_final_call(torch, scorer):
    """Fixtures produced by executing the reference's Final.call (tests/golden/make_golden_head.py)."""
    with np.load(os.path.join(ROOT, "tests", "golden", "final_head.npz"), allow_pickle=False) as z:
        gold = {k: z[k] for k in z.files}
    ran = 0
    for name in gold["cases"]:
        kern = gold[name + ".kernel"]
        if not scorer.head_supported(kern.shape[2]):
            continue
        scorer.prepare_head(kern)
        f = torch.from_numpy(gold[name + ".features"]).cuda()
        for measure in MEASURES:
            _check_maps(scorer.pseudo_annotation_features(f, measure, 0.6), gold[name + ".features"], kern, measure,
                        logits32=gold[name + ".logits"])
        ran += 1
    assert ran >= 2


def test_fused_head_matches_logits_path(torch, scorer):
    """Same ids and scores as scoring the materialised logits with the HBM kernel."""
    from oracle import reference_np as R
    from semanticsegmentationactivelearning_b200 import rank_confidence
    feat, kern = _case(19, 24, 16, 160, seed=7)
    logits = torch.from_numpy(R.final_head(feat, kern)).cuda()
    unl = np.arange(2, 24)
    ids_a, conf_a = rank_confidence(logits, unl, 6, "entropy", scorer=scorer, batch_size=8)
    ids_b, conf_b = rank_confidence(torch.from_numpy(feat).cuda(), unl, 6, "entropy", scorer=scorer, batch_size=8,
                                    head_kernel=kern)
    ids_c, conf_c = rank_confidence(feat, unl, 6, "entropy", scorer=scorer, batch_size=8, head_kernel=kern)  # host batches
    np.testing.assert_allclose(conf_b, conf_a, rtol=RTOL)
    assert np.array_equal(conf_b, conf_c)
    assert sorted(ids_a.tolist()) == sorted(ids_b.tolist()) == sorted(ids_c.tolist())


def test_fused_head_validation(torch, scorer):
    with pytest.raises(ValueError):
        scorer.prepare_head(np.zeros((3, 3, 19, 8), np.float32))
    with pytest.raises(NotImplementedError):
        scorer.prepare_head(np.zeros((3, 3, 33, 16), np.float32))       # fused kernels exist for 2 <= C <= 32
    assert all(scorer.head_supported(c, m) for c in range(2, 33) for m in MEASURES)
    assert not scorer.head_supported(33) and not scorer.head_supported(19, "variance")
    # T > 1 (Monte-Carlo samples): class counts 2..24, all four measures
    assert all(scorer.head_supported(c, m, 8) for c in range(2, 25) for m in MEASURES + ("variance",))
    assert not scorer.head_supported(25, "entropy", 2)
    scorer.prepare_head(np.zeros((3, 3, 19, 16), np.float32))
    f = torch.zeros((1, 4, 4, 16), device="cuda")
    with pytest.raises(NotImplementedError):
        scorer.score_features(f, "bald")
    with pytest.raises(ValueError):
        scorer.score_features(f, "variance")                             # T = 1
    with pytest.raises(NotImplementedError):
        scorer.prepare_head(np.zeros((3, 3, 30, 16), np.float32))
        scorer.score_features(torch.zeros((2, 1, 4, 4, 16), device="cuda"), "entropy")   # T > 1 needs C <= 24
    scorer.prepare_head(np.zeros((3, 3, 19, 16), np.float32))
    with pytest.raises(ValueError):
        scorer.score_features(torch.zeros((1, 4, 4, 8), device="cuda"), "entropy")
    # all-zero kernel -> uniform softmax -> entropy confidence 0, max-prob 1/C
    np.testing.assert_allclose(scorer.score_features(f, "confidence").cpu().numpy(), [1.0 / 19], rtol=1e-6)
