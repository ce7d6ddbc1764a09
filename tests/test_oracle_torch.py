"""The reference arm's arithmetic (oracle/reference_torch.py, what `bench.py --impl reference` and `cpu_baseline`
time) against the pinned NumPy oracle (oracle/reference_np.py, itself bit-for-bit on the executed reference lines,
tests/test_oracle_golden.py).  torch's CPU softmax / log / topk kernels are a different libm from NumPy's, so values
agree to a few fp32 ulps, not bit for bit; the selection is identical.  Also pins tests/_oracle_par.py (the
thread-parallel driver the GPU parity tests use) to the plain oracle, bit for bit."""
import numpy as np
import pytest
import torch

from oracle import reference_np as R
from oracle import reference_torch as RT
from oracle import synth

from _oracle_par import pixel_confidence as par_conf

PIX_ATOL = 5e-7      # a few ulps of O(1) probabilities
SCORE_RTOL = 1e-6


@pytest.mark.parametrize("measure", ["entropy", "margin", "confidence"])
@pytest.mark.parametrize("C", [2, 6, 19, 66])
def test_single_pass_measures(measure, C):
    x = synth.synth_logits(1, 3, 3, 12, 20, C)
    want = R.pixel_confidence(x, measure)
    got = RT.pixel_confidence(torch.from_numpy(x), measure).numpy()
    assert got.dtype == np.float32 and got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=PIX_ATOL)
    np.testing.assert_allclose(RT.score_pool(torch.from_numpy(x), measure).numpy(), R.score_pool(x, measure), rtol=SCORE_RTOL)


@pytest.mark.parametrize("measure", ["entropy", "margin", "confidence", "variance"])
@pytest.mark.parametrize("T,C", [(2, 19), (8, 19), (16, 66), (3, 6)])
def test_monte_carlo_measures(measure, T, C):
    x = synth.synth_logits(T, 0, 2, 8, 12, C)
    want = R.pixel_confidence(x, measure)
    got = RT.pixel_confidence(torch.from_numpy(x), measure).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=PIX_ATOL)
    np.testing.assert_allclose(RT.score_pool(torch.from_numpy(x), measure).numpy(), R.score_pool(x, measure), rtol=SCORE_RTOL)


def test_golden_graph_cases_torch_arm(golden):
    """The fixtures produced by executing the reference's own statements (tests/golden/make_golden.py)."""
    for name in golden["graph_cases"].tolist():
        logits = golden[f"{name}.logits"]
        for measure in ("entropy", "margin", "confidence"):
            got = RT.pixel_confidence(torch.from_numpy(logits), measure).numpy()
            want = golden[f"{name}.{measure}.conf"]
            nan = np.isnan(want)
            assert np.array_equal(np.isnan(got), nan), (name, measure)
            np.testing.assert_allclose(got[~nan], want[~nan], rtol=1e-5, atol=PIX_ATOL, err_msg=f"{name}/{measure}")


def test_unknown_measure_message():
    with pytest.raises(NotImplementedError) as ei:
        RT.pixel_confidence(torch.zeros(1, 2, 2, 3), "bald")
    assert str(ei.value) == "Uncertainty function not implemented."


@pytest.mark.parametrize("measure,T", [("entropy", 1), ("margin", 1), ("confidence", 1), ("variance", 4)])
def test_rank_confidence_same_selection(measure, T):
    N, k = 40, 9
    x = synth.synth_logits(T, 0, N, 8, 16, 19)
    unl = np.sort(np.random.default_rng(2).choice(N, 31, replace=False))
    want_ids, want_u = R.rank_confidence(x, unl, k, measure, batch_size=8)
    got_ids, got_u = RT.rank_confidence(torch.from_numpy(x), unl, k, measure, batch_size=8)
    np.testing.assert_allclose(got_u, want_u, rtol=SCORE_RTOL)
    assert sorted(got_ids.tolist()) == sorted(want_ids.tolist())


def test_rank_confidence_with_head_kernel():
    rng = np.random.default_rng(5)
    feat = rng.standard_normal((6, 5, 7, 16)).astype(np.float32)
    kern = (0.4 * rng.standard_normal((3, 3, 19, 16))).astype(np.float32)
    want = R.score_pool_from_features(feat, kern, "entropy").astype(np.float32)
    ids, u = RT.rank_confidence(torch.from_numpy(feat), np.arange(6), 2, "entropy", head_kernel=torch.from_numpy(kern))
    np.testing.assert_allclose(u, want, rtol=1e-5)
    assert sorted(ids.tolist()) == sorted(np.argsort(want, kind="stable")[:2].tolist())


@pytest.mark.parametrize("shape,measure", [((3, 70, 16, 19), "entropy"), ((4, 2, 45, 8, 6), "variance"),
                                           ((2, 33, 8, 66), "margin"), ((3, 1, 64, 4, 19), "confidence")])
def test_parallel_driver_is_the_oracle(shape, measure):
    T = shape[0] if len(shape) == 5 else 1
    N, H, W, C = shape[-4:]
    x = synth.synth_logits(T, 0, N, H, W, C)
    assert np.array_equal(par_conf(x, measure), R.pixel_confidence(x, measure))
