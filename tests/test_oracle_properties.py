"""Property tests of the CPU oracle (oracle/reference_np.py) against its float64 truth model and against plain
Python orderings -- SURVEY.md section 8(c): analytic known answers, invariances and error budgets.  CPU only."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import reference_np as R

MEASURES = ("entropy", "margin", "confidence")
EPS32 = float(np.finfo(np.float32).eps)


def _logits(seed, n, h, w, c, scale):
    rng = np.random.default_rng(seed)
    return (scale * rng.standard_normal((n, h, w, c))).astype(np.float32)


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), c=st.integers(2, 40), scale=st.sampled_from([0.0, 0.1, 1.0, 4.0, 30.0]),
       measure=st.sampled_from(MEASURES))
def test_fp32_graph_within_budget_of_f64_truth(seed, c, scale, measure):
    x = _logits(seed, 2, 5, 7, c, scale)
    truth = R.pixel_confidence_f64(x, measure)
    for flavour in ("cpu", "gpu"):
        got = R.pixel_confidence(x, measure, flavour)
        # the reference's own op-by-op fp32 evaluation: a few ulp of 1.0 (entropy sums C terms)
        assert np.max(np.abs(got - truth)) <= (8 + c) * EPS32, (flavour, measure, c, scale)
    # the two TF softmax flavours (x * 1/S on the CPU, x / S on the GPU) differ by rounding only
    assert np.max(np.abs(R.pixel_confidence(x, measure, "cpu") - R.pixel_confidence(x, measure, "gpu"))) <= 8 * EPS32


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), c=st.integers(2, 24), t=st.integers(2, 9),
       measure=st.sampled_from(MEASURES + ("variance",)))
def test_mc_samples_within_budget_of_f64_truth(seed, c, t, measure):
    rng = np.random.default_rng(seed)
    x = (2.0 * rng.standard_normal((t, 2, 3, 4, c))).astype(np.float32)
    got = R.pixel_confidence(x, measure)
    truth = R.pixel_confidence_f64(x, measure)
    assert np.max(np.abs(got - truth)) <= (8 + c) * EPS32
    if measure == "variance":
        assert np.all(got <= 1.0 + 4 * EPS32) and np.all(got >= 1.0 / c - 4 * EPS32)     # v in [0, 1 - 1/C]


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), c=st.integers(2, 33), shift=st.floats(-50, 50), measure=st.sampled_from(MEASURES))
def test_shift_and_permutation_invariance(seed, c, shift, measure):
    x = _logits(seed, 1, 4, 6, c, 3.0)
    base = R.pixel_confidence(x, measure)
    shifted = R.pixel_confidence((x + np.float32(shift)).astype(np.float32), measure)
    # adding a constant to every logit changes x - max(x) only through the rounding of x + shift
    tol = 64 * EPS32 * (1.0 + abs(shift))
    assert np.max(np.abs(shifted - base)) <= tol
    perm = np.random.default_rng(seed + 1).permutation(c)
    permuted = R.pixel_confidence(np.ascontiguousarray(x[..., perm]), measure)
    if measure == "entropy":
        assert np.max(np.abs(permuted - base)) <= (4 + c) * EPS32          # sum order
    else:
        assert np.array_equal(permuted, base) or np.max(np.abs(permuted - base)) <= 4 * EPS32


def test_analytic_known_answers():
    """uniform logits, one dominant logit, an exact two-way tie."""
    for c in (2, 6, 19, 66):
        uni = np.full((1, 2, 2, c), 3.25, np.float32)
        assert np.max(np.abs(R.pixel_confidence(uni, "entropy"))) <= 4e-7
        assert np.all(R.pixel_confidence(uni, "margin") == 0)
        np.testing.assert_allclose(R.pixel_confidence(uni, "confidence"), 1.0 / c, rtol=2e-7)
        dom = np.zeros((1, 2, 2, c), np.float32)
        dom[..., 0] = 100.0
        for m in MEASURES:
            np.testing.assert_allclose(R.pixel_confidence(dom, m), 1.0, atol=2e-7)      # 0 * log(tiny) = 0 (:243)
        tie = np.full((1, 1, 3, c), -1e4, np.float32)
        tie[..., :2] = 7.0
        np.testing.assert_allclose(R.pixel_confidence(tie, "entropy"), 1.0 - np.log(2.0) / np.log(float(c)) if c > 2 else 0.0,
                                   atol=4e-7)
        assert np.all(R.pixel_confidence(tie, "margin") == 0)
        np.testing.assert_allclose(R.pixel_confidence(tie, "confidence"), 0.5, rtol=2e-7)
    const = np.full((3, 8, 8), 0.625, np.float32)
    assert np.array_equal(R.image_scores(const), np.full(3, 0.625))                      # mean of a constant map


@settings(max_examples=60, deadline=None)
@given(data=st.data())
def test_total_order_selection_against_python_sort(data):
    m = data.draw(st.integers(1, 60))
    pool = data.draw(st.lists(st.sampled_from([0.0, -0.0, 0.25, 0.25, 0.5, 1.0, float("inf"), float("-inf"), float("nan"), -3.5]),
                              min_size=m, max_size=m))
    conf = np.asarray(pool, np.float32)
    unl = np.asarray(data.draw(st.permutations(list(range(m)))), np.int64)[: data.draw(st.integers(1, m))]
    k = data.draw(st.integers(0, m + 2))
    ids, u = R.select_lowest_total_order(conf, unl, k)
    assert np.array_equal(u, conf[unl], equal_nan=True)
    key = lambda i: (1, 0.0, i) if np.isnan(conf[i]) else (0, float(conf[i]) + 0.0, i)
    assert ids.tolist() == sorted(unl.tolist(), key=key)[: min(k, len(unl))]
    # wherever np.argpartition is defined (k < len) the reference's own selection picks the same score multiset
    if 0 < k < len(unl) and not np.isnan(conf[unl]).any():
        ref_ids, _ = R.select_lowest(conf, unl, k)
        assert sorted(conf[ref_ids].tolist()) == sorted(conf[ids].tolist())


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), c=st.integers(2, 12), h=st.integers(1, 5), w=st.integers(1, 6))
def test_transposed_conv_is_the_adjoint_of_the_strided_conv(seed, c, h, w):
    """<conv2d_transpose(y, K), z> == <y, conv2d(z, K, stride 2, SAME)> -- the defining property of
    tf.nn.conv2d_transpose (Final.call, models/enet/enet_modules.py:1359-1381), checked in float64."""
    rng = np.random.default_rng(seed)
    y = rng.standard_normal((1, h, w, 16))
    kern = rng.standard_normal((3, 3, c, 16))
    z = rng.standard_normal((1, 2 * h, 2 * w, c))
    up = R.conv2d_transpose_same(y, kern, dtype=np.float64)
    # forward SAME conv, stride 2, 3x3, input 2h x 2w -> h x w: pad_total = 1 -> (0 before, 1 after)
    zp = np.pad(z, ((0, 0), (0, 1), (0, 1), (0, 0)))
    down = np.zeros((1, h, w, 16))
    for ky in range(3):
        for kx in range(3):
            patch = zp[:, ky:ky + 2 * h:2, kx:kx + 2 * w:2, :]                  # [1,h,w,c]
            down += np.einsum("nhwc,ck->nhwk", patch, kern[ky, kx])
    np.testing.assert_allclose(np.sum(up * z), np.sum(y * down), rtol=1e-10, atol=1e-9)
