"""Streamed Monte-Carlo accumulation (als_mc_begin / als_mc_add_sample / als_mc_finish): one dropout sample per call,
Welford state resident in HBM.  Must be BIT-IDENTICAL to the resident path (als_score on the stacked [T,N,H,W,C]
samples), which test_gpu_parity.py holds to the oracle's frozen spec (oracle/reference_np.py: welford_mean_m2 /
pixel_confidence; no reference counterpart -- MC dropout does not exist in /root/reference).
Also checked directly against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5
ATOL_PIX = 1e-6
ALL = ("entropy", "margin", "confidence", "variance")


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


SHAPES = [
    # (T, N, H, W, C): 19/6/66 = the BASELINE class counts, 150 = 8 lanes per pixel, 23 = generic kernels
    (4, 3, 16, 32, 19), (3, 2, 12, 20, 6), (4, 2, 8, 16, 66), (2, 2, 8, 8, 150), (3, 2, 7, 9, 23), (8, 2, 16, 16, 19),
    (16, 1, 12, 16, 66), (2, 5, 3, 3, 19), (5, 3, 33, 31, 19),
]


def _device(torch, x, dtype):
    xt = torch.from_numpy(x.view(np.int16) if dtype == "bfloat16" else x).cuda()
    return xt.view(torch.bfloat16) if dtype == "bfloat16" else xt


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "T%d_N%d_%dx%d_C%d" % s)
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_streamed_equals_resident_and_oracle(torch, scorer, shape, dtype):
    from oracle import reference_np as R, synth
    T, N, H, W, C = shape
    x = synth.synth_logits(T, 11, N, H, W, C, dtype=dtype, squeeze_t=False)
    xf = synth.bf16_bits_to_f32(x) if dtype == "bfloat16" else x
    xt = _device(torch, x, dtype)
    for measure in ALL:
        resident = scorer.pseudo_annotation(xt, measure, threshold=0.5)
        label = torch.empty((N, H, W), dtype=torch.uint8, device="cuda")
        scorer.mc_begin((N, H, W, C), dtype, label=label)
        for t in range(T):
            scorer.mc_add_sample(xt[t])
        out = scorer.mc_finish(measure, threshold=0.5, want_maps=True)
        # bit-identical whenever both paths run the same arithmetic.  A stack whose sample planes are not 16-byte
        # aligned (odd N*H*W*C) sends the RESIDENT call to the generic kernel (classic Welford form), while the streamed
        # call copies each misaligned sample into an aligned buffer and keeps the tiled kernels: then only the tolerance
        # against the oracle below applies.
        es = 2 if dtype == "bfloat16" else 4
        if (N * H * W * C * es) % 16 == 0:
            for key in ("pseudo_confidence", "pseudo_mean_confidence", "pseudo_label", "pseudo_mask"):
                assert torch.equal(out[key], resident[key]), "%s: streamed %s differs from the resident path" % (measure, key)
        else:
            assert torch.equal(out["pseudo_label"], resident["pseudo_label"])
        want = R.pixel_confidence(xf, measure)
        got = out["pseudo_confidence"].cpu().numpy()
        assert np.all(np.abs(got - want) <= RTOL * np.abs(want) + ATOL_PIX), measure
        np.testing.assert_allclose(out["pseudo_mean_confidence"].cpu().numpy(), R.image_scores(want), rtol=RTOL, atol=0)
        # scores only (the ranked path): same numbers
        scorer.mc_begin((N, H, W, C), dtype)
        for t in range(T):
            scorer.mc_add_sample(xt[t])
        assert torch.equal(scorer.mc_finish(measure), out["pseudo_mean_confidence"])


def test_streamed_host_samples_and_pool_scatter(torch, scorer):
    """Host samples are staged; finish scatters float32(score) into the pool vector by example index (:700)."""
    from oracle import reference_np as R, synth
    T, N, H, W, C = 4, 6, 16, 24, 19
    x = synth.synth_logits(T, 0, N, H, W, C)
    want = R.score_pool(x, "variance")
    idx = np.array([7, 2, 9, 0, 4, 11])
    scorer.pool_begin(12)
    scorer.mc_begin((N, H, W, C))
    for t in range(T):
        scorer.mc_add_sample(np.ascontiguousarray(x[t]))
    s = scorer.mc_finish("variance", batch_indices=idx)
    np.testing.assert_allclose(s.cpu().numpy(), want, rtol=RTOL)
    pool = scorer.pool_scores(12)
    assert np.array_equal(pool[idx], s.cpu().numpy().astype(np.float32))
    assert np.all(pool[np.setdiff1d(np.arange(12), idx)] == 0)
    ids, u = scorer.pool_select(np.arange(12), 3)
    assert set(ids.tolist()) <= set(np.setdiff1d(np.arange(12), idx).tolist())     # unvisited 0.0 first (:685)


def test_streamed_misaligned_samples(torch, scorer):
    from oracle import reference_np as R, synth
    T, N, H, W, C = 3, 2, 5, 7, 19
    x = synth.synth_logits(T, 0, N, H, W, C)
    flat = torch.zeros(x[0].size + 1, device="cuda")
    scorer.mc_begin((N, H, W, C))
    for t in range(T):
        flat[1:] = torch.from_numpy(x[t]).cuda().reshape(-1)
        scorer.mc_add_sample(flat[1:].view(N, H, W, C))
        torch.cuda.synchronize()
    np.testing.assert_allclose(scorer.mc_finish("variance").cpu().numpy(), R.score_pool(x, "variance"), rtol=RTOL)


def test_streamed_state_errors(torch, scorer):
    x = torch.zeros(2, 4, 4, 19, device="cuda")
    with pytest.raises(RuntimeError):
        scorer._mc_shape, scorer._mc_dtype = (2, 4, 4, 19), 0
        scorer.mc_add_sample(x)                                  # no accumulation open
    scorer.mc_begin((2, 4, 4, 19))
    with pytest.raises(ValueError):
        scorer.mc_add_sample(torch.zeros(2, 4, 4, 6, device="cuda"))
    scorer.mc_add_sample(x)
    with pytest.raises(ValueError):
        scorer.mc_finish("variance")                             # one sample: variance undefined
    scorer.mc_begin((2, 4, 4, 19))
    scorer.mc_add_sample(x)
    with pytest.raises(NotImplementedError):
        scorer.mc_finish("bald")


def test_streamed_full_geometry_one_image(torch, scorer):
    """BASELINE config 2 geometry (T=8 @512x1024, C=19): streamed == resident, bit for bit."""
    x = scorer.synth_logits(8, 3, 2, 512, 1024, 19)
    resident = scorer.score(x, "variance")
    scorer.mc_begin((2, 512, 1024, 19))
    for t in range(8):
        scorer.mc_add_sample(x[t])
    assert torch.equal(scorer.mc_finish("variance"), resident)
