"""CPU checks for the fused classifier-head row:
  * the oracle's restatement of Final.call against the golden fixture made by executing the reference method,
  * the torch CPU restatement (bench baseline) against the NumPy oracle,
  * the host-side operand packing of csrc/head.cu: the four narrow GEMMs it describes, evaluated with NumPy on the
    packed hi + lo weights, must reproduce the transposed convolution (no GPU involved)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import reference_np as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(ROOT, "tests", "golden", "final_head.npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def test_oracle_final_head_matches_reference_final_call(gold):
    assert list(gold["meta_reference_lines"]) == [1359, 1381]
    for name in gold["cases"]:
        got = R.final_head(gold[name + ".features"], gold[name + ".kernel"])
        want = gold[name + ".logits"]
        assert got.shape == want.shape and got.dtype == np.float32
        np.testing.assert_allclose(got, want, rtol=0, atol=4e-6 * max(1.0, float(np.abs(want).max())))


def test_final_head_is_the_adjoint_of_the_same_padded_conv():
    """<conv_same(X), Y> == <X, conv_transpose(Y)> for the stride-2 SAME convolution TF differentiates."""
    rng = np.random.default_rng(3)
    Y = rng.standard_normal((2, 4, 5, 16))
    K = rng.standard_normal((3, 3, 7, 16))
    X = rng.standard_normal((2, 8, 10, 7))
    Xp = np.pad(X, ((0, 0), (0, 1), (0, 1), (0, 0)))          # SAME: one padding row / column at the bottom / right
    conv = np.zeros((2, 4, 5, 16))
    for i in range(4):
        for j in range(5):
            patch = Xp[:, 2 * i:2 * i + 3, 2 * j:2 * j + 3, :]            # [B,3,3,7]
            conv[:, i, j, :] = np.einsum("bklc,klco->bo", patch, K)
    lhs = float((conv * Y).sum())
    rhs = float((X * R.conv2d_transpose_same(Y, K, dtype=np.float64)).sum())
    assert abs(lhs - rhs) <= 1e-9 * max(1.0, abs(lhs))


def test_torch_restatement_matches_numpy_oracle():
    import torch
    from oracle import reference_torch as RT
    rng = np.random.default_rng(5)
    f = rng.standard_normal((2, 6, 9, 16)).astype(np.float32)
    k = (0.4 * rng.standard_normal((3, 3, 19, 16))).astype(np.float32)
    a = R.final_head(f, k)
    b = RT.final_head(torch.from_numpy(f), torch.from_numpy(k)).numpy()
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-5)
    sa = R.score_pool_from_features(f, k, "entropy")
    sb = RT.score_pool(RT.final_head(torch.from_numpy(f), torch.from_numpy(k)), "entropy").numpy()
    np.testing.assert_allclose(sa, sb, rtol=1e-6)


@pytest.mark.parametrize("Cn", [2, 6, 19, 21, 32])
def test_packed_operands_reproduce_the_transposed_convolution(Cn):
    from semanticsegmentationactivelearning_b200 import _lib
    lib = _lib.load()
    geom = (C.c_int32 * 14)()
    rows = C.c_int32()
    _lib.check(lib.als_head_geometry(Cn, geom, C.byref(rows)))
    g = list(geom)
    CB, n, col0, row0, rows = g[1], g[2:6], g[6:10], g[10:14], rows.value
    assert g[0] == Cn and CB % 4 == 0 and CB >= Cn and all(v % 16 == 0 for v in n) and all(v % 4 == 0 for v in col0)
    assert all(col0[o] + n[o] <= n[0] for o in range(4)), "operand 0 must initialise every accumulator column"
    rng = np.random.default_rng(Cn)
    kern = (0.4 * rng.standard_normal((3, 3, Cn, 16))).astype(np.float32)
    packed = np.zeros((2, 4, rows, 4), np.float32)
    _lib.check(lib.als_head_pack_weights(kern.ctypes.data, Cn, packed.ctypes.data, packed.size))
    hi, lo = packed[0], packed[1]
    # hi parts are tf32 numbers (13 low mantissa bits clear) and hi + lo recovers the weight to ~2^-22
    assert not np.any(hi.view(np.uint32) & 0x1fff) and not np.any(lo.view(np.uint32) & 0x1fff)
    B = (hi.astype(np.float64) + lo).transpose(1, 0, 2).reshape(rows, 16)      # [row][channel]
    feat = rng.standard_normal((1, 5, 7, 16)).astype(np.float32)
    h, w = feat.shape[1:3]
    fpad = np.zeros((h + 1, w + 1, 16))
    fpad[1:, 1:] = feat[0]                                                     # fpad[i+1, j+1] = Y[i, j]; row / column 0 = padding
    D = np.zeros((h, w, n[0]))
    for o in range(4):
        oy, ox = o & 1, o >> 1
        A = fpad[1 - oy:1 - oy + h, 1 - ox:1 - ox + w]                         # source pixel (i - oy, j - ox)
        D[:, :, col0[o]:col0[o] + n[o]] += A @ B[row0[o]:row0[o] + n[o]].T
    out = np.zeros((2 * h, 2 * w, Cn))
    for b, (dy, dx) in enumerate([(0, 1), (0, 0), (1, 0), (1, 1)]):
        out[dy::2, dx::2] = D[:, :, b * CB:b * CB + Cn]
    want = R.conv2d_transpose_same(feat, kern, dtype=np.float64)[0]
    np.testing.assert_allclose(out, want, rtol=0, atol=2e-6 * max(1.0, float(np.abs(want).max())))
    # padding columns of every block carry zero weights
    for b in range(4):
        assert not np.any(D[:, :, b * CB + Cn:(b + 1) * CB])


def test_head_support_matrix_is_decided_on_the_host():
    """als_head_supported (class count x measure x Monte-Carlo samples) needs no GPU: 2..32 classes for one sample,
    2..24 with T > 1 (the Welford state of a pixel pair has to fit the register file), variance only with T >= 2."""
    from semanticsegmentationactivelearning_b200 import _lib
    lib = _lib.load()
    ent, mar, con, var = 0, 1, 2, 3
    for c in range(2, 33):
        assert all(lib.als_head_supported(c, m, 1) == 1 for m in (ent, mar, con)), c
        assert lib.als_head_supported(c, var, 1) == 0
        for t in (2, 8, 16):
            want = 1 if c <= 24 else 0
            assert all(lib.als_head_supported(c, m, t) == want for m in (ent, mar, con, var)), (c, t)
    for c in (0, 1, 33, 66):
        assert lib.als_head_supported(c, ent, 1) == 0
    assert lib.als_head_supported(19, ent, 0) == 0 and lib.als_head_supported(19, 7, 1) == 0


def test_torch_restatement_runs_the_layer_per_monte_carlo_sample():
    """reference_torch.final_head on [T,N,h,w,16] = the layer applied to each sample (bench CPU baseline of cfg2h)."""
    import torch
    from oracle import reference_torch as RT
    rng = np.random.default_rng(5)
    feat = rng.standard_normal((3, 2, 4, 5, 16)).astype(np.float32)
    kern = (0.4 * rng.standard_normal((3, 3, 7, 16))).astype(np.float32)
    got = RT.final_head(torch.from_numpy(feat), torch.from_numpy(kern)).numpy()
    want = np.stack([R.final_head(feat[t], kern) for t in range(3)])
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)
    s_t = RT.score_pool(torch.from_numpy(want), "variance").numpy()
    np.testing.assert_allclose(s_t, R.score_pool(want, "variance"), rtol=1e-6)
