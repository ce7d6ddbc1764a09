"""The oracle against the fixtures produced by executing the reference's own source
(tests/golden/make_golden.py; active_learning.py:40, :234-269, :682-715)."""
import numpy as np
import pytest

from oracle import reference_np as R

MEASURES = ("entropy", "margin", "confidence")


def test_fixture_provenance(golden):
    eline, g0, g1, r0, r1 = golden["meta_reference_lines"].tolist()
    assert eline == 40 and (g0, g1) == (234, 269) and (r0, r1) == (682, 715)


@pytest.mark.parametrize("measure", MEASURES)
def test_graph_matches_reference_source(golden, measure):
    for name in golden["graph_cases"].tolist():
        logits = golden[f"{name}.logits"]
        conf = R.pixel_confidence(logits, measure, flavour="gpu")
        want = golden[f"{name}.{measure}.conf"]
        assert conf.dtype == np.float32 and conf.shape == want.shape
        np.testing.assert_array_equal(conf, want, err_msg=f"{name}/{measure}")
        mean = R.image_scores(conf)
        assert mean.dtype == np.float64
        np.testing.assert_array_equal(mean, golden[f"{name}.{measure}.mean"])
        np.testing.assert_array_equal(R.pseudo_mask(conf, 0.9), golden[f"{name}.{measure}.mask"])
        np.testing.assert_array_equal(R.pseudo_label(logits), golden[f"{name}.label"])


def test_softmax_flavours_agree_to_an_ulp(golden):
    logits = golden["rand0.logits"]
    a, b = R.softmax(logits, "gpu"), R.softmax(logits, "cpu")
    assert np.max(np.abs(a - b)) <= 2 * np.finfo(np.float32).eps


def test_unknown_measure(golden):
    with pytest.raises(NotImplementedError) as ei:
        R.pixel_confidence(golden["rand0.logits"], "bald")
    assert str(ei.value) == str(golden["unknown_measure_message"])


def test_known_answers(golden):
    for C in (19, 6, 3):
        x = golden[f"special_c{C}.logits"]
        ent = R.pixel_confidence(x, "entropy")[0, 0]
        mar = R.pixel_confidence(x, "margin")[0, 0]
        con = R.pixel_confidence(x, "confidence")[0, 0]
        # uniform (rows 0, 1)
        assert abs(ent[0]) <= 3e-7 and abs(ent[1]) <= 3e-7
        assert mar[0] == 0 and mar[1] == 0
        np.testing.assert_allclose(con[:2], 1.0 / C, rtol=2e-7)
        # one dominant logit (row 2)
        assert ent[2] == 1 and mar[2] == 1 and con[2] == 1
        # exact two-way tie (row 3)
        np.testing.assert_allclose(ent[3], 1 - np.log(2) / np.log(C), rtol=0, atol=2e-7)
        assert mar[3] == 0 and con[3] == 0.5
        # permutation / shift invariance of the ramp (rows 4, 5, 6)
        np.testing.assert_allclose(ent[4], ent[5], rtol=0, atol=2e-7)
        # (exact up to the class-sum order)
        np.testing.assert_allclose([mar[4], con[4]], [mar[5], con[5]], rtol=0, atol=1e-7)
        np.testing.assert_allclose([ent[6], mar[6], con[6]], [ent[4], mar[4], con[4]], rtol=0, atol=3e-7)
        # masked class behaves like C-1 uniform classes
        np.testing.assert_allclose(con[7], 1.0 / (C - 1), rtol=2e-7)
        assert np.isfinite(ent[7])


def test_rank_confidence_matches_reference_source(golden):
    for name in golden["rank_cases"].tolist():
        scores64 = golden[f"{name}.scores64"]
        order = golden[f"{name}.order"]
        bs = int(golden[f"{name}.batch_size"])
        trunc = int(golden[f"{name}.truncate_after"])
        batches = [(scores64[order][i:i + bs], order[i:i + bs]) for i in range(0, len(order), bs)]
        if trunc >= 0:
            batches = batches[:trunc]
        conf = R.scatter_scores(int(golden[f"{name}.num_examples"]), batches)
        ids, uconf = R.select_lowest(conf, golden[f"{name}.unlabelled"], int(golden[f"{name}.k"]))
        np.testing.assert_array_equal(uconf, golden[f"{name}.uconf"])
        # np.argpartition is deterministic for identical input, so even the order matches
        np.testing.assert_array_equal(ids, golden[f"{name}.ids"])
        # and the deterministic completion picks the same set whenever there is no boundary tie
        ids2, _ = R.select_lowest_total_order(conf, golden[f"{name}.unlabelled"], int(golden[f"{name}.k"]))
        k = len(ids)
        u = {int(i): float(s) for i, s in zip(golden[f"{name}.unlabelled"], uconf)}
        kth = sorted(np.nan_to_num(list(u.values()), nan=np.inf))[k - 1] if k else None
        ties = [i for i, s in u.items() if np.nan_to_num(s, nan=np.inf) == kth]
        if len(ties) <= 1:
            assert sorted(ids.tolist()) == sorted(ids2.tolist()), name
        else:
            strictly = {i for i, s in u.items() if np.nan_to_num(s, nan=np.inf) < kth}
            assert strictly <= set(ids.tolist()) and strictly <= set(ids2.tolist())
            assert set(ids.tolist()) - strictly <= set(ties) and set(ids2.tolist()) - strictly <= set(ties)


def test_rank_k_equal_len_raises_like_reference(golden):
    assert str(golden["rank_k_eq_m_raises"]) == "ValueError"
    with pytest.raises(ValueError):
        R.select_lowest(np.arange(10, dtype=np.float32), np.arange(10), 10)
    ids, _ = R.select_lowest_total_order(np.arange(10, dtype=np.float32), np.arange(10), 10)
    assert ids.tolist() == list(range(10))


def test_variance_spec_against_f64_truth():
    from oracle import synth
    x = synth.synth_logits(8, 0, 3, 8, 16, 19)
    got = R.pixel_confidence(x, "variance")
    truth = R.pixel_confidence_f64(x, "variance")
    np.testing.assert_allclose(got, truth, rtol=1e-5, atol=0)
    np.testing.assert_allclose(1 - got, 1 - truth, rtol=0, atol=2e-7)
    for m in MEASURES:
        np.testing.assert_allclose(R.pixel_confidence(x, m), R.pixel_confidence_f64(x, m), rtol=0, atol=5e-7)
    with pytest.raises(ValueError):
        R.pixel_confidence(x[0], "variance")
