"""CPU-side checks of the boundary: the library loads, exports every declared symbol, and the
host logic behaves like the reference (no compute calls -- there is no GPU here)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from semanticsegmentationactivelearning_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    from semanticsegmentationactivelearning_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "alscore.h")).read()
    declared = set(re.findall(r"^ALS_API\s+[\w\s\*]+?\b(als_\w+)\s*\(", hdr, flags=re.M))
    assert declared, "no prototypes found"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.als_version() == 110


def test_header_cites_reference_lines():
    hdr = open(os.path.join(ROOT, "include", "alscore.h")).read()
    for cite in ("active_learning.py:259-260", "active_learning.py:261-263", "active_learning.py:682-715",
                 "active_learning.py:234-236", "active_learning.py:265-269"):
        assert cite in hdr, cite


def test_measure_names_and_error_message(lib, golden):
    import semanticsegmentationactivelearning_b200 as A
    assert [A.measure_id(m) for m in ("entropy", "margin", "confidence", "variance")] == [0, 1, 2, 3]
    with pytest.raises(NotImplementedError) as ei:
        A.measure_id("bald")
    assert str(ei.value) == str(golden["unknown_measure_message"])


def test_no_cpu_fallback(lib):
    """Without a GPU the product path must fail loudly, never compute on the host."""
    import torch
    import semanticsegmentationactivelearning_b200 as A
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError) as ei:
        A.Scorer(0)
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(RuntimeError):
        A.rank_confidence(np.zeros((2, 4, 4, 19), np.float32), np.arange(2), 1, "entropy")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "semanticsegmentationactivelearning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "reference_np" not in src, f


def test_shard_bounds():
    from semanticsegmentationactivelearning_b200 import shard_bounds
    for n in (0, 1, 7, 64, 2975, 18000):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_bounds(2975, r, 8)[1] - shard_bounds(2975, r, 8)[0] for r in range(8)] == [372] * 7 + [371]
