"""Known-answer vectors from the reference's design chapter (docs/index.rst:1108-1372, figures img29-39), with the
"code wins" corrections of SURVEY.md section 4, against the oracle's restated functions."""
import numpy as np
import pytest

from oracle import codec_oracle as co

ORIG = np.array([60, 40, 20, 40, 20, 60, 20, 20, 60], dtype=np.int64)
PRED = np.array([60, 40, 25, 45, 10, 55, 25, 25, 60], dtype=np.int64)


def _eb_c(orig, diff, mode, value):
    """C restatement on a single 1 x n plane (frame index 1 of a two-frame window)."""
    X = np.zeros((2, 1, orig.size, 1), np.int64)
    D = np.zeros((2, 1, orig.size, 1), np.int64)
    X[1, 0, :, 0], D[1, 0, :, 0] = orig, diff
    co.error_bound_frames(X, D, mode, value)
    return D[1, 0, :, 0].copy()


@pytest.mark.parametrize("impl", ["py", "c"])
def test_error_bound_doc_examples(impl):
    fig_sign = ORIG - PRED          # docs/img/img29-32: error = orig - pred
    code_sign = PRED - ORIG         # compress.py:313: pred - orig
    f = (lambda o, d, m, v: co.error_bound_py(o.copy(), d.copy(), m, v)) if impl == "py" else _eb_c
    # img33 (pwrel 0.1): figure midpoints -3.5.. are truncated toward zero on assignment (compress.py:61,67)
    assert list(f(ORIG, fig_sign, "pwrel", [0.1])) == [-3, -3, -3, -3, 9, 9, -4, -4, -4]
    assert list(f(ORIG, code_sign, "pwrel", [0.1])) == [3, 3, 3, 3, -9, -9, 4, 4, 4]
    for mode, value in (("abs", [5.0]), ("rel", [0.1]), ("absrel", [5.0, 0.1]), ("abs", [2.55])):
        assert list(f(ORIG, fig_sign, mode, value)) == [-2, -2, -2, -2, 7, 7, -2, -2, -2], (mode, value)
    # abs 0 and abs 0.01 return the input unchanged
    assert list(f(ORIG, fig_sign, "abs", [0.0])) == list(fig_sign)
    assert list(f(ORIG, fig_sign, "abs", [0.01])) == list(fig_sign)
    assert list(f(ORIG, fig_sign, "absrel", [3.0, 0.0])) == list(fig_sign)     # compress.py:35


def test_error_bound_c_equals_python_random():
    rng = np.random.default_rng(0)
    for mode, value in (("abs", [0.5]), ("abs", [3.0]), ("abs", [7.9]), ("rel", [0.013]), ("absrel", [4.0, 0.02]),
                        ("absrel", [1.5, 0.5]), ("pwrel", [0.1]), ("pwrel", [0.017])):
        for _ in range(20):
            o = rng.integers(0, 256, 300).astype(np.int64)
            d = rng.integers(-30, 31, 300).astype(np.int64)
            assert np.array_equal(_eb_c(o, d, mode, value), co.error_bound_py(o.copy(), d.copy(), mode, value))


def test_error_bound_effective_guarantee():
    """error <= E for integer E; <= floor(E)+1 otherwise."""
    rng = np.random.default_rng(1)
    for E in (1.0, 2.0, 3.0, 5.0, 0.5, 2.55, 7.9):
        o = rng.integers(0, 256, 2000).astype(np.int64)
        d = rng.integers(-40, 41, 2000).astype(np.int64)
        q = _eb_c(o, d, "abs", [E])
        lim = E if float(E).is_integer() else np.floor(E) + 1
        assert np.abs(q - d).max() <= lim


def test_finding_difference_doc_examples():
    x = np.array([2, 5, 8, 8, 4, 4, 5, 6, 6], np.int16)
    y = co.delta_encode(x)
    assert list(y) == [2, -3, -3, 0, 4, 0, -1, -1, 0]            # img34/39 with the typo at index 4 corrected
    assert list(co.delta_decode(y)) == list(x)                   # img38


def test_replacing_doc_example_and_offset():
    arr = np.array([0, 5, 5, 5, 4, 5, 4, 4, 5], np.int16)
    table = np.array([5, 4, 0], np.int16)
    # without the 1600 offset values and indices collide exactly as docs/index.rst:1218-1232 describes
    assert list(co.replacing_encode(arr, table)) == [2, 2, 2, 2, 1, 2, 1, 1, 2]
    s = (1600 - arr).astype(np.int16)
    t = co.build_table(s)
    enc = co.replacing_encode(s, t)
    assert list(enc) == [2, 0, 0, 0, 1, 0, 1, 1, 0]              # img35/37: with the offset it is a pure LUT
    assert list(1600 - co.replacing_decode(enc, t)) == list(arr)


def test_table_tie_break():
    y = np.array([3, 3, -2, -2, 7, 0, 0, 1], np.int16)
    t = co.build_table((1600 - y).astype(np.int16))
    assert list(t) == [1597, 1600, 1602, 1593, 1599]             # count desc, ties ascending symbol
    assert list(1600 - t) == [3, 0, -2, 7, 1]


def test_padding():
    assert [co.padding_size(n) for n in (1, 7, 8, 9, 128, 130)] == [8, 8, 8, 16, 128, 136]
    X = np.ones((1, 2, 5, 9, 3), np.float32)
    P = co.data_padding(X)
    assert P.shape == (1, 2, 8, 16, 3) and P.dtype == np.float64 and P[0, 0, 5:].sum() == 0 and P[0, 0, :5, :9].all()
