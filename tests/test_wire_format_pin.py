"""Wire-format pin: an independent Python encoder written from the Go sources (tests/golden/pyenc.py) must produce the SAME
BYTES as the oracle's C++ encoders for every page shape the reference can emit without a third-party compressor, and the
oracle's decoders must give the values back.  Two separate readings of the reference agreeing byte for byte is the strongest
pin available without a Go toolchain (the reference itself holds no golden bytes)."""
import os
import sys

import numpy as np
import pytest

import oracle
from opengemini_b200 import _lib as L

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import pyenc  # noqa: E402

T0, SEC = 1_700_000_000_000_000_000, 1_000_000_000


def _float_cases():
    rng = np.random.default_rng(7)
    yield "noise", 100 + rng.random(1000), None
    yield "walk_int_valued", np.cumsum(rng.integers(-2, 3, 1000)).astype(np.float64) + 1000, None
    yield "clz32_wrap_quirk", (np.uint64(0x4059000000000000) + rng.integers(0, 1000, 600).astype(np.uint64)).view(np.float64), None
    yield "same", np.full(333, 3.25), None
    yield "same_zero", np.zeros(100), None
    yield "rle", np.repeat([1.5, 2.5, 0.0, 7.0, -0.0], 120), None
    yield "rle_long_run", np.concatenate([np.full(20000, 2.0), np.full(7, 9.0)]), None
    yield "raw_short", np.array([1.0, 2.5, 3.25]), None
    yield "incompressible", np.where(np.isfinite(v := rng.integers(0, 2**63, 500).view(np.float64)), v, 1.0), None
    yield "nulls", 100 + rng.random(700), (rng.random(700) > 0.15).astype(np.uint8)
    yield "all_null", np.zeros(50), np.zeros(50, np.uint8)
    yield "one_row", np.array([42.5]), None
    yield "negatives_and_sign_flips", rng.standard_normal(400) * 1e6, None
    yield "window_growth", np.ldexp(1.0 + rng.random(300), rng.integers(-20, 20, 300)), None


def _int_cases():
    rng = np.random.default_rng(8)
    yield "const_delta", np.arange(500, dtype=np.int64) * -7 + 3, None
    yield "s8b_walk", np.cumsum(rng.integers(-1000, 1001, 1000)).astype(np.int64), None
    yield "s8b_runs_of_ones", np.concatenate([-np.arange(300, dtype=np.int64), [-295]]), None      # zigzag(delta) == 1 runs: selector 0/1 quirk
    yield "s8b_all_ones_240", -np.arange(242, dtype=np.int64), None
    yield "s8b_mixed_widths", np.cumsum(np.where(rng.random(900) < 0.1, rng.integers(-2**40, 2**40, 900), rng.integers(-3, 4, 900))).astype(np.int64), None
    yield "raw_pair", np.array([5, -9], np.int64), None
    yield "nulls", np.cumsum(rng.integers(-50, 51, 600)).astype(np.int64), (rng.random(600) > 0.2).astype(np.uint8)
    yield "one_row", np.array([-7], np.int64), None


def _bytes(a):
    return bytes(np.asarray(a, np.uint8))


@pytest.mark.parametrize("name,vals,valid", list(_float_cases()), ids=[c[0] for c in _float_cases()])
def test_float_pages_byte_for_byte(name, vals, valid):
    want = pyenc.field_page(pyenc.TYPE_FLOAT, [float(x) for x in vals], None if valid is None else [int(k) for k in valid])
    assert want is not None
    got = oracle.field_page_encode(L.TYPE_FLOAT, vals, valid)
    assert _bytes(got) == want, f"{name}: oracle {len(got)} B vs python {len(want)} B; first difference at {next((i for i, (a, b) in enumerate(zip(_bytes(got), want)) if a != b), None)}"
    dv, dk = oracle.field_page_decode(L.TYPE_FLOAT, got, cap=max(2000, len(vals) + 10))
    k = np.ones(len(vals), bool) if valid is None else valid.astype(bool)
    assert np.array_equal(dk, k) and np.array_equal(dv.view(np.uint64), np.asarray(vals)[k].view(np.uint64))


@pytest.mark.parametrize("name,vals,valid", list(_int_cases()), ids=[c[0] for c in _int_cases()])
def test_int_pages_byte_for_byte(name, vals, valid):
    want = pyenc.field_page(pyenc.TYPE_INT, [int(x) for x in vals], None if valid is None else [int(k) for k in valid])
    assert want is not None
    got = oracle.field_page_encode(L.TYPE_INT, vals, valid)
    assert _bytes(got) == want, name
    dv, dk = oracle.field_page_decode(L.TYPE_INT, got, cap=max(2000, len(vals) + 10))
    k = np.ones(len(vals), bool) if valid is None else valid.astype(bool)
    assert np.array_equal(dk, k) and np.array_equal(dv, np.asarray(vals)[k])


def test_bool_pages_byte_for_byte():
    rng = np.random.default_rng(9)
    for n, nulls in ((9, False), (1000, False), (777, True), (1, False)):
        b = (rng.random(n) < 0.5).astype(np.uint8)
        valid = (rng.random(n) > 0.3).astype(np.uint8) if nulls else None
        want = pyenc.field_page(pyenc.TYPE_BOOL, [int(x) for x in b], None if valid is None else [int(k) for k in valid])
        assert _bytes(oracle.field_page_encode(L.TYPE_BOOL, b, valid)) == want, (n, nulls)


def test_time_pages_byte_for_byte():
    rng = np.random.default_rng(10)
    cases = {"const": T0 + np.arange(1000, dtype=np.int64) * SEC,
             "s8b_scaled_ms": T0 + np.cumsum(rng.integers(1, 50, 1000) * 1_000_000).astype(np.int64),
             "s8b_unscaled": T0 + np.cumsum(rng.integers(1, 5000, 777)).astype(np.int64),
             "s8b_scale_drops": T0 + np.cumsum(np.concatenate([[10**9] * 500, [10**9 + 10**3], [10**9] * 100])).astype(np.int64),
             "raw_two": np.array([T0, T0 + 17], np.int64), "one": np.array([T0 + 5], np.int64),
             "negative_times": np.arange(-500, 500, dtype=np.int64) * 37}
    for name, t in cases.items():
        want = pyenc.time_page([int(x) for x in t])
        assert want is not None, name
        got = oracle.time_page_encode(t)
        assert _bytes(got) == want, name
        assert np.array_equal(oracle.time_page_decode(got), t), name
