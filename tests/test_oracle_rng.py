"""CPU tier: the oracle's RNG restatement (ISAAC64 + Ziggurat) against golden vectors produced by the
reference's own src/rng.c (tests/golden/gen_rng_golden.py) and, when oracle/_ref is present, against the
compiled reference itself."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle_py as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rng_kat.json")))


def _new(L, seed):
    return C.c_void_p(L.orc_rng_new(C.c_uint32(seed)))


@pytest.mark.parametrize("seed", sorted(GOLD["seeds"], key=int))
def test_oracle_rng_matches_reference_golden(seed):
    g = GOLD["seeds"][seed]
    s = int(seed)
    L = O.lib()
    r = _new(L, s)
    assert [L.orc_rng_uint(r) for _ in range(6)] == g["uint"]
    L.orc_rng_free(r)
    r = _new(L, s)
    assert [float(L.orc_rng_dbl(r)).hex() for _ in range(4)] == g["dbl"]
    L.orc_rng_free(r)
    r = _new(L, s)
    assert [float(L.orc_rng_gauss(r)).hex() for _ in range(4)] == g["gauss"]
    L.orc_rng_free(r)
    r = _new(L, s)
    x = np.zeros(1000000)
    L.orc_rng_fill_gauss(r, C.c_void_p(x.ctypes.data), len(x))
    assert hashlib.sha256(x.tobytes()).hexdigest() == g["gauss_1e6_sha256"]
    assert int(L.orc_rng_uses(r)) == g["gauss_1e6_words_used"]
    assert float(x.sum()).hex() == g["gauss_1e6_sum"]
    assert float(np.abs(x).max()).hex() == g["gauss_1e6_absmax"]
    L.orc_rng_free(r)
    r = _new(L, s)
    u = np.zeros(5000, np.uint32)
    L.orc_rng_fill_uint(r, C.c_void_p(u.ctypes.data), len(u))
    assert hashlib.sha256(u.tobytes()).hexdigest() == g["uint_5000_sha256"]
    for k, v in g["uint_at"].items():
        assert int(u[int(k)]) == v
    L.orc_rng_free(r)


def test_survey_a3_known_answers():
    """SURVEY.md §A.3 (decimal form)."""
    L = O.lib()
    r = _new(L, 1)
    assert [L.orc_rng_uint(r) for _ in range(6)] == [3785283026, 3399003949, 2812382471, 3400768338, 2695561778, 3514871572]
    L.orc_rng_free(r)
    r = _new(L, 1)
    assert [L.orc_rng_gauss(r) for _ in range(4)] == [-1.5776702091810464, 1.0105580808054933, 0.38828352164332647, 1.4174081638342191]
    L.orc_rng_free(r)
    r = _new(L, 12345)
    assert [L.orc_rng_dbl(r) for _ in range(4)] == [0.90799094014801085, 0.46296899975277483, 0.53759316471405327, 0.89580700290389359]
    L.orc_rng_free(r)


def test_compiled_reference_rng_agrees_with_golden_and_oracle():
    R = O.ref_rng_lib()
    if R is None:
        pytest.skip("oracle/_ref/librefrng.so not present on this box")
    L = O.lib()
    for seed in (1, 77, 12345):
        a, b = _new(L, seed), C.c_void_p(R.ref_rng_new(seed))
        n = 300000
        x, y = np.zeros(n), np.zeros(n)
        L.orc_rng_fill_gauss(a, C.c_void_p(x.ctypes.data), n)
        R.ref_rng_fill_gauss(b, C.c_void_p(y.ctypes.data), n)
        assert (x == y).all()
        assert int(L.orc_rng_uses(a)) == int(R.ref_rng_uses(b))
        u, v = np.zeros(2000, np.uint32), np.zeros(2000, np.uint32)
        L.orc_rng_fill_uint(a, C.c_void_p(u.ctypes.data), 2000)
        R.ref_rng_fill_uint(b, C.c_void_p(v.ctypes.data), 2000)
        assert (u == v).all()
        L.orc_rng_free(a); R.ref_rng_free(b)


def test_ziggurat_tables_are_the_reference_tables():
    """Both copies of zig_tables.inc (oracle + product) carry the reference's tables bit for bit."""
    z = GOLD["ziggurat"]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in ("oracle/zig_tables.inc", "mcell_b200/csrc/zig_tables.inc"):
        txt = open(os.path.join(root, path)).read()
        import re

        def table(name, conv):
            body = re.search(r"#define MCX_ZIG_%s_INIT \{(.*?)\}" % name, txt, re.S).group(1).replace("\\", "")
            return [conv(t.strip()) for t in body.split(",") if t.strip()]

        Y = np.array(table("YTAB", float.fromhex))
        W = np.array(table("WTAB", float.fromhex))
        K = np.array(table("KTAB", lambda s: int(s.rstrip("u"))), dtype=np.uint64)
        assert hashlib.sha256(Y.tobytes()).hexdigest() == z["ytab_sha256"]
        assert hashlib.sha256(W.tobytes()).hexdigest() == z["wtab_sha256"]
        assert hashlib.sha256(K.tobytes()).hexdigest() == z["ktab_sha256"]
        r = re.search(r"#define MCX_ZIG_R (\S+)", txt).group(1)
        assert float.fromhex(r).hex() == z["R"]


def test_gauss_hard_bound_for_halo_width():
    """|gauss| < 9.89 bounds the per-step displacement (halo width of the slab decomposition)."""
    for g in GOLD["seeds"].values():
        assert float.fromhex(g["gauss_1e6_absmax"]) < 9.89
