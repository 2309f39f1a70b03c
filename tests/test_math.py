"""include/abides_dd_math.h: the table-driven log / exp / pow used by both the CUDA engine and the CPU
oracle, against (a) the slow double-double versions kept in the same header and (b) mpmath at 200 bits.
The production functions must be the correctly rounded double on (almost) every argument: the reference
pushes them through int(round(.)) ("Float paths feeding integer decisions")."""
import numpy as np
import pytest

import oracle

mp = pytest.importorskip("mpmath")


def _ulps(a, b):
    ia = np.ascontiguousarray(a, np.float64).view(np.int64)
    ib = np.ascontiguousarray(b, np.float64).view(np.int64)
    return np.abs(ia - ib)


def _cases():
    rs = np.random.RandomState(5)
    u = rs.random_sample(400000)
    log_x = np.concatenate([1.0 - u,                       # rng_exponential: log(1 - u)
                            u * u + 1e-300,                # gauss: log(r2)
                            np.exp(rs.uniform(-700, 700, 100000)),
                            1.0 + rs.uniform(-1, 1, 100000) * 2.0 ** rs.randint(-52, -1, 100000),
                            [1.0, 0.75, 1.5, 0.5, 2.0, np.nextafter(1.0, 0), np.nextafter(1.0, 2), 1 - 1.67e-15]])
    exp_x = np.concatenate([-rs.exponential(1.0, 300000), rs.uniform(-700, 700, 100000),
                            rs.uniform(-1, 1, 100000) * 2.0 ** rs.randint(-60, 0, 100000), [0.0, -0.0]])
    pow_x = np.concatenate([np.full(200000, 1 - 1.67e-15), np.full(100000, 1 - 1.67e-12), rs.uniform(0.5, 2, 100000)])
    pow_y = np.concatenate([np.floor(rs.uniform(0, 4.7e13, 200000)), np.floor(rs.uniform(0, 4.7e13, 100000)),
                            rs.uniform(-50, 50, 100000)])
    return log_x, exp_x, pow_x, pow_y


def test_fast_matches_slow_double_double():
    log_x, exp_x, pow_x, pow_y = _cases()
    for what, x, y in ((0, log_x, None), (1, exp_x, None), (2, pow_x, pow_y)):
        fast, slow = oracle.math_probe(what, x, y), oracle.math_probe(what + 10, x, y)
        d = _ulps(fast, slow)
        assert d.max() <= 1, (what, x[d.argmax()], fast[d.argmax()], slow[d.argmax()])
        assert (d != 0).mean() < 2e-3, (what, (d != 0).mean())     # both are "almost always correctly rounded"


def test_fast_is_correctly_rounded_vs_mpmath():
    mp.mp.prec = 200
    log_x, exp_x, pow_x, pow_y = _cases()
    rs = np.random.RandomState(6)
    bad = 0
    n = 0
    for what, x, y in ((0, log_x, None), (1, exp_x, None), (2, pow_x, pow_y)):
        idx = rs.choice(len(x), 4000, replace=False)
        got = oracle.math_probe(what, x[idx], None if y is None else y[idx])
        for j, i in enumerate(idx):
            xv = mp.mpf(float(x[i]))
            want = mp.log(xv) if what == 0 else mp.exp(xv) if what == 1 else mp.power(xv, mp.mpf(float(y[i])))
            w = float(want)                      # correctly rounded
            n += 1
            if got[j] != w:
                bad += 1
                assert _ulps(np.array([got[j]]), np.array([w]))[0] <= 1
    assert bad <= n * 2e-3, (bad, n)
