"""SURVEY 8a row a26 / north-star item 4: the dense MeanRevertingOracle series (util/oracle/MeanRevertingOracle.py:49-81)
as a kernel across series, against golden vectors the unmodified reference produced on nanosecond-scale windows
(tests/golden/gen_dense_oracle_golden.py).  Values are integer cents after rounding; the float path feeding them may
differ from glibc in the last bit of ~1e-4 of the normal deviates, hence the north star's 1e-9 relative tolerance on
the comparison (in practice the series are equal)."""
import os

import numpy as np
import pytest

import oracle
from parity_util import GOLDEN

G = np.load(os.path.join(GOLDEN, "dense_oracle.golden.npz"))
CASES = [tuple(r) for r in G["cases"]]


def _state_of(seed):
    np.random.seed(int(seed))
    return np.random.get_state()


def _close(a, b):
    return np.allclose(a.astype(np.float64), b.astype(np.float64), rtol=1e-9, atol=1.0)


@pytest.mark.parametrize("k", range(len(CASES)))
def test_oracle_dense_series_matches_reference(k):
    seed, r_bar, kappa, sigma_s, steps = CASES[k]
    st = _state_of(seed)
    a, st = oracle.dense_fundamental(st, r_bar, kappa, sigma_s, int(steps))
    b, st = oracle.dense_fundamental(st, r_bar / 2, kappa, sigma_s, int(steps))      # the second symbol continues the stream
    assert _close(a, G["case%d_A" % k]) and _close(b, G["case%d_B" % k])
    assert np.array_equal(a, G["case%d_A" % k]) and np.array_equal(b, G["case%d_B" % k])
    rs = np.random.RandomState()
    rs.set_state(st)
    assert rs.randint(0, 2 ** 32, dtype="uint64") == G["case%d_after" % k][0]          # the stream is left where the reference left it


@pytest.mark.gpu
def test_device_dense_series_matches_reference_and_oracle():
    from abides_b200.engine import Engine
    eng = Engine(0)
    for k, (seed, r_bar, kappa, sigma_s, steps) in enumerate(CASES):
        series, (st,) = eng.dense_fundamental([_state_of(seed)], r_bar, kappa, sigma_s, int(steps))
        assert _close(series[0], G["case%d_A" % k])
        assert np.array_equal(series[0], G["case%d_A" % k])
        series_b, (st,) = eng.dense_fundamental([st], r_bar / 2, kappa, sigma_s, int(steps))
        assert np.array_equal(series_b[0], G["case%d_B" % k])
    # many series at once (a seed sweep of oracles), each equal to the oracle's
    seeds = list(range(500, 756))
    series, _ = eng.dense_fundamental(seeds, 100000, 0.05, 1e6, 2000)
    for i in (0, 17, 255):
        want, _ = oracle.dense_fundamental(_state_of(seeds[i]), 100000, 0.05, 1e6, 2000)
        assert np.array_equal(series[i], want)
    assert len({tuple(s[:50]) for s in series}) == len(seeds)


@pytest.mark.gpu
def test_compat_mean_reverting_oracle_class_is_a_drop_in():
    """The reference's constructor signature and attributes; the global np.random stream ends where the reference's does."""
    import sys
    import pandas as pd
    from abides_b200 import runtime
    sys.path.insert(0, runtime.COMPAT_DIR)
    try:
        from util.oracle.MeanRevertingOracle import MeanRevertingOracle
    finally:
        sys.path.remove(runtime.COMPAT_DIR)
    seed, r_bar, kappa, sigma_s, steps = CASES[0]
    np.random.seed(int(seed))
    t0 = pd.Timestamp("2020-06-03 09:30:00")
    o = MeanRevertingOracle(t0, t0 + pd.Timedelta(int(steps), unit="ns"),
                            {"A": dict(r_bar=r_bar, kappa=kappa, sigma_s=sigma_s), "B": dict(r_bar=r_bar / 2, kappa=kappa, sigma_s=sigma_s)})
    assert np.array_equal(o.r["A"].to_numpy(), G["case0_A"]) and np.array_equal(o.r["B"].to_numpy(), G["case0_B"])
    assert np.random.randint(0, 2 ** 32, dtype="uint64") == G["case0_after"][0]
    assert o.getDailyOpenPrice("A") == int(r_bar)
    assert o.observePrice("A", t0 + pd.Timedelta(10, unit="ns"), sigma_n=0) == G["case0_A"][10]
