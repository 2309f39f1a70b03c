#!/usr/bin/env python
"""Golden vectors of the DENSE mean-reverting fundamental from the unmodified reference
(util/oracle/MeanRevertingOracle.py:49-81) on nanosecond-scale toy windows.  Build container only.

    python tests/golden/gen_dense_oracle_golden.py        # writes tests/golden/dense_oracle.golden.npz

Compat shim: pandas 3 renamed date_range(closed=) to inclusive= and the 'N' alias to 'ns'."""
import os
import sys

import numpy as np
import pandas as pd

REF = os.environ.get("ABIDES_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [  # seed, r_bar, kappa, sigma_s, steps
    (1, 100000, 0.05, 1000000, 5000), (2, 100000, 1.67e-3, 400, 20000), (77, 5000, 0.3, 2.5e5, 3000),
    (12345, 100000, 0.05, 4e8, 4000),          # shocks large enough to hit the max(0, .) floor
]


def main():
    sys.path.insert(0, REF)
    orig = pd.date_range

    def date_range(*a, closed=None, freq=None, **kw):
        if closed is not None:
            kw["inclusive"] = closed
        return orig(*a, freq="ns" if freq == "N" else freq, **kw)

    pd.date_range = date_range
    from util.oracle.MeanRevertingOracle import MeanRevertingOracle
    import util.util
    util.util.silent_mode = True
    out = {}
    for k, (seed, r_bar, kappa, sigma_s, steps) in enumerate(CASES):
        np.random.seed(seed)
        t0 = pd.Timestamp("2020-06-03 09:30:00")
        o = MeanRevertingOracle(t0, t0 + pd.Timedelta(steps, unit="ns"),
                                {"A": dict(r_bar=r_bar, kappa=kappa, sigma_s=sigma_s), "B": dict(r_bar=r_bar / 2, kappa=kappa, sigma_s=sigma_s)})
        out["case%d_A" % k] = o.r["A"].to_numpy().astype(np.int64)
        out["case%d_B" % k] = o.r["B"].to_numpy().astype(np.int64)
        out["case%d_after" % k] = np.array([np.random.randint(0, 2 ** 32, dtype="uint64")], np.uint64)   # stream position check
    out["cases"] = np.array(CASES, np.float64)
    np.savez_compressed(os.path.join(HERE, "dense_oracle.golden.npz"), **out)
    print({k: (v.shape, int(v.min()), int(v.max())) for k, v in out.items() if k.endswith("_A")})


if __name__ == "__main__":
    main()
