"""CPU: the C oracle (oracle/abides_oracle.c) against golden vectors produced by the unmodified
reference (tests/golden/gen_golden.py).  Bit-exact: trace, ttl_messages, finals, fundamental log."""
import numpy as np
import pytest

import oracle
from abides_b200 import lowering
from parity_util import load_golden, load_spec, first_trace_mismatch, describe_mismatch


@pytest.mark.parametrize("name", ["zi100_s123456789"])
def test_oracle_matches_reference_run(name):
    g, meta = load_golden(name)
    hb = lowering.HostBatch([load_spec(name)])
    r = oracle.run_instance(hb, 0, trace_cap=meta["delivered"] + 16, flog_cap=len(g["f_time"]) + 16)
    assert r["status"] == 0
    assert r["ttl_messages"] == meta["ttl_messages"]
    assert r["delivered"] == meta["delivered"]
    i = first_trace_mismatch(r["trace"], g["trace"])
    assert i < 0, describe_mismatch(r["trace"], g["trace"], i, ("oracle", "reference"))
    for k in ("cash", "holdings", "mtm"):
        assert np.array_equal(r[k], g[k]), k
    fv, gfv = r["final_valuation"], g["final_valuation"]
    assert np.array_equal(np.isnan(fv), np.isnan(gfv))
    assert np.array_equal(fv[~np.isnan(fv)], gfv[~np.isnan(gfv)])
    assert np.array_equal(r["f_time"], g["f_time"])
    # oracle floats: fundamental values are integral after int(round(.)); 1e-9 relative bound stated by the north star
    assert np.allclose(r["f_value"], g["f_value"], rtol=1e-9, atol=0)
    assert np.array_equal(r["f_value"], g["f_value"])
    assert r["last_trade"] == meta["last_trade"]
    assert r["final_time"] == meta["final_time"]
    assert r["slowest_agent_finish_time"] == meta["slowest_agent_finish_time"]
