"""GPU: the CUDA engine (through the C ABI) against the CPU oracle and the reference goldens."""
import numpy as np
import pytest

import oracle
from abides_b200 import lowering
from parity_util import (load_golden, load_spec, first_trace_mismatch, describe_mismatch,
                         split_device_trace, device_run, agent_slices)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["zi100_s123456789"])
def test_device_matches_reference_golden(name):
    g, meta = load_golden(name)
    cap = meta["delivered"] + len(g["f_time"]) + 64
    hb = lowering.HostBatch([load_spec(name)], trace_capacity=cap)
    eng, res, ares = device_run(hb)
    assert res["status"][0] == 0
    assert res["ttl_messages"][0] == meta["ttl_messages"]
    assert res["delivered"][0] == meta["delivered"]
    tr, n = eng.trace(0)
    assert n <= cap
    disp, fund = split_device_trace(tr)
    i = first_trace_mismatch(disp, g["trace"])
    assert i < 0, describe_mismatch(disp, g["trace"], i, ("device", "reference"))
    assert np.array_equal(fund[:, 0], g["f_time"][1:])
    assert np.array_equal(fund[:, 3].astype(np.float64), g["f_value"][1:])
    sl = agent_slices(hb, 0)
    assert np.array_equal(ares["cash"][sl], g["cash"])
    assert np.array_equal(ares["holdings"][sl], g["holdings"])
    assert np.array_equal(ares["marked_to_market"][sl], g["mtm"])
    fv, gfv = ares["final_valuation"][sl], g["final_valuation"]
    assert np.array_equal(np.isnan(fv), np.isnan(gfv))
    assert np.array_equal(fv[~np.isnan(fv)], gfv[~np.isnan(gfv)])
    assert res["last_trade"][0] == meta["last_trade"]
    assert res["final_time"][0] == meta["final_time"]
    assert res["slowest_agent_finish_time"][0] == meta["slowest_agent_finish_time"]


def test_device_batch_of_replicas_matches_oracle():
    name = "zi100_s123456789"
    spec = load_spec(name)
    hb = lowering.HostBatch([spec] * 6)
    eng, res, ares = device_run(hb)
    ores, oares = oracle.run_batch(hb, n_threads=2)
    for i in range(hb.n_instances):
        assert res["status"][i] == 0
        for f in ("ttl_messages", "delivered", "final_time", "last_trade", "final_fundamental", "n_orders",
                  "n_fills", "traded_volume", "slowest_agent_finish_time"):
            assert res[f][i] == getattr(ores[i], f), (i, f)
    o = np.frombuffer(oares, dtype=ares.dtype)
    for f in ("cash", "holdings", "marked_to_market"):
        assert np.array_equal(ares[f], o[f]), f
