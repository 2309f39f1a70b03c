"""GPU: the CUDA engine (through the C ABI) against the reference goldens and the CPU oracle."""
import hashlib

import numpy as np
import pytest

import oracle
from abides_b200 import lowering, runtime
from parity_util import (load_golden, load_spec, first_trace_mismatch, describe_mismatch,
                         split_device_trace, device_run, agent_slices)
from test_oracle_golden import CASES, build_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_matches_reference_golden(name):
    """Full stochastic run with the reference's own RNG streams: every dispatched event, every fill,
    ttl_messages and every agent's final portfolio are bit-exact."""
    g, meta = load_golden(name)
    cap = meta["delivered"] + len(g["f_time"]) + 64
    hb = lowering.HostBatch([build_case(name)], trace_capacity=cap)
    eng, res, ares = device_run(hb)
    assert res["status"][0] == 0
    assert res["ttl_messages"][0] == meta["ttl_messages"]
    assert res["delivered"][0] == meta["delivered"]
    tr, n = eng.trace(0)
    assert n <= cap
    disp, fund = split_device_trace(tr)
    if "trace" in g.files:
        i = first_trace_mismatch(disp, g["trace"])
        assert i < 0, describe_mismatch(disp, g["trace"], i, ("device", "reference"))
    assert hashlib.sha256(np.ascontiguousarray(disp, np.int64).tobytes()).hexdigest() == meta["trace_sha256"]
    # oracle / fundamental floats: 1e-9 relative per the north star (values are integral, so in fact equal)
    assert np.array_equal(fund[:, 0], g["f_time"][1:])
    assert np.allclose(fund[:, 3].astype(np.float64), g["f_value"][1:], rtol=1e-9, atol=0)
    sl = agent_slices(hb, 0)
    assert np.array_equal(ares["cash"][sl], g["cash"])
    assert np.array_equal(ares["holdings"][sl], g["holdings"])
    assert np.array_equal(ares["marked_to_market"][sl], g["mtm"])
    fv, gfv = ares["final_valuation"][sl], g["final_valuation"]
    assert np.array_equal(np.isnan(fv), np.isnan(gfv))
    assert np.allclose(fv[~np.isnan(fv)], gfv[~np.isnan(gfv)], rtol=1e-9, atol=0)
    assert res["last_trade"][0] == meta["last_trade"]
    assert res["final_time"][0] == meta["final_time"]
    assert res["slowest_agent_finish_time"][0] == meta["slowest_agent_finish_time"]


# the kernel variant each golden's population lands on when it is loaded untraced (abides_engine_info)
WARP_VARIANT = {"zi100_s123456789": 2, "zi1000_s123456789": 1, "rmsc03_s1234_1000": 3, "rmsc03e_s1234_1100": 3,
                "rmsc01s_s77": 3, "rmsc01mm_s33": 4, "rmsc01_s33_1100": 4, "rmsc01_s34_1100": 4, "exec_twap_s6": 6,
                "exec_vwap_s5": 6, "execution_s8_e": 6, "rmsc02_s22": 5, "rmsc02sub_s21_1100": 5, "rmsc02fast_s23_1100": 5,
                "value_noise_s7": 0, "random_fund_value_s7": 3, "random_fund_diverse_s7": 0}
LANE_VARIANT = {"zi100_s123456789": 17, "zi1000_s123456789": 16, "rmsc03_s1234_1000": 18, "rmsc03e_s1234_1100": 18,
                "rmsc01s_s77": 18, "rmsc01mm_s33": 19, "exec_twap_s6": 19, "exec_vwap_s5": 19, "execution_s8_e": 19,
                "value_noise_s7": 19, "random_fund_value_s7": 18, "random_fund_diverse_s7": 19, "marketreplay_s5": 19}


@pytest.mark.parametrize("engine", ["warp", "lane"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_untraced_lean_variant_matches_reference_golden(name, engine):
    """The kernels the bench runs, pinned to the reference directly: each golden loaded WITHOUT a trace lands on its
    lean kernel variant (asserted through abides_engine_info); ttl_messages, deliveries, every agent's final
    portfolio and the last trade must equal what the unmodified reference produced (Kernel.py:204,:318-322)."""
    if engine == "lane" and name not in LANE_VARIANT:
        pytest.skip("population runs on the warp engine only (HBL agents / market-data subscriptions)")
    if engine == "warp" and name not in WARP_VARIANT:
        pytest.skip("population runs on the lane engine only (MarketReplayAgent)")
    g, meta = load_golden(name)
    spec = build_case(name)
    hb = lowering.HostBatch([spec, spec, spec])             # a small homogeneous batch
    eng, res, ares = device_run(hb, engine=engine)
    e, v = eng.engine_info()
    assert v == (WARP_VARIANT if engine == "warp" else LANE_VARIANT)[name], (e, v)
    for i in range(3):
        assert res["status"][i] == 0
        assert res["ttl_messages"][i] == meta["ttl_messages"]
        assert res["delivered"][i] == meta["delivered"]
        assert res["last_trade"][i] == meta["last_trade"]
        assert res["final_time"][i] == meta["final_time"]
        assert res["slowest_agent_finish_time"][i] == meta["slowest_agent_finish_time"]
        sl = agent_slices(hb, i)
        assert np.array_equal(ares["cash"][sl], g["cash"])
        assert np.array_equal(ares["holdings"][sl], g["holdings"])
        assert np.array_equal(ares["marked_to_market"][sl], g["mtm"])
        fv, gfv = ares["final_valuation"][sl], g["final_valuation"]
        assert np.array_equal(np.isnan(fv), np.isnan(gfv))
        assert np.allclose(fv[~np.isnan(fv)], gfv[~np.isnan(gfv)], rtol=1e-9, atol=0)


def _compare_with_oracle(hb, res, ares):
    ores, oares = oracle.run_batch(hb)
    for i in range(hb.n_instances):
        assert res["status"][i] == 0
        for f in ("ttl_messages", "delivered", "final_time", "last_trade", "final_fundamental", "n_orders",
                  "n_fills", "traded_volume", "slowest_agent_finish_time"):
            assert res[f][i] == getattr(ores[i], f), (i, f)
    o = np.frombuffer(oares, dtype=ares.dtype)
    for f in ("cash", "holdings", "marked_to_market"):
        assert np.array_equal(ares[f], o[f]), f
    fv, ofv = ares["final_valuation"], o["final_valuation"]
    assert np.array_equal(np.isnan(fv), np.isnan(ofv))
    assert np.array_equal(fv[~np.isnan(fv)], ofv[~np.isnan(ofv)])


def test_device_seed_sweep_matches_oracle_zi100():
    seeds = runtime.parallel_seeds(11, 24)
    specs = runtime.build_specs_parallel("sparse_zi_100", [["-s", s] for s in seeds])
    hb = lowering.HostBatch(specs)
    eng, res, ares = device_run(hb)
    _compare_with_oracle(hb, res, ares)
    assert len(set(int(x) for x in res["ttl_messages"])) > 1      # different seeds, different days


def test_device_seed_sweep_matches_oracle_zi1000():
    seeds = runtime.parallel_seeds(12, 8)
    specs = runtime.build_specs_parallel("sparse_zi_1000", [["-s", s] for s in seeds])
    hb = lowering.HostBatch(specs)
    eng, res, ares = device_run(hb)
    _compare_with_oracle(hb, res, ares)


@pytest.mark.parametrize("config, extra", [
    ("rmsc01", ["--end-time", "10:30:00"]),                                   # kernel variant v_rmsc01
    ("rmsc02", ["--end-time", "10:30:00"]),                                   # subscriptions (full variant)
    ("execution", ["-t", "ABM", "-d", "20200603", "-e", "--execution-type", "vwap", "--num-noise", "200",
                   "--num-value", "30", "--num-momentum", "8"]),              # schedule execution agent (full variant)
    ("value_noise", []),                                                      # legacy latency + noise / value agents
])
def test_device_seed_sweep_matches_oracle_other_populations(config, extra):
    """Untraced seed sweeps (the lean kernel variants only run without a trace) against the CPU oracle, instance by
    instance, including the seeds on which the reference itself would raise (same status on both sides)."""
    seeds = runtime.parallel_seeds(21, 8)
    specs = runtime.build_specs_parallel(config, [list(extra) + ["-s", s] for s in seeds])
    hb = lowering.HostBatch(specs)
    eng, res, ares = device_run(hb)
    ores, oares = oracle.run_batch(hb)
    o = np.frombuffer(oares, dtype=ares.dtype)
    for i in range(hb.n_instances):
        assert res["status"][i] == ores[i].status, i
        for f in ("ttl_messages", "delivered", "final_time", "last_trade", "n_orders", "n_fills", "traded_volume"):
            assert res[f][i] == getattr(ores[i], f), (i, f)
        if res["status"][i] == 0:
            sl = agent_slices(hb, i)
            for f in ("cash", "holdings", "marked_to_market"):
                assert np.array_equal(ares[f][sl], o[f][sl]), (i, f)


def test_reset_reruns_identically():
    spec = load_spec("zi100_s123456789")
    hb = lowering.HostBatch([spec] * 3)
    eng, res, ares = device_run(hb)
    eng.reset()
    eng.run()
    res2, ares2 = eng.results_arrays()
    assert np.array_equal(res["ttl_messages"], res2["ttl_messages"])
    assert np.array_equal(ares["cash"], ares2["cash"])
    assert res2["ttl_messages"][0] == 21799


def test_device_parameter_grid_matches_oracle():
    """BASELINE configs[4]: an rmsc03 parameter grid is ONE heterogeneous batch (instances differ in agent counts,
    market-maker parameters and seeds)."""
    grid = {"--num-noise": [40, 80], "--mm-num-ticks": [5, 10], "--mm-pov": [0.01, 0.05]}
    args = runtime.grid_arg_lists(["-t", "ABM", "-d", "20200603", "--end-time", "09:45:00", "--num-value", "10",
                                   "--num-momentum", "5"], grid, runtime.parallel_seeds(5, 2))
    specs = runtime.build_specs_parallel("rmsc03", args)
    assert len(specs) == 16 and len(set(s.n_agents for s in specs)) == 2
    hb = lowering.HostBatch(specs)
    eng, res, ares = device_run(hb)
    ok = res["status"] == 0
    ores, oares = oracle.run_batch(hb)
    for i in range(hb.n_instances):
        assert res["status"][i] == ores[i].status                 # incl. instances where the reference would raise
        for f in ("ttl_messages", "delivered", "n_orders", "n_fills", "traded_volume", "last_trade"):
            assert res[f][i] == getattr(ores[i], f), (i, f)
    assert ok.sum() >= 12


def test_ragged_batch_of_different_configs_matches_every_golden():
    """One launch holding instances of DIFFERENT configs (agent classes, latency models, market hours, population
    sizes from 101 to 564 agents, with and without subscriptions / execution agents): each instance ends exactly as
    its reference run did.  The batch falls on the full kernel variant."""
    names = ["zi100_s123456789", "value_noise_s7", "rmsc02sub_s21_1100", "rmsc02fast_s23_1100", "rmsc01_s33_1100",
             "rmsc03_s1234_1000", "exec_vwap_s5", "zi100_s123456789"]
    hb = lowering.HostBatch([build_case(n) for n in names])
    eng, res, ares = device_run(hb)
    for i, n in enumerate(names):
        g, meta = load_golden(n)
        assert res["status"][i] == 0, n
        for f in ("ttl_messages", "delivered", "last_trade", "final_time", "slowest_agent_finish_time"):
            assert res[f][i] == meta[f], (n, f)
        sl = agent_slices(hb, i)
        assert np.array_equal(ares["cash"][sl], g["cash"]), n
        assert np.array_equal(ares["holdings"][sl], g["holdings"]), n
        assert np.array_equal(ares["marked_to_market"][sl], g["mtm"]), n
    # pausing the ragged batch anywhere changes nothing either
    from abides_b200.engine import Engine
    e2 = Engine(0)
    e2.load(hb)
    e2.run(until_time=min(ic.mkt_open for ic in hb.instances) + 20 * 60 * 10 ** 9)
    e2.run()
    r2, a2 = e2.results_arrays()
    assert np.array_equal(res["ttl_messages"], r2["ttl_messages"]) and np.array_equal(ares["cash"], a2["cash"])


def test_pause_and_resume_equals_one_run():
    """abides_run(until_time) parks the batch in HBM mid-day; resuming gives the same day as one uninterrupted run."""
    spec = load_spec("zi100_s123456789")
    hb = lowering.HostBatch([spec] * 2)
    eng, res, ares = device_run(hb)
    from abides_b200.engine import Engine
    e2 = Engine(0)
    e2.load(hb)
    mid = (hb.instances[0].mkt_open + hb.instances[0].mkt_close) // 2
    e2.run(until_time=hb.instances[0].mkt_open)          # before the first trade
    e2.run(until_time=mid)
    e2.run(until_time=mid)                                # idempotent: nothing left below `mid`
    e2.run()
    r2, a2 = e2.results_arrays()
    for f in ("ttl_messages", "delivered", "n_fills", "traded_volume", "last_trade", "final_time"):
        assert np.array_equal(res[f], r2[f]), f
    assert np.array_equal(ares["cash"], a2["cash"]) and np.array_equal(ares["holdings"], a2["holdings"])


def test_full_size_batch_invariants():
    """BASELINE configs[1] at full size (1024 x sparse_zi_1000): size-independent properties of a closed market,
    plus the oracle on a subset.  Cash only moves between agents and shares are only exchanged, so per instance
    sum(cash - starting_cash) == 0 and sum(holdings) == 0; every fill is reported to two agents."""
    n = 1024
    seeds = runtime.parallel_seeds(20261019, n)
    specs = runtime.build_specs_parallel("sparse_zi_1000", [["-s", s] for s in seeds])
    hb = lowering.HostBatch(specs)
    eng, res, ares = device_run(hb)
    assert (res["status"] == 0).all()
    assert (res["ttl_messages"] >= res["delivered"]).all() and (res["delivered"] > 100000).all()
    start = specs[0].types[specs[0].agent_type[1]]["starting_cash"]
    cash = ares["cash"].reshape(n, -1)[:, 1:]
    hold = ares["holdings"].reshape(n, -1)[:, 1:]
    assert ((cash - start).sum(axis=1) == 0).all()
    assert (hold.sum(axis=1) == 0).all()
    assert (np.abs(hold) % 100 == 0).all()                 # ZI agents trade 100-share lots
    assert (res["traded_volume"] == 100 * res["n_fills"]).all()
    # identical re-run after reset (no host traffic)
    eng.reset(); eng.run()
    r2, a2 = eng.results_arrays()
    assert np.array_equal(res["ttl_messages"], r2["ttl_messages"]) and np.array_equal(ares["cash"], a2["cash"])
    # the first 32 instances against the CPU oracle
    sub = lowering.HostBatch(specs[:32])
    ores, oares = oracle.run_batch(sub)
    assert [int(ores[i].ttl_messages) for i in range(32)] == [int(x) for x in res["ttl_messages"][:32]]
    o = np.frombuffer(oares, dtype=ares.dtype)
    assert np.array_equal(o["cash"], ares["cash"][:len(o)])


def test_load_rejects_malformed_batches():
    from abides_b200.engine import Engine, EngineError
    spec = load_spec("zi100_s123456789")
    hb = lowering.HostBatch([spec])
    hb.agent_type[0] = 1                                    # agent 0 must be the exchange
    eng = Engine(0)
    with pytest.raises(EngineError, match="exchange"):
        eng.load(hb)
    hb = lowering.HostBatch([spec])
    hb.c.n_agents_total += 1
    with pytest.raises(EngineError, match="n_agents_total"):
        eng.load(hb)
    with pytest.raises(EngineError, match="no batch loaded"):
        Engine(0).run()
    # inputs the device would dereference are range-checked before anything is allocated or launched
    hb = lowering.HostBatch([spec])
    hb.c.n_agents_total = -5
    with pytest.raises(EngineError, match="out of range"):
        eng.load(hb)
    hb = lowering.HostBatch([spec])
    hb.mt_pos[3] = 700
    with pytest.raises(EngineError, match="mt_pos"):
        eng.load(hb)
    hb = lowering.HostBatch([spec])
    hb.mt_pos[3] = -1
    with pytest.raises(EngineError, match="mt_pos"):
        eng.load(hb)
    z1000 = build_case("zi1000_s123456789")                 # legacy latency matrix + noise
    hb = lowering.HostBatch([z1000])
    hb.instances[0].n_latency_noise = 0
    with pytest.raises(EngineError, match="n_latency_noise"):
        eng.load(hb)
    hb = lowering.HostBatch([spec])
    hb.c.n_theta_total = 10                                 # the agents' theta rows would run past the table
    with pytest.raises(EngineError, match="theta"):
        eng.load(hb)
    hb = lowering.HostBatch([spec])
    eng.load(hb)                                            # and a well-formed batch still loads afterwards
    eng.run()
    assert eng.results_arrays()[0]["ttl_messages"][0] == 21799


def test_philox_throughput_mode_matches_oracle():
    """Counter-based Philox4x32-10 streams keyed by (instance seed, stream): no RNG state is uploaded; device and
    oracle implement the same keyed streams, so the runs are still bit-identical to each other (they are
    statistically, not bit-wise, equivalent to the reference's MT19937 runs)."""
    from abides_b200 import _abi
    seeds = runtime.parallel_seeds(21, 12)
    specs = runtime.build_specs_parallel("sparse_zi_100", [["-s", s] for s in seeds])
    hb = lowering.HostBatch(specs, rng_mode=_abi.RNG_PHILOX)
    assert hb.mt_key is None and hb.h2d_bytes() < 200000
    eng, res, ares = device_run(hb)
    _compare_with_oracle(hb, res, ares)
    mt = lowering.HostBatch(specs)
    _, res_mt, _ = device_run(mt)
    assert not np.array_equal(res["ttl_messages"], res_mt["ttl_messages"])          # different streams ...
    assert abs(res["ttl_messages"].mean() / res_mt["ttl_messages"].mean() - 1) < 0.05   # ... same market


def test_full_size_rmsc03_full_day_matches_oracle_on_the_lean_variant():
    """config/rmsc03.py at full size (5129 agents) over the whole day 09:30-16:00, 32 seeds, untraced -- i.e. on the
    lean rmsc kernel variant the bench runs -- against the oracle, instance by instance (3.6 M messages each)."""
    from abides_b200 import _abi
    seeds = runtime.parallel_seeds(31, 32)
    extra = ["-t", "ABM", "-d", "20200603", "--end-time", "16:00:00"]
    specs = runtime.build_specs_parallel("rmsc03", [extra + ["-s", s] for s in seeds])
    assert specs[0].n_agents == 5129
    hb = lowering.HostBatch(specs)
    eng, res, ares = device_run(hb, engine="warp")
    assert eng.engine_info() == (_abi.ENGINE_WARP, 3)
    ores, oares = oracle.run_batch(hb)
    o = np.frombuffer(oares, dtype=ares.dtype)
    n_ok = 0
    for i in range(hb.n_instances):
        assert res["status"][i] == ores[i].status, i
        for f in ("ttl_messages", "delivered", "final_time", "last_trade", "n_orders", "n_fills", "traded_volume",
                  "slowest_agent_finish_time", "final_fundamental"):
            assert res[f][i] == getattr(ores[i], f), (i, f)
        if res["status"][i] == 0:
            n_ok += 1
            sl = agent_slices(hb, i)
            for f in ("cash", "holdings", "marked_to_market"):
                assert np.array_equal(ares[f][sl], o[f][sl]), (i, f)
    assert n_ok >= 28 and int(res["ttl_messages"].min()) > 3000000
