"""CPU: the C oracle (oracle/abides_oracle.c) against golden vectors produced by the unmodified
reference (tests/golden/gen_golden.py).  Bit-exact: dispatch trace, ttl_messages, per-agent finals,
fundamental log.  The instance inputs come from this repo's own config builders and must hash to the
inputs the reference's builders produced (spec_sha256)."""
import hashlib
import sys
import os

import numpy as np
import pytest

import oracle
from abides_b200 import lowering, runtime
from parity_util import load_golden, load_spec, first_trace_mismatch, describe_mismatch, GOLDEN

sys.path.insert(0, GOLDEN)
from gen_golden import spec_digest  # noqa: E402

CASES = {
    "zi100_s123456789": ("sparse_zi_100", ["-s", "123456789"]),
    "zi1000_s123456789": ("sparse_zi_1000", ["-s", "123456789"]),
    "rmsc03_s1234_1000": ("rmsc03", ["-t", "ABM", "-d", "20200603", "-s", "1234", "--end-time", "10:00:00"]),
    # rmsc03 with the execution agent trading (`-e`): POVExecutionAgent(trade=True) buys 10 % of the
    # transacted volume with MARKET orders every minute from 10:00 to 10:30
    "rmsc03e_s1234_1100": ("rmsc03", ["-t", "ABM", "-d", "20200603", "-s", "1234", "-e", "--end-time", "11:00:00"]),
    # SURVEY 8(f)1 (first half): the reference's rmsc01 population with its MarketMakerAgent (20 quotes a second) and
    # 75 ZI + 24 momentum agents, 09:30-11:00; HBL agents are not on the device yet.  Reference side:
    # tests/golden/refcfg/rmsc01_flags.py
    "rmsc01mm_s33": ("rmsc01", ["-s", "33", "--num-zi", "75", "--num-hbl", "0", "--end-time", "11:00:00"]),
    # the literal rmsc01 population (1 MarketMaker + 50 ZI + 25 HBL + 24 momentum), 09:30-11:00
    "rmsc01_s33_1100": ("rmsc01", ["-s", "33", "--end-time", "11:00:00"]),
    "rmsc01_s34_1100": ("rmsc01", ["-s", "34", "--end-time", "11:00:00"]),
    # BASELINE configs[2] "rmsc01: value + noise agents + market maker" from the supported classes: the rmsc03
    # population at rmsc01 scale (103 agents), one full market day; reference side: tests/golden/refcfg/rmsc_small.py
    "rmsc01s_s77": ("rmsc03", ["-t", "ABM", "-d", "20200603", "-s", "77", "--end-time", "16:00:00",
                               "--num-noise", "50", "--num-value", "25", "--num-momentum", "24"]),
    # more of the reference's stock configs that run on the supported classes (compat/config/ re-states each one;
    # the goldens come from the reference's own config files): legacy latency matrix + noise with Noise / Value
    # agents; full-day 5000 noise + 100 value; and the same plus the polling MarketMakerAgent and 25 momentum agents
    # schedule-driven execution agents (config/execution.py `-e`) in the rmsc03 market at 1/10 scale:
    # a limit order of schedule[minute] shares at the ask every minute 10:00-10:58, the rest as a MARKET order at
    # 10:59.  Reference side: tests/golden/refcfg/execution_flags.py
    "exec_twap_s6": ("execution", ["-t", "ABM", "-d", "20200603", "-s", "6", "-e",
                                   "--num-noise", "500", "--num-value", "50", "--num-momentum", "10"]),
    "exec_vwap_s5": ("execution", ["-t", "ABM", "-d", "20200603", "-s", "5", "-e", "--execution-type", "vwap",
                                   "--num-noise", "500", "--num-value", "50", "--num-momentum", "10"]),
    # SURVEY 8(f)3, market-data subscriptions: config/rmsc02.py (the rmsc01 population with the market maker and the
    # momentum agents on MARKET_DATA subscriptions) over a full day, and 09:30-11:00 with the reference's example
    # SubscriptionAgent added (reference side: tests/golden/refcfg/rmsc02_flags.py)
    "rmsc02_s22": ("rmsc02", ["-s", "22"]),
    "rmsc02sub_s21_1100": ("rmsc02", ["-s", "21", "--end-time", "11:00:00", "--subscription-agent"]),
    # the market maker on 8 levels every 0.1 s: books shallower than the subscription, one-tick-past-the-book quotes
    "rmsc02fast_s23_1100": ("rmsc02", ["-s", "23", "--end-time", "11:00:00", "--mm-subscribe-levels", "8",
                                       "--mm-subscribe-freq", "1e8"]),
    # the reference's config/execution.py itself, full size (5129 agents, TWAP agent trading)
    "execution_s8_e": ("execution", ["-t", "ABM", "-d", "20200603", "-s", "8", "-e"]),
    # config/marketreplay.py -- the reference's MarketReplayAgent replaying an order file inside a running
    # simulation (no oracle, zero latency): limit orders, cancels, quantity modifies, crossing orders, an id that is
    # submitted again after it filled; reference side: tests/golden/refcfg/marketreplay_flags.py
    "marketreplay_s5": ("marketreplay", ["-t", "XYZ", "-d", "20190603", "-s", "5", "--orders-file",
                                         os.path.join(GOLDEN, "replay_orders_file.csv"),
                                         "--processed-orders-folder", os.path.join(GOLDEN, "no_such_cache") + os.sep]),
    "value_noise_s7": ("value_noise", ["-s", "7"]),
    "random_fund_value_s7": ("random_fund_value", ["-s", "7", "-t", "ABM", "-d", "20200603"]),
    "random_fund_diverse_s7": ("random_fund_diverse", ["-s", "7", "-t", "ABM", "-d", "20200603"]),
}


def build_case(name):
    cfg, args = CASES[name]
    return runtime.build_instances(cfg, [args])[0][0]


def check_against_golden(r, g, meta, trace=None):
    assert r["status"] == 0
    assert r["ttl_messages"] == meta["ttl_messages"]
    assert r["delivered"] == meta["delivered"]
    if trace is not None:
        if "trace" in g.files:
            i = first_trace_mismatch(trace, g["trace"])
            assert i < 0, describe_mismatch(trace, g["trace"], i, ("got", "reference"))
        assert hashlib.sha256(np.ascontiguousarray(trace, np.int64).tobytes()).hexdigest() == meta["trace_sha256"]
    for k in ("cash", "holdings", "mtm"):
        assert np.array_equal(r[k], g[k]), k
    fv, gfv = r["final_valuation"], g["final_valuation"]
    assert np.array_equal(np.isnan(fv), np.isnan(gfv))
    # FINAL_VALUATION is a float for Value/Noise agents: north-star tolerance 1e-9 relative (it is in fact exact)
    assert np.allclose(fv[~np.isnan(fv)], gfv[~np.isnan(gfv)], rtol=1e-9, atol=0)
    assert np.array_equal(fv[~np.isnan(fv)], gfv[~np.isnan(gfv)])
    assert r["last_trade"] == meta["last_trade"]
    assert r["final_time"] == meta["final_time"]
    assert r["slowest_agent_finish_time"] == meta["slowest_agent_finish_time"]


@pytest.mark.parametrize("name", sorted(CASES))
def test_builders_reproduce_reference_inputs(name):
    _, meta = load_golden(name)
    assert spec_digest(build_case(name)) == meta["spec_sha256"]


REF_CONFIG = os.path.join(os.environ.get("ABIDES_REFERENCE", "/root/reference"), "config")


@pytest.mark.skipif(not os.path.isdir(REF_CONFIG), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", ["zi100_s123456789", "zi1000_s123456789", "rmsc03_s1234_1000", "rmsc03e_s1234_1100",
                                  "value_noise_s7", "random_fund_value_s7", "random_fund_diverse_s7",
                                  "execution_s8_e", "rmsc02_s22"])
def test_reference_config_files_run_unmodified_on_the_compat_tree(name):
    """The drop-in claim at the config level: the reference's OWN config file, executed against abides_b200/compat
    (its imports resolve to the descriptor classes), lowers to exactly the inputs the reference run received."""
    cfg, args = CASES[name]
    _, meta = load_golden(name)
    ((spec, _),) = runtime.build_instances(os.path.join(REF_CONFIG, cfg + ".py"), [args])
    assert spec_digest(spec) == meta["spec_sha256"]


def test_stored_spec_equals_built_spec():
    assert spec_digest(load_spec("zi100_s123456789")) == spec_digest(build_case("zi100_s123456789"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_run(name):
    g, meta = load_golden(name)
    hb = lowering.HostBatch([build_case(name)])
    r = oracle.run_instance(hb, 0, trace_cap=meta["delivered"] + 16, flog_cap=len(g["f_time"]) + 16)
    check_against_golden(r, g, meta, r["trace"])
    assert np.array_equal(r["f_time"], g["f_time"])
    assert np.allclose(r["f_value"], g["f_value"], rtol=1e-9, atol=0)    # oracle floats (integral after round)
    assert np.array_equal(r["f_value"], g["f_value"])


def test_mean_ending_value_by_type_matches_reference_stdout():
    name = "zi100_s123456789"
    g, meta = load_golden(name)
    spec = build_case(name)
    r = oracle.run_instance(lowering.HostBatch([spec]), 0)
    sums, counts = {}, {}
    start = spec.types[spec.agent_type[1]]["starting_cash"]
    for a in range(1, spec.n_agents):
        t = spec.agent_type_names[a]
        sums[t] = sums.get(t, 0) + int(r["mtm"][a]) - start
        counts[t] = counts.get(t, 0) + 1
    means = {t: int(round(sums[t] / counts[t])) for t in sums}       # Kernel.py:318-322
    assert means == meta["means"]
