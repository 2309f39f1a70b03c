"""GPU: the lane engine (one instance per thread, abides_b200/csrc/lane_device.inc) through the C ABI against the
reference goldens, the C oracle and the warp engine."""
import hashlib

import numpy as np
import pytest

import oracle
from abides_b200 import _abi, lowering, runtime, engine as eng_mod
from parity_util import (load_golden, device_run, agent_slices, split_device_trace, first_trace_mismatch,
                         describe_mismatch, fuzz_book_rows, random_book_rows)
from test_oracle_golden import CASES, build_case
from test_lane_host import LANE_CASES, LEAN, check_golden
from test_book_replay import golden as book_golden, check as book_check

pytestmark = pytest.mark.gpu
LANE_BASE = 16


@pytest.mark.parametrize("name", LANE_CASES)
def test_lane_engine_matches_reference_golden(name):
    """Full stochastic run on the lane engine, traced: every dispatched event, oracle step and final portfolio."""
    g, meta = load_golden(name)
    cap = meta["delivered"] + len(g["f_time"]) + 64
    hb = lowering.HostBatch([build_case(name)], trace_capacity=cap)
    eng, res, ares = device_run(hb, engine="lane")
    e, v = eng.engine_info()
    assert e == _abi.ENGINE_LANE and v == LANE_BASE + LEAN.get(name, 3)
    tr, n = eng.trace(0)
    assert n <= cap
    check_golden(name, res, ares[agent_slices(hb, 0)], tr, g, meta)


def _sweep(config, seeds, extra=()):
    return [s for s, _ in runtime.build_instances(config, [list(extra) + ["-s", str(s)] for s in seeds])]


def _same_results(res, ares, ores, oares):
    for f in ("status", "ttl_messages", "delivered", "final_time", "last_trade", "final_fundamental", "n_orders", "n_fills",
              "traded_volume", "slowest_agent_finish_time"):
        assert np.array_equal(res[f], ores[f]), f
    for f in ("cash", "holdings", "marked_to_market"):
        assert np.array_equal(ares[f], oares[f]), f
    fv, ofv = ares["final_valuation"], oares["final_valuation"]
    assert np.array_equal(np.isnan(fv), np.isnan(ofv))
    assert np.array_equal(fv[~np.isnan(fv)], ofv[~np.isnan(ofv)])


def _oracle_arrays(hb):
    ores, oares = oracle.run_batch(hb)
    return (np.frombuffer(ores, dtype=eng_mod.INSTANCE_RESULT_DTYPE).copy(),
            np.frombuffer(oares, dtype=eng_mod.AGENT_RESULT_DTYPE).copy())


def test_lane_seed_sweep_matches_oracle_and_warp_engine():
    """70 seeds of sparse_zi_100 (more than two warps of lanes, the last one ragged): lane == oracle == warp engine."""
    hb = lowering.HostBatch(_sweep("sparse_zi_100", range(100, 170)))
    eng, res, ares = device_run(hb, engine="lane")
    assert eng.engine_info() == (_abi.ENGINE_LANE, LANE_BASE + 1)
    ores, oares = _oracle_arrays(hb)
    _same_results(res, ares, ores, oares)
    eng2, wres, wares = device_run(hb, engine="warp")
    assert eng2.engine_info()[0] == _abi.ENGINE_WARP
    _same_results(res, ares, wres, wares)


def test_lane_rmsc03_sweep_matches_oracle():
    """rmsc03 09:30-10:00 x 33 seeds on the lean rmsc variant (transacted-volume log, market-maker ladders, deque of 128)."""
    hb = lowering.HostBatch(_sweep("rmsc03", range(40, 73), ["-t", "ABM", "-d", "20200603", "--end-time", "10:00:00"]))
    eng, res, ares = device_run(hb, engine="lane")
    assert eng.engine_info() == (_abi.ENGINE_LANE, LANE_BASE + 2)
    ores, oares = _oracle_arrays(hb)
    _same_results(res, ares, ores, oares)


def test_lane_pause_and_resume_equals_one_run():
    hb = lowering.HostBatch(_sweep("sparse_zi_100", range(5, 9)))
    eng, res, ares = device_run(hb, engine="lane")
    e2 = eng_mod.Engine(0)
    e2.set_engine("lane")
    e2.load(hb)
    t0 = hb.instances[0].mkt_open
    for cut in (t0 + 3600 * 10 ** 9, t0 + 2 * 3600 * 10 ** 9, t0 + 5 * 3600 * 10 ** 9):
        e2.run(cut)
    e2.run()
    r2, a2 = e2.results_arrays()
    _same_results(res, ares, r2, a2)
    e2.reset()
    e2.run()
    r3, a3 = e2.results_arrays()
    _same_results(res, ares, r3, a3)


def test_lane_book_replay_matches_reference_orderbook():
    g, sh = book_golden()
    e = eng_mod.Engine(0)
    e.set_engine("lane")
    events, counts, states = eng_mod.replay_book(e, g["op_offset"], g["ops"], sh, 64)
    book_check(events, counts, states, g)


def test_lane_deep_book_fuzz_matches_oracle():
    rows, offs = fuzz_book_rows(20261020, n_books=48)
    e = eng_mod.Engine(0)
    e.set_engine("lane")
    events, counts, states = eng_mod.replay_book(e, offs, rows, 25, 96)
    ev2, cnt2, st2 = oracle.replay_book(offs, eng_mod.make_book_ops(rows), 25, 96)
    events_o, counts_o, states_o = eng_mod.unpack_replay(ev2, cnt2[:len(rows)], st2, len(rows), 96)
    assert (counts <= 96).all() and np.array_equal(counts, counts_o)
    assert np.array_equal(states, states_o)
    assert np.array_equal(events, events_o)


def test_lane_many_books_property():
    rows, offs = random_book_rows(7, n_books=512)
    e = eng_mod.Engine(0)
    e.set_engine("lane")
    events, counts, states = eng_mod.replay_book(e, offs, rows, 10, 64)
    ok = (states[:, 0] < 0) | (states[:, 2] < 0) | (states[:, 0] < states[:, 2])
    assert ok.all()
    ex = events[events[:, 1] == 12]
    assert ex[ex[:, 3] == 1][:, 6].sum() == ex[ex[:, 3] == 0][:, 6].sum()
    ev2, cnt2, st2 = oracle.replay_book(offs, eng_mod.make_book_ops(rows), 10, 64)
    events_o, counts_o, states_o = eng_mod.unpack_replay(ev2, cnt2[:len(rows)], st2, len(rows), 64)
    assert np.array_equal(states, states_o) and np.array_equal(events, events_o)


@pytest.mark.parametrize("engine", ["lane", "warp"])
@pytest.mark.parametrize("config, extra, seeds", [
    ("sparse_zi_100", [], list(range(200, 240))), ("sparse_zi_1000", [], [5, 6, 7]),
    ("rmsc03", ["-t", "ABM", "-d", "20200603", "--end-time", "10:00:00"], list(range(40, 45)))])
def test_seeded_sweep_equals_config_built_instances_on_device(config, extra, seeds, engine):
    """Streams handed over as seeds (abides_batch.mt_seed): the device seeds them, fast-forwards the global stream past
    the N x N latency draw, draws the ZI private values and momentum sizes, and -- on the lane engine -- keeps the
    noise agents' streams as seeds.  Same days as the config-built, fully uploaded batch, on both engines."""
    from abides_b200 import sweep
    sb = sweep.build_sweep(config, seeds, extra)
    eng, res, ares = device_run(sb, engine=engine)
    hb = lowering.HostBatch(_sweep(config, seeds, extra))
    ores, oares = _oracle_arrays(hb)
    _same_results(res, ares, ores, oares)


@pytest.mark.parametrize("engine", ["lane", "warp"])
def test_philox_throughput_mode_on_device_matches_oracle(engine):
    """Philox4x32-10 keyed streams: lane engine == warp engine == oracle; and the seeded sweep batch in Philox mode
    (no MT19937 state on the device at all) runs the same days."""
    from abides_b200 import sweep
    seeds = list(range(300, 340))
    hb = lowering.HostBatch(_sweep("sparse_zi_100", seeds), rng_mode=_abi.RNG_PHILOX)
    for i, sd in enumerate(seeds):
        hb.instances[i].seed = sd
    eng, res, ares = device_run(hb, engine=engine)
    ores, oares = _oracle_arrays(hb)
    _same_results(res, ares, ores, oares)
    if engine == "lane":
        sb = sweep.build_sweep("sparse_zi_100", seeds, rng_mode=_abi.RNG_PHILOX)
        assert sb.h2d_bytes() < 100 * sb.n_agents_total
        eng2, res2, ares2 = device_run(sb, engine="lane")
        _same_results(res2, ares2, res, ares)


def test_chunked_sweep_driver_on_device_equals_one_batch():
    """runtime.run_sweep cuts the sweep into chunks handed to the device worker(s) by a ticket counter; per-instance
    results equal those of the same seeds run as ONE batch (and, with two GPUs visible, of the two-device run)."""
    import torch
    seeds = runtime.parallel_seeds(77, 150)
    one = runtime.run_sweep("sparse_zi_100", seeds, devices=(0,))
    chunked = runtime.run_sweep("sparse_zi_100", seeds, devices=(0,), chunk_instances=40)
    assert len(chunked.placement) == 4
    for f in ("ttl_messages", "delivered", "last_trade", "n_fills", "final_time"):
        assert np.array_equal(one.res[f], chunked.res[f]), f
    assert np.array_equal(one.ares["cash"], chunked.ares["cash"])
    assert np.array_equal(one.agents(123)["holdings"], chunked.agents(123)["holdings"])
    if torch.cuda.device_count() >= 2:
        two = runtime.run_sweep("sparse_zi_100", seeds, devices=(0, 1), chunk_instances=25)
        assert {p[0] for p in two.placement} == {0, 1}
        assert np.array_equal(one.res["ttl_messages"], two.res["ttl_messages"]) and np.array_equal(one.ares["cash"], two.ares["cash"])
