"""CPU: the LOGIC of the lane engine (abides_b200/csrc/lane_device.inc: one instance per thread) checked against the
reference goldens and the C oracle through a host build of the same source (tests/hostsim/ -- test infrastructure,
never loaded by the package).  The GPU runs of the same code are in tests/test_gpu_lane.py."""
import hashlib

import numpy as np
import pytest

import oracle
import hostsim
from abides_b200 import lowering, runtime, engine as eng_mod
from parity_util import (load_golden, first_trace_mismatch, describe_mismatch, split_device_trace, fuzz_book_rows,
                         random_book_rows)
from test_oracle_golden import CASES, build_case
from test_book_replay import golden as book_golden, check as book_check

# goldens whose population the lane engine carries (HBL agents and market-data subscriptions stay on the warp engine)
LANE_CASES = ["zi100_s123456789", "zi1000_s123456789", "rmsc03_s1234_1000", "rmsc03e_s1234_1100", "rmsc01s_s77",
              "rmsc01mm_s33", "exec_twap_s6", "exec_vwap_s5", "execution_s8_e", "value_noise_s7",
              "random_fund_value_s7", "random_fund_diverse_s7", "marketreplay_s5"]
LEAN = {"zi100_s123456789": hostsim.LV_ZI_CUBIC, "zi1000_s123456789": hostsim.LV_ZI_LEGACY,
        "rmsc03_s1234_1000": hostsim.LV_RMSC, "rmsc03e_s1234_1100": hostsim.LV_RMSC, "rmsc01s_s77": hostsim.LV_RMSC,
        "random_fund_value_s7": hostsim.LV_RMSC}


def check_golden(name, res, ares, tr, g, meta):
    assert res["status"][0] == 0
    assert res["ttl_messages"][0] == meta["ttl_messages"]
    assert res["delivered"][0] == meta["delivered"]
    disp, fund = split_device_trace(tr)
    if "trace" in g.files:
        i = first_trace_mismatch(disp, g["trace"])
        assert i < 0, describe_mismatch(disp, g["trace"], i, ("lane", "reference"))
    assert hashlib.sha256(np.ascontiguousarray(disp, np.int64).tobytes()).hexdigest() == meta["trace_sha256"]
    assert np.array_equal(fund[:, 0], g["f_time"][1:])
    assert np.array_equal(fund[:, 3].astype(np.float64), g["f_value"][1:])
    assert np.array_equal(ares["cash"], g["cash"])
    assert np.array_equal(ares["holdings"], g["holdings"])
    assert np.array_equal(ares["marked_to_market"], g["mtm"])
    fv, gfv = ares["final_valuation"], g["final_valuation"]
    assert np.array_equal(np.isnan(fv), np.isnan(gfv))
    assert np.array_equal(fv[~np.isnan(fv)], gfv[~np.isnan(gfv)])
    assert res["last_trade"][0] == meta["last_trade"]
    assert res["final_time"][0] == meta["final_time"]
    assert res["slowest_agent_finish_time"][0] == meta["slowest_agent_finish_time"]


@pytest.mark.parametrize("name", LANE_CASES)
def test_lane_logic_matches_reference_golden(name):
    """Every dispatched event, every oracle step and every agent's final portfolio of a full stochastic run, on the
    variant the loader picks AND on the all-classes variant."""
    g, meta = load_golden(name)
    cap = meta["delivered"] + len(g["f_time"]) + 64
    hb = lowering.HostBatch([build_case(name)])
    seen = set()
    for variant in (-1, hostsim.LV_FULL):
        res, ares, tr, n, var = hostsim.run(hb, variant=variant, trace_instance=0, trace_cap=cap)
        assert n <= cap
        if variant == -1:
            assert var == LEAN.get(name, hostsim.LV_FULL)
        if var in seen:
            continue
        seen.add(var)
        check_golden(name, res, ares, tr, g, meta)


def test_lane_untraced_batch_matches_oracle():
    """A ragged batch (different populations side by side, as lanes of one warp would hold them) without tracing."""
    specs = [build_case("zi100_s123456789")] + [s for s, _ in runtime.build_instances(
        "sparse_zi_100", [["-s", str(s)] for s in (3, 4, 5)])]
    hb = lowering.HostBatch(specs)
    res, ares, _, _, var = hostsim.run(hb)
    assert var == hostsim.LV_ZI_CUBIC
    ores, oares = oracle.run_batch(hb)
    o = np.frombuffer(oares, dtype=ares.dtype)
    for i in range(hb.n_instances):
        for f in ("ttl_messages", "delivered", "final_time", "last_trade", "final_fundamental", "n_orders", "n_fills",
                  "traded_volume", "slowest_agent_finish_time", "status"):
            assert res[f][i] == getattr(ores[i], f), (i, f)
    for f in ("cash", "holdings", "marked_to_market"):
        assert np.array_equal(ares[f], o[f])


def _unpack(ev, cnt, st, n, k):
    return eng_mod.unpack_replay(ev, cnt[:n], st, n, k)


def test_lane_book_replay_matches_reference_orderbook():
    g, sh = book_golden()
    ops = eng_mod.make_book_ops(g["ops"])
    n = int(g["op_offset"][-1])
    ev, cnt, st = hostsim.replay_book(g["op_offset"], ops, sh, 64)
    book_check(*_unpack(ev, cnt, st, n, 64), g)


@pytest.mark.parametrize("seed", [20261020, 7])
def test_lane_deep_book_fuzz_matches_oracle(seed):
    """Deep, wide books with every message type: the deque's both-ends inserts and removes, spills to the pool, slot
    reuse, refills and level sums that run into the pool, the modify index-0 overwrite, sweeping market orders."""
    rows, offs = fuzz_book_rows(seed, n_books=24)
    ops = eng_mod.make_book_ops(rows)
    n = len(rows)
    a = _unpack(*hostsim.replay_book(offs, ops, 25, 96), n, 96)
    b = _unpack(*oracle.replay_book(offs, ops, 25, 96), n, 96)
    assert (a[1] <= 96).all() and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2], b[2])
    assert np.array_equal(a[0], b[0])
    assert a[2][:, 5].max() > 100 and a[2][:, 6].max() > 100


def test_lane_shallow_books_match_oracle():
    rows, offs = random_book_rows(7, n_books=64)
    ops = eng_mod.make_book_ops(rows)
    n = len(rows)
    a = _unpack(*hostsim.replay_book(offs, ops, 10, 64), n, 64)
    b = _unpack(*oracle.replay_book(offs, ops, 10, 64), n, 64)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


SWEEPS = [("sparse_zi_100", [], [11, 12, 13]), ("sparse_zi_1000", [], [5, 6]),
          ("rmsc03", ["-t", "ABM", "-d", "20200603", "--end-time", "10:00:00"], [40, 41, 42])]


@pytest.mark.parametrize("config, extra, seeds", SWEEPS)
def test_seeded_sweep_equals_config_built_instances(config, extra, seeds):
    """sweep.build_sweep (vectorised config-time draws, streams handed over as seeds, ZI private values / momentum
    sizes / noise agents' few words drawn by the engine itself) runs exactly the days that executing the config
    module once per seed and uploading every MT19937 state does."""
    from abides_b200 import sweep
    sb = sweep.build_sweep(config, seeds, extra)
    assert sb.h2d_bytes() < 80 * sb.n_agents_total + 4096                   # no 2.5 KB states travel
    res, ares, _, _, var = hostsim.run(sb)
    specs = [s for s, _ in runtime.build_instances(config, [list(extra) + ["-s", str(s)] for s in seeds])]
    hb = lowering.HostBatch(specs)
    res2, ares2, _, _, var2 = hostsim.run(hb)
    assert var == var2 and (res["status"] == 0).all()
    for f in res.dtype.names:
        if f != "reserved":
            assert np.array_equal(res[f], res2[f]), f
    for f in ("cash", "holdings", "marked_to_market"):
        assert np.array_equal(ares[f], ares2[f]), f
    # and the oracle agrees on the config-built batch
    ores, oares = oracle.run_batch(hb)
    assert [int(ores[i].ttl_messages) for i in range(len(seeds))] == [int(x) for x in res["ttl_messages"]]


def test_vec_random_state_is_numpy_legacy_random_state():
    from abides_b200.vecrng import VecRandomState
    seeds = np.random.RandomState(5).randint(0, 2 ** 32, size=32, dtype='uint64')
    v = VecRandomState(seeds)
    refs = [np.random.RandomState(seed=s) for s in seeds]
    for it in range(160):
        assert np.array_equal(v.randint(0, 2 ** 32), [r.randint(0, 2 ** 32, dtype='uint64') for r in refs])
        assert np.array_equal(v.rand(), [r.rand() for r in refs])
        assert np.array_equal(v.randint(0, 100), [r.randint(0, 100) for r in refs])
        assert np.array_equal(v.randint(20, 50), [r.randint(20, 50) for r in refs])
        assert np.array_equal(v.uniform(21000, 13000000), [r.uniform(21000, 13000000) for r in refs])
        assert np.array_equal(v.normal(1e3, np.sqrt(5e4)), [r.normal(1e3, np.sqrt(5e4)) for r in refs])
        assert np.array_equal(v.exponential(3.6e12), [r.exponential(3.6e12) for r in refs])
    for i in (0, 31):
        st, mine = refs[i].get_state(), v.numpy_state(i)
        assert np.array_equal(st[1], mine[1]) and tuple(st[2:]) == tuple(mine[2:])


@pytest.mark.parametrize("config, extra, seeds", [SWEEPS[0], SWEEPS[2]])
def test_lane_philox_streams_match_oracle(config, extra, seeds):
    """Throughput mode: counter-based Philox4x32-10 streams keyed by (instance seed, stream).  The lane engine, the
    oracle and the seeded sweep batch (constructor draws still from RandomState(seed), by throw-away generators)
    implement the same keyed streams, so their runs are bit-identical to each other."""
    from abides_b200 import sweep, _abi
    specs = [s for s, _ in runtime.build_instances(config, [list(extra) + ["-s", str(s)] for s in seeds])]
    hb = lowering.HostBatch(specs, rng_mode=_abi.RNG_PHILOX)
    for i, sd in enumerate(seeds):
        hb.instances[i].seed = sd                                # the sweep builder keys each instance by its config seed
    res, ares, _, _, _ = hostsim.run(hb)
    ores, oares = oracle.run_batch(hb)
    o = np.frombuffer(oares, dtype=ares.dtype)
    assert [int(ores[i].ttl_messages) for i in range(len(seeds))] == [int(x) for x in res["ttl_messages"]]
    assert (res["status"] == 0).all()
    for f in ("cash", "holdings", "marked_to_market"):
        assert np.array_equal(ares[f], o[f]), f
    sb = sweep.build_sweep(config, seeds, extra, rng_mode=_abi.RNG_PHILOX)
    res2, ares2, _, _, _ = hostsim.run(sb)
    assert np.array_equal(res2["ttl_messages"], res["ttl_messages"]) and np.array_equal(ares2["cash"], ares["cash"])
    assert len(set(int(x) for x in res["ttl_messages"])) > 1


@pytest.mark.parametrize("caps, name", [
    ("150,0,0,0,0,0", "zi100_s123456789"),        # event heap: 2 n - 1 events at the opening handshake do not fit in 150
    ("0,60,0,0,0,0", "zi100_s123456789"),         # message payload slots
    ("0,0,40,0,0,0", "zi100_s123456789"),         # open-order arena
    ("0,0,0,8,0,0", "zi100_s123456789"),          # book pool: a ZI book rests far more than 32 + 8 orders per side
    ("0,0,0,0,16,0", "rmsc03_s1234_1000"),        # transaction rows between two volume queries
    ("0,0,0,0,0,16", "rmsc03_s1234_1000"),        # unrolled rows inside the look-back window
])
def test_lane_capacity_overflow_stops_the_instance_loudly(caps, name, monkeypatch):
    """Every fixed capacity of the lane engine is checked where it is consumed: an instance that outgrows one ends with
    ABIDES_EOVERFLOW (never a silent wrap or a write past the end), and its neighbours in the batch are unaffected."""
    g, meta = load_golden(name)
    spec = build_case(name)
    hb = lowering.HostBatch([spec])
    monkeypatch.setenv("HOSTSIM_CAPS", caps)
    res, ares, _, _, _ = hostsim.run(hb)
    assert res["status"][0] == _abi_mod().EOVERFLOW
    assert res["ttl_messages"][0] < meta["ttl_messages"]
    monkeypatch.delenv("HOSTSIM_CAPS")
    res, ares, _, _, _ = hostsim.run(hb)
    assert res["status"][0] == 0 and res["ttl_messages"][0] == meta["ttl_messages"]


def _abi_mod():
    from abides_b200 import _abi
    return _abi


def test_lane_event_beyond_the_key_range_is_an_overflow():
    """Event times are stored as 48-bit offsets from the kernel start (78 hours): a wakeup requested past that ends the
    instance with ABIDES_EOVERFLOW instead of wrapping into the past."""
    spec = build_case("rmsc03_s1234_1000")
    spec.wakeup_time = spec.wakeup_time.copy()
    noise = [i for i, t in enumerate(spec.agent_type) if spec.types[t]["klass"] == _abi_mod().NOISE]
    spec.wakeup_time[noise[0]] = spec.cfg["start_time"] + 100 * 3600 * 10 ** 9
    res, ares, _, _, _ = hostsim.run(lowering.HostBatch([spec]))
    assert res["status"][0] == _abi_mod().EOVERFLOW


@pytest.mark.parametrize("defines", ["LANE_STAGE=1", "LANE_STAGE=0 LANE_ONE_CTX=1"])
def test_lane_build_options_do_not_change_results(defines):
    """The lane engine's build options (csrc/lane_types.h: thread-local staging of the hot state, one context per
    thread) are performance choices only: the host build of each setting reproduces the reference goldens event for
    event (a subprocess, because the harness loads one build per process)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, ABIDES_HOSTSIM_DEFINES=defines)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", os.path.abspath(__file__), "-k",
                        "test_lane_logic_matches_reference_golden and (zi100 or rmsc03_s1234 or marketreplay or exec_twap)"],
                       env=env, capture_output=True, text=True, timeout=1200, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "5 passed" in r.stdout and "failed" not in r.stdout, r.stdout[-1500:] + r.stderr[-500:]
