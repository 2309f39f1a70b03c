"""Log archives: the exchange's event log, the ORDERBOOK_<symbol>_FULL depth archive and the trading
agents' logs, rebuilt from the trace of one instance (abides_b200/exchange_log.py), against the files the unmodified
reference wrote for the same run (tests/golden/gen_exchange_log_golden.py).  CPU: the trace comes from the C oracle and
from the host build of the lane engine's device code; GPU: from both device engines through `runtime.run_spec_logged`."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

import hostsim
import oracle
from abides_b200 import _abi, engine as eng_mod, lowering, runtime
from gen_exchange_log_golden import canon
from parity_util import GOLDEN

RMSC03 = ["-t", "ABM", "-d", "20200603"]
CASES = {
    "zi100_s123456789": ("sparse_zi_100", ["-s", "123456789"]),
    "rmsc03_s1234_0940": ("rmsc03", RMSC03 + ["-s", "1234", "--end-time", "09:40:00"]),
    # config/marketreplay.py: exchange log with order dicts incl. MODIFY_ORDER rows, order-book archive through
    # modifications and an order id that is used twice (reference side: tests/golden/refcfg/marketreplay_flags.py)
    "marketreplay_s5": ("marketreplay", ["-t", "XYZ", "-d", "20190603", "-s", "5", "--orders-file",
                                         os.path.join(GOLDEN, "replay_orders_file.csv"),
                                         "--processed-orders-folder", os.path.join(GOLDEN, "no_such_cache") + os.sep]),
}
# CPU-only cases (the device runs of the same populations are covered by the parity tests): the literal rmsc01 population
# -- the MarketMakerAgent's five-level BID_DEPTH / ASK_DEPTH rows, HBL, ZI and momentum agents' logs, exchange log without
# order dicts; reference side: tests/golden/refcfg/rmsc01_flags.py with the long-format book archive skipped
HOST_ONLY_CASES = {
    "rmsc01_s33_0935": ("rmsc01", ["-s", "33", "--end-time", "09:35:00"]),
}
NAT = np.iinfo(np.int64).min
TRACE_CAP = 1 << 20


def check_event_log(df, g, k):
    lo, hi = int(g["log_start"][k]), int(g["log_start"][k + 1])
    name = str(g["log_files"][k])
    assert df.index.name == "EventTime" and list(df.columns) == ["EventType", "Event"], name
    t = df.index.to_numpy("datetime64[ns]")
    ti = t.astype(np.int64)
    ti[np.isnat(t)] = NAT
    assert len(df) == hi - lo, (name, len(df), hi - lo)
    assert np.array_equal(ti, g["log_time"][lo:hi]), name
    assert list(df.EventType) == list(g["log_type"][lo:hi]), name
    got = [canon(x) for x in df.Event]
    want = list(g["log_event"][lo:hi])
    bad = [i for i in range(len(got)) if got[i] != want[i]]
    assert not bad, (name, bad[0], got[bad[0]], want[bad[0]])


def check_directory(path, name):
    g = np.load(os.path.join(GOLDEN, name + ".exlog.npz"))
    want_files = [str(f) for f in g["files"]]
    assert sorted(os.listdir(path)) == want_files
    for k, f in enumerate(g["log_files"]):
        check_event_log(pd.read_pickle(os.path.join(path, str(f)), compression="bz2"), g, k)
    if "book_file" in g.files:
        b = pd.read_pickle(os.path.join(path, str(g["book_file"])), compression="bz2")
        assert b.index.name == "QuoteTime"
        assert np.array_equal(b.index.to_numpy("datetime64[ns]").astype(np.int64), g["book_time"])
        assert np.array_equal(np.array([int(q) for q in b.columns], np.int64), g["book_quotes"])
        assert all(str(d) == "Sparse[int64, 0]" for d in b.dtypes)
        dense = np.asarray(b.sparse.to_coo().todense())
        r, c = np.nonzero(dense)
        assert np.array_equal(r, g["book_row"]) and np.array_equal(c, g["book_col"])
        assert np.array_equal(dense[r, c], g["book_val"])


def build(name):
    cfg, args = (CASES.get(name) or HOST_ONLY_CASES[name])
    ((spec, kernel),) = runtime.build_instances(cfg, [list(args) + ["-l", "x"]])
    return spec, kernel


def finish(kernel, hb, res, ares, trace, tmp_path, name):
    out = runtime.RunOutput(hb, res, ares, 0.0, 1.0)
    runtime.write_back(kernel, out, 0)
    path = runtime.write_logs(kernel, trace, root=str(tmp_path))
    check_directory(path, name)


@pytest.mark.parametrize("source", ["oracle", "lane"])
@pytest.mark.parametrize("name", sorted(CASES) + sorted(HOST_ONLY_CASES))
def test_archives_from_host_traces_reproduce_reference_files(tmp_path, name, source):
    if source == "lane" and name.startswith("rmsc01_"):
        pytest.skip("HBL agents run on the warp engine only")
    spec, kernel = build(name)
    hb = lowering.HostBatch([spec], trace_capacity=TRACE_CAP, trace_flags=_abi.TRACE_EXCHANGE_LOG)
    if source == "lane":
        res, ares, trace, n, _ = hostsim.run(hb, trace_instance=0, trace_cap=TRACE_CAP)
    else:
        d = oracle.run_instance(hb, 0, trace_cap=TRACE_CAP, flog_cap=TRACE_CAP)
        n = d["n_trace"]
        fund = np.zeros((max(0, len(d["f_time"]) - 1), 7), np.int64)          # no oracle, no series (marketreplay)
        fund[:, 0], fund[:, 1], fund[:, 2], fund[:, 3] = d["f_time"][1:], -1, _abi.TRACE_FUNDAMENTAL, d["f_value"][1:]
        trace = np.concatenate([d["trace"], fund])
        res = np.zeros(1, eng_mod.INSTANCE_RESULT_DTYPE)
        for f in ("ttl_messages", "last_trade", "slowest_agent_finish_time"):
            res[f][0] = d[f]
        ares = np.zeros(hb.n_agents_total, eng_mod.AGENT_RESULT_DTYPE)
        ares["cash"], ares["holdings"], ares["marked_to_market"] = d["cash"], d["holdings"], d["mtm"]
        ares["final_valuation"] = d["final_valuation"]
    assert n <= TRACE_CAP
    finish(kernel, hb, res, ares, trace, tmp_path, name)


def test_exchange_log_records_leave_the_dispatch_trace_alone():
    """With the flag on, the dispatched events are the ones of a plain traced run, in the same order."""
    spec, _ = build("zi100_s123456789")
    plain = hostsim.run(lowering.HostBatch([spec], trace_capacity=TRACE_CAP), trace_instance=0, trace_cap=TRACE_CAP)[2]
    full = hostsim.run(lowering.HostBatch([spec], trace_capacity=TRACE_CAP, trace_flags=_abi.TRACE_EXCHANGE_LOG),
                       trace_instance=0, trace_cap=TRACE_CAP)[2]
    assert len(full) > len(plain)
    keep = (full[:, 2] < _abi.TRACE_ORDER_PLACED)
    assert np.array_equal(full[keep], plain)


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["warp", "lane"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_archives_from_device_traces_reproduce_reference_files(tmp_path, name, engine):
    if engine == "warp" and name.startswith("marketreplay"):
        pytest.skip("MarketReplayAgent runs on the lane engine only")
    spec, kernel = build(name)
    out, trace = runtime.run_spec_logged(spec, engine_name=engine)
    runtime.write_back(kernel, out, 0)
    path = runtime.write_logs(kernel, trace, root=str(tmp_path))
    check_directory(path, name)


REFERENCE = os.environ.get("ABIDES_REFERENCE", "/root/reference")
CONSUMER = r"""
import os, sys
import pandas as pd, pandas.io.json
pandas.io.json.json_normalize = pd.json_normalize            # compat shim (SURVEY 8c)
sys.path.insert(0, sys.argv[1]); sys.argv = sys.argv[:1]
from util.formatting import convert_order_stream as cos
df = pd.read_pickle(sys.stdin.readline().strip(), compression="bz2").reset_index()
lob = cos.convert_stream_to_format(df, fmt="plot-scripts")
print(len(lob), sorted(lob.TYPE.value_counts().to_dict().items()), int(lob.PRICE.sum()), int(lob.SIZE.sum()), int(lob.ORDER_ID.sum()))
"""


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree only exists in the build container")
def test_reference_consumers_read_the_archives(tmp_path):
    """The reference's own tools run on the files: util/formatting/convert_order_stream.py (:8-78) turns the exchange
    log into the same LOBSTER-style order stream it extracts from the golden's rows, cli/stats.py (:20-82) summarises
    the sparse_zi_100 directory."""
    import subprocess
    spec, kernel = build("rmsc03_s1234_0940")
    hb = lowering.HostBatch([spec], trace_capacity=TRACE_CAP, trace_flags=_abi.TRACE_EXCHANGE_LOG)
    res, ares, trace, n, _ = hostsim.run(hb, trace_instance=0, trace_cap=TRACE_CAP)
    runtime.write_back(kernel, runtime.RunOutput(hb, res, ares, 0.0, 1.0), 0)
    path = runtime.write_logs(kernel, trace, root=str(tmp_path))
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", CONSUMER, REFERENCE], input=os.path.join(path, "EXCHANGE_AGENT.bz2") + "\n",
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    g = np.load(os.path.join(GOLDEN, "rmsc03_s1234_0940.exlog.npz"))
    lo, hi = int(g["log_start"][0]), int(g["log_start"][1])
    types = g["log_type"][lo:hi]
    want = {1: int((types == "LIMIT_ORDER").sum()), 3: int((types == "ORDER_CANCELLED").sum()), 4: int((types == "ORDER_EXECUTED").sum())}
    names = {1: "LIMIT_ORDER", 3: "ORDER_CANCELLED", 4: "ORDER_EXECUTED"}
    assert r.stdout.split("[")[0].strip() == str(sum(want.values()))
    for k, v in want.items():
        assert "('%s', %d)" % (names[k], v) in r.stdout or "(%d, %d)" % (k, v) in r.stdout, r.stdout
    # cli/stats.py on a sparse_zi_100 directory
    spec, kernel = build("zi100_s123456789")
    hb = lowering.HostBatch([spec], trace_capacity=TRACE_CAP, trace_flags=_abi.TRACE_EXCHANGE_LOG)
    res, ares, trace, n, _ = hostsim.run(hb, trace_instance=0, trace_cap=TRACE_CAP)
    runtime.write_back(kernel, runtime.RunOutput(hb, res, ares, 0.0, 1.0), 0)
    path = runtime.write_logs(kernel, trace, root=str(tmp_path / "zi"))
    r = subprocess.run([sys.executable, "-W", "ignore", os.path.join(REFERENCE, "cli", "stats.py"), path], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Agent Mean" in r.stdout and "Read summary files in 1 log directories." in r.stdout


@pytest.mark.parametrize("name", ["rmsc03e_s1234_1100", "rmsc01_s33_1100", "rmsc02fast_s23_1100", "exec_twap_s6"])
def test_order_book_account_agrees_with_top_of_book_on_other_populations(name):
    """Market orders (POV / TWAP execution), the polling and the subscribing MarketMakerAgent's ladders, HBL agents: the
    depth account kept from the exchange's notifications must show the device's own best bid / ask and their volumes
    after every handleLimitOrder (ExchangeArchive raises otherwise); the traded volume of the LAST_TRADE rows is the
    run's volume."""
    from abides_b200 import exchange_log
    from test_oracle_golden import CASES as GOLDEN_CASES
    cfg, args = GOLDEN_CASES[name]
    ((spec, kernel),) = runtime.build_instances(cfg, [list(args) + ["-l", "x"]])
    hb = lowering.HostBatch([spec], trace_capacity=4 * TRACE_CAP, trace_flags=_abi.TRACE_EXCHANGE_LOG)
    d = oracle.run_instance(hb, 0, trace_cap=4 * TRACE_CAP)
    assert d["n_trace"] <= 4 * TRACE_CAP and d["status"] == 0
    kernel.agents[0].book_freq, kernel.agents[0].wide_book, kernel.agents[0].log_orders = 0, True, True
    ares = np.zeros(hb.n_agents_total, eng_mod.AGENT_RESULT_DTYPE)
    ares["cash"], ares["holdings"], ares["marked_to_market"] = d["cash"], d["holdings"], d["mtm"]
    ares["final_valuation"] = d["final_valuation"]
    res = np.zeros(1, eng_mod.INSTANCE_RESULT_DTYPE)
    for f in ("ttl_messages", "last_trade", "slowest_agent_finish_time"):
        res[f][0] = d[f]
    runtime.write_back(kernel, runtime.RunOutput(hb, res, ares, 0.0, 1.0), 0)
    ar = exchange_log.ExchangeArchive(kernel, d["trace"])
    assert ar.n_snapshots == int((d["trace"][:, 2] == _abi.TRACE_BOOK_TOP).sum()) > 0
    traded = sum(int(r[2].split(",")[0]) for r in ar.exchange_rows if r[1] == "LAST_TRADE")
    assert traded == d["traded_volume"]
    executed = [r[2] for r in ar.exchange_rows if r[1] == "ORDER_EXECUTED"]
    assert len(executed) == 2 * d["n_fills"] and sum(e["quantity"] for e in executed) == 2 * d["traded_volume"]
    bf = ar.book_frame()
    assert bf.index.is_monotonic_increasing and bf.index.is_unique


def test_every_log_of_the_execution_config_matches_reference_hashes():
    """config/execution.py `-e` at 1/10 scale, the whole 09:30-11:30 session, every agent logging its own orders
    (tests/golden/exlog_hashes.json, gen_exchange_log_golden.py --hash-only): the exchange's 381 019 rows (MARKET_ORDER rows
    and the notifications of their child limit orders included), the 30 214 x 1 322 order-book archive and the logs of
    the two adaptive market makers (61 k rows each: ORDER_SUBMITTED, CANCEL_SUBMITTED at the send time recovered from
    the deterministic latency, notifications, five-level depth rows), the ten momentum agents and the TWAP agent
    (whole-book depth rows, its execution report) hash to what the reference wrote -- 539 338 rows in 14 files."""
    import json
    from abides_b200 import exchange_log
    from gen_exchange_log_golden import digest
    from test_oracle_golden import CASES as GOLDEN_CASES
    want = json.load(open(os.path.join(GOLDEN, "exlog_hashes.json")))["exec_twap_s6"]
    cfg, args = GOLDEN_CASES["exec_twap_s6"]
    ((spec, kernel),) = runtime.build_instances(cfg, [list(args) + ["-l", "x"]])
    cap = 8 * TRACE_CAP
    hb = lowering.HostBatch([spec], trace_capacity=cap, trace_flags=_abi.TRACE_EXCHANGE_LOG)
    d = oracle.run_instance(hb, 0, trace_cap=cap)
    assert d["n_trace"] <= cap
    ares = np.zeros(hb.n_agents_total, eng_mod.AGENT_RESULT_DTYPE)
    ares["cash"], ares["holdings"], ares["marked_to_market"] = d["cash"], d["holdings"], d["mtm"]
    ares["final_valuation"] = d["final_valuation"]
    res = np.zeros(1, eng_mod.INSTANCE_RESULT_DTYPE)
    for f in ("ttl_messages", "last_trade", "slowest_agent_finish_time"):
        res[f][0] = d[f]
    runtime.write_back(kernel, runtime.RunOutput(hb, res, ares, 0.0, 1.0), 0)
    ar = exchange_log.ExchangeArchive(kernel, d["trace"], spec)
    assert not ar.skipped_agents
    frames = {name.replace(" ", "") + ".bz2": df for name, df in ar.agent_frames().items()}
    frames[kernel.agents[0].name.replace(" ", "") + ".bz2"] = ar.exchange_frame()
    assert sorted(frames) == sorted(want["logs"])
    for name, df in frames.items():
        t = df.index.to_numpy("datetime64[ns]")
        ti = t.astype(np.int64)
        ti[np.isnat(t)] = NAT
        assert digest(ti, list(df.EventType), [canon(x) for x in df.Event]) == want["logs"][name], name
    b = ar.book_frame()
    m = b.sparse.to_coo().tocsr()
    m.sort_indices()
    r, c = m.nonzero()
    got = digest([], [], [], dict(book_time=b.index.to_numpy("datetime64[ns]").astype(np.int64),
                                  book_quotes=[int(q) for q in b.columns], book_row=r, book_col=c,
                                  book_val=np.asarray(m[r, c]).ravel()))
    assert (got["book_shape"], got["book_sha256"]) == (want["book_shape"], want["book_sha256"])
    # without the lowered instance the send times of cancellations are unknown: those agents get no log
    assert len(exchange_log.ExchangeArchive(kernel, d["trace"]).skipped_agents) == 13
