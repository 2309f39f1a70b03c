"""Host logic: what the lowering accepts and what it refuses loudly (nothing silently runs something else)."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

from abides_b200 import _abi, lowering, runtime

sys.path.insert(0, runtime.COMPAT_DIR)
from agent.examples.MomentumAgent import MomentumAgent                      # noqa: E402
from agent.examples.SubscriptionAgent import SubscriptionAgent              # noqa: E402
from agent.execution.TWAPExecutionAgent import TWAPExecutionAgent           # noqa: E402
from agent.execution.VWAPExecutionAgent import VWAPExecutionAgent           # noqa: E402
from agent.market_makers.MarketMakerAgent import MarketMakerAgent           # noqa: E402
from agent.TradingAgent import TradingAgent                                 # noqa: E402
from util.oracle.ExternalFileOracle import ExternalFileOracle               # noqa: E402

DAY = pd.Timestamp("2020-06-03")


def rs():
    return np.random.RandomState(1)


def horizon(freq="1min", end="10:10:00"):
    return pd.date_range(DAY + pd.to_timedelta("10:00:00"), DAY + pd.to_timedelta(end), freq=freq)


def twap(**kw):
    args = dict(id=1, name="t", type="ExecutionAgent", symbol="ABM", starting_cash=0, direction="BUY", quantity=1100,
                execution_time_horizon=horizon(), freq="1min", trade=True, random_state=rs())
    args.update(kw)
    return TWAPExecutionAgent(**args)


def test_twap_schedule_rows_and_type_row():
    a = twap()
    assert lowering._schedule_rows(a) == [100] * 9                  # int(1100 / 11) in each of the 11 - 2 limit-order slots
    row = lowering._type_row(a, _abi.SCHEDULE_EXEC)
    assert row["wake_up_freq"] == 60 * 10 ** 9 and row["exec_trade"] == 1 and row["exec_is_buy"] == 1
    assert row["exec_end_time"] - row["exec_start_time"] == 10 * 60 * 10 ** 9


def test_vwap_synthetic_profile_is_the_reference_formula():
    a = VWAPExecutionAgent(id=1, name="v", type="ExecutionAgent", symbol="ABM", starting_cash=0, direction="SELL",
                           quantity=12e5, execution_time_horizon=horizon(), freq="1min", volume_profile_path=None,
                           trade=True, random_state=rs())
    grid = pd.date_range(DAY + pd.to_timedelta("09:30:00"), DAY + pd.to_timedelta("16:00:00"), freq="1min")
    xs = np.arange(int(-len(grid) / 2), int(len(grid) / 2))
    w = xs ** 2 + 2 * xs + 2
    want = [round(float(w[list(grid).index(t)]) * (1.0 / w.sum()) * 12e5) for t in horizon()[:-2]]
    assert lowering._schedule_rows(a) == want
    assert lowering._type_row(a, _abi.SCHEDULE_EXEC)["exec_is_buy"] == 0


@pytest.mark.parametrize("bad, why", [
    (dict(execution_time_horizon=horizon("30s")), "no bin"),                 # the reference raises KeyError there
    (dict(execution_time_horizon=horizon()[[0, 1, 3, 4]]), "regular grid"),
    (dict(execution_time_horizon=horizon()[:2]), "three points"),
    (dict(quantity=1100.5), "non-integral"),
])
def test_execution_agents_the_reference_cannot_run_are_refused(bad, why):
    a = twap(**bad)
    with pytest.raises(NotImplementedError, match=why):
        lowering._schedule_rows(a)
        lowering._type_row(a, _abi.SCHEDULE_EXEC)


def test_subscription_rows():
    mm = MarketMakerAgent(1, "m", "MarketMakerAgent", "ABM", 10 ** 7, 500, 1000, subscribe=True, subscribe_freq=1e8,
                          subscribe_num_levels=3, random_state=rs())
    row = lowering._type_row(mm, _abi.MARKET_MAKER)
    assert (row["subscribe"], row["sub_levels"], row["subscribe_freq"]) == (1, 3, 1e8)
    mo = MomentumAgent(2, "mo", "MomentumAgent", "ABM", 10 ** 7, 1, 10, subscribe=True, random_state=rs())
    row = lowering._type_row(mo, _abi.MOMENTUM)
    assert (row["subscribe"], row["sub_levels"], row["subscribe_freq"]) == (1, 1, 10e9)      # MomentumAgent.py:37
    polling = lowering._type_row(MomentumAgent(3, "p", "MomentumAgent", "ABM", 10 ** 7, 1, 10, random_state=rs()),
                                 _abi.MOMENTUM)
    assert polling["subscribe"] == 0 and polling["sub_levels"] == 0
    sub = SubscriptionAgent(4, "s", "SubscriptionAgent", "ABM", 10 ** 7, levels=5, freq=10e9, random_state=rs())
    assert lowering.agent_class(sub) == _abi.SUBSCRIPTION
    with pytest.raises(NotImplementedError, match="levels"):
        lowering._type_row(SubscriptionAgent(5, "s", "SubscriptionAgent", "ABM", 10 ** 7, levels=_abi.MD_MAX_LEVELS + 1,
                                             freq=0, random_state=rs()), _abi.SUBSCRIPTION)


def test_unsupported_agents_and_oracles_are_refused():
    class ImpactAgent(TradingAgent):
        pass

    class MyMomentum(MomentumAgent):                                # a subclass may override the strategy
        pass

    with pytest.raises(NotImplementedError, match="ImpactAgent"):
        lowering.agent_class(ImpactAgent(1, "i", "ImpactAgent", random_state=rs()))
    with pytest.raises(NotImplementedError, match="subclass of MomentumAgent"):
        lowering.agent_class(MyMomentum(1, "m", "MomentumAgent", "ABM", 10 ** 7, 1, 10, random_state=rs()))
    with pytest.raises(NotImplementedError, match="ExternalFileOracle"):
        ExternalFileOracle({"ABM": {}})


def test_message_codes_match_the_header():
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "abides_b200.h")).read()
    for code, name in _abi.MSG_NAMES.items():
        assert "ABIDES_MSG_%s = %d" % (name, code) in text


def _compat_replay_agent(tmp_path, text, name="orders.txt", **kw):
    """compat MarketReplayAgent on an order file written to tmp_path (the compat tree on sys.path as a config would have it)."""
    import importlib
    import sys
    import numpy as np
    import pandas as pd
    from abides_b200 import runtime
    path = tmp_path / name
    path.write_text(text)
    sys.path.insert(0, runtime.COMPAT_DIR)
    try:
        for m in [k for k in sys.modules if k == "agent" or k.startswith("agent.")]:
            del sys.modules[m]
        mod = importlib.import_module("agent.examples.MarketReplayAgent")
    finally:
        sys.path.remove(runtime.COMPAT_DIR)
    day = pd.Timestamp("2019-06-03")
    args = dict(id=1, name="R", type="MarketReplayAgent", symbol="XYZ", date=day, start_time=day + pd.Timedelta("09:30:00"),
                end_time=day + pd.Timedelta("16:00:00"), orders_file_path=str(path),
                processed_orders_folder_path=str(tmp_path) + "/cache_", starting_cash=0,
                random_state=np.random.RandomState(1))
    args.update(kw)
    return mod.MarketReplayAgent(**args), day


def test_market_replay_agent_reads_pipe_and_comma_files(tmp_path):
    """agent/examples/MarketReplayAgent.py:72-125: '|' separated L3 rows with 0 / 1 direction flags, dollars -> integer
    cents by truncation (:113-114), rows outside [start_time, end_time) dropped, timestamps grouped in ascending order."""
    import pandas as pd
    text = ("TIMESTAMP|ORDER_ID|PRICE|SIZE|BUY_SELL_FLAG\n"
            "20190603092959.000000000|7|100.00|10|0\n"           # before start_time
            "20190603093000.500000000|1|100.29|100|0\n"
            "20190603093000.500000000|2|101.10|50|1\n"
            "20190603093001.000000001|1|100.29|0|0\n"
            "20190603160000.000000000|9|100.00|10|1\n")          # at end_time: excluded
    a, day = _compat_replay_agent(tmp_path, text)
    od = a.historical_orders.orders_dict
    t0 = day + pd.Timedelta("09:30:00.5")
    assert list(od) == [t0, day + pd.Timedelta("09:30:01") + pd.Timedelta(1, "ns")] == a.wakeup_times
    assert a.historical_orders.first_wakeup == t0
    assert od[t0] == [{'ORDER_ID': 1, 'PRICE': int(100.29 * 100), 'SIZE': 100, 'BUY_SELL_FLAG': 'BUY'},
                      {'ORDER_ID': 2, 'PRICE': int(101.10 * 100), 'SIZE': 50, 'BUY_SELL_FLAG': 'SELL'}]
    b, _ = _compat_replay_agent(tmp_path, text.replace("|", ","), name="orders.csv")
    assert b.historical_orders.orders_dict == od


def test_market_replay_agent_prefers_the_processed_orders_pickle(tmp_path):
    """:89-92: `<folder>marketreplay_<symbol>_<date>.pkl`, when present, is used instead of the order file."""
    import pickle
    import pandas as pd
    day = pd.Timestamp("2019-06-03")
    cached = {day + pd.Timedelta("10:00:00"): [{'ORDER_ID': 5, 'PRICE': 12345, 'SIZE': 7, 'BUY_SELL_FLAG': 'SELL'}]}
    with open(str(tmp_path) + "/cache_marketreplay_XYZ_2019-06-03.pkl", "wb") as h:
        pickle.dump(cached, h)
    a, _ = _compat_replay_agent(tmp_path, "TIMESTAMP,ORDER_ID,PRICE,SIZE,BUY_SELL_FLAG\n")
    assert a.historical_orders.orders_dict == cached


def test_replay_tape_lowering_and_its_refusals(tmp_path):
    """The tape as the device reads it (six ints per row, 64-bit time split), the first wake-up in the type row, and the
    inputs the lowering refuses: an empty tape (the reference raises IndexError, :84) and an order before the open."""
    import numpy as np
    import pandas as pd
    import pytest
    from abides_b200 import lowering
    text = ("TIMESTAMP,ORDER_ID,PRICE,SIZE,BUY_SELL_FLAG\n20190603093000.000000005,11,100.00,10,BUY\n"
            "20190603093000.000000005,12,100.50,20,SELL\n20190603100000.000000000,11,100.00,0,BUY\n")
    a, day = _compat_replay_agent(tmp_path, text)
    rows, first = lowering._replay_rows(a, (day + pd.Timedelta("09:30:00")).value)
    r = np.array(rows, np.int64).reshape(-1, 6)
    t = ((r[:, 1] << 32) | (r[:, 0] & 0xFFFFFFFF))
    assert first == t[0] == (day + pd.Timedelta("09:30:00")).value + 5 and t[2] == (day + pd.Timedelta("10:00:00")).value
    assert r[:, 2:].tolist() == [[11, 10000, 10, 1], [12, 10050, 20, 0], [11, 10000, 0, 1]]
    assert all(-2 ** 31 <= x < 2 ** 31 for x in rows)
    with pytest.raises(NotImplementedError):
        lowering._replay_rows(a, (day + pd.Timedelta("09:30:01")).value)          # first order before the open
    with pytest.raises(IndexError):
        _compat_replay_agent(tmp_path, "TIMESTAMP,ORDER_ID,PRICE,SIZE,BUY_SELL_FLAG\n", name="empty.csv")
