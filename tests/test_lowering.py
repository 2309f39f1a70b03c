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
