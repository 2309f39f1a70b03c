"""Golden-vector generator input (runs on the UNMODIFIED reference only): the rmsc03 population with the agent
counts as flags, written against the reference's own API.  The reference hard-codes 5000 noise / 100 value / 25
momentum agents (config/rmsc03.py:199,:216,:275); BASELINE.json's "rmsc01: value + noise agents + market maker"
workload is this population at rmsc01 scale (about 100 agents), built from the supported agent classes.  The draw
order (np.random seeds, wake times, latency positions) is that of config/rmsc03.py, so that
`abides_b200/compat/config/rmsc03.py --num-noise .. --num-value .. --num-momentum ..` lowers to the same inputs
(checked through spec_sha256)."""
import argparse
import sys

import numpy as np
import pandas as pd
from dateutil.parser import parse

from Kernel import Kernel
from agent.ExchangeAgent import ExchangeAgent
from agent.NoiseAgent import NoiseAgent
from agent.ValueAgent import ValueAgent
from agent.examples.MomentumAgent import MomentumAgent
from agent.execution.POVExecutionAgent import POVExecutionAgent
from agent.market_makers.AdaptiveMarketMakerAgent import AdaptiveMarketMakerAgent
from model.LatencyModel import LatencyModel
from util import util
from util.oracle.SparseMeanRevertingOracle import SparseMeanRevertingOracle
from util.order import LimitOrder

ap = argparse.ArgumentParser()
ap.add_argument('-c', '--config', required=True)
ap.add_argument('-t', '--ticker', required=True)
ap.add_argument('-d', '--historical-date', required=True, type=parse)
ap.add_argument('--start-time', default='09:30:00', type=parse)
ap.add_argument('--end-time', default='11:30:00', type=parse)
ap.add_argument('-s', '--seed', type=int, default=None)
ap.add_argument('--num-noise', type=int, default=5000)
ap.add_argument('--num-value', type=int, default=100)
ap.add_argument('--num-momentum', type=int, default=25)
args, _ = ap.parse_known_args()

np.random.seed(args.seed)
util.silent_mode = True
LimitOrder.silent_mode = True


def rs(**kw):
    return np.random.RandomState(seed=np.random.randint(low=0, high=2 ** 32, **kw))


day = pd.to_datetime(args.historical_date)
mkt_open = day + pd.to_timedelta(args.start_time.strftime('%H:%M:%S'))
mkt_close = day + pd.to_timedelta(args.end_time.strftime('%H:%M:%S'))
sym, cash, r_bar = args.ticker, 10000000, 1e5
symbols = {sym: {'r_bar': r_bar, 'kappa': 1.67e-16, 'sigma_s': 0, 'fund_vol': 1e-8, 'megashock_lambda_a': 2.77778e-18,
                 'megashock_mean': 1e3, 'megashock_var': 5e4, 'random_state': rs(dtype='uint64')}}
oracle = SparseMeanRevertingOracle(mkt_open, mkt_close, symbols)
agents = [ExchangeAgent(id=0, name="EXCHANGE_AGENT", type="ExchangeAgent", mkt_open=mkt_open, mkt_close=mkt_close,
                        symbols=[sym], log_orders=True, pipeline_delay=0, computation_delay=0, stream_history=25000,
                        book_freq=0, wide_book=True, random_state=rs(dtype='uint64'))]
n_open, n_close = day + pd.to_timedelta("09:00:00"), day + pd.to_timedelta("16:00:00")
for _ in range(args.num_noise):
    j = len(agents)
    agents.append(NoiseAgent(id=j, name="NoiseAgent {}".format(j), type="NoiseAgent", symbol=sym, starting_cash=cash,
                             wakeup_time=util.get_wake_time(n_open, n_close), log_orders=None,
                             random_state=rs(dtype='uint64')))
for _ in range(args.num_value):
    j = len(agents)
    agents.append(ValueAgent(id=j, name="Value Agent {}".format(j), type="ValueAgent", symbol=sym, starting_cash=cash,
                             sigma_n=r_bar / 10, r_bar=r_bar, kappa=1.67e-15, lambda_a=7e-11, log_orders=None,
                             random_state=rs(dtype='uint64')))
for _ in range(2):
    j = len(agents)
    agents.append(AdaptiveMarketMakerAgent(
        id=j, name="ADAPTIVE_POV_MARKET_MAKER_AGENT_{}".format(j), type='AdaptivePOVMarketMakerAgent', symbol=sym,
        starting_cash=cash, pov=0.025, min_order_size=1, window_size='adaptive', num_ticks=10, wake_up_freq='10S',
        cancel_limit_delay=50, skew_beta=0, level_spacing=5, spread_alpha=0.75, backstop_quantity=50000,
        log_orders=None, random_state=rs(dtype='uint64')))
for _ in range(args.num_momentum):
    j = len(agents)
    agents.append(MomentumAgent(id=j, name="MOMENTUM_AGENT_{}".format(j), type="MomentumAgent", symbol=sym,
                                starting_cash=cash, min_size=1, max_size=10, wake_up_freq='20s', log_orders=None,
                                random_state=rs(dtype='uint64')))
agents.append(POVExecutionAgent(
    id=len(agents), name='POV_EXECUTION_AGENT', type='ExecutionAgent', symbol=sym, starting_cash=cash,
    start_time=mkt_open + pd.to_timedelta('00:30:00'), end_time=mkt_close - pd.to_timedelta('00:30:00'),
    freq='1min', lookback_period='1min', pov=0.1, direction="BUY", quantity=12e5, trade=False, log_orders=True,
    random_state=rs(dtype='uint64')))
kernel = Kernel("RMSC small Kernel", random_state=rs(dtype='uint64'))
lat_rs = rs()
lat = util.meters_to_light_ns(util.generate_uniform_random_pairwise_dist_on_line(0.0, 3866660, len(agents),
                                                                                 random_state=lat_rs))
model = LatencyModel(latency_model='deterministic', random_state=lat_rs, kwargs={'connected': True, 'min_latency': lat})
kernel.runner(agents=agents, startTime=day, stopTime=mkt_close + pd.to_timedelta('00:01:00'), agentLatencyModel=model,
              defaultComputationDelay=50, oracle=oracle, log_dir=None)
