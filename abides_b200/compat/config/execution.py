"""execution: the rmsc03 market (exchange + 5000 noise + 100 value + 2 adaptive market makers + 25 momentum,
09:30-11:30) with one schedule-driven execution agent buying 1.2 M shares between 10:00 and 11:00 (reference:
config/execution.py; `-e` makes the agent trade).  Same flags and the same config-time draw order.  Extensions:
--execution-type picks the TWAP (the reference's hard-coded choice) or VWAP agent, --num-* scale the population."""
import argparse
import datetime as dt
import sys

import numpy as np
import pandas as pd
from dateutil.parser import parse

from Kernel import Kernel
from agent.ExchangeAgent import ExchangeAgent
from agent.NoiseAgent import NoiseAgent
from agent.ValueAgent import ValueAgent
from agent.examples.MomentumAgent import MomentumAgent
from agent.execution.TWAPExecutionAgent import TWAPExecutionAgent
from agent.execution.VWAPExecutionAgent import VWAPExecutionAgent
from agent.market_makers.AdaptiveMarketMakerAgent import AdaptiveMarketMakerAgent
from model.LatencyModel import LatencyModel
from util import util
from util.oracle.SparseMeanRevertingOracle import SparseMeanRevertingOracle
from util.order import LimitOrder

ap = argparse.ArgumentParser(description='Detailed options for the execution config.')
ap.add_argument('-c', '--config', required=True)
ap.add_argument('-t', '--ticker', required=True)
ap.add_argument('-d', '--historical-date', required=True, type=parse)
ap.add_argument('-f', '--fundamental-file-path', required=False)
ap.add_argument('-e', '--execution_agents', action='store_true')
ap.add_argument('-s', '--seed', type=int, default=None)
ap.add_argument('-l', '--log_dir', default=None)
ap.add_argument('-v', '--verbose', action='store_true')
ap.add_argument('--config_help', action='store_true')
# extensions (not in the reference)
ap.add_argument('--execution-type', choices=('twap', 'vwap'), default='twap')
ap.add_argument('--num-noise', type=int, default=5000)
ap.add_argument('--num-value', type=int, default=100)
ap.add_argument('--num-momentum', type=int, default=25)
args, _ = ap.parse_known_args()
if args.config_help:
    ap.print_help()
    sys.exit()

seed = args.seed or int(pd.Timestamp.now().timestamp() * 1000000) % (2 ** 32 - 1)
np.random.seed(seed)
util.silent_mode = LimitOrder.silent_mode = not args.verbose
t_begin = dt.datetime.now()
print("Simulation Start Time: {}".format(t_begin))
print("Configuration seed: {}".format(seed))


def fresh_state():
    return np.random.RandomState(seed=np.random.randint(low=0, high=2 ** 32))


day = pd.to_datetime(args.historical_date)
mkt_open, mkt_close = day + pd.to_timedelta('09:30:00'), day + pd.to_timedelta('11:30:00')
sym, cash, r_bar = args.ticker, 10000000, 1e5
symbols = {sym: {'r_bar': r_bar, 'kappa': 1.67e-16, 'sigma_s': 0, 'fund_vol': 1e-8,
                 'megashock_lambda_a': 2.77778e-18, 'megashock_mean': 1e3, 'megashock_var': 5e4,
                 'random_state': fresh_state()}}
oracle = SparseMeanRevertingOracle(mkt_open, mkt_close, symbols)

agents = [ExchangeAgent(id=0, name="ExchangeAgent", type="ExchangeAgent", mkt_open=mkt_open, mkt_close=mkt_close,
                        symbols=[sym], log_orders=True, pipeline_delay=0, computation_delay=0, stream_history=25000,
                        book_freq=0, wide_book=True, random_state=fresh_state())]
noise_open, noise_close = day + pd.to_timedelta("09:00:00"), day + pd.to_timedelta("16:00:00")
for _ in range(args.num_noise):
    j = len(agents)
    when = util.get_wake_time(noise_open, noise_close)           # global draw first, then the agent's seed
    agents.append(NoiseAgent(id=j, name="NoiseAgent_{}".format(j), type="NoiseAgent", symbol=sym, starting_cash=cash,
                             wakeup_time=when, log_orders=False, log_to_file=False, random_state=fresh_state()))
for _ in range(args.num_value):
    j = len(agents)
    agents.append(ValueAgent(id=j, name="ValueAgent_{}".format(j), type="ValueAgent", symbol=sym, starting_cash=cash,
                             sigma_n=r_bar / 10, r_bar=r_bar, kappa=1.67e-15, lambda_a=7e-11, log_orders=False,
                             log_to_file=False, random_state=fresh_state()))
for _ in range(2):
    j = len(agents)
    agents.append(AdaptiveMarketMakerAgent(
        id=j, name="AdaptiveMarketMakerAgent_{}".format(j), type='AdaptiveMarketMakerAgent', symbol=sym,
        starting_cash=cash, pov=0.025, min_order_size=1, window_size='adaptive', num_ticks=10, wake_up_freq='10S',
        cancel_limit_delay=50, skew_beta=0, level_spacing=5, spread_alpha=0.75, backstop_quantity=50000,
        log_orders=True, random_state=fresh_state()))
for _ in range(args.num_momentum):
    j = len(agents)
    agents.append(MomentumAgent(id=j, name="MOMENTUM_AGENT_{}".format(j), type="MomentumAgent", symbol=sym,
                                starting_cash=cash, min_size=1, max_size=10, wake_up_freq='20s', log_orders=True,
                                random_state=fresh_state()))

horizon = pd.date_range(start=day + pd.to_timedelta("10:00:00"), end=day + pd.to_timedelta("11:00:00"), freq='1min')
common = dict(id=len(agents), type='ExecutionAgent', symbol=sym, starting_cash=0, direction="BUY", quantity=12e5,
              execution_time_horizon=horizon, freq='1min', trade=bool(args.execution_agents), log_orders=True)
if args.execution_type == 'twap':
    agents.append(TWAPExecutionAgent(name='TWAPExecutionAgent', random_state=fresh_state(), **common))
else:
    agents.append(VWAPExecutionAgent(name='VWAPExecutionAgent', volume_profile_path=None, random_state=fresh_state(),
                                     **common))

kernel = Kernel("Kernel", random_state=fresh_state())
latency_rstate = fresh_state()
dist = util.exchange_distances_on_line(0.0, 3866660, len(agents), random_state=latency_rstate)
model = LatencyModel(latency_model='deterministic', random_state=latency_rstate,
                     kwargs={'connected': True, 'min_latency': util.meters_to_light_ns(dist)})
kernel.runner(agents=agents, startTime=day, stopTime=mkt_close + pd.to_timedelta('00:01:00'),
              agentLatencyModel=model, defaultComputationDelay=50, oracle=oracle, log_dir=args.log_dir)
print("Simulation End Time: {}".format(dt.datetime.now()))
print("Time taken to run simulation: {}".format(dt.datetime.now() - t_begin))
