"""rmsc03: exchange + 5000 noise + 100 value + 2 adaptive market makers + 25 momentum + 1 POV heartbeat
(reference: config/rmsc03.py).  Same flags and the same config-time draw order; extra flags
--num-noise / --num-value / --num-momentum / --latency-scale parameterise what the reference hard-codes
(used by the parameter-grid workload)."""
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
from agent.execution.POVExecutionAgent import POVExecutionAgent
from agent.market_makers.AdaptiveMarketMakerAgent import AdaptiveMarketMakerAgent
from model.LatencyModel import LatencyModel
from util import util
from util.oracle.SparseMeanRevertingOracle import SparseMeanRevertingOracle
from util.order import LimitOrder

ap = argparse.ArgumentParser(description='Detailed options for RMSC03 config.')
ap.add_argument('-c', '--config', required=True)
ap.add_argument('-t', '--ticker', required=True)
ap.add_argument('-d', '--historical-date', required=True, type=parse)
ap.add_argument('--start-time', default='09:30:00', type=parse)
ap.add_argument('--end-time', default='11:30:00', type=parse)
ap.add_argument('-l', '--log_dir', default=None)
ap.add_argument('-s', '--seed', type=int, default=None)
ap.add_argument('-v', '--verbose', action='store_true')
ap.add_argument('--config_help', action='store_true')
ap.add_argument('-e', '--execution-agents', action='store_true')
ap.add_argument('-p', '--execution-pov', type=float, default=0.1)
ap.add_argument('--mm-pov', type=float, default=0.025)
ap.add_argument('--mm-window-size', type=util.validate_window_size, default='adaptive')
ap.add_argument('--mm-min-order-size', type=int, default=1)
ap.add_argument('--mm-num-ticks', type=int, default=10)
ap.add_argument('--mm-wake-up-freq', type=str, default='10S')
ap.add_argument('--mm-skew-beta', type=float, default=0)
ap.add_argument('--mm-level-spacing', type=float, default=5)
ap.add_argument('--mm-spread-alpha', type=float, default=0.75)
ap.add_argument('--mm-backstop-quantity', type=float, default=50000)
ap.add_argument('--fund-vol', type=float, default=1e-8)
# extensions (not in the reference): population / latency scale for parameter grids
ap.add_argument('--num-noise', type=int, default=5000)
ap.add_argument('--num-value', type=int, default=100)
ap.add_argument('--num-momentum', type=int, default=25)
ap.add_argument('--latency-scale', type=float, default=1.0)
args, _ = ap.parse_known_args()
if args.config_help:
    ap.print_help()
    sys.exit()

seed = args.seed or int(pd.Timestamp.now().timestamp() * 1000000) % (2 ** 32 - 1)
np.random.seed(seed)
util.silent_mode = LimitOrder.silent_mode = not args.verbose
t_begin = dt.datetime.now()
print("Simulation Start Time: {}".format(t_begin))
print("Configuration seed: {}\n".format(seed))


def fresh_state(**kw):
    return np.random.RandomState(seed=np.random.randint(low=0, high=2 ** 32, **kw))


day = pd.to_datetime(args.historical_date)
mkt_open = day + pd.to_timedelta(args.start_time.strftime('%H:%M:%S'))
mkt_close = day + pd.to_timedelta(args.end_time.strftime('%H:%M:%S'))
sym, cash, r_bar = args.ticker, 10000000, 1e5
symbols = {sym: {'r_bar': r_bar, 'kappa': 1.67e-16, 'sigma_s': 0, 'fund_vol': args.fund_vol,
                 'megashock_lambda_a': 2.77778e-18, 'megashock_mean': 1e3, 'megashock_var': 5e4,
                 'random_state': fresh_state(dtype='uint64')}}
oracle = SparseMeanRevertingOracle(mkt_open, mkt_close, symbols)

agents = [ExchangeAgent(id=0, name="EXCHANGE_AGENT", type="ExchangeAgent", mkt_open=mkt_open, mkt_close=mkt_close,
                        symbols=[sym], log_orders=True, pipeline_delay=0, computation_delay=0, stream_history=25000,
                        book_freq=0, wide_book=True, random_state=fresh_state(dtype='uint64'))]
noise_open, noise_close = day + pd.to_timedelta("09:00:00"), day + pd.to_timedelta("16:00:00")
for _ in range(args.num_noise):
    j = len(agents)
    when = util.get_wake_time(noise_open, noise_close)           # global draw first, then the agent's seed
    agents.append(NoiseAgent(id=j, name="NoiseAgent {}".format(j), type="NoiseAgent", symbol=sym, starting_cash=cash,
                             wakeup_time=when, log_orders=None, random_state=fresh_state(dtype='uint64')))
for _ in range(args.num_value):
    j = len(agents)
    agents.append(ValueAgent(id=j, name="Value Agent {}".format(j), type="ValueAgent", symbol=sym, starting_cash=cash,
                             sigma_n=r_bar / 10, r_bar=r_bar, kappa=1.67e-15, lambda_a=7e-11, log_orders=None,
                             random_state=fresh_state(dtype='uint64')))
for _ in range(2):
    j = len(agents)
    agents.append(AdaptiveMarketMakerAgent(
        id=j, name="ADAPTIVE_POV_MARKET_MAKER_AGENT_{}".format(j), type='AdaptivePOVMarketMakerAgent', symbol=sym,
        starting_cash=cash, pov=args.mm_pov, min_order_size=args.mm_min_order_size, window_size=args.mm_window_size,
        num_ticks=args.mm_num_ticks, wake_up_freq=args.mm_wake_up_freq, cancel_limit_delay=50,
        skew_beta=args.mm_skew_beta, level_spacing=args.mm_level_spacing, spread_alpha=args.mm_spread_alpha,
        backstop_quantity=args.mm_backstop_quantity, log_orders=None, random_state=fresh_state(dtype='uint64')))
for _ in range(args.num_momentum):
    j = len(agents)
    agents.append(MomentumAgent(id=j, name="MOMENTUM_AGENT_{}".format(j), type="MomentumAgent", symbol=sym,
                                starting_cash=cash, min_size=1, max_size=10, wake_up_freq='20s', log_orders=None,
                                random_state=fresh_state(dtype='uint64')))
agents.append(POVExecutionAgent(
    id=len(agents), name='POV_EXECUTION_AGENT', type='ExecutionAgent', symbol=sym, starting_cash=cash,
    start_time=mkt_open + pd.to_timedelta('00:30:00'), end_time=mkt_close - pd.to_timedelta('00:30:00'),
    freq='1min', lookback_period='1min', pov=args.execution_pov, direction="BUY", quantity=12e5,
    trade=bool(args.execution_agents), log_orders=True, random_state=fresh_state(dtype='uint64')))

kernel = Kernel("RMSC03 Kernel", random_state=fresh_state(dtype='uint64'))
latency_rstate = fresh_state()
dist = util.exchange_distances_on_line(0.0, 3866660 * args.latency_scale, len(agents), random_state=latency_rstate)
model = LatencyModel(latency_model='deterministic', random_state=latency_rstate,
                     kwargs={'connected': True, 'min_latency': util.meters_to_light_ns(dist)})
kernel.runner(agents=agents, startTime=day, stopTime=mkt_close + pd.to_timedelta('00:01:00'),
              agentLatencyModel=model, defaultComputationDelay=50, oracle=oracle, log_dir=args.log_dir)
print("Simulation End Time: {}".format(dt.datetime.now()))
print("Time taken to run simulation: {}".format(dt.datetime.now() - t_begin))
