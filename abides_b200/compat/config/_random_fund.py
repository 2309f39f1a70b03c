"""Shared body of random_fund_value / random_fund_diverse (reference: config/random_fund_value.py,
config/random_fund_diverse.py): a full 09:30-16:00 day on the sparse mean-reverting fundamental with 5000 noise
and 100 value agents, `diverse` adding one polling MarketMakerAgent and 25 momentum agents.  Same flags, same
order of draws from the global numpy stream."""
import argparse
import datetime as dt
import sys

import numpy as np
import pandas as pd

from Kernel import Kernel
from agent.ExchangeAgent import ExchangeAgent
from agent.NoiseAgent import NoiseAgent
from agent.ValueAgent import ValueAgent
from agent.examples.MomentumAgent import MomentumAgent
from agent.market_makers.MarketMakerAgent import MarketMakerAgent
from model.LatencyModel import LatencyModel
from util import util
from util.oracle.SparseMeanRevertingOracle import SparseMeanRevertingOracle
from util.order import LimitOrder


def run(label, diverse):
    ap = argparse.ArgumentParser(description='Detailed options for {} config.'.format(label))
    ap.add_argument('-c', '--config', required=True)
    ap.add_argument('-t', '--ticker', required=True)
    ap.add_argument('-d', '--historical-date', required=True)
    ap.add_argument('-l', '--log_dir', default=None)
    ap.add_argument('-s', '--seed', type=int, default=None)
    ap.add_argument('-v', '--verbose', action='store_true')
    ap.add_argument('--config_help', action='store_true')
    # extensions (not in the reference): population sizes for tests and grids
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
    print("Configuration seed: {}\n".format(seed))

    def fresh_state(**kw):
        return np.random.RandomState(seed=np.random.randint(low=0, high=2 ** 32, **kw))

    day = pd.to_datetime(args.historical_date)
    mkt_open, mkt_close = day + pd.to_timedelta('09:30:00'), day + pd.to_timedelta('16:00:00')
    sym, cash, r_bar = args.ticker, 10000000, 1e5
    symbols = {sym: {'r_bar': r_bar, 'kappa': 1.67e-12, 'agent_kappa': 1.67e-15, 'sigma_s': 0, 'fund_vol': 1e-8,
                     'megashock_lambda_a': 2.77778e-13, 'megashock_mean': 1e3, 'megashock_var': 5e4,
                     'random_state': fresh_state(dtype='uint64')}}
    oracle = SparseMeanRevertingOracle(mkt_open, mkt_close, symbols)
    agents = [ExchangeAgent(id=0, name="EXCHANGE_AGENT", type="ExchangeAgent", mkt_open=mkt_open,
                            mkt_close=mkt_close, symbols=[sym], log_orders=True, pipeline_delay=0,
                            computation_delay=0, stream_history=10, book_freq=None,
                            random_state=fresh_state(dtype='uint64'))]
    for _ in range(args.num_noise):
        j = len(agents)
        when = util.get_wake_time(mkt_open, mkt_close)             # global draw first, then the agent's seed
        agents.append(NoiseAgent(id=j, name="NoiseAgent {}".format(j), type="NoiseAgent", symbol=sym,
                                 starting_cash=cash, wakeup_time=when, log_orders=False,
                                 random_state=fresh_state(dtype='uint64')))
    for _ in range(args.num_value):
        j = len(agents)
        agents.append(ValueAgent(id=j, name="Value Agent {}".format(j), type="ValueAgent", symbol=sym,
                                 starting_cash=cash, sigma_n=r_bar / 10, r_bar=r_bar, kappa=1.67e-15, lambda_a=1e-12,
                                 random_state=fresh_state(dtype='uint64')))
    if diverse:
        j = len(agents)
        agents.append(MarketMakerAgent(id=j, name="MARKET_MAKER_AGENT_{}".format(j), type='MarketMakerAgent',
                                       symbol=sym, starting_cash=cash, min_size=100, max_size=101,
                                       wake_up_freq="1min", log_orders=False,
                                       random_state=fresh_state(dtype='uint64')))
        for _ in range(args.num_momentum):
            j = len(agents)
            agents.append(MomentumAgent(id=j, name="MOMENTUM_AGENT_{}".format(j), type="MomentumAgent", symbol=sym,
                                        starting_cash=cash, min_size=1, max_size=10, log_orders=False,
                                        random_state=fresh_state(dtype='uint64')))
    kernel = Kernel("{} Kernel".format(label), random_state=fresh_state(dtype='uint64'))
    latency_rstate = fresh_state()
    dist = util.exchange_distances_on_line(0.0, 3866660, len(agents), random_state=latency_rstate)
    model = LatencyModel(latency_model='deterministic', random_state=latency_rstate,
                         kwargs={'connected': True, 'min_latency': util.meters_to_light_ns(dist)})
    kernel.runner(agents=agents, startTime=day, stopTime=mkt_close + pd.to_timedelta('00:01:00'),
                  agentLatencyModel=model, defaultComputationDelay=50, oracle=oracle, log_dir=args.log_dir)
    print("Simulation End Time: {}".format(dt.datetime.now()))
    print("Time taken to run simulation: {}".format(dt.datetime.now() - t_begin))
