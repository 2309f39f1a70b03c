"""Golden-vector generator input (runs on the UNMODIFIED reference only): config/rmsc02.py with the population sizes,
the market close and the (in the reference commented-out) example SubscriptionAgent as flags, written against the
reference's own API.  The draw order is that of config/rmsc02.py, so that `abides_b200/compat/config/rmsc02.py` with
the same flags lowers to the same inputs (checked through spec_sha256)."""
import argparse
import datetime as dt
import sys

import numpy as np
import pandas as pd

from Kernel import Kernel
from agent.ExchangeAgent import ExchangeAgent
from agent.HeuristicBeliefLearningAgent import HeuristicBeliefLearningAgent
from agent.ZeroIntelligenceAgent import ZeroIntelligenceAgent
from agent.examples.MomentumAgent import MomentumAgent
from agent.examples.SubscriptionAgent import SubscriptionAgent
from agent.market_makers.MarketMakerAgent import MarketMakerAgent
from model.LatencyModel import LatencyModel
from util import util
from util.oracle.SparseMeanRevertingOracle import SparseMeanRevertingOracle
from util.order import LimitOrder

ap = argparse.ArgumentParser(description='Detailed options for RMSC-2 (Reference Market Simulation Configuration) config.')
ap.add_argument('-c', '--config', required=True)
ap.add_argument('-l', '--log_dir', default=None)
ap.add_argument('-s', '--seed', type=int, default=None)
ap.add_argument('-v', '--verbose', action='store_true')
ap.add_argument('--config_help', action='store_true')
ap.add_argument('-a', '--agent_name', default=None)
# extensions (not in the reference)
ap.add_argument('--num-zi', type=int, default=50)
ap.add_argument('--num-hbl', type=int, default=25)
ap.add_argument('--num-momentum', type=int, default=24)
ap.add_argument('--end-time', default='16:00:00')
ap.add_argument('--subscription-agent', action='store_true')
ap.add_argument('--mm-subscribe-levels', type=int, default=5)
ap.add_argument('--mm-subscribe-freq', type=float, default=10e9)
args, _ = ap.parse_known_args()
if args.config_help:
    ap.print_help()
    sys.exit()
if args.agent_name:
    raise NotImplementedError("--agent_name loads an arbitrary experimental agent class: not on the B200 path")

seed = args.seed or int(pd.Timestamp.now().timestamp() * 1000000) % (2 ** 32 - 1)
np.random.seed(seed)
util.silent_mode = LimitOrder.silent_mode = not args.verbose
t_begin = dt.datetime.now()
print("Simulation Start Time: {}".format(t_begin))
print("Configuration seed: {}\n".format(seed))


def fresh_state(**kw):
    return np.random.RandomState(seed=np.random.randint(low=0, high=2 ** 32, **kw))


day = pd.to_datetime('2019-06-28')
sym, cash = 'JPM', 10000000
mkt_open, mkt_close = day + pd.to_timedelta('09:30:00'), day + pd.to_timedelta(args.end_time)
agents = [ExchangeAgent(id=0, name="EXCHANGE_AGENT", type="ExchangeAgent", mkt_open=mkt_open, mkt_close=mkt_close,
                        symbols=[sym], log_orders=False, pipeline_delay=0, computation_delay=0, stream_history=10,
                        book_freq=0, random_state=fresh_state(dtype='uint64'))]
agents.append(MarketMakerAgent(id=1, name="MARKET_MAKER_AGENT_1", type='MarketMakerAgent', symbol=sym,
                               starting_cash=cash, min_size=500, max_size=1000, subscribe=True,
                               subscribe_freq=args.mm_subscribe_freq, subscribe_num_levels=args.mm_subscribe_levels,
                               log_orders=False,
                               random_state=fresh_state(dtype='uint64')))
fund = {'r_bar': 1e5, 'kappa': 1.67e-12, 'agent_kappa': 1.67e-15, 'sigma_s': 0, 'fund_vol': 1e-8,
        'megashock_lambda_a': 2.77778e-13, 'megashock_mean': 1e3, 'megashock_var': 5e4,
        'random_state': fresh_state(dtype='uint64')}
oracle = SparseMeanRevertingOracle(mkt_open, mkt_close, {sym: fund})
belief = dict(symbol=sym, starting_cash=cash, sigma_n=10000, sigma_s=fund['fund_vol'], kappa=fund['agent_kappa'],
              r_bar=fund['r_bar'], q_max=10, sigma_pv=5e4, R_min=0, R_max=100, eta=1, lambda_a=1e-12, log_orders=False)
for _ in range(args.num_zi):
    j = len(agents)
    agents.append(ZeroIntelligenceAgent(id=j, name="ZI_AGENT_{}".format(j), type="ZeroIntelligenceAgent",
                                        random_state=fresh_state(dtype='uint64'), **belief))
for _ in range(args.num_hbl):
    j = len(agents)
    agents.append(HeuristicBeliefLearningAgent(id=j, name="HBL_AGENT_{}".format(j), type="HeuristicBeliefLearningAgent",
                                               L=2, random_state=fresh_state(dtype='uint64'), **belief))
for _ in range(args.num_momentum):
    j = len(agents)
    agents.append(MomentumAgent(id=j, name="MOMENTUM_AGENT_{}".format(j), type="MomentumAgent", symbol=sym,
                                starting_cash=cash, min_size=1, max_size=10, subscribe=True, log_orders=False,
                                random_state=fresh_state(dtype='uint64')))
if args.subscription_agent:
    agents.append(SubscriptionAgent(id=len(agents), name="EXAMPLE_SUBSCRIPTION_AGENT", type="SubscriptionAgent",
                                    symbol=sym, starting_cash=cash, levels=5, freq=10e9, log_orders=False,
                                    random_state=fresh_state(dtype='uint64')))

kernel = Kernel("RMSC02 Kernel", random_state=fresh_state(dtype='uint64'))
latency_rstate = fresh_state()
dist = util.generate_uniform_random_pairwise_dist_on_line(0.0, 3866660, len(agents), random_state=latency_rstate)
model = LatencyModel(latency_model='deterministic', random_state=latency_rstate,
                     kwargs={'connected': True, 'min_latency': util.meters_to_light_ns(dist)})
kernel.runner(agents=agents, startTime=day, stopTime=mkt_close + pd.to_timedelta('00:01:00'),
              agentLatencyModel=model, defaultComputationDelay=50, oracle=oracle, log_dir=args.log_dir)
print("Simulation End Time: {}".format(dt.datetime.now()))
print("Time taken to run simulation: {}".format(dt.datetime.now() - t_begin))
